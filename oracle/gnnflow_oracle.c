/*
 * gnnflow_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C (OpenMP) CPU restatement of the reference GNNFlow hot path:
 *   - the block-adjacency-list store and its batched add_edges / offload path,
 *   - SampleLayerRecent / SampleLayerUniform and the SamplingResult assembly.
 * It deliberately keeps the reference's data structure (a doubly linked list of TemporalBlocks per
 * vertex, three SoA arrays per block) and walks it the way the reference kernels do, one
 * (target, fanout-slot) "thread" at a time, so that it can serve as the parity oracle for the
 * B200 implementation, whose data layout is different.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product path (gnnflow_b200/) never does.
 *
 * Parity pinning: tests/test_oracle_golden.py replays every golden vector of the reference's own
 * tests/test_dynamic_graph.py and tests/test_temporal_sampler.py through this file.
 *
 * Each function cites the reference file:line it follows (paths relative to the reference root).
 *
 * Deliberate, documented deviations (all on inputs where the reference is undefined):
 *   D1  uniform sampling with ZERO candidates: the reference computes `curand() % 0` (UB,
 *       gnnflow/csrc/sampling_kernels.cu:202).  Intended semantics implemented here: no output.
 *   D2  RNG: the reference uses cuRAND XORWOW device state (gnnflow/csrc/utils.cu:88-94).  The shared
 *       counter-based stream used by both this oracle and the CUDA path is Philox4x32-10 with
 *       counter = (tid, launch_index, 0, 0), key = (seed_lo, seed_hi), word 0 of the output;
 *       tid = target_index * fanout + slot, launch_index = number of non-empty SampleLayer launches
 *       this sampler performed before this one.
 *   D3  a target id outside [0, max_node_id] reads out of bounds in the reference
 *       (sampling_kernels.cu:43); here it has no neighbours.
 *   D4  out-of-order batches: the reference only CHECKs last_new_ts >= tail.end_ts
 *       (gnnflow/csrc/utils.cu:43, LOG(FATAL) -> abort) and silently stores unsorted blocks
 *       otherwise.  Here a batch whose oldest new edge of any vertex is older than that vertex's
 *       tail end_timestamp is rejected as a whole (return -2) before any mutation.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OG_API __attribute__((visibility("default")))

/* ----------------------------------------------------------------------------------------------
 * types: gnnflow/csrc/common.h:13-48, gnnflow/csrc/doubly_linked_list.h:15-34
 * -------------------------------------------------------------------------------------------- */
typedef int64_t nid_t;
typedef int64_t eid_t;
typedef float ts_t;

typedef struct og_block {
  nid_t *dst_nodes;
  ts_t *timestamps;
  eid_t *eids;
  size_t size;
  size_t capacity;
  ts_t start_timestamp;
  ts_t end_timestamp;
  struct og_block *prev;
  struct og_block *next;
} og_block;

typedef struct {
  og_block *head;
  og_block *tail;
  size_t num_edges;
  size_t num_insertions;
  size_t size;
} og_list;

/* eid -> refcount, the reference's std::unordered_map<EIDType,size_t> edges_ (dynamic_graph.h) */
typedef struct {
  eid_t *keys;
  size_t *vals;
  uint8_t *used;
  size_t cap;   /* power of two */
  size_t slots; /* used slots (incl. zero-count tombstones) */
  size_t live;  /* keys with val > 0 */
} og_eidmap;

typedef struct og_graph {
  og_list *table; /* h_copy_of_d_node_table_ */
  size_t table_len;
  int table_init;
  size_t min_block_size;
  int insertion_policy; /* 0 insert, 1 replace (common.h:73) */
  int adaptive_block_size;
  size_t max_node_id;
  uint8_t *is_node; /* nodes_ (std::set) as a dense flag array, same length as table */
  uint8_t *is_src;  /* src_nodes_ */
  size_t num_nodes, num_src_nodes;
  og_eidmap edges;
  size_t allocated;  /* TemporalBlockAllocator::allocated_ */
  size_t num_blocks; /* h2d_mapping_.size() */
} og_graph;

typedef struct og_sampler {
  const og_graph *graph;
  uint32_t *fanouts;
  uint32_t num_layers;
  int policy; /* 0 recent, 1 uniform (common.h:82) */
  uint32_t num_snapshots;
  float snapshot_time_window;
  int prop_time;
  uint64_t seed;
  uint64_t launch_index;
} og_sampler;

/* ----------------------------------------------------------------------------------------------
 * eid map
 * -------------------------------------------------------------------------------------------- */
static uint64_t og_mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
static void eidmap_init(og_eidmap *m, size_t cap) {
  m->cap = cap; m->slots = 0; m->live = 0;
  m->keys = (eid_t *)malloc(cap * sizeof(eid_t));
  m->vals = (size_t *)calloc(cap, sizeof(size_t));
  m->used = (uint8_t *)calloc(cap, 1);
}
static void eidmap_free(og_eidmap *m) { free(m->keys); free(m->vals); free(m->used); }
static size_t *eidmap_slot(og_eidmap *m, eid_t k, int create);
static void eidmap_grow(og_eidmap *m) {
  og_eidmap n; eidmap_init(&n, m->cap * 2);
  for (size_t i = 0; i < m->cap; i++)
    if (m->used[i] && m->vals[i] > 0) { *eidmap_slot(&n, m->keys[i], 1) = m->vals[i]; n.live++; }
  eidmap_free(m); *m = n;
}
static size_t *eidmap_slot(og_eidmap *m, eid_t k, int create) {
  if (create && (m->slots + 1) * 2 > m->cap) eidmap_grow(m);
  size_t i = og_mix64((uint64_t)k) & (m->cap - 1);
  while (m->used[i]) { if (m->keys[i] == k) return &m->vals[i]; i = (i + 1) & (m->cap - 1); }
  if (!create) return NULL;
  m->used[i] = 1; m->keys[i] = k; m->vals[i] = 0; m->slots++;
  return &m->vals[i];
}
static void eidmap_inc(og_eidmap *m, eid_t k) { size_t *v = eidmap_slot(m, k, 1); if ((*v)++ == 0) m->live++; }
static void eidmap_dec(og_eidmap *m, eid_t k) {
  size_t *v = eidmap_slot(m, k, 0);
  if (v && *v > 0 && --(*v) == 0) m->live--;
}

/* ----------------------------------------------------------------------------------------------
 * TemporalBlockAllocator: gnnflow/csrc/temporal_block_allocator.cu:83-180
 * -------------------------------------------------------------------------------------------- */
/* AlignUp, temporal_block_allocator.cu:83-88 */
static size_t og_align_up(const og_graph *g, size_t size) { return size < g->min_block_size ? g->min_block_size : size; }

/* AllocateInternal, temporal_block_allocator.cu:134-157 */
static void og_allocate_internal(og_graph *g, og_block *b, size_t size) {
  size_t capacity = og_align_up(g, size);
  b->size = 0;
  b->capacity = capacity;
  b->start_timestamp = FLT_MAX;
  b->end_timestamp = 0;
  b->prev = NULL;
  b->next = NULL;
  b->dst_nodes = (nid_t *)malloc(capacity * sizeof(nid_t));
  b->timestamps = (ts_t *)malloc(capacity * sizeof(ts_t));
  b->eids = (eid_t *)malloc(capacity * sizeof(eid_t));
  g->allocated += capacity * (sizeof(nid_t) + sizeof(ts_t) + sizeof(eid_t));
}
/* DeallocateInternal, temporal_block_allocator.cu:159-180 */
static void og_deallocate_internal(og_graph *g, og_block *b) {
  if (b->dst_nodes) { free(b->dst_nodes); b->dst_nodes = NULL; g->allocated -= b->capacity * sizeof(nid_t); }
  if (b->timestamps) { free(b->timestamps); b->timestamps = NULL; g->allocated -= b->capacity * sizeof(ts_t); }
  if (b->eids) { free(b->eids); b->eids = NULL; g->allocated -= b->capacity * sizeof(eid_t); }
  b->size = 0;
  b->capacity = 0;
}
/* Allocate, temporal_block_allocator.cu:90-108 */
static og_block *og_allocate(og_graph *g, size_t size) {
  og_block *b = (og_block *)calloc(1, sizeof(og_block));
  og_allocate_internal(g, b, size);
  return b;
}
/* Reallocate + CopyTemporalBlock, temporal_block_allocator.cu:122-132, utils.cu:9-31 */
static void og_reallocate(og_graph *g, og_block *b, size_t size) {
  og_block tmp;
  og_allocate_internal(g, &tmp, size);
  memcpy(tmp.dst_nodes, b->dst_nodes, b->size * sizeof(nid_t));
  memcpy(tmp.timestamps, b->timestamps, b->size * sizeof(ts_t));
  memcpy(tmp.eids, b->eids, b->size * sizeof(eid_t));
  tmp.size = b->size;
  tmp.start_timestamp = b->start_timestamp;
  tmp.end_timestamp = b->end_timestamp;
  tmp.next = b->next;
  og_block *prev = b->prev; /* CopyTemporalBlock does not copy prev (utils.cu:27-30): `*block = tmp`
                               would null it; a single-block "replace" list never has one. */
  og_deallocate_internal(g, b);
  *b = tmp;
  b->prev = prev;
}

/* ----------------------------------------------------------------------------------------------
 * linked list: gnnflow/csrc/doubly_linked_list.cu:37-81
 * -------------------------------------------------------------------------------------------- */
static void og_list_insert(og_list *l, og_block *b) {
  if (l->tail == NULL) { l->tail = b; l->head = b; b->prev = NULL; b->next = NULL; }
  else { l->tail->next = b; b->prev = l->tail; b->next = NULL; l->tail = b; }
  l->size++;
}
static void og_list_remove(og_list *l, og_block *b) {
  if (b->prev == NULL && b->next == NULL) { l->head = l->tail = NULL; }
  else if (b->prev == NULL) { l->head = b->next; b->next->prev = NULL; }
  else if (b->next == NULL) { l->tail = b->prev; b->prev->next = NULL; }
  else { b->prev->next = b->next; b->next->prev = b->prev; }
  l->size--;
}

/* ----------------------------------------------------------------------------------------------
 * DynamicGraph: gnnflow/csrc/dynamic_graph.cu
 * -------------------------------------------------------------------------------------------- */
OG_API og_graph *og_graph_create(uint64_t min_block_size, int insertion_policy, int adaptive_block_size) {
  og_graph *g = (og_graph *)calloc(1, sizeof(og_graph));
  g->min_block_size = (size_t)min_block_size;
  g->insertion_policy = insertion_policy;
  g->adaptive_block_size = adaptive_block_size;
  g->max_node_id = 0; /* dynamic_graph.cu:33 */
  eidmap_init(&g->edges, 1024);
  return g;
}

OG_API void og_graph_destroy(og_graph *g) {
  if (!g) return;
  for (size_t v = 0; v < g->table_len; v++) {
    og_block *b = g->table[v].head;
    while (b) { og_block *n = b->next; og_deallocate_internal(g, b); free(b); b = n; }
  }
  free(g->table); free(g->is_node); free(g->is_src);
  eidmap_free(&g->edges);
  free(g);
}

/* AddNodes, dynamic_graph.cu:140-147 */
static void og_add_nodes(og_graph *g, size_t max_node) {
  if (g->table_init && max_node < g->max_node_id) return;
  if (!g->table_init || max_node + 1 > g->table_len) {
    size_t n = max_node + 1;
    g->table = (og_list *)realloc(g->table, n * sizeof(og_list));
    g->is_node = (uint8_t *)realloc(g->is_node, n);
    g->is_src = (uint8_t *)realloc(g->is_src, n);
    memset(g->table + g->table_len, 0, (n - g->table_len) * sizeof(og_list));
    memset(g->is_node + g->table_len, 0, n - g->table_len);
    memset(g->is_src + g->table_len, 0, n - g->table_len);
    g->table_len = n;
  }
  g->max_node_id = max_node;
  g->table_init = 1;
}

/* get_next_power_of_two, dynamic_graph.cu:202-204 (n == 1 is UB there: clzl(0); lzcnt gives 1) */
static size_t og_next_pow2(size_t n) {
  if (n <= 1) return 1;
  return (size_t)1 << (64 - __builtin_clzl(n - 1));
}

/* CopyEdgesToBlock, utils.cu:33-63 */
static void og_copy_edges_to_block(og_block *b, const nid_t *dst, const ts_t *ts, const eid_t *eid,
                                   size_t start_idx, size_t n) {
  memcpy(b->dst_nodes + b->size, dst + start_idx, n * sizeof(nid_t));
  memcpy(b->timestamps + b->size, ts + start_idx, n * sizeof(ts_t));
  memcpy(b->eids + b->size, eid + start_idx, n * sizeof(eid_t));
  b->size += n;
  b->start_timestamp = b->start_timestamp < ts[start_idx] ? b->start_timestamp : ts[start_idx];
  b->end_timestamp = ts[start_idx + n - 1];
}

/* AddEdgesForOneNode, dynamic_graph.cu:206-287.  dst/ts/eid: this vertex's edges, sorted by ts. */
static void og_add_edges_for_one_node(og_graph *g, nid_t src, const nid_t *dst, const ts_t *ts,
                                      const eid_t *eid, size_t total) {
  size_t num_edges = total;
  og_list *l = &g->table[src];
  og_block *tail = l->tail;
  og_block *blk = NULL;
  int is_new = 0;
  size_t start_idx = 0;
  if (tail == NULL) {
    /* case 1: empty list */
    blk = og_allocate(g, num_edges);
    is_new = 1;
  } else if (tail->size + num_edges > tail->capacity) {
    /* case 2: not enough space in the current block */
    if (g->insertion_policy == 0) {
      size_t fill = tail->capacity - tail->size;
      if (fill > 0) {
        og_copy_edges_to_block(tail, dst, ts, eid, 0, fill);
        start_idx = fill;
        num_edges -= fill;
      }
      size_t avg = l->num_insertions == 0 ? num_edges : l->num_edges / l->num_insertions;
      size_t new_size;
      if (g->adaptive_block_size) {
        new_size = num_edges > avg ? num_edges : avg;
        new_size = og_next_pow2(new_size);
      } else {
        new_size = num_edges;
      }
      blk = og_allocate(g, new_size);
      is_new = 1;
    } else {
      og_reallocate(g, tail, tail->size + num_edges);
    }
  }
  if (!is_new) blk = tail; /* case 3 */
  og_copy_edges_to_block(blk, dst, ts, eid, start_idx, num_edges);
  if (is_new) { og_list_insert(l, blk); g->num_blocks++; }
  l->num_edges += total;
  l->num_insertions++;
}

/* stable merge sort of an index permutation by (src, ts): equals "group by src, then stable_sort by
 * timestamp within each group" (dynamic_graph.cu:105-128, utils.h:16-38). */
static void og_msort(uint32_t *idx, uint32_t *tmp, size_t n, const nid_t *src, const ts_t *ts) {
  if (n < 2) return;
  size_t h = n / 2;
  og_msort(idx, tmp, h, src, ts);
  og_msort(idx + h, tmp, n - h, src, ts);
  size_t i = 0, j = h, k = 0;
  while (i < h && j < n) {
    uint32_t a = idx[i], b = idx[j];
    int b_less = (src[b] < src[a]) || (src[b] == src[a] && ts[b] < ts[a]);
    tmp[k++] = b_less ? idx[j++] : idx[i++];
  }
  while (i < h) tmp[k++] = idx[i++];
  while (j < n) tmp[k++] = idx[j++];
  memcpy(idx, tmp, n * sizeof(uint32_t));
}

/* AddEdges, dynamic_graph.cu:77-138.  returns 0 ok, -1 bad argument, -2 out-of-order batch (D4) */
OG_API int og_graph_add_edges(og_graph *g, const nid_t *src, const nid_t *dst, const ts_t *ts,
                              const eid_t *eid, size_t n) {
  if (n == 0 || n > 0xffffffffu) return -1; /* CHECK_GT(src_nodes.size(), 0) */
  nid_t max_node = 0;
  for (size_t i = 0; i < n; i++) {
    if (src[i] < 0 || dst[i] < 0) return -1;
    if (src[i] > max_node) max_node = src[i];
    if (dst[i] > max_node) max_node = dst[i];
  }
  uint32_t *idx = (uint32_t *)malloc(n * sizeof(uint32_t));
  uint32_t *tmp = (uint32_t *)malloc(n * sizeof(uint32_t));
  for (size_t i = 0; i < n; i++) idx[i] = (uint32_t)i;
  og_msort(idx, tmp, n, src, ts);
  free(tmp);
  /* D4 validation before any mutation */
  for (size_t i = 0; i < n; i++) {
    if (i > 0 && src[idx[i]] == src[idx[i - 1]]) continue;
    nid_t v = src[idx[i]];
    if ((size_t)v < g->table_len && g->table[v].tail && ts[idx[i]] < g->table[v].tail->end_timestamp) {
      free(idx);
      return -2;
    }
  }
  og_add_nodes(g, (size_t)max_node);
  for (size_t i = 0; i < n; i++) {
    if (!g->is_src[src[i]]) { g->is_src[src[i]] = 1; g->num_src_nodes++; }
    if (!g->is_node[src[i]]) { g->is_node[src[i]] = 1; g->num_nodes++; }
    if (!g->is_node[dst[i]]) { g->is_node[dst[i]] = 1; g->num_nodes++; }
    eidmap_inc(&g->edges, eid[i]);
  }
  nid_t *sdst = (nid_t *)malloc(n * sizeof(nid_t));
  ts_t *sts = (ts_t *)malloc(n * sizeof(ts_t));
  eid_t *seid = (eid_t *)malloc(n * sizeof(eid_t));
  for (size_t i = 0; i < n; i++) { sdst[i] = dst[idx[i]]; sts[i] = ts[idx[i]]; seid[i] = eid[idx[i]]; }
  size_t s = 0;
  while (s < n) {
    size_t e = s + 1;
    nid_t v = src[idx[s]];
    while (e < n && src[idx[e]] == v) e++;
    og_add_edges_for_one_node(g, v, sdst + s, sts + s, seid + s, e - s);
    s = e;
  }
  free(sdst); free(sts); free(seid); free(idx);
  return 0;
}

/* OffloadOldBlocks, dynamic_graph.cu:382-411 (to_file only changes where the payload goes) */
OG_API size_t og_graph_offload_old_blocks(og_graph *g, float timestamp) {
  size_t num_blocks = 0;
  for (size_t v = 0; v < g->table_len; v++) {
    if (!g->is_node[v]) continue;
    og_list *l = &g->table[v];
    og_block *cur = l->head;
    while (cur) {
      og_block *next = cur->next;
      if (cur->end_timestamp < timestamp) {
        for (size_t i = 0; i < cur->size; i++) eidmap_dec(&g->edges, cur->eids[i]);
        og_list_remove(l, cur);
        g->num_blocks--;
        og_deallocate_internal(g, cur);
        free(cur);
        num_blocks++;
      }
      cur = next;
    }
  }
  return num_blocks;
}

OG_API size_t og_graph_num_nodes(const og_graph *g) { return g->num_nodes; }          /* dynamic_graph.cu:149 */
OG_API size_t og_graph_num_src_nodes(const og_graph *g) { return g->num_src_nodes; }  /* :150 */
OG_API size_t og_graph_num_edges(const og_graph *g) { return g->edges.live; }         /* :151 */
OG_API int64_t og_graph_max_node_id(const og_graph *g) { return (int64_t)g->max_node_id; } /* :357 */

/* out_degree, dynamic_graph.cu:289-297 */
OG_API void og_graph_out_degree(const og_graph *g, const nid_t *nodes, size_t n, uint64_t *out) {
  for (size_t i = 0; i < n; i++)
    out[i] = (nodes[i] >= 0 && (size_t)nodes[i] < g->table_len) ? g->table[nodes[i]].num_edges : 0;
}

/* get_temporal_neighbors, dynamic_graph.cu:299-337: newest block first, each block reversed.
 * Two-call: returns the total; fills at most cap entries. */
OG_API size_t og_graph_get_temporal_neighbors(const og_graph *g, nid_t node, nid_t *dst, ts_t *ts,
                                              eid_t *eid, size_t cap) {
  size_t k = 0;
  if (node < 0 || (size_t)node >= g->table_len) return 0;
  for (og_block *b = g->table[node].tail; b; b = b->prev)
    for (size_t i = b->size; i-- > 0;) {
      if (k < cap) { dst[k] = b->dst_nodes[i]; ts[k] = b->timestamps[i]; eid[k] = b->eids[i]; }
      k++;
    }
  return k;
}

/* nodes / src_nodes / edges, dynamic_graph.cu:343-355 (std::set order = ascending; the reference's
 * edges() order is unordered_map order, here ascending) */
OG_API size_t og_graph_nodes(const og_graph *g, nid_t *out, size_t cap) {
  size_t k = 0;
  for (size_t v = 0; v < g->table_len; v++) if (g->is_node[v]) { if (k < cap) out[k] = (nid_t)v; k++; }
  return k;
}
OG_API size_t og_graph_src_nodes(const og_graph *g, nid_t *out, size_t cap) {
  size_t k = 0;
  for (size_t v = 0; v < g->table_len; v++) if (g->is_src[v]) { if (k < cap) out[k] = (nid_t)v; k++; }
  return k;
}
static int og_cmp_i64(const void *a, const void *b) {
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return (x > y) - (x < y);
}
OG_API size_t og_graph_edges(const og_graph *g, eid_t *out, size_t cap) {
  size_t k = 0;
  for (size_t i = 0; i < g->edges.cap; i++)
    if (g->edges.used[i] && g->edges.vals[i] > 0) { if (k < cap) out[k] = g->edges.keys[i]; k++; }
  if (k <= cap) qsort(out, k, sizeof(eid_t), og_cmp_i64);
  return k;
}

/* avg_linked_list_length, dynamic_graph.cu:359-366 */
OG_API float og_graph_avg_linked_list_length(const og_graph *g) {
  float sum = 0;
  for (size_t v = 0; v < g->table_len; v++) if (g->is_node[v]) sum += g->table[v].size;
  return sum / g->num_nodes;
}
/* graph_mem_usage, dynamic_graph.cu:368-370 */
OG_API float og_graph_mem_usage(const og_graph *g) { return (float)g->allocated; }
/* graph_metadata_mem_usage, dynamic_graph.cu:372-380: 64 B per block (sizeof(TemporalBlock)) + 8 B per table entry */
OG_API float og_graph_metadata_mem_usage(const og_graph *g) {
  float sum = 0;
  sum += 64 * g->num_blocks; /* sizeof(TemporalBlock) == 64 */
  sum += 8 * g->table_len;
  return sum;
}
/* per-vertex block shape, oldest -> newest; for checking the block-sizing policy (a5) */
OG_API size_t og_graph_block_shapes(const og_graph *g, nid_t node, uint64_t *sizes, uint64_t *caps,
                                    float *start_ts, float *end_ts, size_t cap) {
  size_t k = 0;
  if (node < 0 || (size_t)node >= g->table_len) return 0;
  for (og_block *b = g->table[node].head; b; b = b->next) {
    if (k < cap) { sizes[k] = b->size; caps[k] = b->capacity; start_ts[k] = b->start_timestamp; end_ts[k] = b->end_timestamp; }
    k++;
  }
  return k;
}

/* ----------------------------------------------------------------------------------------------
 * sampling kernels: gnnflow/csrc/sampling_kernels.cu, gnnflow/csrc/utils.cu:96-109
 * -------------------------------------------------------------------------------------------- */
/* LowerBound, utils.cu:96-109 */
static void og_lower_bound(const ts_t *timestamps, int num_edges, ts_t timestamp, int *res) {
  int left = 0, right = num_edges;
  while (left < right) {
    int mid = (left + right) / 2;
    if (timestamps[mid] < timestamp) left = mid + 1; else right = mid;
  }
  *res = left;
}

/* window arithmetic, sampling_kernels.cu:27-40.  The reference is built with --use_fast_math, so the
 * multi-snapshot `root - float(u) * w` is a single fused multiply-add. */
static void og_window(ts_t root_timestamp, uint32_t snapshot_idx, uint32_t num_snapshots, ts_t w,
                      ts_t *start_timestamp, ts_t *end_timestamp) {
  if (num_snapshots == 1) {
    if ((double)fabsf(w) < 1e-6) *start_timestamp = 0; else *start_timestamp = root_timestamp - w;
    *end_timestamp = root_timestamp;
  } else {
    *end_timestamp = fmaf(-(float)(num_snapshots - snapshot_idx - 1), w, root_timestamp);
    *start_timestamp = *end_timestamp - w;
  }
}

/* the per-block [start_idx, end_idx) computation shared by both kernels, sampling_kernels.cu:66-86 */
static void og_block_range(const og_block *b, ts_t start_timestamp, ts_t end_timestamp, int *start_idx, int *end_idx) {
  if (start_timestamp >= b->start_timestamp && end_timestamp <= b->end_timestamp) {
    og_lower_bound(b->timestamps, (int)b->size, start_timestamp, start_idx);
    og_lower_bound(b->timestamps, (int)b->size, end_timestamp, end_idx);
  } else if (start_timestamp < b->start_timestamp && end_timestamp <= b->end_timestamp) {
    *start_idx = 0;
    og_lower_bound(b->timestamps, (int)b->size, end_timestamp, end_idx);
  } else if (start_timestamp > b->start_timestamp && end_timestamp > b->end_timestamp) {
    og_lower_bound(b->timestamps, (int)b->size, start_timestamp, start_idx);
    *end_idx = (int)b->size;
  } else {
    *start_idx = 0;
    *end_idx = (int)b->size;
  }
}

typedef struct { nid_t nbr; eid_t eid; ts_t ts; ts_t dt; } og_slot;

static const og_block *og_tail_of(const og_graph *g, nid_t nid) {
  if (nid < 0 || (size_t)nid >= g->table_len) return NULL; /* D3 */
  return g->table[nid].tail;
}

/* one "thread" of SampleLayerRecentKernel, sampling_kernels.cu:11-107.  returns 1 if a neighbour was
 * emitted, 0 if the slot is invalid (kInvalidNID). */
static int og_recent_thread(const og_sampler *s, nid_t nid, ts_t root_timestamp, uint32_t snapshot_idx,
                            uint32_t sample_index, og_slot *out) {
  ts_t start_timestamp, end_timestamp;
  og_window(root_timestamp, snapshot_idx, s->num_snapshots, s->snapshot_time_window, &start_timestamp, &end_timestamp);
  const og_block *curr = og_tail_of(s->graph, nid);
  int start_idx, end_idx;
  int index = (int)sample_index;
  while (curr != NULL) {
    if (curr->capacity == 0) { curr = curr->prev; continue; }
    if (end_timestamp < curr->start_timestamp) { curr = curr->prev; continue; }
    if (start_timestamp > curr->end_timestamp) break;
    og_block_range(curr, start_timestamp, end_timestamp, &start_idx, &end_idx);
    int32_t i = end_idx - 1 - index;
    if (i < start_idx) {
      index -= end_idx - start_idx;
      curr = curr->prev;
      continue;
    } else {
      out->nbr = curr->dst_nodes[i];
      out->eid = curr->eids[i];
      ts_t timestamp = curr->timestamps[i];
      out->ts = s->prop_time ? root_timestamp : timestamp;
      out->dt = root_timestamp - timestamp;
      return 1;
    }
  }
  return 0;
}

/* Philox4x32-10 (Salmon et al., SC'11), D2 */
static inline void og_philox_round(uint32_t c[4], const uint32_t k[2]) {
  uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
  uint32_t n1 = (uint32_t)p1;
  uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
  uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
OG_API uint32_t og_philox_u32(uint64_t seed, uint32_t tid, uint64_t launch_index) {
  uint32_t c[4] = {tid, (uint32_t)launch_index, (uint32_t)(launch_index >> 32), 0};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int r = 0; r < 10; r++) {
    og_philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  return c[0];
}

/* one "thread" of SampleLayerUniformKernel, sampling_kernels.cu:109-273 (the shared-memory range
 * cache :192-195,233-236 is a pure optimisation and is not restated) */
static int og_uniform_thread(const og_sampler *s, nid_t nid, ts_t root_timestamp, uint32_t snapshot_idx,
                             uint32_t tid, uint64_t launch_index, og_slot *out) {
  ts_t start_timestamp, end_timestamp;
  og_window(root_timestamp, snapshot_idx, s->num_snapshots, s->snapshot_time_window, &start_timestamp, &end_timestamp);
  const og_block *tail = og_tail_of(s->graph, nid);
  uint32_t num_candidates = 0;
  const og_block *curr = tail;
  int start_idx, end_idx;
  while (curr != NULL) {
    if (curr->capacity == 0) { curr = curr->prev; continue; }
    if (end_timestamp < curr->start_timestamp) { curr = curr->prev; continue; }
    if (start_timestamp > curr->end_timestamp) break;
    og_block_range(curr, start_timestamp, end_timestamp, &start_idx, &end_idx);
    num_candidates += end_idx - start_idx;
    curr = curr->prev;
  }
  if (num_candidates == 0) return 0; /* D1 */
  uint32_t index = og_philox_u32(s->seed, tid, launch_index) % num_candidates; /* D2 */
  curr = tail;
  while (curr != NULL) {
    if (curr->capacity == 0) { curr = curr->prev; continue; }
    if (end_timestamp < curr->start_timestamp) { curr = curr->prev; continue; }
    if (start_timestamp > curr->end_timestamp) break;
    og_block_range(curr, start_timestamp, end_timestamp, &start_idx, &end_idx);
    int32_t i = (int32_t)((uint32_t)end_idx - 1u - index);
    if (i < start_idx) {
      index -= end_idx - start_idx;
      curr = curr->prev;
      continue;
    } else {
      out->nbr = curr->dst_nodes[i];
      out->eid = curr->eids[i];
      ts_t timestamp = curr->timestamps[i];
      out->ts = s->prop_time ? root_timestamp : timestamp;
      out->dt = root_timestamp - timestamp;
      return 1;
    }
  }
  return 0;
}

/* ----------------------------------------------------------------------------------------------
 * TemporalSampler: gnnflow/csrc/temporal_sampler.cu
 * -------------------------------------------------------------------------------------------- */
OG_API og_sampler *og_sampler_create(const og_graph *g, const uint32_t *fanouts, uint32_t num_layers, int policy,
                                     uint32_t num_snapshots, float snapshot_time_window, int prop_time, uint64_t seed) {
  og_sampler *s = (og_sampler *)calloc(1, sizeof(og_sampler));
  s->graph = g;
  s->fanouts = (uint32_t *)malloc(num_layers * sizeof(uint32_t));
  memcpy(s->fanouts, fanouts, num_layers * sizeof(uint32_t));
  s->num_layers = num_layers;
  s->policy = policy;
  s->num_snapshots = num_snapshots;
  s->snapshot_time_window = snapshot_time_window;
  s->prop_time = prop_time;
  s->seed = seed;
  s->launch_index = 0;
  return s;
}
OG_API void og_sampler_destroy(og_sampler *s) { if (s) { free(s->fanouts); free(s); } }
OG_API uint64_t og_sampler_launch_index(const og_sampler *s) { return s->launch_index; }
OG_API void og_sampler_set_launch_index(og_sampler *s, uint64_t v) { s->launch_index = v; }

/*
 * SampleLayer, temporal_sampler.cu:97-277.
 * Kernel launch over T*F threads, thrust::remove_if compaction (stable: target-major, slot order),
 * num_sampled[] -> row.  Outputs (caller allocated, capacity T*F):
 *   out_nbr/out_eid/out_ts/out_dt : the compacted neighbour arrays; all_nodes = roots ++ out_nbr,
 *   all_timestamps = root_ts ++ out_ts, delta_timestamps = out_dt, eids = out_eid,
 *   out_row[j] = target index of edge j, col[j] = T + j.
 *   out_num_sampled[T] (may be NULL).
 * Returns the number of sampled neighbours S, or (size_t)-1 on a bad layer / snapshot.
 */
OG_API size_t og_sampler_sample_layer(og_sampler *s, const nid_t *nodes, const ts_t *timestamps, size_t T,
                                      uint32_t layer, uint32_t snapshot, nid_t *out_nbr, eid_t *out_eid,
                                      ts_t *out_ts, ts_t *out_dt, int64_t *out_row, uint32_t *out_num_sampled) {
  if (layer >= s->num_layers || snapshot >= s->num_snapshots) return (size_t)-1;
  if (T == 0) return 0; /* temporal_sampler.cu:107-114 */
  const uint32_t fanout = s->fanouts[layer];
  const uint64_t launch_index = s->launch_index++;
  uint32_t *cnt = (uint32_t *)malloc(T * sizeof(uint32_t));
  og_slot *slots = (og_slot *)malloc(T * (size_t)fanout * sizeof(og_slot));
  uint8_t *valid = (uint8_t *)malloc(T * (size_t)fanout);
#pragma omp parallel for schedule(dynamic, 256)
  for (size_t i = 0; i < T; i++) {
    uint32_t c = 0;
    for (uint32_t k = 0; k < fanout; k++) {
      size_t tid = i * fanout + k;
      int ok = s->policy == 0
                   ? og_recent_thread(s, nodes[i], timestamps[i], snapshot, k, &slots[tid])
                   : og_uniform_thread(s, nodes[i], timestamps[i], snapshot, (uint32_t)tid, launch_index, &slots[tid]);
      valid[tid] = (uint8_t)ok;
      c += ok;
    }
    cnt[i] = c;
  }
  /* row_offsets, temporal_sampler.cu:263-267 */
  size_t *off = (size_t *)malloc((T + 1) * sizeof(size_t));
  off[0] = 0;
  for (size_t i = 0; i < T; i++) off[i + 1] = off[i] + cnt[i];
  size_t S = off[T];
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < T; i++) {
    size_t o = off[i];
    for (uint32_t k = 0; k < fanout; k++) {
      size_t tid = i * fanout + k;
      if (!valid[tid]) continue;
      out_nbr[o] = slots[tid].nbr; out_eid[o] = slots[tid].eid;
      out_ts[o] = slots[tid].ts; out_dt[o] = slots[tid].dt;
      out_row[o] = (int64_t)i;
      o++;
    }
    if (out_num_sampled) out_num_sampled[i] = cnt[i];
  }
  free(off); free(valid); free(slots); free(cnt);
  return S;
}

OG_API int og_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
