// Dynamic-graph store: per-vertex time-sorted TemporalBlocks and the batched edge-insert path, entirely on
// the device.  Replaces reference gnnflow/csrc/dynamic_graph.cu, temporal_block_allocator.cu,
// doubly_linked_list.cu and the host loops of DynamicGraph::AddEdges (dynamic_graph.cu:77-138,206-287).
//
// add_edges = 2 host synchronisations and ~15 kernel launches per batch regardless of how many vertices the
// batch touches (the reference issues ~5 CUDA API calls per distinct source vertex):
//   batch_stats -> [sync: grow vertex table / edge-id table] -> radix sort by (src, ts) -> segment heads ->
//   plan (block-sizing policy per vertex, validation) -> scan of allocation sizes -> [sync: grow arena] ->
//   commit (descriptors, directories) -> scatter (payload) .
#include <cfloat>
#include <cstdio>
#include <cstdlib>

#include <algorithm>

#include "gf_primitives.cuh"
#include "gf_store.cuh"

namespace gf {

// ------------------------------------------------------------------------------------------ error text
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char *get_error() { return g_err; }
std::atomic<unsigned long long> g_launch_count{0};

// ------------------------------------------------------------------------------------------ kernels
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) batch_stats_kernel(const int64_t *__restrict__ src,
                                                               const int64_t *__restrict__ dst,
                                                               const float *__restrict__ ts,
                                                               const int64_t *__restrict__ eid, uint64_t n,
                                                               GraphStats *stats) {
  long long mn = INT64_MAX, mx = INT64_MIN, emn = INT64_MAX, emx = INT64_MIN;
  unsigned unsorted = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    long long s = src[i], d = dst[i], e = eid[i];
    mn = min(mn, min(s, d));
    mx = max(mx, max(s, d));
    emn = min(emn, e);
    emx = max(emx, e);
    if (i + 1 < n && ts[i + 1] < ts[i]) unsorted = 1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    emn = min(emn, __shfl_xor_sync(0xffffffffu, emn, o));
    emx = max(emx, __shfl_xor_sync(0xffffffffu, emx, o));
    unsorted |= __shfl_xor_sync(0xffffffffu, unsorted, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&stats->batch_min_id, mn);
    atomicMax(&stats->batch_max_id, mx);
    atomicMin(&stats->batch_min_eid, emn);
    atomicMax(&stats->batch_max_eid, emx);
    if (unsorted) atomicOr(&stats->ts_unsorted, 1u);
  }
}

__global__ void stats_reset_kernel(GraphStats *stats) {
  stats->batch_min_id = INT64_MAX;
  stats->batch_max_id = INT64_MIN;
  stats->batch_min_eid = INT64_MAX;
  stats->batch_max_eid = INT64_MIN;
  stats->ts_unsorted = 0;
  stats->num_segments = 0;
  stats->error_flags = 0;
  stats->total_units = 0;
  stats->call_count = 0;
}

__global__ void keys_from_ts_kernel(const float *__restrict__ ts, uint64_t n, uint32_t *keys, uint32_t *vals) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t = ts[i];
  keys[i] = t == 0.0f ? orderable_f32(0.0f) : orderable_f32(t);  // -0.0 == +0.0 under operator<
  vals[i] = (uint32_t)i;
}
// keys[i] = src[vals[i]] (vals == nullptr: identity permutation, also written to vals_out)
__global__ void keys_from_src_kernel(const int64_t *__restrict__ src, const uint32_t *__restrict__ vals_in, uint64_t n,
                                     uint32_t *keys, uint32_t *vals_out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t j = vals_in ? vals_in[i] : (uint32_t)i;
  keys[i] = (uint32_t)src[j];
  if (vals_out) vals_out[i] = j;
}

__global__ void seg_heads_kernel(const uint32_t *__restrict__ keys, uint64_t n, uint32_t *flags) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
// flags -> segment id of every element; seg_start[s] = first element of segment s; seg_start[U] = n
__global__ void seg_starts_kernel(uint32_t *flags_to_segid, const uint32_t *__restrict__ excl, uint64_t n,
                                  uint32_t *seg_start, GraphStats *stats) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t f = flags_to_segid[i];
  uint32_t sid = excl[i] + f - 1;
  if (f) seg_start[sid] = (uint32_t)i;
  flags_to_segid[i] = sid;
  if (i == n - 1) {
    seg_start[sid + 1] = (uint32_t)n;
    stats->num_segments = sid + 1;
  }
}

struct SegPlan {
  uint32_t fill;        // edges appended to the existing tail block
  uint32_t newcap;      // capacity of the block to allocate (0 = none)
  uint32_t dir_newcap;  // capacity of the new directory (0 = keep)
  uint32_t flags;       // kPlanNew | kPlanRealloc
};
enum : uint32_t { kPlanNew = 1u, kPlanRealloc = 2u };

struct SegInfo {  // where the scatter kernel writes the edges of one segment
  uint64_t p0, p1;
  uint32_t cap0, cap1;
  uint32_t off0, off1;
  uint32_t fill;
  uint32_t old_size;  // realloc only: elements to copy from old_payload
  uint64_t old_payload;
  uint32_t old_cap;
  uint32_t pad;
};

struct StoreParams {
  uint32_t min_block;
  int policy;
  int adaptive;
};

__device__ __forceinline__ uint32_t next_pow2_u32(uint32_t n) {  // dynamic_graph.cu:202-204
  return n <= 1 ? 1u : 1u << (32 - __clz(n - 1));
}

// One thread per source vertex of the batch: validation + the block-sizing policy of
// DynamicGraph::AddEdgesForOneNode (dynamic_graph.cu:206-287) + TemporalBlockAllocator::AlignUp
// (temporal_block_allocator.cu:83-88).  Nothing is mutated here.
__global__ void __launch_bounds__(kThreads) plan_kernel(const uint32_t *__restrict__ keys,
                                                        const uint32_t *__restrict__ perm,
                                                        const uint32_t *__restrict__ seg_start,
                                                        const float *__restrict__ ts, uint64_t n,
                                                        const NodeEntry *__restrict__ table, StoreParams sp,
                                                        SegPlan *plans, uint32_t *units, GraphStats *stats) {
  uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  if (s >= stats->num_segments) {
    units[s] = 0;
    return;
  }
  uint32_t b = seg_start[s], e = seg_start[s + 1];
  uint32_t cnt = e - b;
  uint32_t v = keys[b];
  float first_ts = ts[perm[b]];
  NodeEntry ent = table[v];
  bool live = ent.end > ent.first;
  SegPlan p = {0, 0, 0, 0};
  if (!live) {
    p.newcap = max(cnt, sp.min_block);
    p.flags = kPlanNew;
  } else {
    BlockDesc t = reinterpret_cast<const BlockDesc *>(ent.dir)[ent.end - 1];
    if (first_ts < t.end_ts) atomicOr(&stats->error_flags, kErrOutOfOrder);
    if ((uint64_t)t.size + cnt > t.capacity) {
      if (sp.policy == GF_INSERTION_INSERT) {
        p.fill = t.capacity - t.size;
        uint32_t rem = cnt - p.fill;
        uint64_t avg = ent.num_insertions == 0 ? rem : ent.num_edges / ent.num_insertions;
        uint32_t ns = sp.adaptive ? next_pow2_u32((uint32_t)max((uint64_t)rem, avg)) : rem;
        p.newcap = max(ns, sp.min_block);
        p.flags = kPlanNew;
      } else {
        p.newcap = max(t.size + cnt, sp.min_block);
        p.flags = kPlanRealloc;
      }
    } else {
      p.fill = cnt;
    }
  }
  uint32_t u = 0;
  if (p.flags & kPlanNew) {
    if (ent.end == ent.dir_cap) {
      uint32_t nlive = ent.end - ent.first;
      p.dir_newcap = max(4u, (2 * (nlive + 1) + 3) & ~3u);
      u += dir_units(p.dir_newcap);
    }
  }
  if (p.newcap) u += payload_units(p.newcap);
  plans[s] = p;
  units[s] = u;
}

// One thread per source vertex: applies the plan (InsertBlock / Reallocate / CopyEdgesToBlock header updates,
// dynamic_graph.cu:153-174, temporal_block_allocator.cu:122-132, utils.cu:58-62) and tells the scatter kernel
// where the payload goes.
__global__ void __launch_bounds__(kThreads) commit_kernel(const uint32_t *__restrict__ keys,
                                                          const uint32_t *__restrict__ perm,
                                                          const uint32_t *__restrict__ seg_start,
                                                          const float *__restrict__ ts, uint32_t num_segments,
                                                          NodeEntry *table, const SegPlan *__restrict__ plans,
                                                          const uint32_t *__restrict__ unit_off, uint64_t arena_base,
                                                          SegInfo *infos, uint8_t *is_src, GraphStats *stats) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= num_segments) return;
  uint32_t b = seg_start[s], e = seg_start[s + 1];
  uint32_t cnt = e - b;
  uint32_t v = keys[b];
  SegPlan p = plans[s];
  NodeEntry ent = table[v];
  float first_ts = ts[perm[b]];
  float last_ts = ts[perm[e - 1]];
  uint64_t addr = arena_base + (uint64_t)unit_off[s] * kUnit;
  unsigned long long dead = 0;
  if (p.dir_newcap) {
    BlockDesc *nd = reinterpret_cast<BlockDesc *>(addr);
    const BlockDesc *od = reinterpret_cast<const BlockDesc *>(ent.dir);
    uint32_t nlive = ent.end - ent.first;
    for (uint32_t i = 0; i < nlive; i++) nd[i] = od[ent.first + i];
    if (ent.dir) dead += dir_units(ent.dir_cap);
    ent.dir = addr;
    ent.first = 0;
    ent.end = nlive;
    ent.dir_cap = p.dir_newcap;
    addr += (uint64_t)dir_units(p.dir_newcap) * kUnit;
  }
  BlockDesc *dir = reinterpret_cast<BlockDesc *>(ent.dir);
  bool live = ent.end > ent.first;
  BlockDesc *tail = live ? &dir[ent.end - 1] : nullptr;
  SegInfo info;
  memset(&info, 0, sizeof(info));
  info.fill = p.fill;
  if (p.fill) {
    info.p0 = tail->payload;
    info.cap0 = tail->capacity;
    info.off0 = tail->size;
    tail->size += p.fill;
    tail->start_ts = fminf(tail->start_ts, first_ts);
    tail->end_ts = ts[perm[b + p.fill - 1]];
  }
  if (p.flags & kPlanNew) {
    BlockDesc d;
    d.payload = addr;
    d.size = cnt - p.fill;
    d.capacity = p.newcap;
    d.start_ts = fminf(FLT_MAX, ts[perm[b + p.fill]]);
    d.end_ts = last_ts;
    d.cum_before = live ? tail->cum_before + tail->size : 0u;
    d.reserved = 0;
    dir[ent.end] = d;
    ent.end++;
    info.p1 = addr;
    info.cap1 = p.newcap;
    info.off1 = 0;
    atomicAdd(&stats->num_blocks, 1ull);
    atomicAdd(&stats->allocated_elems, (unsigned long long)p.newcap);
  } else if (p.flags & kPlanRealloc) {
    info.old_payload = tail->payload;
    info.old_cap = tail->capacity;
    info.old_size = tail->size;
    info.p1 = addr;
    info.cap1 = p.newcap;
    info.off1 = tail->size;
    dead += payload_units(tail->capacity);
    atomicAdd(&stats->allocated_elems, (unsigned long long)p.newcap - tail->capacity);
    tail->payload = addr;
    tail->capacity = p.newcap;
    tail->size += cnt;
    tail->start_ts = fminf(tail->start_ts, first_ts);
    tail->end_ts = last_ts;
  }
  ent.num_edges += cnt;
  ent.num_insertions += 1;
  table[v] = ent;
  infos[s] = info;
  is_src[v] = 1;
  if (dead) atomicAdd(&stats->dead_units, dead);
}

// replace policy only: move the old payload of a reallocated block (CopyTemporalBlock, utils.cu:9-31)
__global__ void __launch_bounds__(kThreads) realloc_copy_kernel(const SegInfo *__restrict__ infos, uint32_t num_segments) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (s >= num_segments) return;
  SegInfo f = infos[s];
  if (!f.old_payload) return;
  const float *ots = blk_ts(f.old_payload);
  const int64_t *od = blk_dst(f.old_payload, f.old_cap), *oe = blk_eid(f.old_payload, f.old_cap);
  float *nts = const_cast<float *>(blk_ts(f.p1));
  int64_t *nd = const_cast<int64_t *>(blk_dst(f.p1, f.cap1)), *ne = const_cast<int64_t *>(blk_eid(f.p1, f.cap1));
  for (uint32_t i = lane; i < f.old_size; i += 32) {
    const float t = ots[i];
    nts[i] = t;
    nd[i] = od[i];
    ne[i] = oe[i];
    blk_store_pivots(f.p1, f.cap1, i, t);  // the new capacity has its own pivot geometry
  }
}

// One thread per edge in (src, ts) order: payload append + vertex / edge-id bookkeeping
// (CopyEdgesToBlock, utils.cu:45-57; nodes_/edges_ upkeep, dynamic_graph.cu:89-97).
__global__ void __launch_bounds__(kThreads) scatter_kernel(const uint32_t *__restrict__ perm,
                                                           const uint32_t *__restrict__ segid,
                                                           const uint32_t *__restrict__ seg_start,
                                                           const SegInfo *__restrict__ infos,
                                                           const int64_t *__restrict__ src,
                                                           const int64_t *__restrict__ dst,
                                                           const float *__restrict__ ts,
                                                           const int64_t *__restrict__ eid, uint64_t n,
                                                           uint8_t *is_node, uint32_t *eid_ref, GraphStats *stats) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool fresh = false;
  if (i < n) {
    uint32_t s = segid[i];
    uint32_t r = (uint32_t)i - seg_start[s];
    SegInfo f = infos[s];
    uint64_t p;
    uint32_t cap, pos;
    if (r < f.fill) {
      p = f.p0; cap = f.cap0; pos = f.off0 + r;
    } else {
      p = f.p1; cap = f.cap1; pos = f.off1 + (r - f.fill);
    }
    uint32_t j = perm[i];
    int64_t d = dst[j], e = eid[j];
    const float t = ts[j];
    const_cast<float *>(blk_ts(p))[pos] = t;
    blk_store_pivots(p, cap, pos, t);
    const_cast<int64_t *>(blk_dst(p, cap))[pos] = d;
    const_cast<int64_t *>(blk_eid(p, cap))[pos] = e;
    is_node[src[j]] = 1;
    is_node[d] = 1;
    fresh = atomicAdd(&eid_ref[e], 1u) == 0;
  }
  unsigned m = __ballot_sync(0xffffffffu, fresh);
  if (m && (threadIdx.x & 31) == 0) atomicAdd(&stats->num_edges, (unsigned long long)__popc(m));
}

// DynamicGraph::OffloadOldBlocks, dynamic_graph.cu:382-411: one warp per vertex, oldest block first.
// `drops` (optional) records (vertex, dir index) of every dropped block for the to_file path.
__global__ void __launch_bounds__(kThreads) offload_kernel(NodeEntry *table, const uint8_t *__restrict__ is_node,
                                                           uint64_t table_len, float timestamp, uint32_t *eid_ref,
                                                           GraphStats *stats, uint2 *drops, uint32_t drops_cap) {
  uint64_t v = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (v >= table_len || !is_node[v]) return;
  NodeEntry ent = table[v];
  const BlockDesc *dir = reinterpret_cast<const BlockDesc *>(ent.dir);
  uint32_t first = ent.first;
  unsigned long long dropped = 0, gone_edges = 0, cap_sum = 0, dead = 0;
  while (first < ent.end) {
    BlockDesc d = dir[first];
    if (!(d.end_ts < timestamp)) break;
    const int64_t *e = blk_eid(d.payload, d.capacity);
    for (uint32_t i = lane; i < d.size; i += 32)
      if (atomicSub(&eid_ref[e[i]], 1u) == 1u) gone_edges++;
    if (lane == 0 && drops) {
      unsigned long long k = atomicAdd(&stats->call_count, 1ull);
      if (k < drops_cap) drops[k] = make_uint2((uint32_t)v, first);
    }
    dropped++;
    cap_sum += d.capacity;
    dead += payload_units(d.capacity);
    first++;
  }
  if (!dropped) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gone_edges += __shfl_xor_sync(0xffffffffu, gone_edges, o);
  if (lane == 0) {
    table[v].first = first;
    if (!drops) atomicAdd(&stats->call_count, dropped);
    atomicAdd(&stats->num_blocks, 0ull - dropped);
    atomicAdd(&stats->allocated_elems, 0ull - cap_sum);
    atomicAdd(&stats->num_edges, 0ull - gone_edges);
    atomicAdd(&stats->dead_units, dead);
  }
}

__global__ void count_flags_kernel(const uint8_t *__restrict__ flags, uint64_t n, unsigned long long *out) {
  unsigned long long c = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    c += flags[i] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

__global__ void out_degree_kernel(const NodeEntry *__restrict__ table, uint64_t table_len, const int64_t *__restrict__ ids,
                                  uint64_t n, uint64_t *out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t v = ids[i];
  out[i] = (v >= 0 && (uint64_t)v < table_len) ? table[v].num_edges : 0;
}

// ------------------------------------------------------------------------------------------ host helpers
static int set_device(const gf_graph *g) {
  GF_CUDA(cudaSetDevice(g->cfg.device));
  return GF_OK;
}

static int ensure_table(gf_graph *g, int64_t max_id, cudaStream_t st) {
  size_t need = (size_t)max_id + 1;
  if (need > g->table_cap) {
    size_t cap = g->table_cap ? g->table_cap * 2 : 1024;
    if (cap < need) cap = need;
    NodeEntry *nt;
    uint8_t *nn, *ns;
    GF_CUDA(cudaMallocAsync(&nt, cap * sizeof(NodeEntry), st));
    GF_CUDA(cudaMallocAsync(&nn, cap, st));
    GF_CUDA(cudaMallocAsync(&ns, cap, st));
    size_t old = g->table_cap;
    if (old) {
      GF_CUDA(cudaMemcpyAsync(nt, g->d_table, old * sizeof(NodeEntry), cudaMemcpyDeviceToDevice, st));
      GF_CUDA(cudaMemcpyAsync(nn, g->d_is_node, old, cudaMemcpyDeviceToDevice, st));
      GF_CUDA(cudaMemcpyAsync(ns, g->d_is_src, old, cudaMemcpyDeviceToDevice, st));
      GF_CUDA(cudaFreeAsync(g->d_table, st));
      GF_CUDA(cudaFreeAsync(g->d_is_node, st));
      GF_CUDA(cudaFreeAsync(g->d_is_src, st));
    }
    GF_CUDA(cudaMemsetAsync(nt + old, 0, (cap - old) * sizeof(NodeEntry), st));
    GF_CUDA(cudaMemsetAsync(nn + old, 0, cap - old, st));
    GF_CUDA(cudaMemsetAsync(ns + old, 0, cap - old, st));
    g->d_table = nt;
    g->d_is_node = nn;
    g->d_is_src = ns;
    g->table_cap = cap;
  }
  // DynamicGraph::AddNodes, dynamic_graph.cu:140-147
  if (!g->has_nodes || max_id > g->max_node_id) g->max_node_id = max_id;
  g->has_nodes = true;
  return GF_OK;
}

constexpr uint64_t kMaxEid = 1ull << 31;

static int ensure_eids(gf_graph *g, int64_t max_eid, cudaStream_t st) {
  size_t need = (size_t)max_eid + 1;
  if (need <= g->eid_cap) return GF_OK;
  size_t cap = g->eid_cap ? g->eid_cap * 2 : 4096;
  if (cap < need) cap = need;
  uint32_t *nr;
  GF_CUDA(cudaMallocAsync(&nr, cap * sizeof(uint32_t), st));
  if (g->eid_cap) {
    GF_CUDA(cudaMemcpyAsync(nr, g->d_eid_ref, g->eid_cap * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    GF_CUDA(cudaFreeAsync(g->d_eid_ref, st));
  }
  GF_CUDA(cudaMemsetAsync(nr + g->eid_cap, 0, (cap - g->eid_cap) * sizeof(uint32_t), st));
  g->d_eid_ref = nr;
  g->eid_cap = cap;
  return GF_OK;
}

// bump allocation of `units` contiguous arena units; grows the arena chunk-wise up to maximum_pool_size
// (the reference's rmm pool_memory_resource(initial, maximum), temporal_block_allocator.cu:27-65)
static int arena_alloc(gf_graph *g, uint64_t units, uint64_t *base) {
  size_t bytes = units * kUnit;
  if (!g->chunks.empty()) {
    ArenaChunk &c = g->chunks.back();
    if (c.size - c.used >= bytes) {
      *base = (uint64_t)(uintptr_t)(c.base + c.used);
      c.used += bytes;
      return GF_OK;
    }
  }
  size_t maxp = g->cfg.maximum_pool_size ? g->cfg.maximum_pool_size : SIZE_MAX;
  size_t want = g->chunks.empty() ? (size_t)g->cfg.initial_pool_size : g->arena_total;  // double
  if (want < bytes) want = bytes;
  if (want < (1u << 20)) want = 1u << 20;
  if (g->arena_total + want > maxp) want = maxp > g->arena_total ? maxp - g->arena_total : 0;
  want = want / kUnit * kUnit;
  if (want < bytes)
    GF_FAIL(GF_ENOMEM, "edge pool exhausted: need %zu more bytes, pool holds %zu of maximum_pool_size %zu", bytes,
            g->arena_total, (size_t)g->cfg.maximum_pool_size);
  char *p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    GF_FAIL(GF_ENOMEM, "cudaMalloc(%zu) for the edge pool failed: %s", want, cudaGetErrorString(e));
  }
  g->chunks.push_back({p, want, bytes});
  g->arena_total += want;
  *base = (uint64_t)(uintptr_t)p;
  return GF_OK;
}

static int pull_stats(gf_graph *g, cudaStream_t st) {
  GF_CUDA(cudaMemcpyAsync(g->h_stats, g->d_stats, sizeof(GraphStats), cudaMemcpyDeviceToHost, st));
  GF_CUDA(cudaStreamSynchronize(st));
  return GF_OK;
}

static int bit_width_u64(uint64_t x) {
  int b = 0;
  while (x) {
    b++;
    x >>= 1;
  }
  return b;
}

static int add_edges_impl(gf_graph *g, const int64_t *src, const int64_t *dst, const float *ts, const int64_t *eid,
                          uint64_t n, int ptr_kind, cudaStream_t st) {
  if (n == 0) GF_FAIL(GF_EINVAL, "add_edges: empty batch (reference: CHECK_GT(src_nodes.size(), 0))");
  if (n >= (1ull << 31)) GF_FAIL(GF_EINVAL, "add_edges: batch of %llu edges exceeds 2^31-1", (unsigned long long)n);
  if (!src || !dst || !ts || !eid) GF_FAIL(GF_EINVAL, "add_edges: null array");
  GF_TRY(set_device(g));
  g->prof.begin(st);
  // ---- stage host input
  if (ptr_kind == GF_PTR_HOST) {
    size_t off_dst = align_up(n * 8, 256), off_eid = 2 * off_dst, off_ts = 3 * off_dst;
    GF_TRY(g->s_in.reserve(off_ts + align_up(n * 4, 256), st));
    char *b = g->s_in.as<char>();
    GF_CUDA(cudaMemcpyAsync(b, src, n * 8, cudaMemcpyHostToDevice, st));
    GF_CUDA(cudaMemcpyAsync(b + off_dst, dst, n * 8, cudaMemcpyHostToDevice, st));
    GF_CUDA(cudaMemcpyAsync(b + off_eid, eid, n * 8, cudaMemcpyHostToDevice, st));
    GF_CUDA(cudaMemcpyAsync(b + off_ts, ts, n * 4, cudaMemcpyHostToDevice, st));
    src = (const int64_t *)b;
    dst = (const int64_t *)(b + off_dst);
    eid = (const int64_t *)(b + off_eid);
    ts = (const float *)(b + off_ts);
  } else if (ptr_kind != GF_PTR_DEVICE) {
    GF_FAIL(GF_EINVAL, "add_edges: bad ptr_kind %d", ptr_kind);
  }
  const unsigned nb = cdiv(n, kThreads);
  // ---- pass 0: id range, eid range, is the batch already in time order?
  gf::launch(stats_reset_kernel, 1, 1, 0, st, g->d_stats);
  gf::launch(batch_stats_kernel, min(nb, 148u * 8), kThreads, 0, st, src, dst, ts, eid, n, g->d_stats);
  GF_CUDA(cudaGetLastError());
  GF_TRY(pull_stats(g, st));
  GraphStats hs = *g->h_stats;
  if (hs.batch_min_id < 0) GF_FAIL(GF_EINVAL, "add_edges: negative vertex id %lld", hs.batch_min_id);
  if (hs.batch_max_id >= (1ll << 32)) GF_FAIL(GF_EINVAL, "add_edges: vertex id %lld >= 2^32", hs.batch_max_id);
  if (hs.batch_min_eid < 0 || (uint64_t)hs.batch_max_eid >= kMaxEid)
    GF_FAIL(GF_EINVAL, "add_edges: edge ids must lie in [0, 2^31); got [%lld, %lld]", hs.batch_min_eid,
            hs.batch_max_eid);
  const int64_t old_max = g->max_node_id;
  const bool old_has = g->has_nodes;
  GF_TRY(ensure_table(g, std::max<int64_t>(hs.batch_max_id, g->has_nodes ? g->max_node_id : 0), st));
  GF_TRY(ensure_eids(g, hs.batch_max_eid, st));
  g->prof.end(0, st);
  // ---- sort by (src, ts), stable: LSD = [ts pass if needed] then src
  size_t sort_elems = 4 * align_up(n, 64) + radix_tmp_elems(n);
  GF_TRY(g->s_sort.reserve(sort_elems * 4, st));
  uint32_t *k0 = g->s_sort.as<uint32_t>(), *v0 = k0 + align_up(n, 64), *k1 = v0 + align_up(n, 64),
           *v1 = k1 + align_up(n, 64), *stmp = v1 + align_up(n, 64);
  bool in0 = true;
  if (hs.ts_unsorted) {
    gf::launch(keys_from_ts_kernel, nb, kThreads, 0, st, ts, n, k0, v0);
    GF_TRY(radix_sort_pairs(k0, v0, k1, v1, n, 0, 32, stmp, &in0, st));
    uint32_t *vs = in0 ? v0 : v1;
    // regenerate keys from src in time order; keep values where they are
    gf::launch(keys_from_src_kernel, nb, kThreads, 0, st, src, vs, n, in0 ? k0 : k1, nullptr);
  } else {
    gf::launch(keys_from_src_kernel, nb, kThreads, 0, st, src, nullptr, n, k0, v0);
  }
  {
    int bits = bit_width_u64((uint64_t)hs.batch_max_id);
    if (bits < 1) bits = 1;
    bool r0;
    uint32_t *ka = in0 ? k0 : k1, *va = in0 ? v0 : v1, *kb = in0 ? k1 : k0, *vb = in0 ? v1 : v0;
    GF_TRY(radix_sort_pairs(ka, va, kb, vb, n, 0, (bits + 7) / 8 * 8, stmp, &r0, st));
    if (!r0) { uint32_t *t = ka; ka = kb; kb = t; t = va; va = vb; vb = t; }
    k0 = ka; v0 = va; k1 = kb; v1 = vb;  // (k0, v0) = sorted keys + permutation; (k1, v1) free
  }
  const uint32_t *keys = k0, *perm = v0;
  g->prof.end(1, st);
  // ---- segments (one per distinct source vertex)
  size_t nseg = align_up(n + 1, 64);
  size_t seg_bytes = nseg * 4 * 4 + scan_tmp_elems(n) * 4 + nseg * sizeof(SegPlan) + nseg * sizeof(SegInfo);
  GF_TRY(g->s_seg.reserve(seg_bytes, st));
  uint32_t *segid = g->s_seg.as<uint32_t>(), *excl = segid + nseg, *seg_start = excl + nseg, *units = seg_start + nseg,
           *sctmp = units + nseg;
  SegPlan *plans = reinterpret_cast<SegPlan *>(sctmp + align_up(scan_tmp_elems(n), 64));
  SegInfo *infos = reinterpret_cast<SegInfo *>(plans + nseg);
  gf::launch(seg_heads_kernel, nb, kThreads, 0, st, keys, n, segid);
  GF_TRY(exclusive_scan_u32(segid, excl, n, nullptr, sctmp, st));
  gf::launch(seg_starts_kernel, nb, kThreads, 0, st, segid, excl, n, seg_start, g->d_stats);
  // ---- plan + allocation sizes
  StoreParams sp = {(uint32_t)g->cfg.minimum_block_size, g->cfg.insertion_policy, g->cfg.adaptive_block_size};
  gf::launch(plan_kernel, nb, kThreads, 0, st, keys, perm, seg_start, ts, n, g->d_table, sp, plans, units, g->d_stats);
  uint32_t *unit_off = excl;  // excl is dead after seg_starts
  GF_TRY(exclusive_scan_u32(units, unit_off, n, &g->d_stats->total_units, sctmp, st));
  GF_CUDA(cudaGetLastError());
  GF_TRY(pull_stats(g, st));
  g->prof.end(2, st);
  hs = *g->h_stats;
  if (hs.error_flags & kErrOutOfOrder) {
    g->prof.stop();
    g->max_node_id = old_max;
    g->has_nodes = old_has;
    GF_FAIL(GF_EORDER, "add_edges: timestamps are older than the existing edges in the graph");
  }
  uint64_t base = 0;
  if (hs.total_units) {
    int rc = arena_alloc(g, hs.total_units, &base);
    if (rc != GF_OK) {
      g->prof.stop();
      g->max_node_id = old_max;
      g->has_nodes = old_has;
      return rc;
    }
  }
  // ---- commit + scatter
  const uint32_t U = hs.num_segments;
  gf::launch(commit_kernel, cdiv(U, kThreads), kThreads, 0, st, keys, perm, seg_start, ts, U, g->d_table, plans, unit_off, base,
                                                        infos, g->d_is_src, g->d_stats);
  g->prof.end(3, st);
  if (g->cfg.insertion_policy == GF_INSERTION_REPLACE)
    gf::launch(realloc_copy_kernel, cdiv((uint64_t)U * 32, kThreads), kThreads, 0, st, infos, U);
  gf::launch(scatter_kernel, nb, kThreads, 0, st, perm, segid, seg_start, infos, src, dst, ts, eid, n, g->d_is_node,
                                          g->d_eid_ref, g->d_stats);
  GF_CUDA(cudaGetLastError());
  g->prof.end(4, st, false);
  g->counts_dirty = true;
  // the reference returns after cudaStreamSynchronize (dynamic_graph.cu:135-137); host buffers were staged,
  // so only the stats mirror needs the sync
  GF_TRY(pull_stats(g, st));
  return GF_OK;
}

static int refresh_counts(gf_graph *g) {
  if (!g->counts_dirty) return GF_OK;
  GF_TRY(set_device(g));
  cudaStream_t st = 0;
  size_t len = g->table_len();
  unsigned long long *d = &g->d_stats->call_count;
  unsigned long long h[2] = {0, 0};
  for (int k = 0; k < 2; k++) {
    GF_CUDA(cudaMemsetAsync(d, 0, sizeof(*d), st));
    if (len) gf::launch(count_flags_kernel, min(cdiv(len, kThreads), 148u * 8), kThreads, 0, st, k ? g->d_is_src : g->d_is_node, len, d);
    GF_CUDA(cudaMemcpyAsync(&h[k], d, sizeof(*d), cudaMemcpyDeviceToHost, st));
    GF_CUDA(cudaStreamSynchronize(st));
  }
  g->num_nodes = h[0];
  g->num_src_nodes = h[1];
  g->counts_dirty = false;
  return GF_OK;
}

static int read_entry(gf_graph *g, int64_t v, NodeEntry *ent, std::vector<BlockDesc> *descs) {
  memset(ent, 0, sizeof(*ent));
  descs->clear();
  if (v < 0 || (size_t)v >= g->table_len()) return GF_OK;
  GF_TRY(set_device(g));
  GF_CUDA(cudaDeviceSynchronize());
  GF_CUDA(cudaMemcpy(ent, g->d_table + v, sizeof(NodeEntry), cudaMemcpyDeviceToHost));
  uint32_t nlive = ent->end - ent->first;
  if (nlive) {
    descs->resize(nlive);
    GF_CUDA(cudaMemcpy(descs->data(), reinterpret_cast<const BlockDesc *>(ent->dir) + ent->first,
                       nlive * sizeof(BlockDesc), cudaMemcpyDeviceToHost));
  }
  return GF_OK;
}

static int flags_to_list(gf_graph *g, const uint8_t *d_flags, size_t len, int64_t *out, uint64_t cap, uint64_t *count) {
  GF_TRY(set_device(g));
  std::vector<uint8_t> h(len);
  GF_CUDA(cudaDeviceSynchronize());
  if (len) GF_CUDA(cudaMemcpy(h.data(), d_flags, len, cudaMemcpyDeviceToHost));
  uint64_t k = 0;
  for (size_t i = 0; i < len; i++)
    if (h[i]) {
      if (out && k < cap) out[k] = (int64_t)i;
      k++;
    }
  *count = k;
  if (out && cap < k) GF_FAIL(GF_ECAPACITY, "output buffer holds %llu entries, %llu needed", (unsigned long long)cap, (unsigned long long)k);
  return GF_OK;
}

// SaveToFile, temporal_block_allocator.cu:182-222
static int save_block_file(gf_graph *g, int64_t v, const BlockDesc &d, uint64_t prev, uint64_t next) {
  if ((size_t)v >= g->saved_blocks_per_node.size()) g->saved_blocks_per_node.resize(v + 1, 0);
  char name[128];
  snprintf(name, sizeof(name), "temporal_block_%lld-%u.bin", (long long)v, g->saved_blocks_per_node[v]);
  std::vector<int64_t> hd(d.size), he(d.size);
  std::vector<float> ht(d.size);
  GF_CUDA(cudaMemcpy(hd.data(), (const void *)(d.payload + payload_ts_bytes(d.capacity)), d.size * 8ull, cudaMemcpyDeviceToHost));
  GF_CUDA(cudaMemcpy(ht.data(), (const void *)d.payload, d.size * 4ull, cudaMemcpyDeviceToHost));
  GF_CUDA(cudaMemcpy(he.data(), (const void *)(d.payload + payload_ts_bytes(d.capacity) + payload_i64_bytes(d.capacity)),
                     d.size * 8ull, cudaMemcpyDeviceToHost));
  FILE *f = fopen(name, "wb");
  if (!f) GF_FAIL(GF_EINVAL, "cannot open %s for writing", name);
  size_t size = d.size, capacity = d.capacity;
  fwrite(&size, sizeof(size), 1, f);
  fwrite(&capacity, sizeof(capacity), 1, f);
  fwrite(&d.start_ts, 4, 1, f);
  fwrite(&d.end_ts, 4, 1, f);
  fwrite(hd.data(), 8, d.size, f);
  fwrite(ht.data(), 4, d.size, f);
  fwrite(he.data(), 8, d.size, f);
  fwrite(&prev, 8, 1, f);
  fwrite(&next, 8, 1, f);
  fclose(f);
  g->saved_blocks_per_node[v]++;
  return GF_OK;
}

}  // namespace gf

using namespace gf;

// ============================================================================================== C ABI
GF_EXPORT const char *gf_last_error(void) { return gf::get_error(); }
GF_EXPORT int gf_abi_version(void) { return GF_ABI_VERSION; }

GF_EXPORT int gf_graph_create(const gf_graph_config *cfg, gf_graph **out) {
  if (!cfg || !out) GF_FAIL(GF_EINVAL, "gf_graph_create: null argument");
  if (cfg->insertion_policy != GF_INSERTION_INSERT && cfg->insertion_policy != GF_INSERTION_REPLACE)
    GF_FAIL(GF_EINVAL, "Invalid insertion policy: %d", cfg->insertion_policy);
  if (cfg->mem_resource_type < GF_MEM_CUDA || cfg->mem_resource_type > GF_MEM_SHARED)
    GF_FAIL(GF_EINVAL, "Invalid memory resource type: %d", cfg->mem_resource_type);
  if (cfg->minimum_block_size >= (1ull << 31)) GF_FAIL(GF_EINVAL, "minimum_block_size too large");
  if (cfg->maximum_pool_size && cfg->maximum_pool_size < cfg->initial_pool_size)
    GF_FAIL(GF_EINVAL, "maximum_pool_size < initial_pool_size");
  int ndev = 0;
  GF_CUDA(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) GF_FAIL(GF_EINVAL, "device %d out of range (%d devices)", cfg->device, ndev);
  GF_CUDA(cudaSetDevice(cfg->device));
  gf_graph *g = new gf_graph();
  g->cfg = *cfg;
  cudaError_t e = cudaMalloc(&g->d_stats, sizeof(GraphStats));
  if (e == cudaSuccess) e = cudaMemset(g->d_stats, 0, sizeof(GraphStats));
  if (e == cudaSuccess) e = cudaMallocHost(&g->h_stats, sizeof(GraphStats));
  if (e != cudaSuccess) {
    delete g;
    GF_FAIL(GF_ECUDA, "gf_graph_create: %s", cudaGetErrorString(e));
  }
  memset(g->h_stats, 0, sizeof(GraphStats));
  g->prof.init(GF_GRAPH_PHASES);
  *out = g;
  return GF_OK;
}

GF_EXPORT int gf_graph_destroy(gf_graph *g) {
  if (!g) return GF_OK;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    if (--g->refs > 0) return GF_OK;
  }
  cudaSetDevice(g->cfg.device);
  cudaDeviceSynchronize();
  g->prof.destroy();
  for (auto &c : g->chunks) cudaFree(c.base);
  if (g->d_table) cudaFree(g->d_table);
  if (g->d_is_node) cudaFree(g->d_is_node);
  if (g->d_is_src) cudaFree(g->d_is_src);
  if (g->d_eid_ref) cudaFree(g->d_eid_ref);
  if (g->d_stats) cudaFree(g->d_stats);
  if (g->h_stats) cudaFreeHost(g->h_stats);
  g->s_in.release();
  g->s_sort.release();
  g->s_seg.release();
  g->s_misc.release();
  delete g;
  return GF_OK;
}

GF_EXPORT int gf_graph_add_edges(gf_graph *g, const int64_t *src, const int64_t *dst, const float *ts,
                                 const int64_t *eid, uint64_t n, int ptr_kind, void *stream) {
  if (!g) GF_FAIL(GF_EINVAL, "null graph");
  std::lock_guard<std::mutex> lk(g->mu);
  return add_edges_impl(g, src, dst, ts, eid, n, ptr_kind, (cudaStream_t)stream);
}

GF_EXPORT int gf_graph_clear(gf_graph *g, void *stream) {
  if (!g) GF_FAIL(GF_EINVAL, "null graph");
  std::lock_guard<std::mutex> lk(g->mu);
  cudaStream_t st = (cudaStream_t)stream;
  GF_TRY(set_device(g));
  size_t len = g->table_len();
  if (len) {
    GF_CUDA(cudaMemsetAsync(g->d_table, 0, len * sizeof(NodeEntry), st));
    GF_CUDA(cudaMemsetAsync(g->d_is_node, 0, len, st));
    GF_CUDA(cudaMemsetAsync(g->d_is_src, 0, len, st));
  }
  if (g->eid_cap) GF_CUDA(cudaMemsetAsync(g->d_eid_ref, 0, g->eid_cap * sizeof(uint32_t), st));
  GF_CUDA(cudaMemsetAsync(g->d_stats, 0, sizeof(GraphStats), st));
  memset(g->h_stats, 0, sizeof(GraphStats));
  for (auto &c : g->chunks) c.used = 0;
  // bump allocation only ever looks at the last chunk: keep the largest one last
  std::sort(g->chunks.begin(), g->chunks.end(), [](const ArenaChunk &a, const ArenaChunk &b) { return a.size < b.size; });
  g->max_node_id = 0;
  g->has_nodes = false;
  g->counts_dirty = false;
  g->num_nodes = g->num_src_nodes = 0;
  g->saved_blocks_per_node.clear();
  return GF_OK;
}

GF_EXPORT int gf_graph_offload_old_blocks(gf_graph *g, float timestamp, int to_file, uint64_t *num_blocks,
                                          void *stream) {
  if (!g) GF_FAIL(GF_EINVAL, "null graph");
  std::lock_guard<std::mutex> lk(g->mu);
  cudaStream_t st = (cudaStream_t)stream;
  GF_TRY(set_device(g));
  if (num_blocks) *num_blocks = 0;
  size_t len = g->table_len();
  if (!len) return GF_OK;
  GF_CUDA(cudaMemsetAsync(&g->d_stats->call_count, 0, sizeof(unsigned long long), st));
  uint2 *drops = nullptr;
  uint32_t drops_cap = 0;
  std::vector<NodeEntry> before;
  if (to_file) {
    // the dropped descriptors stay readable in the directories; remember the pre-offload `first` per vertex
    GF_TRY(pull_stats(g, st));
    drops_cap = (uint32_t)g->h_stats->num_blocks;
    GF_TRY(g->s_misc.reserve((size_t)drops_cap * sizeof(uint2) + 16, st));
    drops = g->s_misc.as<uint2>();
  }
  gf::launch(offload_kernel, cdiv(len * 32, kThreads), kThreads, 0, st, g->d_table, g->d_is_node, len, timestamp, g->d_eid_ref,
                                                                g->d_stats, drops, drops_cap);
  GF_CUDA(cudaGetLastError());
  GF_TRY(pull_stats(g, st));
  uint64_t nd = g->h_stats->call_count;
  if (num_blocks) *num_blocks = nd;
  if (to_file && nd) {
    std::vector<uint2> h(nd);
    GF_CUDA(cudaMemcpy(h.data(), drops, nd * sizeof(uint2), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end(), [](const uint2 &a, const uint2 &b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    for (auto &d : h) {
      NodeEntry ent;
      GF_CUDA(cudaMemcpy(&ent, g->d_table + d.x, sizeof(ent), cudaMemcpyDeviceToHost));
      BlockDesc bd;
      const BlockDesc *dir = reinterpret_cast<const BlockDesc *>(ent.dir);
      GF_CUDA(cudaMemcpy(&bd, dir + d.y, sizeof(bd), cudaMemcpyDeviceToHost));
      uint64_t prev = d.y > 0 ? (uint64_t)(uintptr_t)(dir + d.y - 1) : 0;
      uint64_t next = d.y + 1 < ent.end ? (uint64_t)(uintptr_t)(dir + d.y + 1) : 0;
      GF_TRY(save_block_file(g, d.x, bd, prev, next));
    }
  }
  return GF_OK;
}

GF_EXPORT int gf_graph_num_vertices(gf_graph *g, uint64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(refresh_counts(g));
  *out = g->num_nodes;
  return GF_OK;
}
GF_EXPORT int gf_graph_num_source_vertices(gf_graph *g, uint64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(refresh_counts(g));
  *out = g->num_src_nodes;
  return GF_OK;
}
GF_EXPORT int gf_graph_num_edges(gf_graph *g, uint64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  *out = g->h_stats->num_edges;
  return GF_OK;
}
GF_EXPORT int gf_graph_max_vertex_id(gf_graph *g, int64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  *out = g->max_node_id;
  return GF_OK;
}
GF_EXPORT int gf_graph_avg_linked_list_length(gf_graph *g, float *out) {  // dynamic_graph.cu:359-366
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(refresh_counts(g));
  float sum = (float)g->h_stats->num_blocks;
  *out = sum / (float)g->num_nodes;
  return GF_OK;
}
GF_EXPORT int gf_graph_memory_usage(gf_graph *g, float *out) {  // temporal_block_allocator.cu:155-156
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  *out = (float)(g->h_stats->allocated_elems * 20ull);
  return GF_OK;
}
GF_EXPORT int gf_graph_metadata_memory_usage(gf_graph *g, float *out) {  // dynamic_graph.cu:372-380
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  float sum = 0;
  sum += 64 * g->h_stats->num_blocks;  // sizeof(TemporalBlock) == 64 (common.h:35-48)
  sum += 8 * g->table_len();
  *out = sum;
  return GF_OK;
}
GF_EXPORT int gf_graph_device_bytes(gf_graph *g, uint64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  *out = g->arena_total + g->table_cap * (sizeof(NodeEntry) + 2) + g->eid_cap * 4 + g->s_in.cap + g->s_sort.cap +
         g->s_seg.cap + g->s_misc.cap;
  return GF_OK;
}

GF_EXPORT int gf_graph_out_degree(gf_graph *g, const int64_t *ids, uint64_t n, uint64_t *out) {
  if (!g || (n && (!ids || !out))) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  if (!n) return GF_OK;
  GF_TRY(set_device(g));
  cudaStream_t st = 0;
  GF_TRY(g->s_misc.reserve(n * 16, st));
  int64_t *d_ids = g->s_misc.as<int64_t>();
  uint64_t *d_out = reinterpret_cast<uint64_t *>(d_ids + n);
  GF_CUDA(cudaMemcpyAsync(d_ids, ids, n * 8, cudaMemcpyHostToDevice, st));
  gf::launch(out_degree_kernel, cdiv(n, kThreads), kThreads, 0, st, g->d_table, g->table_len(), d_ids, n, d_out);
  GF_CUDA(cudaMemcpyAsync(out, d_out, n * 8, cudaMemcpyDeviceToHost, st));
  GF_CUDA(cudaStreamSynchronize(st));
  return GF_OK;
}

GF_EXPORT int gf_graph_nodes(gf_graph *g, int64_t *out, uint64_t cap, uint64_t *count) {
  if (!g || !count) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  return flags_to_list(g, g->d_is_node, g->table_len(), out, cap, count);
}
GF_EXPORT int gf_graph_src_nodes(gf_graph *g, int64_t *out, uint64_t cap, uint64_t *count) {
  if (!g || !count) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  return flags_to_list(g, g->d_is_src, g->table_len(), out, cap, count);
}
GF_EXPORT int gf_graph_edges(gf_graph *g, int64_t *out, uint64_t cap, uint64_t *count) {
  if (!g || !count) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(set_device(g));
  std::vector<uint32_t> h(g->eid_cap);
  GF_CUDA(cudaDeviceSynchronize());
  if (g->eid_cap) GF_CUDA(cudaMemcpy(h.data(), g->d_eid_ref, g->eid_cap * 4, cudaMemcpyDeviceToHost));
  uint64_t k = 0;
  for (size_t i = 0; i < h.size(); i++)
    if (h[i]) {
      if (out && k < cap) out[k] = (int64_t)i;
      k++;
    }
  *count = k;
  if (out && cap < k) GF_FAIL(GF_ECAPACITY, "output buffer too small");
  return GF_OK;
}

GF_EXPORT int gf_graph_get_temporal_neighbors(gf_graph *g, int64_t vertex, int64_t *dst, float *ts, int64_t *eid,
                                              uint64_t cap, uint64_t *count) {
  if (!g || !count) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  NodeEntry ent;
  std::vector<BlockDesc> descs;
  GF_TRY(read_entry(g, vertex, &ent, &descs));
  uint64_t total = 0;
  for (auto &d : descs) total += d.size;
  *count = total;
  if (!dst || !ts || !eid) return GF_OK;
  if (cap < total) GF_FAIL(GF_ECAPACITY, "output buffer too small");
  uint64_t k = 0;
  std::vector<int64_t> hd, he;
  std::vector<float> ht;
  for (size_t b = descs.size(); b-- > 0;) {  // newest block first, each block reversed (dynamic_graph.cu:305-333)
    const BlockDesc &d = descs[b];
    hd.resize(d.size);
    he.resize(d.size);
    ht.resize(d.size);
    if (!d.size) continue;
    GF_CUDA(cudaMemcpy(ht.data(), (const void *)d.payload, d.size * 4ull, cudaMemcpyDeviceToHost));
    GF_CUDA(cudaMemcpy(hd.data(), (const void *)(d.payload + payload_ts_bytes(d.capacity)), d.size * 8ull, cudaMemcpyDeviceToHost));
    GF_CUDA(cudaMemcpy(he.data(), (const void *)(d.payload + payload_ts_bytes(d.capacity) + payload_i64_bytes(d.capacity)),
                       d.size * 8ull, cudaMemcpyDeviceToHost));
    for (uint32_t i = d.size; i-- > 0;) {
      dst[k] = hd[i];
      ts[k] = ht[i];
      eid[k] = he[i];
      k++;
    }
  }
  return GF_OK;
}

GF_EXPORT int gf_graph_block_shapes(gf_graph *g, int64_t vertex, uint64_t *sizes, uint64_t *caps, float *start_ts,
                                    float *end_ts, uint64_t cap, uint64_t *count) {
  if (!g || !count) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  NodeEntry ent;
  std::vector<BlockDesc> descs;
  GF_TRY(read_entry(g, vertex, &ent, &descs));
  *count = descs.size();
  if (!sizes) return GF_OK;
  if (cap < descs.size()) GF_FAIL(GF_ECAPACITY, "output buffer too small");
  for (size_t i = 0; i < descs.size(); i++) {
    sizes[i] = descs[i].size;
    caps[i] = descs[i].capacity;
    start_ts[i] = descs[i].start_ts;
    end_ts[i] = descs[i].end_ts;
  }
  return GF_OK;
}

GF_EXPORT int gf_graph_set_profiling(gf_graph *g, int on) {
  if (!g) GF_FAIL(GF_EINVAL, "null graph");
  g->prof.on = on != 0;
  return GF_OK;
}
GF_EXPORT int gf_graph_get_profile(gf_graph *g, double *ms, uint64_t *count, int reset) {
  if (!g || !ms || !count) GF_FAIL(GF_EINVAL, "null argument");
  g->prof.collect();
  for (int i = 0; i < GF_GRAPH_PHASES; i++) { ms[i] = g->prof.ms[i]; count[i] = g->prof.count[i]; }
  if (reset) g->prof.reset();
  return GF_OK;
}
GF_EXPORT uint64_t gf_debug_launch_count(void) { return gf::g_launch_count.load(); }
