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

// CTA-wide sums of K u64 values (kThreads threads, all of them must call); thread 0 ends up with the totals.
// Counters shared by the whole graph get ONE atomic per CTA: per-warp atomics on a single address serialise in L2
// and were 80 % of the commit / scatter time at multi-million-edge batches.
template <int K>
__device__ __forceinline__ void block_sum_u64(unsigned long long (&v)[K]) {
  __shared__ unsigned long long part[K][kThreads / 32];
#pragma unroll
  for (int k = 0; k < K; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0) part[k][threadIdx.x >> 5] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      unsigned long long t = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; w++) t += part[k][w];
      v[k] = t;
    }
  }
}

// Pass 0 over the batch: validation flags, id ranges, sort keys (src) + identity permutation; clears the scratch
// slot of the next call.
__global__ void __launch_bounds__(kThreads) prep_kernel(const int64_t *__restrict__ src, const int64_t *__restrict__ dst,
                                                        const float *__restrict__ ts, const int64_t *__restrict__ eid,
                                                        uint64_t n, uint64_t table_cap, uint64_t eid_cap,
                                                        int assume_sorted, uint32_t *__restrict__ keys,
                                                        uint32_t *__restrict__ vals, CallScratch *cur, CallScratch *nxt) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    nxt->max_id = nxt->max_eid = 0;
    nxt->error_flags = nxt->num_segments = nxt->total_units = nxt->accepted = nxt->unsorted = 0;
  }
  long long mx = 0, emx = 0;
  unsigned flags = 0;
  if (i < n) {
    const long long s = src[i], d = dst[i], e = eid[i];
    mx = max(s, d);
    emx = e;
    if (s < 0 || d < 0 || mx >= (1ll << 32)) flags |= kErrBadId;
    else if ((uint64_t)mx >= table_cap) flags |= kErrTableSmall;
    if (e < 0 || e >= (1ll << 31)) flags |= kErrBadEid;
    else if ((uint64_t)e >= eid_cap) flags |= kErrEidSmall;
    if (i + 1 < n && ts[i + 1] < ts[i]) flags |= kErrUnsorted;
    keys[i] = (uint32_t)s;
    vals[i] = (uint32_t)i;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    emx = max(emx, __shfl_xor_sync(0xffffffffu, emx, o));
    flags |= __shfl_xor_sync(0xffffffffu, flags, o);
  }
  __shared__ long long s_mx[kThreads / 32], s_emx[kThreads / 32];
  __shared__ unsigned s_flags[kThreads / 32];
  if ((threadIdx.x & 31) == 0) {
    s_mx[threadIdx.x >> 5] = mx;
    s_emx[threadIdx.x >> 5] = emx;
    s_flags[threadIdx.x >> 5] = flags;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < kThreads / 32; w++) {
      mx = max(mx, s_mx[w]);
      emx = max(emx, s_emx[w]);
      flags |= s_flags[w];
    }
    // the maxima only grow: a stale read can only cause a redundant atomic
    if (mx > *(volatile long long *)&cur->max_id) atomicMax(&cur->max_id, mx);
    if (emx > *(volatile long long *)&cur->max_eid) atomicMax(&cur->max_eid, emx);
    if (flags & kErrUnsorted) {
      cur->unsorted = 1;
      if (!assume_sorted) flags &= ~kErrUnsorted;  // the timestamp sort pass is already scheduled
    }
    if (flags) atomicOr(&cur->error_flags, flags);
  }
}

__global__ void keys_from_ts_kernel(const float *__restrict__ ts, uint64_t n, uint32_t *keys, uint32_t *vals) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t = ts[i];
  keys[i] = t == 0.0f ? orderable_f32(0.0f) : orderable_f32(t);  // -0.0 == +0.0 under operator<
  vals[i] = (uint32_t)i;
}
// keys[i] = src[vals[i]]
__global__ void keys_from_src_kernel(const int64_t *__restrict__ src, const uint32_t *__restrict__ vals_in, uint64_t n,
                                     uint32_t *keys) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = (uint32_t)src[vals_in[i]];
}

// segments of equal keys, fused into one look-back scan: in(i) = "element i starts a segment",
// out: segid[i], seg_start[segment], seg_start[U] = n, num_segments = U
struct SegIn {
  const uint32_t *keys;
  __device__ uint32_t operator()(uint64_t i) const { return (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u; }
};
struct SegOut {
  uint32_t *segid, *seg_start;
  uint64_t n;
  CallScratch *cur;
  __device__ void operator()(uint64_t i, uint32_t excl, uint32_t head) const {
    const uint32_t sid = excl + head - 1;
    segid[i] = sid;
    if (head) seg_start[sid] = (uint32_t)i;
    if (i == n - 1) {
      seg_start[sid + 1] = (uint32_t)n;
      cur->num_segments = sid + 1;
    }
  }
};

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

// One element per source vertex of the batch: validation + the block-sizing policy of
// DynamicGraph::AddEdgesForOneNode (dynamic_graph.cu:206-287) + TemporalBlockAllocator::AlignUp
// (temporal_block_allocator.cu:83-88), fused into the look-back scan of the allocation sizes: in(s) plans segment s
// and returns its arena units, out(s, offset) records where its allocation starts.  Nothing is mutated here.
struct PlanIn {
  const uint32_t *keys, *perm, *seg_start;
  const float *ts;
  const NodeEntry *table;
  StoreParams sp;
  SegPlan *plans;
  CallScratch *cur;
  __device__ uint32_t operator()(uint64_t s) const {
    if (cur->error_flags & ~kErrOutOfOrder) return 0;  // ids may be out of range: do not touch the table
    if (s >= cur->num_segments) return 0;
    const uint32_t b = seg_start[s], e = seg_start[s + 1];
    const uint32_t cnt = e - b;
    const uint32_t v = keys[b];
    const float first_ts = ts[perm[b]];
    const NodeEntry ent = table[v];
    const bool live = ent.end > ent.first;
    SegPlan p = {0, 0, 0, 0};
    if (!live) {
      p.newcap = max(cnt, sp.min_block);
      p.flags = kPlanNew;
    } else {
      const BlockDesc t = reinterpret_cast<const BlockDesc *>(ent.dir)[ent.end - 1];
      if (first_ts < t.end_ts) atomicOr(&cur->error_flags, kErrOutOfOrder);
      if ((uint64_t)t.size + cnt > t.capacity) {
        if (sp.policy == GF_INSERTION_INSERT) {
          p.fill = t.capacity - t.size;
          const uint32_t rem = cnt - p.fill;
          const uint64_t avg = ent.num_insertions == 0 ? rem : ent.num_edges / ent.num_insertions;
          const uint32_t ns = sp.adaptive ? next_pow2_u32((uint32_t)max((uint64_t)rem, avg)) : rem;
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
    if ((p.flags & kPlanNew) && ent.end == ent.dir_cap) {
      const uint32_t nlive = ent.end - ent.first;
      p.dir_newcap = max(4u, (2 * (nlive + 1) + 3) & ~3u);
      u += dir_units(p.dir_newcap);
    }
    if (p.newcap) u += payload_units(p.newcap);
    plans[s] = p;
    return u;
  }
};
struct PlanOut {
  uint32_t *unit_off;
  __device__ void operator()(uint64_t s, uint32_t excl, uint32_t) const { unit_off[s] = excl; }
};

// the commit kernel decides: any flag, or an arena chunk that cannot hold the batch => nothing is changed; the
// decision is recorded in cur->accepted for the kernels after it (which must not re-read the bump pointer)
__device__ __forceinline__ bool batch_rejected(GraphStats *stats, CallScratch *cur, bool reporter, int async) {
  // asynchronous ingest: once a queued batch is rejected every later one must be a no-op too (the host replays them
  // in order at the next flush); synchronous calls never see the flag set
  bool rejected = (cur->error_flags & ~kErrArena) != 0 || (async && stats->poison);
  if (!rejected && stats->arena_cur + (unsigned long long)cur->total_units * kUnit > stats->arena_end) {
    if (reporter) atomicOr(&cur->error_flags, kErrArena);
    rejected = true;
  }
  if (reporter) {
    cur->accepted = rejected ? 0u : 1u;
    if (rejected && async) stats->poison = 1u;
  }
  return rejected;
}

// One thread per source vertex: applies the plan (InsertBlock / Reallocate / CopyEdgesToBlock header updates,
// dynamic_graph.cu:153-174, temporal_block_allocator.cu:122-132, utils.cu:58-62) and tells the scatter kernel
// where the payload goes.
__global__ void __launch_bounds__(kThreads) commit_kernel(const uint32_t *__restrict__ keys,
                                                          const uint32_t *__restrict__ perm,
                                                          const uint32_t *__restrict__ seg_start,
                                                          const float *__restrict__ ts, NodeEntry *table,
                                                          const SegPlan *__restrict__ plans,
                                                          const uint32_t *__restrict__ unit_off, SegInfo *infos,
                                                          uint8_t *is_src, uint8_t *is_node, GraphStats *stats,
                                                          CallScratch *cur, int async) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (batch_rejected(stats, cur, s == 0, async)) return;
  if ((uint64_t)blockIdx.x * blockDim.x >= cur->num_segments) return;  // whole CTA idle
  unsigned long long agg[3] = {0, 0, 0};  // new blocks, added capacity, dead arena units
  if (s < cur->num_segments) {
  const uint64_t arena_base = stats->arena_cur;
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
    // positions (cum_before) are relative to the oldest LIVE block: re-base them, since `first` restarts at 0 and the
    // sampler reads "window starts before the oldest stored edge" as position 0 (blocks dropped by offload_old_blocks
    // must not count)
    const uint32_t rebase = nlive ? od[ent.first].cum_before : 0u;
    for (uint32_t i = 0; i < nlive; i++) {
      BlockDesc c = od[ent.first + i];
      c.cum_before -= rebase;
      nd[i] = c;
    }
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
    d.min_ts = live ? tail->min_ts : d.start_ts;  // start_ts of the vertex's oldest live block travels with the tail
    dir[ent.end] = d;
    ent.end++;
    info.p1 = addr;
    info.cap1 = p.newcap;
    info.off1 = 0;
    agg[0] = 1;
    agg[1] = p.newcap;
  } else if (p.flags & kPlanRealloc) {
    info.old_payload = tail->payload;
    info.old_cap = tail->capacity;
    info.old_size = tail->size;
    info.p1 = addr;
    info.cap1 = p.newcap;
    info.off1 = tail->size;
    dead += payload_units(tail->capacity);
    agg[1] = (unsigned long long)p.newcap - tail->capacity;
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
  is_node[v] = 1;
  agg[2] = dead;
  }
  block_sum_u64(agg);
  if (threadIdx.x == 0) {
    if (agg[0]) atomicAdd(&stats->num_blocks, agg[0]);
    if (agg[1]) atomicAdd(&stats->allocated_elems, agg[1]);
    if (agg[2]) atomicAdd(&stats->dead_units, agg[2]);
  }
}

// replace policy only: move the old payload of a reallocated block (CopyTemporalBlock, utils.cu:9-31)
__global__ void __launch_bounds__(kThreads) realloc_copy_kernel(const SegInfo *__restrict__ infos, const GraphStats *stats,
                                                                CallScratch *cur) {
  uint32_t s = (uint32_t)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  if (!cur->accepted) return;
  if (s >= cur->num_segments) return;
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
                                                           const int64_t *__restrict__ dst,
                                                           const float *__restrict__ ts,
                                                           const int64_t *__restrict__ eid, uint64_t n,
                                                           uint8_t *is_node, uint32_t *eid_ref, GraphStats *stats,
                                                           CallScratch *cur) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (!cur->accepted) return;
  if (i == 0) stats->arena_cur += (unsigned long long)cur->total_units * kUnit;  // nobody reads it after the commit kernel
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
    if (!is_node[d]) is_node[d] = 1;  // hot vertices: test first, thousands of identical byte stores serialise in L2
                                      // (the source vertex was flagged once per segment by the commit kernel)
    fresh = atomicAdd(&eid_ref[e], 1u) == 0;
  }
  unsigned long long nf[1] = {fresh ? 1ull : 0ull};
  block_sum_u64(nf);
  if (threadIdx.x == 0 && nf[0]) atomicAdd(&stats->num_edges, nf[0]);
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
    if (first < ent.end)  // the newest descriptor carries the oldest live timestamp (sampler's window-start shortcut)
      const_cast<BlockDesc *>(dir)[ent.end - 1].min_ts = dir[first].start_ts;
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

// host reads (getters) must see every mutation enqueued so far: all entry points but gf_graph_clear end with a stream
// synchronisation of their own, so only that one stream can still be busy
static int settle(gf_graph *g) {
  if (g->unsettled) {
    GF_CUDA(cudaStreamSynchronize(g->unsettled_stream));
    g->unsettled = false;
  }
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

// The payload arena is a bump allocator whose pointer lives on the device (GraphStats::arena_cur/arena_end), so a
// batch is planned, sized and committed without a host round trip.  The host only adds a chunk when the device
// reports kErrArena: `bytes` more are needed; chunks double up to maximum_pool_size (the reference's rmm
// pool_memory_resource(initial, maximum), temporal_block_allocator.cu:27-65).
static int arena_add_chunk(gf_graph *g, size_t bytes, cudaStream_t st) {
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
  g->chunks.push_back({p, want, 0});
  g->arena_total += want;
  unsigned long long ptrs[2] = {(unsigned long long)(uintptr_t)p, (unsigned long long)(uintptr_t)(p + want)};
  GF_CUDA(cudaMemcpyAsync(&g->d_stats->arena_cur, ptrs, sizeof(ptrs), cudaMemcpyHostToDevice, st));
  GF_CUDA(cudaStreamSynchronize(st));  // ptrs is a stack variable
  return GF_OK;
}

static int pull_stats(gf_graph *g, cudaStream_t st) {
  GF_CUDA(cudaMemcpyAsync(g->h_stats, g->d_stats, sizeof(GraphStats), cudaMemcpyDeviceToHost, st));
  GF_CUDA(cudaStreamSynchronize(st));
  if (!g->chunks.empty() && g->h_stats->arena_cur) g->chunks.back().used = (size_t)(g->h_stats->arena_cur - (uintptr_t)g->chunks.back().base);
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

static int ensure_lb(gf_graph *g, uint64_t tiles, cudaStream_t st) {
  if (tiles > g->lb_tiles || !g->s_lb.ptr) {
    size_t want = std::max<size_t>(tiles * 2, 1024);
    Scratch n;
    GF_TRY(n.reserve(256 + want * 8, st));
    GF_CUDA(cudaMemsetAsync(n.ptr, 0, n.cap, st));  // generation 0 == never written
    if (g->s_lb.ptr) GF_CUDA(cudaFreeAsync(g->s_lb.ptr, st));
    g->s_lb = n;
    g->lb_tiles = want;
  }
  if (++g->lb_gen >= (1ull << 30)) {
    GF_CUDA(cudaMemsetAsync(g->s_lb.ptr, 0, g->s_lb.cap, st));
    g->lb_gen = 1;
  }
  return GF_OK;
}
static LookbackCtl lb_ctl(gf_graph *g) {
  return {g->s_lb.as<unsigned int>(), reinterpret_cast<unsigned long long *>(g->s_lb.as<char>() + 256), g->lb_gen};
}

// One attempt = 5 + 3 * (sort passes) kernel launches and ONE host synchronisation, whatever the batch touches:
//   prep -> radix sort by (src, ts) -> segments (fused look-back scan) -> plan + allocation offsets (fused look-back
//   scan) -> commit -> scatter.  Capacity problems (vertex table, edge-id table, arena chunk) and a batch that is
//   not in time order are detected on the device, leave the graph untouched, and make the host fix the cause and
//   replay the batch.
static int add_edges_impl(gf_graph *g, const int64_t *src, const int64_t *dst, const float *ts, const int64_t *eid,
                          uint64_t n, int ptr_kind, cudaStream_t st, bool async);

// Look at the outcome of the batches queued by gf_graph_add_edges_async (one synchronisation for all of them).  The
// first rejected batch -- table / edge-id / arena capacity, a batch out of time order, a bad id -- and every batch
// queued after it changed nothing on the device (GraphStats::poison); they are replayed here, in order, through the
// synchronous path, which fixes the cause or reports the error.
static int flush_pending(gf_graph *g) {
  if (g->pending.empty()) return GF_OK;
  cudaStream_t st = g->pending_stream;
  GF_TRY(set_device(g));
  GF_TRY(pull_stats(g, st));
  std::vector<gf_graph::Pending> q;
  q.swap(g->pending);
  size_t j = 0;
  for (; j < q.size(); j++) {
    const CallScratch hs = g->h_stats->call[q[j].slot];
    if (!hs.accepted) break;
    if (!g->has_nodes || hs.max_id > g->max_node_id) g->max_node_id = hs.max_id;
    g->has_nodes = true;
    g->counts_dirty = true;
  }
  if (j == q.size()) return GF_OK;
  GF_CUDA(cudaMemsetAsync(&g->d_stats->poison, 0, sizeof(unsigned int), st));
  for (; j < q.size(); j++) {
    const int rc = add_edges_impl(g, q[j].src, q[j].dst, q[j].ts, q[j].eid, q[j].n, GF_PTR_DEVICE, st, false);
    if (rc != GF_OK) return rc;  // the batches queued after the failing one are dropped
  }
  return GF_OK;
}

static int add_edges_impl(gf_graph *g, const int64_t *src, const int64_t *dst, const float *ts, const int64_t *eid,
                          uint64_t n, int ptr_kind, cudaStream_t st, bool async) {
  if (n == 0) GF_FAIL(GF_EINVAL, "add_edges: empty batch (reference: CHECK_GT(src_nodes.size(), 0))");
  if (n >= (1ull << 30)) GF_FAIL(GF_EINVAL, "add_edges: batch of %llu edges exceeds 2^30-1", (unsigned long long)n);
  if (!src || !dst || !ts || !eid) GF_FAIL(GF_EINVAL, "add_edges: null array");
  if (async && ptr_kind != GF_PTR_DEVICE) GF_FAIL(GF_EINVAL, "add_edges_async: device arrays only");
  if (async && !g->pending.empty() && g->pending_stream != st) GF_TRY(flush_pending(g));  // one stream per queue
  if (async && g->pending.size() + 2 >= kCallRing) GF_TRY(flush_pending(g));              // a slot per queued batch
  if (!async) GF_TRY(flush_pending(g));
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
  const uint64_t scan_tiles = (n + kScanTile - 1) / kScanTile;
  // ---- scratch
  const size_t na = align_up(n + 1, 64);
  GF_TRY(g->s_sort.reserve((4 * na + radix_tmp_elems(n)) * 4, st));
  GF_TRY(g->s_seg.reserve(na * 3 * 4 + na * sizeof(SegPlan) + na * sizeof(SegInfo), st));
  uint32_t *segid = g->s_seg.as<uint32_t>(), *seg_start = segid + na, *unit_off = seg_start + na;
  SegPlan *plans = reinterpret_cast<SegPlan *>(unit_off + na);
  SegInfo *infos = reinterpret_cast<SegInfo *>(plans + na);
  const StoreParams sp = {(uint32_t)g->cfg.minimum_block_size, g->cfg.insertion_policy, g->cfg.adaptive_block_size};

  for (int attempt = 0; attempt < 8; attempt++) {
    const unsigned parity = g->call_parity;
    g->call_parity = (parity + 1) % kCallRing;
    CallScratch *cur = &g->d_stats->call[parity], *nxt = &g->d_stats->call[g->call_parity];
    uint32_t *k0 = g->s_sort.as<uint32_t>(), *v0 = k0 + na, *k1 = v0 + na, *v1 = k1 + na, *stmp = v1 + na;
    const bool fast = !g->expect_unsorted;
    // ---- pass 0: validation flags, id ranges, keys = src, identity permutation
    gf::launch(prep_kernel, nb, kThreads, 0, st, src, dst, ts, eid, n, (uint64_t)g->table_cap, (uint64_t)g->eid_cap,
               fast ? 1 : 0, k0, v0, cur, nxt);
    g->prof.end(0, st);
    // ---- sort by (src, ts), stable: LSD = [ts pass unless the batch is in time order] then src
    bool in0 = true;
    if (!fast) {
      gf::launch(keys_from_ts_kernel, nb, kThreads, 0, st, ts, n, k0, v0);
      GF_TRY(radix_sort_pairs(k0, v0, k1, v1, n, 0, 32, stmp, &in0, st));
      gf::launch(keys_from_src_kernel, nb, kThreads, 0, st, src, in0 ? v0 : v1, n, in0 ? k0 : k1);
    }
    {
      int bits = bit_width_u64(g->table_cap ? (uint64_t)g->table_cap - 1 : 0);
      if (bits < 1) bits = 1;
      bool r0;
      uint32_t *ka = in0 ? k0 : k1, *va = in0 ? v0 : v1, *kb = in0 ? k1 : k0, *vb = in0 ? v1 : v0;
      GF_TRY(radix_sort_pairs(ka, va, kb, vb, n, 0, (bits + 7) / 8 * 8, stmp, &r0, st));
      if (!r0) { uint32_t *t = ka; ka = kb; kb = t; t = va; va = vb; vb = t; }
      k0 = ka; v0 = va;  // sorted keys + permutation
    }
    const uint32_t *keys = k0, *perm = v0;
    g->prof.end(1, st);
    // ---- segments (one per distinct source vertex), then plan + allocation offsets
    GF_TRY(ensure_lb(g, scan_tiles, st));
    gf::launch(scan_lookback_kernel<SegIn, SegOut>, (unsigned)scan_tiles, kScanThreads, 0, st, n, SegIn{keys},
               SegOut{segid, seg_start, n, cur}, lb_ctl(g), (uint32_t *)nullptr);
    GF_TRY(ensure_lb(g, scan_tiles, st));
    gf::launch(scan_lookback_kernel<PlanIn, PlanOut>, (unsigned)scan_tiles, kScanThreads, 0, st, n,
               PlanIn{keys, perm, seg_start, ts, g->d_table, sp, plans, cur}, PlanOut{unit_off}, lb_ctl(g),
               &cur->total_units);
    g->prof.end(2, st);
    // ---- commit + scatter (no-ops when any flag is up or the arena chunk is too small)
    gf::launch(commit_kernel, nb, kThreads, 0, st, keys, perm, seg_start, ts, g->d_table, plans, unit_off, infos,
               g->d_is_src, g->d_is_node, g->d_stats, cur, async ? 1 : 0);
    g->prof.end(3, st);
    if (g->cfg.insertion_policy == GF_INSERTION_REPLACE)
      gf::launch(realloc_copy_kernel, cdiv(n * 32, kThreads), kThreads, 0, st, infos, g->d_stats, cur);
    gf::launch(scatter_kernel, nb, kThreads, 0, st, perm, segid, seg_start, infos, dst, ts, eid, n, g->d_is_node,
               g->d_eid_ref, g->d_stats, cur);
    GF_CUDA(cudaGetLastError());
    g->prof.end(4, st, false);
    if (async) {  // the caller keeps the arrays alive until the next flush; the outcome is looked at there
      g->pending.push_back({src, dst, ts, eid, n, parity});
      g->pending_stream = st;
      return GF_OK;
    }
    // the reference returns after cudaStreamSynchronize (dynamic_graph.cu:135-137); this is the only sync
    GF_TRY(pull_stats(g, st));
    const CallScratch hs = g->h_stats->call[parity];
    const uint32_t f = hs.error_flags;
    if (f & kErrBadId) GF_FAIL(GF_EINVAL, "add_edges: vertex ids must lie in [0, 2^32)");
    if (f & kErrBadEid) GF_FAIL(GF_EINVAL, "add_edges: edge ids must lie in [0, 2^31)");
    if (f & (kErrTableSmall | kErrEidSmall | kErrUnsorted | kErrArena)) {  // fix the cause, replay the batch
      if (f & kErrTableSmall) {
        const int64_t keep_max = g->max_node_id;
        const bool keep_has = g->has_nodes;
        GF_TRY(ensure_table(g, hs.max_id, st));
        g->max_node_id = keep_max;  // the table grew; the graph has not changed yet
        g->has_nodes = keep_has;
      }
      if (f & kErrEidSmall) GF_TRY(ensure_eids(g, hs.max_eid, st));
      if (f & kErrUnsorted) g->expect_unsorted = true;
      if ((f & kErrArena) && !(f & (kErrTableSmall | kErrEidSmall | kErrUnsorted)))
        GF_TRY(arena_add_chunk(g, (size_t)hs.total_units * kUnit, st));
      if (g->prof.on) g->prof.begin(st);
      continue;
    }
    if (f & kErrOutOfOrder) GF_FAIL(GF_EORDER, "add_edges: timestamps are older than the existing edges in the graph");
    if (!hs.accepted) GF_FAIL(GF_ECUDA, "add_edges: internal error (batch neither accepted nor flagged)");
    // success: host mirrors (DynamicGraph::AddNodes, dynamic_graph.cu:140-147)
    if (!g->has_nodes || hs.max_id > g->max_node_id) g->max_node_id = hs.max_id;
    g->has_nodes = true;
    g->counts_dirty = true;
    if (!fast && !hs.unsorted) g->expect_unsorted = false;  // a time-ordered stream resumes the fast path
    return GF_OK;
  }
  GF_FAIL(GF_ECUDA, "add_edges: the batch could not be applied after 8 attempts");
}

static int refresh_counts(gf_graph *g) {
  if (!g->counts_dirty) return GF_OK;
  GF_TRY(set_device(g));
  GF_TRY(settle(g));
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
  GF_TRY(settle(g));
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
  GF_TRY(settle(g));
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
  if (cfg->initial_pool_size) {  // the reference's pool resource reserves initial_pool_size up front as well
    int rc = arena_add_chunk(g, 0, 0);
    if (rc != GF_OK) {
      gf_graph_destroy(g);
      return rc;
    }
  }
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
  g->s_lb.release();
  delete g;
  return GF_OK;
}

GF_EXPORT int gf_graph_add_edges(gf_graph *g, const int64_t *src, const int64_t *dst, const float *ts,
                                 const int64_t *eid, uint64_t n, int ptr_kind, void *stream) {
  if (!g) GF_FAIL(GF_EINVAL, "null graph");
  std::lock_guard<std::mutex> lk(g->mu);
  return add_edges_impl(g, src, dst, ts, eid, n, ptr_kind, (cudaStream_t)stream, false);
}

GF_EXPORT int gf_graph_add_edges_async(gf_graph *g, const int64_t *src, const int64_t *dst, const float *ts,
                                       const int64_t *eid, uint64_t n, void *stream) {
  if (!g) GF_FAIL(GF_EINVAL, "null graph");
  std::lock_guard<std::mutex> lk(g->mu);
  return add_edges_impl(g, src, dst, ts, eid, n, GF_PTR_DEVICE, (cudaStream_t)stream, true);
}

GF_EXPORT int gf_graph_flush(gf_graph *g) {
  if (!g) GF_FAIL(GF_EINVAL, "null graph");
  std::lock_guard<std::mutex> lk(g->mu);
  return flush_pending(g);
}

// for the sampler and the cache (other translation units): settle queued batches before the graph is read
int gf_graph_flush_internal(gf_graph *g) {
  std::lock_guard<std::mutex> lk(g->mu);
  return flush_pending(g);
}

GF_EXPORT int gf_graph_clear(gf_graph *g, void *stream) {
  if (!g) GF_FAIL(GF_EINVAL, "null graph");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
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
  if (!g->chunks.empty()) {
    const ArenaChunk &c = g->chunks.back();
    g->h_stats->arena_cur = (unsigned long long)(uintptr_t)c.base;
    g->h_stats->arena_end = (unsigned long long)(uintptr_t)(c.base + c.size);
    GF_CUDA(cudaMemcpyAsync(&g->d_stats->arena_cur, &g->h_stats->arena_cur, 16, cudaMemcpyHostToDevice, st));
  }
  g->unsettled_stream = st;
  g->unsettled = true;
  g->call_parity = 0;
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
  GF_TRY(flush_pending(g));
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

GF_EXPORT int gf_block_file_read(const char *path, uint64_t *size, uint64_t *capacity, float *start_ts, float *end_ts,
                                 int64_t *dst, float *ts, int64_t *eid, uint64_t cap) {
  if (!path || !size || !capacity || !start_ts || !end_ts) GF_FAIL(GF_EINVAL, "gf_block_file_read: null argument");
  if ((dst || ts || eid) && !(dst && ts && eid)) GF_FAIL(GF_EINVAL, "gf_block_file_read: dst / ts / eid go together");
  FILE *f = fopen(path, "rb");
  if (!f) GF_FAIL(GF_EINVAL, "cannot open %s", path);
  uint64_t hdr[2];
  float tt[2];
  bool ok = fread(hdr, 8, 2, f) == 2 && fread(tt, 4, 2, f) == 2;
  if (ok) {
    fseek(f, 0, SEEK_END);
    ok = hdr[0] <= hdr[1] && (uint64_t)ftell(f) == 24 + hdr[0] * 20 + 16;  // header, three arrays, prev / next
  }
  if (!ok) {
    fclose(f);
    GF_FAIL(GF_EINVAL, "%s is not a temporal block file", path);
  }
  *size = hdr[0];
  *capacity = hdr[1];
  *start_ts = tt[0];
  *end_ts = tt[1];
  if (dst) {
    if (cap < hdr[0]) {
      fclose(f);
      GF_FAIL(GF_ECAPACITY, "output arrays hold %llu entries, the block has %llu", (unsigned long long)cap, (unsigned long long)hdr[0]);
    }
    fseek(f, 24, SEEK_SET);
    ok = fread(dst, 8, hdr[0], f) == hdr[0] && fread(ts, 4, hdr[0], f) == hdr[0] && fread(eid, 8, hdr[0], f) == hdr[0];
  }
  fclose(f);
  if (!ok) GF_FAIL(GF_EINVAL, "short read from %s", path);
  return GF_OK;
}

GF_EXPORT int gf_graph_num_vertices(gf_graph *g, uint64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  GF_TRY(refresh_counts(g));
  *out = g->num_nodes;
  return GF_OK;
}
GF_EXPORT int gf_graph_num_source_vertices(gf_graph *g, uint64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  GF_TRY(refresh_counts(g));
  *out = g->num_src_nodes;
  return GF_OK;
}
GF_EXPORT int gf_graph_num_edges(gf_graph *g, uint64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  *out = g->h_stats->num_edges;
  return GF_OK;
}
GF_EXPORT int gf_graph_max_vertex_id(gf_graph *g, int64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  GF_TRY(gf_graph_flush_internal(g));
  *out = g->max_node_id;
  return GF_OK;
}
GF_EXPORT int gf_graph_avg_linked_list_length(gf_graph *g, float *out) {  // dynamic_graph.cu:359-366
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  GF_TRY(refresh_counts(g));
  float sum = (float)g->h_stats->num_blocks;
  *out = sum / (float)g->num_nodes;
  return GF_OK;
}
GF_EXPORT int gf_graph_memory_usage(gf_graph *g, float *out) {  // temporal_block_allocator.cu:155-156
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  GF_TRY(gf_graph_flush_internal(g));
  *out = (float)(g->h_stats->allocated_elems * 20ull);
  return GF_OK;
}
GF_EXPORT int gf_graph_metadata_memory_usage(gf_graph *g, float *out) {  // dynamic_graph.cu:372-380
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  GF_TRY(gf_graph_flush_internal(g));
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
  GF_TRY(flush_pending(g));
  if (!n) return GF_OK;
  GF_TRY(set_device(g));
  GF_TRY(settle(g));
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
  GF_TRY(flush_pending(g));
  return flags_to_list(g, g->d_is_node, g->table_len(), out, cap, count);
}
GF_EXPORT int gf_graph_src_nodes(gf_graph *g, int64_t *out, uint64_t cap, uint64_t *count) {
  if (!g || !count) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  return flags_to_list(g, g->d_is_src, g->table_len(), out, cap, count);
}
GF_EXPORT int gf_graph_edges(gf_graph *g, int64_t *out, uint64_t cap, uint64_t *count) {
  if (!g || !count) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  GF_TRY(set_device(g));
  std::vector<uint32_t> h(g->eid_cap);
  GF_TRY(settle(g));
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
  GF_TRY(flush_pending(g));
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
  GF_TRY(flush_pending(g));
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
