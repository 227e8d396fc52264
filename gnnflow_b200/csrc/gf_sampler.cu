// Temporal neighbour sampler: SampleLayerRecent / SampleLayerUniform and the SamplingResult assembly, fully
// on the device.  Replaces reference gnnflow/csrc/sampling_kernels.cu:11-273, utils.cu:96-109 (LowerBound),
// temporal_sampler.cu:97-305 (SampleLayer / Sample, incl. the thrust::remove_if compaction and the host-side
// row/col assembly).
//
// One sampling step = three launches, no host round trip:
//   locate  : each warp owns a target (node, timestamp): window arithmetic, then a warp-cooperative search of
//             the vertex's block directory and of one block's timestamps gives the position range
//             [P_lo, P_hi) of the in-window edges -> (descriptor address, idx_hi, #candidates), and the
//             number of neighbours the target will emit.
//   scan    : exclusive scan of the per-target counts = output offsets (this IS the compaction; there are no
//             "-1" slots to remove).
//   emit    : each warp owns 32 consecutive targets and walks their concatenated output slots 32 at a time,
//             so every store (neighbour id, eid, ts, dt, row, col) is a full coalesced warp store.
#include <algorithm>

#include "gf_primitives.cuh"
#include "gf_store.cuh"

namespace gf {

constexpr int kSThreads = 256;
constexpr int kSWarps = kSThreads / 32;

struct TargetLoc {   // 16 bytes, one per target
  uint64_t desc;     // address of the BlockDesc that holds the newest in-window edge boundary (0 = none)
  uint32_t idx_hi;   // in that block: edges [0, idx_hi) are older than the window end
  uint32_t ncand;    // number of in-window edges (P_hi - P_lo)
};

struct SampleParams {
  const NodeEntry *table;
  uint64_t table_len;
  uint32_t fanout;
  uint32_t num_snapshots;
  uint32_t snapshot_idx;
  float window;
  int prop_time;
  int policy;
  uint64_t seed;
  uint64_t launch_index;
};

// window arithmetic of sampling_kernels.cu:27-40.  The reference is compiled with --use_fast_math, which
// contracts `root - float(u) * w` into one FFMA: written out explicitly here.
__device__ __forceinline__ void window_of(float root, const SampleParams &p, float &start, float &end) {
  if (p.num_snapshots == 1) {
    start = ((double)fabsf(p.window) < 1e-6) ? 0.0f : __fsub_rn(root, p.window);
    end = root;
  } else {
    end = __fmaf_rn(-(float)(p.num_snapshots - p.snapshot_idx - 1), p.window, root);
    start = __fsub_rn(end, p.window);
  }
}

// Philox4x32-10, the shared counter-based stream (oracle/gnnflow_oracle.c D2)
__device__ __forceinline__ uint32_t philox_u32(uint64_t seed, uint32_t tid, uint64_t launch_index) {
  uint32_t c0 = tid, c1 = (uint32_t)launch_index, c2 = (uint32_t)(launch_index >> 32), c3 = 0;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c0;
}

// ------------------------------------------------------------------------------------ search helpers
// first index in [0, n) with key(idx) >= x, keys non-decreasing.  Scalar version (one thread).
template <class KeyAt>
__device__ __forceinline__ uint32_t lower_bound_scalar(KeyAt key_at, uint32_t n, float x) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (key_at(mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Warp-cooperative version: every lane calls with the same arguments and gets the same answer.  Each round
// issues 32 independent probes (one per lane) and narrows the range 33-fold, so a block of n timestamps costs
// ceil(log33(n)) dependent memory round trips instead of log2(n).
template <class KeyAt>
__device__ __forceinline__ uint32_t lower_bound_warp(KeyAt key_at, uint32_t n, float x, int lane) {
  uint32_t lo = 0, hi = n;  // answer in [lo, hi]
  while (hi - lo > 32) {
    uint32_t len = hi - lo;
    uint32_t step = (len + 32) / 33;  // 32 probes cut [lo,hi) into <= 33 pieces of <= step
    uint32_t idx = lo + (lane + 1) * step - 1;
    bool less = idx < hi && key_at(idx) < x;
    uint32_t c = __popc(__ballot_sync(0xffffffffu, less));  // probes are monotone: c leading trues
    uint32_t nlo = lo + c * step;
    uint32_t nhi = min(hi, lo + (c + 1) * step - 1);
    lo = nlo;
    hi = nhi < nlo ? nlo : nhi;
  }
  uint32_t idx = lo + lane;
  bool less = idx < hi && key_at(idx) < x;
  return lo + __popc(__ballot_sync(0xffffffffu, less));
}

// timestamps of one block: 128-bit loads, 128 timestamps per round trip (the array is 16-byte aligned and
// padded to 16 bytes, gf_common.cuh)
__device__ __forceinline__ uint32_t lower_bound_ts_warp(const float *ts, uint32_t n, float x, int lane) {
  uint32_t lo = 0, hi = n;
  while (hi - (lo & ~3u) > 128) {
    uint32_t len = hi - lo;
    uint32_t step = (len + 32) / 33;
    uint32_t idx = lo + (lane + 1) * step - 1;
    bool less = idx < hi && __ldg(ts + idx) < x;
    uint32_t c = __popc(__ballot_sync(0xffffffffu, less));
    uint32_t nlo = lo + c * step;
    uint32_t nhi = min(hi, lo + (c + 1) * step - 1);
    lo = nlo;
    hi = nhi < nlo ? nlo : nhi;
  }
  uint32_t base = lo & ~3u;  // aligned window [base, base + 128) covers [lo, hi)
  uint32_t i0 = base + lane * 4;
  uint32_t c = 0;
  if (i0 < hi) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(ts + i0));
    c += (i0 + 0 >= lo && i0 + 0 < hi && v.x < x);
    c += (i0 + 1 >= lo && i0 + 1 < hi && v.y < x);
    c += (i0 + 2 >= lo && i0 + 2 < hi && v.z < x);
    c += (i0 + 3 >= lo && i0 + 3 < hi && v.w < x);
  }
  return lo + __reduce_add_sync(0xffffffffu, c);
}

struct Located {
  uint64_t desc;
  uint32_t idx;
  uint32_t pos;  // global position = cum_before + idx
  uint32_t rel;  // index of the block relative to the oldest live block
};

// Position of the first edge with ts >= x in the live blocks [first, end) of one vertex.
//   b* = oldest block with end_ts >= x; position = cum_before[b*] + lower_bound(ts[b*], x).
//   no such block -> position = total, reported against the tail block.
// Equivalent to summing the reference's per-block LowerBound over its tail->prev walk
// (sampling_kernels.cu:44-93) because block time ranges are ordered (add_edges rejects out-of-order batches).
template <bool WARP>
__device__ __forceinline__ Located locate_pos(const BlockDesc *dir, uint32_t first, uint32_t end, const BlockDesc &tail,
                                              float x, int lane) {
  Located r;
  const BlockDesc *d;
  BlockDesc blk;
  if (tail.end_ts < x) {  // everything stored is older than x
    r.desc = (uint64_t)(uintptr_t)(dir + end - 1);
    r.idx = tail.size;
    r.pos = tail.cum_before + tail.size;
    r.rel = end - 1 - first;
    return r;
  }
  if (end - first == 1) {
    d = dir + first;
    blk = tail;
  } else {
    auto end_ts_at = [&](uint32_t i) { return __ldg(&dir[first + i].end_ts); };
    uint32_t b = WARP ? lower_bound_warp(end_ts_at, end - first - 1, x, lane)
                      : lower_bound_scalar(end_ts_at, end - first - 1, x);  // tail.end_ts >= x already known
    d = dir + first + b;
    if (first + b == end - 1) {
      blk = tail;
    } else {
      const uint4 *q = reinterpret_cast<const uint4 *>(d);
      uint4 a = __ldg(q), c = __ldg(q + 1);
      blk.payload = ((uint64_t)a.y << 32) | a.x;
      blk.size = a.z;
      blk.capacity = a.w;
      blk.start_ts = __uint_as_float(c.x);
      blk.end_ts = __uint_as_float(c.y);
      blk.cum_before = c.z;
      blk.min_ts = __uint_as_float(c.w);
    }
  }
  uint32_t idx;
  if (x <= blk.start_ts) {
    idx = 0;  // the whole block is >= x (shortcut cases of sampling_kernels.cu:66-86)
  } else {
    const float *ts = blk_ts(blk.payload);
    idx = WARP ? lower_bound_ts_warp(ts, blk.size, x, lane)
               : lower_bound_scalar([&](uint32_t i) { return __ldg(ts + i); }, blk.size, x);
  }
  r.desc = (uint64_t)(uintptr_t)d;
  r.idx = idx;
  r.pos = blk.cum_before + idx;
  r.rel = (uint32_t)(d - (dir + first));
  return r;
}

__device__ __forceinline__ BlockDesc load_desc(const BlockDesc *d) {
  const U8x32 q = ldg256_b32(d);
  BlockDesc b;
  b.payload = ((uint64_t)q.w[1] << 32) | q.w[0];
  b.size = q.w[2];
  b.capacity = q.w[3];
  b.start_ts = __uint_as_float(q.w[4]);
  b.end_ts = __uint_as_float(q.w[5]);
  b.cum_before = q.w[6];
  b.min_ts = __uint_as_float(q.w[7]);
  return b;
}
// POLICY template arguments: -1 = read SampleParams::policy at run time; GF_SAMPLING_RECENT / GF_SAMPLING_UNIFORM = the
// kernel is compiled for that policy alone (the persistent kernel: the other policy's code and registers are gone)
template <int POLICY>
__device__ __forceinline__ bool is_uniform(const SampleParams &p) {
  return POLICY < 0 ? p.policy == GF_SAMPLING_UNIFORM : POLICY == GF_SAMPLING_UNIFORM;
}
template <int POLICY = -1>
__device__ __forceinline__ uint32_t count_of(const SampleParams &p, uint32_t ncand) {
  // recent: slot k is valid iff k < #in-window edges (sampling_kernels.cu:88-105);
  // uniform: with replacement, every slot valid iff there is a candidate (:202, oracle D1)
  return !is_uniform<POLICY>(p) ? min(p.fanout, ncand) : (ncand ? p.fanout : 0u);
}

// batch lookup for the multi-batch launch: largest b with batch_offsets[b] <= i
__device__ __forceinline__ uint32_t batch_of(const uint64_t *__restrict__ batch_offsets, uint32_t num_batches, uint64_t i) {
  uint32_t lo = 0, hi = num_batches;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (batch_offsets[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------ locate
// One target, one thread: window arithmetic + scalar searches.  Returns the number of neighbours the target emits.
__device__ __forceinline__ uint32_t locate_thread(const SampleParams &p, int64_t nid, float root, TargetLoc &loc,
                                                  uint32_t &back) {
  loc.desc = 0; loc.idx_hi = 0; loc.ncand = 0;
  back = 0;
  float start, end;
  window_of(root, p, start, end);
  if (nid < 0 || (uint64_t)nid >= p.table_len) return 0;  // oracle D3
  NodeEntry ent = load_entry64(p.table + nid);
  if (ent.end <= ent.first) return 0;
  const BlockDesc *dir = reinterpret_cast<const BlockDesc *>(ent.dir());
  BlockDesc tail = ent.tail;
  Located hi = locate_pos<false>(dir, ent.first, ent.end, tail, end, 0);
  Located lo = locate_pos<false>(dir, ent.first, ent.end, tail, start, 0);
  loc.desc = hi.desc;
  loc.idx_hi = hi.idx;
  loc.ncand = hi.pos > lo.pos ? hi.pos - lo.pos : 0u;
  back = hi.rel;
  return count_of(p, loc.ncand);
}

// variant 0: warp-cooperative.  Each warp takes `tpw` (<= 32) consecutive targets: lanes first fetch their own
// target, vertex entry and tail descriptor (32 independent dependent-load chains in flight), then the warp
// resolves the targets that have edges one at a time with cooperative searches.
__global__ void __launch_bounds__(kSThreads) locate_warp_kernel(SampleParams p, const int64_t *__restrict__ nodes,
                                                                const float *__restrict__ root_ts, uint64_t T_bound,
                                                                const uint32_t *__restrict__ T_dev, int tpw,
                                                                TargetLoc *__restrict__ locs,
                                                                uint32_t *__restrict__ counts,
                                                                uint32_t *__restrict__ nback) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t T = T_dev ? (uint64_t)*T_dev : T_bound;
  const uint64_t i = warp * tpw + lane;
  const bool in_range = lane < tpw && i < T_bound;
  const bool valid = in_range && i < T;
  float start = 0.f, end = 0.f;
  NodeEntry ent;
  ent.dir_tagged = 0; ent.first = ent.end = 0;
  BlockDesc tail;
  tail.size = 0; tail.cum_before = 0; tail.end_ts = 0.f; tail.start_ts = 0.f; tail.payload = 0; tail.capacity = 0;
  if (valid) {
    int64_t nid = nodes[i];
    window_of(root_ts[i], p, start, end);
    if (nid >= 0 && (uint64_t)nid < p.table_len) ent = load_entry64(p.table + nid);
    if (ent.end > ent.first) tail = ent.tail;
  }
  const uint64_t my_dir = ent.dir();
  TargetLoc mine = {0, 0, 0};
  uint32_t my_back = 0;
  unsigned todo = __ballot_sync(0xffffffffu, valid && ent.end > ent.first);
  while (todo) {
    int j = __ffs(todo) - 1;
    todo &= todo - 1;
    const BlockDesc *dir = reinterpret_cast<const BlockDesc *>(__shfl_sync(0xffffffffu, my_dir, j));
    uint32_t first = __shfl_sync(0xffffffffu, ent.first, j), last = __shfl_sync(0xffffffffu, ent.end, j);
    float s = __shfl_sync(0xffffffffu, start, j), e = __shfl_sync(0xffffffffu, end, j);
    BlockDesc t;
    t.payload = __shfl_sync(0xffffffffu, tail.payload, j);
    t.size = __shfl_sync(0xffffffffu, tail.size, j);
    t.capacity = __shfl_sync(0xffffffffu, tail.capacity, j);
    t.start_ts = __shfl_sync(0xffffffffu, tail.start_ts, j);
    t.end_ts = __shfl_sync(0xffffffffu, tail.end_ts, j);
    t.cum_before = __shfl_sync(0xffffffffu, tail.cum_before, j);
    Located hi = locate_pos<true>(dir, first, last, t, e, lane);
    Located lo = locate_pos<true>(dir, first, last, t, s, lane);
    if (lane == j) {
      mine.desc = hi.desc;
      mine.idx_hi = hi.idx;
      mine.ncand = hi.pos > lo.pos ? hi.pos - lo.pos : 0u;
      my_back = hi.rel;
    }
  }
  if (in_range) {
    locs[i] = mine;
    counts[i] = valid ? count_of(p, mine.ncand) : 0u;
    if (p.policy == GF_SAMPLING_UNIFORM) nback[i] = my_back;
  }
}

// variant 1: one thread per target, scalar binary searches.
__global__ void __launch_bounds__(kSThreads) locate_thread_kernel(SampleParams p, const int64_t *__restrict__ nodes,
                                                                  const float *__restrict__ root_ts, uint64_t T_bound,
                                                                  const uint32_t *__restrict__ T_dev,
                                                                  TargetLoc *__restrict__ locs,
                                                                  uint32_t *__restrict__ counts,
                                                                  uint32_t *__restrict__ nback) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T_bound) return;
  const uint64_t T = T_dev ? (uint64_t)*T_dev : T_bound;
  TargetLoc mine = {0, 0, 0};
  uint32_t cnt = 0, my_back = 0;
  if (i < T) cnt = locate_thread(p, nodes[i], root_ts[i], mine, my_back);
  locs[i] = mine;
  counts[i] = cnt;
  if (p.policy == GF_SAMPLING_UNIFORM) nback[i] = my_back;
}

// -------------------------------------------------------------------------------------------- emit
struct EmitOut {
  int64_t *all_nodes;  // [T + S] (nullable: batched mode writes neighbours to nbr instead)
  float *all_ts;       // [T + S]
  int64_t *nbr;        // [S] batched mode
  float *nbr_ts;       // [S] batched mode
  float *dt;           // [S]
  int64_t *eid;        // [S]
  int64_t *row;        // [S]
  int64_t *col;        // [S] nullable
};

// A warp owns 32 consecutive targets (lane j holds target j's location, output offset and count) and walks their
// concatenated output slots 32 at a time: slot q belongs to the last lane whose relative offset is <= q, so all
// stores of one iteration hit 32 consecutive elements of every output array.
__device__ __forceinline__ void emit_warp(const SampleParams &p, const EmitOut &out, uint64_t T, bool valid,
                                          const TargetLoc &loc, uint32_t off, uint32_t cnt, uint32_t back, float root,
                                          uint64_t local_i, uint32_t batch, int lane) {
  const uint32_t base = __shfl_sync(0xffffffffu, off, 0);
  const uint32_t rel = valid ? off - base : 0xffffffffu;
  const uint32_t total = __reduce_add_sync(0xffffffffu, cnt);
  for (uint32_t q0 = 0; q0 < total; q0 += 32) {
    const uint32_t q = q0 + lane;
    int j = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      int cand = j + step;
      uint32_t r = __shfl_sync(0xffffffffu, rel, cand & 31);
      if (r <= q) j = cand;
    }
    const uint32_t rel_j = __shfl_sync(0xffffffffu, rel, j);
    const uint64_t desc_addr = __shfl_sync(0xffffffffu, loc.desc, j);
    uint32_t avail = __shfl_sync(0xffffffffu, loc.idx_hi, j);
    const uint32_t ncand = __shfl_sync(0xffffffffu, loc.ncand, j);
    const float root_j = __shfl_sync(0xffffffffu, root, j);
    const uint64_t li = __shfl_sync(0xffffffffu, local_i, j);
    const uint32_t batch_j = __shfl_sync(0xffffffffu, batch, j);
    const uint32_t back_j = __shfl_sync(0xffffffffu, back, j);
    if (q < total) {
      const uint32_t k = q - rel_j;
      uint32_t kk = k;  // distance (in edges) back from the newest in-window edge
      const BlockDesc *d = reinterpret_cast<const BlockDesc *>(desc_addr);
      BlockDesc blk = load_desc(d);
      if (p.policy == GF_SAMPLING_UNIFORM) {
        kk = philox_u32(p.seed, (uint32_t)(li * p.fanout + k), p.launch_index + batch_j) % ncand;
        if (kk >= avail) {
          // the drawn position lies in an older block; positions are cum_before + idx and the directory is
          // contiguous: smallest step back s in [1, back_j] with (d - s)->cum_before <= pos
          const uint32_t pos = blk.cum_before + avail - 1 - kk;
          uint32_t lo = 1, hi = back_j;
          while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&(d - mid)->cum_before) <= pos) hi = mid; else lo = mid + 1;
          }
          d -= lo;
          blk = load_desc(d);
          avail = pos - blk.cum_before + 1;
          kk = 0;
        }
      } else {
        while (kk >= avail) {  // recent: at most a few blocks back (sampling_kernels.cu:88-92)
          kk -= avail;
          d -= 1;
          blk = load_desc(d);
          avail = blk.size;
        }
      }
      const uint32_t idx = avail - 1 - kk;
      const EdgeRec rec = ld_rec(blk_rec(blk.payload, blk.capacity) + idx);
      const float t = rec.ts;
      const int64_t nb = (int64_t)rec.dst, ed = rec.eid;
      const uint64_t o = (uint64_t)base + q;
      const float ots = p.prop_time ? root_j : t;
      if (out.all_nodes) {
        out.all_nodes[T + o] = nb;
        out.all_ts[T + o] = ots;
      } else {
        out.nbr[o] = nb;
        out.nbr_ts[o] = ots;
      }
      out.dt[o] = __fsub_rn(root_j, t);
      out.eid[o] = ed;
      out.row[o] = (int64_t)li;
      if (out.col) out.col[o] = (int64_t)(T + o);
    }
  }
}

__global__ void __launch_bounds__(kSThreads) emit_kernel(SampleParams p, const int64_t *__restrict__ nodes,
                                                         const float *__restrict__ root_ts, uint64_t T_bound,
                                                         const uint32_t *__restrict__ T_dev,
                                                         const TargetLoc *__restrict__ locs,
                                                         const uint32_t *__restrict__ nback,
                                                         const uint32_t *__restrict__ offsets,  // [T_bound + 1]
                                                         const uint64_t *__restrict__ batch_offsets, uint32_t num_batches,
                                                         EmitOut out) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t T = T_dev ? (uint64_t)*T_dev : T_bound;
  const uint64_t i = warp * 32 + lane;
  if (warp * 32 >= T) return;
  const bool valid = i < T;
  TargetLoc loc = {0, 0, 0};
  uint32_t off = 0, cnt = 0, back = 0;
  float root = 0.f;
  uint64_t local_i = i;
  uint32_t batch = 0;
  if (valid) {
    loc = locs[i];
    if (p.policy == GF_SAMPLING_UNIFORM) back = nback[i];
    off = offsets[i];
    cnt = offsets[i + 1] - off;
    root = root_ts[i];
    if (out.all_nodes) {  // roots are the first T rows of the MFG source arrays (temporal_sampler.cu:242-243)
      out.all_nodes[i] = nodes[i];
      out.all_ts[i] = root;
    }
    if (batch_offsets) {
      batch = batch_of(batch_offsets, num_batches, i);
      local_i = i - batch_offsets[batch];
    }
  }
  emit_warp(p, out, T, valid, loc, off, cnt, back, root, local_i, batch, lane);
}

struct FusedMeta {
  uint32_t *meta_dev;   // {T, S, T + S} (device; feeds the next layer)
  uint32_t *meta_host;  // same, mapped pinned host memory (nullable)
  uint64_t *edge_offsets;  // batched mode: [num_batches + 1]
};

// variant 3 (default).  Persistent CTAs (one wave: #SMs x resident CTAs) pull tiles of 256 consecutive targets from an
// atomic ticket counter; the ticket of the NEXT tile is requested before the current tile is processed, so its
// latency is off the critical path.  Per tile:
//   locate : one thread per target -- window arithmetic, vertex entry, tail + first descriptor (two independent
//            loads), scalar searches -> block holding the newest in-window edge, idx_hi, #candidates, #emitted
//   scan   : block-wide exclusive scan of the emit counts; warp 0 publishes the tile aggregate and resolves the
//            tile's global output offset with a WARP-PARALLEL decoupled look-back (32 predecessor status words per
//            round trip; words are {generation, flag, value}, so no memset between launches)
//   emit   : per-target records and a slot -> owner map are staged in shared memory; then one thread per OUTPUT SLOT:
//            the 256 threads of the CTA write 256 consecutive elements of every output array per iteration
//            (full-line coalesced, streaming stores).
// Per-target state never touches HBM.
#ifndef GF_PTHREADS
// worker threads (= targets per tile) of the persistent kernel; + one control warp.  224 + 32 = 256 threads per CTA: four
// resident CTAs leave 64 registers per thread.  With 256 + 32 the budget is 56, which the round-2 locate (64-byte vertex
// entries, interpolation search, position bookkeeping) no longer fits without spilling: measured on one box, REDDIT headline
// launch 0.160 ms (256 workers, 36-92 B of spills) vs 0.1515 ms (224) vs 0.152 ms (192, 70 registers, no spills).
#define GF_PTHREADS 224
#endif
constexpr int kPThreads = GF_PTHREADS;
using OwnerT = uint8_t;   // slot -> owning target within the tile
constexpr uint32_t kMaxOwnerFanout = 128;  // slot -> owner map is kPThreads * fanout bytes of dynamic shared memory

struct PersistCtl {
  unsigned int *ticket;
  unsigned long long *status;  // [tiles]  (gen << 34) | (flag << 32) | value ; flag 1 = aggregate, 2 = inclusive prefix
  unsigned long long gen;
  unsigned long long *gstatus;  // [tiles / 32]  the same word for GROUPS of 32 consecutive tiles (two-level look-back)
};

struct LocatedT {   // what one thread keeps about its target between locate and emit
  uint64_t desc;    // BlockDesc holding the newest in-window edge (0 = none)
  uint64_t payload; // that block's payload
  uint32_t cap;     // ... and capacity
  uint32_t idx_hi;  // edges [0, idx_hi) of that block are older than the window end
  uint32_t ncand;   // in-window edges
  uint32_t back;    // live blocks older than that block
  uint32_t cum_d;   // position of that block's first edge
  uint32_t cum_f;   // position of the vertex's oldest live edge
};

__device__ __forceinline__ uint32_t lower_bound_ts_scalar(const float *ts, uint32_t n, float x) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(ts + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

#ifndef GF_DIR_BISECT
#define GF_DIR_BISECT 0  // experiment knob, see find_pos
#endif
// position (edges of this vertex older than x) + where it falls; `tail`/`head` are the newest / oldest live descriptors
struct Pos {
  const BlockDesc *d;
  BlockDesc blk;
  uint32_t idx;
};
// Directory search by time: the oldest live block whose end_ts >= x.  A vertex of the GDELT shape has hundreds to
// thousands of blocks; a binary search over their descriptors is that many DEPENDENT 32-byte loads per target.  Block
// time ranges are ordered and edges arrive roughly evenly in time, so two interpolation steps (each loads one whole
// descriptor: start_ts and end_ts decide three ways) usually land on the block; bisection takes over after that, which
// bounds the worst case at log2(#blocks) + 2.  Exact whatever the guesses were: the bracket [lo, hi] always contains the
// answer, and "start_ts < x <= end_ts" identifies it (the block before ends at or before start_ts < x).
__device__ __forceinline__ Pos find_pos(const BlockDesc *dir, uint32_t first, uint32_t end, const BlockDesc &tail, float x) {
  Pos r;
  r.d = dir + end - 1;
  r.blk = tail;
  if (tail.end_ts < x) {  // everything stored is older than x
    r.idx = tail.size;
    return r;
  }
#if GF_DIR_BISECT
  if (end - first > 1) {  // experiment knob: plain bisection over end_ts, one 4-byte load per probe
    uint32_t lo = 0, hi = end - first - 1;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (__ldg(&dir[first + mid].end_ts) < x) lo = mid + 1; else hi = mid;
    }
    if (first + lo != end - 1) {
      r.d = dir + first + lo;
      r.blk = load_desc(r.d);
    }
  }
#else
  if (end - first > 1 && !(x > tail.start_ts)) {  // x > tail.start_ts: the block before the newest ends before x
    uint32_t lo = 0, hi = end - first - 1;        // answer (relative to `first`) in [lo, hi]; hi = the newest block
    float t_lo = tail.min_ts, t_hi = tail.start_ts;  // times at the start of block lo / block hi
    uint32_t have = hi;                            // r.blk holds the descriptor of block `have`
    for (int it = 0; lo < hi; it++) {
      uint32_t g = (lo + hi) >> 1;
      if (it < 2 && t_hi > t_lo) {
        const float f = (x - t_lo) / (t_hi - t_lo);
        g = lo + (uint32_t)fminf(fmaxf(f, 0.f) * (float)(hi - lo), (float)(hi - lo - 1));
      }
      const BlockDesc b = load_desc(dir + first + g);
      if (x > b.end_ts) {
        lo = g + 1;
        t_lo = b.end_ts;
      } else {
        r.blk = b;
        have = g;
        if (x <= b.start_ts) {
          hi = g;
          t_hi = b.start_ts;
        } else {
          lo = hi = g;
        }
      }
    }
    if (have != lo) r.blk = load_desc(dir + first + lo);  // (only when lo stepped past the last loaded block onto hi's old value)
    r.d = dir + first + lo;
  }
#endif
  r.idx = x <= r.blk.start_ts ? 0u : blk_lower_bound(r.blk.payload, r.blk.capacity, r.blk.size, x);
  return r;
}

// the part of locate after the vertex entry and its newest descriptor have arrived (both searches + the counts)
template <int POLICY = -1>
__device__ __forceinline__ uint32_t locate_rest(const SampleParams &p, const NodeEntry &ent, const BlockDesc &tail, float root,
                                                LocatedT &loc) {
  float start, end;
  window_of(root, p, start, end);
  const BlockDesc *dir = reinterpret_cast<const BlockDesc *>(ent.dir());
  const Pos hi = find_pos(dir, ent.first, ent.end, tail, end);
  uint32_t pos_lo;
  if (start <= tail.min_ts) {
    // the window starts before the oldest stored edge (the newest descriptor carries that timestamp): no search and
    // no load -- the vertex entry carries the position of the oldest live edge
    pos_lo = ent.cum_first;
  } else {
    const Pos lo = find_pos(dir, ent.first, ent.end, tail, start);
    pos_lo = lo.blk.cum_before + lo.idx;
  }
  const uint32_t pos_hi = hi.blk.cum_before + hi.idx;
  loc.desc = (uint64_t)(uintptr_t)hi.d;
  loc.payload = hi.blk.payload;
  loc.cap = hi.blk.capacity;
  loc.idx_hi = hi.idx;
  loc.ncand = pos_hi > pos_lo ? pos_hi - pos_lo : 0u;
  loc.back = (uint32_t)(hi.d - (dir + ent.first));
  if (is_uniform<POLICY>(p)) {  // positions: only the uniform draw maps one back to a block
    loc.cum_d = hi.blk.cum_before;
    loc.cum_f = ent.cum_first;
  }
  return count_of<POLICY>(p, loc.ncand);
}

__device__ __forceinline__ uint32_t locate_target(const SampleParams &p, int64_t nid, float root, LocatedT &loc) {
  loc.desc = 0; loc.payload = 0; loc.cap = 0; loc.idx_hi = 0; loc.ncand = 0; loc.back = 0; loc.cum_d = 0; loc.cum_f = 0;
  if (nid < 0 || (uint64_t)nid >= p.table_len) return 0;  // oracle D3
  const NodeEntry ent = load_entry64(p.table + nid);
  if (ent.end <= ent.first) return 0;
  return locate_rest(p, ent, ent.tail, root, loc);
}

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

#ifndef GF_LOOKBACK_SLEEP
#define GF_LOOKBACK_SLEEP 300  // ns to back off after a poll that found a needed predecessor unpublished (+1-4 %)
#endif
// Two-level look-back (persistent kernel).  With ~600 tiles in flight and ~100 tiles retired per microsecond the
// nearest predecessor that already knows its inclusive prefix is hundreds of tiles back: the flat look-back paid a
// dozen dependent L2 round trips per tile and the workers waited for it (25 % of all stall samples in launches of
// short tiles, profiles/r01_s8_lookback.txt).  Here the last tile of every group of 32 also publishes a GROUP word
// (aggregate as soon as its 31 group predecessors have published theirs, inclusive prefix when it has resolved), so
// a tile sums <= 31 tile words of its own group and then walks 32 groups (1024 tiles) per round trip; the first
// loads of both levels are in flight together.
__device__ __forceinline__ uint32_t lookback_grouped(const PersistCtl &ctl, uint32_t tile, uint32_t total, int lane) {
  const uint32_t g = tile >> 5, r = tile & 31;  // r = predecessors inside the tile's own group
  const unsigned long long tag = ctl.gen << 34;
  auto flag_of = [](unsigned long long w) { return (uint32_t)(w >> 32) & 3u; };
  unsigned long long wg;
  {
    const int64_t q = (int64_t)g - 1 - lane;
    wg = q >= 0 ? ld_status(ctl.gstatus + q) : (2ull << 32);  // groups before the first: inclusive prefix 0
  }
  // ---- level 1: the tiles of the own group
  uint32_t sum_a = 0;
  bool closed = false;  // an inclusive prefix was found inside the group: nothing older is needed
  while (true) {
    unsigned long long w = 0;
    bool ready = true;
    if ((uint32_t)lane < r) {
      w = ld_status(ctl.status + tile - 1 - lane);
      ready = (w >> 34) == ctl.gen && flag_of(w) != 0;
    }
    const unsigned incl = __ballot_sync(0xffffffffu, (uint32_t)lane < r && ready && flag_of(w) == 2u);
    const unsigned nready = __ballot_sync(0xffffffffu, !ready);
    const int stop = incl ? __ffs(incl) - 1 : 31;
    const unsigned need = stop == 31 ? 0xffffffffu : ((2u << stop) - 1u);
    if (nready & need) {
#if GF_LOOKBACK_SLEEP
      __nanosleep(GF_LOOKBACK_SLEEP);
#endif
      continue;
    }
    sum_a = __reduce_add_sync(0xffffffffu, ((uint32_t)lane < r && lane <= stop) ? (uint32_t)w : 0u);
    closed = incl != 0;
    break;
  }
  if (r == 31 && lane == 0) st_status(ctl.gstatus + g, tag | ((closed ? 2ull : 1ull) << 32) | (sum_a + total));
  if (closed) return sum_a;
  // ---- level 2: whole groups, nearest first
  uint32_t part = 0;
  int64_t q0 = (int64_t)g - 1;
  bool have = true;  // wg holds the words of the current window
  while (true) {
    const int64_t q = q0 - lane;
    if (!have) wg = q >= 0 ? ld_status(ctl.gstatus + q) : (2ull << 32);
    have = false;
    const bool ready = q < 0 || ((wg >> 34) == ctl.gen && flag_of(wg) != 0);
    const unsigned incl = __ballot_sync(0xffffffffu, ready && flag_of(wg) == 2u);
    const unsigned nready = __ballot_sync(0xffffffffu, !ready);
    const int stop = incl ? __ffs(incl) - 1 : 31;
    const unsigned need = stop == 31 ? 0xffffffffu : ((2u << stop) - 1u);
    if (nready & need) {
#if GF_LOOKBACK_SLEEP
      __nanosleep(GF_LOOKBACK_SLEEP);
#endif
      continue;
    }
    if (lane <= stop) part += (uint32_t)wg;
    if (incl) break;
    q0 -= 32;
  }
  const uint32_t excl = sum_a + __reduce_add_sync(0xffffffffu, part);
  if (r == 31 && lane == 0) st_status(ctl.gstatus + g, tag | (2ull << 32) | (excl + total));
  return excl;
}

// where one output slot reads from: (block payload, capacity, element index); `k` is the slot of the target, `li`
// the target's index for the counter-based RNG (sampling_kernels.cu:88-105 / :202-270)
struct Slot {
  uint64_t payload;
  uint32_t cap, idx, li;
  float root;
};
template <int POLICY = -1>
__device__ __forceinline__ Slot resolve_slot(const SampleParams &p, uint64_t payload, uint32_t cap, uint32_t idx_hi,
                                             uint32_t ncand, uint32_t back, uint64_t desc, uint32_t cum_d, uint32_t cum_f,
                                             uint32_t li, uint32_t k, uint32_t batch, float root) {
  Slot r;
  r.payload = payload;
  r.cap = cap;
  r.root = root;
  r.li = li;
  uint32_t avail = idx_hi;
  uint32_t kk = k;  // distance (in edges) back from the newest in-window edge
  if (is_uniform<POLICY>(p))
    kk = philox_u32(p.seed, (uint32_t)((uint64_t)li * p.fanout + k), p.launch_index + batch) % ncand;
  if (kk >= avail) {  // the edge lies in an older block
    const BlockDesc *d = reinterpret_cast<const BlockDesc *>(desc);
    BlockDesc blk;
    if (is_uniform<POLICY>(p)) {
      // The drawn edge is at position pos (positions are cum_before + idx and contiguous over the directory) in one
      // of the `back` older live blocks.  Blocks of one vertex are mostly of one size (the adaptive block policy), so
      // interpolating over the positions usually hits the block with the first descriptor load (a descriptor gives
      // both ends of its position range); bisection after two misses bounds the worst case.  The reference walks the
      // whole list for every draw (sampling_kernels.cu:207-270).
      const uint32_t pos = cum_d + avail - 1 - kk;
      const BlockDesc *d0 = d - back;            // oldest live block
      uint32_t lo = 0, hi = back - 1;            // block index relative to d0, answer in [lo, hi]
      uint32_t c_lo = cum_f, c_hi = cum_d;       // positions [c_lo, c_hi) are what blocks lo .. hi hold
      for (int it = 0;; it++) {
        uint32_t g = (lo + hi) >> 1;
        if (it < 2)
          g = lo + (uint32_t)min((uint64_t)(hi - lo), (uint64_t)(pos - c_lo) * (hi - lo + 1) / (uint64_t)(c_hi - c_lo));
        blk = load_desc(d0 + g);
        if (pos < blk.cum_before) {
          hi = g - 1;
          c_hi = blk.cum_before;
        } else if (pos >= blk.cum_before + blk.size) {
          lo = g + 1;
          c_lo = blk.cum_before + blk.size;
        } else {
          break;
        }
      }
      avail = pos - blk.cum_before + 1;
      kk = 0;
    } else {
      uint32_t left = back;  // live blocks older than d: never walk past the oldest one
      do {  // recent: at most a few blocks back (sampling_kernels.cu:88-92)
        if (left-- == 0) { kk = avail - 1; break; }  // unreachable when ncand is consistent with the directory
        kk -= avail;
        d -= 1;
        blk = load_desc(d);
        avail = blk.size;
      } while (kk >= avail);
    }
    r.payload = blk.payload;
    r.cap = blk.capacity;
  }
  r.idx = avail - 1 - kk;
  return r;
}

// named barriers (PTX barrier.sync / barrier.arrive): the control warp and the 8 worker warps hand tiles to each other
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
  __threadfence_block();
  asm volatile("barrier.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#ifndef GF_PERSIST_STAGES
#define GF_PERSIST_STAGES 2  // tiles in flight per CTA between locate and emit (2 or 3; 3 helps only launches of short tiles and costs 5 % elsewhere)
#endif
constexpr int kStages = GF_PERSIST_STAGES;
enum : int { kBarTile = 1, kBarCounts = 4, kBarBase = 7, kBarBatch = 10, kBarWorkers = 13 };  // + pipeline stage (0 .. 2)
constexpr int kPAll = kPThreads + 32;  // 8 worker warps + the control warp
constexpr uint32_t kNoTile = 0xffffffffu;

// exclusive scan over the 256 worker threads (named barrier: the control warp does not take part)
__device__ __forceinline__ uint32_t worker_excl_scan(uint32_t v, uint32_t *warp_sums, uint32_t *total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t incl = warp_incl_scan(v, lane);
  if (lane == 31) warp_sums[w] = incl;
  bar_sync(kBarWorkers, kPThreads);
  uint32_t s = lane < kPThreads / 32 ? warp_sums[lane] : 0;
  const uint32_t si = warp_incl_scan(s, lane);  // every warp redundantly scans the 8 warp sums
  const uint32_t before = __shfl_sync(0xffffffffu, si - s, w);
  if (threadIdx.x == kPThreads - 1) *total = incl + before;
  return incl - v + before;  // warp_sums / total belong to one pipeline stage: not reused before the tile after next
}

struct TileStage {  // per-target records of one tile in flight between locate and emit (shared memory)
  static constexpr int N = kPThreads;
  uint64_t desc[N], payload[N];
  uint32_t cap[N], idx_hi[N], ncand[N], back[N], cum_d[N], cum_f[N], loff[N], li[N], batch[N];
  uint32_t pstart[N];  // compacted launches: first batch whose edge offset this target reports (> batch: none)
  float root[N];
  uint32_t warp_sums[kPThreads / 32];
  uint32_t tile, total, base, batch0;
};

#ifndef GF_PERSIST_OCC
#define GF_PERSIST_OCC 4  // resident CTAs per SM the register budget is sized for (build-time experiment knob)
#endif
#ifndef GF_ENTRY_COND
// 1: the second sector of a vertex entry (the copy of its newest block descriptor) is requested only once the first
// says the vertex has blocks.  Measured (profiles/r02_c4_sampler_variants.json): REDDIT headline launch 0.154 vs
// 0.161 ms with both sectors requested up front (36 % of its targets have no out-edges), GDELT shapes equal.
#define GF_ENTRY_COND 1
#endif
#ifndef GF_ANNOUNCE_LATE
#define GF_ANNOUNCE_LATE 0  // 1 = control warp resolves the current tile before the batch lookup of the next one (measured: 0-5 % slower)
#endif
#ifndef GF_CTL_PREFETCH
#define GF_CTL_PREFETCH 1  // control warp prefetches the next tile's roots into L2: +1.5-3 % (profiles/r01_s8_experiments.json)
#endif

// TT_: targets per tile.  kPThreads (one per worker thread) for launches that fill the GPU.  kSmallTile for launches of a
// few thousand targets -- the per-batch calls of a training loop: a quarter of the worker threads locate, ALL of them
// emit, so the batch spreads over four times as many SMs and a thread resolves one or two output slots instead of five
// (the uniform policy's per-slot directory search is a chain of dependent loads that nothing else hides at one tile per SM).
#ifndef GF_SMALL_TILE
#define GF_SMALL_TILE 56  // build-time experiment knob
#endif
constexpr int kSmallTile = GF_SMALL_TILE;
template <bool LIST, int OCC, int POLICY, int TT_ = kPThreads>
__global__ void __launch_bounds__(kPAll, OCC)
    sample_persistent_kernel(SampleParams p, const int64_t *__restrict__ nodes, const float *__restrict__ root_ts,
                             uint64_t T_bound, const uint32_t *__restrict__ T_dev,
                             const uint64_t *__restrict__ batch_offsets, uint32_t num_batches, EmitOut out,
                             PersistCtl ctl, FusedMeta meta, const uint32_t *__restrict__ active_arg,
                             const uint32_t *__restrict__ A_dev) {
  const uint32_t *const active = LIST ? active_arg : nullptr;  // LIST = false: the list code folds away
  // `active` (optional, multi-batch launches): ascending indices of the targets that can have neighbours at all (their
  // vertex has out-edges); the tiles then run over this compacted list.  Targets without edges emit nothing, so the
  // output is the same as without the list; they just no longer take a slot of the ordered tile pipeline.
  using Stage = TileStage;
  using Owner = OwnerT;
  constexpr uint32_t TT = TT_;  // targets per tile (worker threads tid < TT locate one each)
  extern __shared__ __align__(16) uint8_t s_dyn[];  // kStages x Stage, then kStages x slot -> owner map [TT * fanout]
  Stage *stages = reinterpret_cast<Stage *>(s_dyn);
  Owner *owners = reinterpret_cast<Owner *>(s_dyn + kStages * sizeof(Stage));
  const int tid = threadIdx.x, lane = tid & 31;
  const uint64_t T = T_dev ? (uint64_t)*T_dev : T_bound;
  const uint32_t N = active ? *A_dev : (uint32_t)T;  // entries the tiles run over (a launch has < 2^32 targets)
  const uint32_t ntiles = (uint32_t)(((uint64_t)N + TT - 1) / TT);

  // Roles.  The LEADER (last worker thread, the one that ends up holding the tile's total) publishes the tile
  // AGGREGATE the moment the scan has produced it; the CONTROL warp draws the tickets (the latency of that contended
  // atomic must stay off the workers' path: drawn by the leader it cost 10-50 %, profiles/r01_s8_experiments.json),
  // looks up the first batch of the next tile and resolves: look-back, inclusive prefix, output offset.  (Until session
  // 8 the control warp also published the aggregate, i.e. only after it had finished the look-back of the CTA's
  // PREVIOUS tile: look-backs waited on aggregates that waited on look-backs.)
  const bool leader = tid == kPThreads - 1;

  if (tid >= kPThreads) {
    // ================================================================= control warp: batches + look-back
    // batch of the tile's first target (largest b with batch_offsets[b] <= i): 32 probes per round trip, done by the
    // control warp so that no worker's locate is delayed by it
    auto batch0_of = [&](uint32_t t) -> uint32_t {
      const uint64_t i = active ? (uint64_t)active[(uint64_t)t * TT] : (uint64_t)t * TT;
      uint32_t lo = 0, hi = num_batches;  // answer in [lo, hi)
      while (hi - lo > 1) {
        const uint32_t len = hi - lo, step = (len + 32) / 33;
        const uint32_t idx = lo + (lane + 1) * step;
        const bool le = idx < hi && batch_offsets[idx] <= i;
        const uint32_t c = __popc(__ballot_sync(0xffffffffu, le));  // probes are monotone: c leading trues
        const uint32_t nlo = lo + c * step;
        hi = min(hi, nlo + step);
        lo = nlo;
      }
      return lo;
    };
    // the roots of a tile the workers are about to reach: pull their lines into L2 now
    auto prefetch_roots = [&](uint32_t t) {
#if GF_CTL_PREFETCH
      if (active) return;  // the roots of a compacted tile are not contiguous
      const uint64_t i0 = (uint64_t)t * TT;
      const uint64_t n = min((uint64_t)TT, T - i0);
      for (uint64_t k = (uint64_t)lane * 16; k < n; k += 32 * 16)  // 128-byte lines of the node ids
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nodes + i0 + k));
      for (uint64_t k = (uint64_t)lane * 32; k < n; k += 32 * 32)  // ... and of the timestamps
        asm volatile("prefetch.global.L2 [%0];" ::"l"(root_ts + i0 + k));
#endif
    };
    auto announce = [&](uint32_t t, int sg) {  // what the workers need for tile t beyond its id (after their searches)
      if (t == kNoTile) return;
      if (batch_offsets) {
        const uint32_t b0 = batch0_of(t);
        if (lane == 0) stages[sg].batch0 = b0;
        bar_arrive(kBarBatch + sg, kPAll);
      }
    };
    bool drained = false;  // this CTA has drawn its end-of-work ticket: it draws no more (exactly one per CTA)
    auto draw = [&]() -> uint32_t {
      if (drained) return kNoTile;
      uint32_t t = 0;
      if (lane == 0) {
        t = atomicAdd(ctl.ticket, 1u);
        if (t == ntiles + gridDim.x - 1) *ctl.ticket = 0;  // the last ticket of this launch: re-arm for the next one
        if (ntiles == 0 && t == 0) {                         // empty launch: nobody else reports the totals
          meta.meta_dev[0] = meta.meta_dev[2] = (uint32_t)T;
          meta.meta_dev[1] = 0;
          if (meta.meta_host) {
            meta.meta_host[0] = meta.meta_host[2] = (uint32_t)T;
            meta.meta_host[1] = 0;
            __threadfence_system();
            *(volatile uint32_t *)(meta.meta_host + 3) = 1u;  // the word the host polls (wait_flag_or_sync)
          }
          if (meta.edge_offsets)
            for (uint32_t b = 0; b <= num_batches; b++) meta.edge_offsets[b] = 0;
        }
      }
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= ntiles) drained = true;
      return t < ntiles ? t : kNoTile;
    };
    uint32_t tile = draw();
    if (lane == 0) stages[0].tile = tile;
    bar_arrive(kBarTile + 0, kPAll);
    if (tile != kNoTile) prefetch_roots(tile);
    announce(tile, 0);
    // NOTE (measured, profiles/r01_s6_runahead_experiment.json): drawing the ticket one tile EARLIER (so that the
    // hand-over never waits for the atomic) is 14 % slower -- a tile is then claimed ~2 tile times before its
    // aggregate is published and every later tile's emit waits on the slowest such claim.  Claim late.
    int st = 0;
    uint32_t it = 0;
    for (; tile != kNoTile; it++) {
      const int sn = st + 1 == kStages ? 0 : st + 1;
      bar_sync(kBarCounts + st, kPAll);  // the workers have staged tile `tile`
      const uint32_t total = stages[st].total;
      // hand the workers their next tile before resolving this one: the look-back overlaps their work
      const uint32_t next = draw();
      if (lane == 0) stages[sn].tile = next;
      bar_arrive(kBarTile + sn, kPAll);
      if (next != kNoTile) prefetch_roots(next);
#if !GF_ANNOUNCE_LATE
      announce(next, sn);
#endif
      uint32_t excl = 0;
      if (tile != 0) {
        excl = lookback_grouped(ctl, tile, total, lane);
        if (lane == 0) st_status(ctl.status + tile, (ctl.gen << 34) | (2ull << 32) | (excl + total));
      }
      if (lane == 0) {
        stages[st].base = excl;
        if (tile == ntiles - 1) {  // the last tile's inclusive prefix is the number of sampled neighbours
          const uint32_t S = excl + total;
          meta.meta_dev[0] = (uint32_t)T;
          meta.meta_dev[1] = S;
          meta.meta_dev[2] = (uint32_t)T + S;
          if (meta.meta_host) {
            // the sizes are known as soon as the last tile's prefix is: a host that only needs them (device-resident
            // results) builds its views while the tiles still emit
            meta.meta_host[0] = (uint32_t)T;
            meta.meta_host[1] = S;
            meta.meta_host[2] = (uint32_t)T + S;
            __threadfence_system();
            *(volatile uint32_t *)(meta.meta_host + 3) = 1u;  // the word the host polls (wait_flag_or_sync)
          }
          if (meta.edge_offsets) {
            meta.edge_offsets[num_batches] = S;
            if (active) {  // the batches after the one of the last listed target hold no edges
              const uint64_t last = active[N - 1];
              uint32_t lo = 0, hi = num_batches;
              while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (batch_offsets[mid] <= last) lo = mid; else hi = mid;
              }
              for (uint32_t b = lo + 1; b < num_batches; b++) meta.edge_offsets[b] = S;
            }
          }
        }
      }
      bar_arrive(kBarBase + st, kPAll);
#if GF_ANNOUNCE_LATE
      announce(next, sn);  // the workers ask for it a whole locate phase after the hand-over: resolve this tile first
#endif
      tile = next;
      st = sn;
    }
    // the workers run kStages - 1 iterations past their last tile to drain the pipeline: the first hand-over of "no
    // tile" was made above, the others here
    if (it > 0) {
      for (int k = 1; k < kStages - 1; k++) {
        const int sg = (st + k) % kStages;
        if (lane == 0) stages[sg].tile = kNoTile;
        bar_arrive(kBarTile + sg, kPAll);
      }
    }
    return;
  }

  // ============================================= worker warps: locate(it), emit(it - (kStages - 1))
  // With three stages the look-back of a tile has two locate phases to finish before its emit asks for the result.
  uint32_t hist[kStages - 1];  // hist[k] = tile of iteration it - 1 - k
#pragma unroll
  for (int k = 0; k < kStages - 1; k++) hist[k] = kNoTile;
  bar_sync(kBarTile + 0, kPAll);
  for (int st = 0;; st = st + 1 == kStages ? 0 : st + 1) {
    Stage &S = stages[st];
    const int sn = st + 1 == kStages ? 0 : st + 1;
    const uint32_t tile = S.tile;
    if (tile != kNoTile) {
      const uint32_t c = tile * TT + tid;  // this thread's entry
      const bool live = (TT == kPThreads || tid < (int)TT) && c < N;
      const uint32_t oi = live ? (active ? active[c] : c) : 0u;  // target index (the entry itself without a list)
      // ---- front of the chain: root, vertex entry, newest descriptor
      int64_t nid = -1;
      float root = 0.f;
      if (live) {
        nid = __ldcs(nodes + oi);
        root = __ldcs(root_ts + oi);
      }
      // the vertex entry carries a copy of its newest block descriptor: one dependent load, two sectors of one line
      NodeEntry ent;
      ent.dir_tagged = 0; ent.first = 0; ent.end = 0;
#if GF_ENTRY_COND
      if (nid >= 0 && (uint64_t)nid < p.table_len) {
        const U8x32 q = ldg256_b32(p.table + nid);
        ent.dir_tagged = ((uint64_t)q.w[1] << 32) | q.w[0];
        ent.first = q.w[2]; ent.end = q.w[3]; ent.cum_first = q.w[4];
        if (ent.end > ent.first) ent.tail = load_desc(&(p.table + nid)->tail);
      }
#else
      if (nid >= 0 && (uint64_t)nid < p.table_len) ent = load_entry64(p.table + nid);  // oracle D3
#endif
      // ---- the searches
      LocatedT loc;
      loc.desc = 0; loc.payload = 0; loc.cap = 0; loc.idx_hi = 0; loc.ncand = 0; loc.back = 0; loc.cum_d = 0; loc.cum_f = 0;
      const uint32_t cnt = ent.end > ent.first ? locate_rest<POLICY>(p, ent, ent.tail, root, loc) : 0u;
      if (out.all_nodes && live) {  // roots are the first T rows of the MFG source arrays (temporal_sampler.cu:242-243)
        out.all_nodes[oi] = nid;
        out.all_ts[oi] = root;
      }
      const uint32_t loff = worker_excl_scan(cnt, S.warp_sums, &S.total);
      if (leader)  // this thread's inclusive prefix is the tile's total: publish the aggregate right away
        st_status(ctl.status + tile, (ctl.gen << 34) | ((tile == 0 ? 2ull : 1ull) << 32) | (loff + cnt));
      Owner *own = owners + (size_t)st * TT * p.fanout;
      if (batch_offsets) bar_sync(kBarBatch + st, kPAll);  // the control warp has looked up the tile's first batch
      {
        const uint64_t i = oi;
        const uint32_t j = tid;
        uint64_t local_i = i;
        uint32_t b = 0, pstart = 0xffffffffu;
        if (batch_offsets && live) {
          b = S.batch0;
          while (b + 1 < num_batches && i >= batch_offsets[b + 1]) b++;
          local_i = i - batch_offsets[b];
          if (active) {  // does this target open its batch in the list?  then it reports the offsets of the batches
            pstart = 0;  // between the previous listed target's batch (exclusive) and its own (inclusive)
            if (c > 0) {
              const uint64_t prev = active[c - 1];
              uint32_t pb = b;
              while (pb > 0 && batch_offsets[pb] > prev) pb--;
              pstart = pb + 1;
            }
          }
        }
        S.desc[j] = loc.desc;
        S.payload[j] = loc.payload;
        S.cap[j] = loc.cap;
        S.idx_hi[j] = loc.idx_hi;
        S.ncand[j] = loc.ncand;
        S.back[j] = loc.back;
        if (POLICY == GF_SAMPLING_UNIFORM) {
          S.cum_d[j] = loc.cum_d;
          S.cum_f[j] = loc.cum_f;
        }
        S.loff[j] = loff;
        S.li[j] = (uint32_t)local_i;
        S.batch[j] = b;
        S.pstart[j] = pstart;
        S.root[j] = root;
        for (uint32_t k = 0; k < cnt; k++) own[loff + k] = (Owner)j;
      }
      bar_arrive(kBarCounts + st, kPAll);
    }
    const uint32_t prev_tile = hist[kStages - 2];
    if (prev_tile != kNoTile) {
      // ---- emit tile it - (kStages - 1): one thread per output slot
      const int ps = st + 1 == kStages ? 0 : st + 1;
      const Stage &P = stages[ps];
      const Owner *own = owners + (size_t)ps * TT * p.fanout;
      // the control warp has resolved the tile's output offset; it did so after every worker had staged its
      // record (kBarCounts), so the records are visible too
      bar_sync(kBarBase + ps, kPAll);
      const uint32_t total = P.total;
      const uint64_t base = P.base;
      if (meta.edge_offsets) {
        const uint32_t j = tid;
        const uint64_t i = (TT == kPThreads || j < TT) ? (uint64_t)prev_tile * TT + j : ~0ull;  // threads without a target
        if (active) {
          if (i < N)
            for (uint32_t b = P.pstart[j]; b <= P.batch[j]; b++) meta.edge_offsets[b] = base + P.loff[j];
        } else if (i < T && P.li[j] == 0) {
          const uint32_t b0 = P.batch[j];
          meta.edge_offsets[b0] = base + P.loff[j];
          for (uint32_t b = b0; b > 0 && batch_offsets[b - 1] == i; --b) meta.edge_offsets[b - 1] = base + P.loff[j];  // empty batches
        }
      }
      auto resolve = [&](uint32_t q) -> Slot {
        const uint32_t j = own[q];
        constexpr bool uni = POLICY == GF_SAMPLING_UNIFORM;  // only the uniform policy looks at positions
        return resolve_slot<POLICY>(p, P.payload[j], P.cap[j], P.idx_hi[j], P.ncand[j], P.back[j], P.desc[j],
                                    uni ? P.cum_d[j] : 0u, uni ? P.cum_f[j] : 0u, P.li[j], q - P.loff[j], P.batch[j], P.root[j]);
      };
      auto store = [&](uint32_t q, const Slot &r, float t, int64_t nb, int64_t ed) {
        const uint64_t o = base + q;
        const float ots = p.prop_time ? r.root : t;
        if (out.all_nodes) {  // re-read by the next layer: default caching
          out.all_nodes[T + o] = nb;
          out.all_ts[T + o] = ots;
        } else {
          __stcs(out.nbr + o, nb);
          __stcs(out.nbr_ts + o, ots);
        }
        __stcs(out.dt + o, __fsub_rn(r.root, t));
        __stcs(out.eid + o, ed);
        __stcs(out.row + o, (int64_t)r.li);
        if (out.col) __stcs(out.col + o, (int64_t)(T + o));
      };
      // two slots per thread and iteration: two independent 128-bit gathers in flight before the first store
      for (uint32_t q = tid; q < total; q += 2 * kPThreads) {
        const uint32_t q2 = q + kPThreads;
        const bool two = q2 < total;
        const Slot a = resolve(q);
        const Slot b = two ? resolve(q2) : a;
        const EdgeRec ra = ld_rec(blk_rec(a.payload, a.cap) + a.idx);
        const EdgeRec rb = ld_rec(blk_rec(b.payload, b.cap) + b.idx);
        store(q, a, ra.ts, (int64_t)ra.dst, ra.eid);
        if (two) store(q2, b, rb.ts, (int64_t)rb.dst, rb.eid);
      }
    }
    bool idle = tile == kNoTile;
#pragma unroll
    for (int k = kStages - 2; k > 0; k--) {
      hist[k] = hist[k - 1];
      idle = idle && hist[k] == kNoTile;
    }
    hist[0] = tile;
    if (idle) return;  // nothing left in flight
    bar_sync(kBarTile + sn, kPAll);  // the control warp has handed over the next iteration's tile
  }
}

// chaining: meta[0] = T, meta[1] = S of the step just finished; next step's T = T + S
__global__ void chain_meta_kernel(const uint32_t *T_dev, uint64_t T_host, const uint32_t *S_dev, uint32_t *meta_out,
                                  uint32_t *meta_host) {
  uint32_t T = T_dev ? *T_dev : (uint32_t)T_host;
  meta_out[0] = T;
  meta_out[1] = *S_dev;
  meta_out[2] = T + *S_dev;
  if (meta_host) {
    meta_host[0] = T;
    meta_host[1] = *S_dev;
    meta_host[2] = T + *S_dev;
    __threadfence_system();
    *(volatile uint32_t *)(meta_host + 3) = 1u;
  }
}

__global__ void gather_edge_offsets_kernel(const uint32_t *__restrict__ offsets, const uint64_t *__restrict__ batch_offsets,
                                           uint32_t num_batches, uint64_t *edge_offsets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= num_batches) edge_offsets[b] = offsets[batch_offsets[b]];
}

// Targets of the next layer of a multi-batch launch: for every batch b its roots followed by the neighbours the layer
// just sampled for it (what TemporalSampler::Sample chains for one batch, temporal_sampler.cu:242-262, 279-305).
// Element e of the output: batch b = the batch of e in the OUTPUT offsets (batch_offsets[b] + edge_offsets[b]); the
// first size(b) elements of the batch are its roots, the rest its neighbours.  One warp per 32 consecutive output
// elements: lane 0 finds the batch of the first one, the lanes walk forward from there.
__global__ void __launch_bounds__(256) chain_batched_kernel(const int64_t *__restrict__ nodes, const float *__restrict__ ts,
                                                            uint64_t T, const uint64_t *__restrict__ batch_offsets,
                                                            uint32_t num_batches, const int64_t *__restrict__ nbr,
                                                            const float *__restrict__ nbr_ts,
                                                            const uint64_t *__restrict__ edge_offsets, int64_t *nodes_out,
                                                            float *ts_out, uint64_t *batch_offsets_out) {
  const uint64_t total = T + edge_offsets[num_batches];
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= num_batches; b += stride)
    batch_offsets_out[b] = batch_offsets[b] + edge_offsets[b];
  // every warp streams one contiguous range of the output: the batch of its first element costs one binary search, after
  // that the warp walks forward (a search per 32 elements made the pass latency-bound: 0.33 of HBM)
  const uint64_t nwarps = stride >> 5, warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  constexpr int U = 4;  // elements per lane and iteration: 2 * U independent loads in flight before the first store
  const uint64_t chunk = ((total + nwarps - 1) / nwarps + 32 * U - 1) / (32 * U) * (32 * U);
  const uint64_t begin = warp * chunk, end = min(total, begin + chunk);
  if (begin >= end) return;
  uint32_t b = 0;
  {
    uint32_t hi = num_batches;  // largest b with out_offset(b) <= begin (same answer in every lane: broadcast loads)
    while (hi - b > 1) {
      const uint32_t mid = (b + hi) >> 1;
      if (batch_offsets[mid] + edge_offsets[mid] <= begin) b = mid; else hi = mid;
    }
  }
  uint64_t r0 = batch_offsets[b], q0 = edge_offsets[b], r1 = batch_offsets[b + 1], q1 = edge_offsets[b + 1];
  for (uint64_t e0 = begin; e0 < end; e0 += 32 * U) {
    while (e0 >= r1 + q1) {  // warp-uniform: the warp's first element has left batch b
      b++;
      r0 = r1; q0 = q1;
      r1 = batch_offsets[b + 1]; q1 = edge_offsets[b + 1];
    }
    const int64_t *sn[U];
    const float *st[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t e = e0 + u * 32 + lane;
      sn[u] = nullptr;
      st[u] = nullptr;
      if (e >= end) continue;
      uint64_t lr0 = r0, lq0 = q0, lr1 = r1;
      if (e >= r1 + q1) {  // already in a later batch (rare: batches are thousands of elements)
        uint32_t lb = b + 1;
        while (lb + 1 < num_batches && batch_offsets[lb + 1] + edge_offsets[lb + 1] <= e) lb++;
        lr0 = batch_offsets[lb]; lq0 = edge_offsets[lb];
        lr1 = batch_offsets[lb + 1];
      }
      const uint64_t k = e - (lr0 + lq0), nroots = lr1 - lr0;
      sn[u] = k < nroots ? nodes + lr0 + k : nbr + lq0 + (k - nroots);
      st[u] = k < nroots ? ts + lr0 + k : nbr_ts + lq0 + (k - nroots);
    }
    int64_t vn[U];
    float vt[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      vn[u] = sn[u] ? __ldcs(sn[u]) : 0;
      vt[u] = st[u] ? __ldcs(st[u]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t e = e0 + u * 32 + lane;
      if (sn[u]) {
        nodes_out[e] = vn[u];
        ts_out[e] = vt[u];
      }
    }
  }
}

// List of the targets whose vertex has out-edges at all.  The flag of target i comes from the store's is_src bytes (a
// superset after offloading: such targets are then simply not skipped).  One launch: a CTA takes a chunk of 16 384
// consecutive targets from a ticket, every thread flags its 64 consecutive targets into a 64-bit mask, the counts
// are scanned in the block, the chunk's offset comes from a decoupled look-back over the chunks (a handful per
// launch: the 1024-element tiles of the generic scan cost 45 us on 2 M targets, all of it look-back latency), and
// every thread writes the indices of its set bits in order -- the list is ascending.
constexpr int kActPer = 64, kActChunk = kScanThreads * kActPer;
__global__ void __launch_bounds__(kScanThreads) active_list_kernel(const int64_t *__restrict__ nodes, uint64_t T,
                                                                   const uint8_t *__restrict__ is_src, uint64_t table_len,
                                                                   uint32_t *__restrict__ list, LookbackCtl ctl,
                                                                   uint32_t *count_out) {
  __shared__ uint32_t s_chunk, s_base, s_total;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ctl.ticket, 1u);
    if (t == gridDim.x - 1) *ctl.ticket = 0;  // every chunk of this launch has its ticket: re-arm
    s_chunk = t;
  }
  __syncthreads();
  const uint32_t chunk = s_chunk;
  const uint64_t first = (uint64_t)chunk * kActChunk + (uint64_t)threadIdx.x * kActPer;
  unsigned long long mask = 0;
  auto flag = [&](int64_t v) -> bool { return v >= 0 && (uint64_t)v < table_len && __ldg(is_src + v) != 0; };
  if (first + kActPer <= T && ((uintptr_t)nodes & 15) == 0) {  // 64 whole targets, 16-byte aligned: two ids per load
    const longlong2 *q = reinterpret_cast<const longlong2 *>(nodes + first);
#pragma unroll 8
    for (int k = 0; k < kActPer / 2; k++) {
      const longlong2 v = __ldg(q + k);
      if (flag(v.x)) mask |= 1ull << (2 * k);
      if (flag(v.y)) mask |= 1ull << (2 * k + 1);
    }
  } else {
    for (int k = 0; k < kActPer; k++)
      if (first + k < T && flag(__ldg(nodes + first + k))) mask |= 1ull << k;
  }
  uint32_t off = block_excl_scan((uint32_t)__popcll(mask), &s_total);
  if (threadIdx.x < 32) {
    const uint32_t total = s_total;
    const unsigned long long tag = ctl.gen << 34;
    uint32_t excl = 0;
    if (chunk == 0) {
      if (lane == 0) lb_store(ctl.status, tag | (2ull << 32) | total);
    } else {
      if (lane == 0) lb_store(ctl.status + chunk, tag | (1ull << 32) | total);
      excl = lb_lookback_warp(ctl, chunk, lane);
      if (lane == 0) lb_store(ctl.status + chunk, tag | (2ull << 32) | (excl + total));
    }
    if (lane == 0) {
      s_base = excl;
      if (chunk == gridDim.x - 1) *count_out = excl + total;
    }
  }
  __syncthreads();
  off += s_base;
  while (mask) {
    const int k = __ffsll((long long)mask) - 1;
    mask &= mask - 1;
    list[off++] = (uint32_t)(first + k);
  }
}

}  // namespace gf

using namespace gf;

struct gf_sampler {
  gf_graph *graph;
  std::vector<uint32_t> fanouts;
  int policy;
  uint32_t num_snapshots;
  float window;
  int prop_time;
  uint64_t seed;
  uint64_t launch_index = 0;
  int variant = 3;
  int host_out_mode = 0;         // 0: auto; 1: always device mirror + D2H copies; 2: pinned host outputs written in place
  gf::Scratch compact;          // [ticket, count | scan status words | active list] of the compacted multi-batch launches
  long long compact_min = -1;   // launches with at least this many targets are compacted (-1: not read yet; 0: never)
  std::vector<std::pair<uint64_t, unsigned>> persist_cfg;  // (instantiation << 32 | fan-out) -> #SMs x resident CTAs
  size_t persist_dyn[64] = {};  // per instantiation: dynamic shared memory limit set so far
  int occ = 0;                  // launch-bound instantiation of the persistent kernel (0: not chosen yet)
  Scratch ws;      // 3-kernel pipeline: locs | counts | offsets | scan tmp
  Scratch in;      // staged host input (device)
  Scratch outbuf;  // device copy of host-bound outputs
  Scratch meta;    // per-step {T, S, T + S, scratch}
  Scratch fused;   // ticket + tile status words
  size_t fused_tiles = 0;
  unsigned long long fused_gen = 0;
  uint32_t *h_meta = nullptr;  // pinned + mapped
  size_t h_meta_cap = 0;
  char *h_in = nullptr;  // pinned + mapped staging for small host inputs
  size_t h_in_cap = 0;
  const void *pinned_lo = nullptr, *pinned_hi = nullptr;  // last host output range known to be pinned
  PhaseProf prof;
};

namespace gf {

struct StepBuffers {
  TargetLoc *locs;
  uint32_t *counts;
  uint32_t *offsets;
  uint32_t *nback;
  uint32_t *scan_tmp;
};

static size_t step_ws_bytes(uint64_t T) {
  uint64_t n = align_up(T + 1, 64);
  return n * sizeof(TargetLoc) + 3 * n * 4 + scan_tmp_elems(T + 1) * 4 + 256;
}
static StepBuffers carve(void *ws, uint64_t T) {
  uint64_t n = align_up(T + 1, 64);
  StepBuffers b;
  b.locs = reinterpret_cast<TargetLoc *>(ws);
  b.counts = reinterpret_cast<uint32_t *>(b.locs + n);
  b.offsets = b.counts + n;
  b.nback = b.offsets + n;
  b.scan_tmp = b.nback + n;
  return b;
}

static int choose_tpw(uint64_t T) {
  // enough warps to cover the 148 SMs a few times over before each warp takes a full 32 targets
  int tpw = 32;
  while (tpw > 1 && (T + tpw - 1) / tpw < 148ull * 16) tpw >>= 1;
  return tpw;
}

static int ensure_fused(gf_sampler *s, uint64_t tiles, cudaStream_t st) {
  if (tiles > s->fused_tiles || !s->fused.ptr) {
    size_t want = std::max<size_t>(tiles * 2, 1024);
    Scratch n;
    GF_TRY(n.reserve(256 + want * 8 + (want / 32 + 1) * 8, st));  // ticket | tile words | group words
    GF_CUDA(cudaMemsetAsync(n.ptr, 0, n.cap, st));  // generation 0 == never written
    if (s->fused.ptr) GF_CUDA(cudaFreeAsync(s->fused.ptr, st));
    s->fused = n;
    s->fused_tiles = want;
  }
  if (++s->fused_gen >= (1ull << 30)) {  // generation tag wrapped: start over
    GF_CUDA(cudaMemsetAsync(s->fused.ptr, 0, s->fused.cap, st));
    s->fused_gen = 1;
  }
  return GF_OK;
}

// one (layer, snapshot) step, everything on `st`.  meta_dev receives {T, S, T + S}.
static int launch_step(gf_sampler *s, const SampleParams &p, const int64_t *d_nodes, const float *d_ts, uint64_t T_bound,
                       const uint32_t *T_dev, const uint64_t *batch_offsets, uint32_t num_batches, EmitOut out,
                       uint32_t *meta_dev, uint32_t *meta_host, uint64_t *edge_offsets, cudaStream_t st,
                       const uint32_t *active = nullptr, const uint32_t *A_dev = nullptr) {
  if (s->variant == 3 && p.fanout <= kMaxOwnerFanout) {
    // launches whose small tiles are all resident at once (GF_PERSIST_OCC per SM): the small-tile instantiation
    static const bool no_small = getenv("GNNFLOW_B200_NO_SMALL_TILES") != nullptr;  // evidence / test knobs
    static const uint64_t small_max = getenv("GNNFLOW_B200_SMALL_TILES_MAX") ? strtoull(getenv("GNNFLOW_B200_SMALL_TILES_MAX"), nullptr, 10)
                                                                             : 148ull * GF_PERSIST_OCC;
    const bool small = !active && !no_small && T_bound <= small_max * kSmallTile;
    const uint64_t tile_targets = small ? kSmallTile : kPThreads;
    const uint64_t tiles = (T_bound + tile_targets - 1) / tile_targets;
    GF_TRY(ensure_fused(s, tiles, st));
    // resident CTAs per SM the register budget is sized for: 4 (56 registers) by default, 3 (80 registers, no spills)
    // as a measured alternative (GNNFLOW_B200_OCC)
    if (s->occ == 0) {
      const char *e = getenv("GNNFLOW_B200_OCC");
      s->occ = e && atoi(e) == 3 ? 3 : GF_PERSIST_OCC;
    }
    constexpr int RC = GF_SAMPLING_RECENT, UN = GF_SAMPLING_UNIFORM;
    const bool uni = p.policy == GF_SAMPLING_UNIFORM;
    auto kern = small
        ? (uni ? sample_persistent_kernel<false, GF_PERSIST_OCC, UN, kSmallTile> : sample_persistent_kernel<false, GF_PERSIST_OCC, RC, kSmallTile>)
        : s->occ == 3
        ? (uni ? (active ? sample_persistent_kernel<true, 3, UN> : sample_persistent_kernel<false, 3, UN>)
               : (active ? sample_persistent_kernel<true, 3, RC> : sample_persistent_kernel<false, 3, RC>))
        : (uni ? (active ? sample_persistent_kernel<true, GF_PERSIST_OCC, UN> : sample_persistent_kernel<false, GF_PERSIST_OCC, UN>)
               : (active ? sample_persistent_kernel<true, GF_PERSIST_OCC, RC> : sample_persistent_kernel<false, GF_PERSIST_OCC, RC>));
    const size_t dyn = kStages * (sizeof(TileStage) + (size_t)kPThreads * p.fanout * sizeof(OwnerT));
    const int kern_id = (active ? 1 : 0) + 2 * s->occ + (uni ? 16 : 0) + (small ? 32 : 0);
    // grid size per (instantiation, fan-out), set up once each: layers with different fan-outs alternate between
    // configurations on every call.  The dynamic shared memory limit of an instantiation only ever grows.
    const uint64_t cfg_key = ((uint64_t)kern_id << 32) | p.fanout;
    unsigned persist_grid = 0;
    for (const auto &c : s->persist_cfg)
      if (c.first == cfg_key) persist_grid = c.second;
    if (!persist_grid) {
      int occ = 0, sms = 0, dev = 0;
      GF_CUDA(cudaGetDevice(&dev));
      GF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      size_t &limit = s->persist_dyn[kern_id];
      if (dyn > limit) {
        GF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        limit = dyn;
      }
      GF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kPAll, dyn));
      persist_grid = (unsigned)std::max(1, occ * sms);
      s->persist_cfg.emplace_back(cfg_key, persist_grid);
    }
    PersistCtl ctl = {s->fused.as<unsigned int>(),
                      reinterpret_cast<unsigned long long *>(s->fused.as<char>() + 256), s->fused_gen,
                      reinterpret_cast<unsigned long long *>(s->fused.as<char>() + 256) + s->fused_tiles};
    FusedMeta fm = {meta_dev, meta_host, edge_offsets};
    s->prof.begin(st);
    gf::launch(kern, (unsigned)std::min<uint64_t>(tiles, persist_grid), kPAll, dyn, st, p, d_nodes,
               d_ts, T_bound, T_dev, batch_offsets, num_batches, out, ctl, fm, active, A_dev);
    s->prof.end(2, st, false);
    GF_CUDA(cudaGetLastError());
    return GF_OK;
  }
  GF_TRY(s->ws.reserve(step_ws_bytes(T_bound), st));
  StepBuffers b = carve(s->ws.ptr, T_bound);
  s->prof.begin(st);
  if (s->variant == 1) {
    gf::launch(locate_thread_kernel, cdiv(T_bound, kSThreads), kSThreads, 0, st, p, d_nodes, d_ts, T_bound, T_dev, b.locs,
               b.counts, b.nback);
  } else {
    int tpw = choose_tpw(T_bound);
    uint64_t warps = (T_bound + tpw - 1) / tpw;
    gf::launch(locate_warp_kernel, cdiv(warps, kSWarps), kSThreads, 0, st, p, d_nodes, d_ts, T_bound, T_dev, tpw, b.locs,
               b.counts, b.nback);
  }
  // offsets[0..T_bound]: scan over T_bound + 1 entries (the extra entry is a zero) so that offsets[T] is valid
  // for every T <= T_bound
  s->prof.end(0, st);
  GF_CUDA(cudaMemsetAsync(b.counts + T_bound, 0, 4, st));
  GF_TRY(exclusive_scan_u32(b.counts, b.offsets, T_bound + 1, meta_dev + 3, b.scan_tmp, st));
  s->prof.end(1, st);
  gf::launch(emit_kernel, cdiv((T_bound + 31) / 32, kSWarps), kSThreads, 0, st, p, d_nodes, d_ts, T_bound, T_dev, b.locs,
             b.nback, b.offsets, batch_offsets, num_batches, out);
  s->prof.end(2, st, false);
  gf::launch(chain_meta_kernel, 1, 1, 0, st, T_dev, T_bound, meta_dev + 3, meta_dev, meta_host);
  if (edge_offsets)
    gf::launch(gather_edge_offsets_kernel, cdiv((uint64_t)num_batches + 1, 256), 256, 0, st, b.offsets, batch_offsets,
               num_batches, edge_offsets);
  GF_CUDA(cudaGetLastError());
  return GF_OK;
}

static SampleParams make_params(gf_sampler *s, uint32_t layer, uint32_t snapshot) {
  SampleParams p;
  p.table = s->graph->d_table;
  p.table_len = s->graph->table_len();
  p.fanout = s->fanouts[layer];
  p.num_snapshots = s->num_snapshots;
  p.snapshot_idx = snapshot;
  p.window = s->window;
  p.prop_time = s->prop_time;
  p.policy = s->policy;
  p.seed = s->seed;
  p.launch_index = s->launch_index;
  return p;
}

static int ensure_h_meta(gf_sampler *s, size_t n) {
  if (n <= s->h_meta_cap) return GF_OK;
  if (s->h_meta) cudaFreeHost(s->h_meta);
  s->h_meta = nullptr;
  GF_CUDA(cudaHostAlloc(&s->h_meta, n * sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable));
  s->h_meta_cap = n;
  return GF_OK;
}

constexpr size_t kInPlaceOutputBytes = 4u << 20;  // multi-batch host call: larger outputs go through the copy engine
constexpr size_t kSmallInputBytes = 1u << 20;  // host inputs up to 1 MiB are read by the kernel straight from pinned memory

static bool host_range_is_pinned(gf_sampler *s, const void *p, size_t bytes) {
  const char *lo = (const char *)p, *hi = lo + bytes;
  if (s->pinned_lo && lo >= (const char *)s->pinned_lo && hi <= (const char *)s->pinned_hi) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (a.type != cudaMemoryTypeHost || a.devicePointer != p) return false;
  cudaPointerAttributes b;
  if (bytes && (cudaPointerGetAttributes(&b, hi - 1) != cudaSuccess || b.type != cudaMemoryTypeHost)) {
    cudaGetLastError();
    return false;
  }
  return true;
}

}  // namespace gf

GF_EXPORT int gf_sampler_create(gf_graph *g, const uint32_t *fanouts, uint32_t num_layers, int sampling_policy,
                                uint32_t num_snapshots, float snapshot_time_window, int prop_time, uint64_t seed,
                                gf_sampler **out) {
  if (!g || !fanouts || !out || num_layers == 0) GF_FAIL(GF_EINVAL, "gf_sampler_create: bad argument");
  if (sampling_policy != GF_SAMPLING_RECENT && sampling_policy != GF_SAMPLING_UNIFORM)
    GF_FAIL(GF_EINVAL, "strategy must be 'recent' or 'uniform'");
  if (num_snapshots == 0) GF_FAIL(GF_EINVAL, "num_snapshots must be >= 1");
  for (uint32_t l = 0; l < num_layers; l++)
    if (fanouts[l] == 0) GF_FAIL(GF_EINVAL, "fanout of layer %u is 0", l);
  gf_sampler *s = new gf_sampler();
  s->graph = g;
  s->fanouts.assign(fanouts, fanouts + num_layers);
  s->policy = sampling_policy;
  s->num_snapshots = num_snapshots;
  s->window = snapshot_time_window;
  s->prop_time = prop_time ? 1 : 0;
  s->seed = seed;
  s->prof.init(GF_SAMPLER_PHASES);
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->refs++;
  }
  *out = s;
  return GF_OK;
}

GF_EXPORT int gf_sampler_destroy(gf_sampler *s) {
  if (!s) return GF_OK;
  cudaSetDevice(s->graph->cfg.device);
  cudaDeviceSynchronize();
  s->prof.destroy();
  s->ws.release();
  s->in.release();
  s->outbuf.release();
  s->meta.release();
  s->fused.release();
  s->compact.release();
  if (s->h_meta) cudaFreeHost(s->h_meta);
  if (s->h_in) cudaFreeHost(s->h_in);
  gf_graph_destroy(s->graph);
  delete s;
  return GF_OK;
}

GF_EXPORT int gf_sampler_get_launch_index(gf_sampler *s, uint64_t *out) {
  if (!s || !out) GF_FAIL(GF_EINVAL, "null argument");
  *out = s->launch_index;
  return GF_OK;
}
GF_EXPORT int gf_sampler_set_launch_index(gf_sampler *s, uint64_t v) {
  if (!s) GF_FAIL(GF_EINVAL, "null argument");
  s->launch_index = v;
  return GF_OK;
}
GF_EXPORT int gf_sampler_set_variant(gf_sampler *s, int variant) {
  if (!s || (variant != 0 && variant != 1 && variant != 3)) GF_FAIL(GF_EINVAL, "variant must be 0, 1 or 3");
  s->variant = variant;
  return GF_OK;
}

namespace gf {

// shared implementation of sample_layer / sample: layers [layer0, layer0 + nlayers) x snapshots
static int sample_impl(gf_sampler *s, const int64_t *nodes, const float *timestamps, uint64_t T0, uint32_t layer0,
                       uint32_t nlayers, uint32_t snap0, uint32_t nsnaps, gf_sampling_result *results, int in_kind,
                       int out_kind, cudaStream_t st) {
  if ((in_kind != GF_PTR_HOST && in_kind != GF_PTR_DEVICE) || (out_kind != GF_PTR_HOST && out_kind != GF_PTR_DEVICE))
    GF_FAIL(GF_EINVAL, "bad ptr kind");
  if (T0 && (!nodes || !timestamps)) GF_FAIL(GF_EINVAL, "null input");
  gf_graph *g = s->graph;
  GF_TRY(gf_graph_flush_internal(g));  // batches queued by gf_graph_add_edges_async
  std::lock_guard<std::mutex> lk(g->mu);
  GF_CUDA(cudaSetDevice(g->cfg.device));
  const uint32_t nsteps = nlayers * nsnaps;
  // capacity checks + bounds per layer
  std::vector<uint64_t> bound(nlayers);
  for (uint32_t l = 0; l < nlayers; l++) {
    bound[l] = l == 0 ? T0 : bound[l - 1] * (1 + (uint64_t)s->fanouts[layer0 + l - 1]);
    if (bound[l] * (1 + (uint64_t)s->fanouts[layer0 + l]) >= (1ull << 32))
      GF_FAIL(GF_EINVAL, "sampling step too large: %llu targets x fanout %u", (unsigned long long)bound[l], s->fanouts[layer0 + l]);
    for (uint32_t k = 0; k < nsnaps; k++) {
      gf_sampling_result &r = results[l * nsnaps + k];
      if (r.capacity_dst < bound[l])
        GF_FAIL(GF_ECAPACITY, "result[%u][%u].capacity_dst=%llu < %llu", l, k, (unsigned long long)r.capacity_dst, (unsigned long long)bound[l]);
      if (bound[l] && (!r.all_nodes || !r.all_timestamps || !r.delta_timestamps || !r.eids || !r.row))
        GF_FAIL(GF_EINVAL, "result[%u][%u]: null output array", l, k);
    }
  }
  if (T0 == 0) {  // temporal_sampler.cu:107-114: empty input -> empty results
    for (uint32_t i = 0; i < nsteps; i++) results[i].num_dst = results[i].num_edges = 0;
    return GF_OK;
  }
  // ---- input: device pointer, or (host) small -> pinned staging read in place by the kernel, large -> H2D copy
  const int64_t *d_nodes = nodes;
  const float *d_ts = timestamps;
  if (in_kind == GF_PTR_HOST) {
    size_t off = align_up(T0 * 8, 256), bytes = off + T0 * 4;
    if (bytes <= kSmallInputBytes) {
      if (bytes > s->h_in_cap) {
        if (s->h_in) cudaFreeHost(s->h_in);
        s->h_in = nullptr;
        GF_CUDA(cudaHostAlloc(&s->h_in, kSmallInputBytes, cudaHostAllocMapped | cudaHostAllocPortable));
        s->h_in_cap = kSmallInputBytes;
      }
      // the previous call on this sampler synchronised before returning, so the staging area is free
      memcpy(s->h_in, nodes, T0 * 8);
      memcpy(s->h_in + off, timestamps, T0 * 4);
      d_nodes = reinterpret_cast<const int64_t *>(s->h_in);
      d_ts = reinterpret_cast<const float *>(s->h_in + off);
    } else {
      GF_TRY(s->in.reserve(bytes, st));
      GF_CUDA(cudaMemcpyAsync(s->in.ptr, nodes, T0 * 8, cudaMemcpyHostToDevice, st));
      GF_CUDA(cudaMemcpyAsync(s->in.as<char>() + off, timestamps, T0 * 4, cudaMemcpyHostToDevice, st));
      d_nodes = s->in.as<int64_t>();
      d_ts = reinterpret_cast<const float *>(s->in.as<char>() + off);
    }
  }
  // ---- output: caller's device arrays; caller's PINNED host arrays (written in place over PCIe by the kernel);
  //      or an internal device mirror + D2H copies for pageable host arrays
  bool direct_host = false;
  if (out_kind == GF_PTR_HOST && s->host_out_mode != 1) {
    direct_host = true;
    for (uint32_t l = 0; l < nlayers && direct_host; l++) {
      uint64_t cap_src = bound[l] * (1 + (uint64_t)s->fanouts[layer0 + l]), cap_e = bound[l] * s->fanouts[layer0 + l];
      for (uint32_t k = 0; k < nsnaps && direct_host; k++) {
        gf_sampling_result &r = results[l * nsnaps + k];
        direct_host = host_range_is_pinned(s, r.all_nodes, cap_src * 8) && host_range_is_pinned(s, r.all_timestamps, cap_src * 4) &&
                      host_range_is_pinned(s, r.delta_timestamps, cap_e * 4) && host_range_is_pinned(s, r.eids, cap_e * 8) &&
                      host_range_is_pinned(s, r.row, cap_e * 8) && (!r.col || host_range_is_pinned(s, r.col, cap_e * 8));
      }
    }
  }
  std::vector<EmitOut> outs(nsteps);
  std::vector<size_t> mirror_off(nsteps, 0);
  const bool mirror = out_kind == GF_PTR_HOST && !direct_host;
  if (mirror) {
    size_t total = 0;
    for (uint32_t l = 0; l < nlayers; l++) {
      uint64_t cap_src = bound[l] * (1 + (uint64_t)s->fanouts[layer0 + l]), cap_e = bound[l] * s->fanouts[layer0 + l];
      for (uint32_t k = 0; k < nsnaps; k++) {
        mirror_off[l * nsnaps + k] = total;
        total += align_up(cap_src * 8, 256) + align_up(cap_src * 4, 256) + align_up(cap_e * 4, 256) + 3 * align_up(cap_e * 8, 256);
      }
    }
    GF_TRY(s->outbuf.reserve(total, st));
  }
  for (uint32_t l = 0; l < nlayers; l++) {
    uint64_t cap_src = bound[l] * (1 + (uint64_t)s->fanouts[layer0 + l]), cap_e = bound[l] * s->fanouts[layer0 + l];
    for (uint32_t k = 0; k < nsnaps; k++) {
      uint32_t i = l * nsnaps + k;
      EmitOut &o = outs[i];
      memset(&o, 0, sizeof(o));
      if (!mirror) {
        o.all_nodes = results[i].all_nodes;
        o.all_ts = results[i].all_timestamps;
        o.dt = results[i].delta_timestamps;
        o.eid = results[i].eids;
        o.row = results[i].row;
        o.col = results[i].col;
      } else {
        char *b = s->outbuf.as<char>() + mirror_off[i];
        o.all_nodes = (int64_t *)b; b += align_up(cap_src * 8, 256);
        o.all_ts = (float *)b; b += align_up(cap_src * 4, 256);
        o.dt = (float *)b; b += align_up(cap_e * 4, 256);
        o.eid = (int64_t *)b; b += align_up(cap_e * 8, 256);
        o.row = (int64_t *)b; b += align_up(cap_e * 8, 256);
        o.col = results[i].col ? (int64_t *)b : nullptr;
      }
    }
  }
  GF_TRY(s->meta.reserve((size_t)nsteps * 4 * sizeof(uint32_t) + 64, st));
  uint32_t *d_meta = s->meta.as<uint32_t>();  // per step: {T, S, T + S, scratch}
  GF_TRY(ensure_h_meta(s, (size_t)nsteps * 4));
  for (uint32_t i = 0; i < nsteps; i++) *(volatile uint32_t *)(s->h_meta + i * 4 + 3) = 0u;  // set by step i's report
  for (uint32_t l = 0; l < nlayers; l++) {
    for (uint32_t k = 0; k < nsnaps; k++) {
      uint32_t i = l * nsnaps + k;
      SampleParams p = make_params(s, layer0 + l, snap0 + k);
      const int64_t *in_n = d_nodes;
      const float *in_t = d_ts;
      const uint32_t *T_dev = nullptr;
      if (l > 0) {  // temporal_sampler.cu:295-299: previous layer's all_nodes / all_timestamps, same snapshot
        uint32_t pi = (l - 1) * nsnaps + k;
        in_n = outs[pi].all_nodes;
        in_t = outs[pi].all_ts;
        T_dev = d_meta + pi * 4 + 2;
      }
      GF_TRY(launch_step(s, p, in_n, in_t, bound[l], T_dev, nullptr, 0, outs[i], d_meta + i * 4,
                         s->h_meta + i * 4, nullptr, st));  // {T, S} land straight in mapped host memory
      s->launch_index++;
    }
  }
  // Device arrays in, device arrays out: the host needs the sizes only, and those of the last step are reported when its
  // last tile's prefix is known (the steps run in stream order, so the earlier reports are complete by then); whatever
  // consumes the results is ordered behind the kernels on `st`.  Host arrays: the stream is synchronised (results written
  // in place over PCIe, the staging area of small inputs).
  if (in_kind == GF_PTR_DEVICE && out_kind == GF_PTR_DEVICE) GF_TRY(gf::wait_flag_or_sync(s->h_meta + (nsteps - 1) * 4 + 3, st));
  else GF_CUDA(cudaStreamSynchronize(st));
  for (uint32_t i = 0; i < nsteps; i++) {
    results[i].num_dst = s->h_meta[i * 4];
    results[i].num_edges = s->h_meta[i * 4 + 1];
  }
  if (mirror) {
    for (uint32_t i = 0; i < nsteps; i++) {
      uint64_t T = results[i].num_dst, S = results[i].num_edges;
      EmitOut &o = outs[i];
      GF_CUDA(cudaMemcpyAsync(results[i].all_nodes, o.all_nodes, (T + S) * 8, cudaMemcpyDeviceToHost, st));
      GF_CUDA(cudaMemcpyAsync(results[i].all_timestamps, o.all_ts, (T + S) * 4, cudaMemcpyDeviceToHost, st));
      if (S) {
        GF_CUDA(cudaMemcpyAsync(results[i].delta_timestamps, o.dt, S * 4, cudaMemcpyDeviceToHost, st));
        GF_CUDA(cudaMemcpyAsync(results[i].eids, o.eid, S * 8, cudaMemcpyDeviceToHost, st));
        GF_CUDA(cudaMemcpyAsync(results[i].row, o.row, S * 8, cudaMemcpyDeviceToHost, st));
        if (o.col) GF_CUDA(cudaMemcpyAsync(results[i].col, o.col, S * 8, cudaMemcpyDeviceToHost, st));
      }
    }
    GF_CUDA(cudaStreamSynchronize(st));
  }
  return GF_OK;
}

}  // namespace gf

GF_EXPORT int gf_sampler_sample_layer(gf_sampler *s, const int64_t *nodes, const float *timestamps, uint64_t num_targets,
                                      uint32_t layer, uint32_t snapshot, gf_sampling_result *result, int in_kind,
                                      int out_kind, void *stream) {
  if (!s || !result) GF_FAIL(GF_EINVAL, "null argument");
  if (layer >= s->fanouts.size()) GF_FAIL(GF_EINVAL, "layer %u out of range", layer);
  if (snapshot >= s->num_snapshots) GF_FAIL(GF_EINVAL, "snapshot %u out of range", snapshot);
  return sample_impl(s, nodes, timestamps, num_targets, layer, 1, snapshot, 1, result, in_kind, out_kind,
                     (cudaStream_t)stream);
}

GF_EXPORT int gf_sampler_sample(gf_sampler *s, const int64_t *nodes, const float *timestamps, uint64_t num_targets,
                                gf_sampling_result *results, int in_kind, int out_kind, void *stream) {
  if (!s || !results) GF_FAIL(GF_EINVAL, "null argument");
  return sample_impl(s, nodes, timestamps, num_targets, 0, (uint32_t)s->fanouts.size(), 0, s->num_snapshots, results,
                     in_kind, out_kind, (cudaStream_t)stream);
}

// Multi-batch launches of at least `compact_min` targets (GNNFLOW_B200_COMPACT_MIN, default 8 000 000; 0 = never) first
// list the targets whose vertex has out-edges (one launch) and run the tiles over that list.  On a directed bipartite
// stream 85-90 % of a second layer's targets are destination vertices without out-edges: each used to take a slot of
// the ordered tile pipeline (~28 ps) to emit nothing -- REDDIT two-layer replay, second layer: 0.476 -> 0.193 ms recent,
// 0.807 -> 0.257 ms uniform.  The listing pass costs 30-45 us whatever the size (profiles/r01_s9_compaction.json), which
// a launch without such targets does not get back (WIKI second layer, 4 M targets: +13 %; headline, 2 M: +20 %), and
// the density is not known before the pass: hence the size threshold.  Sets *active / *A_dev (nullptr: all targets).
static int build_active_list(gf_sampler *s, const int64_t *d_nodes, uint64_t T, const uint32_t **active,
                             const uint32_t **A_dev, cudaStream_t st) {
  *active = nullptr;
  *A_dev = nullptr;
  if (s->compact_min < 0) {
    const char *e = getenv("GNNFLOW_B200_COMPACT_MIN");
    s->compact_min = e ? atoll(e) : 8000000;
    if (s->compact_min < 0) s->compact_min = 0;
  }
  gf_graph *g = s->graph;
  if (s->compact_min == 0 || T < (uint64_t)s->compact_min || s->variant != 3 || !g->d_is_src || !g->table_len()) return GF_OK;
  const uint64_t chunks = (T + kActChunk - 1) / kActChunk;
  const size_t o_status = 256, o_list = o_status + align_up(chunks * 8, 256);
  GF_TRY(s->compact.reserve(o_list + align_up(T * 4, 256), st));
  char *b = s->compact.as<char>();
  GF_CUDA(cudaMemsetAsync(b, 0, o_list, st));  // ticket, count, status words (generation 1 below)
  uint32_t *ctl = reinterpret_cast<uint32_t *>(b);
  uint32_t *list = reinterpret_cast<uint32_t *>(b + o_list);
  LookbackCtl lb = {ctl, reinterpret_cast<unsigned long long *>(b + o_status), 1ull};
  gf::launch(active_list_kernel, (unsigned)chunks, kScanThreads, 0, st, d_nodes, T, g->d_is_src, (uint64_t)g->table_len(), list,
             lb, ctl + 1);
  GF_CUDA(cudaGetLastError());
  *active = list;
  *A_dev = ctl + 1;
  return GF_OK;
}

GF_EXPORT int gf_sampler_chain_batched(const int64_t *nodes, const float *timestamps, uint64_t num_targets,
                                       const uint64_t *batch_offsets, uint64_t num_batches, const int64_t *nbr,
                                       const float *nbr_ts, const uint64_t *edge_offsets, uint64_t max_edges,
                                       int64_t *nodes_out, float *timestamps_out, uint64_t *batch_offsets_out,
                                       void *stream) {
  if (!batch_offsets || !edge_offsets || !batch_offsets_out) GF_FAIL(GF_EINVAL, "chain_batched: null offsets");
  if (num_batches == 0 || num_batches >= (1ull << 31)) GF_FAIL(GF_EINVAL, "chain_batched: bad num_batches");
  if ((num_targets + max_edges) && (!nodes || !timestamps || !nodes_out || !timestamps_out))
    GF_FAIL(GF_EINVAL, "chain_batched: null array");
  if (max_edges && (!nbr || !nbr_ts)) GF_FAIL(GF_EINVAL, "chain_batched: null neighbour array");
  const uint64_t bound = num_targets + max_edges + num_batches + 1;  // the exact total is edge_offsets[num_batches], on the device
  const unsigned blocks = (unsigned)std::min<uint64_t>(cdiv(bound, 256), 148ull * 16);
  gf::launch(chain_batched_kernel, blocks, 256, 0, (cudaStream_t)stream, nodes, timestamps, num_targets, batch_offsets,
             (uint32_t)num_batches, nbr, nbr_ts, edge_offsets, nodes_out, timestamps_out, batch_offsets_out);
  GF_CUDA(cudaGetLastError());
  return GF_OK;
}

namespace gf {
// neighbour ids and rows of a multi-batch launch, 64 -> 32 bits (vertex ids are < 2^32 by the store's contract, rows < the
// number of targets): the narrow copies are what crosses PCIe in the ids32 host call
__global__ void __launch_bounds__(256) narrow_ids_kernel(const int64_t *__restrict__ nbr, const int64_t *__restrict__ row,
                                                         const uint64_t *__restrict__ total, uint32_t *__restrict__ nbr32,
                                                         uint32_t *__restrict__ row32) {
  const uint64_t S = *total, quads = S / 4, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += stride) {
    const longlong2 a0 = __ldg(reinterpret_cast<const longlong2 *>(nbr) + 2 * q), a1 = __ldg(reinterpret_cast<const longlong2 *>(nbr) + 2 * q + 1);
    const longlong2 b0 = __ldg(reinterpret_cast<const longlong2 *>(row) + 2 * q), b1 = __ldg(reinterpret_cast<const longlong2 *>(row) + 2 * q + 1);
    reinterpret_cast<uint4 *>(nbr32)[q] = make_uint4((uint32_t)a0.x, (uint32_t)a0.y, (uint32_t)a1.x, (uint32_t)a1.y);
    reinterpret_cast<uint4 *>(row32)[q] = make_uint4((uint32_t)b0.x, (uint32_t)b0.y, (uint32_t)b1.x, (uint32_t)b1.y);
  }
  if (blockIdx.x == 0 && threadIdx.x < (S & 3)) {
    const uint64_t i = quads * 4 + threadIdx.x;
    nbr32[i] = (uint32_t)nbr[i];
    row32[i] = (uint32_t)row[i];
  }
}
}  // namespace gf

// out_nbr / out_row: int64 arrays, or (ids32, host arrays only) uint32 arrays
static int sample_layer_batched_impl(gf_sampler *s, const int64_t *nodes, const float *timestamps, uint64_t num_targets,
                                     const uint64_t *batch_offsets, uint64_t num_batches, uint32_t layer, uint32_t snapshot,
                                     void *out_nbr_, float *out_ts, float *out_dt, int64_t *out_eid, void *out_row_,
                                     uint64_t *edge_offsets, int ptr_kind, bool ids32, void *stream) {
  int64_t *out_nbr = reinterpret_cast<int64_t *>(out_nbr_), *out_row = reinterpret_cast<int64_t *>(out_row_);
  if (!s || !batch_offsets || !edge_offsets) GF_FAIL(GF_EINVAL, "null argument");
  if (layer >= s->fanouts.size() || snapshot >= s->num_snapshots) GF_FAIL(GF_EINVAL, "layer/snapshot out of range");
  if (ptr_kind != GF_PTR_DEVICE && ptr_kind != GF_PTR_HOST) GF_FAIL(GF_EINVAL, "bad ptr kind");
  if (ids32 && ptr_kind != GF_PTR_HOST) GF_FAIL(GF_EINVAL, "32-bit ids are a host-array format");
  if (num_batches == 0 || num_batches >= (1ull << 31)) GF_FAIL(GF_EINVAL, "bad num_batches");
  cudaStream_t st = (cudaStream_t)stream;
  gf_graph *g = s->graph;
  GF_TRY(gf_graph_flush_internal(g));  // batches queued by gf_graph_add_edges_async
  std::lock_guard<std::mutex> lk(g->mu);
  GF_CUDA(cudaSetDevice(g->cfg.device));
  const uint64_t T = num_targets;
  const uint64_t F = s->fanouts[layer];
  if (T * (1 + F) >= (1ull << 32)) GF_FAIL(GF_EINVAL, "too many targets in one launch");
  const bool host = ptr_kind == GF_PTR_HOST;
  if (T == 0) {
    if (host) memset(edge_offsets, 0, (num_batches + 1) * 8);
    else GF_CUDA(cudaMemsetAsync(edge_offsets, 0, (num_batches + 1) * 8, st));
    return GF_OK;
  }
  if (!nodes || !timestamps || !out_nbr || !out_ts || !out_dt || !out_eid || !out_row) GF_FAIL(GF_EINVAL, "null array");
  GF_TRY(s->meta.reserve(64, st));
  SampleParams p = make_params(s, layer, snapshot);
  EmitOut o;
  memset(&o, 0, sizeof(o));
  uint32_t *nbr32 = nullptr, *row32 = nullptr;
  if (!host) {
    o.nbr = out_nbr;
    o.nbr_ts = out_ts;
    o.dt = out_dt;
    o.eid = out_eid;
    o.row = out_row;
    const uint32_t *active, *A_dev;
    GF_TRY(build_active_list(s, nodes, T, &active, &A_dev, st));
    GF_TRY(launch_step(s, p, nodes, timestamps, T, nullptr, batch_offsets, (uint32_t)num_batches, o, s->meta.as<uint32_t>(),
                       nullptr, edge_offsets, st, active, A_dev));
    s->launch_index += num_batches;
    return GF_OK;
  }
  // ---- HOST arrays: inputs are copied to the device (the locate phase is latency-bound, it must not read over PCIe);
  //      outputs are written by the kernel IN PLACE into the caller's arrays when those are pinned (full-line
  //      streaming stores over PCIe, no staging copy and no second pass), else through a device mirror + D2H copies.
  const size_t o_ts = align_up(T * 8, 256), o_bo = o_ts + align_up(T * 4, 256), o_eo = o_bo + align_up((num_batches + 1) * 8, 256);
  GF_TRY(s->in.reserve(o_eo + align_up((num_batches + 1) * 8, 256), st));
  char *din = s->in.as<char>();
  GF_CUDA(cudaMemcpyAsync(din, nodes, T * 8, cudaMemcpyHostToDevice, st));
  GF_CUDA(cudaMemcpyAsync(din + o_ts, timestamps, T * 4, cudaMemcpyHostToDevice, st));
  GF_CUDA(cudaMemcpyAsync(din + o_bo, batch_offsets, (num_batches + 1) * 8, cudaMemcpyHostToDevice, st));
  uint64_t *d_eo = reinterpret_cast<uint64_t *>(din + o_eo);
  const uint64_t cap_e = T * F;
  // measured on B200 / PCIe 5 x16 (profiles/r01_bench_s5_hostmode*.json): SM-issued stores reach 40 GB/s, the copy engine
  // 54 GB/s, so in-place output only pays when the arrays are small enough for the extra copy launch to matter
  const bool want_direct = s->host_out_mode == 2 || (s->host_out_mode == 0 && cap_e * 32 <= kInPlaceOutputBytes);
  const bool direct = !ids32 && want_direct && host_range_is_pinned(s, out_nbr, cap_e * 8) &&
                      host_range_is_pinned(s, out_ts, cap_e * 4) && host_range_is_pinned(s, out_dt, cap_e * 4) &&
                      host_range_is_pinned(s, out_eid, cap_e * 8) && host_range_is_pinned(s, out_row, cap_e * 8);
  if (direct) {
    o.nbr = out_nbr;
    o.nbr_ts = out_ts;
    o.dt = out_dt;
    o.eid = out_eid;
    o.row = out_row;
  } else {
    const size_t a8 = align_up(cap_e * 8, 256), a4 = align_up(cap_e * 4, 256);
    GF_TRY(s->outbuf.reserve(3 * a8 + (ids32 ? 4 : 2) * a4, st));
    char *b = s->outbuf.as<char>();
    o.nbr = (int64_t *)b; b += a8;
    o.eid = (int64_t *)b; b += a8;
    o.row = (int64_t *)b; b += a8;
    o.nbr_ts = (float *)b; b += a4;
    o.dt = (float *)b; b += a4;
    nbr32 = (uint32_t *)b; b += a4;
    row32 = (uint32_t *)b;
  }
  const uint32_t *active, *A_dev;
  GF_TRY(build_active_list(s, reinterpret_cast<const int64_t *>(din), T, &active, &A_dev, st));
  GF_TRY(launch_step(s, p, reinterpret_cast<const int64_t *>(din), reinterpret_cast<const float *>(din + o_ts), T, nullptr,
                     reinterpret_cast<const uint64_t *>(din + o_bo), (uint32_t)num_batches, o, s->meta.as<uint32_t>(),
                     nullptr, d_eo, st, active, A_dev));
  s->launch_index += num_batches;
  GF_CUDA(cudaMemcpyAsync(edge_offsets, d_eo, (num_batches + 1) * 8, cudaMemcpyDeviceToHost, st));
  if (ids32)  // queued behind the sampling launch: done by the time the host has looked at the offsets
    gf::launch(gf::narrow_ids_kernel, (unsigned)std::min<uint64_t>(cdiv(cap_e, 1024), 148ull * 8), 256, 0, st, o.nbr, o.row,
               d_eo + num_batches, nbr32, row32);
  GF_CUDA(cudaStreamSynchronize(st));
  if (!direct) {
    const uint64_t S = edge_offsets[num_batches];
    if (S) {
      if (ids32) {
        GF_CUDA(cudaMemcpyAsync(out_nbr_, nbr32, S * 4, cudaMemcpyDeviceToHost, st));
        GF_CUDA(cudaMemcpyAsync(out_row_, row32, S * 4, cudaMemcpyDeviceToHost, st));
      } else {
        GF_CUDA(cudaMemcpyAsync(out_nbr, o.nbr, S * 8, cudaMemcpyDeviceToHost, st));
        GF_CUDA(cudaMemcpyAsync(out_row, o.row, S * 8, cudaMemcpyDeviceToHost, st));
      }
      GF_CUDA(cudaMemcpyAsync(out_eid, o.eid, S * 8, cudaMemcpyDeviceToHost, st));
      GF_CUDA(cudaMemcpyAsync(out_ts, o.nbr_ts, S * 4, cudaMemcpyDeviceToHost, st));
      GF_CUDA(cudaMemcpyAsync(out_dt, o.dt, S * 4, cudaMemcpyDeviceToHost, st));
      GF_CUDA(cudaStreamSynchronize(st));
    }
  }
  return GF_OK;
}

GF_EXPORT int gf_sampler_sample_layer_batched(gf_sampler *s, const int64_t *nodes, const float *timestamps,
                                              uint64_t num_targets, const uint64_t *batch_offsets, uint64_t num_batches,
                                              uint32_t layer, uint32_t snapshot, int64_t *out_nbr, float *out_ts,
                                              float *out_dt, int64_t *out_eid, int64_t *out_row, uint64_t *edge_offsets,
                                              int ptr_kind, void *stream) {
  return sample_layer_batched_impl(s, nodes, timestamps, num_targets, batch_offsets, num_batches, layer, snapshot, out_nbr,
                                   out_ts, out_dt, out_eid, out_row, edge_offsets, ptr_kind, false, stream);
}

GF_EXPORT int gf_sampler_sample_layer_batched_ids32(gf_sampler *s, const int64_t *nodes, const float *timestamps,
                                                      uint64_t num_targets, const uint64_t *batch_offsets,
                                                      uint64_t num_batches, uint32_t layer, uint32_t snapshot,
                                                      uint32_t *out_nbr, float *out_ts, float *out_dt, int64_t *out_eid,
                                                      uint32_t *out_row, uint64_t *edge_offsets, void *stream) {
  return sample_layer_batched_impl(s, nodes, timestamps, num_targets, batch_offsets, num_batches, layer, snapshot, out_nbr,
                                   out_ts, out_dt, out_eid, out_row, edge_offsets, GF_PTR_HOST, true, stream);
}

// The caller promises that [ptr, ptr + bytes) is ONE pinned, mapped host allocation that stays alive (and pinned) until it
// is unbound (ptr == NULL), re-bound, or the sampler is destroyed: result arrays inside it are then written in place
// without asking the driver about every array on every call (6 x cudaPointerGetAttributes per step, 8-12 us per
// per-batch call).  The range is verified here, page by page, once.
GF_EXPORT int gf_sampler_bind_host_outputs(gf_sampler *s, const void *ptr, uint64_t bytes) {
  if (!s) GF_FAIL(GF_EINVAL, "null sampler");
  s->pinned_lo = s->pinned_hi = nullptr;
  if (!ptr || !bytes) return GF_OK;
  if (bytes > (256ull << 20)) GF_FAIL(GF_EUNSUPPORTED, "gf_sampler_bind_host_outputs: ranges above 256 MiB are checked per call");
  const char *lo = (const char *)ptr, *hi = lo + bytes;
  for (const char *q = lo; q < hi; q = (q + 4096 < hi || q == hi - 1) ? q + 4096 : hi - 1) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, q) != cudaSuccess || a.type != cudaMemoryTypeHost || a.devicePointer != q) {
      cudaGetLastError();
      GF_FAIL(GF_EINVAL, "gf_sampler_bind_host_outputs: %p + %llu is not pinned, mapped host memory", ptr,
              (unsigned long long)(q - lo));
    }
  }
  s->pinned_lo = lo;
  s->pinned_hi = hi;
  return GF_OK;
}

GF_EXPORT int gf_sampler_set_host_output_mode(gf_sampler *s, int mode) {
  if (!s) GF_FAIL(GF_EINVAL, "null sampler");
  if (mode < 0 || mode > 2) GF_FAIL(GF_EINVAL, "host output mode must be 0 (auto), 1 (device mirror + copies) or 2 (in place when pinned)");
  s->host_out_mode = mode;
  return GF_OK;
}

GF_EXPORT int gf_sampler_set_profiling(gf_sampler *s, int on) {
  if (!s) GF_FAIL(GF_EINVAL, "null sampler");
  s->prof.on = on != 0;
  return GF_OK;
}
GF_EXPORT int gf_sampler_get_profile(gf_sampler *s, double *ms, uint64_t *count, int reset) {
  if (!s || !ms || !count) GF_FAIL(GF_EINVAL, "null argument");
  s->prof.collect();
  for (int i = 0; i < GF_SAMPLER_PHASES; i++) { ms[i] = s->prof.ms[i]; count[i] = s->prof.count[i]; }
  if (reset) s->prof.reset();
  return GF_OK;
}

// =====================================================================================================
// Partitioned sampling over NVLink peer memory (include/gnnflow_b200.h: gf_peer_*, gf_sampler_sample_layer_partitioned).
//
// Every rank owns one exchange WINDOW (a single cudaMalloc, shared with the other processes of the box through
// cudaIpc handles and mapped into every process).  One (layer, snapshot) step on every rank, all stream-ordered:
//   route   : owner of every target; requests {nid, ts, target index} are written straight into the OWNER's window
//             (st.global over NVLink) at [requester][stable position]; the last CTA publishes count + flag (release.sys)
//   sample  : waits for the P request flags, samples every request against the local partition and writes the
//             neighbours straight into the REQUESTER's window at [owner][position * F + slot]; flag when done
//   merge   : waits for the P response flags, one look-back scan over the targets in their original order
//             compacts the padded responses into the single-GPU (target-major) SamplingResult.
// Flags carry a step generation, so nothing is reset between steps; a window is reused only after its reader has
// left the previous step (stream order + the flags make that transitive).
// =====================================================================================================
namespace gf {

constexpr uint32_t kMaxPeers = 8;

struct ReqRec {  // 16 B, one 128-bit store over NVLink
  int64_t nid;
  float ts;
  uint32_t idx;
};

struct WinLayout {  // byte offsets inside a window; identical on every rank
  uint64_t req_flag, resp_flag, req_count, req, resp_cnt, resp_nbr, resp_eid, resp_ts, resp_dt, total;
  uint64_t cap;
  uint32_t P, F;
};

static WinLayout make_layout(uint32_t P, uint64_t cap, uint32_t F) {
  WinLayout L;
  uint64_t o = 0;
  auto take = [&](uint64_t bytes) {
    uint64_t at = o;
    o += align_up(bytes, 256);
    return at;
  };
  L.req_flag = take(kMaxPeers * 8);
  L.resp_flag = take(kMaxPeers * 8);
  L.req_count = take(kMaxPeers * 4);
  L.req = take((uint64_t)P * cap * sizeof(ReqRec));
  L.resp_cnt = take((uint64_t)P * cap * 4);
  L.resp_nbr = take((uint64_t)P * cap * F * 8);
  L.resp_eid = take((uint64_t)P * cap * F * 8);
  L.resp_ts = take((uint64_t)P * cap * F * 4);
  L.resp_dt = take((uint64_t)P * cap * F * 4);
  L.total = o;
  L.cap = cap;
  L.P = P;
  L.F = F;
  return L;
}

struct PeerView {
  char *win[kMaxPeers];  // every rank's window as mapped in this process; win[rank] is the local one
  WinLayout L;
  uint32_t rank;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ int owner_of_vertex(int64_t v, const int8_t *__restrict__ table, uint64_t table_len, uint32_t P) {
  if (table) return (v >= 0 && (uint64_t)v < table_len) ? (int)table[v] : -1;
  if (v < 0) return -1;
  unsigned long long z = (unsigned long long)v + 0x9E3779B97F4A7C15ull;  // splitmix64, gnnflow_b200/distributed.py
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (int)(z % P);
}

constexpr int kRThreads = 256;
constexpr int kQThreads = 256;  // sample_partition_kernel

// route: ONE kernel.  A CTA takes a ticket (tile = kRItems groups of kRThreads consecutive targets), counts its targets
// per owner, publishes the counts, resolves the counts of the tiles before it with a per-owner decoupled look-back (the
// onesweep status words of gf_primitives.cuh, thread q follows owner q), and writes its request records -- 16 bytes, one
// 128-bit store over NVLink -- at their final positions in the owners' windows; the order of one rank's requests at an
// owner is the order of its targets.  The CTA that finishes last publishes the totals and raises this rank's flag.
constexpr int kRItems = 4;
__global__ void __launch_bounds__(kRThreads) route_kernel(const int64_t *__restrict__ nodes, const float *__restrict__ ts,
                                                          uint64_t T, const int8_t *__restrict__ table, uint64_t table_len,
                                                          PeerView pv, int32_t *__restrict__ owner_local,
                                                          uint32_t *__restrict__ pos_local, uint32_t *ticket,
                                                          uint32_t *status /* [tiles][kMaxPeers] */, uint32_t *totals,
                                                          unsigned int *done, unsigned long long gen, uint32_t *overflow) {
  __shared__ uint32_t warp_cnt[kRThreads / 32][kMaxPeers];
  __shared__ uint32_t run[kMaxPeers];   // requests of this CTA's earlier groups, per owner
  __shared__ uint32_t base[kMaxPeers];  // requests of the earlier tiles, per owner
  __shared__ uint32_t s_tile;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t P = pv.L.P;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  if (threadIdx.x < kMaxPeers) run[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t tile = s_tile, ntiles = gridDim.x;
  const uint64_t tile_first = (uint64_t)tile * kRItems * kRThreads;
  // ---- owners of this tile's targets (kept in registers), counts per owner
  int8_t own[kRItems];
  uint32_t mine[kMaxPeers];
#pragma unroll
  for (int q = 0; q < (int)kMaxPeers; q++) mine[q] = 0;
#pragma unroll
  for (int it = 0; it < kRItems; it++) {
    const uint64_t i = tile_first + (uint64_t)it * kRThreads + threadIdx.x;
    int o = i < T ? owner_of_vertex(nodes[i], table, table_len, P) : -1;
    if (o >= (int)P) o = -1;
    own[it] = (int8_t)o;
#pragma unroll
    for (int q = 0; q < (int)kMaxPeers; q++) mine[q] += __popc(__ballot_sync(0xffffffffu, o == q));
  }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < (int)kMaxPeers; q++)
      if (mine[q]) atomicAdd(&run[q], mine[q]);
  }
  __syncthreads();
  if (w < (int)kMaxPeers) {  // warp q: look-back over the earlier tiles for owner q, 32 tiles per round trip
    const uint32_t q = w, c = run[q];
    uint32_t *my = status + (uint64_t)tile * kMaxPeers + q;
    uint32_t excl = 0;
    if (tile == 0) {
      if (lane == 0) os_store(my, kOsIncl | c);
    } else {
      if (lane == 0) os_store(my, kOsAgg | c);
      int64_t t0 = (int64_t)tile - 1;
      while (true) {
        const int64_t t = t0 - lane;
        const uint32_t sw = t >= 0 ? os_load(status + (uint64_t)t * kMaxPeers + q) : kOsIncl;  // before the first tile: nothing
        const bool ready = (sw & (kOsAgg | kOsIncl)) != 0;
        const unsigned incl = __ballot_sync(0xffffffffu, ready && (sw & kOsIncl));
        const unsigned nready = __ballot_sync(0xffffffffu, !ready);
        const int stop = incl ? __ffs(incl) - 1 : 31;  // nearest inclusive predecessor, or the whole window
        const unsigned need = stop == 31 ? 0xffffffffu : ((2u << stop) - 1u);
        if (nready & need) continue;  // a needed predecessor has not published yet: poll again
        excl += __reduce_add_sync(0xffffffffu, lane <= stop ? (sw & kOsValue) : 0u);
        if (incl) break;
        t0 -= 32;
      }
      if (lane == 0) os_store(my, kOsIncl | (excl + c));
    }
    if (lane == 0) {
      base[q] = excl;
      if (tile == ntiles - 1) totals[q] = excl + c;
    }
  }
  __syncthreads();
  if (threadIdx.x < kMaxPeers) run[threadIdx.x] = 0;
  __syncthreads();
  // ---- request records at their final positions
#pragma unroll 1
  for (int it = 0; it < kRItems; it++) {
    if (tile_first + (uint64_t)it * kRThreads >= T) break;  // uniform over the CTA
    const uint64_t i = tile_first + (uint64_t)it * kRThreads + threadIdx.x;
    int o = -1;
#pragma unroll
    for (int j = 0; j < kRItems; j++)
      if (j == it) o = own[j];  // static indexing keeps own[] in registers
    uint32_t rank_in_warp = 0;
    for (uint32_t q = 0; q < P; q++) {  // stable rank among the targets of the same owner
      const unsigned m = __ballot_sync(0xffffffffu, o == (int)q);
      if (o == (int)q) rank_in_warp = __popc(m & ((1u << lane) - 1u));
      if (lane == 0) warp_cnt[w][q] = __popc(m);
    }
    __syncthreads();
    if (o >= 0) {
      uint32_t before = base[o] + run[o];
      for (int ww = 0; ww < w; ww++) before += warp_cnt[ww][o];
      const uint32_t pos = before + rank_in_warp;
      if (pos < pv.L.cap) {
        ReqRec r = {nodes[i], ts[i], (uint32_t)i};
        ReqRec *dst = reinterpret_cast<ReqRec *>(pv.win[o] + pv.L.req) + ((uint64_t)pv.rank * pv.L.cap + pos);
        *reinterpret_cast<int4 *>(dst) = *reinterpret_cast<const int4 *>(&r);
      } else {
        *overflow = 1;
      }
      pos_local[i] = pos;
    }
    if (i < T) owner_local[i] = o;
    __syncthreads();
    if (threadIdx.x < P) {
      uint32_t c = 0;
      for (int ww = 0; ww < kRThreads / 32; ww++) c += warp_cnt[ww][threadIdx.x];
      run[threadIdx.x] += c;
    }
    __syncthreads();
  }
  // ---- the last CTA to finish publishes the request counts and raises this rank's flag in every owner's window
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned s_last;
  if (threadIdx.x == 0) {
    const unsigned k = atomicAdd(done, 1u);
    s_last = k == gridDim.x - 1;
    if (s_last) *done = 0;
  }
  __syncthreads();
  if (s_last && threadIdx.x < P) {
    const uint32_t q = threadIdx.x;
    const uint32_t tot = *reinterpret_cast<volatile uint32_t *>(totals + q);
    reinterpret_cast<uint32_t *>(pv.win[q] + pv.L.req_count)[pv.rank] = min(tot, (uint32_t)pv.L.cap);
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long *>(pv.win[q] + pv.L.req_flag) + pv.rank, gen);
  }
}

// stream-ordered rendezvous: returns once every peer has raised its flag for this step
__global__ void wait_flags_kernel(const unsigned long long *flags, uint32_t P, unsigned long long gen) {
  if (threadIdx.x < P)
    while (ld_acquire_sys(flags + threadIdx.x) < gen) __nanosleep(64);
}

// the owner side: every request of every rank, against the local partition; padded (F slots per target) output
// straight into the requesters' windows
__global__ void __launch_bounds__(kQThreads, 4) sample_partition_kernel(SampleParams p, PeerView pv, unsigned int *done,
                                                                      unsigned long long gen) {
  __shared__ uint64_t s_desc[kQThreads], s_payload[kQThreads];
  __shared__ uint32_t s_cap[kQThreads], s_idx_hi[kQThreads], s_ncand[kQThreads], s_back[kQThreads], s_cnt[kQThreads],
      s_idx[kQThreads], s_slot0[kQThreads], s_cumd[kQThreads], s_cumf[kQThreads];
  __shared__ uint8_t s_req[kQThreads];
  __shared__ float s_root[kQThreads];
  __shared__ uint32_t s_prefix[kMaxPeers + 1];
  const int tid = threadIdx.x;
  const uint32_t P = pv.L.P, F = p.fanout;
  char *mine = pv.win[pv.rank];
  if (tid == 0) {
    uint32_t run = 0;
    const uint32_t *c = reinterpret_cast<const uint32_t *>(mine + pv.L.req_count);
    for (uint32_t r = 0; r < P; r++) {
      s_prefix[r] = run;
      run += c[r];
    }
    for (uint32_t r = P; r <= kMaxPeers; r++) s_prefix[r] = run;
  }
  __syncthreads();
  const uint32_t R = s_prefix[kMaxPeers];
  const uint32_t ntiles = (R + kQThreads - 1) / kQThreads;
  const ReqRec *req = reinterpret_cast<const ReqRec *>(mine + pv.L.req);
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint32_t v = tile * kQThreads + tid;
    LocatedT loc;
    loc.desc = 0; loc.payload = 0; loc.cap = 0; loc.idx_hi = 0; loc.ncand = 0; loc.back = 0; loc.cum_d = 0; loc.cum_f = 0;
    uint32_t cnt = 0, r = 0, j = 0, idx = 0;
    float root = 0.f;
    if (v < R) {
      while (r + 1 < P && v >= s_prefix[r + 1]) r++;
      j = v - s_prefix[r];
      const int4 raw = *reinterpret_cast<const int4 *>(req + ((uint64_t)r * pv.L.cap + j));
      const ReqRec rec = *reinterpret_cast<const ReqRec *>(&raw);
      root = rec.ts;
      idx = rec.idx;
      cnt = locate_target(p, rec.nid, root, loc);
      reinterpret_cast<uint32_t *>(pv.win[r] + pv.L.resp_cnt)[(uint64_t)pv.rank * pv.L.cap + j] = cnt;
    }
    s_desc[tid] = loc.desc; s_payload[tid] = loc.payload; s_cap[tid] = loc.cap; s_idx_hi[tid] = loc.idx_hi;
    s_ncand[tid] = loc.ncand; s_back[tid] = loc.back; s_cnt[tid] = cnt; s_idx[tid] = idx; s_root[tid] = root;
    s_cumd[tid] = loc.cum_d; s_cumf[tid] = loc.cum_f;
    s_req[tid] = (uint8_t)r;
    s_slot0[tid] = j;
    __syncthreads();
    const uint32_t nslots = kQThreads * F;
    for (uint32_t q = tid; q < nslots; q += kQThreads) {
      const uint32_t jt = q / F, k = q - jt * F;
      if (k >= s_cnt[jt]) continue;
      const Slot sl = resolve_slot(p, s_payload[jt], s_cap[jt], s_idx_hi[jt], s_ncand[jt], s_back[jt], s_desc[jt],
                                   s_cumd[jt], s_cumf[jt], s_idx[jt], k, 0, s_root[jt]);
      const EdgeRec rec = ld_rec(blk_rec(sl.payload, sl.cap) + sl.idx);
      const float t = rec.ts;
      const int64_t nb = (int64_t)rec.dst, ed = rec.eid;
      char *w = pv.win[s_req[jt]];
      const uint64_t o = ((uint64_t)pv.rank * pv.L.cap + s_slot0[jt]) * pv.L.F + k;
      reinterpret_cast<int64_t *>(w + pv.L.resp_nbr)[o] = nb;
      reinterpret_cast<int64_t *>(w + pv.L.resp_eid)[o] = ed;
      reinterpret_cast<float *>(w + pv.L.resp_ts)[o] = p.prop_time ? sl.root : t;
      reinterpret_cast<float *>(w + pv.L.resp_dt)[o] = __fsub_rn(sl.root, t);
    }
    __syncthreads();
  }
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned s_last;
  if (tid == 0) {
    const unsigned k = atomicAdd(done, 1u);
    s_last = k == gridDim.x - 1;
    if (s_last) *done = 0;
  }
  __syncthreads();
  if (s_last && tid < (int)P)
    st_release_sys(reinterpret_cast<unsigned long long *>(pv.win[tid] + pv.L.resp_flag) + pv.rank, gen);
}

// merge: compaction of the padded responses in the original target order.  One tile = kScanThreads targets: counts ->
// tile scan -> decoupled look-back for the tile's first output slot -> one thread per OUTPUT SLOT copies a neighbour
// from the response area of this rank's window (slot -> target by a search of the tile's offsets in shared memory), so
// that every output array is written in full lines.
struct MergeArgs {
  const int64_t *nodes;
  const float *ts;
  const int32_t *owner;
  const uint32_t *pos;
  const uint32_t *resp_cnt;
  const int64_t *r_nbr, *r_eid;
  const float *r_ts, *r_dt;
  uint64_t cap, T;
  uint32_t F;
  EmitOut out;
};
__global__ void __launch_bounds__(kScanThreads) merge_kernel(MergeArgs m, LookbackCtl ctl, uint32_t *total_out) {
  __shared__ uint32_t s_off[kScanThreads + 1];
  __shared__ uint64_t s_src[kScanThreads];
  __shared__ uint32_t s_tile, s_base, s_total;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) {
    const unsigned t = atomicAdd(ctl.ticket, 1u);
    if (t == gridDim.x - 1) *ctl.ticket = 0;  // every tile of this launch has its ticket: re-arm
    s_tile = t;
  }
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t i = (uint64_t)tile * kScanThreads + tid;
  uint32_t c = 0;
  uint64_t src = 0;
  if (i < m.T) {
    const int o = m.owner[i];
    if (o >= 0) {
      src = (uint64_t)o * m.cap + m.pos[i];
      c = m.resp_cnt[src];
      src *= m.F;
    }
    m.out.all_nodes[i] = m.nodes[i];
    m.out.all_ts[i] = m.ts[i];
  }
  const uint32_t off = block_excl_scan(c, &s_total);
  s_off[tid] = off;
  s_src[tid] = src;
  if (tid < 32) {
    const uint32_t total = s_total;
    const unsigned long long tag = ctl.gen << 34;
    uint32_t excl = 0;
    if (tile == 0) {
      if (lane == 0) lb_store(ctl.status, tag | (2ull << 32) | total);
    } else {
      if (lane == 0) lb_store(ctl.status + tile, tag | (1ull << 32) | total);
      excl = lb_lookback_warp(ctl, tile, lane);
      if (lane == 0) lb_store(ctl.status + tile, tag | (2ull << 32) | (excl + total));
    }
    if (lane == 0) {
      s_base = excl;
      s_off[kScanThreads] = total;
      if (total_out && tile == gridDim.x - 1) *total_out = excl + total;
    }
  }
  __syncthreads();
  const uint32_t total = s_total;
  const uint64_t base = s_base, row0 = (uint64_t)tile * kScanThreads;
  for (uint32_t q = tid; q < total; q += kScanThreads) {
    uint32_t lo = 0, hi = kScanThreads;  // last j with s_off[j] <= q  (targets without neighbours share an offset)
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (s_off[mid] <= q) lo = mid; else hi = mid;
    }
    const uint64_t from = s_src[lo] + (q - s_off[lo]), o = base + q;
    m.out.all_nodes[m.T + o] = m.r_nbr[from];
    m.out.all_ts[m.T + o] = m.r_ts[from];
    m.out.dt[o] = m.r_dt[from];
    m.out.eid[o] = m.r_eid[from];
    m.out.row[o] = (int64_t)(row0 + lo);
    if (m.out.col) m.out.col[o] = (int64_t)(m.T + o);
  }
}

// edge dispatch by owner (reference gnnflow/distributed/dispatcher.py:41-100 groups a batch by the partition of its
// source vertex on the host and sends each group over RPC): the rows whose source this rank owns, compacted in place
// order by one look-back scan
struct DispatchIn {
  const int64_t *src;
  const int8_t *table;
  uint64_t table_len;
  uint32_t rank, P;
  __device__ uint32_t operator()(uint64_t i) const { return owner_of_vertex(src[i], table, table_len, P) == (int)rank ? 1u : 0u; }
};
struct DispatchOut {
  const int64_t *src, *dst;
  const float *ts;
  const int64_t *eid;
  int64_t *osrc, *odst;
  float *ots;
  int64_t *oeid;
  __device__ void operator()(uint64_t i, uint32_t off, uint32_t keep) const {
    if (!keep) return;
    osrc[off] = src[i];
    odst[off] = dst[i];
    ots[off] = ts[i];
    oeid[off] = eid[i];
  }
};

}  // namespace gf

GF_EXPORT int gf_dispatch_edges(const int64_t *src, const int64_t *dst, const float *ts, const int64_t *eid, uint64_t n,
                                const int8_t *partition_table, uint64_t table_len, uint32_t rank, uint32_t world,
                                int64_t *out_src, int64_t *out_dst, float *out_ts, int64_t *out_eid, uint64_t *count,
                                void *stream) {
  if (!count || world == 0 || rank >= world) GF_FAIL(GF_EINVAL, "gf_dispatch_edges: bad argument");
  *count = 0;
  if (!n) return GF_OK;
  if (n >= (1ull << 32)) GF_FAIL(GF_EINVAL, "gf_dispatch_edges: batch of %llu edges exceeds 2^32-1", (unsigned long long)n);
  if (!src || !dst || !ts || !eid || !out_src || !out_dst || !out_ts || !out_eid) GF_FAIL(GF_EINVAL, "gf_dispatch_edges: null array");
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
  // ticket | tile status words | total (mapped back to the host below)
  const size_t bytes = 256 + tiles * 8 + 8;
  char *ws = nullptr;
  GF_CUDA(cudaMallocAsync(&ws, bytes, st));
  GF_CUDA(cudaMemsetAsync(ws, 0, bytes, st));
  LookbackCtl ctl = {reinterpret_cast<unsigned int *>(ws), reinterpret_cast<unsigned long long *>(ws + 256), 1ull};
  uint32_t *d_total = reinterpret_cast<uint32_t *>(ws + 256 + tiles * 8);
  DispatchIn in = {src, partition_table, table_len, rank, world};
  DispatchOut out = {src, dst, ts, eid, out_src, out_dst, out_ts, out_eid};
  gf::launch(scan_lookback_kernel<DispatchIn, DispatchOut>, (unsigned)tiles, kScanThreads, 0, st, n, in, out, ctl, d_total);
  GF_CUDA(cudaGetLastError());
  uint32_t h_total = 0;
  GF_CUDA(cudaMemcpyAsync(&h_total, d_total, 4, cudaMemcpyDeviceToHost, st));
  GF_CUDA(cudaFreeAsync(ws, st));
  GF_CUDA(cudaStreamSynchronize(st));
  *count = h_total;
  return GF_OK;
}

struct gf_peer {
  int device = 0;
  uint32_t rank = 0, world = 1;
  uint64_t cap = 0;
  uint32_t max_fanout = 0;
  gf::WinLayout L;
  char *window = nullptr;
  char *win[gf::kMaxPeers] = {nullptr};
  bool connected = false;
  unsigned long long gen = 0;
  gf::Scratch scratch;  // block counts, totals, owner / pos per target, done counters, overflow flag
  unsigned grid = 0;
};

GF_EXPORT int gf_peer_create(int device, uint32_t rank, uint32_t world, uint64_t max_targets, uint32_t max_fanout,
                             gf_peer **out) {
  if (!out || world == 0 || world > kMaxPeers || rank >= world || max_targets == 0 || max_fanout == 0)
    GF_FAIL(GF_EINVAL, "gf_peer_create: bad argument (world must be 1..%u)", kMaxPeers);
  if (max_targets >= (1ull << 31)) GF_FAIL(GF_EINVAL, "gf_peer_create: max_targets too large");
  GF_CUDA(cudaSetDevice(device));
  gf_peer *p = new gf_peer();
  p->device = device;
  p->rank = rank;
  p->world = world;
  p->cap = max_targets;
  p->max_fanout = max_fanout;
  p->L = make_layout(world, max_targets, max_fanout);
  cudaError_t e = cudaMalloc(&p->window, p->L.total);
  if (e == cudaSuccess) e = cudaMemset(p->window, 0, p->L.total);
  if (e != cudaSuccess) {
    cudaGetLastError();
    const unsigned long long want = p->L.total;
    delete p;
    GF_FAIL(GF_ENOMEM, "gf_peer_create: exchange window of %llu bytes: %s", want, cudaGetErrorString(e));
  }
  p->win[rank] = p->window;
  p->connected = world == 1;
  *out = p;
  return GF_OK;
}

GF_EXPORT int gf_peer_export(gf_peer *p, void *handle_out) {
  if (!p || !handle_out) GF_FAIL(GF_EINVAL, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == GF_PEER_HANDLE_BYTES, "handle size");
  GF_CUDA(cudaSetDevice(p->device));
  cudaIpcMemHandle_t h;
  GF_CUDA(cudaIpcGetMemHandle(&h, p->window));
  memcpy(handle_out, &h, sizeof(h));
  return GF_OK;
}

GF_EXPORT int gf_peer_connect(gf_peer *p, const void *handles) {
  if (!p || !handles) GF_FAIL(GF_EINVAL, "null argument");
  GF_CUDA(cudaSetDevice(p->device));
  for (uint32_t r = 0; r < p->world; r++) {
    if (r == p->rank || p->win[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + (size_t)r * GF_PEER_HANDLE_BYTES, sizeof(h));
    void *ptr = nullptr;
    GF_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->win[r] = (char *)ptr;
  }
  p->connected = true;
  return GF_OK;
}

GF_EXPORT int gf_peer_destroy(gf_peer *p) {
  if (!p) return GF_OK;
  cudaSetDevice(p->device);
  cudaDeviceSynchronize();
  for (uint32_t r = 0; r < p->world; r++)
    if (r != p->rank && p->win[r]) cudaIpcCloseMemHandle(p->win[r]);
  if (p->window) cudaFree(p->window);
  p->scratch.release();
  delete p;
  return GF_OK;
}

GF_EXPORT int gf_sampler_sample_layer_partitioned(gf_sampler *s, gf_peer *pr, const int64_t *nodes, const float *timestamps,
                                                  uint64_t T, const int8_t *partition_table, uint64_t table_len,
                                                  uint32_t layer, uint32_t snapshot, gf_sampling_result *result,
                                                  void *stream) {
  if (!s || !pr || !result) GF_FAIL(GF_EINVAL, "null argument");
  if (!pr->connected) GF_FAIL(GF_EINVAL, "gf_peer_connect has not been called");
  if (layer >= s->fanouts.size() || snapshot >= s->num_snapshots) GF_FAIL(GF_EINVAL, "layer/snapshot out of range");
  const uint32_t F = s->fanouts[layer];
  if (F > pr->max_fanout) GF_FAIL(GF_ECAPACITY, "fanout %u exceeds the window's max_fanout %u", F, pr->max_fanout);
  if (T > pr->cap) GF_FAIL(GF_ECAPACITY, "%llu targets exceed the window's max_targets %llu", (unsigned long long)T, (unsigned long long)pr->cap);
  if (T && (!nodes || !timestamps)) GF_FAIL(GF_EINVAL, "null input");
  if (result->capacity_dst < T) GF_FAIL(GF_ECAPACITY, "result.capacity_dst too small");
  if (T && (!result->all_nodes || !result->all_timestamps || !result->delta_timestamps || !result->eids || !result->row))
    GF_FAIL(GF_EINVAL, "null output array");
  if (pr->device != s->graph->cfg.device) GF_FAIL(GF_EINVAL, "peer window and graph live on different devices");
  cudaStream_t st = (cudaStream_t)stream;
  gf_graph *g = s->graph;
  GF_TRY(gf_graph_flush_internal(g));  // batches queued by gf_graph_add_edges_async
  std::lock_guard<std::mutex> lk(g->mu);
  GF_CUDA(cudaSetDevice(g->cfg.device));
  const unsigned long long gen = ++pr->gen;
  const uint32_t nblk = (uint32_t)std::max<uint64_t>(1, (T + (uint64_t)kRThreads * kRItems - 1) / ((uint64_t)kRThreads * kRItems));
  // scratch: [done_route, done_sample, overflow, pad] | totals[8] | owner[T] | pos[T] | ticket + status[nblk * 8] (zeroed per step)
  const size_t off_tot = 64, off_own = off_tot + 64, off_pos = off_own + align_up(T * 4, 256),
               off_lb = off_pos + align_up(T * 4, 256), lb_bytes = 256 + align_up((size_t)nblk * kMaxPeers * 4, 256),
               total = off_lb + lb_bytes;
  if (total > pr->scratch.cap) {
    GF_TRY(pr->scratch.reserve(total, st));
    GF_CUDA(cudaMemsetAsync(pr->scratch.ptr, 0, 64, st));  // done counters start at zero
  }
  char *sc = pr->scratch.as<char>();
  unsigned int *done_route = reinterpret_cast<unsigned int *>(sc), *done_sample = done_route + 1;
  uint32_t *overflow = reinterpret_cast<uint32_t *>(sc) + 2;
  uint32_t *totals = reinterpret_cast<uint32_t *>(sc + off_tot);
  int32_t *owner_local = reinterpret_cast<int32_t *>(sc + off_own);
  uint32_t *pos_local = reinterpret_cast<uint32_t *>(sc + off_pos);
  uint32_t *route_ticket = reinterpret_cast<uint32_t *>(sc + off_lb), *route_status = reinterpret_cast<uint32_t *>(sc + off_lb + 256);
  PeerView pv;
  for (uint32_t r = 0; r < kMaxPeers; r++) pv.win[r] = r < pr->world ? pr->win[r] : nullptr;
  pv.L = pr->L;
  pv.L.F = pr->max_fanout;
  pv.rank = pr->rank;
  // ---- route
  s->prof.begin(st);
  GF_CUDA(cudaMemsetAsync(sc + off_lb, 0, lb_bytes, st));
  gf::launch(route_kernel, nblk, kRThreads, 0, st, nodes, timestamps, T, partition_table, table_len, pv, owner_local, pos_local,
             route_ticket, route_status, totals, done_route, gen, overflow);
  s->prof.end(0, st);
  // ---- sample what the ranks asked of this partition
  gf::launch(wait_flags_kernel, 1, 32, 0, st, reinterpret_cast<const unsigned long long *>(pr->window + pr->L.req_flag),
             pr->world, gen);
  SampleParams p = make_params(s, layer, snapshot);
  if (!pr->grid) {
    int occ = 0, sms = 0;
    GF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pr->device));
    GF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sample_partition_kernel, kQThreads, 0));
    pr->grid = (unsigned)std::max(1, occ * sms);
  }
  gf::launch(sample_partition_kernel, pr->grid, kQThreads, 0, st, p, pv, done_sample, gen);
  s->launch_index++;
  s->prof.end(1, st);
  // ---- merge the responses in the original target order
  gf::launch(wait_flags_kernel, 1, 32, 0, st, reinterpret_cast<const unsigned long long *>(pr->window + pr->L.resp_flag),
             pr->world, gen);
  GF_TRY(ensure_h_meta(s, 8));
  s->h_meta[0] = 0;
  if (T) {
    const uint64_t tiles = (T + kScanThreads - 1) / kScanThreads;
    GF_TRY(ensure_fused(s, tiles, st));
    LookbackCtl ctl = {s->fused.as<unsigned int>(), reinterpret_cast<unsigned long long *>(s->fused.as<char>() + 256),
                       s->fused_gen};
    EmitOut eo;
    memset(&eo, 0, sizeof(eo));
    eo.all_nodes = result->all_nodes;
    eo.all_ts = result->all_timestamps;
    eo.dt = result->delta_timestamps;
    eo.eid = result->eids;
    eo.row = result->row;
    eo.col = result->col;
    const char *w = pr->window;
    MergeArgs ma = {nodes, timestamps, owner_local, pos_local, reinterpret_cast<const uint32_t *>(w + pr->L.resp_cnt),
                    reinterpret_cast<const int64_t *>(w + pr->L.resp_nbr), reinterpret_cast<const int64_t *>(w + pr->L.resp_eid),
                    reinterpret_cast<const float *>(w + pr->L.resp_ts), reinterpret_cast<const float *>(w + pr->L.resp_dt),
                    pr->cap, T, pr->max_fanout, eo};
    gf::launch(merge_kernel, (unsigned)tiles, kScanThreads, 0, st, ma, ctl, s->h_meta);
  }
  GF_CUDA(cudaGetLastError());
  s->prof.end(2, st, false);
  uint32_t h_over = 0;
  GF_CUDA(cudaMemcpyAsync(&h_over, overflow, 4, cudaMemcpyDeviceToHost, st));
  GF_CUDA(cudaStreamSynchronize(st));
  if (h_over) GF_FAIL(GF_ECAPACITY, "partitioned sampling: more requests for one owner than the window holds");
  result->num_dst = T;
  result->num_edges = T ? s->h_meta[0] : 0;
  return GF_OK;
}
