// Kernels of the batched edge-insert path (DynamicGraph::AddEdges + AddEdgesForOneNode + TemporalBlockAllocator,
// reference dynamic_graph.cu:77-138,153-287, temporal_block_allocator.cu:83-180, utils.cu:9-63) and of the block
// allocator behind it.  One add_edges attempt is 3 + P launches whatever the batch touches (P = radix passes over the
// source-vertex bits: 2 for tables of <= 65 536 vertices, 3 up to 16.7 M):
//
//   ingest_prep_kernel   validation flags, id ranges, digit histograms of every pass            (reads 28 B / edge)
//   ingest_sort_kernel   x P: stable LSD radix sort by source vertex that CARRIES the 20-byte payload, so that nothing
//                        after it gathers through a permutation; the first pass reads the caller's arrays directly
//   ingest_plan_kernel   segments (one per source vertex), the block-sizing policy of AddEdgesForOneNode per segment,
//                        allocation by size class (free lists first, then the bump pointer), accept / reject
//   ingest_apply_kernel  one thread per edge: payload append (+ pivots); the thread of a segment's first edge also
//                        applies the segment's plan to the vertex entry / directory; bookkeeping of vertex flags and
//                        edge-id reference counts; cleans the control words for the next call; the last CTA reports to
//                        the host through mapped pinned memory
//
// A batch that is not in time order is first reordered by timestamp (the generic key/value radix sort of
// gf_primitives.cuh + one gather), then takes the same path.
#pragma once
#include <cfloat>
#include <cstdlib>

#include "gf_primitives.cuh"
#include "gf_store.cuh"

namespace gf {

constexpr int kThreads = 256;
constexpr unsigned kNoteEidsNotIncreasing = 1u << 30;  // prep pass, thread-local flag bit (never reaches error_flags)

// CTA-wide sums of K u64 values (kThreads threads, all of them must call); thread 0 ends up with the totals.
// Counters shared by the whole graph get ONE atomic per CTA: per-warp atomics on a single address serialise in L2.
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

struct StoreParams {
  uint32_t min_block;
  int policy;
  int adaptive;
};
// Replace policy (TemporalBlockAllocator::Reallocate, temporal_block_allocator.cu:122-132): the reference re-allocates
// a vertex's single block to EXACTLY size + n on every append, so its capacity is always max(size, minimum_block_size).
// That value is what the getters report (`logical` capacity); physically the block is given a quarter as much again, so
// that a hot vertex is copied O(log n) times instead of once per batch and the blocks it leaves behind fall into the
// same size classes as everybody else's (exact sizes would never be asked for again: O(n^2) dead memory).
__host__ __device__ inline uint32_t replace_logical_cap(uint32_t size, uint32_t min_block) { return size > min_block ? size : min_block; }
__host__ __device__ inline uint32_t replace_physical_cap(uint32_t logical) {
  const uint64_t c = (uint64_t)logical + logical / 4;
  return c > 0x7fffffffull ? 0x7fffffffu : (uint32_t)c;
}

__device__ __forceinline__ unsigned int ld_flag(const unsigned int *p) { return *reinterpret_cast<const volatile unsigned int *>(p); }

__device__ __forceinline__ uint32_t next_pow2_u32(uint32_t n) {  // dynamic_graph.cu:202-204
  return n <= 1 ? 1u : 1u << (32 - __clz(n - 1));
}

// ------------------------------------------------------------------------------------------------ prep
// Pass 0 over the batch: validation flags, id ranges, the digit histograms of all sort passes (keys = low 32 bits of the
// source vertex); clears the CallScratch slot of the next call.  Grid-stride: a CTA adds its histograms to the global
// ones once.
__device__ __forceinline__ void ingest_prep_body(const int64_t *__restrict__ src, const int64_t *__restrict__ dst,
                                                 const float *__restrict__ ts, const int64_t *__restrict__ eid, uint64_t n,
                                                 uint64_t table_cap, long long eid_base, uint64_t eid_cap, int assume_sorted,
                                                 int passes, uint32_t *ghist, CallScratch *cur, CallScratch *nxt) {
  __shared__ uint32_t hist[kSortMaxPasses][256];
  for (int i = threadIdx.x; i < kSortMaxPasses * 256; i += kThreads) (&hist[0][0])[i] = 0;
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    nxt->max_id = nxt->max_eid = nxt->max_neg_eid = 0;
    nxt->error_flags = nxt->num_segments = nxt->total_units = nxt->accepted = nxt->unsorted = nxt->done_ctas = 0;
    nxt->eids_not_increasing = 0;
  }
  long long mx = 0, emx = 0, enmx = 0;  // enmx = max(LLONG_MAX - eid): zero is its identity, LLONG_MAX - enmx the minimum eid
  unsigned flags = 0;
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const long long s = src[i], d = dst[i], e = eid[i];
    const long long m = max(s, d);
    mx = max(mx, m);
    emx = max(emx, e);
    if (s < 0 || d < 0 || m >= (1ll << 32)) flags |= kErrBadId;
    else if ((uint64_t)m >= table_cap) flags |= kErrTableSmall;
    // the reference counts live in a dense table over [eid_base, eid_base + eid_cap): the host moves / grows it and replays
    if (e < 0) flags |= kErrBadEid;
    else {
      enmx = max(enmx, 0x7fffffffffffffffll - e);
      if (e < eid_base) flags |= kErrEidLow;
      else if ((uint64_t)(e - eid_base) >= eid_cap) flags |= kErrEidSmall;
    }
    if (i + 1 < n && ts[i + 1] < ts[i]) flags |= kErrUnsorted;
    if (i + 1 < n && eid[i + 1] <= e) flags |= kNoteEidsNotIncreasing;
    const uint32_t key = (uint32_t)s;
    for (int p = 0; p < passes; p++) atomicAdd(&hist[p][(key >> (8 * p)) & 255u], 1u);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    emx = max(emx, __shfl_xor_sync(0xffffffffu, emx, o));
    enmx = max(enmx, __shfl_xor_sync(0xffffffffu, enmx, o));
    flags |= __shfl_xor_sync(0xffffffffu, flags, o);
  }
  __shared__ long long s_mx[kThreads / 32], s_emx[kThreads / 32], s_enmx[kThreads / 32];
  __shared__ unsigned s_flags[kThreads / 32];
  if ((threadIdx.x & 31) == 0) {
    s_mx[threadIdx.x >> 5] = mx;
    s_emx[threadIdx.x >> 5] = emx;
    s_enmx[threadIdx.x >> 5] = enmx;
    s_flags[threadIdx.x >> 5] = flags;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < kThreads / 32; w++) {
      mx = max(mx, s_mx[w]);
      emx = max(emx, s_emx[w]);
      enmx = max(enmx, s_enmx[w]);
      flags |= s_flags[w];
    }
    // fire-and-forget reductions (no value comes back: nothing here waits for an L2 round trip)
    if (mx) atomicMax(&cur->max_id, mx);
    if (emx) atomicMax(&cur->max_eid, emx);
    if (enmx) atomicMax(&cur->max_neg_eid, enmx);
    if (flags & kNoteEidsNotIncreasing) {  // a note for the reference-count upkeep, not an error
      cur->eids_not_increasing = 1;
      flags &= ~kNoteEidsNotIncreasing;
    }
    if (flags & kErrUnsorted) {
      cur->unsorted = 1;
      if (!assume_sorted) flags &= ~kErrUnsorted;  // the caller has already put the batch in time order
    }
    if (flags) atomicOr(&cur->error_flags, flags);
  }
  for (int p = 0; p < passes; p++) {
    const uint32_t c = hist[p][threadIdx.x];
    if (c) atomicAdd(&ghist[p * 256 + threadIdx.x], c);
  }
}
__global__ void __launch_bounds__(kThreads) ingest_prep_kernel(const int64_t *__restrict__ src, const int64_t *__restrict__ dst,
                                                               const float *__restrict__ ts, const int64_t *__restrict__ eid,
                                                               uint64_t n, uint64_t table_cap, long long eid_base,
                                                               uint64_t eid_cap, int assume_sorted, int passes,
                                                               uint32_t *ghist, CallScratch *cur, CallScratch *nxt) {
  pdl_trigger();  // the sort pass may be scheduled already; it waits for this grid before it reads anything
  ingest_prep_body(src, dst, ts, eid, n, table_cap, eid_base, eid_cap, assume_sorted, passes, ghist, cur, nxt);
}

// ------------------------------------------------------------------------------------------------ sort
// One pass of the stable LSD radix sort by source vertex, 8 bits, "onesweep" (see gf_primitives.cuh for the scheme), moving
// {key, ts, dst, eid} = 24 bytes per edge: 24 B read + 24 B written per pass and edge, all of it in full lines -- the tile
// is reordered in shared memory and leaves it digit run by digit run.  FIRST: reads the caller's src / dst / ts / eid.
struct SortSrc {
  const int64_t *src;   // FIRST pass: the key is the low 32 bits of src
  const uint32_t *key;  // later passes
  const float *ts;
  const int64_t *dst, *eid;
};
struct SortDst {
  uint32_t *key;
  float *ts;
  int64_t *dst, *eid;
};
constexpr int kIngestRoundsSmall = 4, kIngestRoundsBig = 8;  // tile = 1024 / 2048 edges (24 / 48 KB of shared memory)
constexpr uint64_t kIngestBigN = 1u << 19;                    // below this, small tiles spread the batch over more SMs
inline int ingest_rounds(uint64_t n) {
  static const bool small_only = getenv("GNNFLOW_B200_SORT_SMALL_TILES") != nullptr;  // experiment knob
  return n < kIngestBigN || small_only ? kIngestRoundsSmall : kIngestRoundsBig;
}
inline uint32_t ingest_sort_tiles(uint64_t n) {
  const uint64_t tile = (uint64_t)kSortThreads * ingest_rounds(n);
  return (uint32_t)((n + tile - 1) / tile);
}

// COHERENT: read the inputs through the L2 only (for callers that run several passes inside one launch)
template <int ROUNDS, bool FIRST, bool COHERENT>
__device__ __forceinline__ void ingest_sort_tile(const SortSrc &in, const SortDst &out, uint64_t n, int shift,
                                                 const uint32_t *ghist, uint32_t *status, uint32_t tile, uint8_t *s_dyn) {
  constexpr int TILE = kSortThreads * ROUNDS;
  int64_t *sd = reinterpret_cast<int64_t *>(s_dyn), *se = sd + TILE;
  uint32_t *sk = reinterpret_cast<uint32_t *>(se + TILE);
  float *st = reinterpret_cast<float *>(sk + TILE);
  __shared__ uint32_t cnt[kSortWarps][256];  // per-warp digit counts -> exclusive prefix over warps
  __shared__ uint32_t dstart[256];           // first local slot of each digit in the reordered tile
  __shared__ uint32_t gbase[256];            // global position of local slot e with digit d = gbase[d] + e
  __shared__ uint32_t s_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const uint64_t tile_base = (uint64_t)tile * TILE;
  const uint32_t tile_n = (uint32_t)min((uint64_t)TILE, n - tile_base);
  const uint64_t base = tile_base + (uint64_t)w * (32 * ROUNDS) + lane;
  uint32_t k[ROUNDS];
  float t[ROUNDS];
  int64_t d[ROUNDS], e[ROUNDS];
  uint16_t rk[ROUNDS];  // stable rank among the same digit inside this warp's strip
#pragma unroll
  for (int r = 0; r < ROUNDS; r++) {
    const uint64_t i = base + r * 32;
    const bool valid = i < n;
    if (COHERENT) {
      k[r] = valid ? __ldcg(in.key + i) : 0u;
      t[r] = valid ? __ldcg(in.ts + i) : 0.f;
      d[r] = valid ? __ldcg(reinterpret_cast<const long long *>(in.dst) + i) : 0;
      e[r] = valid ? __ldcg(reinterpret_cast<const long long *>(in.eid) + i) : 0;
    } else {
      k[r] = valid ? (FIRST ? (uint32_t)__ldg(in.src + i) : __ldg(in.key + i)) : 0u;
      t[r] = valid ? __ldg(in.ts + i) : 0.f;
      d[r] = valid ? __ldg(in.dst + i) : 0;
      e[r] = valid ? __ldg(in.eid + i) : 0;
    }
  }
#pragma unroll
  for (int r = 0; r < ROUNDS; r++) {
    const bool valid = base + r * 32 < n;
    const uint32_t dg = (k[r] >> shift) & 255u;
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? dg : (0x100u | lane));
    const uint32_t before = cnt[w][dg];
    __syncwarp();
    if (valid && (peers & lt_mask) == 0) cnt[w][dg] = before + __popc(peers);
    __syncwarp();
    rk[r] = (uint16_t)(before + __popc(peers & lt_mask));
  }
  __syncthreads();
  // thread dg owns digit dg from here on
  const uint32_t dg = threadIdx.x;
  uint32_t mine = 0;
#pragma unroll
  for (int ww = 0; ww < kSortWarps; ww++) {
    const uint32_t c = cnt[ww][dg];
    cnt[ww][dg] = mine;
    mine += c;
  }
  uint32_t *my_status = status + (uint64_t)tile * 256 + dg;
  os_store(my_status, (tile == 0 ? kOsIncl : kOsAgg) | mine);
  const uint32_t local_start = block_excl_scan(mine, &s_total);
  dstart[dg] = local_start;
  const uint32_t digit_base = block_excl_scan(__ldcg(ghist + dg), &s_total);  // global start of digit dg
  __syncthreads();
  // reorder the tile in shared memory
#pragma unroll
  for (int r = 0; r < ROUNDS; r++) {
    if (base + r * 32 < n) {
      const uint32_t dd = (k[r] >> shift) & 255u;
      const uint32_t pos = dstart[dd] + cnt[w][dd] + rk[r];
      sk[pos] = k[r];
      st[pos] = t[r];
      sd[pos] = d[r];
      se[pos] = e[r];
    }
  }
  // per-digit look-back over the earlier tiles
  uint32_t excl = 0;
  if (tile > 0) excl = os_lookback<OsWindow<ROUNDS>::value>(my_status, tile, mine);
  gbase[dg] = digit_base + excl - local_start;
  __syncthreads();
  for (uint32_t x = threadIdx.x; x < tile_n; x += kSortThreads) {
    const uint32_t key = sk[x];
    const uint32_t pos = gbase[(key >> shift) & 255u] + x;
    out.key[pos] = key;
    out.ts[pos] = st[x];
    out.dst[pos] = sd[x];
    out.eid[pos] = se[x];
  }
}
template <int ROUNDS, bool FIRST>
__global__ void __launch_bounds__(kSortThreads) ingest_sort_kernel(SortSrc in, SortDst out, uint64_t n, int shift,
                                                                   const uint32_t *__restrict__ ghist, uint32_t *ticket,
                                                                   uint32_t *status) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  __shared__ uint32_t s_tile;
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  ingest_sort_tile<ROUNDS, FIRST, false>(in, out, n, shift, ghist, status, s_tile, s_dyn);
}

// out[i] = in[perm[i]] for the four arrays of a batch (slow path: the batch was not in time order)
__global__ void __launch_bounds__(kThreads) ingest_permute_kernel(const uint32_t *__restrict__ perm, uint64_t n,
                                                                  const int64_t *__restrict__ src, const int64_t *__restrict__ dst,
                                                                  const float *__restrict__ ts, const int64_t *__restrict__ eid,
                                                                  int64_t *osrc, int64_t *odst, float *ots, int64_t *oeid) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t j = perm[i];
  osrc[i] = src[j];
  odst[i] = dst[j];
  ots[i] = ts[j];
  oeid[i] = eid[j];
}
__global__ void keys_from_ts_kernel(const float *__restrict__ ts, uint64_t n, uint32_t *keys, uint32_t *vals) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t = ts[i];
  keys[i] = t == 0.0f ? orderable_f32(0.0f) : orderable_f32(t);  // -0.0 == +0.0 under operator<
  vals[i] = (uint32_t)i;
}

// ------------------------------------------------------------------------------------------------ plan
enum : uint32_t { kPlanNew = 1u, kPlanRealloc = 2u, kPlanDir = 4u };
struct __align__(16) SegRec {  // one per segment (= source vertex of the batch): written by plan, read by apply
  uint32_t v;       // vertex
  uint32_t start;   // first edge of the segment in the sorted batch
  uint32_t cnt;     // edges
  uint32_t fill;    // edges appended to the existing newest block
  uint64_t p0;      // newest block before the batch: payload, capacity, size
  uint32_t cap0, off0;
  uint32_t newcap;  // capacity of the block to allocate (0 = none)
  uint32_t flags;   // kPlan* | payload class << 8 | directory class << 16
  uint32_t prank;   // rank of the payload request among the batch's requests of its class
  uint32_t drank;   // ... of the directory request
};
static_assert(sizeof(SegRec) == 48, "SegRec is three 16-byte loads");

constexpr int kPlanEpt = 4;      // consecutive edges per thread of the plan kernel; tile = kThreads * edges per thread
constexpr int kPlanEptBig = 16;  // batches of 2^20 edges and more: a tile's fixed cost (ticket, two scans, look-back, the class
                                 // atomics, the arrival) is paid a quarter as often (9 341 tiles of 12 us each in four-deep
                                 // waves were 188 us of a 9.56 M-edge batch)
constexpr int kPlanTile = kThreads * kPlanEpt;
constexpr unsigned long long kPlAgg = 1ull << 62, kPlIncl = 2ull << 62;  // status A: flag | heads << 31 | last head + 1

struct PlanArgs {
  const uint32_t *keys;  // sorted
  const float *ts;       // sorted along
  uint64_t n;
  const NodeEntry *table;
  StoreParams sp;
  uint32_t *segid;  // [n]
  SegRec *recs;     // [segments]
  GraphStats *stats;
  CallScratch *cur;
  CallClasses *cls;
  uint32_t *ticket;
  unsigned long long *stat_a;  // [tiles]
  uint32_t *gcls;              // [kNumClasses] requests per size class so far | [kNumClasses] tiles done
  int async;
};

// exclusive max-scan over the 256 threads of the CTA (values are non-decreasing where non-zero); *all = max of all
__device__ __forceinline__ uint32_t block_excl_max_scan(uint32_t v, uint32_t *all) {
  __shared__ uint32_t wmax[kThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl = max(incl, t);
  }
  if (lane == 31) wmax[w] = incl;
  __syncthreads();
  uint32_t before = 0;
#pragma unroll
  for (int ww = 0; ww < kThreads / 32; ww++)
    if (ww < w) before = max(before, wmax[ww]);
  if (threadIdx.x == kThreads - 1) *all = max(before, incl);
  uint32_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0;
  __syncthreads();
  return max(before, excl);
}

// one tile of the plan: segments, policy, ranks of the allocation requests (nothing is mutated but the per-call scratch)
template <bool COHERENT, int EPT>
__device__ __forceinline__ void ingest_plan_tile(const PlanArgs &a, uint32_t tile, uint32_t ntiles) {
  static_assert(EPT % 4 == 0 && EPT <= 16, "segment ids leave in groups of four; request words hold 13 bits of segment");
  constexpr int TILE = kThreads * EPT;
  constexpr bool LISTED = EPT > 4;  // the tile's allocation requests go through a list in shared memory, not registers
  // keys of the tile and one neighbour on either side.  A thread reads EPT consecutive keys: with 16 per thread the plain
  // layout puts the 32 threads of a warp on two banks (28 % of the kernel's stall samples), so one word of padding
  // follows every 16 keys
  constexpr int SKPAD = LISTED ? (TILE + 2 + 15) / 16 : 0;
  __shared__ uint32_t sk_[TILE + 2 + SKPAD];
  auto sk = [&](uint32_t j) -> uint32_t & { return sk_[LISTED ? j + (j >> 4) : j]; };
  __shared__ uint32_t s_req[LISTED ? TILE : 1];  // segment - sid_base | (payload class + 1) << 13 | (directory class + 1) << 21
  __shared__ uint32_t cls_cnt[kNumClasses], cls_excl[kNumClasses];
  __shared__ uint32_t s_total, s_last1, s_excl_heads, s_prev_last1, s_nreq;
  const int tid = threadIdx.x, lane = tid & 31;
  CallScratch *cur = a.cur;
  for (int c = tid; c < (int)kNumClasses; c += kThreads) cls_cnt[c] = 0;
  if (tid == 0) s_nreq = 0;
  const uint64_t n = a.n, base = (uint64_t)tile * TILE;
  for (int j = tid; j < TILE + 2; j += kThreads) {
    const int64_t gi = (int64_t)base - 1 + j;
    sk(j) = (gi >= 0 && (uint64_t)gi < n) ? (COHERENT ? __ldcg(a.keys + gi) : __ldg(a.keys + gi)) : 0u;
  }
  __syncthreads();
  // ---- heads / tails of the segments among this thread's EPT consecutive edges
  const uint32_t j0 = tid * EPT;
  unsigned heads = 0, tails = 0, valid = 0;
  uint32_t last1 = 0;  // index + 1 of this thread's last head
#pragma unroll
  for (int k = 0; k < EPT; k++) {
    const uint64_t i = base + j0 + k;
    if (i >= n) break;
    valid |= 1u << k;
    if (i == 0 || sk(j0 + k + 1) != sk(j0 + k)) {
      heads |= 1u << k;
      last1 = (uint32_t)i + 1;
    }
    if (i == n - 1 || sk(j0 + k + 1) != sk(j0 + k + 2)) tails |= 1u << k;
  }
  const uint32_t heads_before = block_excl_scan(__popc(heads), &s_total);
  const uint32_t last1_before = block_excl_max_scan(last1, &s_last1);
  // ---- look-back A over the earlier tiles: (segment heads so far, position of the last head)
  if (tid < 32) {
    const uint32_t total = s_total, tl1 = s_last1;
    uint32_t eh = 0, pl = 0;
    if (tile == 0) {
      if (lane == 0) lb_store(a.stat_a, kPlIncl | ((unsigned long long)total << 31) | tl1);
    } else {
      if (lane == 0) lb_store(a.stat_a + tile, kPlAgg | ((unsigned long long)total << 31) | tl1);
      int64_t q0 = (int64_t)tile - 1;
      while (true) {
        const int64_t q = q0 - lane;
        const unsigned long long w = q >= 0 ? lb_load(a.stat_a + q) : kPlIncl;  // before the first tile: nothing
        const unsigned flag = (unsigned)(w >> 62);
        const unsigned incl = __ballot_sync(0xffffffffu, flag == 2u);
        const unsigned nready = __ballot_sync(0xffffffffu, flag == 0u);
        const int stop = incl ? __ffs(incl) - 1 : 31;  // nearest inclusive predecessor, or the whole window
        const unsigned need = stop == 31 ? 0xffffffffu : ((2u << stop) - 1u);
        if (nready & need) continue;  // a needed predecessor has not published yet: poll again
        const bool use = lane <= stop;
        eh += __reduce_add_sync(0xffffffffu, use ? (uint32_t)((w >> 31) & 0x7fffffffull) : 0u);
        pl = max(pl, __reduce_max_sync(0xffffffffu, use ? (uint32_t)(w & 0x7fffffffull) : 0u));
        if (incl) break;
        q0 -= 32;
      }
      if (lane == 0) lb_store(a.stat_a + tile, kPlIncl | ((unsigned long long)(eh + total) << 31) | max(pl, tl1));
    }
    if (lane == 0) {
      s_excl_heads = eh;
      s_prev_last1 = pl;
    }
  }
  __syncthreads();
  // ---- segment ids; the thread that holds a segment's LAST edge knows its extent and plans it
  uint32_t incl_heads = s_excl_heads + heads_before;
  uint32_t cur_last1 = max(s_prev_last1, last1_before);
  const uint32_t sid_base = s_excl_heads - 1;  // a tile's first edges may close the segment before its first head (wraps at 0)
  uint32_t sid4[4] = {0, 0, 0, 0};
  uint32_t cls4[4] = {0, 0, 0, 0};  // per tail: payload class + 1 | (directory class + 1) << 16
  uint32_t prk4[4] = {0, 0, 0, 0}, drk4[4] = {0, 0, 0, 0};  // ... and the ranks of its requests within the tile
#pragma unroll
  for (int k = 0; k < EPT; k++) {
    const int k4 = k & 3;
    if (valid & (1u << k)) {
    const uint32_t i = (uint32_t)(base + j0 + k);
    if (heads & (1u << k)) {
      incl_heads++;
      cur_last1 = i + 1;
    }
    const uint32_t sid = incl_heads - 1;
    sid4[k4] = sid;
    if (tails & (1u << k)) {
    // DynamicGraph::AddEdgesForOneNode (dynamic_graph.cu:206-287) + TemporalBlockAllocator::AlignUp (:83-88); nothing
    // is mutated here
    const uint32_t start = cur_last1 - 1, cnt = i - start + 1, v = sk(j0 + k + 1);
    const NodeEntry ent = load_entry64(a.table + v);
    const bool live = ent.end > ent.first;
    const float first_ts = COHERENT ? __ldcg(a.ts + start) : __ldg(a.ts + start);
    SegRec r;
    r.v = v; r.start = start; r.cnt = cnt; r.fill = 0; r.p0 = 0; r.cap0 = 0; r.off0 = 0; r.newcap = 0; r.flags = 0;
    r.prank = 0; r.drank = 0;
    if (!live) {
      r.newcap = max(cnt, a.sp.min_block);
      if (a.sp.policy == GF_INSERTION_REPLACE) r.newcap = replace_physical_cap(r.newcap);
      r.flags = kPlanNew;
    } else {
      const BlockDesc t = ent.tail;
      if (first_ts < t.end_ts) atomicOr(&cur->error_flags, kErrOutOfOrder);
      r.p0 = t.payload; r.cap0 = t.capacity; r.off0 = t.size;
      if ((uint64_t)t.size + cnt > t.capacity) {
        if (a.sp.policy == GF_INSERTION_INSERT) {
          r.fill = t.capacity - t.size;
          const uint32_t rem = cnt - r.fill;
          const uint64_t avg = ent.num_insertions == 0 ? rem : ent.num_edges / ent.num_insertions;
          const uint32_t ns = a.sp.adaptive ? next_pow2_u32((uint32_t)max((uint64_t)rem, avg)) : rem;
          r.newcap = max(ns, a.sp.min_block);
          r.flags = kPlanNew;
        } else {
          r.newcap = replace_physical_cap(replace_logical_cap(t.size + cnt, a.sp.min_block));
          r.flags = kPlanRealloc;
        }
      } else {
        r.fill = cnt;
      }
    }
    uint32_t pc = 0, dc = 0;
    if (r.newcap) {
      pc = class_of_units(payload_units(r.newcap)) + 1;
      r.prank = atomicAdd(&cls_cnt[pc - 1], 1u);
    }
    if ((r.flags & kPlanNew) && ent.end == ent.dir_cap()) {  // the directory is full (or there is none yet)
      const uint32_t nlive = ent.end - ent.first;
      const uint32_t dcap = max(4u, next_pow2_u32(2 * (nlive + 1)));
      dc = class_of_units(dir_units(dcap)) + 1;
      r.flags |= kPlanDir;
      r.drank = atomicAdd(&cls_cnt[dc - 1], 1u);
    }
    r.flags |= (pc ? (pc - 1) << 8 : 0u) | (dc ? (dc - 1) << 16 : 0u);
    if (LISTED) {
      if (pc | dc) s_req[atomicAdd(&s_nreq, 1u)] = (sid - sid_base) | (pc << 13) | (dc << 21);
    } else {
      cls4[k4] = pc | (dc << 16);
      prk4[k4] = r.prank;
      drk4[k4] = r.drank;
    }
    a.recs[sid] = r;
    }  // tail
    }  // valid
    if (k4 == 3) {  // segment ids of four edges: one 16-byte store
      const unsigned v4 = (valid >> (k - 3)) & 0xfu;
      if (v4 == 0xfu) {
        *reinterpret_cast<uint4 *>(a.segid + base + j0 + (k - 3)) = make_uint4(sid4[0], sid4[1], sid4[2], sid4[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (v4 & (1u << q)) a.segid[base + j0 + (k - 3) + q] = sid4[q];
      }
    }
  }
  __syncthreads();
  // ---- ranks of the tile's requests among the batch's requests of their class: ONE atomic per (tile, class in use).
  //      (The order in which tiles arrive only decides which address a block gets, never what the store contains; a
  //      per-class look-back over the tiles here cost 17 of the kernel's 20 us at 100 000-edge batches.)
  if (tid < (int)kNumClasses) {
    const uint32_t c = cls_cnt[tid];
    cls_excl[tid] = c ? atomicAdd(&a.gcls[tid], c) : 0u;
  }
  __syncthreads();
  if (LISTED) {  // ranks within the tile -> ranks within the batch, request by request (the records were written by this CTA)
    const uint32_t nreq = s_nreq;
    for (uint32_t q = tid; q < nreq; q += kThreads) {
      const uint32_t w = s_req[q], sid = sid_base + (w & 0x1fffu), pc = (w >> 13) & 0xffu, dc = w >> 21;
      if (pc) a.recs[sid].prank += cls_excl[pc - 1];
      if (dc) a.recs[sid].drank += cls_excl[dc - 1];
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t pc = cls4[k] & 0xffffu, dc = cls4[k] >> 16;
      if (pc) a.recs[sid4[k]].prank = prk4[k] + cls_excl[pc - 1];
      if (dc) a.recs[sid4[k]].drank = drk4[k] + cls_excl[dc - 1];
    }
  }
  if (tile == ntiles - 1 && tid == 0) cur->num_segments = s_excl_heads + s_total;
}

// after every tile of the plan: pop the free lists, lay out the bump region, accept or reject the batch (one CTA)
__device__ __forceinline__ void ingest_plan_finalize(const PlanArgs &a) {
  const int tid = threadIdx.x, lane = tid & 31;
  CallScratch *cur = a.cur;
  const uint32_t total_c = tid < (int)kNumClasses ? *(volatile uint32_t *)&a.gcls[tid] : 0u;  // requests of class tid in the batch
  ArenaState *ar = &a.stats->arena;
  uint32_t take = 0, have = 0, fbase = 0;
  if (tid < (int)kNumClasses) {
    have = ar->free_cnt[tid];
    fbase = ar->free_base[tid];
    take = min(total_c, have);
  }
  const uint32_t cu = tid < (int)kNumClasses ? class_units(tid) : 0u;
  const unsigned long long bump_units_c = (unsigned long long)(total_c - take) * cu;
  // exclusive scan of the bump units over the classes (u64: a batch may need more than 2^32 units in theory)
  __shared__ unsigned long long s_wsum[kThreads / 32];
  __shared__ unsigned int s_accept;
  unsigned long long incl = bump_units_c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_wsum[tid >> 5] = incl;
  __syncthreads();
  unsigned long long bump_before = incl - bump_units_c, bump_total = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; w++) {
    if (w < (tid >> 5)) bump_before += s_wsum[w];
    bump_total += s_wsum[w];
  }
  if (tid == 0) {
    const unsigned flags = *(volatile unsigned int *)&cur->error_flags;
    bool rejected = flags != 0 || (a.async && a.stats->poison);
    if (!rejected) {  // the bump region of this batch: the current one if it has room, else the first that has
      const unsigned long long need = bump_total * kUnit;
      unsigned int r = ar->cur_region;
      if (bump_total >= (1ull << 32) || r >= ar->num_regions || ar->regions[r].cur + need > ar->regions[r].end) {
        r = ar->num_regions;
        if (bump_total < (1ull << 32))
          for (unsigned int k = 0; k < ar->num_regions; k++)
            if (ar->regions[k].cur + need <= ar->regions[k].end) {
              r = k;
              break;
            }
      }
      if (r >= ar->num_regions) {
        atomicOr(&cur->error_flags, kErrArena);
        rejected = true;
      } else {
        ar->cur_region = r;
      }
    }
    cur->total_units = (unsigned int)min(bump_total, 0xffffffffull);
    cur->accepted = rejected ? 0u : 1u;
    if (rejected && a.async) a.stats->poison = 1u;
    s_accept = rejected ? 0u : 1u;
  }
  __syncthreads();
  if (!s_accept) return;
  unsigned long long taken_units[2] = {(unsigned long long)take * cu, (unsigned long long)take};
  if (tid < (int)kNumClasses) {
    a.cls->take[tid] = take;
    a.cls->top[tid] = fbase + have;
    a.cls->bump_base[tid] = (unsigned int)bump_before;
    ar->free_cnt[tid] = have - take;
  }
  block_sum_u64(taken_units);
  const unsigned long long taken_cnt = taken_units[1];
  if (tid == 0) {
    ArenaRegion *reg = &ar->regions[ar->cur_region];
    a.cls->arena_base = reg->cur;
    reg->cur += bump_total * kUnit;
    ar->free_units -= taken_units[0];
    ar->sorted_cnt -= (unsigned int)taken_cnt;
  }
}
// ids may be out of range (flags raised by the prep pass): the table is not touched, the batch changes nothing
__device__ __forceinline__ void ingest_plan_reject(const PlanArgs &a) {
  if (threadIdx.x == 0) {
    a.cur->accepted = 0;
    if (a.async) a.stats->poison = 1u;
  }
}

template <int EPT>
__global__ void __launch_bounds__(kThreads) ingest_plan_kernel(PlanArgs a) {
  __shared__ uint32_t s_tile, s_skip;
  __shared__ unsigned int s_is_last;
  const int tid = threadIdx.x;
  pdl_wait();
  pdl_trigger();
  if (tid == 0) {
    s_tile = atomicAdd(a.ticket, 1u);
    // flags raised by the prep kernel (bad / out-of-range ids, batch not in time order) are final here; the
    // out-of-order flag is being raised by this very kernel and is looked at by the last tile only
    s_skip = (a.cur->error_flags & ~kErrOutOfOrder) != 0;
  }
  __syncthreads();
  const uint32_t tile = s_tile, ntiles = gridDim.x;
  if (s_skip) {
    if (tile == ntiles - 1) ingest_plan_reject(a);
    return;
  }
  ingest_plan_tile<false, EPT>(a, tile, ntiles);
  // ---- the tile that finishes last sees every total
  if (tid == 0) {
    __threadfence();  // the out-of-order flags, the class counts and num_segments travel with the arrival
    s_is_last = atomicAdd(&a.gcls[kNumClasses], 1u) == ntiles - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  ingest_plan_finalize(a);
}

// ------------------------------------------------------------------------------------------------ apply
__device__ __forceinline__ uint64_t class_addr(const CallClasses *cls, const unsigned long long *sorted, uint32_t c,
                                               uint32_t rank) {
  const uint32_t take = cls->take[c];
  if (rank < take) return sorted[cls->top[c] - 1 - rank];
  return cls->arena_base + ((uint64_t)cls->bump_base[c] + (uint64_t)(rank - take) * class_units(c)) * kUnit;
}
__device__ __forceinline__ void free_push(ArenaState *ar, FreeRec *log, uint64_t addr, uint32_t c) {
  const unsigned int k = atomicAdd(&ar->log_cnt, 1u);
  log[k].addr = addr;
  log[k].cls = c;
  atomicAdd(&ar->free_units, (unsigned long long)class_units(c));
}
__device__ __forceinline__ SegRec load_rec(const SegRec *p) {
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  const uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  SegRec r;
  r.v = a.x; r.start = a.y; r.cnt = a.z; r.fill = a.w;
  r.p0 = ((uint64_t)b.y << 32) | b.x; r.cap0 = b.z; r.off0 = b.w;
  r.newcap = c.x; r.flags = c.y; r.prank = c.z; r.drank = c.w;
  return r;
}

struct ApplyArgs {
  const float *ts;  // the batch sorted by (source vertex, time)
  const int64_t *dst, *eid;
  const int64_t *dst_orig, *eid_orig;  // ... and as given (bookkeeping in input order: consecutive eids coalesce)
  uint64_t n;
  const uint32_t *segid;
  const SegRec *recs;
  NodeEntry *table;
  uint8_t *is_src, *is_node;
  uint32_t *eid_ref;  // reference count of edge id e at [e - eid_base]
  long long eid_base;
  GraphStats *stats;
  CallScratch *cur;
  const CallClasses *cls;
  const unsigned long long *sorted;
  FreeRec *log;
  HostResult *hres;
  uint32_t *ctl;  // control words of this call: zeroed here for the next one
  uint64_t ctl_words;
  StoreParams sp;
  int separate_bookkeeping;  // vertex flags / reference counts are kept by ingest_bookkeep_kernel (large batches)
  uint32_t *book_bitmaps;    // ... through one bitmap of book_words words per CTA of that kernel (or null)
  uint32_t book_words;
};

// replace policy only: move the old payload of a reallocated block (TemporalBlockAllocator::Reallocate,
// temporal_block_allocator.cu:122-132 / CopyTemporalBlock, utils.cu:9-31); one CTA per segment at a time
__device__ __forceinline__ void ingest_realloc_copy_body(const SegRec *__restrict__ recs, const CallScratch *cur,
                                                         const CallClasses *cls, const unsigned long long *sorted) {
  if (!ld_flag(&cur->accepted)) return;
  const uint32_t nseg = ld_flag(&cur->num_segments);
  for (uint32_t s = blockIdx.x; s < nseg; s += gridDim.x) {
    const SegRec r = load_rec(recs + s);
    if (!(r.flags & kPlanRealloc)) continue;
    const uint64_t np = class_addr(cls, sorted, (r.flags >> 8) & 0xffu, r.prank);
    const float *ots = blk_ts(r.p0);
    const longlong2 *ode = reinterpret_cast<const longlong2 *>(blk_rec(r.p0, r.cap0));  // records move as 16-byte blobs
    float *nts = const_cast<float *>(blk_ts(np));
    longlong2 *nde = reinterpret_cast<longlong2 *>(const_cast<EdgeRec *>(blk_rec(np, r.newcap)));
    for (uint32_t i = threadIdx.x; i < r.off0; i += kThreads) {
      const float t = ots[i];
      nts[i] = t;
      nde[i] = ode[i];
      blk_store_pivots(np, r.newcap, i, t);  // the new capacity has its own pivot geometry
    }
  }
}
__global__ void __launch_bounds__(kThreads) ingest_realloc_copy_kernel(const SegRec *__restrict__ recs, const CallScratch *cur,
                                                                       const CallClasses *cls,
                                                                       const unsigned long long *sorted) {
  pdl_wait();
  pdl_trigger();
  ingest_realloc_copy_body(recs, cur, cls, sorted);
}

// the i-th edge of the sorted batch; agg += {new blocks, added capacity, first-time edge ids}
__device__ __forceinline__ void ingest_apply_edge(const ApplyArgs &a, uint64_t i, unsigned long long (&agg)[3]) {
  // Everything that does not depend on the segment's plan is requested FIRST: the edge itself, the flag of its destination
  // and the reference count of its id (an atomic whose old value is needed) -- their round trips then run beside the
  // segid -> plan record -> block address chain instead of after it (they were 33 % of the kernel's stall samples at
  // 16 M-edge batches: profiles/r02_c53_ncu_ingest_apply_16k.txt).
  const bool book = !a.separate_bookkeeping;
  const int64_t d_orig = book ? __ldg(a.dst_orig + i) : 0, e_orig = book ? __ldg(a.eid_orig + i) : 0;
  const uint32_t s = __ldg(a.segid + i);
  const float t = __ldg(a.ts + i);
  const int64_t dst_i = __ldg(a.dst + i), eid_i = __ldg(a.eid + i);
  const uint8_t node_seen = book ? *reinterpret_cast<volatile uint8_t *>(a.is_node + d_orig) : (uint8_t)1;
  const bool ref_new = book ? atomicAdd(&a.eid_ref[e_orig - a.eid_base], 1u) == 0 : false;
  {
    const SegRec r = load_rec(a.recs + s);
    const uint32_t k = (uint32_t)i - r.start;
    const uint32_t pcls = (r.flags >> 8) & 0xffu;
    uint64_t np = 0;  // payload of the new (or reallocated) block
    if (r.newcap && (k >= r.fill || k == 0)) np = class_addr(a.cls, a.sorted, pcls, r.prank);
    // ---- payload append (CopyEdgesToBlock, utils.cu:45-57)
    {
      uint64_t p;
      uint32_t cap, pos;
      if (k < r.fill) {
        p = r.p0; cap = r.cap0; pos = r.off0 + k;
      } else {
        p = np; cap = r.newcap; pos = ((r.flags & kPlanRealloc) ? r.off0 : 0u) + (k - r.fill);
      }
      const_cast<float *>(blk_ts(p))[pos] = t;
      blk_store_pivots(p, cap, pos, t);
      // {dst (< 2^32, checked by the prep pass), ts, eid} in one 128-bit store
      const unsigned long long lo = ((unsigned long long)__float_as_uint(t) << 32) | ((unsigned long long)dst_i & 0xffffffffull);
      reinterpret_cast<longlong2 *>(const_cast<EdgeRec *>(blk_rec(p, cap)))[pos] = make_longlong2((long long)lo, eid_i);
    }
    // ---- the segment's first edge applies the plan to the vertex entry and its directory (InsertBlock / Reallocate /
    //      header updates: dynamic_graph.cu:153-174, temporal_block_allocator.cu:122-132, utils.cu:58-62)
    if (k == 0) {
      ArenaState *ar = &a.stats->arena;
      NodeEntry ent = a.table[r.v];
      const float first_ts = a.ts[r.start], last_ts = a.ts[r.start + r.cnt - 1];
      if (r.flags & kPlanDir) {
        const uint32_t dcls = (r.flags >> 16) & 0xffu;
        const uint64_t naddr = class_addr(a.cls, a.sorted, dcls, r.drank);
        BlockDesc *nd = reinterpret_cast<BlockDesc *>(naddr);
        const BlockDesc *od = reinterpret_cast<const BlockDesc *>(ent.dir());
        const uint32_t nlive = ent.end - ent.first;
        // positions (cum_before) are relative to the oldest LIVE block: re-base them while copying, so that they stay
        // bounded by the live range however long the stream runs
        const uint32_t rebase = nlive ? od[ent.first].cum_before : 0u;
        for (uint32_t j = 0; j < nlive; j++) {
          BlockDesc c = od[ent.first + j];
          c.cum_before -= rebase;
          nd[j] = c;
        }
        if (ent.dir_tagged) free_push(ar, a.log, ent.dir(), class_of_units(dir_units(ent.dir_cap())));
        const uint32_t dcap = class_units(dcls) * (kUnit / (uint32_t)sizeof(BlockDesc));
        ent.dir_tagged = naddr | (uint64_t)ilog2_u32(dcap);
        ent.first = 0;
        ent.end = nlive;
        ent.cum_first = 0;
        ent.tail.cum_before -= rebase;
      }
      BlockDesc *dir = reinterpret_cast<BlockDesc *>(ent.dir());
      const bool live = ent.end > ent.first;
      BlockDesc tail = ent.tail;  // == dir[end - 1] when live
      const bool replace = a.sp.policy == GF_INSERTION_REPLACE;
      if (replace)  // what the reference's capacity grows by: max(size, minimum block size) before / after
        agg[1] += (unsigned long long)replace_logical_cap((live ? tail.size : 0u) + r.cnt, a.sp.min_block) -
                  (live ? replace_logical_cap(tail.size, a.sp.min_block) : 0u);
      if (r.fill) {
        tail.size += r.fill;
        tail.start_ts = fminf(tail.start_ts, first_ts);
        tail.end_ts = a.ts[r.start + r.fill - 1];
        dir[ent.end - 1] = tail;
      }
      if (r.flags & kPlanNew) {
        BlockDesc d;
        d.payload = np;
        d.size = r.cnt - r.fill;
        d.capacity = r.newcap;
        d.start_ts = fminf(FLT_MAX, a.ts[r.start + r.fill]);
        d.end_ts = last_ts;
        d.cum_before = live ? tail.cum_before + tail.size : 0u;
        d.min_ts = live ? tail.min_ts : d.start_ts;  // start_ts of the vertex's oldest live block travels with the newest
        if (!live) ent.cum_first = 0;
        dir[ent.end] = d;
        ent.end++;
        tail = d;
        agg[0] += 1;
        if (!replace) agg[1] += r.newcap;
      } else if (r.flags & kPlanRealloc) {
        free_push(ar, a.log, tail.payload, class_of_units(payload_units(tail.capacity)));
        tail.payload = np;
        tail.capacity = r.newcap;
        tail.size += r.cnt;
        tail.start_ts = fminf(tail.start_ts, first_ts);
        tail.end_ts = last_ts;
        dir[ent.end - 1] = tail;
      }
      ent.tail = tail;
      ent.num_edges += r.cnt;
      ent.num_insertions += 1;
      a.table[r.v] = ent;
      a.is_src[r.v] = 1;
      a.is_node[r.v] = 1;
    }
    // ---- vertex flags / edge-id reference counts (nodes_ / edges_ upkeep, dynamic_graph.cu:89-97), for the i-th edge
    //      of the batch AS GIVEN
    if (!node_seen) a.is_node[d_orig] = 1;  // hot vertices: test first, thousands of identical byte stores serialise in L2
    agg[2] += ref_new ? 1ull : 0ull;
  }
}
// Large batches: the upkeep of nodes_ / edges_ (dynamic_graph.cu:89-97) for the batch AS GIVEN in a pass of its own -- a
// pure stream over dst / eid at full occupancy -- so that the apply pass is left with the append.  Runs between the
// plan (which accepts or rejects the batch) and the apply pass (whose last CTA reports the counters to the host).
//
// Vertex flags.  Every thread that reads a flag as 0 stores a 1, and byte stores into one 32-byte sector queue up in its
// L2 slice: the first batch of a stream on a graph of a few thousand hot vertices put 1.2 M stores into 520 sectors and
// spent 0.45 ms on them.  BITMAP: a CTA marks its destinations in a bitmap in shared memory and writes the bitmap to a
// slab of its own; ingest_flags_merge_kernel ORs the slabs and sets each new flag (nearly) once.  (Vertex tables beyond
// kBookMaxWords * 32 entries leave the upkeep to the apply pass.)
//
// Reference counts.  Edge ids that increase strictly through the batch and span exactly n values (default ids, row
// numbers of a dataset: the prep pass has looked) are one contiguous run of the table, each word touched by one
// thread: a plain read-modify-write stream (14 us for 9.56 M ids), no atomics and no second read of the ids.  Otherwise
// an atomic per edge, whose old value says whether the id is new (205 us).
constexpr uint32_t kBookMaxWords = 8192;  // 32 KB of shared memory: vertex tables of up to 262 144 entries
constexpr uint32_t kBookMergeSlices = 32;  // 74 -> 19 dependent-free loads per thread at 592 slabs (the kernel is a chain of round trips)
__global__ void __launch_bounds__(kThreads) ingest_bookkeep_kernel(ApplyArgs a) {
  extern __shared__ uint32_t s_bm[];
  pdl_wait();
  pdl_trigger();
  unsigned long long fresh[1] = {0};
  if (a.cur->accepted != 0) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int U = 4;
    const long long emax = a.cur->max_eid, emin = 0x7fffffffffffffffll - a.cur->max_neg_eid;
    const bool contiguous = a.cur->eids_not_increasing == 0 && (unsigned long long)(emax - emin) + 1ull == a.n;
    {
      const uint32_t W = a.book_words;
      for (uint32_t w = threadIdx.x; w < W; w += kThreads) s_bm[w] = 0;
      __syncthreads();
      for (uint64_t i0 = t0; i0 < a.n; i0 += stride * U) {
        int64_t d[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const uint64_t i = i0 + (uint64_t)u * stride;
          d[u] = i < a.n ? __ldg(a.dst_orig + i) : -1;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          if (d[u] < 0) continue;
          const uint32_t w = (uint32_t)(d[u] >> 5), bit = 1u << (d[u] & 31);
          if (!(*reinterpret_cast<volatile uint32_t *>(s_bm + w) & bit)) atomicOr(s_bm + w, bit);
        }
      }
      __syncthreads();
      uint32_t *slab = a.book_bitmaps + (size_t)blockIdx.x * W;
      for (uint32_t w = threadIdx.x; w < W; w += kThreads) slab[w] = s_bm[w];
    }
    if (contiguous) {
      uint32_t *ref = a.eid_ref + (emin - a.eid_base);
      for (uint64_t i0 = t0; i0 < a.n; i0 += stride * U) {
        uint32_t v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const uint64_t i = i0 + (uint64_t)u * stride;
          v[u] = i < a.n ? __ldcg(ref + i) : 1u;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const uint64_t i = i0 + (uint64_t)u * stride;
          if (i < a.n) ref[i] = v[u] + 1u;
          fresh[0] += v[u] == 0 ? 1ull : 0ull;
        }
      }
    } else {
      for (uint64_t i = t0; i < a.n; i += stride) {
        const int64_t e = __ldg(a.eid_orig + i);
        fresh[0] += atomicAdd(&a.eid_ref[e - a.eid_base], 1u) == 0 ? 1ull : 0ull;
      }
    }
  }
  block_sum_u64(fresh);
  if (threadIdx.x == 0 && fresh[0]) atomicAdd(&a.stats->num_edges, fresh[0]);
}
// grid (words / kThreads, slices): a thread ORs one word of its slice of the slabs and sets the flags that are still 0
// (a flag may be stored once per slice)
__global__ void __launch_bounds__(kThreads) ingest_flags_merge_kernel(ApplyArgs a, uint32_t slabs, uint64_t table_cap) {
  pdl_wait();
  pdl_trigger();
  if (a.cur->accepted == 0) return;
  const uint32_t W = a.book_words, w = blockIdx.x * kThreads + threadIdx.x;
  if (w >= W) return;
  const uint32_t per = (slabs + gridDim.y - 1) / gridDim.y, g0 = blockIdx.y * per, g1 = min(slabs, g0 + per);
  uint32_t m = 0;
#pragma unroll 8
  for (uint32_t g = g0; g < g1; g++) m |= __ldcg(a.book_bitmaps + (size_t)g * W + w);
  if (!m) return;
  const uint64_t v0 = (uint64_t)w * 32;
  if (v0 + 32 <= table_cap) {  // the 32 flags of the word in two 16-byte loads
    const uint4 lo = __ldcg(reinterpret_cast<const uint4 *>(a.is_node + v0)), hi = __ldcg(reinterpret_cast<const uint4 *>(a.is_node + v0) + 1);
    const uint32_t f[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t set = 0;
#pragma unroll
    for (int k = 0; k < 8; k++)
#pragma unroll
      for (int q = 0; q < 4; q++)
        if ((f[k] >> (8 * q)) & 0xffu) set |= 1u << (4 * k + q);
    m &= ~set;
  }
  while (m) {
    const uint32_t bit = __ffs(m) - 1;
    m &= m - 1;
    if (v0 + bit < table_cap) a.is_node[v0 + bit] = 1;
  }
}
// after a CTA's edges: counters (ONE atomic per CTA and counter), and the last CTA reports to the host
__device__ __forceinline__ void ingest_apply_finish(const ApplyArgs &a, unsigned long long (&agg)[3]) {
  CallScratch *cur = a.cur;
  block_sum_u64(agg);
  if (threadIdx.x == 0) {
    if (agg[0]) atomicAdd(&a.stats->num_blocks, agg[0]);
    if (agg[1]) atomicAdd(&a.stats->allocated_elems, agg[1]);
    if (agg[2]) atomicAdd(&a.stats->num_edges, agg[2]);
    // the last CTA to get here reports to the host (mapped pinned memory: no copy to enqueue, no second sync)
    __threadfence();
    if (atomicAdd(&cur->done_ctas, 1u) == gridDim.x - 1) {
      __threadfence();
      volatile GraphStats *st = a.stats;
      HostResult h;
      h.call.max_id = *(volatile long long *)&cur->max_id;
      h.call.max_eid = *(volatile long long *)&cur->max_eid;
      h.call.max_neg_eid = *(volatile long long *)&cur->max_neg_eid;
      h.call.error_flags = *(volatile unsigned int *)&cur->error_flags;
      h.call.num_segments = *(volatile unsigned int *)&cur->num_segments;
      h.call.total_units = *(volatile unsigned int *)&cur->total_units;
      h.call.accepted = *(volatile unsigned int *)&cur->accepted;
      h.call.unsorted = *(volatile unsigned int *)&cur->unsorted;
      h.call.done_ctas = 0;
      h.num_edges = st->num_edges;
      h.num_blocks = st->num_blocks;
      h.allocated_elems = st->allocated_elems;
      h.free_units = st->arena.free_units;
      h.log_cnt = st->arena.log_cnt;
      h.sorted_cnt = st->arena.sorted_cnt;
      *a.hres = h;
      // the word the host polls (wait_flag_or_sync) goes last: every other CTA has finished its edges (it got here through
      // a fence), so when the host sees it the batch is applied and the report above is complete
      __threadfence_system();
      *(volatile unsigned int *)&a.hres->call.done_ctas = gridDim.x;
    }
  }
}
#ifndef GF_APPLY_OCC
#define GF_APPLY_OCC 1  // minimum resident CTAs per SM the apply kernel's register budget is sized for (build-time knob)
#endif
// A CTA applies `ept` x 256 consecutive edges (thread t: edges t, t + 256, ...).  ept = 1 for the reference's 100 000-edge
// batches (every SM gets work); large batches use several edges per thread: every CTA ends with a barrier, three atomics
// on the graph's counters, a fence and the done-counter round trip, and 37 000 one-edge CTAs queued on those few
// addresses.
__global__ void __launch_bounds__(kThreads, GF_APPLY_OCC) ingest_apply_kernel(ApplyArgs a, uint32_t ept) {
  const uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x * ept + threadIdx.x;
  pdl_wait();
  // the control words (histograms, tickets, look-back status) have been consumed by the kernels before this one
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < a.ctl_words; w += (uint64_t)gridDim.x * blockDim.x)
    a.ctl[w] = 0u;
  unsigned long long agg[3] = {0, 0, 0};
  if (a.cur->accepted != 0)
    for (uint32_t j = 0; j < ept; j++) {
      const uint64_t i = i0 + (uint64_t)j * blockDim.x;
      if (i < a.n) ingest_apply_edge(a, i, agg);
    }
  ingest_apply_finish(a, agg);
}

// ------------------------------------------------------------------------------------------------ free-list merge
// Folds the free log into the class-sorted array (counting sort by class into the other buffer).
struct MergeArgs {
  ArenaState *ar;
  const FreeRec *log;
  const unsigned long long *sorted_in;
  unsigned long long *sorted_out;
  unsigned int *work;  // [3 * kNumClasses] zeroed: new counts | new bases | cursors
};
__global__ void __launch_bounds__(kThreads) merge_count_kernel(MergeArgs m) {
  __shared__ unsigned int h[kNumClasses];
  for (int c = threadIdx.x; c < (int)kNumClasses; c += kThreads) h[c] = 0;
  __syncthreads();
  const unsigned int nlog = m.ar->log_cnt;
  for (unsigned int i = blockIdx.x * kThreads + threadIdx.x; i < nlog; i += gridDim.x * kThreads)
    atomicAdd(&h[m.log[i].cls], 1u);
  __syncthreads();
  for (int c = threadIdx.x; c < (int)kNumClasses; c += kThreads) {
    unsigned int v = h[c] + (blockIdx.x == 0 ? m.ar->free_cnt[c] : 0u);
    if (v) atomicAdd(&m.work[c], v);
  }
}
// one CTA per class: new bases (every CTA scans the 224 counts itself), then the class's surviving entries move over
__global__ void __launch_bounds__(kThreads) merge_move_kernel(MergeArgs m) {
  __shared__ unsigned int s_total;
  const unsigned int c = blockIdx.x;
  const unsigned int mine = threadIdx.x < kNumClasses ? m.work[threadIdx.x] : 0u;
  const unsigned int base = block_excl_scan(mine, &s_total);
  __shared__ unsigned int s_base[kNumClasses];
  if (threadIdx.x < kNumClasses) s_base[threadIdx.x] = base;
  __syncthreads();
  const unsigned int nb = s_base[c], old_base = m.ar->free_base[c], old_cnt = m.ar->free_cnt[c];
  for (unsigned int j = threadIdx.x; j < old_cnt; j += kThreads) m.sorted_out[nb + j] = m.sorted_in[old_base + j];
  if (threadIdx.x == 0) {
    m.work[kNumClasses + c] = nb;
    m.work[2 * kNumClasses + c] = nb + old_cnt;  // the log's entries of this class follow the survivors
  }
}
__global__ void __launch_bounds__(kThreads) merge_scatter_kernel(MergeArgs m) {
  const unsigned int nlog = m.ar->log_cnt;
  for (unsigned int i = blockIdx.x * kThreads + threadIdx.x; i < nlog; i += gridDim.x * kThreads) {
    const FreeRec r = m.log[i];
    m.sorted_out[atomicAdd(&m.work[2 * kNumClasses + r.cls], 1u)] = r.addr;
  }
}
__global__ void __launch_bounds__(kThreads) merge_finish_kernel(MergeArgs m) {
  __shared__ unsigned int s_total;
  const unsigned int c = threadIdx.x;
  const unsigned int cnt = c < kNumClasses ? m.work[c] : 0u;
  block_excl_scan(cnt, &s_total);
  if (c < kNumClasses) {
    m.ar->free_cnt[c] = cnt;
    m.ar->free_base[c] = m.work[kNumClasses + c];
  }
  if (c == 0) {
    m.ar->log_cnt = 0;
    m.ar->sorted_cnt = s_total;
  }
}

// ------------------------------------------------------------------------------------------------ offload
// DynamicGraph::OffloadOldBlocks, dynamic_graph.cu:382-411: one warp per vertex, oldest block first; the payloads go to
// the free log (TemporalBlockAllocator::Deallocate, temporal_block_allocator.cu:110-113,159-180).
// `drops` (optional) records (vertex, dir index) of every dropped block for the to_file path; the dropped
// descriptors stay readable in the directory and their payloads are not reused before the next merge.
__global__ void __launch_bounds__(kThreads) offload_kernel(NodeEntry *table, const uint8_t *__restrict__ is_node,
                                                           uint64_t table_len, float timestamp, uint32_t *eid_ref,
                                                           long long eid_base, GraphStats *stats, FreeRec *log,
                                                           uint2 *drops, uint32_t drops_cap, StoreParams sp) {
  uint64_t v = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (v >= table_len || !is_node[v]) return;
  NodeEntry ent = table[v];
  BlockDesc *dir = reinterpret_cast<BlockDesc *>(ent.dir());
  uint32_t first = ent.first;
  unsigned long long dropped = 0, gone_edges = 0, cap_sum = 0;
  while (first < ent.end) {
    BlockDesc d = dir[first];
    if (!(d.end_ts < timestamp)) break;
    const EdgeRec *de = blk_rec(d.payload, d.capacity);
    for (uint32_t i = lane; i < d.size; i += 32)
      if (atomicSub(&eid_ref[de[i].eid - eid_base], 1u) == 1u) gone_edges++;
    if (lane == 0) {
      if (drops) {
        unsigned long long k = atomicAdd(&stats->call_count, 1ull);
        if (k < drops_cap) drops[k] = make_uint2((uint32_t)v, first);
      }
      free_push(&stats->arena, log, d.payload, class_of_units(payload_units(d.capacity)));
    }
    dropped++;
    cap_sum += sp.policy == GF_INSERTION_REPLACE ? replace_logical_cap(d.size, sp.min_block) : d.capacity;
    first++;
  }
  if (!dropped) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gone_edges += __shfl_xor_sync(0xffffffffu, gone_edges, o);
  if (lane == 0) {
    ent.first = first;
    if (first < ent.end) {  // the newest descriptor carries the oldest live timestamp (sampler's window-start shortcut)
      ent.cum_first = dir[first].cum_before;
      ent.tail.min_ts = dir[first].start_ts;
      dir[ent.end - 1].min_ts = ent.tail.min_ts;
    } else {
      ent.cum_first = 0;
    }
    table[v] = ent;
    if (!drops) atomicAdd(&stats->call_count, dropped);
    atomicAdd(&stats->num_blocks, 0ull - dropped);
    atomicAdd(&stats->allocated_elems, 0ull - cap_sum);
    atomicAdd(&stats->num_edges, 0ull - gone_edges);
  }
}

// ------------------------------------------------------------------------------------------------ checkpoint
// After a checkpoint was loaded into freshly allocated chunks: every device address stored in the graph (directory and
// payload addresses in the vertex entries and descriptors, the free lists, the free log) moves from the chunk it was
// saved in to the chunk that replaced it.
struct RelocMap {
  unsigned int n;
  unsigned long long old_base[kMaxRegions], size[kMaxRegions], new_base[kMaxRegions];
};
__device__ __forceinline__ unsigned long long reloc(const RelocMap &m, unsigned long long a) {
  for (unsigned int k = 0; k < m.n; k++)
    if (a >= m.old_base[k] && a - m.old_base[k] < m.size[k]) return a - m.old_base[k] + m.new_base[k];
  return a;  // 0, or not an arena address
}
// one warp per vertex: the entry, its newest-descriptor copy, and every descriptor of the directory ([0, end): the
// offloaded ones are only read again by a to_file sweep, but stay consistent)
__global__ void __launch_bounds__(kThreads) reloc_table_kernel(NodeEntry *table, uint64_t table_len, const RelocMap *mp) {
  const uint64_t v = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= table_len) return;
  const RelocMap &m = *mp;
  NodeEntry ent = table[v];
  if (!ent.dir_tagged) return;
  const uint64_t ndir = reloc(m, ent.dir());
  BlockDesc *dir = reinterpret_cast<BlockDesc *>(ndir);
  for (uint32_t j = lane; j < ent.end; j += 32) dir[j].payload = reloc(m, dir[j].payload);
  if (lane == 0) {
    ent.dir_tagged = ndir | (ent.dir_tagged & 127ull);
    ent.tail.payload = reloc(m, ent.tail.payload);
    table[v] = ent;
  }
}
__global__ void __launch_bounds__(kThreads) reloc_lists_kernel(unsigned long long *sorted, uint64_t nsorted, FreeRec *log,
                                                               uint64_t nlog, ArenaState *ar, const RelocMap *mp) {
  const RelocMap &m = *mp;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t k = i; k < nsorted; k += stride) sorted[k] = reloc(m, sorted[k]);
  for (uint64_t k = i; k < nlog; k += stride) log[k].addr = reloc(m, log[k].addr);
  for (uint64_t k = i; k < ar->num_regions; k += stride) {
    const unsigned long long len = ar->regions[k].end - ar->regions[k].cur;
    // an exhausted region's `cur` equals the end of its chunk, which no chunk contains: move the pair by its start
    const unsigned long long nc = len ? reloc(m, ar->regions[k].cur) : reloc(m, ar->regions[k].cur - kUnit) + kUnit;
    ar->regions[k].cur = nc;
    ar->regions[k].end = nc + len;
  }
}

// index of the first non-zero reference count (n if none): where the live edge ids start
__global__ void first_nonzero_kernel(const uint32_t *__restrict__ ref, uint64_t n, unsigned long long *out) {
  unsigned long long m = n;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n && i < m; i += (uint64_t)gridDim.x * blockDim.x)
    if (ref[i]) m = i;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m < n) atomicMin(out, m);
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

}  // namespace gf
