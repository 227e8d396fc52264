// Hand-written device-wide primitives used by the ingest and cache paths: exclusive scan (u32) and a
// stable LSD radix sort of (u32 key, u32 value) pairs.  No Thrust / CUB.
#pragma once
#include "gf_common.cuh"

namespace gf {

// =====================================================================================================
// exclusive scan, u32.  Tiles of 1024 elements (256 threads x uint4), three phases, recursive on tile sums.
// =====================================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanTile = kScanThreads * 4;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread (256 threads); returns exclusive prefix, *total = sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total) {
  __shared__ uint32_t warp_sums[kScanThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t incl = warp_incl_scan(v, lane);
  if (lane == 31) warp_sums[w] = incl;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
    uint32_t si = warp_incl_scan(s, lane);
    if (lane < kScanThreads / 32) warp_sums[lane] = si - s;
    if (lane == kScanThreads / 32 - 1) *total = si;
  }
  __syncthreads();
  uint32_t r = incl - v + warp_sums[w];
  __syncthreads();
  return r;
}

__device__ __forceinline__ uint4 load4_guard(const uint32_t *in, uint64_t base, uint64_t n) {
  uint4 x = make_uint4(0, 0, 0, 0);
  if (base + 3 < n) {
    x = *reinterpret_cast<const uint4 *>(in + base);
  } else {
    if (base < n) x.x = in[base];
    if (base + 1 < n) x.y = in[base + 1];
    if (base + 2 < n) x.z = in[base + 2];
  }
  return x;
}

static __global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                                   uint32_t *__restrict__ tile_sums) {
  __shared__ uint32_t total;
  uint64_t base = (uint64_t)blockIdx.x * kScanTile + threadIdx.x * 4;
  uint4 x = load4_guard(in, base, n);
  block_excl_scan(x.x + x.y + x.z + x.w, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// tile_offsets == nullptr: single tile.  total_out (optional) receives the grand total (written by the last tile).
static __global__ void __launch_bounds__(kScanThreads) scan_downsweep_kernel(const uint32_t *in, uint32_t *out, uint64_t n,
                                                                      const uint32_t *__restrict__ tile_offsets,
                                                                      uint32_t *total_out) {
  __shared__ uint32_t total;
  uint64_t base = (uint64_t)blockIdx.x * kScanTile + threadIdx.x * 4;
  uint4 x = load4_guard(in, base, n);
  uint32_t mine = x.x + x.y + x.z + x.w;
  uint32_t off = block_excl_scan(mine, &total) + (tile_offsets ? tile_offsets[blockIdx.x] : 0u);
  uint4 y;
  y.x = off;
  y.y = y.x + x.x;
  y.z = y.y + x.y;
  y.w = y.z + x.z;
  if (base + 3 < n) {
    *reinterpret_cast<uint4 *>(out + base) = y;
  } else {
    if (base < n) out[base] = y.x;
    if (base + 1 < n) out[base + 1] = y.y;
    if (base + 2 < n) out[base + 2] = y.z;
  }
  if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0)
    *total_out = (tile_offsets ? tile_offsets[blockIdx.x] : 0u) + total;
}

// small arrays: one CTA of 1024 threads walks the array in chunks of 4096 (one launch instead of three)
constexpr int kScanSoloThreads = 1024;
constexpr uint64_t kScanSoloMax = 1u << 16;
static __global__ void __launch_bounds__(kScanSoloThreads) scan_solo_kernel(const uint32_t *in, uint32_t *out, uint64_t n,
                                                                      uint32_t *total_out) {
  __shared__ uint32_t warp_sums[kScanSoloThreads / 32];
  __shared__ uint32_t carry;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint64_t base0 = 0; base0 < n; base0 += (uint64_t)kScanSoloThreads * 4) {
    const uint64_t base = base0 + threadIdx.x * 4;
    const uint4 x = load4_guard(in, base, n);
    const uint32_t mine = x.x + x.y + x.z + x.w;
    const uint32_t incl = warp_incl_scan(mine, lane);
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
      const uint32_t sv = warp_sums[lane];
      const uint32_t si = warp_incl_scan(sv, lane);
      warp_sums[lane] = si - sv;
    }
    __syncthreads();
    const uint32_t c = carry;
    uint4 y;
    y.x = c + warp_sums[w] + incl - mine;
    y.y = y.x + x.x;
    y.z = y.y + x.y;
    y.w = y.z + x.z;
    if (base + 3 < n) {
      *reinterpret_cast<uint4 *>(out + base) = y;
    } else {
      if (base < n) out[base] = y.x;
      if (base + 1 < n) out[base + 1] = y.y;
      if (base + 2 < n) out[base + 2] = y.z;
    }
    __syncthreads();
    if (threadIdx.x == kScanSoloThreads - 1) carry = y.w + x.w;
    __syncthreads();
  }
  if (total_out && threadIdx.x == 0) *total_out = carry;
}

inline size_t scan_tmp_elems(uint64_t n) {
  size_t t = 0;
  while (n > (uint64_t)kScanTile) {
    n = (n + kScanTile - 1) / kScanTile;
    t += align_up(n, 64);
  }
  return t + 64;
}

// in/out may alias; both 16-byte aligned.  tmp: scan_tmp_elems(n) u32.  total_out: device pointer or nullptr.
inline int exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *total_out, uint32_t *tmp,
                              cudaStream_t st) {
  if (n == 0) {
    if (total_out) GF_CUDA(cudaMemsetAsync(total_out, 0, sizeof(uint32_t), st));
    return GF_OK;
  }
  uint64_t tiles = (n + kScanTile - 1) / kScanTile;
  if (tiles == 1) {
    gf::launch(scan_downsweep_kernel, 1, kScanThreads, 0, st, in, out, n, nullptr, total_out);
  } else if (n <= kScanSoloMax) {
    gf::launch(scan_solo_kernel, 1, kScanSoloThreads, 0, st, in, out, n, total_out);
  } else {
    gf::launch(scan_reduce_kernel, (unsigned)tiles, kScanThreads, 0, st, in, n, tmp);
    GF_TRY(exclusive_scan_u32(tmp, tmp, tiles, nullptr, tmp + align_up(tiles, 64), st));
    gf::launch(scan_downsweep_kernel, (unsigned)tiles, kScanThreads, 0, st, in, out, n, tmp, total_out);
  }
  GF_CUDA(cudaGetLastError());
  return GF_OK;
}

// =====================================================================================================
// single-pass exclusive scan with decoupled look-back (one launch for any n).  The input of element i is produced
// by `in(i)` and the result consumed by `out(i, exclusive_prefix, value)`, so that flagging / compaction steps fuse
// into the scan.  Tiles of 1024 elements (256 threads x 4); tile ids come from an atomic ticket (a tile's
// predecessors have always started); status words are {generation, flag, value} so nothing is reset between uses.
// =====================================================================================================
struct LookbackCtl {
  unsigned int *ticket;
  unsigned long long *status;  // [tiles]  (gen << 34) | (flag << 32) | value ; flag 1 = aggregate, 2 = inclusive prefix
  unsigned long long gen;
};
__device__ __forceinline__ unsigned long long lb_load(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void lb_store(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// exclusive prefix of tile `tile`; called by all 32 lanes of one warp after the tile's aggregate was published
__device__ __forceinline__ uint32_t lb_lookback_warp(const LookbackCtl &ctl, uint32_t tile, int lane) {
  uint32_t excl = 0;
  int64_t q0 = (int64_t)tile - 1;
  while (true) {
    const int64_t q = q0 - lane;
    unsigned long long w = 2ull << 32;  // tiles before the first: inclusive prefix 0
    bool ready = true;
    if (q >= 0) {
      w = lb_load(ctl.status + q);
      ready = (w >> 34) == ctl.gen && ((w >> 32) & 3ull) != 0;
    }
    const unsigned incl = __ballot_sync(0xffffffffu, ready && ((w >> 32) & 3ull) == 2ull);
    const unsigned nready = __ballot_sync(0xffffffffu, !ready);
    const int stop = incl ? __ffs(incl) - 1 : 31;  // nearest inclusive predecessor, or the whole window
    const unsigned need = stop == 31 ? 0xffffffffu : ((2u << stop) - 1u);
    if (nready & need) continue;  // a needed predecessor has not published yet: poll again
    excl += __reduce_add_sync(0xffffffffu, lane <= stop ? (uint32_t)w : 0u);
    if (incl) return excl;
    q0 -= 32;
  }
}

template <class In, class Out>
static __global__ void __launch_bounds__(kScanThreads) scan_lookback_kernel(uint64_t n, In in, Out out, LookbackCtl ctl,
                                                                     uint32_t *total_out) {
  __shared__ uint32_t s_tile, s_base, s_total;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(ctl.ticket, 1u);
    if (t == gridDim.x - 1) *ctl.ticket = 0;  // every tile of this launch has its ticket: re-arm
    s_tile = t;
  }
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t base = (uint64_t)tile * kScanTile + threadIdx.x * 4;
  uint32_t v[4];
#pragma unroll
  for (int k = 0; k < 4; k++) v[k] = base + k < n ? in(base + k) : 0u;
  const uint32_t mine = v[0] + v[1] + v[2] + v[3];
  uint32_t off = block_excl_scan(mine, &s_total);
  if (threadIdx.x < 32) {
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
      if (total_out && tile == gridDim.x - 1) *total_out = excl + total;
    }
  }
  __syncthreads();
  off += s_base;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (base + k < n) out(base + k, off, v[k]);
    off += v[k];
  }
}

// =====================================================================================================
// stable LSD radix sort of (u32 key, u32 value), 8 bits per pass, ONE kernel per pass ("onesweep"):
//   * radix_hist_all_kernel reads the keys once and builds the global digit histogram of every pass;
//   * radix_onesweep_kernel (one per pass): a tile takes a ticket, ranks its pairs per digit (warp match +
//     per-warp counters, input order preserved: warp w owns the contiguous strip w of the tile and walks it in
//     rounds of 32), publishes its 256 digit counts, resolves the counts of all earlier tiles with a per-digit
//     decoupled look-back (thread d follows digit d) while the pairs are being reordered in shared memory, and
//     writes them out digit run by digit run (consecutive threads -> consecutive addresses).
// Per pass and pair: 8 B read + 8 B written; the 3-kernel version re-read the keys for the histogram, scanned a
// tiles x 256 table in up to 3 more launches and scattered straight from registers (one 4-byte store per sector).
// Control words live in `tmp` and are cleared by one memset per sort.
// =====================================================================================================
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRoundsBig = 16, kSortRoundsSmall = 4;  // tile = 4096 / 1024 pairs
constexpr uint64_t kSortSmallN = 1u << 20;                // below this, small tiles spread the work over more SMs
constexpr int kSortMaxPasses = 4;
constexpr uint32_t kOsAgg = 1u << 30, kOsIncl = 2u << 30, kOsValue = (1u << 30) - 1u;  // status word = flag | count
inline int sort_rounds(uint64_t n) {
  static const uint64_t small_n = getenv("GNNFLOW_B200_SORT_SMALL_N") ? strtoull(getenv("GNNFLOW_B200_SORT_SMALL_N"), nullptr, 10)
                                                                      : kSortSmallN;  // experiment knob
  return n < small_n ? kSortRoundsSmall : kSortRoundsBig;
}
inline uint32_t sort_tiles(uint64_t n) {
  uint64_t tile = (uint64_t)kSortThreads * sort_rounds(n);
  return (uint32_t)((n + tile - 1) / tile);
}
// tmp layout (u32): ghist[kSortMaxPasses][256] | ticket[64] | status[pass][tile][256]
constexpr size_t kOsCtlElems = kSortMaxPasses * 256 + 64;
inline size_t radix_tmp_elems(uint64_t n) { return kOsCtlElems + (size_t)kSortMaxPasses * 256 * sort_tiles(n); }

__device__ __forceinline__ uint32_t os_load(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void os_store(uint32_t *p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Per-digit decoupled look-back of the onesweep passes: exclusive prefix of this thread's digit over the tiles before
// `tile` (my_status = this tile's word of the digit; a tile's words are 256 apart); publishes the inclusive prefix.
// W predecessor words are requested at once: when all tiles of a pass are resident together (batches of up to a
// few hundred thousand edges) no predecessor has an inclusive prefix yet and the walk goes back over every aggregate --
// one dependent L2 round trip per tile before (11 us per pass at 98 tiles), 1 / W of that now.
template <int W = 8>
__device__ __forceinline__ uint32_t os_lookback(uint32_t *my_status, uint32_t tile, uint32_t mine) {
  uint32_t excl = 0;
  int64_t q = (int64_t)tile - 1;
  bool done = false;
  while (!done) {
    uint32_t w[W];
#pragma unroll
    for (int j = 0; j < W; j++) w[j] = q - j >= 0 ? os_load(my_status - (int64_t)(tile - (q - j)) * 256) : kOsIncl;
#pragma unroll
    for (int j = 0; j < W; j++) {
      if (done) break;
      uint32_t sw = w[j];
      while (!(sw & (kOsAgg | kOsIncl))) sw = os_load(my_status - (int64_t)(tile - (q - j)) * 256);  // not published yet
      excl += sw & kOsValue;
      done = (sw & kOsIncl) != 0;
    }
    q -= W;
  }
  os_store(my_status, kOsIncl | (excl + mine));
  return excl;
}
// look-back window: small tiles are all resident at once (the walk covers every tile before) -> 32 words per round trip
#ifndef GF_OS_WINDOW_BIG
#define GF_OS_WINDOW_BIG 8  // big tiles (build-time experiment knob)
#endif
template <int ROUNDS>
struct OsWindow { static constexpr int value = ROUNDS <= 4 ? 32 : GF_OS_WINDOW_BIG; };

static __global__ void __launch_bounds__(kSortThreads) radix_hist_all_kernel(const uint32_t *__restrict__ keys, uint64_t n,
                                                                      int begin_bit, int passes,
                                                                      uint32_t *__restrict__ ghist) {
  __shared__ uint32_t hist[kSortMaxPasses][256];
  for (int i = threadIdx.x; i < kSortMaxPasses * 256; i += kSortThreads) (&hist[0][0])[i] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * kSortThreads * 4;
  for (uint64_t base = ((uint64_t)blockIdx.x * kSortThreads + threadIdx.x) * 4; base < n; base += stride) {
    const uint4 x = load4_guard(keys, base, n);
    const uint32_t k[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (base + j < n) {
        for (int p = 0; p < passes; p++) atomicAdd(&hist[p][(k[j] >> (begin_bit + 8 * p)) & 255u], 1u);
      }
    }
  }
  __syncthreads();
  for (int p = 0; p < passes; p++) {
    const uint32_t c = hist[p][threadIdx.x];
    if (c) atomicAdd(&ghist[p * 256 + threadIdx.x], c);
  }
}

template <int kSortRounds>
static __global__ void __launch_bounds__(kSortThreads) radix_onesweep_kernel(const uint32_t *__restrict__ keys_in,
                                                                      const uint32_t *__restrict__ vals_in,
                                                                      uint32_t *__restrict__ keys_out,
                                                                      uint32_t *__restrict__ vals_out, uint64_t n,
                                                                      int shift, const uint32_t *__restrict__ ghist,
                                                                      uint32_t *ticket, uint32_t *status,
                                                                      const uint32_t *active_passes = nullptr, int pass = 0) {
  constexpr int kSortTile = kSortThreads * kSortRounds;
  // a caller whose key range is only known on the device launches the passes of the worst case; the ones beyond
  // *active_passes have nothing to order (their digits are all zero) and return at once, leaving the data where it is
  if (active_passes && (uint32_t)pass >= *active_passes) return;
  __shared__ uint32_t cnt[kSortWarps][256];  // per-warp digit counts -> exclusive prefix over warps
  __shared__ uint32_t dstart[256];           // first local slot of each digit in the reordered tile
  __shared__ uint32_t gbase[256];            // global position of local slot e with digit d = gbase[d] + e
  __shared__ uint32_t sk[kSortTile], sv[kSortTile];
  __shared__ uint32_t s_tile, s_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t tile_base = (uint64_t)tile * kSortTile;
  const uint32_t tile_n = (uint32_t)min((uint64_t)kSortTile, n - tile_base);
  const uint64_t base = tile_base + (uint64_t)w * (32 * kSortRounds) + lane;
  uint32_t k[kSortRounds], v[kSortRounds];
  uint16_t rk[kSortRounds];  // stable rank among the same digit inside this warp's strip
#pragma unroll
  for (int r = 0; r < kSortRounds; r++) {
    const uint64_t i = base + r * 32;
    const bool valid = i < n;
    k[r] = valid ? keys_in[i] : 0u;
    v[r] = valid ? vals_in[i] : 0u;
  }
#pragma unroll
  for (int r = 0; r < kSortRounds; r++) {
    const bool valid = base + r * 32 < n;
    const uint32_t d = (k[r] >> shift) & 255u;
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : (0x100u | lane));
    const uint32_t before = cnt[w][d];
    __syncwarp();
    if (valid && (peers & lt_mask) == 0) cnt[w][d] = before + __popc(peers);
    __syncwarp();
    rk[r] = (uint16_t)(before + __popc(peers & lt_mask));
  }
  __syncthreads();
  // thread d owns digit d from here on
  const uint32_t d = threadIdx.x;
  uint32_t mine = 0;
#pragma unroll
  for (int ww = 0; ww < kSortWarps; ww++) {
    const uint32_t t = cnt[ww][d];
    cnt[ww][d] = mine;
    mine += t;
  }
  uint32_t *my_status = status + (uint64_t)tile * 256 + d;
  os_store(my_status, (tile == 0 ? kOsIncl : kOsAgg) | mine);
  const uint32_t local_start = block_excl_scan(mine, &s_total);
  dstart[d] = local_start;
  const uint32_t digit_base = block_excl_scan(ghist[d], &s_total);  // global start of digit d
  __syncthreads();
  // reorder the tile in shared memory
#pragma unroll
  for (int r = 0; r < kSortRounds; r++) {
    if (base + r * 32 < n) {
      const uint32_t dd = (k[r] >> shift) & 255u;
      const uint32_t pos = dstart[dd] + cnt[w][dd] + rk[r];
      sk[pos] = k[r];
      sv[pos] = v[r];
    }
  }
  // per-digit look-back over the earlier tiles
  uint32_t excl = 0;
  if (tile > 0) excl = os_lookback<OsWindow<kSortRounds>::value>(my_status, tile, mine);
  gbase[d] = digit_base + excl - local_start;
  __syncthreads();
  for (uint32_t e = threadIdx.x; e < tile_n; e += kSortThreads) {
    const uint32_t key = sk[e];
    const uint32_t pos = gbase[(key >> shift) & 255u] + e;
    keys_out[pos] = key;
    vals_out[pos] = sv[e];
  }
}

// The passes of a sort whose control words (tmp: radix_tmp_elems(n) u32) are ALREADY cleared and whose digit histograms
// (tmp[p * 256 + d], pass p, digit d) are already built -- by radix_sort_pairs below, or by a caller that produces the
// keys and counts their digits in the same kernel (gf_cache.cu).
inline size_t radix_ctl_bytes(uint64_t n, int passes) { return (kOsCtlElems + (size_t)passes * 256 * sort_tiles(n)) * sizeof(uint32_t); }
inline int radix_sort_pairs_prepared(uint32_t *k0, uint32_t *v0, uint32_t *k1, uint32_t *v1, uint64_t n, int begin_bit,
                                     int passes, uint32_t *tmp, bool *result_in_0, cudaStream_t st,
                                     const uint32_t *active_passes = nullptr) {
  *result_in_0 = true;
  const uint32_t tiles = sort_tiles(n);
  const bool small = sort_rounds(n) == kSortRoundsSmall;
  uint32_t *ghist = tmp, *ticket = tmp + kSortMaxPasses * 256, *status = tmp + kOsCtlElems;
  uint32_t *ki = k0, *vi = v0, *ko = k1, *vo = v1;
  for (int p = 0; p < passes; p++) {
    const int shift = begin_bit + 8 * p;
    uint32_t *stat = status + (size_t)p * 256 * tiles;
    if (small)
      gf::launch(radix_onesweep_kernel<kSortRoundsSmall>, tiles, kSortThreads, 0, st, (const uint32_t *)ki, (const uint32_t *)vi, ko,
                 vo, n, shift, (const uint32_t *)(ghist + p * 256), ticket + p, stat, active_passes, p);
    else
      gf::launch(radix_onesweep_kernel<kSortRoundsBig>, tiles, kSortThreads, 0, st, (const uint32_t *)ki, (const uint32_t *)vi, ko,
                 vo, n, shift, (const uint32_t *)(ghist + p * 256), ticket + p, stat, active_passes, p);
    GF_CUDA(cudaGetLastError());
    uint32_t *t;
    t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
    *result_in_0 = !*result_in_0;
  }
  return GF_OK;
}

// Sorts bits [begin_bit, end_bit) of the keys.  Ping-pongs between (k0,v0) and (k1,v1); *result_in_0 tells
// which pair holds the sorted output.  tmp: radix_tmp_elems(n) u32.
inline int radix_sort_pairs(uint32_t *k0, uint32_t *v0, uint32_t *k1, uint32_t *v1, uint64_t n, int begin_bit,
                            int end_bit, uint32_t *tmp, bool *result_in_0, cudaStream_t st) {
  *result_in_0 = true;
  if (n == 0 || end_bit <= begin_bit) return GF_OK;
  if (n >= (1ull << 30)) GF_FAIL(GF_EINVAL, "radix_sort_pairs: %llu pairs exceed 2^30-1", (unsigned long long)n);
  const int passes = (end_bit - begin_bit + 7) / 8;
  if (passes > kSortMaxPasses) GF_FAIL(GF_EINVAL, "radix_sort_pairs: more than %d passes", kSortMaxPasses);
  GF_CUDA(cudaMemsetAsync(tmp, 0, radix_ctl_bytes(n, passes), st));
  gf::launch(radix_hist_all_kernel, std::min<unsigned>(cdiv(n, kSortThreads * 16), 148u * 8), kSortThreads, 0, st, k0, n,
             begin_bit, passes, tmp);
  return radix_sort_pairs_prepared(k0, v0, k1, v1, n, begin_bit, passes, tmp, result_in_0, st);
}

// float -> u32 whose unsigned order equals the float order (negatives flipped)
__host__ __device__ inline uint32_t orderable_f32(float f) {
  uint32_t b;
#ifdef __CUDA_ARCH__
  b = __float_as_uint(f);
#else
  memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

}  // namespace gf
