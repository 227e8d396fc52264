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
  } else {
    gf::launch(scan_reduce_kernel, (unsigned)tiles, kScanThreads, 0, st, in, n, tmp);
    GF_TRY(exclusive_scan_u32(tmp, tmp, tiles, nullptr, tmp + align_up(tiles, 64), st));
    gf::launch(scan_downsweep_kernel, (unsigned)tiles, kScanThreads, 0, st, in, out, n, tmp, total_out);
  }
  GF_CUDA(cudaGetLastError());
  return GF_OK;
}

// =====================================================================================================
// stable LSD radix sort of (u32 key, u32 value), 8 bits per pass.
// A tile is 4096 pairs; warp w owns the contiguous 512-pair strip w of the tile and walks it in 16 rounds of
// 32, so (warp, round, lane) order == input order and per-digit ranks are stable.
// =====================================================================================================
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRounds = 16;
constexpr int kSortTile = kSortThreads * kSortRounds;

static __global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint32_t *__restrict__ keys, uint64_t n,
                                                                  int shift, uint32_t *__restrict__ tile_hist,
                                                                  uint32_t num_tiles) {
  __shared__ uint32_t hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint64_t base = (uint64_t)blockIdx.x * kSortTile + (uint64_t)w * (32 * kSortRounds) + lane;
#pragma unroll 4
  for (int r = 0; r < kSortRounds; r++) {
    uint64_t i = base + r * 32;
    if (i < n) atomicAdd(&hist[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  tile_hist[(uint64_t)threadIdx.x * num_tiles + blockIdx.x] = hist[threadIdx.x];  // digit-major
}

static __global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint32_t *__restrict__ keys_in,
                                                                     const uint32_t *__restrict__ vals_in,
                                                                     uint32_t *__restrict__ keys_out,
                                                                     uint32_t *__restrict__ vals_out, uint64_t n,
                                                                     int shift,
                                                                     const uint32_t *__restrict__ tile_offsets,
                                                                     uint32_t num_tiles) {
  __shared__ uint32_t cnt[kSortWarps][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  uint64_t base = (uint64_t)blockIdx.x * kSortTile + (uint64_t)w * (32 * kSortRounds) + lane;
  uint32_t k[kSortRounds], v[kSortRounds];
  // pass A: per-warp digit counts
#pragma unroll
  for (int r = 0; r < kSortRounds; r++) {
    uint64_t i = base + r * 32;
    bool valid = i < n;
    k[r] = valid ? keys_in[i] : 0u;
    v[r] = valid ? vals_in[i] : 0u;
    uint32_t d = (k[r] >> shift) & 255u;
    unsigned peers = __match_any_sync(0xffffffffu, valid ? d : (0x100u | lane));
    if (valid && (peers & lt_mask) == 0) cnt[w][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {  // exclusive prefix over warps per digit, seeded with this tile's global offset for the digit
    uint32_t d = threadIdx.x;
    uint32_t run = tile_offsets[(uint64_t)d * num_tiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < kSortWarps; ww++) {
      uint32_t t = cnt[ww][d];
      cnt[ww][d] = run;
      run += t;
    }
  }
  __syncthreads();
  // pass B: stable ranks, scatter
#pragma unroll
  for (int r = 0; r < kSortRounds; r++) {
    uint64_t i = base + r * 32;
    bool valid = i < n;
    uint32_t d = (k[r] >> shift) & 255u;
    unsigned peers = __match_any_sync(0xffffffffu, valid ? d : (0x100u | lane));
    if (valid) {
      uint32_t pos = cnt[w][d] + __popc(peers & lt_mask);
      keys_out[pos] = k[r];
      vals_out[pos] = v[r];
    }
    __syncwarp();
    if (valid && (peers & lt_mask) == 0) cnt[w][d] += __popc(peers);
    __syncwarp();
  }
}

inline size_t radix_hist_elems(uint64_t n) {
  uint64_t tiles = (n + kSortTile - 1) / kSortTile;
  return align_up(256 * tiles, 64);
}
inline size_t radix_tmp_elems(uint64_t n) { return radix_hist_elems(n) + scan_tmp_elems(radix_hist_elems(n)); }

// Sorts bits [begin_bit, end_bit) of the keys.  Ping-pongs between (k0,v0) and (k1,v1); *result_in_0 tells
// which pair holds the sorted output.  tmp: radix_tmp_elems(n) u32.
inline int radix_sort_pairs(uint32_t *k0, uint32_t *v0, uint32_t *k1, uint32_t *v1, uint64_t n, int begin_bit,
                            int end_bit, uint32_t *tmp, bool *result_in_0, cudaStream_t st) {
  *result_in_0 = true;
  if (n == 0) return GF_OK;
  uint32_t tiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
  size_t hist_elems = radix_hist_elems(n);
  uint32_t *hist = tmp;
  uint32_t *scan_tmp = tmp + hist_elems;
  uint32_t *ki = k0, *vi = v0, *ko = k1, *vo = v1;
  for (int shift = begin_bit; shift < end_bit; shift += 8) {
    gf::launch(radix_hist_kernel, tiles, kSortThreads, 0, st, ki, n, shift, hist, tiles);
    GF_TRY(exclusive_scan_u32(hist, hist, 256ull * tiles, nullptr, scan_tmp, st));
    gf::launch(radix_scatter_kernel, tiles, kSortThreads, 0, st, ki, vi, ko, vo, n, shift, hist, tiles);
    GF_CUDA(cudaGetLastError());
    uint32_t *t;
    t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
    *result_in_0 = !*result_in_0;
  }
  return GF_OK;
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
