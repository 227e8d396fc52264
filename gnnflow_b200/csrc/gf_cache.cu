// Feature-cache gather and LRU / FIFO policy updates as CUDA kernels.  Replaces the ~12 torch index kernels +
// torch.unique + host index_select + H2D per MFG block of reference gnnflow/cache/cache.py:255-413 and the
// policy updates of lru_cache.py:121-201 / fifo_cache.py:77-161.
#include <algorithm>

#include "gf_primitives.cuh"

namespace gf {

constexpr int kCThreads = 256;

template <typename V>
__device__ __forceinline__ V ld_stream(const V *p) { return __ldg(p); }

// out[i,:] = flag[id] ? buffer[map[id],:] : features[id,:].  One warp per row, ROWS rows in flight per warp so
// that the dependent id -> flag -> map -> row chain of one row overlaps the row copy of another.
template <typename V, int ROWS>
__global__ void __launch_bounds__(kCThreads) cache_gather_kernel(const int64_t *__restrict__ ids, uint64_t n,
                                                                 const uint8_t *__restrict__ flag,
                                                                 const int64_t *__restrict__ map,
                                                                 const V *__restrict__ buffer,
                                                                 const V *__restrict__ features, uint32_t nvec,
                                                                 V *__restrict__ out, uint8_t *__restrict__ hit_mask,
                                                                 unsigned long long *num_hits) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  unsigned hits = 0;
  for (uint64_t r0 = warp * ROWS; r0 < n; r0 += nwarps * ROWS) {
    const V *srow[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      uint64_t i = r0 + r;
      srow[r] = nullptr;
      if (i < n) {
        int64_t id = __ldg(ids + i);
        bool hit = flag ? __ldg(flag + id) != 0 : false;
        srow[r] = hit ? buffer + (uint64_t)__ldg(map + id) * nvec : features + (uint64_t)id * nvec;
        if (lane == 0) {
          hits += hit;
          if (hit_mask) hit_mask[i] = hit;
        }
      }
    }
    for (uint32_t c = lane; c < nvec; c += 32) {
      V v[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; r++)
        if (srow[r]) v[r] = ld_stream(srow[r] + c);
#pragma unroll
      for (int r = 0; r < ROWS; r++)
        if (srow[r]) out[(r0 + r) * nvec + c] = v[r];
    }
  }
  if (num_hits && lane == 0 && hits) atomicAdd(num_hits, (unsigned long long)hits);
}

template <typename V>
static int launch_gather(const int64_t *ids, uint64_t n, const uint8_t *flag, const int64_t *map, const float *buffer,
                         const float *features, uint32_t dim, float *out, uint8_t *hit_mask, uint64_t *num_hits,
                         cudaStream_t st) {
  constexpr int ROWS = 4;
  uint32_t nvec = dim / (sizeof(V) / 4);
  uint64_t warps = (n + ROWS - 1) / ROWS;
  unsigned blocks = (unsigned)std::min<uint64_t>((warps + kCThreads / 32 - 1) / (kCThreads / 32), 148ull * 16);
  gf::launch(cache_gather_kernel<V, ROWS>, blocks, kCThreads, 0, st, ids, n, flag, map, (const V *)buffer, (const V *)features, nvec,
                                                             (V *)out, hit_mask, (unsigned long long *)num_hits);
  GF_CUDA(cudaGetLastError());
  return GF_OK;
}

static bool aligned(const void *p, size_t a) { return ((uintptr_t)p % a) == 0; }

static int gather_dispatch(const int64_t *ids, uint64_t n, const uint8_t *flag, const int64_t *map, const float *buffer,
                           const float *features, uint32_t dim, float *out, uint8_t *hit_mask, uint64_t *num_hits,
                           cudaStream_t st) {
  if (n == 0) return GF_OK;
  if (!ids || !features || !out || dim == 0) GF_FAIL(GF_EINVAL, "gather: null argument");
  if (flag && (!map || !buffer)) GF_FAIL(GF_EINVAL, "gather: cache_flag without cache_map / cache_buffer");
  bool a16 = aligned(features, 16) && aligned(out, 16) && (!flag || aligned(buffer, 16));
  bool a8 = aligned(features, 8) && aligned(out, 8) && (!flag || aligned(buffer, 8));
  if (dim % 4 == 0 && a16) return launch_gather<float4>(ids, n, flag, map, buffer, features, dim, out, hit_mask, num_hits, st);
  if (dim % 2 == 0 && a8) return launch_gather<float2>(ids, n, flag, map, buffer, features, dim, out, hit_mask, num_hits, st);
  return launch_gather<float>(ids, n, flag, map, buffer, features, dim, out, hit_mask, num_hits, st);
}

// out[i,:] = shards[owner[id]][local_index[id],:]: rows of other ranks are read over NVLink through IPC-mapped
// pointers.  Same shape as the cache gather: one warp per row, ROWS rows in flight per warp.
template <typename V, int ROWS>
__global__ void __launch_bounds__(kCThreads) partitioned_gather_kernel(const int64_t *__restrict__ ids, uint64_t n,
                                                                       const int8_t *__restrict__ owner,
                                                                       const int32_t *__restrict__ local_index,
                                                                       uint64_t num_items,
                                                                       const float *const *__restrict__ shards,
                                                                       uint32_t world, uint32_t nvec, V *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r0 = warp * ROWS; r0 < n; r0 += nwarps * ROWS) {
    const V *srow[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      const uint64_t i = r0 + r;
      srow[r] = nullptr;
      if (i < n) {
        const int64_t id = __ldg(ids + i);
        if (id >= 0 && (uint64_t)id < num_items) {
          const int o = __ldg(owner + id);
          if (o >= 0 && o < (int)world)
            srow[r] = reinterpret_cast<const V *>(shards[o]) + (uint64_t)__ldg(local_index + id) * nvec;
        }
      }
    }
    for (uint32_t c = lane; c < nvec; c += 32) {
      V v[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; r++) {
        memset(&v[r], 0, sizeof(V));
        if (srow[r]) v[r] = srow[r][c];  // plain load: peer memory is not cached in the local L2 anyway
      }
#pragma unroll
      for (int r = 0; r < ROWS; r++)
        if (r0 + r < n) out[(r0 + r) * nvec + c] = v[r];
    }
  }
}

// ---------------------------------------------------------------------------------------- policy updates
struct UpdCtl {       // device control block inside the scratch area
  uint32_t num_miss;  // misses (with duplicates)
  uint32_t num_uniq;  // unique misses
  uint32_t k;         // admitted = min(num_uniq, capacity)
  uint32_t pad;
};

// compact the missed ids (as u32 keys); also LRU bookkeeping for the hits
__global__ void upd_collect_kernel(const int64_t *__restrict__ ids, const uint8_t *__restrict__ hit_mask, uint64_t n,
                                   uint32_t *miss_keys, uint32_t *miss_vals, UpdCtl *ctl) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool miss = i < n && !hit_mask[i];
  unsigned m = __ballot_sync(0xffffffffu, miss);
  int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == 0 && m) base = atomicAdd(&ctl->num_miss, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (miss) {
    uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
    miss_keys[pos] = (uint32_t)ids[i];
    miss_vals[pos] = pos;
  }
}
__global__ void upd_pad_kernel(uint32_t *keys, uint64_t n, const UpdCtl *ctl) {  // unused tail sorts last
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && i >= ctl->num_miss) keys[i] = 0xffffffffu;
}
__global__ void upd_unique_flags_kernel(const uint32_t *__restrict__ keys, uint64_t n, const UpdCtl *ctl, uint32_t *flags) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[i] = (i < ctl->num_miss && (i == 0 || keys[i] != keys[i - 1])) ? 1u : 0u;
}
__global__ void upd_unique_compact_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ flags_excl,
                                          uint64_t n, uint32_t *uniq, UpdCtl *ctl, uint32_t capacity) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || i >= ctl->num_miss) return;
  bool head = i == 0 || keys[i] != keys[i - 1];
  if (head) uniq[flags_excl[i]] = keys[i];
  if (i == ctl->num_miss - 1) {
    uint32_t u = flags_excl[i] + (head ? 1u : 0u);
    ctl->num_uniq = u;
    ctl->k = min(u, capacity);
  }
}
// LRU: count -= 1 everywhere, hits -> 0 (lru_cache.py:142-145); only when this fetch had a miss (cache.py:317)
__global__ void lru_age_kernel(int32_t *count, uint64_t capacity, const UpdCtl *ctl) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < capacity && ctl->num_miss) count[i] -= 1;
}
__global__ void lru_touch_kernel(const int64_t *__restrict__ ids, const uint8_t *__restrict__ hit_mask, uint64_t n,
                                 const int64_t *__restrict__ map, int32_t *count, const UpdCtl *ctl) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && ctl->num_miss && hit_mask[i]) count[map[ids[i]]] = 0;
}
__global__ void lru_keys_kernel(const int32_t *__restrict__ count, uint64_t capacity, uint32_t *keys, uint32_t *vals) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= capacity) return;
  keys[i] = (uint32_t)count[i] ^ 0x80000000u;  // signed order -> unsigned order
  vals[i] = (uint32_t)i;
}
// LFU: count[hit slots] += 1, ONCE per distinct slot however often the slot was hit in this fetch (the reference's
// `count[cached_index] += 1` is a non-accumulating index_put, lfu_cache.py:158): hits set a mark bit, the sweep over
// the slots that builds the sort keys folds the mark into the count.  Counts stay < 2^30.
constexpr int32_t kLfuMark = 1 << 30;
__global__ void lfu_mark_kernel(const int64_t *__restrict__ ids, const uint8_t *__restrict__ hit_mask, uint64_t n,
                                const int64_t *__restrict__ map, int32_t *count, const UpdCtl *ctl) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && ctl->num_miss && hit_mask[i]) atomicOr(count + map[ids[i]], kLfuMark);
}
__global__ void lfu_keys_kernel(int32_t *count, uint64_t capacity, uint32_t *keys, uint32_t *vals) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= capacity) return;
  int32_t c = count[i];
  if (c & kLfuMark) {
    c = (c & ~kLfuMark) + 1;
    count[i] = c;
  }
  keys[i] = (uint32_t)c ^ 0x80000000u;
  vals[i] = (uint32_t)i;
}
// admit uniq[j] into slot victim(j), j < k  (lru_cache.py:151-160 / fifo_cache.py:106-116 / lfu_cache.py:161-172).
// One warp per admission.
template <bool FIFO>
__global__ void __launch_bounds__(kCThreads) upd_apply_kernel(gf_cache_state c, const uint32_t *__restrict__ uniq,
                                                              const uint32_t *__restrict__ victims,
                                                              const float *__restrict__ features, const UpdCtl *ctl,
                                                              const int64_t *fifo_ptr, int32_t admit_count) {
  const int lane = threadIdx.x & 31;
  const uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t k = ctl->k;
  if (j >= k) return;
  uint64_t slot;
  if (FIFO) {
    const int64_t ptr = *fifo_ptr, cap = (int64_t)c.capacity;
    if (ptr + (int64_t)k < cap) {
      slot = (uint64_t)(ptr + 1 + (int64_t)j);
    } else {  // wrap: [0, r) ++ [ptr + 1, cap), fifo_cache.py:101-103
      const int64_t r = (int64_t)k - (cap - 1 - ptr);
      slot = (int64_t)j < r ? j : (uint64_t)(ptr + 1 + ((int64_t)j - r));
    }
  } else {
    slot = victims[j];
  }
  const int64_t new_id = uniq[j];
  if (lane == 0) {
    const int64_t old_id = c.index_to_id[slot];
    if (old_id >= 0) {
      c.flag[old_id] = 0;
      c.map[old_id] = -1;
    }
  }
  __syncwarp();
  const float *src = features + (uint64_t)new_id * c.dim;
  float *dst = c.buffer + slot * c.dim;
  for (uint32_t d = lane; d < c.dim; d += 32) dst[d] = __ldg(src + d);
  if (lane == 0) {
    if (!FIFO) c.count[slot] = admit_count;  // LRU: 0 (lru_cache.py:153), LFU: 1 (lfu_cache.py:166)
    c.index_to_id[slot] = new_id;
  }
}
// second phase so that an id evicted and an id admitted never race on flag/map (they are disjoint sets, but
// two admissions may evict/admit in any order)
__global__ void upd_publish_kernel(gf_cache_state c, const uint32_t *__restrict__ uniq, const UpdCtl *ctl,
                                   const uint32_t *__restrict__ victims, int fifo, int64_t *fifo_ptr) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t k = ctl->k;
  if (j < k) {
    uint64_t slot;
    if (fifo) {
      const int64_t ptr = *fifo_ptr, cap = (int64_t)c.capacity;
      if (ptr + (int64_t)k < cap) slot = (uint64_t)(ptr + 1 + (int64_t)j);
      else {
        const int64_t r = (int64_t)k - (cap - 1 - ptr);
        slot = (int64_t)j < r ? j : (uint64_t)(ptr + 1 + ((int64_t)j - r));
      }
    } else slot = victims[j];
    const int64_t id = uniq[j];
    c.flag[id] = 1;
    c.map[id] = (int64_t)slot;
  }
}
__global__ void fifo_advance_kernel(int64_t *fifo_ptr, const UpdCtl *ctl, uint64_t capacity) {
  const int64_t ptr = *fifo_ptr, cap = (int64_t)capacity, k = ctl->k;
  if (k == 0) return;
  if (ptr + k < cap) *fifo_ptr = ptr + k;
  else *fifo_ptr = k - (cap - 1 - ptr) - 1;  // fifo_cache.py:104-105
}

struct UpdScratch {
  UpdCtl *ctl;
  uint32_t *k0, *v0, *k1, *v1, *flags, *uniq, *tmp;
};
static uint64_t upd_elems(uint64_t n, uint64_t capacity) { return align_up(std::max(n, capacity) + 1, 64); }
static UpdScratch carve_upd(void *scratch, uint64_t n, uint64_t capacity) {
  uint64_t m = upd_elems(n, capacity);
  UpdScratch s;
  s.ctl = reinterpret_cast<UpdCtl *>(scratch);
  s.k0 = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(scratch) + 256);
  s.v0 = s.k0 + m;
  s.k1 = s.v0 + m;
  s.v1 = s.k1 + m;
  s.flags = s.v1 + m;
  s.uniq = s.flags + m;
  s.tmp = s.uniq + m;
  return s;
}

enum { kPolicyLru = 0, kPolicyFifo = 1, kPolicyLfu = 2 };
static int cache_update(gf_cache_state *c, const int64_t *ids, const uint8_t *hit_mask, uint64_t n, const float *features,
                        int policy, int64_t *fifo_ptr, void *scratch, uint64_t scratch_bytes, cudaStream_t st) {
  const bool fifo = policy == kPolicyFifo, lfu = policy == kPolicyLfu;
  if (!c || !ids || !hit_mask || !features || !scratch) GF_FAIL(GF_EINVAL, "cache update: null argument");
  if (!c->buffer || !c->flag || !c->map || !c->index_to_id || (!fifo && !c->count) || (fifo && !fifo_ptr))
    GF_FAIL(GF_EINVAL, "cache update: incomplete cache state");
  if (c->num_items >= (1ull << 32) || c->capacity >= (1ull << 31)) GF_FAIL(GF_EINVAL, "cache too large");
  if (scratch_bytes < gf_cache_update_scratch_bytes(n, c->capacity)) GF_FAIL(GF_ECAPACITY, "cache update: scratch too small");
  if (n == 0 || c->capacity == 0) return GF_OK;
  UpdScratch s = carve_upd(scratch, n, c->capacity);
  const unsigned nb = cdiv(n, kCThreads), cb = cdiv(c->capacity, kCThreads);
  GF_CUDA(cudaMemsetAsync(s.ctl, 0, sizeof(UpdCtl), st));
  gf::launch(upd_collect_kernel, nb, kCThreads, 0, st, ids, hit_mask, n, s.k0, s.v0, s.ctl);
  gf::launch(upd_pad_kernel, nb, kCThreads, 0, st, s.k0, n, s.ctl);
  // torch.unique(sorted=True) of the missed ids (cache.py:290,379)
  int bits = 1;
  while (bits < 32 && (1ull << bits) < c->num_items) bits++;
  bool in0;
  GF_TRY(radix_sort_pairs(s.k0, s.v0, s.k1, s.v1, n, 0, 32, s.tmp, &in0, st));  // full 32 bits: the 0xffffffff pad sorts last
  (void)bits;
  uint32_t *sk = in0 ? s.k0 : s.k1;
  uint32_t *free_k = in0 ? s.k1 : s.k0, *free_v = in0 ? s.v1 : s.v0, *free_v2 = in0 ? s.v0 : s.v1;
  gf::launch(upd_unique_flags_kernel, nb, kCThreads, 0, st, sk, n, s.ctl, s.flags);
  GF_TRY(exclusive_scan_u32(s.flags, s.flags, n, nullptr, s.tmp, st));
  gf::launch(upd_unique_compact_kernel, nb, kCThreads, 0, st, sk, s.flags, n, s.uniq, s.ctl, (uint32_t)c->capacity);
  const uint32_t *victims = nullptr;
  if (!fifo) {
    // k smallest water levels / use counts, ties -> lowest slot (stable sort of slots by count)
    // sk (sorted miss keys) is dead after the compaction; reuse the two free buffers + sk's partner
    uint32_t *ck0 = free_k, *cv0 = free_v, *ck1 = sk, *cv1 = free_v2;
    if (lfu) {
      gf::launch(lfu_mark_kernel, nb, kCThreads, 0, st, ids, hit_mask, n, c->map, c->count, s.ctl);
      gf::launch(lfu_keys_kernel, cb, kCThreads, 0, st, c->count, c->capacity, ck0, cv0);
    } else {
      gf::launch(lru_age_kernel, cb, kCThreads, 0, st, c->count, c->capacity, s.ctl);
      gf::launch(lru_touch_kernel, nb, kCThreads, 0, st, ids, hit_mask, n, c->map, c->count, s.ctl);
      gf::launch(lru_keys_kernel, cb, kCThreads, 0, st, c->count, c->capacity, ck0, cv0);
    }
    bool r0;
    GF_TRY(radix_sort_pairs(ck0, cv0, ck1, cv1, c->capacity, 0, 32, s.tmp, &r0, st));
    victims = r0 ? cv0 : cv1;
  }
  const uint64_t kmax = std::min<uint64_t>(n, c->capacity);
  if (fifo)
    gf::launch(upd_apply_kernel<true>, cdiv(kmax * 32, kCThreads), kCThreads, 0, st, *c, s.uniq, victims, features, s.ctl, fifo_ptr, 0);
  else
    gf::launch(upd_apply_kernel<false>, cdiv(kmax * 32, kCThreads), kCThreads, 0, st, *c, s.uniq, victims, features, s.ctl, fifo_ptr,
               lfu ? 1 : 0);
  gf::launch(upd_publish_kernel, cdiv(kmax, kCThreads), kCThreads, 0, st, *c, s.uniq, s.ctl, victims, fifo ? 1 : 0, fifo_ptr);
  if (fifo) gf::launch(fifo_advance_kernel, 1, 1, 0, st, fifo_ptr, s.ctl, c->capacity);
  GF_CUDA(cudaGetLastError());
  return GF_OK;
}

// ---------------------------------------------------------------------------------------- GNNLab static cache
// pre-sampling statistics: counts[id] += 1 once per distinct id of one block (gnnlab_static_cache.py:104-111, the
// non-accumulating `count[ids] += 1`): mark, then the first thread to clear an id's mark increments it.
__global__ void static_mark_kernel(const int64_t *__restrict__ ids, uint64_t n, int32_t *counts, uint64_t num_items) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (uint64_t)ids[i] < num_items) atomicOr(counts + ids[i], kLfuMark);
}
__global__ void static_fold_kernel(const int64_t *__restrict__ ids, uint64_t n, int32_t *counts, uint64_t num_items) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (uint64_t)ids[i] >= num_items) return;
  int32_t old = atomicAnd(counts + ids[i], ~kLfuMark);
  if (old & kLfuMark) atomicAdd(counts + ids[i], 1);
}
__global__ void static_keys_kernel(const int32_t *__restrict__ counts, uint64_t num_items, uint32_t *keys, uint32_t *vals) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_items) return;
  keys[i] = ~(uint32_t)counts[i];  // ascending sort of ~count == descending count; stable -> ties keep the lowest id
  vals[i] = (uint32_t)i;
}
__global__ void static_clear_kernel(gf_cache_state c) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c.num_items) {
    c.flag[i] = 0;
    c.map[i] = -1;
  }
}
// slot j <- the id with the j-th highest count (gnnlab_static_cache.py:130-141, 160-168).  One warp per slot.
__global__ void __launch_bounds__(kCThreads) static_fill_kernel(gf_cache_state c, const uint32_t *__restrict__ order,
                                                                const float *__restrict__ features) {
  const int lane = threadIdx.x & 31;
  const uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= c.capacity) return;
  const uint64_t id = order[j];
  const float *src = features + id * c.dim;
  float *dst = c.buffer + j * c.dim;
  for (uint32_t d = lane; d < c.dim; d += 32) dst[d] = __ldg(src + d);
  if (lane == 0) {
    c.flag[id] = 1;
    c.map[id] = (int64_t)j;
    if (c.index_to_id) c.index_to_id[j] = (int64_t)id;
  }
}

}  // namespace gf

using namespace gf;

GF_EXPORT int gf_cache_update_lfu(gf_cache_state *c, const int64_t *ids, const uint8_t *hit_mask, uint64_t n,
                                  const float *features, void *scratch, uint64_t scratch_bytes, void *stream) {
  return cache_update(c, ids, hit_mask, n, features, kPolicyLfu, nullptr, scratch, scratch_bytes, (cudaStream_t)stream);
}

GF_EXPORT int gf_cache_count_distinct(const int64_t *ids, uint64_t n, int32_t *counts, uint64_t num_items, void *stream) {
  if (n && (!ids || !counts)) GF_FAIL(GF_EINVAL, "count_distinct: null argument");
  if (n == 0) return GF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  gf::launch(static_mark_kernel, cdiv(n, kCThreads), kCThreads, 0, st, ids, n, counts, num_items);
  gf::launch(static_fold_kernel, cdiv(n, kCThreads), kCThreads, 0, st, ids, n, counts, num_items);
  GF_CUDA(cudaGetLastError());
  return GF_OK;
}

GF_EXPORT int gf_cache_fill_topk(gf_cache_state *c, const int32_t *counts, const float *features, void *scratch,
                                 uint64_t scratch_bytes, void *stream) {
  if (!c || !counts || !features || !scratch) GF_FAIL(GF_EINVAL, "fill_topk: null argument");
  if ((c->capacity && !c->buffer) || !c->flag || !c->map) GF_FAIL(GF_EINVAL, "fill_topk: incomplete cache state");
  if (c->num_items >= (1ull << 32) || c->capacity > c->num_items) GF_FAIL(GF_EINVAL, "fill_topk: bad capacity / num_items");
  if (scratch_bytes < gf_cache_update_scratch_bytes(c->num_items, c->capacity)) GF_FAIL(GF_ECAPACITY, "fill_topk: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->num_items == 0) return GF_OK;
  UpdScratch s = carve_upd(scratch, c->num_items, c->capacity);
  const unsigned ib = cdiv(c->num_items, kCThreads);
  gf::launch(static_clear_kernel, ib, kCThreads, 0, st, *c);
  if (c->capacity == 0) return GF_OK;
  gf::launch(static_keys_kernel, ib, kCThreads, 0, st, counts, c->num_items, s.k0, s.v0);
  bool r0;
  GF_TRY(radix_sort_pairs(s.k0, s.v0, s.k1, s.v1, c->num_items, 0, 32, s.tmp, &r0, st));
  gf::launch(static_fill_kernel, cdiv(c->capacity * 32, kCThreads), kCThreads, 0, st, *c, r0 ? s.v0 : s.v1, features);
  GF_CUDA(cudaGetLastError());
  return GF_OK;
}

GF_EXPORT int gf_cache_gather(const int64_t *ids, uint64_t n, const uint8_t *cache_flag, const int64_t *cache_map,
                              const float *cache_buffer, const float *features, uint32_t dim, float *out,
                              uint8_t *hit_mask, uint64_t *num_hits, void *stream) {
  return gather_dispatch(ids, n, cache_flag, cache_map, cache_buffer, features, dim, out, hit_mask, num_hits,
                         (cudaStream_t)stream);
}

GF_EXPORT int gf_gather_rows(const int64_t *ids, uint64_t n, const float *features, uint32_t dim, float *out, void *stream) {
  return gather_dispatch(ids, n, nullptr, nullptr, nullptr, features, dim, out, nullptr, nullptr, (cudaStream_t)stream);
}

GF_EXPORT uint64_t gf_cache_update_scratch_bytes(uint64_t n, uint64_t capacity) {
  uint64_t m = upd_elems(n, capacity);
  return 256 + (6 * m + radix_tmp_elems(std::max(n, capacity)) + scan_tmp_elems(std::max(n, capacity)) + 64) * 4;
}

GF_EXPORT int gf_cache_update_lru(gf_cache_state *c, const int64_t *ids, const uint8_t *hit_mask, uint64_t n,
                                  const float *features, void *scratch, uint64_t scratch_bytes, void *stream) {
  return cache_update(c, ids, hit_mask, n, features, kPolicyLru, nullptr, scratch, scratch_bytes, (cudaStream_t)stream);
}

GF_EXPORT int gf_cache_update_fifo(gf_cache_state *c, const int64_t *ids, const uint8_t *hit_mask, uint64_t n,
                                   const float *features, int64_t *pointer, void *scratch, uint64_t scratch_bytes,
                                   void *stream) {
  return cache_update(c, ids, hit_mask, n, features, kPolicyFifo, pointer, scratch, scratch_bytes, (cudaStream_t)stream);
}

GF_EXPORT int gf_host_register(void *ptr, uint64_t bytes, int *owned) {
  if (!ptr || !bytes || !owned) GF_FAIL(GF_EINVAL, "gf_host_register: null argument");
  *owned = 0;
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
  if (e == cudaSuccess) {
    void *dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, ptr, 0) != cudaSuccess || dp != ptr) {  // kernels dereference the host address
      cudaGetLastError();
      cudaHostUnregister(ptr);
      GF_FAIL(GF_EUNSUPPORTED, "registered host memory is not addressable by its host pointer on this platform");
    }
    *owned = 1;
    return GF_OK;
  }
  cudaGetLastError();  // a failed registration must not leak into the caller's next CUDA error check
  if (e == cudaErrorHostMemoryAlreadyRegistered) return GF_OK;
  GF_FAIL(GF_ECUDA, "cudaHostRegister(%p, %llu) failed: %s", ptr, (unsigned long long)bytes, cudaGetErrorString(e));
}

GF_EXPORT int gf_host_unregister(void *ptr) {
  if (!ptr) return GF_OK;
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    GF_FAIL(GF_ECUDA, "cudaHostUnregister(%p) failed: %s", ptr, cudaGetErrorString(e));
  }
  return GF_OK;
}

GF_EXPORT int gf_shared_alloc(int device, uint64_t bytes, void **ptr) {
  if (!ptr || !bytes) GF_FAIL(GF_EINVAL, "gf_shared_alloc: bad argument");
  GF_CUDA(cudaSetDevice(device));
  cudaError_t e = cudaMalloc(ptr, bytes);  // plain cudaMalloc: pool / async allocations cannot be exported over IPC
  if (e != cudaSuccess) {
    cudaGetLastError();
    GF_FAIL(GF_ENOMEM, "gf_shared_alloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e));
  }
  return GF_OK;
}
GF_EXPORT int gf_shared_free(void *ptr) {
  if (ptr) GF_CUDA(cudaFree(ptr));
  return GF_OK;
}
GF_EXPORT int gf_shared_export(void *ptr, void *handle_out) {
  if (!ptr || !handle_out) GF_FAIL(GF_EINVAL, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == GF_PEER_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h;
  GF_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle_out, &h, sizeof(h));
  return GF_OK;
}
GF_EXPORT int gf_shared_open(int device, const void *handle, void **ptr) {
  if (!handle || !ptr) GF_FAIL(GF_EINVAL, "null argument");
  GF_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  GF_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return GF_OK;
}
GF_EXPORT int gf_shared_close(void *ptr) {
  if (ptr) GF_CUDA(cudaIpcCloseMemHandle(ptr));
  return GF_OK;
}

GF_EXPORT int gf_gather_rows_partitioned(const int64_t *ids, uint64_t n, const int8_t *owner, const int32_t *local_index,
                                         uint64_t num_items, const float *const *shards, uint32_t world, uint32_t dim,
                                         float *out, void *stream) {
  if (n == 0) return GF_OK;
  if (!ids || !owner || !local_index || !shards || !out || dim == 0 || world == 0)
    GF_FAIL(GF_EINVAL, "gf_gather_rows_partitioned: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int ROWS = 4;
  const uint64_t warps = (n + ROWS - 1) / ROWS;
  const unsigned blocks = (unsigned)std::min<uint64_t>((warps + kCThreads / 32 - 1) / (kCThreads / 32), 148ull * 16);
  // shard buffers come from cudaMalloc (256-byte aligned); rows are 16-byte aligned iff dim % 4 == 0
  if (dim % 4 == 0 && aligned(out, 16))
    gf::launch(partitioned_gather_kernel<float4, ROWS>, blocks, kCThreads, 0, st, ids, n, owner, local_index, num_items, shards,
               world, dim / 4, (float4 *)out);
  else
    gf::launch(partitioned_gather_kernel<float, ROWS>, blocks, kCThreads, 0, st, ids, n, owner, local_index, num_items, shards,
               world, dim, out);
  GF_CUDA(cudaGetLastError());
  return GF_OK;
}
