// Feature-cache gather and LRU / FIFO policy updates as CUDA kernels.  Replaces the ~12 torch index kernels +
// torch.unique + host index_select + H2D per MFG block of reference gnnflow/cache/cache.py:255-413 and the
// policy updates of lru_cache.py:121-201 / fifo_cache.py:77-161.
#include <algorithm>

#include "gf_primitives.cuh"

namespace gf {

constexpr int kCThreads = 256;

template <typename V>
__device__ __forceinline__ V ld_stream(const V *p) { return __ldg(p); }
template <typename V>
__device__ __forceinline__ void st_stream(V *p, const V &v) { *p = v; }  // default caching: the rows are consumed next

// ---- control block and collect pass of the policy updates (declared first: the gather kernel can run the collect pass
// of a fused fetch itself; the update pipeline is described further down)
struct UpdCtl {        // device control block at the start of the scratch area (zeroed by the call's one memset)
  uint32_t num_miss;   // != 0 iff this fetch had a miss (the reference only updates then, cache.py:317)
  uint32_t num_uniq;   // unique misses
  uint32_t ticket;     // tile ticket of the look-back scan (id spaces too large for the in-kernel scan)
  uint32_t done;       // CTAs of the collect pass that have finished: the last one ranks the bitmap's chunks
  uint32_t done_apply; // CTAs of the apply pass that have finished: the last one advances the FIFO ring pointer
  uint32_t active_passes;  // passes of the victim sort that really run (the key range is known on the device only)
  unsigned long long hits;  // hits of the fused gather (copied to the caller's counter by its last CTA)
  int32_t new_min;          // LRU with a count floor: smallest folded water level of this update (<= 0)
};
constexpr int32_t kLfuMark = 1 << 30;  // static cache statistics: "id seen in this block" (counts stay < 2^30)
enum { kPolicyLru = 0, kPolicyFifo = 1, kPolicyLfu = 2 };
constexpr uint64_t kFusedScanMaxChunks = 8192;  // id spaces of up to 2 M ids: the rank pass's CTAs scan the chunks themselves

struct ChunkPopc {  // scan input: set bits of 256-bit chunk i
  const uint32_t *bitmap;
  __device__ uint32_t operator()(uint64_t i) const {
    const uint4 *p = reinterpret_cast<const uint4 *>(bitmap + i * 8);
    const uint4 a = __ldcg(p), b = __ldcg(p + 1);
    return __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
  }
};
struct ChunkPrefixOut {
  uint32_t *prefix;
  __device__ void operator()(uint64_t i, uint32_t excl, uint32_t) const { prefix[i] = excl; }
};

// What the collect pass leaves behind for one fetch: the misses' bits in a bitmap over the id space (ranking its set
// bits gives torch.unique's sorted, de-duplicated list without a sort), the hit slots' bits in a bitmap over the slots
// (LRU "refresh" / LFU "+1 once per slot"); the CTA that finishes last hands the fetch's hit count to the caller.
struct CollectCtx {
  uint32_t *bitmap;       // null: nothing to collect (gather without a policy update)
  uint32_t *slotbits;     // null for FIFO
  UpdCtl *ctl;            // null: no counters at all
  uint64_t num_items;     // id space of the policy state (ids beyond it are counted by the gather, never admitted)
  uint64_t capacity;      // slots (a hit whose id no longer maps to a slot -- a stale hit mask -- is ignored)
  unsigned long long *hits_out;  // optional: receives ctl->hits (plain store by the last CTA)
};
__device__ __forceinline__ void collect_one(const CollectCtx &cx, uint64_t id, bool hit, const int64_t *__restrict__ map) {
  if (id >= cx.num_items) return;
  if (hit) {
    if (cx.slotbits) {
      const uint64_t slot = (uint64_t)__ldg(map + id);
      if (slot >= cx.capacity) return;
      uint32_t *w = cx.slotbits + (slot >> 5);
      const uint32_t bit = 1u << (slot & 31);
      if (!(*reinterpret_cast<volatile uint32_t *>(w) & bit)) atomicOr(w, bit);
    }
  } else {
    uint32_t *w = cx.bitmap + (id >> 5);
    const uint32_t bit = 1u << (id & 31);
    // hot ids: test first -- thousands of atomics on one word serialise in L2 (a stale read only costs a redundant atomic)
    if (!(*reinterpret_cast<volatile uint32_t *>(w) & bit)) atomicOr(w, bit);
  }
}
// End of the collect pass, called by every thread of every CTA (kCThreads threads).  any_miss: this thread saw a miss.
__device__ __forceinline__ void collect_finish(const CollectCtx &cx, bool any_miss) {
  if (!cx.ctl) return;
  __shared__ uint32_t s_last;
  if (cx.bitmap && __any_sync(0xffffffffu, any_miss) && (threadIdx.x & 31) == 0) cx.ctl->num_miss = 1;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&cx.ctl->done, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (cx.hits_out && threadIdx.x == 0) *cx.hits_out = *reinterpret_cast<volatile unsigned long long *>(&cx.ctl->hits);
}

// out[i,:] = flag[id] ? buffer[map[id],:] : features[id,:].
// A warp takes G consecutive rows: lane l < G resolves row l's source (coalesced id load, then the dependent
// flag -> map chain, once per G rows and in parallel across lanes), then the warp copies the rows ROWS at a time
// with 2 x ROWS independent 16-byte loads in flight per lane before the first store.  G = 32 for large gathers, 8 for
// small ones (so that a batch-sized gather still covers every SM).
template <typename V, int ROWS, int G>
__global__ void __launch_bounds__(kCThreads) cache_gather_kernel(const int64_t *__restrict__ ids, uint64_t n,
                                                                 const uint8_t *__restrict__ flag,
                                                                 const int64_t *__restrict__ map,
                                                                 const V *__restrict__ buffer,
                                                                 const V *__restrict__ features, uint64_t num_items,
                                                                 uint32_t nvec, V *__restrict__ out,
                                                                 uint8_t *__restrict__ hit_mask,
                                                                 unsigned long long *num_hits, unsigned int *num_bad,
                                                                 CollectCtx cx) {
  static_assert(G % ROWS == 0 && G <= 32, "row group");
  bool any_miss = false;
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  unsigned hits = 0;
  for (uint64_t r0 = warp * G; r0 < n; r0 += nwarps * G) {
    const uint64_t i = r0 + lane;
    unsigned long long mine = 0;
    bool bad = false;
    if (lane < G && i < n) {
      const int64_t id = __ldg(ids + i);
      bad = id < 0 || (uint64_t)id >= num_items;  // torch indexing raises IndexError here (cache.py:283): zero row + counter
      const bool hit = !bad && flag ? __ldg(flag + id) != 0 : false;
      mine = bad ? 1ull  // odd: no row address is
                 : (unsigned long long)(uintptr_t)(hit ? buffer + (uint64_t)__ldg(map + id) * nvec : features + (uint64_t)id * nvec);
      hits += hit;
      if (hit_mask) hit_mask[i] = hit;
      if (cx.bitmap && !bad) {  // fused fetch: this kernel is also the collect pass of the policy update
        collect_one(cx, (uint64_t)id, hit, map);
        any_miss |= !hit && (uint64_t)id < cx.num_items;
      }
    }
    if (__any_sync(0xffffffffu, bad) && num_bad && lane == 0) atomicAdd(num_bad, 1u);
    const int cnt = (int)min((uint64_t)G, n - r0);
    for (int rb = 0; rb < cnt; rb += ROWS) {
      const V *srow[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; r++)
        srow[r] = reinterpret_cast<const V *>((uintptr_t)__shfl_sync(0xffffffffu, mine, rb + r));  // null past the end
      V *orow = out + (r0 + rb) * nvec;
      for (uint32_t c = lane; c < nvec; c += 64) {
        const bool two = c + 32 < nvec;
        V v0[ROWS], v1[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; r++)
          if ((uintptr_t)srow[r] == 1) {  // id out of range
            memset(&v0[r], 0, sizeof(V));
            memset(&v1[r], 0, sizeof(V));
          } else if (srow[r]) {
            v0[r] = ld_stream(srow[r] + c);
            if (two) v1[r] = ld_stream(srow[r] + c + 32);
          }
#pragma unroll
        for (int r = 0; r < ROWS; r++)
          if (srow[r]) {
            st_stream(orow + (uint64_t)r * nvec + c, v0[r]);
            if (two) st_stream(orow + (uint64_t)r * nvec + c + 32, v1[r]);
          }
      }
    }
  }
  hits = __reduce_add_sync(0xffffffffu, hits);
  if (num_hits && lane == 0 && hits) atomicAdd(num_hits, (unsigned long long)hits);
  collect_finish(cx, any_miss);
}

template <typename V>
static int launch_gather(const int64_t *ids, uint64_t n, uint64_t num_items, const uint8_t *flag, const int64_t *map,
                         const float *buffer, const float *features, uint32_t dim, float *out, uint8_t *hit_mask,
                         uint64_t *num_hits, uint32_t *num_bad, const CollectCtx &cx, cudaStream_t st) {
  constexpr int ROWS = 4;
  const uint32_t nvec = dim / (sizeof(V) / 4);
  const bool big = n >= 148ull * 64 * 32;  // every SM gets a full complement of 32-row warps
  const uint64_t warps = (n + (big ? 32 : 8) - 1) / (big ? 32 : 8);
  const unsigned blocks = (unsigned)std::min<uint64_t>((warps + kCThreads / 32 - 1) / (kCThreads / 32), 148ull * 16);
  if (big)
    gf::launch(cache_gather_kernel<V, ROWS, 32>, blocks, kCThreads, 0, st, ids, n, flag, map, (const V *)buffer,
               (const V *)features, num_items, nvec, (V *)out, hit_mask, (unsigned long long *)num_hits, num_bad, cx);
  else
    gf::launch(cache_gather_kernel<V, ROWS, 8>, blocks, kCThreads, 0, st, ids, n, flag, map, (const V *)buffer,
               (const V *)features, num_items, nvec, (V *)out, hit_mask, (unsigned long long *)num_hits, num_bad, cx);
  GF_CUDA(cudaGetLastError());
  return GF_OK;
}

static bool aligned(const void *p, size_t a) { return ((uintptr_t)p % a) == 0; }

static int gather_dispatch(const int64_t *ids, uint64_t n, uint64_t num_items, const uint8_t *flag, const int64_t *map,
                           const float *buffer, const float *features, uint32_t dim, float *out, uint8_t *hit_mask,
                           uint64_t *num_hits, uint32_t *num_bad, cudaStream_t st, const CollectCtx *collect = nullptr) {
  if (n == 0) return GF_OK;
  const CollectCtx cx = collect ? *collect : CollectCtx{nullptr, nullptr, nullptr, 0, 0, nullptr};
  if (!ids || !features || !out || dim == 0) GF_FAIL(GF_EINVAL, "gather: null argument");
  if (flag && (!map || !buffer)) GF_FAIL(GF_EINVAL, "gather: cache_flag without cache_map / cache_buffer");
  bool a16 = aligned(features, 16) && aligned(out, 16) && (!flag || aligned(buffer, 16));
  bool a8 = aligned(features, 8) && aligned(out, 8) && (!flag || aligned(buffer, 8));
  if (dim % 4 == 0 && a16)
    return launch_gather<float4>(ids, n, num_items, flag, map, buffer, features, dim, out, hit_mask, num_hits, num_bad, cx, st);
  if (dim % 2 == 0 && a8)
    return launch_gather<float2>(ids, n, num_items, flag, map, buffer, features, dim, out, hit_mask, num_hits, num_bad, cx, st);
  return launch_gather<float>(ids, n, num_items, flag, map, buffer, features, dim, out, hit_mask, num_hits, num_bad, cx, st);
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
// One update = de-duplicated, ascending list of the missed ids (torch.unique, cache.py:290,379) admitted over the
// policy's victims.  The unique list comes from a BITMAP over the id space instead of a sort: misses set their bit,
// one single-pass scan over the 256-bit chunks' popcounts ranks every set bit, and each miss reads its rank back
// (duplicates write the same slot).  The victims of LRU / LFU come from a stable radix sort of the slots by count
// restricted to the bits the counts can occupy (`count_bound`), typically 1-2 passes.  Launches per update (+ one
// memset): FIFO 3, LRU / LFU 3 + P sort passes; inside gf_cache_fetch the first of them is the gather itself.
// collect pass of the stand-alone update calls (the fused fetch does this inside its gather kernel)
__global__ void __launch_bounds__(kCThreads) upd_collect_kernel(const int64_t *__restrict__ ids,
                                                                const uint8_t *__restrict__ hit_mask, uint64_t n,
                                                                const int64_t *__restrict__ map, CollectCtx cx) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool miss = false;
  if (i < n) {
    const uint64_t id = (uint64_t)ids[i];
    const bool hit = hit_mask[i] != 0;
    miss = !hit && id < cx.num_items;  // ids outside the table: counted by the gather
    collect_one(cx, id, hit, map);
  }
  collect_finish(cx, miss);
}

// Second pass, two roles in one launch (both only read what the collect pass wrote):
//   CTAs [0, id_ctas): misses: uniq[rank of the id among the set bits] = id (ranks >= kmax are not admitted);
//   the others (LRU / LFU): fold this fetch into the counts -- LRU: count -= 1 everywhere, hit slots -> 0
//   (lru_cache.py:142-145); LFU: hit slots += 1, once per slot (the reference's non-accumulating index_put,
//   lfu_cache.py:158) -- and emit the victim sort's keys together with the digit histograms of all its passes.
//   key = count + bound (all counts lie in [-bound, bound]); bound == 0: full 32-bit order.
struct RankKeysArgs {
  const int64_t *ids;
  const uint8_t *hit_mask;
  uint64_t n;
  const uint32_t *bitmap, *chunk_prefix, *slotbits;
  uint32_t *uniq;
  uint32_t kmax;
  uint64_t num_items;
  int32_t *count;
  uint64_t capacity;
  UpdCtl *ctl;
  int policy;
  uint32_t bound;
  int passes;
  uint32_t *keys, *vals, *ghist;
  unsigned id_ctas;
  const int32_t *floor;  // LRU, optional: every water level is >= *floor before this update
  uint64_t chunks;       // 256-bit chunks of the bitmap
  int local_scan;        // every id CTA scans the chunks itself (else chunk_prefix holds the result of scan_lookback_kernel)
};
__global__ void __launch_bounds__(kCThreads) upd_rank_keys_kernel(RankKeysArgs a) {
  if (!a.ctl->num_miss) return;
  if (blockIdx.x < a.id_ctas) {
    // Exclusive prefix of the bitmap's set bits per 256-bit chunk.  Id spaces of up to kFusedScanMaxChunks chunks: EVERY id
    // CTA scans all chunks itself into shared memory (coalesced popcounts, one block scan: a few L2 round trips run
    // side by side in all CTAs -- the same scan done once by the last CTA of the collect pass was a 10 us serial tail);
    // larger id spaces: scan_lookback_kernel has written chunk_prefix.
    __shared__ uint32_t s_pre[kFusedScanMaxChunks];
    __shared__ uint32_t s_total;
    const bool local = a.local_scan != 0;
    if (local) {
      const ChunkPopc popc{a.bitmap};
#pragma unroll 4
      for (uint64_t c = threadIdx.x; c < a.chunks; c += kCThreads) s_pre[c] = popc(c);
      __syncthreads();
      const uint64_t per = (a.chunks + kCThreads - 1) / kCThreads;
      const uint64_t c0 = min(a.chunks, (uint64_t)threadIdx.x * per), c1 = min(a.chunks, c0 + per);
      uint32_t mine = 0;
      for (uint64_t c = c0; c < c1; c++) mine += s_pre[c];
      uint32_t off = block_excl_scan(mine, &s_total);
      for (uint64_t c = c0; c < c1; c++) {
        const uint32_t v = s_pre[c];
        s_pre[c] = off;
        off += v;
      }
      __syncthreads();
      if (blockIdx.x == 0 && threadIdx.x == 0) a.ctl->num_uniq = s_total;
    }
    const uint64_t i = (uint64_t)blockIdx.x * kCThreads + threadIdx.x;
    if (i >= a.n || a.hit_mask[i]) return;
    const uint64_t id = (uint64_t)a.ids[i];
    if (id >= a.num_items) return;
    const uint64_t w = id >> 5, c = w >> 3;
    uint32_t rank = local ? s_pre[c] : a.chunk_prefix[c];
    for (uint64_t q = c << 3; q < w; q++) rank += __popc(a.bitmap[q]);  // same 32-byte sector as word w
    rank += __popc(a.bitmap[w] & ((1u << (id & 31)) - 1u));
    if (rank < a.kmax) a.uniq[rank] = (uint32_t)id;
    return;
  }
  __shared__ uint32_t hist[kSortMaxPasses][256];
  for (int i = threadIdx.x; i < kSortMaxPasses * 256; i += kCThreads) (&hist[0][0])[i] = 0;
  __syncthreads();
  // LRU with a count floor: the folded water levels lie in [lo, 0], a range that is usually far narrower than the static
  // bound (slots are evicted oldest first: the oldest live one is about capacity / admissions-per-update updates old),
  // so the sort needs fewer passes than the host had to launch; the surplus ones return at once.
  const bool dyn = a.floor != nullptr && a.policy == kPolicyLru;
  const int32_t lo = dyn ? *a.floor - 1 : 0;
  int passes = a.passes;
  if (dyn) {
    int bits = 1;
    while ((1ll << bits) <= -(long long)lo) bits++;
    passes = min(passes, (bits + 7) / 8);
  }
  if (blockIdx.x == a.id_ctas && threadIdx.x == 0) a.ctl->active_passes = (uint32_t)passes;
  int32_t mn = 0;
  const uint64_t stride = (uint64_t)(gridDim.x - a.id_ctas) * kCThreads;
  for (uint64_t i = (uint64_t)(blockIdx.x - a.id_ctas) * kCThreads + threadIdx.x; i < a.capacity; i += stride) {
    int32_t c = a.count[i];
    const bool hit = (a.slotbits[i >> 5] >> (i & 31)) & 1u;
    if (a.policy == kPolicyLru) c = hit ? 0 : c - 1;
    else c += hit ? 1 : 0;
    a.count[i] = c;
    mn = min(mn, c);
    const uint32_t key = dyn ? (uint32_t)(c - lo) : (a.bound ? (uint32_t)(c + (int32_t)a.bound) : ((uint32_t)c ^ 0x80000000u));
    a.keys[i] = key;
    a.vals[i] = (uint32_t)i;
    for (int p = 0; p < passes; p++) atomicAdd(&hist[p][(key >> (8 * p)) & 255u], 1u);
  }
  if (dyn) {
    mn = __reduce_min_sync(0xffffffffu, mn);
    if ((threadIdx.x & 31) == 0 && mn < 0) atomicMin(&a.ctl->new_min, mn);
  }
  __syncthreads();
  for (int p = 0; p < passes; p++) {
    const uint32_t c = hist[p][threadIdx.x];
    if (c) atomicAdd(&a.ghist[p * 256 + threadIdx.x], c);
  }
}
__device__ __forceinline__ uint64_t fifo_slot(int64_t ptr, int64_t cap, int64_t k, int64_t j) {
  if (ptr + k < cap) return (uint64_t)(ptr + 1 + j);
  const int64_t r = k - (cap - 1 - ptr);  // wrap: [0, r) ++ [ptr + 1, cap), fifo_cache.py:101-103
  return j < r ? (uint64_t)j : (uint64_t)(ptr + 1 + (j - r));
}
// admit uniq[j] into slot victim(j), j < k  (lru_cache.py:151-160 / fifo_cache.py:106-116 / lfu_cache.py:161-172).
// One warp per admission: evict the slot's old id, copy the row, publish the new id.  Evicted ids (cached) and admitted
// ids (missed) are disjoint sets, so the flag / map writes of different warps never touch the same id; the eviction is
// guarded (the old id must still map to this slot) so that a stale index_to_id entry cannot unpublish anybody.
// FIFO: the last CTA to finish advances the ring pointer (every warp has read it by then).
template <bool FIFO>
__global__ void __launch_bounds__(kCThreads) upd_apply_kernel(gf_cache_state c, const uint32_t *__restrict__ uniq,
                                                              const uint32_t *__restrict__ v0, const uint32_t *__restrict__ v1,
                                                              const float *__restrict__ features, UpdCtl *ctl,
                                                              int64_t *fifo_ptr, int32_t admit_count, int32_t *floor) {
  const int lane = threadIdx.x & 31;
  // the sorted slots are in (k1, v1) after an odd number of passes
  const uint32_t *victims = FIFO ? nullptr : ((ctl->active_passes & 1u) ? v1 : v0);
  if (!FIFO && floor && blockIdx.x == 0 && threadIdx.x == 0 && ctl->num_miss) *floor = ctl->new_min;
  const uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t k = (uint32_t)min((uint64_t)ctl->num_uniq, c.capacity);
  const int64_t ptr = FIFO ? *reinterpret_cast<volatile int64_t *>(fifo_ptr) : 0;
  if (j < k) {
    const uint64_t slot = FIFO ? fifo_slot(ptr, (int64_t)c.capacity, k, (int64_t)j) : victims[j];
    const int64_t new_id = uniq[j];
    if (lane == 0) {
      const int64_t old_id = c.index_to_id[slot];
      if (old_id >= 0 && (uint64_t)old_id < c.num_items && c.flag[old_id] && c.map[old_id] == (int64_t)slot) {
        c.flag[old_id] = 0;
        c.map[old_id] = -1;
      }
    }
    const float *src = features + (uint64_t)new_id * c.dim;
    float *dst = c.buffer + slot * c.dim;
    for (uint32_t d = lane; d < c.dim; d += 32) dst[d] = __ldg(src + d);
    if (lane == 0) {
      if (!FIFO) c.count[slot] = admit_count;  // LRU: 0 (lru_cache.py:153), LFU: 1 (lfu_cache.py:166)
      c.index_to_id[slot] = new_id;
      c.flag[new_id] = 1;
      c.map[new_id] = (int64_t)slot;
    }
  }
  if (FIFO) {  // fifo_cache.py:98-105
    __shared__ uint32_t s_last;
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ctl->done_apply, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
      const int64_t cap = (int64_t)c.capacity;
      *fifo_ptr = k == 0 ? ptr : (ptr + k < cap ? ptr + k : k - (cap - 1 - ptr) - 1);
    }
  }
}

struct UpdScratch {
  UpdCtl *ctl;
  unsigned long long *status;  // look-back status words of the chunk scan (large id spaces)
  uint32_t *bitmap;            // one bit per id, in 256-bit chunks
  uint32_t *slotbits;          // one bit per slot
  uint32_t *sort_tmp;          // control words + digit histograms of the victim sort (radix_sort_pairs_prepared)
  size_t zero_bytes;           // everything above: cleared by ONE memset per call
  uint8_t *hit_mask;           // fused fetch: the gather's hit mask stays in here
  uint32_t *chunk_prefix, *uniq, *k0, *v0, *k1, *v1;
  size_t total_bytes;
};
static UpdScratch carve_upd(void *scratch, uint64_t n, uint64_t capacity, uint64_t num_items, int passes = kSortMaxPasses) {
  const uint64_t chunks = (num_items + 255) / 256, tiles = (chunks + kScanTile - 1) / kScanTile;
  const uint64_t kmax = std::min(n, capacity), m = align_up(capacity + 1, 64);
  UpdScratch s;
  char *p = reinterpret_cast<char *>(scratch);
  s.ctl = reinterpret_cast<UpdCtl *>(p); p += 256;
  s.status = reinterpret_cast<unsigned long long *>(p); p += align_up(tiles * 8, 256);
  s.bitmap = reinterpret_cast<uint32_t *>(p); p += align_up(chunks * 32, 256);
  s.slotbits = reinterpret_cast<uint32_t *>(p); p += align_up((capacity + 31) / 32 * 4 + 4, 256);
  s.sort_tmp = reinterpret_cast<uint32_t *>(p);
  s.zero_bytes = (size_t)(p - reinterpret_cast<char *>(scratch)) + radix_ctl_bytes(capacity, passes);  // the passes in use
  p += align_up((radix_tmp_elems(capacity) + 64) * 4, 256);
  s.hit_mask = reinterpret_cast<uint8_t *>(p); p += align_up(n + 1, 256);
  s.chunk_prefix = reinterpret_cast<uint32_t *>(p); p += align_up(chunks * 4, 256);
  s.uniq = reinterpret_cast<uint32_t *>(p); p += align_up((kmax + 1) * 4, 256);
  s.k0 = reinterpret_cast<uint32_t *>(p); p += m * 4;
  s.v0 = reinterpret_cast<uint32_t *>(p); p += m * 4;
  s.k1 = reinterpret_cast<uint32_t *>(p); p += m * 4;
  s.v1 = reinterpret_cast<uint32_t *>(p); p += m * 4;
  s.total_bytes = (size_t)(p - reinterpret_cast<char *>(scratch));
  return s;
}

static int check_update_args(gf_cache_state *c, int policy, int64_t *fifo_ptr, const void *scratch) {
  const bool fifo = policy == kPolicyFifo;
  if (!c->buffer || !c->flag || !c->map || !c->index_to_id || (!fifo && !c->count) || (fifo && !fifo_ptr))
    GF_FAIL(GF_EINVAL, "cache update: incomplete cache state");
  if (c->num_items >= (1ull << 32) || c->capacity >= (1ull << 31)) GF_FAIL(GF_EINVAL, "cache too large");
  if (((uintptr_t)scratch & 255) != 0) GF_FAIL(GF_EINVAL, "cache update: scratch must be 256-byte aligned");
  return GF_OK;
}
// bits the counts can occupy, and the matching passes of the victim sort
static void victim_sort_shape(int policy, uint64_t count_bound, uint32_t *bound, int *passes) {
  int bits = 32;
  *bound = 0;
  if (count_bound && count_bound < (1ull << 29)) {
    *bound = (uint32_t)count_bound;
    bits = 1;
    while ((1ull << bits) <= 2ull * *bound) bits++;
  }
  *passes = policy == kPolicyFifo ? 0 : (bits + 7) / 8;
}
static CollectCtx make_collect(const UpdScratch &s, const gf_cache_state *c, int policy, unsigned long long *hits_out) {
  CollectCtx cx = {s.bitmap, policy == kPolicyFifo ? nullptr : s.slotbits, s.ctl, c->num_items, c->capacity, hits_out};
  return cx;
}
// Everything of an update after the collect pass (which has run on `st`, over a scratch area cleared by the caller):
// [chunk scan for large id spaces] -> rank + keys -> victim sort passes -> apply.  2 + P launches (FIFO: 2).
static int cache_update_tail(gf_cache_state *c, const UpdScratch &s, const int64_t *ids, const uint8_t *hit_mask, uint64_t n,
                             const float *features, int policy, int64_t *fifo_ptr, uint64_t count_bound, int32_t *floor,
                             cudaStream_t st) {
  const bool fifo = policy == kPolicyFifo, lfu = policy == kPolicyLfu;
  if (policy != kPolicyLru) floor = nullptr;
  const uint64_t chunks = (c->num_items + 255) / 256, kmax = std::min<uint64_t>(n, c->capacity);
  const unsigned nb = cdiv(n, kCThreads);
  // chunk prefixes: scanned by every id CTA of the rank pass itself (batch-sized fetches over id spaces of up to 2 M ids:
  // no launch, no serial tail), or once by the look-back scan in its own launch (large fetches, where thousands of id CTAs
  // would each re-read the whole bitmap, and large id spaces)
  const bool local_scan = chunks <= kFusedScanMaxChunks && nb <= 2u * 148u;
  if (!local_scan) {
    LookbackCtl lb = {&s.ctl->ticket, s.status, 1ull};
    gf::launch(scan_lookback_kernel<ChunkPopc, ChunkPrefixOut>, cdiv(chunks, kScanTile), kScanThreads, 0, st, chunks,
               ChunkPopc{s.bitmap}, ChunkPrefixOut{s.chunk_prefix}, lb, &s.ctl->num_uniq);
  }
  uint32_t bound;
  int passes;
  victim_sort_shape(policy, count_bound, &bound, &passes);
  const unsigned cb = fifo ? 0u : std::min<unsigned>(cdiv(c->capacity, kCThreads), 148u * 4);
  RankKeysArgs a = {ids, hit_mask, n, s.bitmap, s.chunk_prefix, s.slotbits, s.uniq, (uint32_t)kmax, c->num_items, c->count,
                    c->capacity, s.ctl, policy, bound, passes, s.k0, s.v0, s.sort_tmp, nb, floor, chunks, local_scan ? 1 : 0};
  gf::launch(upd_rank_keys_kernel, nb + cb, kCThreads, 0, st, a);
  if (!fifo) {  // k smallest water levels / use counts, ties -> lowest slot: stable sort of the slots by count
    bool r0;
    GF_TRY(radix_sort_pairs_prepared(s.k0, s.v0, s.k1, s.v1, c->capacity, 0, passes, s.sort_tmp, &r0, st, &s.ctl->active_passes));
  }
  if (fifo)
    gf::launch(upd_apply_kernel<true>, cdiv(kmax * 32, kCThreads), kCThreads, 0, st, *c, s.uniq, (const uint32_t *)nullptr,
               (const uint32_t *)nullptr, features, s.ctl, fifo_ptr, 0, (int32_t *)nullptr);
  else
    gf::launch(upd_apply_kernel<false>, cdiv(kmax * 32, kCThreads), kCThreads, 0, st, *c, s.uniq, (const uint32_t *)s.v0,
               (const uint32_t *)s.v1, features, s.ctl, fifo_ptr, lfu ? 1 : 0, floor);
  GF_CUDA(cudaGetLastError());
  return GF_OK;
}

static int cache_update(gf_cache_state *c, const int64_t *ids, const uint8_t *hit_mask, uint64_t n, const float *features,
                        int policy, int64_t *fifo_ptr, uint64_t count_bound, void *scratch, uint64_t scratch_bytes,
                        cudaStream_t st) {
  if (!c || !ids || !hit_mask || !features || !scratch) GF_FAIL(GF_EINVAL, "cache update: null argument");
  GF_TRY(check_update_args(c, policy, fifo_ptr, scratch));
  if (n == 0 || c->capacity == 0) return GF_OK;
  uint32_t bound_;
  int passes_;
  victim_sort_shape(policy, count_bound, &bound_, &passes_);
  UpdScratch s = carve_upd(scratch, n, c->capacity, c->num_items, passes_);
  if (scratch_bytes < s.total_bytes) GF_FAIL(GF_ECAPACITY, "cache update: scratch too small");
  GF_CUDA(cudaMemsetAsync(scratch, 0, s.zero_bytes, st));
  gf::launch(upd_collect_kernel, cdiv(n, kCThreads), kCThreads, 0, st, ids, hit_mask, n, c->map,
             make_collect(s, c, policy, nullptr));
  return cache_update_tail(c, s, ids, hit_mask, n, features, policy, fifo_ptr, count_bound, nullptr, st);
}

// ------------------------------------------------------------------------------- sorted unique + inverse map
// torch.unique(ids, return_inverse=True) over a bounded id space (Memory.prepare_input, models/modules/memory.py:170-171,
// where the reference first copies all_nodes to the host; cache.py:290,355,379) with the bitmap ranking above instead of
// a sort: every id sets its bit, one single-pass scan ranks the set bits, every id reads its rank back.  3 launches +
// one memset, O(n + num_items / 8) bytes, ascending output like torch.unique(sorted=True).
__global__ void __launch_bounds__(kCThreads) uniq_collect_kernel(const int64_t *__restrict__ ids, uint64_t n,
                                                                 uint64_t num_items, uint32_t *bitmap) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t id = (uint64_t)ids[i];
  if (id >= num_items) return;
  uint32_t *w = bitmap + (id >> 5);
  const uint32_t bit = 1u << (id & 31);
  // hot ids: test first -- thousands of atomics on one word serialise in L2 (a stale read only costs a redundant atomic)
  if (!(*reinterpret_cast<volatile uint32_t *>(w) & bit)) atomicOr(w, bit);
}
__global__ void __launch_bounds__(kCThreads) uniq_rank_kernel(const int64_t *__restrict__ ids, uint64_t n,
                                                              uint64_t num_items, const uint32_t *__restrict__ bitmap,
                                                              const uint32_t *__restrict__ chunk_prefix,
                                                              int64_t *unique_out, int64_t *inverse_out,
                                                              const uint32_t *num_uniq, uint64_t *count_out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && count_out) *count_out = *num_uniq;
  if (i >= n) return;
  const uint64_t id = (uint64_t)ids[i];
  if (id >= num_items) {  // outside the id space: no rank (precondition violated by the caller)
    if (inverse_out) inverse_out[i] = -1;
    return;
  }
  const uint64_t w = id >> 5, c = w >> 3;
  uint32_t rank = chunk_prefix[c];
  for (uint64_t q = c << 3; q < w; q++) rank += __popc(bitmap[q]);  // same 32-byte sector as word w
  rank += __popc(bitmap[w] & ((1u << (id & 31)) - 1u));
  if (inverse_out) inverse_out[i] = (int64_t)rank;
  if (unique_out) unique_out[rank] = (int64_t)id;  // duplicates write the same value
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

GF_EXPORT uint64_t gf_unique_scratch_bytes(uint64_t num_items) {
  const uint64_t chunks = (num_items + 255) / 256, tiles = (chunks + kScanTile - 1) / kScanTile;
  return 256 + align_up(tiles * 8, 256) + align_up(chunks * 32, 256) + align_up(chunks * 4, 256) + 256;
}

GF_EXPORT int gf_unique_inverse(const int64_t *ids, uint64_t n, uint64_t num_items, int64_t *unique_out, int64_t *inverse_out,
                                uint64_t *count_out, void *scratch, uint64_t scratch_bytes, void *stream) {
  if (n && !ids) GF_FAIL(GF_EINVAL, "unique_inverse: null ids");
  if (!scratch || ((uintptr_t)scratch & 255) != 0) GF_FAIL(GF_EINVAL, "unique_inverse: scratch must be 256-byte aligned");
  if (num_items >= (1ull << 32)) GF_FAIL(GF_EINVAL, "unique_inverse: id space too large");
  if (scratch_bytes < gf_unique_scratch_bytes(num_items)) GF_FAIL(GF_ECAPACITY, "unique_inverse: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t chunks = (num_items + 255) / 256, tiles = (chunks + kScanTile - 1) / kScanTile;
  char *p = reinterpret_cast<char *>(scratch);
  UpdCtl *ctl = reinterpret_cast<UpdCtl *>(p); p += 256;
  unsigned long long *status = reinterpret_cast<unsigned long long *>(p); p += align_up(tiles * 8, 256);
  uint32_t *bitmap = reinterpret_cast<uint32_t *>(p); p += align_up(chunks * 32, 256);
  const size_t zero_bytes = (size_t)(p - reinterpret_cast<char *>(scratch));
  uint32_t *chunk_prefix = reinterpret_cast<uint32_t *>(p);
  GF_CUDA(cudaMemsetAsync(scratch, 0, zero_bytes, st));
  if (n == 0 || chunks == 0) {
    if (count_out) GF_CUDA(cudaMemsetAsync(count_out, 0, sizeof(uint64_t), st));
    return GF_OK;
  }
  const unsigned nb = cdiv(n, kCThreads);
  gf::launch(uniq_collect_kernel, nb, kCThreads, 0, st, ids, n, num_items, bitmap);
  LookbackCtl lb = {&ctl->ticket, status, 1ull};
  gf::launch(scan_lookback_kernel<ChunkPopc, ChunkPrefixOut>, cdiv(chunks, kScanTile), kScanThreads, 0, st, chunks,
             ChunkPopc{bitmap}, ChunkPrefixOut{chunk_prefix}, lb, &ctl->num_uniq);
  gf::launch(uniq_rank_kernel, nb, kCThreads, 0, st, ids, n, num_items, bitmap, chunk_prefix, unique_out, inverse_out,
             &ctl->num_uniq, count_out);
  GF_CUDA(cudaGetLastError());
  return GF_OK;
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

GF_EXPORT uint64_t gf_cache_fill_scratch_bytes(uint64_t num_items) {
  return (4 * align_up(num_items + 1, 64) + radix_tmp_elems(num_items) + 64) * 4 + 256;
}

GF_EXPORT int gf_cache_fill_topk(gf_cache_state *c, const int32_t *counts, const float *features, void *scratch,
                                 uint64_t scratch_bytes, void *stream) {
  if (!c || !counts || !features || !scratch) GF_FAIL(GF_EINVAL, "fill_topk: null argument");
  if ((c->capacity && !c->buffer) || !c->flag || !c->map) GF_FAIL(GF_EINVAL, "fill_topk: incomplete cache state");
  if (c->num_items >= (1ull << 32) || c->capacity > c->num_items) GF_FAIL(GF_EINVAL, "fill_topk: bad capacity / num_items");
  if (scratch_bytes < gf_cache_fill_scratch_bytes(c->num_items)) GF_FAIL(GF_ECAPACITY, "fill_topk: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->num_items == 0) return GF_OK;
  const uint64_t m = align_up(c->num_items + 1, 64);
  uint32_t *k0 = reinterpret_cast<uint32_t *>(scratch), *v0 = k0 + m, *k1 = v0 + m, *v1 = k1 + m, *tmp = v1 + m;
  const unsigned ib = cdiv(c->num_items, kCThreads);
  gf::launch(static_clear_kernel, ib, kCThreads, 0, st, *c);
  if (c->capacity == 0) return GF_OK;
  gf::launch(static_keys_kernel, ib, kCThreads, 0, st, counts, c->num_items, k0, v0);
  bool r0;
  GF_TRY(radix_sort_pairs(k0, v0, k1, v1, c->num_items, 0, 32, tmp, &r0, st));
  gf::launch(static_fill_kernel, cdiv(c->capacity * 32, kCThreads), kCThreads, 0, st, *c, r0 ? v0 : v1, features);
  GF_CUDA(cudaGetLastError());
  return GF_OK;
}

GF_EXPORT int gf_cache_gather(const int64_t *ids, uint64_t n, uint64_t num_items, const uint8_t *cache_flag,
                              const int64_t *cache_map, const float *cache_buffer, const float *features, uint32_t dim,
                              float *out, uint8_t *hit_mask, uint64_t *num_hits, uint32_t *num_bad, void *stream) {
  return gather_dispatch(ids, n, num_items, cache_flag, cache_map, cache_buffer, features, dim, out, hit_mask, num_hits,
                         num_bad, (cudaStream_t)stream);
}

GF_EXPORT int gf_gather_rows(const int64_t *ids, uint64_t n, uint64_t num_items, const float *features, uint32_t dim,
                             float *out, uint32_t *num_bad, void *stream) {
  return gather_dispatch(ids, n, num_items, nullptr, nullptr, nullptr, features, dim, out, nullptr, nullptr, num_bad,
                         (cudaStream_t)stream);
}

GF_EXPORT uint64_t gf_cache_update_scratch_bytes(uint64_t n, uint64_t capacity, uint64_t num_items) {
  return carve_upd(nullptr, n, capacity, num_items).total_bytes;
}

GF_EXPORT int gf_cache_update_lru(gf_cache_state *c, const int64_t *ids, const uint8_t *hit_mask, uint64_t n,
                                  const float *features, uint64_t count_bound, void *scratch, uint64_t scratch_bytes,
                                  void *stream) {
  return cache_update(c, ids, hit_mask, n, features, kPolicyLru, nullptr, count_bound, scratch, scratch_bytes, (cudaStream_t)stream);
}

GF_EXPORT int gf_cache_update_lfu(gf_cache_state *c, const int64_t *ids, const uint8_t *hit_mask, uint64_t n,
                                  const float *features, uint64_t count_bound, void *scratch, uint64_t scratch_bytes,
                                  void *stream) {
  return cache_update(c, ids, hit_mask, n, features, kPolicyLfu, nullptr, count_bound, scratch, scratch_bytes, (cudaStream_t)stream);
}

GF_EXPORT int gf_cache_update_fifo(gf_cache_state *c, const int64_t *ids, const uint8_t *hit_mask, uint64_t n,
                                   const float *features, int64_t *pointer, void *scratch, uint64_t scratch_bytes,
                                   void *stream) {
  return cache_update(c, ids, hit_mask, n, features, kPolicyFifo, pointer, 0, scratch, scratch_bytes, (cudaStream_t)stream);
}

GF_EXPORT int gf_cache_fetch(gf_cache_state *c, const int64_t *ids, uint64_t n, const float *features, uint64_t feature_rows,
                             int policy, int64_t *fifo_pointer, uint64_t count_bound, int32_t *count_floor, int update,
                             float *out, uint64_t *hits_out, uint32_t *num_bad, void *scratch, uint64_t scratch_bytes,
                             void *stream) {
  if (!c || (n && (!ids || !features || !out))) GF_FAIL(GF_EINVAL, "cache fetch: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) return GF_OK;
  const uint64_t limit = std::min<uint64_t>(c->num_items, feature_rows);
  const bool cached = c->capacity > 0 && c->flag && c->map && c->buffer;
  if (!cached) {  // nothing can hit
    if (hits_out) GF_CUDA(cudaMemsetAsync(hits_out, 0, sizeof(uint64_t), st));
    return gather_dispatch(ids, n, limit, nullptr, nullptr, nullptr, features, c->dim, out, nullptr, nullptr, num_bad, st);
  }
  if (policy != kPolicyLru && policy != kPolicyFifo && policy != kPolicyLfu) update = 0;  // static cache: gather only
  if (!scratch) GF_FAIL(GF_EINVAL, "cache fetch: null scratch");
  if (update) GF_TRY(check_update_args(c, policy, fifo_pointer, scratch));
  uint32_t bound_;
  int passes_;
  victim_sort_shape(policy, count_bound, &bound_, &passes_);
  UpdScratch s = carve_upd(scratch, n, c->capacity, c->num_items, passes_);
  if (scratch_bytes < (update ? s.total_bytes : (size_t)256)) GF_FAIL(GF_ECAPACITY, "cache fetch: scratch too small");
  GF_CUDA(cudaMemsetAsync(scratch, 0, update ? s.zero_bytes : (size_t)256, st));
  CollectCtx cx = make_collect(s, c, policy, reinterpret_cast<unsigned long long *>(hits_out));
  if (!update) cx.bitmap = cx.slotbits = nullptr;
  GF_TRY(gather_dispatch(ids, n, limit, c->flag, c->map, c->buffer, features, c->dim, out, update ? s.hit_mask : nullptr,
                         reinterpret_cast<uint64_t *>(&s.ctl->hits), num_bad, st, &cx));
  if (!update) return GF_OK;
  return cache_update_tail(c, s, ids, s.hit_mask, n, features, policy, fifo_pointer, count_bound, count_floor, st);
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
