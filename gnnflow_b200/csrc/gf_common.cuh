// Shared definitions of the B200-native store / sampler / cache library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gnnflow_b200.h"

#define GF_EXPORT extern "C" __attribute__((visibility("default")))

namespace gf {

// ---- error plumbing ---------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
const char *get_error();

#define GF_FAIL(code, ...)      \
  do {                          \
    gf::set_error(__VA_ARGS__); \
    return (code);              \
  } while (0)

#define GF_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      gf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);   \
      return _e == cudaErrorMemoryAllocation ? GF_ENOMEM : GF_ECUDA;                               \
    }                                                                                              \
  } while (0)

#define GF_TRY(expr)           \
  do {                         \
    int _rc = (expr);          \
    if (_rc != GF_OK) return _rc; \
  } while (0)

// ---- device data layout -----------------------------------------------------------------------------
// One TemporalBlock (reference gnnflow/csrc/common.h:35-48) is a 32-byte descriptor -- exactly one DRAM
// sector -- plus one contiguous 128-byte-aligned payload:  ts[cap] | {dst, eid}[cap]  (the timestamps stay a dense
// array for the searches; a neighbour's id and edge id sit side by side, one 128-bit load and ONE line instead of two).  The descriptors of one vertex sit in a
// per-vertex directory array ordered oldest -> newest (instead of the reference's prev/next pointers), and
// carry the running edge count of the older blocks so that a window [start, end) maps to one contiguous
// range of "positions" without walking the list.
struct __align__(32) BlockDesc {
  uint64_t payload;     // device address of the payload
  uint32_t size;        // edges stored
  uint32_t capacity;    // edges that fit
  float start_ts;       // min timestamp in the block (FLT_MAX when empty)
  float end_ts;         // max (= last) timestamp in the block
  uint32_t cum_before;  // edges stored in the older blocks of this vertex (position of element 0)
  float min_ts;         // in the NEWEST descriptor of a vertex: start_ts of its oldest live block (window-start shortcut)
};
static_assert(sizeof(BlockDesc) == 32, "BlockDesc must be one 32-byte sector");

// One vertex (reference DoublyLinkedList + HostDoublyLinkedList, doubly_linked_list.h:15-34): 64 bytes = two sectors of
// one 128-byte line.  The second sector is a copy of the vertex's NEWEST block descriptor (dir[end - 1]): a sampling
// target then reaches the block that almost always holds its neighbours with ONE dependent load after the vertex id
// (root -> entry) instead of two (root -> entry -> descriptor); the ingest path, which rewrites the entry of every
// vertex it touches anyway, keeps the copy in sync.
struct __align__(64) NodeEntry {
  uint64_t dir_tagged;      // device address of the BlockDesc directory (128-byte aligned) | log2(dir_cap); 0 = never had a block
  uint32_t first;           // oldest live block (blocks before it were offloaded)
  uint32_t end;             // one past the newest block; live blocks are [first, end)
  uint32_t cum_first;       // cum_before of dir[first]: position of the oldest stored edge (positions are relative)
  uint32_t num_insertions;  // HostDoublyLinkedList::num_insertions
  uint64_t num_edges;       // HostDoublyLinkedList::num_edges == out_degree (never decremented)
  BlockDesc tail;           // == dir[end - 1] when end > first

  __host__ __device__ uint64_t dir() const { return dir_tagged & ~127ull; }
  __host__ __device__ uint32_t dir_cap() const { return dir_tagged ? 1u << (uint32_t)(dir_tagged & 127ull) : 0u; }
};
static_assert(sizeof(NodeEntry) == 64, "NodeEntry must be two 32-byte sectors");

constexpr uint32_t kUnit = 128;  // allocation granule of the payload arena, bytes

// Payload of one block (128-byte aligned):
//   ts[cap] | {dst, eid}[cap] | pivot levels K, K-1, ..., 1
// (a DRAM miss brings in the whole 128-byte line whatever the request -- measured, profiles/r02_ncu_sampler_hbm_gdelt16k.txt:
// 3.4 sectors per L2 read request of the uniform sampler on a graph that does not fit the L2 -- so what one sampled
// neighbour needs should lie in as few lines as possible: two here, three with separate dst[] and eid[] arrays)
// Pivot level k holds every 8^k-th timestamp, piv_k[j] = ts[(j + 1) * 8^k - 1] (the last element of the j-th complete
// run of 8^k edges), so that a lower-bound search touches ONE 32-byte sector per level instead of one sector per
// binary-search probe: the top level has at most kPivTop entries (two sectors, loaded together), every level below
// it is one 256-bit load.  cap <= 16: no pivots; <= 128: 1 level; <= 1024: 2; <= 8192: 3.  Overhead ~0.6 B / edge.
// Levels are stored top level first so that a descending search just advances a pointer.
constexpr uint32_t kPivTop = 16;

__host__ __device__ inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// Size classes of the payload / directory arena (TemporalBlockAllocator::Allocate / Deallocate,
// temporal_block_allocator.cu:90-180, on top of rmm's pool): an allocation of u units of 128 bytes is rounded up to the
// class size -- exact up to 16 units, then 8 steps per power of two (<= 12.5 % slack, 0 for the minimum-size blocks of
// the named datasets) -- so that a freed block serves any later request of its class.
constexpr uint32_t kNumClasses = 224;
__host__ __device__ inline uint32_t ilog2_u32(uint32_t x) {  // floor(log2(x)), x >= 1
  uint32_t r = 0;
  while (x >>= 1) r++;
  return r;
}
__host__ __device__ inline uint32_t class_of_units(uint32_t u) {  // u >= 1
  if (u <= 16) return u - 1;
  const uint32_t e = ilog2_u32(u - 1);  // u in (2^e, 2^(e+1)], e >= 4
  const uint32_t step = 1u << (e - 3);
  const uint32_t k = (u - (1u << e) + step - 1) / step;  // 1 .. 8
  return 16 + (e - 4) * 8 + (k - 1);
}
__host__ __device__ inline uint32_t class_units(uint32_t c) {
  if (c < 16) return c + 1;
  const uint32_t e = 4 + (c - 16) / 8, k = (c - 16) % 8 + 1;
  return (1u << e) + k * (1u << (e - 3));
}
__host__ __device__ inline uint32_t piv_levels(uint32_t cap) {
  uint32_t k = 0;
  while (cap > kPivTop) {
    cap >>= 3;
    k++;
  }
  return k;
}
__host__ __device__ inline uint32_t piv_level_elems(uint32_t cap, uint32_t k) { return ((cap >> (3 * k)) + 7u) & ~7u; }
__host__ __device__ inline uint64_t payload_ts_bytes(uint32_t cap) { return align_up((uint64_t)cap * 4, 32); }
__host__ __device__ inline uint64_t payload_pair_bytes(uint32_t cap) { return (uint64_t)cap * 16; }
__host__ __device__ inline uint64_t payload_piv_off(uint32_t cap) {
  return align_up(payload_ts_bytes(cap) + payload_pair_bytes(cap), 32);
}
__host__ __device__ inline uint64_t payload_piv_bytes(uint32_t cap) {
  uint64_t b = 0;
  for (uint32_t k = piv_levels(cap); k >= 1; k--) b += (uint64_t)piv_level_elems(cap, k) * 4;
  return b;
}
__host__ __device__ inline uint64_t payload_bytes(uint32_t cap) {
  return align_up(payload_piv_off(cap) + payload_piv_bytes(cap), kUnit);
}
__host__ __device__ inline uint32_t payload_units(uint32_t cap) { return (uint32_t)(payload_bytes(cap) / kUnit); }
__host__ __device__ inline uint32_t dir_units(uint32_t dir_cap) {
  return (uint32_t)(align_up((uint64_t)dir_cap * sizeof(BlockDesc), kUnit) / kUnit);
}

__device__ __forceinline__ const float *blk_ts(uint64_t payload) { return reinterpret_cast<const float *>(payload); }
// The 16-byte record of an edge behind the block's ts[] array: {neighbour id (vertex ids are < 2^32: ingest rejects
// larger ones), a COPY of the edge's timestamp, edge id}.  Everything an emitted neighbour needs is in one 128-bit load
// of one line -- a random draw on a graph much larger than the L2 pays one DRAM line instead of two (ts[] + {dst, eid}),
// and the recent policy's emit issues one gather less per slot.  ts[] stays: the searches want 32 timestamps per line.
struct __align__(16) EdgeRec {
  uint32_t dst;
  float ts;
  int64_t eid;
};
static_assert(sizeof(EdgeRec) == 16, "EdgeRec is one 128-bit load");
__device__ __forceinline__ const EdgeRec *blk_rec(uint64_t payload, uint32_t cap) {
  return reinterpret_cast<const EdgeRec *>(payload + payload_ts_bytes(cap));
}
__device__ __forceinline__ EdgeRec ld_rec(const EdgeRec *p) {
  const longlong2 v = __ldg(reinterpret_cast<const longlong2 *>(p));
  EdgeRec r;
  r.dst = (uint32_t)((unsigned long long)v.x & 0xffffffffull);
  r.ts = __uint_as_float((uint32_t)((unsigned long long)v.x >> 32));
  r.eid = v.y;
  return r;
}
__host__ inline void unpack_rec_host(const int64_t *pair, int64_t *dst, int64_t *eid) {  // a record copied to the host
  *dst = (int64_t)((uint64_t)pair[0] & 0xffffffffull);
  *eid = pair[1];
}
// pivot level k (1 <= k <= piv_levels(cap)) of a block
__device__ __forceinline__ float *blk_piv(uint64_t payload, uint32_t cap, uint32_t k) {
  uint64_t off = payload_piv_off(cap);
  for (uint32_t m = piv_levels(cap); m > k; m--) off += (uint64_t)piv_level_elems(cap, m) * 4;
  return reinterpret_cast<float *>(payload + off);
}
// the thread that stores ts[pos] also stores the pivots that end at pos
__device__ __forceinline__ void blk_store_pivots(uint64_t payload, uint32_t cap, uint32_t pos, float t) {
  const uint32_t K = piv_levels(cap);
  uint32_t p1 = pos + 1;
  for (uint32_t k = 1; k <= K && (p1 & 7u) == 0; k++) {
    p1 >>= 3;
    blk_piv(payload, cap, k)[p1 - 1] = t;
  }
}

struct F8 {
  float v[8];
};
struct U8x32 {
  uint32_t w[8];
};
// one whole 32-byte record (BlockDesc / NodeEntry) in ONE load instruction: a per-lane scattered access costs one L1
// wavefront per lane and instruction, so two 128-bit loads of the same sector would pay twice
__device__ __forceinline__ U8x32 ldg256_b32(const void *p) {
  U8x32 r;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
               : "l"(p));
  return r;
}
// one 32-byte sector per lane (LDG.E.256 on sm_100a), read-only path
__device__ __forceinline__ F8 ldg256(const float *p) {
  F8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}
// a whole vertex entry: the two sectors of one line, two 256-bit loads issued back to back
__device__ __forceinline__ NodeEntry load_entry64(const NodeEntry *e) {
  const U8x32 a = ldg256_b32(e), b = ldg256_b32(reinterpret_cast<const char *>(e) + 32);
  NodeEntry n;
  n.dir_tagged = ((uint64_t)a.w[1] << 32) | a.w[0];
  n.first = a.w[2];
  n.end = a.w[3];
  n.cum_first = a.w[4];
  n.num_insertions = a.w[5];
  n.num_edges = ((uint64_t)a.w[7] << 32) | a.w[6];
  n.tail.payload = ((uint64_t)b.w[1] << 32) | b.w[0];
  n.tail.size = b.w[2];
  n.tail.capacity = b.w[3];
  n.tail.start_ts = __uint_as_float(b.w[4]);
  n.tail.end_ts = __uint_as_float(b.w[5]);
  n.tail.cum_before = b.w[6];
  n.tail.min_ts = __uint_as_float(b.w[7]);
  return n;
}

__device__ __forceinline__ uint32_t count_lt(const F8 &a, float x, uint32_t nvalid) {
  uint32_t c = 0;
#pragma unroll
  for (uint32_t m = 0; m < 8; m++) c += (m < nvalid && a.v[m] < x) ? 1u : 0u;
  return c;
}
// first index in [0, size) with ts[idx] >= x (size if none), ts non-decreasing: one sector per pivot level
__device__ __forceinline__ uint32_t blk_lower_bound(uint64_t payload, uint32_t cap, uint32_t size, float x) {
  const uint32_t K = piv_levels(cap);
  const float *ts = blk_ts(payload);
  const float *lp = K ? reinterpret_cast<const float *>(payload + payload_piv_off(cap)) : ts;
  const uint32_t nk = size >> (3 * K);  // valid entries of the top level (<= kPivTop)
  uint32_t j = 0;
  if (nk) {
    const F8 a = ldg256(lp);
    F8 b = a;
    if (nk > 8) b = ldg256(lp + 8);
    j = count_lt(a, x, nk) + (nk > 8 ? count_lt(b, x, nk - 8) : 0u);
  }
  for (uint32_t k = K; k-- > 0;) {
    lp = k ? lp + piv_level_elems(cap, k + 1) : ts;
    const uint32_t base = 8 * j, n = size >> (3 * k);
    j = base;
    if (n > base) j += count_lt(ldg256(lp + base), x, n - base);
  }
  return j;
}

// see gf_l2_fetch_granularity (include/gnnflow_b200.h); called once per device by the create functions
int apply_l2_fetch_default(int device);

// ---- stream-ordered scratch buffer that only ever grows ---------------------------------------------------
struct Scratch {
  void *ptr = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes, cudaStream_t st) {
    if (bytes <= cap) return GF_OK;
    size_t want = bytes + bytes / 2 + 4096;
    void *p = nullptr;
    GF_CUDA(cudaMallocAsync(&p, want, st));
    if (ptr) GF_CUDA(cudaFreeAsync(ptr, st));
    ptr = p;
    cap = want;
    return GF_OK;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
  template <class T>
  T *as() const { return reinterpret_cast<T *>(ptr); }
};

inline unsigned cdiv(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

// Wait for a word in mapped pinned memory that a kernel on `st` sets (after a system-wide fence) once what the host is
// about to read is complete: the host sees the store a microsecond or two after it is made, where the wake-up out of
// cudaStreamSynchronize takes 8 - 9 us -- a tenth of a 100 000-edge add_edges.  After 2 ms of polling (large batches, or
// a kernel that never gets there because of an error) the stream is synchronised as before.
inline int wait_flag_or_sync(const volatile unsigned int *flag, cudaStream_t st) {
  static const bool no_spin = getenv("GNNFLOW_B200_NO_SPIN") != nullptr;  // evidence knob
  if (!no_spin) {
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned it = 1;; it++) {
      if (*flag) {
        std::atomic_thread_fence(std::memory_order_acquire);
        return GF_OK;
      }
      if ((it & 255u) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) break;
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
  }
  GF_CUDA(cudaStreamSynchronize(st));
  return GF_OK;
}

// every kernel launch of the library goes through here so that launches can be counted (bench.py gpu_launches)
extern std::atomic<unsigned long long> g_launch_count;
template <class... KArgs, class... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  kernel<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
}

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl may be scheduled while its predecessor in
// the stream is still running; it must call pdl_wait() before touching anything the predecessor wrote (all kernels here
// do so first thing, so only launch latency overlaps), and a predecessor calls pdl_trigger() to allow it.
#ifndef GF_PDL
#define GF_PDL 1
#endif
__device__ __forceinline__ void pdl_wait() {
#if GF_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_trigger() {
#if GF_PDL
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
#if GF_PDL
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
#else
  launch(kernel, grid, block, smem, st, std::forward<Args>(args)...);
#endif
}

// CUDA-event phase timer (off by default): used by bench.py to time individual kernels inside the timed region
struct PhaseProf {
  bool on = false;
  int nphases = 0;
  struct Rec { int phase; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  std::vector<double> ms;
  std::vector<unsigned long long> count;
  cudaEvent_t pending = nullptr;
  void init(int n) { nphases = n; ms.assign(n, 0.0); count.assign(n, 0); }
  cudaEvent_t get() {
    cudaEvent_t e;
    if (!pool.empty()) { e = pool.back(); pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
  }
  void begin(cudaStream_t st) {
    if (!on) return;
    pending = get();
    cudaEventRecord(pending, st);
  }
  // closes the phase opened by begin() (or by the previous end(): phases are back to back)
  void end(int phase, cudaStream_t st, bool chain = true) {
    if (!on || !pending) return;
    cudaEvent_t b = get();
    cudaEventRecord(b, st);
    recs.push_back({phase, pending, b});
    if (chain) { pending = get(); cudaEventRecord(pending, st); } else pending = nullptr;
  }
  void stop() { if (pending) { pool.push_back(pending); pending = nullptr; } }
  void collect() {
    for (auto &r : recs) {
      cudaEventSynchronize(r.b);
      float t = 0;
      if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.phase] += t; count[r.phase]++; }
      pool.push_back(r.a); pool.push_back(r.b);
    }
    recs.clear();
  }
  void reset() { collect(); init(nphases); }
  void destroy() { collect(); stop(); for (auto e : pool) cudaEventDestroy(e); pool.clear(); }
};

}  // namespace gf
