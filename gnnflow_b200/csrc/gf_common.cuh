// Shared definitions of the B200-native store / sampler / cache library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
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
// sector -- plus one contiguous 128-byte-aligned payload:  ts[cap] | dst[cap] | eid[cap]  (each sub-array
// padded to 16 B so that 128-bit loads are always legal).  The descriptors of one vertex sit in a
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
  uint32_t reserved;
};
static_assert(sizeof(BlockDesc) == 32, "BlockDesc must be one 32-byte sector");

// One vertex (reference DoublyLinkedList + HostDoublyLinkedList, doubly_linked_list.h:15-34): 32 bytes.
struct __align__(32) NodeEntry {
  uint64_t dir;             // device address of the BlockDesc directory (0 = never had a block)
  uint32_t first;           // oldest live block (blocks before it were offloaded)
  uint32_t end;             // one past the newest block; live blocks are [first, end)
  uint32_t dir_cap;         // descriptors that fit in `dir`
  uint32_t num_insertions;  // HostDoublyLinkedList::num_insertions
  uint64_t num_edges;       // HostDoublyLinkedList::num_edges == out_degree (never decremented)
};
static_assert(sizeof(NodeEntry) == 32, "NodeEntry must be one 32-byte sector");

constexpr uint32_t kUnit = 128;  // allocation granule of the payload arena, bytes

__host__ __device__ inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline uint64_t payload_ts_bytes(uint32_t cap) { return align_up((uint64_t)cap * 4, 16); }
__host__ __device__ inline uint64_t payload_i64_bytes(uint32_t cap) { return align_up((uint64_t)cap * 8, 16); }
__host__ __device__ inline uint64_t payload_bytes(uint32_t cap) {
  return align_up(payload_ts_bytes(cap) + 2 * payload_i64_bytes(cap), kUnit);
}
__host__ __device__ inline uint32_t payload_units(uint32_t cap) { return (uint32_t)(payload_bytes(cap) / kUnit); }
__host__ __device__ inline uint32_t dir_units(uint32_t dir_cap) {
  return (uint32_t)(align_up((uint64_t)dir_cap * sizeof(BlockDesc), kUnit) / kUnit);
}

__device__ __forceinline__ const float *blk_ts(uint64_t payload) { return reinterpret_cast<const float *>(payload); }
__device__ __forceinline__ const int64_t *blk_dst(uint64_t payload, uint32_t cap) {
  return reinterpret_cast<const int64_t *>(payload + payload_ts_bytes(cap));
}
__device__ __forceinline__ const int64_t *blk_eid(uint64_t payload, uint32_t cap) {
  return reinterpret_cast<const int64_t *>(payload + payload_ts_bytes(cap) + payload_i64_bytes(cap));
}

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

// every kernel launch of the library goes through here so that launches can be counted (bench.py gpu_launches)
extern std::atomic<unsigned long long> g_launch_count;
template <class... KArgs, class... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  kernel<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
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
