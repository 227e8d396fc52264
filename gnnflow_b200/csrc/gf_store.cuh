// Host-side state of one dynamic graph (opaque gf_graph handle).
#pragma once
#include "gf_common.cuh"

namespace gf {

// per-call error / retry flags raised on the device (any flag set => the batch changes nothing)
enum : uint32_t {
  kErrOutOfOrder = 1u,   // a vertex would receive edges older than its newest stored edge -> GF_EORDER
  kErrBadEid = 2u,       // negative edge id or >= 2^31                                     -> GF_EINVAL
  kErrBadId = 4u,        // negative vertex id or >= 2^32                                   -> GF_EINVAL
  kErrTableSmall = 8u,   // vertex id beyond the table capacity   -> host grows the table and replays the batch
  kErrEidSmall = 16u,    // edge id beyond the refcount capacity  -> host grows it and replays
  kErrUnsorted = 32u,    // batch not in time order on the fast path -> replay with the timestamp sort pass
  kErrArena = 64u,       // free lists + current arena chunk too small -> host reclaims / adds a chunk and replays
  kErrEidLow = 128u      // edge id below the base of the refcount table -> host moves the base down and replays
};

// Per-call scratch; all-zero is the identity, and the prep kernel of call k clears the slot of call k + 1.  A ring:
// the asynchronous ingest keeps up to kCallRing - 2 batches in flight before the host looks at their slots.
constexpr unsigned kCallRing = 16;
struct CallScratch {
  long long max_id, max_eid;
  long long max_neg_eid;  // max(LLONG_MAX - eid) over the batch: LLONG_MAX - this = the smallest edge id
  unsigned int error_flags;
  unsigned int num_segments;
  unsigned int total_units;  // arena units the batch takes from the bump pointer (what is missing, on kErrArena)
  unsigned int accepted;     // set by the plan kernel: the batch passed every check and is being applied
  unsigned int unsorted;     // the batch was not in time order (informational on the slow path)
  unsigned int done_ctas;    // apply kernel: CTAs that have finished (the last one reports to the host)
  unsigned int eids_not_increasing;  // set by the prep pass unless eid[i + 1] > eid[i] throughout (then ids are distinct)
};

// Allocator state on the device (TemporalBlockAllocator, temporal_block_allocator.cu:83-180): a bump pointer into the
// newest chunk plus free lists per size class.  `sorted` holds the free blocks grouped by class (class c owns
// sorted[free_base[c] .. free_base[c] + free_cnt[c]), popped from the top by the plan kernel); blocks freed by offload /
// reallocation / directory growth / chunk tails are appended to `log` and folded into `sorted` by the merge kernels.
constexpr unsigned kMaxRegions = 256;
struct ArenaRegion {
  unsigned long long cur, end;  // bump pointer / end (device addresses)
};
struct ArenaState {
  // bump regions: one per cudaMalloc'd chunk, plus the runs of adjacent free blocks the host's defragmentation pass
  // (arena_defrag) turns back into regions.  A batch's bump allocations go to the current region if they fit, else to
  // the first region that has room (the remainder of the old one stays available to later, smaller batches); only when
  // no region fits does the host step in (kErrArena): fold the free log, defragment, or add a chunk -- and replay.
  ArenaRegion regions[kMaxRegions];
  unsigned int num_regions, cur_region;
  unsigned long long free_units;   // units held by sorted + log
  unsigned int log_cnt;            // entries of the free log
  unsigned int sorted_cnt;         // entries of the sorted array (sum of free_cnt)
  unsigned int free_cnt[kNumClasses];
  unsigned int free_base[kNumClasses];
};
struct FreeRec {  // one entry of the free log
  unsigned long long addr;
  unsigned int cls;
  unsigned int pad;
};

// per-call class table written by the plan kernel for the apply kernel: which requests of class c pop a free block
// (rank < take[c]: sorted[top[c] - 1 - rank]) and where the others start in the bump region
struct CallClasses {
  unsigned int take[kNumClasses];
  unsigned int top[kNumClasses];
  unsigned int bump_base[kNumClasses];  // units from arena_base
  unsigned long long arena_base;
  unsigned long long pad;
};

// Lives in device memory; the persistent counters replace the reference's host-side std::set / unordered_map
// bookkeeping (dynamic_graph.cu:89-97).
struct GraphStats {
  unsigned long long num_edges;        // distinct edge ids currently stored
  unsigned long long num_blocks;       // live blocks (== sum of list lengths)
  unsigned long long allocated_elems;  // sum of live block capacities
  unsigned long long call_count;       // generic counter result (offloaded blocks, flag counts ...)
  unsigned int poison;  // asynchronous ingest: an earlier queued batch was rejected -> later ones must change nothing
  unsigned int pad;
  CallScratch call[kCallRing];
  ArenaState arena;
};

// what the host reads after a call: written into mapped pinned memory by the last CTA of the apply kernel
struct HostResult {
  CallScratch call;
  unsigned long long num_edges, num_blocks, allocated_elems;
  unsigned long long free_units;
  unsigned int log_cnt, sorted_cnt;
};

struct ArenaChunk {
  char *base;
  size_t size;
};

}  // namespace gf

struct gf_graph {
  gf_graph_config cfg;
  std::mutex mu;
  int refs = 1;
  // payload + directory arena
  std::vector<gf::ArenaChunk> chunks;  // every chunk starts as one bump region of ArenaState
  size_t arena_total = 0;
  unsigned num_regions = 0;            // host mirror of ArenaState::num_regions
  uint64_t calls_since_defrag = ~0ull; // add_edges calls since arena_defrag last ran (it is rationed)
  gf::FreeRec *d_log = nullptr;  // free log
  size_t log_cap = 0;
  unsigned long long *d_sorted[2] = {nullptr, nullptr};  // class-sorted free blocks (double-buffered for the merge)
  size_t sorted_cap = 0;
  int sorted_cur = 0;
  uint64_t log_upper = 0;  // upper bound of arena.log_cnt (exact after every host synchronisation)
  uint64_t sorted_upper = 0;
  gf::CallClasses *d_classes = nullptr;  // [kCallRing]
  // vertex table
  gf::NodeEntry *d_table = nullptr;
  uint8_t *d_is_node = nullptr;
  uint8_t *d_is_src = nullptr;
  size_t table_cap = 0;
  int64_t max_node_id = 0;  // reference DynamicGraph::max_node_id_ (0 for an empty graph)
  bool has_nodes = false;
  // edge-id reference counts: dense table over [eid_base, eid_base + eid_cap).  The base follows the live ids up when
  // old blocks are offloaded (a stream whose ids grow with time keeps the table at the size of its live window) and down
  // when a batch brings smaller ids; the span of live ids must stay below 2^31.
  uint32_t *d_eid_ref = nullptr;
  size_t eid_cap = 0;
  long long eid_base = 0;
  gf::GraphStats *d_stats = nullptr;
  gf::GraphStats *h_stats = nullptr;  // pinned mirror (pull_stats)
  gf::HostResult *h_res = nullptr;    // pinned + mapped, [kCallRing]
  // lazily recomputed distinct-vertex counts
  bool counts_dirty = false;
  uint64_t num_nodes = 0, num_src_nodes = 0;
  // offload-to-file ordinal per vertex (temporal_block_allocator.cu:189-191)
  std::vector<uint32_t> saved_blocks_per_node;
  gf::Scratch s_in, s_sort, s_seg, s_misc, s_ctl, s_pre, s_book;
  unsigned call_parity = 0;  // which CallScratch slot (of the ring) the next add_edges attempt uses
  // batches queued by gf_graph_add_edges_async whose outcome the host has not looked at yet
  struct Pending {
    const int64_t *src, *dst;
    const float *ts;
    const int64_t *eid;
    uint64_t n;
    unsigned slot;
  };
  std::vector<Pending> pending;
  cudaStream_t pending_stream = nullptr;
  // a mutation was enqueued on this stream without a host synchronisation after it (gf_graph_clear): the getters, which
  // read on the legacy default stream, wait for exactly this stream instead of the whole device
  cudaStream_t unsettled_stream = nullptr;
  bool unsettled = false;
  bool expect_unsorted = false;  // the previous batch was not in time order: run the timestamp sort pass up front
  gf::PhaseProf prof;

  size_t table_len() const { return has_nodes ? (size_t)max_node_id + 1 : 0; }
};

// settle the batches queued by gf_graph_add_edges_async before another translation unit reads the graph
int gf_graph_flush_internal(gf_graph *g);
