// Host-side state of one dynamic graph (opaque gf_graph handle).
#pragma once
#include "gf_common.cuh"

namespace gf {

// per-call error / retry flags raised on the device (any flag set => the commit and scatter kernels change nothing)
enum : uint32_t {
  kErrOutOfOrder = 1u,   // a vertex would receive edges older than its newest stored edge -> GF_EORDER
  kErrBadEid = 2u,       // negative edge id or >= 2^31                                     -> GF_EINVAL
  kErrBadId = 4u,        // negative vertex id or >= 2^32                                   -> GF_EINVAL
  kErrTableSmall = 8u,   // vertex id beyond the table capacity   -> host grows the table and replays the batch
  kErrEidSmall = 16u,    // edge id beyond the refcount capacity  -> host grows it and replays
  kErrUnsorted = 32u,    // batch not in time order on the fast path -> replay with the timestamp sort pass
  kErrArena = 64u        // current arena chunk too small          -> host adds a chunk and replays
};

// Per-call scratch; all-zero is the identity, and the prep kernel of call k clears the slot of call k + 1.  A ring:
// the asynchronous ingest keeps up to kCallRing - 1 batches in flight before the host looks at their slots.
constexpr unsigned kCallRing = 16;
struct CallScratch {
  long long max_id, max_eid;
  unsigned int error_flags;
  unsigned int num_segments;
  unsigned int total_units;
  unsigned int accepted;  // set by the commit kernel: the batch passed every check and is being applied
  unsigned int unsorted;  // the batch was not in time order (informational on the slow path)
  unsigned int pad;
};

// Lives in device memory with a pinned host mirror; the persistent counters replace the reference's
// host-side std::set / unordered_map bookkeeping (dynamic_graph.cu:89-97).
struct GraphStats {
  unsigned long long num_edges;        // distinct edge ids currently stored
  unsigned long long num_blocks;       // live blocks (== sum of list lengths)
  unsigned long long allocated_elems;  // sum of live block capacities
  unsigned long long dead_units;       // arena units no longer referenced (offloaded / reallocated)
  unsigned long long arena_cur;        // bump pointer (device address) into the current arena chunk
  unsigned long long arena_end;        // end of the current arena chunk
  CallScratch call[kCallRing];
  unsigned long long call_count;  // generic counter result (offloaded blocks, flag counts ...)
  unsigned int poison;  // asynchronous ingest: an earlier queued batch was rejected -> later ones must change nothing
  unsigned int pad;
};

struct ArenaChunk {
  char *base;
  size_t size;
  size_t used;
};

}  // namespace gf

struct gf_graph {
  gf_graph_config cfg;
  std::mutex mu;
  int refs = 1;
  // payload + directory arena
  std::vector<gf::ArenaChunk> chunks;
  size_t arena_total = 0;
  // vertex table
  gf::NodeEntry *d_table = nullptr;
  uint8_t *d_is_node = nullptr;
  uint8_t *d_is_src = nullptr;
  size_t table_cap = 0;
  int64_t max_node_id = 0;  // reference DynamicGraph::max_node_id_ (0 for an empty graph)
  bool has_nodes = false;
  // edge-id reference counts (dense)
  uint32_t *d_eid_ref = nullptr;
  size_t eid_cap = 0;
  gf::GraphStats *d_stats = nullptr;
  gf::GraphStats *h_stats = nullptr;  // pinned
  // lazily recomputed distinct-vertex counts
  bool counts_dirty = false;
  uint64_t num_nodes = 0, num_src_nodes = 0;
  // offload-to-file ordinal per vertex (temporal_block_allocator.cu:189-191)
  std::vector<uint32_t> saved_blocks_per_node;
  gf::Scratch s_in, s_sort, s_seg, s_misc, s_lb;  // s_lb: ticket + tile status words of the look-back scans
  size_t lb_tiles = 0;
  unsigned long long lb_gen = 0;
  unsigned call_parity = 0;      // which CallScratch slot (of the ring) the next add_edges attempt uses
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
