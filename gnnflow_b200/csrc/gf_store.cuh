// Host-side state of one dynamic graph (opaque gf_graph handle).
#pragma once
#include "gf_common.cuh"

namespace gf {

enum : uint32_t { kErrOutOfOrder = 1u, kErrBadEid = 2u };

// Lives in device memory with a pinned host mirror; the persistent counters replace the reference's
// host-side std::set / unordered_map bookkeeping (dynamic_graph.cu:89-97).
struct GraphStats {
  unsigned long long num_edges;        // distinct edge ids currently stored
  unsigned long long num_blocks;       // live blocks (== sum of list lengths)
  unsigned long long allocated_elems;  // sum of live block capacities
  unsigned long long dead_units;       // arena units no longer referenced (offloaded / reallocated)
  // per-call scratch
  long long batch_min_id, batch_max_id, batch_min_eid, batch_max_eid;
  unsigned int ts_unsorted;
  unsigned int num_segments;
  unsigned int error_flags;
  unsigned int total_units;
  unsigned long long call_count;  // generic counter result (offloaded blocks, flag counts ...)
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
  gf::Scratch s_in, s_sort, s_seg, s_misc;
  gf::PhaseProf prof;

  size_t table_len() const { return has_nodes ? (size_t)max_node_id + 1 : 0; }
};
