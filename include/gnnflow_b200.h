/*
 * gnnflow_b200.h -- C ABI of the B200-native dynamic-graph store, temporal sampler and feature-cache
 * gather.  This is the drop-in boundary for the path the reference exposes through its pybind11 module
 * `libgnnflow` (reference gnnflow/csrc/api.cc:26-128); each entry point cites what it replaces.
 *
 * Conventions
 *   - every function returns GF_OK (0) or a negative gf_status; the message of the last failure on the
 *     calling thread is available from gf_last_error().  No C++ exception crosses this boundary, nothing
 *     aborts the process (the reference's CHECK_* / LOG(FATAL) do: gnnflow/csrc/logging.cc:53).
 *   - plain pointers and sizes only; `ptr_kind` says whether array arguments live in host or device memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the CUDA legacy default stream).  All device work
 *     is enqueued on it; calls that return host-visible results synchronise that stream only.
 *   - types: node id / edge id = int64_t, timestamp = float (gnnflow/csrc/common.h:13-15).
 *   - the caller allocates every output array; the library only owns opaque handles.
 *   - one handle must not be used from two threads at once (same as the reference,
 *     gnnflow/csrc/temporal_sampler.h:73-77); different handles may.
 */
#ifndef GNNFLOW_B200_H_
#define GNNFLOW_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 3: + gf_graph_add_edges_async / gf_graph_flush, gf_sampler_chain_batched, gf_unique_inverse (additions only)
 * 5: + gf_graph_save / gf_graph_load, gf_graph_memory_breakdown, gf_l2_fetch_granularity (additions only)
 * 6: + gf_cache_fetch, gf_sampler_bind_host_outputs (additions only)
 * 7: + gf_sampler_sample_layer_batched_ids32 (addition only) */
#define GF_ABI_VERSION 7

typedef enum gf_status {
  GF_OK = 0,
  GF_EINVAL = -1,       /* bad argument (shape, enum, negative id, n == 0 ...) */
  GF_EORDER = -2,       /* add_edges: a vertex would receive edges older than its newest stored edge */
  GF_ENOMEM = -3,       /* pool exhausted (maximum_pool_size) or cudaMalloc failure */
  GF_ECUDA = -4,        /* CUDA runtime error */
  GF_ECAPACITY = -5,    /* caller-provided output buffer too small */
  GF_EUNSUPPORTED = -6  /* valid in the reference, not provided by this build */
} gf_status;

/* gnnflow/csrc/common.h:66-90 */
typedef enum { GF_INSERTION_INSERT = 0, GF_INSERTION_REPLACE = 1 } gf_insertion_policy;
typedef enum { GF_SAMPLING_RECENT = 0, GF_SAMPLING_UNIFORM = 1 } gf_sampling_policy;
typedef enum { GF_MEM_CUDA = 0, GF_MEM_UNIFIED = 1, GF_MEM_PINNED = 2, GF_MEM_SHARED = 3 } gf_mem_resource_type;
typedef enum { GF_PTR_HOST = 0, GF_PTR_DEVICE = 1 } gf_ptr_kind;

typedef struct gf_graph gf_graph;
typedef struct gf_sampler gf_sampler;

/* ctor arguments of _DynamicGraph, gnnflow/csrc/api.cc:41-47 / gnnflow/csrc/dynamic_graph.cu:24-51 */
typedef struct gf_graph_config {
  uint64_t initial_pool_size;     /* bytes reserved for edge payload at creation */
  uint64_t maximum_pool_size;     /* hard cap on edge payload bytes; exceeding it -> GF_ENOMEM */
  int32_t mem_resource_type;      /* gf_mem_resource_type; all four place the store in B200 HBM */
  uint64_t minimum_block_size;    /* smallest TemporalBlock capacity (edges) */
  uint64_t blocks_to_preallocate; /* hint: block descriptors to reserve */
  int32_t insertion_policy;       /* gf_insertion_policy */
  int32_t device;                 /* CUDA device ordinal */
  int32_t adaptive_block_size;    /* bool */
} gf_graph_config;

const char *gf_last_error(void);
int gf_abi_version(void);

/* ------------------------------------------------------------------------------------------------
 * dynamic graph  (replaces class _DynamicGraph, gnnflow/csrc/api.cc:41-85)
 * ---------------------------------------------------------------------------------------------- */
int gf_graph_create(const gf_graph_config *cfg, gf_graph **out);
int gf_graph_destroy(gf_graph *g);

/* DynamicGraph::AddEdges, gnnflow/csrc/dynamic_graph.cu:77-138 (api.cc:48-50).  Groups the batch by source
 * vertex, orders each group by timestamp (stable) and appends it to the vertex's block list with the
 * reference's block-sizing policy (dynamic_graph.cu:206-287).  Returns after the batch has been applied (the last CTA
 * of the last kernel has reported; every later call on any stream sees the batch).  On GF_EORDER / GF_EINVAL /
 * GF_ENOMEM the graph is unchanged. */
int gf_graph_add_edges(gf_graph *g, const int64_t *src, const int64_t *dst, const float *ts, const int64_t *eid,
                       uint64_t n, int ptr_kind, void *stream);
/* The same, without the host synchronisation every call of the reference ends with (cudaStreamSynchronize,
 * dynamic_graph.cu:135-137): the batch (DEVICE arrays, which must stay valid until the next flush) is queued on
 * `stream` and the call returns at once; up to 14 batches stay in flight.  Their outcome is looked at by gf_graph_flush,
 * which every other entry point that reads or changes the graph (the samplers included) calls first: a batch that needs a
 * larger table / arena, or that is invalid, changed nothing on the device, and neither did the batches queued after it;
 * the flush replays them in order through gf_graph_add_edges and returns the first error (batches after it are
 * dropped).  The resulting graph is identical to the one gf_graph_add_edges builds. */
int gf_graph_add_edges_async(gf_graph *g, const int64_t *src, const int64_t *dst, const float *ts, const int64_t *eid,
                             uint64_t n, void *stream);
int gf_graph_flush(gf_graph *g);

/* DynamicGraph::OffloadOldBlocks, dynamic_graph.cu:382-411 (api.cc:51-52): drops every block whose
 * end_timestamp < timestamp.  to_file != 0 additionally writes each dropped block as
 * temporal_block_<src>-<k>.bin in the reference's format (temporal_block_allocator.cu:182-222). */
int gf_graph_offload_old_blocks(gf_graph *g, float timestamp, int to_file, uint64_t *num_blocks, void *stream);
/* Reader of that file format (TemporalBlockAllocator::ReadFromFile, temporal_block_allocator.cu:223-256; the reference
 * never calls it).  Always fills the four header fields; dst / ts / eid (host arrays of `cap` entries, all three or
 * none) receive the block's edges, oldest first.  GF_ECAPACITY if cap < size, GF_EINVAL for a malformed file. */
int gf_block_file_read(const char *path, uint64_t *size, uint64_t *capacity, float *start_ts, float *end_ts,
                       int64_t *dst, float *ts, int64_t *eid, uint64_t cap);

/* Whole-graph checkpoint (not in the reference): gf_graph_save writes an image of the store (blocks,
 * directories, vertex table, edge-id reference counts, allocator state) to `path`; gf_graph_load creates a NEW graph on
 * `device` from it that is indistinguishable from the saved one -- getters, sampling results, block shapes, and the
 * result of every later add_edges / offload_old_blocks.  The file is specific to this library version (GF_EINVAL on a
 * foreign or truncated file). */
int gf_graph_save(gf_graph *g, const char *path);
int gf_graph_load(const char *path, int device, gf_graph **out);

/* not in the reference API: empties the graph but keeps every device allocation (vertex table, edge pool) for
 * reuse, so that a replay can start over without paying cudaMalloc again. */
int gf_graph_clear(gf_graph *g, void *stream);

/* scalar getters, api.cc:53-55,67-68,77-85 */
int gf_graph_num_vertices(gf_graph *g, uint64_t *out);         /* distinct src U dst ids seen */
int gf_graph_num_source_vertices(gf_graph *g, uint64_t *out);  /* distinct src ids seen */
int gf_graph_num_edges(gf_graph *g, uint64_t *out);            /* distinct edge ids currently stored */
int gf_graph_max_vertex_id(gf_graph *g, int64_t *out);
int gf_graph_avg_linked_list_length(gf_graph *g, float *out);
int gf_graph_memory_usage(gf_graph *g, float *out);            /* sum of block capacity * 20 B */
int gf_graph_metadata_memory_usage(gf_graph *g, float *out);   /* reference formula: 64 B/block + 8 B/vertex */
int gf_graph_device_bytes(gf_graph *g, uint64_t *out);         /* what this implementation really holds in HBM */
/* ... itemised, out[8]: pool chunks | handed out by the bump pointers | free (lists + log) | vertex table | edge-id
 * reference counts | per-call scratch | allocator book-keeping | number of free blocks on record */
int gf_graph_memory_breakdown(gf_graph *g, uint64_t *out);

/* api.cc:56-59.  ids/out are HOST arrays. */
int gf_graph_out_degree(gf_graph *g, const int64_t *ids, uint64_t n, uint64_t *out);

/* api.cc:60-66: two-call protocol.  *count receives the number of entries; if out != NULL and cap >= *count
 * the entries are written (ascending; HOST array). */
int gf_graph_nodes(gf_graph *g, int64_t *out, uint64_t cap, uint64_t *count);
int gf_graph_src_nodes(gf_graph *g, int64_t *out, uint64_t cap, uint64_t *count);
int gf_graph_edges(gf_graph *g, int64_t *out, uint64_t cap, uint64_t *count);

/* api.cc:69-76 / dynamic_graph.cu:299-337: neighbours of one vertex, newest first.  Two-call, HOST arrays. */
int gf_graph_get_temporal_neighbors(gf_graph *g, int64_t vertex, int64_t *dst, float *ts, int64_t *eid,
                                    uint64_t cap, uint64_t *count);

/* not in the reference API: per-vertex block shapes oldest -> newest (sizes, capacities, start/end
 * timestamps), used by the parity tests to check the block-sizing policy.  Two-call, HOST arrays. */
int gf_graph_block_shapes(gf_graph *g, int64_t vertex, uint64_t *sizes, uint64_t *caps, float *start_ts,
                          float *end_ts, uint64_t cap, uint64_t *count);

/* ------------------------------------------------------------------------------------------------
 * temporal sampler  (replaces class _TemporalSampler + SamplingResult, api.cc:87-120)
 * ---------------------------------------------------------------------------------------------- */
/* TemporalSampler ctor, gnnflow/csrc/temporal_sampler.cu:31-54.  The sampler keeps a reference on `g`. */
int gf_sampler_create(gf_graph *g, const uint32_t *fanouts, uint32_t num_layers, int sampling_policy,
                      uint32_t num_snapshots, float snapshot_time_window, int prop_time, uint64_t seed,
                      gf_sampler **out);
int gf_sampler_destroy(gf_sampler *s);

/* Output of one (layer, snapshot) sampling step == the reference's SamplingResult (common.h:51-60):
 *   all_nodes      [num_dst + num_edges]  roots followed by the sampled neighbours
 *   all_timestamps [num_dst + num_edges]  root timestamps followed by neighbour edge timestamps
 *                                         (root timestamp again if prop_time)
 *   delta_timestamps[num_edges], eids[num_edges]
 *   row[num_edges]  index of the target each edge belongs to (non-decreasing)
 *   col[num_edges]  num_dst + j
 * Arrays must have room for `capacity_dst` roots and capacity_dst * fanout neighbours.  col may be NULL. */
typedef struct gf_sampling_result {
  int64_t *all_nodes;
  float *all_timestamps;
  float *delta_timestamps;
  int64_t *eids;
  int64_t *row;
  int64_t *col;
  uint64_t capacity_dst; /* in: number of roots the arrays were sized for */
  uint64_t num_dst;      /* out */
  uint64_t num_edges;    /* out */
} gf_sampling_result;

/* TemporalSampler::SampleLayer, temporal_sampler.cu:97-277 (api.cc:119-120). */
int gf_sampler_sample_layer(gf_sampler *s, const int64_t *nodes, const float *timestamps, uint64_t num_targets,
                            uint32_t layer, uint32_t snapshot, gf_sampling_result *result, int in_kind,
                            int out_kind, void *stream);

/* TemporalSampler::Sample, temporal_sampler.cu:279-305 (api.cc:116-118): every layer and snapshot in one call.
 * results[layer * num_snapshots + snapshot]; layer l samples the previous layer's all_nodes/all_timestamps,
 * so results[l] needs capacity_dst >= capacity_dst[l-1] * (1 + fanout[l-1]).  Layers are chained on the
 * device; the host waits once at the end to fill num_dst / num_edges.  With HOST arrays on either side the call returns
 * after `stream` has been synchronised (results complete in host memory).  With DEVICE arrays in and out it returns as
 * soon as the sizes of the last step are known -- the kernels may still be writing the result arrays: consumers on
 * `stream` are ordered behind them as usual, work on OTHER streams (readers of the results, calls that mutate the
 * graph) must be ordered by the caller, exactly as for gf_sampler_sample_layer_batched.  GNNFLOW_B200_NO_SPIN=1 in the
 * environment restores the stream synchronisation. */
int gf_sampler_sample(gf_sampler *s, const int64_t *nodes, const float *timestamps, uint64_t num_targets,
                      gf_sampling_result *results, int in_kind, int out_kind, void *stream);

/* Several independent root batches sampled by ONE launch per layer (1 snapshot): batch b covers targets
 * [batch_offsets[b], batch_offsets[b+1]).  Output of batch b is identical to gf_sampler_sample_layer on that
 * batch; the arrays of all batches are concatenated, result->row holds the batch-local target index and
 * edge_offsets[b] (HOST or DEVICE per out_kind, num_batches+1 entries) the first edge of batch b.  This is
 * how a replay of many training batches saturates the GPU (benchmarks/benchmark_sampler.py:70-92).
 * Uniform policy: batch b draws from the counter-based stream with launch index `current + b` (empty batches
 * included), and the call advances the sampler's launch index by num_batches.
 * ptr_kind == GF_PTR_HOST: every array (inputs, outputs, batch_offsets, edge_offsets) is host memory; outputs need
 * room for num_targets * fanout elements.  Inputs are copied to the device; outputs are written by the kernel in
 * place over PCIe when the arrays are pinned (else through a device mirror + copies); the call returns after the
 * host arrays are complete. */
int gf_sampler_sample_layer_batched(gf_sampler *s, const int64_t *nodes, const float *timestamps, uint64_t num_targets,
                                    const uint64_t *batch_offsets, uint64_t num_batches, uint32_t layer,
                                    uint32_t snapshot, int64_t *out_nbr, float *out_ts, float *out_dt,
                                    int64_t *out_eid, int64_t *out_row, uint64_t *edge_offsets,
                                    int ptr_kind, void *stream);

/* The same call for HOST arrays with 32-bit neighbour ids and rows (vertex ids are < 2^32 by the store's contract, rows
 * < num_targets): 24 instead of 32 bytes per sampled neighbour cross PCIe, which is what bounds the host-array call
 * (the narrowing runs on the device, behind the sampling launch).  Values equal the low 32 bits of what
 * gf_sampler_sample_layer_batched returns; out_ts / out_dt / out_eid / edge_offsets are unchanged. */
int gf_sampler_sample_layer_batched_ids32(gf_sampler *s, const int64_t *nodes, const float *timestamps,
                                            uint64_t num_targets, const uint64_t *batch_offsets, uint64_t num_batches,
                                            uint32_t layer, uint32_t snapshot, uint32_t *out_nbr, float *out_ts,
                                            float *out_dt, int64_t *out_eid, uint32_t *out_row, uint64_t *edge_offsets,
                                            void *stream);

/* The targets of the NEXT layer of such a multi-batch launch, built on the device: for every batch b its roots followed
 * by the neighbours just sampled for it -- what TemporalSampler::Sample chains for a single batch
 * (temporal_sampler.cu:242-262, 279-305; `all_nodes = roots || neighbours`, timestamps = the neighbours' edge
 * timestamps, or the root timestamps when the layer ran with prop_time, as returned in nbr_ts).  All arrays DEVICE.
 * nbr / nbr_ts / edge_offsets: outputs of gf_sampler_sample_layer_batched for (nodes, timestamps, batch_offsets);
 * max_edges >= edge_offsets[num_batches] (e.g. num_targets * fanout) only sizes the launch.  nodes_out / timestamps_out
 * need num_targets + edge_offsets[num_batches] entries, batch_offsets_out num_batches + 1
 * (= batch_offsets[b] + edge_offsets[b]).  Asynchronous on `stream`. */
int gf_sampler_chain_batched(const int64_t *nodes, const float *timestamps, uint64_t num_targets,
                             const uint64_t *batch_offsets, uint64_t num_batches, const int64_t *nbr, const float *nbr_ts,
                             const uint64_t *edge_offsets, uint64_t max_edges, int64_t *nodes_out, float *timestamps_out,
                             uint64_t *batch_offsets_out, void *stream);

/* position in the shared counter-based RNG stream (number of non-empty SampleLayer launches so far) */
int gf_sampler_get_launch_index(gf_sampler *s, uint64_t *out);
int gf_sampler_set_launch_index(gf_sampler *s, uint64_t v);
/* which kernels run a sampling step: 3 (default) = the persistent single-launch kernel (fan-outs <= 128); 0 / 1 = the
 * three-launch pipeline (locate, scan, emit) with a warp-cooperative / one-thread-per-target locate, which is also what
 * fan-outs > 128 use.  Other values are rejected. */
int gf_sampler_set_variant(gf_sampler *s, int variant);
/* host output arrays: 0 (default) = auto: pinned arrays are written in place over PCIe by the kernel (per-batch calls,
 * and multi-batch calls with <= 4 MiB of output), larger multi-batch outputs and pageable arrays go through a device
 * mirror + cudaMemcpyAsync; 1 = always mirror; 2 = always in place when pinned (evidence knobs) */
int gf_sampler_set_host_output_mode(gf_sampler *s, int mode);
/* Declare [ptr, ptr + bytes) to be ONE pinned, mapped host allocation that outlives its use as an output area (verified
 * here, once): host result arrays inside it are written in place without a per-array, per-call driver query (6 x
 * cudaPointerGetAttributes per step -- 8-12 us of a 35 us per-batch call).  ptr == NULL unbinds.  The reference's own
 * per-batch call returns freshly allocated vectors (api.cc:116-118); TemporalSampler.sample_numpy keeps one such area. */
int gf_sampler_bind_host_outputs(gf_sampler *s, const void *ptr, uint64_t bytes);

/* ------------------------------------------------------------------------------------------------
 * partitioned sampling over NVLink peer memory (one process per GPU, all GPUs of one box).  Replaces the per-layer
 * RPC fan-out of gnnflow/distributed/dist_sampler.py:159-314 (targets scattered by partition table ->
 * rpc_async(sample_layer_local) -> merge): every rank writes its requests straight into the owner's exchange
 * window and the owner's sampling kernel writes the neighbours straight back into the requester's window -- no
 * collective library call, no host synchronisation between the three phases.
 * ---------------------------------------------------------------------------------------------- */
/* Edge dispatch of the partitioned store (replaces the host-side grouping + RPC of gnnflow/distributed/dispatcher.py:41-100):
 * every rank is handed the same batch (DEVICE arrays) and keeps, in order, the rows whose SOURCE vertex it owns
 * (partition_table / splitmix64 as below).  out_* are DEVICE arrays of n entries; *count (host) receives how many were
 * kept.  One launch, one host synchronisation. */
int gf_dispatch_edges(const int64_t *src, const int64_t *dst, const float *ts, const int64_t *eid, uint64_t n,
                      const int8_t *partition_table, uint64_t table_len, uint32_t rank, uint32_t world,
                      int64_t *out_src, int64_t *out_dst, float *out_ts, int64_t *out_eid, uint64_t *count, void *stream);

typedef struct gf_peer gf_peer;
#define GF_PEER_HANDLE_BYTES 64 /* one cudaIpcMemHandle_t */

/* Allocates this rank's exchange window, sized for `max_targets` targets per step and rank and fan-outs up to
 * `max_fanout`.  Export the handle, all-gather the `world` handles with any transport (torch.distributed), connect. */
int gf_peer_create(int device, uint32_t rank, uint32_t world, uint64_t max_targets, uint32_t max_fanout, gf_peer **out);
int gf_peer_export(gf_peer *p, void *handle_out);
int gf_peer_connect(gf_peer *p, const void *handles /* world * GF_PEER_HANDLE_BYTES, rank-major */);
int gf_peer_destroy(gf_peer *p); /* every rank must have left its last step (barrier first) */

/* One (layer, snapshot) step with vertices partitioned over the ranks; COLLECTIVE: every rank calls it the same
 * number of times, each with its own targets (possibly none).  owner(v) = partition_table[v] (int8, DEVICE, -1 or
 * beyond table_len = unassigned -> no neighbours; dist_sampler.py:174-236), or splitmix64(v) % world when
 * partition_table == NULL.  `s` samples this rank's part of the graph.  nodes / timestamps / result arrays are
 * DEVICE memory; the result is identical to gf_sampler_sample_layer on the unpartitioned graph (both policies:
 * the request carries the target's index, so the counter-based RNG draws the same numbers). */
int gf_sampler_sample_layer_partitioned(gf_sampler *s, gf_peer *p, const int64_t *nodes, const float *timestamps,
                                        uint64_t num_targets, const int8_t *partition_table, uint64_t table_len,
                                        uint32_t layer, uint32_t snapshot, gf_sampling_result *result, void *stream);

/* ------------------------------------------------------------------------------------------------
 * feature cache  (replaces the torch index ops of gnnflow/cache/cache.py:255-413, lru_cache.py:121-201,
 * fifo_cache.py:77-161).  All array arguments are DEVICE pointers unless stated.
 * ---------------------------------------------------------------------------------------------- */
/* out[i, :] = cache_flag[ids[i]] ? cache_buffer[cache_map[ids[i]], :] : features[ids[i], :]
 * (cache.py:275-313 / 336-390).  `features` may be a device table or a pinned, device-mapped host table
 * (zero-copy miss path).  hit_mask (uint8[n], optional) receives the flags; *num_hits (device, optional)
 * is incremented by the number of cached ids. */
int gf_cache_gather(const int64_t *ids, uint64_t n, uint64_t num_items, const uint8_t *cache_flag,
                    const int64_t *cache_map, const float *cache_buffer, const float *features, uint32_t dim,
                    float *out, uint8_t *hit_mask, uint64_t *num_hits, uint32_t *num_bad, void *stream);
/* num_items = rows of `features` (and entries of cache_flag / cache_map).  An id outside [0, num_items) -- where the
 * reference's torch indexing raises IndexError -- is never dereferenced: its row is zero-filled, it counts as a miss,
 * and *num_bad (device counter, optional) is incremented; the policy updates ignore such ids. */

/* plain row gather out[i,:] = features[ids[i],:] (cache.py:411 target_edge_features, utils.py:465-475) */
int gf_gather_rows(const int64_t *ids, uint64_t n, uint64_t num_items, const float *features, uint32_t dim, float *out,
                   uint32_t *num_bad, void *stream);

typedef struct gf_cache_state {
  float *buffer;        /* [capacity, dim] */
  uint8_t *flag;        /* [num_items] */
  int64_t *map;         /* [num_items]  id -> slot or -1 */
  int64_t *index_to_id; /* [capacity]   slot -> id or -1 */
  int32_t *count;       /* [capacity]   LRU water level / LFU use count (NULL for FIFO, static) */
  uint64_t capacity;
  uint64_t num_items;
  uint32_t dim;
} gf_cache_state;

/* Policy updates for one fetch.  ids: the ids of the fetch (device, n entries); hit_mask as produced by
 * gf_cache_gather.  The reference only updates when the fetch had a miss (cache.py:317); then the misses are
 * de-duplicated and sorted ascending (torch.unique) and the first min(#unique, capacity) are admitted over the policy's
 * victims.  scratch: 256-byte aligned device workspace of gf_cache_update_scratch_bytes(n, capacity, num_items) bytes.
 * count_bound (LRU / LFU): every entry of c->count lies in [-count_bound, count_bound] -- e.g. the number of updates
 * made so far + 1 -- which limits the victim sort to the bits the counts occupy; 0 = unknown (full 32-bit sort).
 *
 * LRUCache.update_{node,edge}_cache, lru_cache.py:121-201: every slot's water level -= 1, hit slots -> 0, victims =
 * the smallest water levels, ties broken by lowest slot index (torch.topk leaves ties unspecified); admitted -> 0. */
int gf_cache_update_lru(gf_cache_state *c, const int64_t *ids, const uint8_t *hit_mask, uint64_t n,
                        const float *features, uint64_t count_bound, void *scratch, uint64_t scratch_bytes, void *stream);
/* FIFOCache.update_{node,edge}_cache, fifo_cache.py:77-161; *pointer is the ring pointer, a DEVICE int64 that the
 * kernels read and advance (no host synchronisation). */
int gf_cache_update_fifo(gf_cache_state *c, const int64_t *ids, const uint8_t *hit_mask, uint64_t n,
                         const float *features, int64_t *pointer, void *scratch, uint64_t scratch_bytes,
                         void *stream);
/* LFUCache.update_{node,edge}_cache, lfu_cache.py:133-210: count[hit slots] += 1 once per distinct slot, victims = the
 * smallest counts (ties -> lowest slot), admitted slots start at count 1.  c->count is int32[capacity]. */
int gf_cache_update_lfu(gf_cache_state *c, const int64_t *ids, const uint8_t *hit_mask, uint64_t n,
                        const float *features, uint64_t count_bound, void *scratch, uint64_t scratch_bytes, void *stream);
uint64_t gf_cache_update_scratch_bytes(uint64_t n, uint64_t capacity, uint64_t num_items);
/* One fetch of cache.py:255-413 for one MFG block in ONE call (Cache.fetch_feature's loop body: gather through the cache,
 * hit statistics, policy update): out[i,:] = features[ids[i],:] (served from the cache where ids[i] is cached), then --
 * if `update` and the fetch had a miss -- the policy update of gf_cache_update_<policy> with the hit mask the gather
 * just produced.  The gather kernel itself is the update's collect pass, so a fetch is 3 + P launches (FIFO 3; P = sort
 * passes over the count bits) and one memset instead of gather + 9-11.  policy: GF_CACHE_LRU / FIFO / LFU, or
 * GF_CACHE_STATIC (never updated).  feature_rows: rows of `features` (ids >= min(c->num_items, feature_rows) give a zero
 * row and bump *num_bad, as gf_cache_gather).  hits_out (optional, DEVICE uint64): receives the number of rows served
 * from the cache by this fetch (plain store, no need to clear it).  count_floor (LRU, optional, DEVICE int32): a lower
 * bound of every water level in c->count, read and advanced on the device -- the victim sort then orders only the bits
 * of [floor, 0] (usually one pass; count_bound, the static bound, still sizes the launches).  Whoever changes c->count
 * behind the library's back (reset, resize, gf_cache_update_lru) re-establishes the floor.  scratch as for
 * gf_cache_update_*.  Asynchronous. */
enum { GF_CACHE_LRU = 0, GF_CACHE_FIFO = 1, GF_CACHE_LFU = 2, GF_CACHE_STATIC = 3 };
int gf_cache_fetch(gf_cache_state *c, const int64_t *ids, uint64_t n, const float *features, uint64_t feature_rows, int policy,
                   int64_t *fifo_pointer, uint64_t count_bound, int32_t *count_floor, int update, float *out,
                   uint64_t *hits_out, uint32_t *num_bad, void *scratch, uint64_t scratch_bytes, void *stream);
/* GNNLab static cache (gnnlab_static_cache.py:87-168).  Pre-sampling statistics: counts[id] += 1 once per distinct id
 * of one sampled block (ids outside [0, num_items) are ignored); then the `capacity` ids with the highest counts
 * (ties -> lowest id) are loaded into slots 0..capacity-1 and flag / map rebuilt.  c->index_to_id and c->count may be
 * NULL.  scratch: gf_cache_fill_scratch_bytes(c->num_items) bytes. */
int gf_cache_count_distinct(const int64_t *ids, uint64_t n, int32_t *counts, uint64_t num_items, void *stream);
int gf_cache_fill_topk(gf_cache_state *c, const int32_t *counts, const float *features, void *scratch,
                       uint64_t scratch_bytes, void *stream);
uint64_t gf_cache_fill_scratch_bytes(uint64_t num_items);

/* Sorted de-duplication with inverse map over a bounded id space -- torch.unique(ids, return_inverse=True) as used on
 * the MFG hand-off into the models (Memory.prepare_input, gnnflow/models/modules/memory.py:170-171, which first copies
 * all_nodes to the host) and by the cache's miss path (cache.py:290,355,379).  ids: DEVICE int64[n], every id in
 * [0, num_items).  unique_out (capacity min(n, num_items), optional): the distinct ids ascending; inverse_out (int64[n],
 * optional): inverse_out[i] = position of ids[i] in unique_out (-1 for an id outside the id space); count_out (DEVICE
 * uint64, optional): number of distinct ids.  scratch: 256-byte aligned device workspace of
 * gf_unique_scratch_bytes(num_items) bytes.  Asynchronous on `stream`. */
int gf_unique_inverse(const int64_t *ids, uint64_t n, uint64_t num_items, int64_t *unique_out, int64_t *inverse_out,
                      uint64_t *count_out, void *scratch, uint64_t scratch_bytes, void *stream);
uint64_t gf_unique_scratch_bytes(uint64_t num_items);

/* Feature rows partitioned over the GPUs of one box (replaces KVStoreClient.pull / the RPC feature fetch,
 * gnnflow/distributed/kvstore.py:251-394, graph_services.py:320-357): every rank keeps the rows it owns in a buffer
 * that the other processes map through CUDA IPC, and one gather kernel reads remote rows with plain loads over NVLink.
 *   out[i,:] = shards[owner[ids[i]]][local_index[ids[i]], :]
 * owner (int8, -1 = nobody: the row is zero-filled) and local_index (int32) are DEVICE arrays over the id space;
 * shards is a DEVICE array of `world` pointers (own buffer + gf_shared_open'ed peers). */
int gf_shared_alloc(int device, uint64_t bytes, void **ptr);
int gf_shared_free(void *ptr);
int gf_shared_export(void *ptr, void *handle_out /* GF_PEER_HANDLE_BYTES */);
int gf_shared_open(int device, const void *handle, void **ptr);
int gf_shared_close(void *ptr);
int gf_gather_rows_partitioned(const int64_t *ids, uint64_t n, const int8_t *owner, const int32_t *local_index,
                               uint64_t num_items, const float *const *shards, uint32_t world, uint32_t dim, float *out,
                               void *stream);

/* Zero-copy miss path: make a pageable HOST feature table readable by the gather kernel in place
 * (cudaHostRegister, no copy), the replacement for the reference's index_select -> pinned buffer -> H2D miss path
 * (cache.py:293-313,351-390).  *owned = 1 if this call created the registration (the caller must then call
 * gf_host_unregister before the memory is freed), 0 if the range was already registered / pinned. */
int gf_host_register(void *ptr, uint64_t bytes, int *owned);
int gf_host_unregister(void *ptr);

/* ------------------------------------------------------------------------------------------------
 * measurement hooks (no equivalent in the reference).  Profiling brackets the kernels of each phase with CUDA
 * events on the caller's stream; it is off by default and costs two cudaEventRecord per phase when on.
 * ---------------------------------------------------------------------------------------------- */
#define GF_SAMPLER_PHASES 3 /* 0 locate, 1 scan (count -> offset), 2 emit */
#define GF_GRAPH_PHASES 5   /* 0 stage + prep, 1 sort by source vertex, 2 plan, 3 realloc copy (replace policy), 4 apply */
int gf_sampler_set_profiling(gf_sampler *s, int on);
int gf_sampler_get_profile(gf_sampler *s, double *ms, uint64_t *count, int reset);
int gf_graph_set_profiling(gf_graph *g, int on);
int gf_graph_get_profile(gf_graph *g, double *ms, uint64_t *count, int reset);
/* kernels launched by this library in this process so far */
uint64_t gf_debug_launch_count(void);

/* L2 fetch granularity of `device` (cudaLimitMaxL2FetchGranularity): how many bytes the L2 brings in from HBM for one
 * missing 32-byte sector.  Measured on B200: every random sector the sampler touches on a graph that does not fit the L2
 * costs ~3.4 sectors of DRAM traffic, and the limit -- a hint -- changes nothing (profiles/r02_l2_fetch_granularity_ab.json),
 * so the library leaves the device alone unless the environment says GNNFLOW_B200_L2_FETCH = 32 | 64 | 128; the data
 * layout keeps what one sample needs in as few 128-byte lines as it can instead.  bytes == 0 only reads the value back. */
int gf_l2_fetch_granularity(int device, uint64_t bytes, uint64_t *current);

#ifdef __cplusplus
}
#endif
#endif /* GNNFLOW_B200_H_ */
