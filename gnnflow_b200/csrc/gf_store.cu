// Dynamic-graph store: per-vertex time-sorted TemporalBlocks and the batched edge-insert path, entirely on
// the device.  Replaces reference gnnflow/csrc/dynamic_graph.cu, temporal_block_allocator.cu,
// doubly_linked_list.cu and the host loops of DynamicGraph::AddEdges (dynamic_graph.cu:77-138,206-287).
//
// add_edges = ONE host synchronisation and 3 + P kernel launches per batch (P = 8-bit radix passes over the source-vertex
// bits) regardless of how many vertices the batch touches (the reference issues ~5 CUDA API calls per distinct source
// vertex): prep -> P payload-carrying sort passes -> plan (segments, block-sizing policy, allocation by size class,
// accept / reject) -> apply (payload, descriptors, directories, bookkeeping, report).  Kernels: gf_ingest.cuh.
#include <cfloat>
#include <cstdio>
#include <cstdlib>

#include <algorithm>

#include "gf_ingest.cuh"

namespace gf {

// ------------------------------------------------------------------------------------------ error text
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char *get_error() { return g_err; }
std::atomic<unsigned long long> g_launch_count{0};

// ------------------------------------------------------------------------------------------ L2 fetch granularity
int apply_l2_fetch_default(int device) {
  static std::mutex mu;
  static bool done[64] = {false};
  std::lock_guard<std::mutex> lk(mu);
  if (device < 0 || device >= 64 || done[device]) return GF_OK;
  done[device] = true;
  const char *e = getenv("GNNFLOW_B200_L2_FETCH");
  const long v = e ? atol(e) : 0;
  if (v != 32 && v != 64 && v != 128) return GF_OK;  // 0 / anything else: leave the device as it is
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != device) cudaSetDevice(device);
  cudaError_t err = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v);
  if (err != cudaSuccess) cudaGetLastError();  // a hint: devices may clamp or ignore it
  if (cur != device) cudaSetDevice(cur);
  return GF_OK;
}

// ------------------------------------------------------------------------------------------ host helpers
static int set_device(const gf_graph *g) {
  GF_CUDA(cudaSetDevice(g->cfg.device));
  return GF_OK;
}

// host reads (getters) must see every mutation enqueued so far: all entry points but gf_graph_clear end with a stream
// synchronisation of their own, so only that one stream can still be busy
static int settle(gf_graph *g) {
  if (g->unsettled) {
    GF_CUDA(cudaStreamSynchronize(g->unsettled_stream));
    g->unsettled = false;
  }
  return GF_OK;
}

static int ensure_table(gf_graph *g, int64_t max_id, cudaStream_t st) {
  size_t need = (size_t)max_id + 1;
  if (need > g->table_cap) {
    size_t cap = g->table_cap ? g->table_cap * 2 : 1024;
    if (cap < need) cap = need;
    NodeEntry *nt;
    uint8_t *nn, *ns;
    GF_CUDA(cudaMallocAsync(&nt, cap * sizeof(NodeEntry), st));
    GF_CUDA(cudaMallocAsync(&nn, cap, st));
    GF_CUDA(cudaMallocAsync(&ns, cap, st));
    size_t old = g->table_cap;
    if (old) {
      GF_CUDA(cudaMemcpyAsync(nt, g->d_table, old * sizeof(NodeEntry), cudaMemcpyDeviceToDevice, st));
      GF_CUDA(cudaMemcpyAsync(nn, g->d_is_node, old, cudaMemcpyDeviceToDevice, st));
      GF_CUDA(cudaMemcpyAsync(ns, g->d_is_src, old, cudaMemcpyDeviceToDevice, st));
      GF_CUDA(cudaFreeAsync(g->d_table, st));
      GF_CUDA(cudaFreeAsync(g->d_is_node, st));
      GF_CUDA(cudaFreeAsync(g->d_is_src, st));
    }
    GF_CUDA(cudaMemsetAsync(nt + old, 0, (cap - old) * sizeof(NodeEntry), st));
    GF_CUDA(cudaMemsetAsync(nn + old, 0, cap - old, st));
    GF_CUDA(cudaMemsetAsync(ns + old, 0, cap - old, st));
    g->d_table = nt;
    g->d_is_node = nn;
    g->d_is_src = ns;
    g->table_cap = cap;
  }
  // DynamicGraph::AddNodes, dynamic_graph.cu:140-147
  if (!g->has_nodes || max_id > g->max_node_id) g->max_node_id = max_id;
  g->has_nodes = true;
  return GF_OK;
}

constexpr long long kEidAlign = 4096;
constexpr uint64_t kEidSpanMax = 1ull << 31;

// make [min_eid, max_eid] fit the reference-count table: move the base down and / or grow the capacity
static int ensure_eids(gf_graph *g, int64_t min_eid, int64_t max_eid, cudaStream_t st) {
  long long base = g->eid_cap ? g->eid_base : (min_eid / kEidAlign * kEidAlign);
  if (min_eid < base) base = min_eid / kEidAlign * kEidAlign;
  const uint64_t shift = g->eid_cap ? (uint64_t)(g->eid_base - base) : 0;  // entries to add in front
  const uint64_t need = std::max<uint64_t>((uint64_t)(max_eid - base) + 1, g->eid_cap + shift);
  if (need > kEidSpanMax)
    GF_FAIL(GF_EINVAL, "add_edges: the edge ids in the graph would span %llu > 2^31 (ids %lld .. %lld; the reference counts "
            "live in a dense table over the live id range)", (unsigned long long)need, base, (long long)max_eid);
  if (shift == 0 && need <= g->eid_cap) return GF_OK;
  size_t cap = g->eid_cap ? g->eid_cap * 2 : 4096;
  if (cap < need) cap = need;
  if (cap > kEidSpanMax) cap = kEidSpanMax;
  uint32_t *nr;
  GF_CUDA(cudaMallocAsync(&nr, cap * sizeof(uint32_t), st));
  GF_CUDA(cudaMemsetAsync(nr, 0, cap * sizeof(uint32_t), st));
  if (g->eid_cap) {
    GF_CUDA(cudaMemcpyAsync(nr + shift, g->d_eid_ref, g->eid_cap * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    GF_CUDA(cudaFreeAsync(g->d_eid_ref, st));
  }
  g->d_eid_ref = nr;
  g->eid_cap = cap;
  g->eid_base = base;
  return GF_OK;
}

// after blocks were offloaded: if the front quarter of the table holds no live id any more, move the base up
static int compact_eids(gf_graph *g, cudaStream_t st) {
  if (g->eid_cap < (1u << 16)) return GF_OK;
  unsigned long long *d = &g->d_stats->call_count, h = g->eid_cap;
  GF_CUDA(cudaMemcpyAsync(d, &h, sizeof(h), cudaMemcpyHostToDevice, st));
  gf::launch(first_nonzero_kernel, std::min(cdiv(g->eid_cap, kThreads * 8), 148u * 4), kThreads, 0, st, g->d_eid_ref,
             (uint64_t)g->eid_cap, d);
  GF_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, st));
  GF_CUDA(cudaStreamSynchronize(st));
  const uint64_t shift = h / kEidAlign * kEidAlign;
  if (h >= g->eid_cap || shift < g->eid_cap / 4) return GF_OK;  // nothing live (keep the base) or not worth a move
  uint32_t *nr;
  GF_CUDA(cudaMallocAsync(&nr, g->eid_cap * sizeof(uint32_t), st));
  GF_CUDA(cudaMemcpyAsync(nr, g->d_eid_ref + shift, (g->eid_cap - shift) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  GF_CUDA(cudaMemsetAsync(nr + (g->eid_cap - shift), 0, shift * sizeof(uint32_t), st));
  GF_CUDA(cudaFreeAsync(g->d_eid_ref, st));
  g->d_eid_ref = nr;
  g->eid_base += (long long)shift;
  return GF_OK;
}

static int pull_stats(gf_graph *g, cudaStream_t st) {
  GF_CUDA(cudaMemcpyAsync(g->h_stats, g->d_stats, sizeof(GraphStats), cudaMemcpyDeviceToHost, st));
  GF_CUDA(cudaStreamSynchronize(st));
  g->log_upper = g->h_stats->arena.log_cnt;
  g->sorted_upper = g->h_stats->arena.sorted_cnt;
  return GF_OK;
}

// ------------------------------------------------------------------------------------------ allocator (host side)
// The arena is a list of cudaMalloc'd chunks (initial_pool_size, then growing by an eighth of what is held, within
// maximum_pool_size: the reference's rmm pool_memory_resource(initial, maximum), temporal_block_allocator.cu:27-65).
// Allocation happens on the device (ingest_plan_kernel): free lists per size class first, then a bump pointer into the
// newest chunk.  The host only steps in when the device reports kErrArena: it folds the free log into the lists, or
// adds a chunk -- the unused tail of the previous chunk goes to the free lists, nothing is stranded.

// room for `pushes` more entries in the free log
static int ensure_log(gf_graph *g, uint64_t pushes, cudaStream_t st) {
  const uint64_t need = g->log_upper + pushes;
  if (need <= g->log_cap) return GF_OK;
  const size_t cap = std::max<size_t>(need + need / 2, 4096);
  FreeRec *nl;
  GF_CUDA(cudaMallocAsync(&nl, cap * sizeof(FreeRec), st));
  if (g->d_log) {
    if (g->log_upper) GF_CUDA(cudaMemcpyAsync(nl, g->d_log, g->log_upper * sizeof(FreeRec), cudaMemcpyDeviceToDevice, st));
    GF_CUDA(cudaFreeAsync(g->d_log, st));
  }
  g->d_log = nl;
  g->log_cap = cap;
  return GF_OK;
}

// fold the free log into the class-sorted free lists (4 small launches; off the per-batch path)
static int arena_merge(gf_graph *g, cudaStream_t st) {
  if (g->log_upper == 0) return GF_OK;
  const uint64_t need = g->sorted_upper + g->log_upper;
  if (need > g->sorted_cap) {
    const size_t cap = std::max<size_t>(need + need / 2, 4096);
    unsigned long long *a, *b;
    GF_CUDA(cudaMallocAsync(&a, cap * 8, st));
    GF_CUDA(cudaMallocAsync(&b, cap * 8, st));
    if (g->d_sorted[0]) {
      // free_base[] indexes the current buffer: keep its whole extent
      if (g->sorted_cap) GF_CUDA(cudaMemcpyAsync(a, g->d_sorted[g->sorted_cur], g->sorted_cap * 8, cudaMemcpyDeviceToDevice, st));
      GF_CUDA(cudaFreeAsync(g->d_sorted[0], st));
      GF_CUDA(cudaFreeAsync(g->d_sorted[1], st));
    }
    g->d_sorted[0] = a;
    g->d_sorted[1] = b;
    g->sorted_cur = 0;
    g->sorted_cap = cap;
  }
  GF_TRY(g->s_misc.reserve(3 * kNumClasses * 4 + 256, st));
  GF_CUDA(cudaMemsetAsync(g->s_misc.ptr, 0, 3 * kNumClasses * 4, st));
  MergeArgs m = {&g->d_stats->arena, g->d_log, g->d_sorted[g->sorted_cur], g->d_sorted[g->sorted_cur ^ 1], g->s_misc.as<unsigned int>()};
  const unsigned nb = (unsigned)std::min<uint64_t>(cdiv(g->log_upper, kThreads), 148ull * 4);
  gf::launch(merge_count_kernel, nb, kThreads, 0, st, m);
  gf::launch(merge_move_kernel, kNumClasses, kThreads, 0, st, m);
  gf::launch(merge_scatter_kernel, nb, kThreads, 0, st, m);
  gf::launch(merge_finish_kernel, 1, kThreads, 0, st, m);
  GF_CUDA(cudaGetLastError());
  g->sorted_cur ^= 1;
  g->sorted_upper = need;
  g->log_upper = 0;
  return GF_OK;
}

// Defragmentation (host, off the per-batch path; rationed by the caller): size classes never coalesce, so a workload
// whose requests keep growing -- the replace policy's hot vertices -- leaves free blocks nobody asks for again.  Gather
// ALL free space (class lists, log, the remainders of the bump regions), sort it by address, merge neighbours into runs;
// the largest runs become the bump regions, the rest goes back to the class lists in class-sized pieces.  Nothing that is
// live moves.
constexpr unsigned kDefragRegions = 64;
static int arena_defrag(gf_graph *g, cudaStream_t st) {
  GF_TRY(arena_merge(g, st));
  GF_TRY(pull_stats(g, st));
  ArenaState ar = g->h_stats->arena;
  std::vector<unsigned long long> h_sorted(g->sorted_cap);
  if (g->sorted_cap) GF_CUDA(cudaMemcpy(h_sorted.data(), g->d_sorted[g->sorted_cur], g->sorted_cap * 8, cudaMemcpyDeviceToHost));
  struct Run { unsigned long long addr, units; };
  std::vector<Run> runs;
  runs.reserve(ar.sorted_cnt + ar.num_regions);
  for (unsigned c = 0; c < kNumClasses; c++)
    for (unsigned j = 0; j < ar.free_cnt[c]; j++) runs.push_back({h_sorted[ar.free_base[c] + j], class_units(c)});
  for (unsigned k = 0; k < ar.num_regions; k++)
    if (ar.regions[k].end > ar.regions[k].cur) runs.push_back({ar.regions[k].cur, (ar.regions[k].end - ar.regions[k].cur) / kUnit});
  std::sort(runs.begin(), runs.end(), [](const Run &a, const Run &b) { return a.addr < b.addr; });
  std::vector<unsigned long long> bounds;  // chunk starts: runs never grow across two cudaMalloc'd chunks
  for (auto &c : g->chunks) bounds.push_back((unsigned long long)(uintptr_t)c.base);
  std::sort(bounds.begin(), bounds.end());
  std::vector<Run> merged;
  for (const Run &r : runs) {
    if (!merged.empty() && merged.back().addr + merged.back().units * kUnit == r.addr &&
        !std::binary_search(bounds.begin(), bounds.end(), r.addr))
      merged.back().units += r.units;
    else
      merged.push_back(r);
  }
  // the kDefragRegions largest runs (of at least 1 MB) become bump regions
  std::vector<size_t> order(merged.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = i;
  const size_t nreg = std::min<size_t>(kDefragRegions, order.size());
  std::partial_sort(order.begin(), order.begin() + nreg, order.end(),
                    [&](size_t a, size_t b) { return merged[a].units > merged[b].units; });
  std::vector<bool> is_region(merged.size(), false);
  memset(ar.regions, 0, sizeof(ar.regions));
  ar.num_regions = 0;
  for (size_t k = 0; k < nreg; k++) {
    const Run &r = merged[order[k]];
    if (k > 0 && r.units * kUnit < (1u << 20)) break;
    is_region[order[k]] = true;
    ar.regions[ar.num_regions++] = {r.addr, r.addr + r.units * kUnit};
  }
  ar.cur_region = 0;  // the largest
  // everything else: class-sized pieces, largest first
  std::vector<std::vector<unsigned long long>> lists(kNumClasses);
  unsigned long long free_units = 0, pieces = 0;
  for (size_t i = 0; i < merged.size(); i++) {
    if (is_region[i]) continue;
    unsigned long long addr = merged[i].addr, u = merged[i].units;
    while (u) {
      uint32_t c = class_of_units((uint32_t)std::min<unsigned long long>(u, 0x40000000ull));
      if (class_units(c) > u) c--;  // the largest class that fits (classes are exact up to 16 units, so c >= 0 here)
      lists[c].push_back(addr);
      addr += (unsigned long long)class_units(c) * kUnit;
      u -= class_units(c);
      free_units += class_units(c);
      pieces++;
    }
  }
  if (pieces > g->sorted_cap) {
    const size_t cap = pieces + pieces / 2 + 4096;
    unsigned long long *a, *b;
    GF_CUDA(cudaMallocAsync(&a, cap * 8, st));
    GF_CUDA(cudaMallocAsync(&b, cap * 8, st));
    if (g->d_sorted[0]) {
      GF_CUDA(cudaFreeAsync(g->d_sorted[0], st));
      GF_CUDA(cudaFreeAsync(g->d_sorted[1], st));
    }
    GF_CUDA(cudaStreamSynchronize(st));
    g->d_sorted[0] = a;
    g->d_sorted[1] = b;
    g->sorted_cur = 0;
    g->sorted_cap = cap;
  }
  h_sorted.assign(pieces, 0);
  unsigned pos = 0;
  for (unsigned c = 0; c < kNumClasses; c++) {
    ar.free_base[c] = pos;
    ar.free_cnt[c] = (unsigned)lists[c].size();
    for (unsigned long long a : lists[c]) h_sorted[pos++] = a;
  }
  ar.free_units = free_units;
  ar.sorted_cnt = (unsigned)pieces;
  ar.log_cnt = 0;
  if (pieces) GF_CUDA(cudaMemcpy(g->d_sorted[g->sorted_cur], h_sorted.data(), pieces * 8, cudaMemcpyHostToDevice));
  GF_CUDA(cudaMemcpy(&g->d_stats->arena, &ar, sizeof(ar), cudaMemcpyHostToDevice));
  g->h_stats->arena = ar;
  g->num_regions = ar.num_regions;
  g->log_upper = 0;
  g->sorted_upper = pieces;
  g->calls_since_defrag = 0;
  return GF_OK;
}

// No bump region has room for `bytes`: add a chunk = a new region (the device moves into it on the replay).
static int arena_add_chunk(gf_graph *g, size_t bytes, cudaStream_t st) {
  if (g->num_regions >= kMaxRegions) GF_TRY(arena_defrag(g, st));  // folds exhausted and small regions away
  if (g->num_regions >= kMaxRegions || g->chunks.size() >= kMaxRegions)
    GF_FAIL(GF_ENOMEM, "edge pool: more than %u regions", kMaxRegions);
  size_t maxp = g->cfg.maximum_pool_size ? g->cfg.maximum_pool_size : SIZE_MAX;
  size_t want = g->chunks.empty() ? (size_t)g->cfg.initial_pool_size
                                  : std::min<size_t>(std::max<size_t>(g->arena_total / 8, 8u << 20), 4ull << 30);
  if (want < bytes) want = bytes;
  if (want < (1u << 20)) want = 1u << 20;
  if (g->arena_total + want > maxp) want = maxp > g->arena_total ? maxp - g->arena_total : 0;
  want = want / kUnit * kUnit;
  if (want < bytes || want == 0)
    GF_FAIL(GF_ENOMEM, "edge pool exhausted: need %zu more bytes, pool holds %zu of maximum_pool_size %zu", bytes,
            g->arena_total, (size_t)g->cfg.maximum_pool_size);
  char *p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    GF_FAIL(GF_ENOMEM, "cudaMalloc(%zu) for the edge pool failed: %s", want, cudaGetErrorString(e));
  }
  const unsigned int k = g->num_regions++;
  g->chunks.push_back({p, want});
  g->arena_total += want;
  ArenaRegion reg = {(unsigned long long)(uintptr_t)p, (unsigned long long)(uintptr_t)(p + want)};
  unsigned int hdr[2] = {k + 1, k};  // num_regions, cur_region
  GF_CUDA(cudaMemcpyAsync(&g->d_stats->arena.regions[k], &reg, sizeof(reg), cudaMemcpyHostToDevice, st));
  GF_CUDA(cudaMemcpyAsync(&g->d_stats->arena.num_regions, hdr, sizeof(hdr), cudaMemcpyHostToDevice, st));
  GF_CUDA(cudaStreamSynchronize(st));  // stack sources
  return GF_OK;
}

static int bit_width_u64(uint64_t x) {
  int b = 0;
  while (x) {
    b++;
    x >>= 1;
  }
  return b;
}

static int add_edges_impl(gf_graph *g, const int64_t *src, const int64_t *dst, const float *ts, const int64_t *eid,
                          uint64_t n, int ptr_kind, cudaStream_t st, bool async);

// Look at the outcome of the batches queued by gf_graph_add_edges_async (one synchronisation for all of them).  The
// first rejected batch -- table / edge-id / arena capacity, a batch out of time order, a bad id -- and every batch
// queued after it changed nothing on the device (GraphStats::poison); they are replayed here, in order, through the
// synchronous path, which fixes the cause or reports the error.
static int flush_pending(gf_graph *g) {
  if (g->pending.empty()) return GF_OK;
  cudaStream_t st = g->pending_stream;
  GF_TRY(set_device(g));
  GF_CUDA(cudaStreamSynchronize(st));
  std::vector<gf_graph::Pending> q;
  q.swap(g->pending);
  size_t j = 0;
  for (; j < q.size(); j++) {
    const HostResult &hr = g->h_res[q[j].slot];
    if (!hr.call.accepted) break;
    if (!g->has_nodes || hr.call.max_id > g->max_node_id) g->max_node_id = hr.call.max_id;
    g->has_nodes = true;
    g->counts_dirty = true;
    g->h_stats->num_edges = hr.num_edges;
    g->h_stats->num_blocks = hr.num_blocks;
    g->h_stats->allocated_elems = hr.allocated_elems;
    g->log_upper = hr.log_cnt;
    g->sorted_upper = hr.sorted_cnt;
  }
  if (j == q.size()) return GF_OK;
  GF_TRY(pull_stats(g, st));  // exact allocator state after the accepted prefix
  GF_CUDA(cudaMemsetAsync(&g->d_stats->poison, 0, sizeof(unsigned int), st));
  for (; j < q.size(); j++) {
    const int rc = add_edges_impl(g, q[j].src, q[j].dst, q[j].ts, q[j].eid, q[j].n, GF_PTR_DEVICE, st, false);
    if (rc != GF_OK) return rc;  // the batches queued after the failing one are dropped
  }
  return GF_OK;
}

template <int ROUNDS, bool FIRST>
static int launch_sort_pass(const SortSrc &in, const SortDst &out, uint64_t n, int shift, const uint32_t *ghist,
                            uint32_t *ticket, uint32_t *status, uint32_t tiles, cudaStream_t st) {
  constexpr size_t dyn = (size_t)kSortThreads * ROUNDS * 24;
  static bool configured = false;  // per instantiation; the attribute is per function and device-independent here
  if (!configured && dyn > 32 * 1024) {
    GF_CUDA(cudaFuncSetAttribute(ingest_sort_kernel<ROUNDS, FIRST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    configured = true;
  }
  gf::launch_pdl(ingest_sort_kernel<ROUNDS, FIRST>, tiles, kSortThreads, dyn, st, in, out, n, shift, ghist, ticket, status);
  return GF_OK;
}

// One attempt = 3 + P kernel launches and ONE host synchronisation, whatever the batch touches.  Capacity problems
// (vertex table, edge-id table, arena) and a batch that is not in time order are detected on the device, leave the
// graph untouched, and make the host fix the cause and replay the batch.
static int add_edges_impl(gf_graph *g, const int64_t *src, const int64_t *dst, const float *ts, const int64_t *eid,
                          uint64_t n, int ptr_kind, cudaStream_t st, bool async) {
  if (n == 0) GF_FAIL(GF_EINVAL, "add_edges: empty batch (reference: CHECK_GT(src_nodes.size(), 0))");
  if (n >= (1ull << 30)) GF_FAIL(GF_EINVAL, "add_edges: batch of %llu edges exceeds 2^30-1", (unsigned long long)n);
  if (!src || !dst || !ts || !eid) GF_FAIL(GF_EINVAL, "add_edges: null array");
  if (async && ptr_kind != GF_PTR_DEVICE) GF_FAIL(GF_EINVAL, "add_edges_async: device arrays only");
  if (async && !g->pending.empty() && g->pending_stream != st) GF_TRY(flush_pending(g));  // one stream per queue
  if (async && g->pending.size() + 2 >= kCallRing) GF_TRY(flush_pending(g));              // a slot per queued batch
  if (!async) GF_TRY(flush_pending(g));
  GF_TRY(set_device(g));
  g->prof.begin(st);
  // ---- stage host input
  if (ptr_kind == GF_PTR_HOST) {
    size_t off_dst = align_up(n * 8, 256), off_eid = 2 * off_dst, off_ts = 3 * off_dst;
    GF_TRY(g->s_in.reserve(off_ts + align_up(n * 4, 256), st));
    char *b = g->s_in.as<char>();
    GF_CUDA(cudaMemcpyAsync(b, src, n * 8, cudaMemcpyHostToDevice, st));
    GF_CUDA(cudaMemcpyAsync(b + off_dst, dst, n * 8, cudaMemcpyHostToDevice, st));
    GF_CUDA(cudaMemcpyAsync(b + off_eid, eid, n * 8, cudaMemcpyHostToDevice, st));
    GF_CUDA(cudaMemcpyAsync(b + off_ts, ts, n * 4, cudaMemcpyHostToDevice, st));
    src = (const int64_t *)b;
    dst = (const int64_t *)(b + off_dst);
    eid = (const int64_t *)(b + off_eid);
    ts = (const float *)(b + off_ts);
  } else if (ptr_kind != GF_PTR_DEVICE) {
    GF_FAIL(GF_EINVAL, "add_edges: bad ptr_kind %d", ptr_kind);
  }
  const size_t na = align_up(n + 1, 64);
  const StoreParams sp = {(uint32_t)g->cfg.minimum_block_size, g->cfg.insertion_policy, g->cfg.adaptive_block_size};
  const uint32_t tiles_p = cdiv(n, kPlanTile), tiles_s = ingest_sort_tiles(n);
  const bool big = ingest_rounds(n) == kIngestRoundsBig;
  const int64_t *src_in = src, *dst_in = dst, *eid_in = eid;
  const float *ts_in = ts;

  if (g->calls_since_defrag != ~0ull) g->calls_since_defrag++;
  for (int attempt = 0; attempt < 12; attempt++) {
    const unsigned parity = g->call_parity;
    g->call_parity = (parity + 1) % kCallRing;
    CallScratch *cur = &g->d_stats->call[parity], *nxt = &g->d_stats->call[g->call_parity];
    int bits = bit_width_u64(g->table_cap ? (uint64_t)g->table_cap - 1 : 0);
    if (bits < 1) bits = 1;
    const int passes = (bits + 7) / 8;
    // ---- scratch (sizes depend on the table capacity, which a replay may have grown)
    GF_TRY(g->s_sort.reserve(2 * na * 24, st));
    GF_TRY(g->s_seg.reserve(na * 4 + na * sizeof(SegRec), st));
    const size_t w_hist = kSortMaxPasses * 256, w_tick = 64, w_sort = (size_t)passes * tiles_s * 256,
                 w_a = 2 * (size_t)tiles_p, w_b = 2 * kNumClasses;
    const size_t ctl_words = w_hist + w_tick + w_sort + w_a + w_b;
    if (ctl_words * 4 > g->s_ctl.cap) {
      GF_TRY(g->s_ctl.reserve(ctl_words * 4, st));
      GF_CUDA(cudaMemsetAsync(g->s_ctl.ptr, 0, g->s_ctl.cap, st));  // afterwards every call leaves it clean
    }
    uint32_t *ghist = g->s_ctl.as<uint32_t>(), *tickets = ghist + w_hist, *sort_status = tickets + w_tick;
    unsigned long long *stat_a = reinterpret_cast<unsigned long long *>(sort_status + w_sort);  // even word offset
    uint32_t *gcls = reinterpret_cast<uint32_t *>(stat_a + tiles_p);
    // pushes this batch may make: the old directory and / or the old payload of each of its source vertices
    const uint64_t max_push = 2 * std::min<uint64_t>(n, g->table_cap ? g->table_cap : n) + 64;
    GF_TRY(ensure_log(g, max_push, st));
    g->log_upper += max_push;
    char *sb = g->s_sort.as<char>();
    SortDst set[2];
    for (int k = 0; k < 2; k++) {
      char *b = sb + (size_t)k * na * 24;
      set[k].dst = reinterpret_cast<int64_t *>(b);
      set[k].eid = set[k].dst + na;
      set[k].key = reinterpret_cast<uint32_t *>(set[k].eid + na);
      set[k].ts = reinterpret_cast<float *>(set[k].key + na);
    }
    uint32_t *segid = g->s_seg.as<uint32_t>();
    SegRec *recs = reinterpret_cast<SegRec *>(segid + na);
    const bool fast = !g->expect_unsorted;
    // ---- slow path: put the batch in time order first (stable), then it is an ordinary batch
    src = src_in; dst = dst_in; eid = eid_in; ts = ts_in;
    if (!fast) {
      GF_TRY(g->s_pre.reserve(4 * na * 4 + align_up(radix_tmp_elems(n), 64) * 4 + 3 * na * 8 + na * 4, st));
      uint32_t *k0 = g->s_pre.as<uint32_t>(), *v0 = k0 + na, *k1 = v0 + na, *v1 = k1 + na, *stmp = v1 + na;
      char *pb = reinterpret_cast<char *>(stmp + align_up(radix_tmp_elems(n), 64));
      int64_t *psrc = reinterpret_cast<int64_t *>(pb), *pdst = psrc + na, *peid = pdst + na;
      float *pts = reinterpret_cast<float *>(peid + na);
      const unsigned nb = cdiv(n, kThreads);
      gf::launch(keys_from_ts_kernel, nb, kThreads, 0, st, ts_in, n, k0, v0);
      bool in0 = true;
      GF_TRY(radix_sort_pairs(k0, v0, k1, v1, n, 0, 32, stmp, &in0, st));
      gf::launch(ingest_permute_kernel, nb, kThreads, 0, st, in0 ? v0 : v1, n, src_in, dst_in, ts_in, eid_in, psrc, pdst,
                 pts, peid);
      src = psrc; dst = pdst; eid = peid; ts = pts;
    }
    // ---- pass 0: validation flags, id ranges, digit histograms
    const unsigned prep_blocks = std::max(1u, std::min(cdiv(n, kThreads * 2), 148u * 4));
    gf::launch(ingest_prep_kernel, prep_blocks, kThreads, 0, st, src, dst, ts, eid, n, (uint64_t)g->table_cap,
               g->eid_base, (uint64_t)g->eid_cap, fast ? 1 : 0, passes, ghist, cur, nxt);
    g->prof.end(0, st);
    // ---- stable sort by source vertex, the payload travelling along
    {
      SortSrc in = {src, nullptr, ts, dst, eid};
      for (int p = 0; p < passes; p++) {
        const SortDst &out = set[p & 1];
        uint32_t *stat = sort_status + (size_t)p * tiles_s * 256;
        if (p == 0) {
          if (big) GF_TRY((launch_sort_pass<kIngestRoundsBig, true>(in, out, n, 0, ghist, tickets, stat, tiles_s, st)));
          else GF_TRY((launch_sort_pass<kIngestRoundsSmall, true>(in, out, n, 0, ghist, tickets, stat, tiles_s, st)));
        } else {
          if (big) GF_TRY((launch_sort_pass<kIngestRoundsBig, false>(in, out, n, 8 * p, ghist + p * 256, tickets + p, stat, tiles_s, st)));
          else GF_TRY((launch_sort_pass<kIngestRoundsSmall, false>(in, out, n, 8 * p, ghist + p * 256, tickets + p, stat, tiles_s, st)));
        }
        in = SortSrc{nullptr, out.key, out.ts, out.dst, out.eid};
      }
    }
    const SortDst &sorted = set[(passes - 1) & 1];
    g->prof.end(1, st);
    // ---- segments, block-sizing policy, allocation, accept / reject
    PlanArgs pa = {sorted.key, sorted.ts, n, g->d_table, sp, segid, recs, g->d_stats, cur, g->d_classes + parity,
                   tickets + kSortMaxPasses, stat_a, gcls, async ? 1 : 0};
    static const bool plan_small_tiles = getenv("GNNFLOW_B200_PLAN_SMALL_TILES") != nullptr;  // evidence knob
    // Large batches of long segments (>= 64 edges per entry of the vertex table; the table of a graph that has just been
    // cleared keeps its size): big plan tiles and the upkeep pass of its own.  With the one-edge segments of the
    // 16.7 M-vertex shape a thread would plan up to 16 segments in turn (5 % slower) and the upkeep pass costs 2 %.
    const bool long_segments = n >= (1u << 20) && (uint64_t)g->table_cap * 64 <= n;
    if (long_segments && !plan_small_tiles)
      gf::launch_pdl(ingest_plan_kernel<kPlanEptBig>, cdiv(n, kThreads * kPlanEptBig), kThreads, 0, st, pa);
    else
      gf::launch_pdl(ingest_plan_kernel<kPlanEpt>, tiles_p, kThreads, 0, st, pa);
    g->prof.end(2, st);
    if (g->cfg.insertion_policy == GF_INSERTION_REPLACE)
      gf::launch_pdl(ingest_realloc_copy_kernel, std::min(cdiv(n, 4), 148u * 8), kThreads, 0, st, recs, cur, g->d_classes + parity,
                     g->d_sorted[g->sorted_cur]);
    g->prof.end(3, st);
    // ---- payload, descriptors, directories, bookkeeping, report
    ApplyArgs aa = {sorted.ts, sorted.dst, sorted.eid, dst, eid, n, segid, recs, g->d_table, g->d_is_src, g->d_is_node,
                    g->d_eid_ref, g->eid_base, g->d_stats, cur, g->d_classes + parity, g->d_sorted[g->sorted_cur], g->d_log,
                    g->h_res + parity, g->s_ctl.as<uint32_t>(), (uint64_t)ctl_words, sp, 0, nullptr, 0};
    static const bool no_split = getenv("GNNFLOW_B200_NO_BOOKKEEP_SPLIT") != nullptr;  // evidence knob
    const uint32_t book_words = std::max(1u, cdiv(g->table_cap, 32));
    if (long_segments && !no_split && book_words <= kBookMaxWords) {
      aa.separate_bookkeeping = 1;
      const unsigned slabs = std::min(cdiv(n, kThreads * 16), 148u * 4);
      GF_TRY(g->s_book.reserve((size_t)slabs * book_words * 4, st));
      aa.book_bitmaps = g->s_book.as<uint32_t>();
      aa.book_words = book_words;
      gf::launch_pdl(ingest_bookkeep_kernel, slabs, kThreads, (size_t)book_words * 4, st, aa);
      gf::launch_pdl(ingest_flags_merge_kernel, dim3(cdiv(book_words, kThreads), kBookMergeSlices), kThreads, 0, st, aa,
                     (uint32_t)slabs, (uint64_t)g->table_cap);
    }
    static const uint32_t ept_big = getenv("GNNFLOW_B200_APPLY_EPT") ? (uint32_t)atoi(getenv("GNNFLOW_B200_APPLY_EPT")) : 4u;  // knob
    const uint32_t ept = n >= (1u << 20) ? std::max(1u, ept_big) : 1u;
    *(volatile unsigned int *)&g->h_res[parity].call.done_ctas = 0;  // set again by the last CTA of the apply kernel
    gf::launch_pdl(ingest_apply_kernel, cdiv(n, (uint64_t)kThreads * ept), kThreads, 0, st, aa, ept);
    GF_CUDA(cudaGetLastError());
    g->prof.end(4, st, false);
    if (async) {  // the caller keeps the arrays alive until the next flush; the outcome is looked at there
      g->pending.push_back({src_in, dst_in, ts_in, eid_in, n, parity});
      g->pending_stream = st;
      return GF_OK;
    }
    // the reference returns after cudaStreamSynchronize (dynamic_graph.cu:135-137); this is the only wait: for the report
    // of the apply kernel's last CTA, which is written when the batch has been applied
    GF_TRY(gf::wait_flag_or_sync(&g->h_res[parity].call.done_ctas, st));
    const HostResult hr = g->h_res[parity];
    const CallScratch hs = hr.call;
    g->log_upper = hr.log_cnt;
    g->sorted_upper = hr.sorted_cnt;
    g->h_stats->num_edges = hr.num_edges;
    g->h_stats->num_blocks = hr.num_blocks;
    g->h_stats->allocated_elems = hr.allocated_elems;
    const uint32_t f = hs.error_flags;
    if (f & kErrBadId) GF_FAIL(GF_EINVAL, "add_edges: vertex ids must lie in [0, 2^32)");
    if (f & kErrBadEid) GF_FAIL(GF_EINVAL, "add_edges: edge ids must not be negative");
    if (f & (kErrTableSmall | kErrEidSmall | kErrEidLow | kErrUnsorted | kErrArena)) {  // fix the cause, replay the batch
      if (f & kErrTableSmall) {
        const int64_t keep_max = g->max_node_id;
        const bool keep_has = g->has_nodes;
        GF_TRY(ensure_table(g, hs.max_id, st));
        g->max_node_id = keep_max;  // the table grew; the graph has not changed yet
        g->has_nodes = keep_has;
      }
      if (f & (kErrEidSmall | kErrEidLow)) GF_TRY(ensure_eids(g, 0x7fffffffffffffffll - hs.max_neg_eid, hs.max_eid, st));
      if (f & kErrUnsorted) g->expect_unsorted = true;
      if ((f & kErrArena) && !(f & (kErrTableSmall | kErrEidSmall | kErrEidLow | kErrUnsorted))) {
        const uint64_t need = (uint64_t)hs.total_units * kUnit, free_bytes = hr.free_units * kUnit;
        if (hr.log_cnt) {
          GF_TRY(arena_merge(g, st));  // blocks freed since the last merge may be all that is missing
        } else if (g->calls_since_defrag >= 32 && free_bytes >= need && free_bytes >= g->arena_total / 8) {
          GF_TRY(arena_defrag(g, st));  // plenty is free, only not in the classes being asked for
        } else {
          GF_TRY(arena_add_chunk(g, (size_t)need, st));
        }
      }
      if (g->prof.on) g->prof.begin(st);
      continue;
    }
    if (f & kErrOutOfOrder) GF_FAIL(GF_EORDER, "add_edges: timestamps are older than the existing edges in the graph");
    if (!hs.accepted) GF_FAIL(GF_ECUDA, "add_edges: internal error (batch neither accepted nor flagged)");
    // success: host mirrors (DynamicGraph::AddNodes, dynamic_graph.cu:140-147)
    if (!g->has_nodes || hs.max_id > g->max_node_id) g->max_node_id = hs.max_id;
    g->has_nodes = true;
    g->counts_dirty = true;
    if (!fast && !hs.unsorted) g->expect_unsorted = false;  // a time-ordered stream resumes the fast path
    // blocks freed by this batch (replace policy, directory growth) become allocatable once enough have piled up
    if (hr.log_cnt > 4096 && hr.log_cnt > g->sorted_upper / 4) GF_TRY(arena_merge(g, st));
    return GF_OK;
  }
  GF_FAIL(GF_ECUDA, "add_edges: the batch could not be applied after 12 attempts");
}

static int refresh_counts(gf_graph *g) {
  if (!g->counts_dirty) return GF_OK;
  GF_TRY(set_device(g));
  GF_TRY(settle(g));
  cudaStream_t st = 0;
  size_t len = g->table_len();
  unsigned long long *d = &g->d_stats->call_count;
  unsigned long long h[2] = {0, 0};
  for (int k = 0; k < 2; k++) {
    GF_CUDA(cudaMemsetAsync(d, 0, sizeof(*d), st));
    if (len) gf::launch(count_flags_kernel, min(cdiv(len, kThreads), 148u * 8), kThreads, 0, st, k ? g->d_is_src : g->d_is_node, len, d);
    GF_CUDA(cudaMemcpyAsync(&h[k], d, sizeof(*d), cudaMemcpyDeviceToHost, st));
    GF_CUDA(cudaStreamSynchronize(st));
  }
  g->num_nodes = h[0];
  g->num_src_nodes = h[1];
  g->counts_dirty = false;
  return GF_OK;
}

static int read_entry(gf_graph *g, int64_t v, NodeEntry *ent, std::vector<BlockDesc> *descs) {
  memset(ent, 0, sizeof(*ent));
  descs->clear();
  if (v < 0 || (size_t)v >= g->table_len()) return GF_OK;
  GF_TRY(set_device(g));
  GF_TRY(settle(g));
  GF_CUDA(cudaMemcpy(ent, g->d_table + v, sizeof(NodeEntry), cudaMemcpyDeviceToHost));
  uint32_t nlive = ent->end - ent->first;
  if (nlive) {
    descs->resize(nlive);
    GF_CUDA(cudaMemcpy(descs->data(), reinterpret_cast<const BlockDesc *>(ent->dir()) + ent->first,
                       nlive * sizeof(BlockDesc), cudaMemcpyDeviceToHost));
  }
  return GF_OK;
}

static int flags_to_list(gf_graph *g, const uint8_t *d_flags, size_t len, int64_t *out, uint64_t cap, uint64_t *count) {
  GF_TRY(set_device(g));
  std::vector<uint8_t> h(len);
  GF_TRY(settle(g));
  if (len) GF_CUDA(cudaMemcpy(h.data(), d_flags, len, cudaMemcpyDeviceToHost));
  uint64_t k = 0;
  for (size_t i = 0; i < len; i++)
    if (h[i]) {
      if (out && k < cap) out[k] = (int64_t)i;
      k++;
    }
  *count = k;
  if (out && cap < k) GF_FAIL(GF_ECAPACITY, "output buffer holds %llu entries, %llu needed", (unsigned long long)cap, (unsigned long long)k);
  return GF_OK;
}

// SaveToFile, temporal_block_allocator.cu:182-222
static int save_block_file(gf_graph *g, int64_t v, const BlockDesc &d, uint64_t prev, uint64_t next) {
  if ((size_t)v >= g->saved_blocks_per_node.size()) g->saved_blocks_per_node.resize(v + 1, 0);
  char name[128];
  snprintf(name, sizeof(name), "temporal_block_%lld-%u.bin", (long long)v, g->saved_blocks_per_node[v]);
  std::vector<int64_t> hd(d.size), he(d.size), hp(2 * (size_t)d.size);
  std::vector<float> ht(d.size);
  GF_CUDA(cudaMemcpy(hp.data(), (const void *)(d.payload + payload_ts_bytes(d.capacity)), d.size * 16ull, cudaMemcpyDeviceToHost));
  GF_CUDA(cudaMemcpy(ht.data(), (const void *)d.payload, d.size * 4ull, cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < d.size; i++) unpack_rec_host(&hp[2 * (size_t)i], &hd[i], &he[i]);
  FILE *f = fopen(name, "wb");
  if (!f) GF_FAIL(GF_EINVAL, "cannot open %s for writing", name);
  size_t size = d.size, capacity = g->cfg.insertion_policy == GF_INSERTION_REPLACE
                                       ? replace_logical_cap(d.size, (uint32_t)g->cfg.minimum_block_size) : d.capacity;
  fwrite(&size, sizeof(size), 1, f);
  fwrite(&capacity, sizeof(capacity), 1, f);
  fwrite(&d.start_ts, 4, 1, f);
  fwrite(&d.end_ts, 4, 1, f);
  fwrite(hd.data(), 8, d.size, f);
  fwrite(ht.data(), 4, d.size, f);
  fwrite(he.data(), 8, d.size, f);
  fwrite(&prev, 8, 1, f);
  fwrite(&next, 8, 1, f);
  fclose(f);
  g->saved_blocks_per_node[v]++;
  return GF_OK;
}

}  // namespace gf

using namespace gf;

// ============================================================================================== C ABI
GF_EXPORT const char *gf_last_error(void) { return gf::get_error(); }
GF_EXPORT int gf_abi_version(void) { return GF_ABI_VERSION; }

GF_EXPORT int gf_graph_create(const gf_graph_config *cfg, gf_graph **out) {
  if (!cfg || !out) GF_FAIL(GF_EINVAL, "gf_graph_create: null argument");
  if (cfg->insertion_policy != GF_INSERTION_INSERT && cfg->insertion_policy != GF_INSERTION_REPLACE)
    GF_FAIL(GF_EINVAL, "Invalid insertion policy: %d", cfg->insertion_policy);
  if (cfg->mem_resource_type < GF_MEM_CUDA || cfg->mem_resource_type > GF_MEM_SHARED)
    GF_FAIL(GF_EINVAL, "Invalid memory resource type: %d", cfg->mem_resource_type);
  if (cfg->minimum_block_size >= (1ull << 31)) GF_FAIL(GF_EINVAL, "minimum_block_size too large");
  if (cfg->maximum_pool_size && cfg->maximum_pool_size < cfg->initial_pool_size)
    GF_FAIL(GF_EINVAL, "maximum_pool_size < initial_pool_size");
  int ndev = 0;
  GF_CUDA(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) GF_FAIL(GF_EINVAL, "device %d out of range (%d devices)", cfg->device, ndev);
  GF_CUDA(cudaSetDevice(cfg->device));
  apply_l2_fetch_default(cfg->device);
  gf_graph *g = new gf_graph();
  g->cfg = *cfg;
  cudaError_t e = cudaMalloc(&g->d_stats, sizeof(GraphStats));
  if (e == cudaSuccess) e = cudaMemset(g->d_stats, 0, sizeof(GraphStats));
  if (e == cudaSuccess) e = cudaMalloc(&g->d_classes, sizeof(CallClasses) * kCallRing);
  if (e == cudaSuccess) e = cudaMallocHost(&g->h_stats, sizeof(GraphStats));
  if (e == cudaSuccess) e = cudaHostAlloc(&g->h_res, sizeof(HostResult) * kCallRing, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    gf_graph_destroy(g);
    GF_FAIL(GF_ECUDA, "gf_graph_create: %s", cudaGetErrorString(e));
  }
  memset(g->h_stats, 0, sizeof(GraphStats));
  memset(g->h_res, 0, sizeof(HostResult) * kCallRing);
  g->prof.init(GF_GRAPH_PHASES);
  if (cfg->initial_pool_size) {  // the reference's pool resource reserves initial_pool_size up front as well
    int rc = arena_add_chunk(g, 0, 0);
    if (rc != GF_OK) {
      gf_graph_destroy(g);
      return rc;
    }
  }
  *out = g;
  return GF_OK;
}

GF_EXPORT int gf_graph_destroy(gf_graph *g) {
  if (!g) return GF_OK;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    if (--g->refs > 0) return GF_OK;
  }
  cudaSetDevice(g->cfg.device);
  cudaDeviceSynchronize();
  g->prof.destroy();
  for (auto &c : g->chunks) cudaFree(c.base);
  if (g->d_table) cudaFree(g->d_table);
  if (g->d_is_node) cudaFree(g->d_is_node);
  if (g->d_is_src) cudaFree(g->d_is_src);
  if (g->d_eid_ref) cudaFree(g->d_eid_ref);
  if (g->d_stats) cudaFree(g->d_stats);
  if (g->d_classes) cudaFree(g->d_classes);
  if (g->d_log) cudaFree(g->d_log);
  if (g->d_sorted[0]) cudaFree(g->d_sorted[0]);
  if (g->d_sorted[1]) cudaFree(g->d_sorted[1]);
  if (g->h_stats) cudaFreeHost(g->h_stats);
  if (g->h_res) cudaFreeHost(g->h_res);
  g->s_in.release();
  g->s_sort.release();
  g->s_seg.release();
  g->s_misc.release();
  g->s_ctl.release();
  g->s_pre.release();
  g->s_book.release();
  delete g;
  return GF_OK;
}

GF_EXPORT int gf_graph_add_edges(gf_graph *g, const int64_t *src, const int64_t *dst, const float *ts,
                                 const int64_t *eid, uint64_t n, int ptr_kind, void *stream) {
  if (!g) GF_FAIL(GF_EINVAL, "null graph");
  std::lock_guard<std::mutex> lk(g->mu);
  return add_edges_impl(g, src, dst, ts, eid, n, ptr_kind, (cudaStream_t)stream, false);
}

GF_EXPORT int gf_graph_add_edges_async(gf_graph *g, const int64_t *src, const int64_t *dst, const float *ts,
                                       const int64_t *eid, uint64_t n, void *stream) {
  if (!g) GF_FAIL(GF_EINVAL, "null graph");
  std::lock_guard<std::mutex> lk(g->mu);
  return add_edges_impl(g, src, dst, ts, eid, n, GF_PTR_DEVICE, (cudaStream_t)stream, true);
}

GF_EXPORT int gf_graph_flush(gf_graph *g) {
  if (!g) GF_FAIL(GF_EINVAL, "null graph");
  std::lock_guard<std::mutex> lk(g->mu);
  return flush_pending(g);
}

// for the sampler and the cache (other translation units): settle queued batches before the graph is read
int gf_graph_flush_internal(gf_graph *g) {
  std::lock_guard<std::mutex> lk(g->mu);
  return flush_pending(g);
}

GF_EXPORT int gf_graph_clear(gf_graph *g, void *stream) {
  if (!g) GF_FAIL(GF_EINVAL, "null graph");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  cudaStream_t st = (cudaStream_t)stream;
  GF_TRY(set_device(g));
  size_t len = g->table_len();
  if (len) {
    GF_CUDA(cudaMemsetAsync(g->d_table, 0, len * sizeof(NodeEntry), st));
    GF_CUDA(cudaMemsetAsync(g->d_is_node, 0, len, st));
    GF_CUDA(cudaMemsetAsync(g->d_is_src, 0, len, st));
  }
  if (g->eid_cap) GF_CUDA(cudaMemsetAsync(g->d_eid_ref, 0, g->eid_cap * sizeof(uint32_t), st));
  GF_CUDA(cudaMemsetAsync(g->d_stats, 0, sizeof(GraphStats), st));
  memset(g->h_stats, 0, sizeof(GraphStats));
  g->log_upper = g->sorted_upper = 0;
  // every chunk is a whole bump region again (the memset above zeroed the device's list)
  if (!g->chunks.empty()) {
    std::vector<ArenaRegion> regs;
    unsigned int hdr[2] = {(unsigned int)g->chunks.size(), 0};
    for (size_t k = 0; k < g->chunks.size(); k++) {
      regs.push_back({(unsigned long long)(uintptr_t)g->chunks[k].base, (unsigned long long)(uintptr_t)(g->chunks[k].base + g->chunks[k].size)});
      if (g->chunks[k].size > g->chunks[hdr[1]].size) hdr[1] = (unsigned int)k;
    }
    GF_CUDA(cudaMemcpyAsync(g->d_stats->arena.regions, regs.data(), regs.size() * sizeof(ArenaRegion), cudaMemcpyHostToDevice, st));
    GF_CUDA(cudaMemcpyAsync(&g->d_stats->arena.num_regions, hdr, sizeof(hdr), cudaMemcpyHostToDevice, st));
    GF_CUDA(cudaStreamSynchronize(st));  // host sources
    g->num_regions = hdr[0];
  }
  g->unsettled_stream = st;
  g->unsettled = true;
  g->call_parity = 0;
  g->max_node_id = 0;
  g->has_nodes = false;
  g->counts_dirty = false;
  g->num_nodes = g->num_src_nodes = 0;
  g->saved_blocks_per_node.clear();
  return GF_OK;
}

GF_EXPORT int gf_graph_offload_old_blocks(gf_graph *g, float timestamp, int to_file, uint64_t *num_blocks,
                                          void *stream) {
  if (!g) GF_FAIL(GF_EINVAL, "null graph");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  cudaStream_t st = (cudaStream_t)stream;
  GF_TRY(set_device(g));
  if (num_blocks) *num_blocks = 0;
  size_t len = g->table_len();
  if (!len) return GF_OK;
  GF_CUDA(cudaMemsetAsync(&g->d_stats->call_count, 0, sizeof(unsigned long long), st));
  uint2 *drops = nullptr;
  uint32_t drops_cap = 0;
  GF_TRY(pull_stats(g, st));  // how many blocks there are = how many may be freed
  GF_TRY(ensure_log(g, g->h_stats->num_blocks + 64, st));
  if (to_file) {  // the dropped descriptors stay readable in the directories
    drops_cap = (uint32_t)g->h_stats->num_blocks;
    GF_TRY(g->s_misc.reserve((size_t)drops_cap * sizeof(uint2) + 16, st));
    drops = g->s_misc.as<uint2>();
  }
  gf::launch(offload_kernel, cdiv(len * 32, kThreads), kThreads, 0, st, g->d_table, g->d_is_node, len, timestamp, g->d_eid_ref,
             g->eid_base, g->d_stats, g->d_log, drops, drops_cap,
             StoreParams{(uint32_t)g->cfg.minimum_block_size, g->cfg.insertion_policy, g->cfg.adaptive_block_size});
  GF_CUDA(cudaGetLastError());
  GF_TRY(pull_stats(g, st));
  uint64_t nd = g->h_stats->call_count;
  if (num_blocks) *num_blocks = nd;
  if (to_file && nd) {
    std::vector<uint2> h(nd);
    GF_CUDA(cudaMemcpy(h.data(), drops, nd * sizeof(uint2), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end(), [](const uint2 &a, const uint2 &b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    for (auto &d : h) {
      NodeEntry ent;
      GF_CUDA(cudaMemcpy(&ent, g->d_table + d.x, sizeof(ent), cudaMemcpyDeviceToHost));
      BlockDesc bd;
      const BlockDesc *dir = reinterpret_cast<const BlockDesc *>(ent.dir());
      GF_CUDA(cudaMemcpy(&bd, dir + d.y, sizeof(bd), cudaMemcpyDeviceToHost));
      uint64_t prev = d.y > 0 ? (uint64_t)(uintptr_t)(dir + d.y - 1) : 0;
      uint64_t next = d.y + 1 < ent.end ? (uint64_t)(uintptr_t)(dir + d.y + 1) : 0;
      GF_TRY(save_block_file(g, d.x, bd, prev, next));
    }
  }
  // TemporalBlockAllocator::Deallocate (temporal_block_allocator.cu:110-113): the dropped payloads are allocatable
  // again from the next batch on
  GF_TRY(arena_merge(g, st));
  if (nd) GF_TRY(compact_eids(g, st));  // edges_.erase(eid), dynamic_graph.cu:394-396: dead ids stop costing memory
  return GF_OK;
}

// ---------------------------------------------------------------------------------------------- checkpoint
// Whole-graph checkpoint (the reference has none): a raw image of what the store holds in HBM -- the
// used part of every arena chunk, the vertex table and flags, the edge-id reference counts, the allocator state with its
// free lists -- plus the host mirrors.  Loading allocates chunks of the same sizes and moves every stored device address
// from the old chunk to the new one (reloc_*_kernel); the loaded graph is bit-identical to the saved one, offloaded
// prefixes, free blocks and block capacities included, and goes on ingesting where the saved one stopped.
namespace {
struct CkptHeader {
  char magic[8];  // "GFB200CK"
  uint32_t version, num_chunks;
  gf_graph_config cfg;
  uint64_t table_cap, eid_cap, log_cap, sorted_cap, log_cnt, sorted_cnt, num_nodes, num_src_nodes, saved_nodes;
  int64_t max_node_id, eid_base;
  uint32_t has_nodes, counts_dirty, expect_unsorted, sorted_cur;
};
struct CkptChunk {
  uint64_t base, size, used;
};
constexpr uint32_t kCkptVersion = 2;  // 2: payload records {dst u32, ts, eid}
constexpr size_t kCkptStage = 32u << 20;

int dev_to_file(FILE *f, const void *dev, size_t bytes, void *stage) {
  const char *p = static_cast<const char *>(dev);
  while (bytes) {
    const size_t c = std::min(bytes, kCkptStage);
    GF_CUDA(cudaMemcpy(stage, p, c, cudaMemcpyDeviceToHost));
    if (fwrite(stage, 1, c, f) != c) GF_FAIL(GF_EINVAL, "checkpoint: short write");
    p += c;
    bytes -= c;
  }
  return GF_OK;
}
int file_to_dev(FILE *f, void *dev, size_t bytes, void *stage) {
  char *p = static_cast<char *>(dev);
  while (bytes) {
    const size_t c = std::min(bytes, kCkptStage);
    if (fread(stage, 1, c, f) != c) GF_FAIL(GF_EINVAL, "checkpoint: short read");
    GF_CUDA(cudaMemcpy(p, stage, c, cudaMemcpyHostToDevice));
    p += c;
    bytes -= c;
  }
  return GF_OK;
}
}  // namespace

GF_EXPORT int gf_graph_save(gf_graph *g, const char *path) {
  if (!g || !path) GF_FAIL(GF_EINVAL, "gf_graph_save: null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  GF_TRY(set_device(g));
  GF_TRY(settle(g));
  GF_CUDA(cudaDeviceSynchronize());  // a checkpoint is a cold path: whatever stream the last batch ran on has drained
  GF_TRY(refresh_counts(g));
  GF_TRY(pull_stats(g, 0));
  const ArenaState &ar = g->h_stats->arena;
  CkptHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "GFB200CK", 8);
  h.version = kCkptVersion;
  h.num_chunks = (uint32_t)g->chunks.size();
  h.cfg = g->cfg;
  h.table_cap = g->table_cap; h.eid_cap = g->eid_cap; h.log_cap = g->log_cap; h.sorted_cap = g->sorted_cap;
  h.log_cnt = ar.log_cnt; h.sorted_cnt = ar.sorted_cnt;
  h.num_nodes = g->num_nodes; h.num_src_nodes = g->num_src_nodes; h.saved_nodes = g->saved_blocks_per_node.size();
  h.max_node_id = g->max_node_id; h.eid_base = g->eid_base;
  h.has_nodes = g->has_nodes; h.counts_dirty = g->counts_dirty; h.expect_unsorted = g->expect_unsorted;
  h.sorted_cur = (uint32_t)g->sorted_cur;
  std::vector<CkptChunk> ck;
  for (auto &c : g->chunks) {
    const uint64_t base = (uint64_t)(uintptr_t)c.base;
    uint64_t used = c.size;  // a bump region that still ends where the chunk ends: nothing beyond its pointer is in use
    for (unsigned k = 0; k < ar.num_regions; k++)
      if (ar.regions[k].end == base + c.size && ar.regions[k].cur >= base) used = ar.regions[k].cur - base;
    ck.push_back({base, (uint64_t)c.size, used});
  }
  FILE *f = fopen(path, "wb");
  if (!f) GF_FAIL(GF_EINVAL, "cannot open %s for writing", path);
  void *stage = nullptr;
  int rc = GF_OK;
  auto body = [&]() -> int {
    GF_CUDA(cudaMallocHost(&stage, kCkptStage));
    if (fwrite(&h, sizeof(h), 1, f) != 1) GF_FAIL(GF_EINVAL, "checkpoint: short write");
    if (!ck.empty() && fwrite(ck.data(), sizeof(CkptChunk), ck.size(), f) != ck.size()) GF_FAIL(GF_EINVAL, "checkpoint: short write");
    if (fwrite(g->h_stats, sizeof(GraphStats), 1, f) != 1) GF_FAIL(GF_EINVAL, "checkpoint: short write");
    if (h.saved_nodes && fwrite(g->saved_blocks_per_node.data(), 4, h.saved_nodes, f) != h.saved_nodes)
      GF_FAIL(GF_EINVAL, "checkpoint: short write");
    for (auto &c : ck) GF_TRY(dev_to_file(f, (const void *)(uintptr_t)c.base, c.used, stage));
    if (g->table_cap) {
      GF_TRY(dev_to_file(f, g->d_table, g->table_cap * sizeof(NodeEntry), stage));
      GF_TRY(dev_to_file(f, g->d_is_node, g->table_cap, stage));
      GF_TRY(dev_to_file(f, g->d_is_src, g->table_cap, stage));
    }
    if (g->eid_cap) GF_TRY(dev_to_file(f, g->d_eid_ref, g->eid_cap * 4, stage));
    if (h.log_cnt) GF_TRY(dev_to_file(f, g->d_log, h.log_cnt * sizeof(FreeRec), stage));
    if (g->sorted_cap) GF_TRY(dev_to_file(f, g->d_sorted[g->sorted_cur], g->sorted_cap * 8, stage));
    return GF_OK;
  };
  rc = body();
  if (stage) cudaFreeHost(stage);
  if (fclose(f) != 0 && rc == GF_OK) {
    set_error("checkpoint: closing %s failed", path);
    rc = GF_EINVAL;
  }
  return rc;
}

GF_EXPORT int gf_graph_load(const char *path, int device, gf_graph **out) {
  if (!path || !out) GF_FAIL(GF_EINVAL, "gf_graph_load: null argument");
  FILE *f = fopen(path, "rb");
  if (!f) GF_FAIL(GF_EINVAL, "cannot open %s", path);
  CkptHeader h;
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "GFB200CK", 8) != 0 || h.version != kCkptVersion ||
      h.num_chunks > kMaxRegions) {
    fclose(f);
    GF_FAIL(GF_EINVAL, "%s is not a graph checkpoint of this version", path);
  }
  gf_graph_config cfg = h.cfg;
  cfg.device = device;
  const uint64_t initial = cfg.initial_pool_size;
  cfg.initial_pool_size = 0;  // the chunks come from the file
  gf_graph *g = nullptr;
  int rc = gf_graph_create(&cfg, &g);
  if (rc != GF_OK) {
    fclose(f);
    return rc;
  }
  g->cfg.initial_pool_size = initial;
  void *stage = nullptr;
  RelocMap *d_map = nullptr;
  auto body = [&]() -> int {
    GF_CUDA(cudaMallocHost(&stage, kCkptStage));
    std::vector<CkptChunk> ck(h.num_chunks);
    if (h.num_chunks && fread(ck.data(), sizeof(CkptChunk), ck.size(), f) != ck.size()) GF_FAIL(GF_EINVAL, "checkpoint: short read");
    if (fread(g->h_stats, sizeof(GraphStats), 1, f) != 1) GF_FAIL(GF_EINVAL, "checkpoint: short read");
    g->saved_blocks_per_node.resize(h.saved_nodes);
    if (h.saved_nodes && fread(g->saved_blocks_per_node.data(), 4, h.saved_nodes, f) != h.saved_nodes)
      GF_FAIL(GF_EINVAL, "checkpoint: short read");
    RelocMap m;
    memset(&m, 0, sizeof(m));
    m.n = h.num_chunks;
    for (uint32_t k = 0; k < h.num_chunks; k++) {
      char *p = nullptr;
      cudaError_t e = cudaMalloc(&p, ck[k].size);
      if (e != cudaSuccess) {
        cudaGetLastError();
        GF_FAIL(GF_ENOMEM, "cudaMalloc(%llu) for the edge pool failed: %s", (unsigned long long)ck[k].size, cudaGetErrorString(e));
      }
      g->chunks.push_back({p, (size_t)ck[k].size});
      g->arena_total += ck[k].size;
      m.old_base[k] = ck[k].base; m.size[k] = ck[k].size; m.new_base[k] = (unsigned long long)(uintptr_t)p;
      GF_TRY(file_to_dev(f, p, ck[k].used, stage));
    }
    g->table_cap = h.table_cap;
    if (h.table_cap) {
      GF_CUDA(cudaMallocAsync(&g->d_table, h.table_cap * sizeof(NodeEntry), 0));
      GF_CUDA(cudaMallocAsync(&g->d_is_node, h.table_cap, 0));
      GF_CUDA(cudaMallocAsync(&g->d_is_src, h.table_cap, 0));
      GF_TRY(file_to_dev(f, g->d_table, h.table_cap * sizeof(NodeEntry), stage));
      GF_TRY(file_to_dev(f, g->d_is_node, h.table_cap, stage));
      GF_TRY(file_to_dev(f, g->d_is_src, h.table_cap, stage));
    }
    g->eid_cap = h.eid_cap;
    g->eid_base = h.eid_base;
    if (h.eid_cap) {
      GF_CUDA(cudaMallocAsync(&g->d_eid_ref, h.eid_cap * 4, 0));
      GF_TRY(file_to_dev(f, g->d_eid_ref, h.eid_cap * 4, stage));
    }
    g->log_cap = h.log_cap;
    if (h.log_cap) GF_CUDA(cudaMallocAsync(&g->d_log, h.log_cap * sizeof(FreeRec), 0));
    if (h.log_cnt) GF_TRY(file_to_dev(f, g->d_log, h.log_cnt * sizeof(FreeRec), stage));
    g->sorted_cap = h.sorted_cap;
    g->sorted_cur = 0;
    if (h.sorted_cap) {
      GF_CUDA(cudaMallocAsync(&g->d_sorted[0], h.sorted_cap * 8, 0));
      GF_CUDA(cudaMallocAsync(&g->d_sorted[1], h.sorted_cap * 8, 0));
      GF_TRY(file_to_dev(f, g->d_sorted[0], h.sorted_cap * 8, stage));
    }
    // device state: counters + allocator as saved (call ring and poison cleared), then the relocation
    GraphStats *hs = g->h_stats;
    hs->poison = 0;
    memset(hs->call, 0, sizeof(hs->call));
    GF_CUDA(cudaMemcpy(g->d_stats, hs, sizeof(GraphStats), cudaMemcpyHostToDevice));
    GF_CUDA(cudaMalloc(&d_map, sizeof(RelocMap)));
    GF_CUDA(cudaMemcpy(d_map, &m, sizeof(m), cudaMemcpyHostToDevice));
    if (h.table_cap) gf::launch(reloc_table_kernel, cdiv(h.table_cap * 32, kThreads), kThreads, 0, 0, g->d_table, (uint64_t)h.table_cap, d_map);
    gf::launch(reloc_lists_kernel, 148u * 4, kThreads, 0, 0, g->d_sorted[0], (uint64_t)h.sorted_cap, g->d_log, (uint64_t)h.log_cnt,
               &g->d_stats->arena, d_map);
    GF_CUDA(cudaGetLastError());
    GF_TRY(pull_stats(g, 0));
    g->num_regions = g->h_stats->arena.num_regions;
    g->max_node_id = h.max_node_id;
    g->has_nodes = h.has_nodes != 0;
    g->counts_dirty = h.counts_dirty != 0;
    g->expect_unsorted = h.expect_unsorted != 0;
    g->num_nodes = h.num_nodes;
    g->num_src_nodes = h.num_src_nodes;
    return GF_OK;
  };
  rc = body();
  if (d_map) cudaFree(d_map);
  if (stage) cudaFreeHost(stage);
  fclose(f);
  if (rc != GF_OK) {
    gf_graph_destroy(g);
    return rc;
  }
  *out = g;
  return GF_OK;
}

GF_EXPORT int gf_block_file_read(const char *path, uint64_t *size, uint64_t *capacity, float *start_ts, float *end_ts,
                                 int64_t *dst, float *ts, int64_t *eid, uint64_t cap) {
  if (!path || !size || !capacity || !start_ts || !end_ts) GF_FAIL(GF_EINVAL, "gf_block_file_read: null argument");
  if ((dst || ts || eid) && !(dst && ts && eid)) GF_FAIL(GF_EINVAL, "gf_block_file_read: dst / ts / eid go together");
  FILE *f = fopen(path, "rb");
  if (!f) GF_FAIL(GF_EINVAL, "cannot open %s", path);
  uint64_t hdr[2];
  float tt[2];
  bool ok = fread(hdr, 8, 2, f) == 2 && fread(tt, 4, 2, f) == 2;
  if (ok) {
    fseek(f, 0, SEEK_END);
    ok = hdr[0] <= hdr[1] && (uint64_t)ftell(f) == 24 + hdr[0] * 20 + 16;  // header, three arrays, prev / next
  }
  if (!ok) {
    fclose(f);
    GF_FAIL(GF_EINVAL, "%s is not a temporal block file", path);
  }
  *size = hdr[0];
  *capacity = hdr[1];
  *start_ts = tt[0];
  *end_ts = tt[1];
  if (dst) {
    if (cap < hdr[0]) {
      fclose(f);
      GF_FAIL(GF_ECAPACITY, "output arrays hold %llu entries, the block has %llu", (unsigned long long)cap, (unsigned long long)hdr[0]);
    }
    fseek(f, 24, SEEK_SET);
    ok = fread(dst, 8, hdr[0], f) == hdr[0] && fread(ts, 4, hdr[0], f) == hdr[0] && fread(eid, 8, hdr[0], f) == hdr[0];
  }
  fclose(f);
  if (!ok) GF_FAIL(GF_EINVAL, "short read from %s", path);
  return GF_OK;
}

GF_EXPORT int gf_graph_num_vertices(gf_graph *g, uint64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  GF_TRY(refresh_counts(g));
  *out = g->num_nodes;
  return GF_OK;
}
GF_EXPORT int gf_graph_num_source_vertices(gf_graph *g, uint64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  GF_TRY(refresh_counts(g));
  *out = g->num_src_nodes;
  return GF_OK;
}
GF_EXPORT int gf_graph_num_edges(gf_graph *g, uint64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  *out = g->h_stats->num_edges;
  return GF_OK;
}
GF_EXPORT int gf_graph_max_vertex_id(gf_graph *g, int64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  GF_TRY(gf_graph_flush_internal(g));
  *out = g->max_node_id;
  return GF_OK;
}
GF_EXPORT int gf_graph_avg_linked_list_length(gf_graph *g, float *out) {  // dynamic_graph.cu:359-366
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  GF_TRY(refresh_counts(g));
  float sum = (float)g->h_stats->num_blocks;
  *out = sum / (float)g->num_nodes;
  return GF_OK;
}
GF_EXPORT int gf_graph_memory_usage(gf_graph *g, float *out) {  // temporal_block_allocator.cu:155-156
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  GF_TRY(gf_graph_flush_internal(g));
  *out = (float)(g->h_stats->allocated_elems * 20ull);
  return GF_OK;
}
GF_EXPORT int gf_graph_metadata_memory_usage(gf_graph *g, float *out) {  // dynamic_graph.cu:372-380
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  GF_TRY(gf_graph_flush_internal(g));
  float sum = 0;
  sum += 64 * g->h_stats->num_blocks;  // sizeof(TemporalBlock) == 64 (common.h:35-48)
  sum += 8 * g->table_len();
  *out = sum;
  return GF_OK;
}
GF_EXPORT int gf_graph_device_bytes(gf_graph *g, uint64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  *out = g->arena_total + g->table_cap * (sizeof(NodeEntry) + 2) + g->eid_cap * 4 + g->s_in.cap + g->s_sort.cap +
         g->s_seg.cap + g->s_misc.cap + g->s_ctl.cap + g->s_pre.cap + g->s_book.cap + g->log_cap * sizeof(FreeRec) + 2 * g->sorted_cap * 8;
  return GF_OK;
}

GF_EXPORT int gf_graph_memory_breakdown(gf_graph *g, uint64_t *out) {
  if (!g || !out) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  GF_TRY(set_device(g));
  GF_TRY(settle(g));
  GF_TRY(pull_stats(g, 0));
  const ArenaState &ar = g->h_stats->arena;
  uint64_t bump_used = g->arena_total;
  for (unsigned k = 0; k < ar.num_regions; k++) bump_used -= ar.regions[k].end - ar.regions[k].cur;
  out[0] = g->arena_total;                                    // bytes of all chunks
  out[1] = bump_used;                                         // ... handed out by the bump pointers so far
  out[2] = (uint64_t)ar.free_units * kUnit;                   // ... of which sitting in the free lists / log
  out[3] = g->table_cap * (sizeof(NodeEntry) + 2);            // vertex table + flags
  out[4] = g->eid_cap * 4;                                    // edge-id reference counts
  out[5] = g->s_in.cap + g->s_sort.cap + g->s_seg.cap + g->s_misc.cap + g->s_ctl.cap + g->s_pre.cap + g->s_book.cap;  // per-call scratch
  out[6] = g->log_cap * sizeof(FreeRec) + 2 * g->sorted_cap * 8;  // allocator book-keeping
  out[7] = (uint64_t)ar.log_cnt + ar.sorted_cnt;              // free blocks on record
  return GF_OK;
}

GF_EXPORT int gf_graph_out_degree(gf_graph *g, const int64_t *ids, uint64_t n, uint64_t *out) {
  if (!g || (n && (!ids || !out))) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  if (!n) return GF_OK;
  GF_TRY(set_device(g));
  GF_TRY(settle(g));
  cudaStream_t st = 0;
  GF_TRY(g->s_misc.reserve(n * 16, st));
  int64_t *d_ids = g->s_misc.as<int64_t>();
  uint64_t *d_out = reinterpret_cast<uint64_t *>(d_ids + n);
  GF_CUDA(cudaMemcpyAsync(d_ids, ids, n * 8, cudaMemcpyHostToDevice, st));
  gf::launch(out_degree_kernel, cdiv(n, kThreads), kThreads, 0, st, g->d_table, g->table_len(), d_ids, n, d_out);
  GF_CUDA(cudaMemcpyAsync(out, d_out, n * 8, cudaMemcpyDeviceToHost, st));
  GF_CUDA(cudaStreamSynchronize(st));
  return GF_OK;
}

GF_EXPORT int gf_graph_nodes(gf_graph *g, int64_t *out, uint64_t cap, uint64_t *count) {
  if (!g || !count) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  return flags_to_list(g, g->d_is_node, g->table_len(), out, cap, count);
}
GF_EXPORT int gf_graph_src_nodes(gf_graph *g, int64_t *out, uint64_t cap, uint64_t *count) {
  if (!g || !count) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  return flags_to_list(g, g->d_is_src, g->table_len(), out, cap, count);
}
GF_EXPORT int gf_graph_edges(gf_graph *g, int64_t *out, uint64_t cap, uint64_t *count) {
  if (!g || !count) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  GF_TRY(set_device(g));
  std::vector<uint32_t> h(g->eid_cap);
  GF_TRY(settle(g));
  if (g->eid_cap) GF_CUDA(cudaMemcpy(h.data(), g->d_eid_ref, g->eid_cap * 4, cudaMemcpyDeviceToHost));
  uint64_t k = 0;
  for (size_t i = 0; i < h.size(); i++)
    if (h[i]) {
      if (out && k < cap) out[k] = (int64_t)i + g->eid_base;
      k++;
    }
  *count = k;
  if (out && cap < k) GF_FAIL(GF_ECAPACITY, "output buffer too small");
  return GF_OK;
}

GF_EXPORT int gf_graph_get_temporal_neighbors(gf_graph *g, int64_t vertex, int64_t *dst, float *ts, int64_t *eid,
                                              uint64_t cap, uint64_t *count) {
  if (!g || !count) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  NodeEntry ent;
  std::vector<BlockDesc> descs;
  GF_TRY(read_entry(g, vertex, &ent, &descs));
  uint64_t total = 0;
  for (auto &d : descs) total += d.size;
  *count = total;
  if (!dst || !ts || !eid) return GF_OK;
  if (cap < total) GF_FAIL(GF_ECAPACITY, "output buffer too small");
  uint64_t k = 0;
  std::vector<int64_t> hp;
  std::vector<float> ht;
  for (size_t b = descs.size(); b-- > 0;) {  // newest block first, each block reversed (dynamic_graph.cu:305-333)
    const BlockDesc &d = descs[b];
    hp.resize(2 * (size_t)d.size);
    ht.resize(d.size);
    if (!d.size) continue;
    GF_CUDA(cudaMemcpy(ht.data(), (const void *)d.payload, d.size * 4ull, cudaMemcpyDeviceToHost));
    GF_CUDA(cudaMemcpy(hp.data(), (const void *)(d.payload + payload_ts_bytes(d.capacity)), d.size * 16ull, cudaMemcpyDeviceToHost));
    for (uint32_t i = d.size; i-- > 0;) {
      unpack_rec_host(&hp[2 * (size_t)i], &dst[k], &eid[k]);
      ts[k] = ht[i];
      k++;
    }
  }
  return GF_OK;
}

GF_EXPORT int gf_graph_block_shapes(gf_graph *g, int64_t vertex, uint64_t *sizes, uint64_t *caps, float *start_ts,
                                    float *end_ts, uint64_t cap, uint64_t *count) {
  if (!g || !count) GF_FAIL(GF_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(g->mu);
  GF_TRY(flush_pending(g));
  NodeEntry ent;
  std::vector<BlockDesc> descs;
  GF_TRY(read_entry(g, vertex, &ent, &descs));
  *count = descs.size();
  if (!sizes) return GF_OK;
  if (cap < descs.size()) GF_FAIL(GF_ECAPACITY, "output buffer too small");
  for (size_t i = 0; i < descs.size(); i++) {
    sizes[i] = descs[i].size;
    caps[i] = g->cfg.insertion_policy == GF_INSERTION_REPLACE
                  ? replace_logical_cap(descs[i].size, (uint32_t)g->cfg.minimum_block_size) : descs[i].capacity;
    start_ts[i] = descs[i].start_ts;
    end_ts[i] = descs[i].end_ts;
  }
  return GF_OK;
}

GF_EXPORT int gf_graph_set_profiling(gf_graph *g, int on) {
  if (!g) GF_FAIL(GF_EINVAL, "null graph");
  g->prof.on = on != 0;
  return GF_OK;
}
GF_EXPORT int gf_graph_get_profile(gf_graph *g, double *ms, uint64_t *count, int reset) {
  if (!g || !ms || !count) GF_FAIL(GF_EINVAL, "null argument");
  g->prof.collect();
  for (int i = 0; i < GF_GRAPH_PHASES; i++) { ms[i] = g->prof.ms[i]; count[i] = g->prof.count[i]; }
  if (reset) g->prof.reset();
  return GF_OK;
}
GF_EXPORT uint64_t gf_debug_launch_count(void) { return gf::g_launch_count.load(); }

GF_EXPORT int gf_l2_fetch_granularity(int device, uint64_t bytes, uint64_t *current) {
  int ndev = 0, cur = 0;
  GF_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) GF_FAIL(GF_EINVAL, "device %d out of range (%d devices)", device, ndev);
  if (bytes != 0 && bytes != 32 && bytes != 64 && bytes != 128) GF_FAIL(GF_EINVAL, "L2 fetch granularity must be 32, 64 or 128");
  GF_CUDA(cudaGetDevice(&cur));
  GF_CUDA(cudaSetDevice(device));
  if (bytes) GF_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
  size_t v = 0;
  GF_CUDA(cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity));
  if (current) *current = v;
  GF_CUDA(cudaSetDevice(cur));
  return GF_OK;
}
