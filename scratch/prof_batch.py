"""Per-batch path, decomposed: raw C-ABI call vs Python glue vs GPU time (events), sample / gather / policy update."""
import ctypes as C, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from gnnflow_b200 import DynamicGraph, TemporalSampler, _lib
from gnnflow_b200._lib import SamplingResultC, GF_PTR_HOST, GF_PTR_DEVICE
from gnnflow_b200.cache import LRUCache
from gnnflow_b200.synth import synth, tgn_batches
dev = torch.device("cuda", 0)
L = _lib.lib()
stream = synth("REDDIT", seed=42)
nodes, rts, offs = tgn_batches(stream, B.BATCH, seed=7)
n = len(stream["src"])
g = DynamicGraph(**B.graph_config(stream))
for lo in range(0, n, B.INGEST_BATCH):
    sl = slice(lo, lo + B.INGEST_BATCH)
    g.add_edges(stream["src"][sl], stream["dst"][sl], stream["ts"][sl], stream["eid"][sl])
res = {}
NB = 300
B0 = 600
def wall(fn, reps=NB, sync_each=False):
    for b in range(B0, B0 + 30): fn(b)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for b in range(B0, B0 + reps): fn(b)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6
def gpu(fn, reps=NB):
    """GPU-side time per call when the calls are queued back to back (no host sync inside fn)"""
    for b in range(B0, B0 + 30): fn(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2000000)
    e0.record()
    for b in range(B0, B0 + reps): fn(b)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
dn, dt = torch.from_numpy(nodes).to(dev), torch.from_numpy(rts).to(dev)
sl_n = [dn[int(offs[b]):int(offs[b + 1])] for b in range(len(offs) - 1)]
sl_t = [dt[int(offs[b]):int(offs[b + 1])] for b in range(len(offs) - 1)]
h_n = [nodes[int(offs[b]):int(offs[b + 1])] for b in range(len(offs) - 1)]
h_t = [rts[int(offs[b]):int(offs[b + 1])] for b in range(len(offs) - 1)]
for strat, fan in (("recent", [10]), ("uniform", [10, 10])):
    tag = "%s%s" % (strat, fan)
    smp = TemporalSampler(g, fan, strat)
    r = {}
    r["sample_numpy_us"] = wall(lambda b: smp.sample_numpy(h_n[b], h_t[b]))
    r["sample_blocks_us"] = wall(lambda b: smp.sample(sl_n[b], sl_t[b]))
    r["sample_results_us"] = wall(lambda b: smp._sample_results(sl_n[b], sl_t[b]))
    # raw C call, host in / pinned host out, args prebuilt
    smp.sample_numpy(h_n[B0], h_t[B0])
    _, bufs, arr = smp._pinned
    st = C.c_void_p(torch.cuda.current_stream(0).cuda_stream)
    def raw_host(b):
        L.gf_sampler_sample(smp._h, h_n[b].ctypes.data, h_t[b].ctypes.data, h_n[b].shape[0], arr, GF_PTR_HOST, GF_PTR_HOST, st)
    r["raw_c_call_host_us"] = wall(raw_host)
    # raw C call, device in / device out
    pool_d, _offs, arr_d = smp._alloc_steps([(2000, fan[0])] + ([(2000 * 11, fan[1])] if len(fan) > 1 else []))
    smp._plan = None  # the sampler builds its own array for later calls
    ptrs = [(sl_n[b].data_ptr(), sl_t[b].data_ptr(), sl_n[b].shape[0]) for b in range(len(sl_n))]
    def raw_dev(b):
        p = ptrs[b]
        L.gf_sampler_sample(smp._h, p[0], p[1], p[2], arr_d, GF_PTR_DEVICE, GF_PTR_DEVICE, st)
    r["raw_c_call_device_us"] = wall(raw_dev)
    smp.set_profiling(True); smp.get_profile(True)
    for b in range(B0, B0 + NB): raw_dev(b)
    pr = smp.get_profile(True); smp.set_profiling(False)
    r["kernel_us_per_launch"] = pr["emit"][0] / max(1, pr["emit"][1]) * 1e3
    r["launches_per_call"] = pr["emit"][1] / NB
    res[tag] = r
    # cache legs on this sampler's blocks
    efeat = torch.randn(n, 172, device=dev)
    cache = LRUCache(0.2, 0.2, stream["num_nodes"], n, dev, None, efeat, 0, 172)
    cache.init_cache()
    for b in range(0, B0):  # warm the cache state to steady state
        cache.fetch_feature(smp.sample(sl_n[b], sl_t[b]))
    torch.cuda.synchronize()
    mf = [smp.sample(sl_n[b], sl_t[b]) for b in range(B0, B0 + NB + 40)]
    ids = [[blk.edata['ID'] for lay in m for blk in lay] for m in mf]
    torch.cuda.synchronize()
    c = {}
    c["edges_per_batch"] = float(np.mean([sum(int(x.shape[0]) for x in i) for i in ids]))
    def gather_only(b):
        for x in ids[b - B0]:
            cache._gather("edge", x)
    c["gather_gpu_us"] = gpu(gather_only); c["gather_wall_us"] = wall(gather_only)
    hits = {}
    def gather_keep(b):
        hits[b] = [cache._gather("edge", x) for x in ids[b - B0]]
    for b in range(B0, B0 + NB + 30): gather_keep(b)
    torch.cuda.synchronize()
    c["hit_rate"] = float(np.mean([float(h[3].item()) / max(1, h[0].shape[0]) for b in hits for h in hits[b]]))
    def update_only(b):
        for (i, f, h, nh) in hits[b]:
            cache.update_edge_cache(i, h)
    l0 = L.gf_debug_launch_count()
    c["update_gpu_us"] = gpu(update_only)
    c["launches_per_update_call"] = (L.gf_debug_launch_count() - l0) / (NB + 30) / len(hits[B0])
    c["update_wall_us"] = wall(update_only)
    l0 = L.gf_debug_launch_count()
    for b in range(B0, B0 + 10): cache.fetch_feature(mf[b - B0])
    c["launches_per_fetch_block"] = (L.gf_debug_launch_count() - l0) / 10 / len(hits[B0])
    def fetch(b): cache.fetch_feature(mf[b - B0])
    c["fetch_gpu_us"] = gpu(fetch); c["fetch_wall_us"] = wall(fetch)
    def both(b): cache.fetch_feature(smp.sample(sl_n[b], sl_t[b]))
    c["sample+fetch_wall_us"] = wall(both)
    res[tag + "_cache"] = c
    del cache, efeat, mf, ids, hits
print(json.dumps(res, indent=1))
