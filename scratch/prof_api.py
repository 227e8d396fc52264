"""Where does the per-batch time of the public API go?  cProfile over sample() and fetch_feature()."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from gnnflow_b200 import DynamicGraph, TemporalSampler
from gnnflow_b200.cache import LRUCache
from gnnflow_b200.synth import synth, tgn_batches
dev = torch.device("cuda", 0)
stream = synth("REDDIT", seed=42)
nodes, rts, offs = tgn_batches(stream, B.BATCH, seed=7)
n = len(stream["src"])
g = DynamicGraph(**B.graph_config(stream))
for lo in range(0, n, B.INGEST_BATCH):
    sl = slice(lo, lo + B.INGEST_BATCH)
    g.add_edges(stream["src"][sl], stream["dst"][sl], stream["ts"][sl], stream["eid"][sl])
efeat = torch.randn(n, 172, device=dev)
cache = LRUCache(0.2, 0.2, stream["num_nodes"], n, dev, None, efeat, 0, 172)
cache.init_cache()
dn, dt = torch.from_numpy(nodes).to(dev), torch.from_numpy(rts).to(dev)
for strat, fan in (("recent", [10]), ("uniform", [10, 10])):
    smp = TemporalSampler(g, fan, strat)
    def run(nb, fetch):
        for b in range(600, 600 + nb):
            lo, hi = int(offs[b]), int(offs[b + 1])
            m = smp.sample(dn[lo:hi], dt[lo:hi])
            if fetch:
                cache.fetch_feature(m)
        torch.cuda.synchronize()
    for fetch in (False, True):
        run(50, fetch)
        t0 = time.perf_counter(); run(300, fetch); t1 = time.perf_counter()
        print("== %s %s fetch=%s: %.1f us/batch" % (strat, fan, fetch, (t1 - t0) / 300 * 1e6))
        pr = cProfile.Profile(); pr.enable(); run(300, fetch); pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(12)
