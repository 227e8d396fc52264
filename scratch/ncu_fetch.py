"""Launch list of one TGAT batch (sample + fetch_feature with an LRU cache) in steady state:
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python scratch/ncu_fetch.py"""
import os, sys
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
dn, dt = torch.from_numpy(nodes).to(dev), torch.from_numpy(rts).to(dev)
strat, fan = (sys.argv[1], [int(x) for x in sys.argv[2].split(",")]) if len(sys.argv) > 2 else ("uniform", [10, 10])
smp = TemporalSampler(g, fan, strat)
efeat = torch.randn(n, 172, device=dev)
cache = LRUCache(0.2, 0.2, stream["num_nodes"], n, dev, None, efeat, 0, 172)
cache.init_cache()
def batch(b):
    lo, hi = int(offs[b]), int(offs[b + 1])
    cache.fetch_feature(smp.sample(dn[lo:hi], dt[lo:hi]))
for b in range(0, 600):
    batch(b)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for b in range(600, 603):
    batch(b)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
