"""Sampling-kernel timing for the experiment matrix: REDDIT headline launch + GDELT hbm-bound launches."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
import bench_configs as BC
from gnnflow_b200 import DynamicGraph, TemporalSampler
from gnnflow_b200.synth import synth, tgn_batches

dev = torch.device("cuda", 0)
out = {"lib": os.environ.get("GNNFLOW_B200_LIB", "default"), "occ": os.environ.get("GNNFLOW_B200_OCC", "4")}
stream = synth("REDDIT")
nodes, rts, offs = tgn_batches(stream, 600, 7)
g = DynamicGraph(**B.graph_config(stream))
for lo in range(0, len(stream["src"]), 100000):
    sl = slice(lo, lo + 100000)
    g.add_edges(stream["src"][sl], stream["dst"][sl], stream["ts"][sl], stream["eid"][sl])
dn, dt, do = [torch.from_numpy(x).to(dev) for x in (nodes, rts, offs)]
for strat in ("recent", "uniform"):
    smp = TemporalSampler(g, [10], strat)
    o = smp.sample_layer_batched(dn, dt, do)
    for _ in range(5):
        smp.sample_layer_batched(dn, dt, do, out=o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        smp.sample_layer_batched(dn, dt, do, out=o)
    e1.record(); torch.cuda.synchronize()
    out["reddit_" + strat + "_ms"] = e0.elapsed_time(e1) / 20
del g
for shape in ("GDELT-16.7K", "GDELT-16.7M"):
    r = BC.hbm_bound_leg(dev, 0, shape, float(sys.argv[1]) if len(sys.argv) > 1 else 0.25, steps=5, warmup=3)
    for l in r["launches"]:
        if l["layer"] != "chain":
            out["%s_%s_l%s" % (shape, l["strategy"], l["layer"])] = (round(l["ms_per_launch"], 4), round(l["frac"], 3))
print(json.dumps(out))
