"""add_edges at the reference's 100 000-edge batches, device-resident arrays: sync (reference-shaped) and queued."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_configs as BC
from gnnflow_b200 import DynamicGraph
dev = torch.device("cuda", 0)
res = {"lib": os.environ.get("GNNFLOW_B200_LIB", "default")}
SHAPES = (("REDDIT", 1), ("GDELT-16.7K", 0.05), ("GDELT-16.7M", 0.05))
if os.environ.get("GF_SHAPE"):
    SHAPES = tuple(x for x in SHAPES if x[0] == os.environ["GF_SHAPE"])
for shape, scale in SHAPES:
    if shape == "REDDIT":
        from gnnflow_b200.synth import synth
        s = synth(shape)
        st = {k: torch.from_numpy(s[k]).to(dev) for k in ("src", "dst", "ts", "eid")}
        st.update(n=len(s["src"]), minimum_block_size=s["minimum_block_size"])
    else:
        st = BC.synth_gpu(shape, scale, dev)
    n, bs = st["n"], int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    g = DynamicGraph(initial_pool_size=256 << 20, maximum_pool_size=100 << 30, mem_resource_type="cuda",
                     minimum_block_size=st["minimum_block_size"], blocks_to_preallocate=1024, insertion_policy="insert")
    for mode in ("sync", "async"):
        def run():
            g.clear()
            for lo in range(0, n, bs):
                sl = slice(lo, min(n, lo + bs))
                (g.add_edges if mode == "sync" else g.add_edges_async)(st["src"][sl], st["dst"][sl], st["ts"][sl], st["eid"][sl])
            g.flush()
        run(); run(); torch.cuda.synchronize()
        if os.environ.get("GF_NCU_RANGE"):  # ncu --profile-from-start off: one run of the synchronous mode
            if mode == "sync":
                torch.cuda.profiler.start(); run(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
            continue
        best = 1e30
        for _ in range(5):
            t0 = time.perf_counter(); run(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        nb = (n + bs - 1) // bs
        res["%s_%s" % (shape, mode)] = {"us_per_batch": round(best * 1e6 / nb, 2), "Gedges_per_s": round(n / best / 1e9, 3)}
    del g
print(json.dumps(res))
