"""Ingest throughput through the reference-shaped call (DynamicGraph.add_edges, one host sync per batch) and the
queued variant, device-resident inputs, batch size swept; phase split from CUDA events inside the library."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_configs as BC  # noqa: E402


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "GDELT-16.7K"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
    dev = torch.device("cuda", 0)
    from gnnflow_b200 import DynamicGraph
    if shape in ("REDDIT", "WIKI"):
        from gnnflow_b200.synth import synth
        s = synth(shape)
        st = {k: torch.from_numpy(s[k]).to(dev) for k in ("src", "dst", "ts", "eid")}
        st.update(n=len(s["src"]), minimum_block_size=s["minimum_block_size"], num_nodes=s["num_nodes"])
    else:
        st = BC.synth_gpu(shape, scale, dev)
    n = st["n"]
    peak, _ = BC.peak_hbm()
    cfg = dict(initial_pool_size=256 << 20, maximum_pool_size=150 << 30, mem_resource_type="cuda",
               minimum_block_size=st["minimum_block_size"], blocks_to_preallocate=1024, insertion_policy="insert")
    g = DynamicGraph(**cfg)
    rows = []
    for bs in (1000, 10000, 100000, 1000000, 4000000, 16000000):
        if bs > n:
            break
        m = min(n, max(bs * 8, min(n, 32_000_000)))
        for mode in ("sync", "async"):
            def run():
                g.clear()
                for lo in range(0, m, bs):
                    sl = slice(lo, min(m, lo + bs))
                    if mode == "sync":
                        g.add_edges(st["src"][sl], st["dst"][sl], st["ts"][sl], st["eid"][sl])
                    else:
                        g.add_edges_async(st["src"][sl], st["dst"][sl], st["ts"][sl], st["eid"][sl])
                g.flush()
            run(); run()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record(); run(); e1.record(); torch.cuda.synchronize()
                wall = (time.perf_counter() - t0) * 1e3
                best = min(best, max(e0.elapsed_time(e1), 0.0))
            ms = best
            g.set_profiling(True); g.get_profile(True); run(); torch.cuda.synchronize()
            ph = {k: round(v[0] / max(1, v[1]) * 1e3, 2) for k, v in g.get_profile(True).items()}
            g.set_profiling(False)
            nb = (m + bs - 1) // bs
            rows.append({"batch_edges": bs, "mode": mode, "edges": m, "edges_per_s": m / (ms * 1e-3), "us_per_batch": ms * 1e3 / nb,
                         "wall_us_per_batch": wall * 1e3 / nb, "algorithmic_GBps": m * 48 / (ms * 1e-3) / 1e9,
                         "frac": m * 48 / (ms * 1e-3) / 1e9 / peak, "phase_us_per_batch": ph,
                         "device_bytes": g.get_device_memory_usage(), "payload_bytes": g.get_graph_memory_usage()})
            print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
    print(json.dumps({"shape": shape, "scale": scale, "num_nodes": st["num_nodes"], "peak": peak, "sweep": rows}))


if __name__ == "__main__":
    main()
