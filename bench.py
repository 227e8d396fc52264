#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric: sampled neighbors/s + edges inserted/s).

Workload (BASELINE.json configs[1]): REDDIT-shaped synthetic stream (10,984 nodes, 672,447 edges, directed,
minimum_block_size 62), TGN 1-layer recent sampling fanout 10 in batches of 600 edges (roots = src || dst || random
negatives = 1,800 targets per batch) + edge insert in 100,000-edge add_edges batches.

One step = one replay of the whole stream:
  ingest phase : empty the graph, add_edges the 672,447 edges in 7 batches            -> edges inserted / s
  sample phase : sample all 1,121 batches of 600 (2,017,341 targets)                  -> sampled neighbors / s
`value` (device-resident inputs) issues the 1,121 batches as ONE multi-batch launch (gf_sampler_sample_layer_batched);
`e2e` is the same multi-batch launch through the public host-array API (TemporalSampler.sample_layer_batched_numpy ->
gf_sampler_sample_layer_batched with GF_PTR_HOST): pinned host arrays in, pinned host arrays out, H2D + D2H inside the
timed region; `e2e.per_batch` is the synchronous per-batch API (sample_numpy once per batch of 600), which is how the
reference arm is called.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FANOUT = 10
BATCH = 600
INGEST_BATCH = 100000
METRIC = "sampled_neighbors_per_s"
UNIT = "neighbors/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-real"])
    p.add_argument("--dataset", default="REDDIT")
    p.add_argument("--e2e-steps", type=int, default=10)
    p.add_argument("--e2e-int64", action="store_true", help="e2e leg with int64 neighbour ids / rows (32 B per neighbour) instead of uint32 (24 B)")
    p.add_argument("--host-out-mode", type=int, default=0, help="0: auto; 1: device mirror + D2H; 2: kernel writes pinned host outputs in place")
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--ref-mem", default="cuda", choices=["cuda", "pinned"],
                   help="reference arm: where the reference keeps its graph (MemoryResourceType)")
    p.add_argument("--variant", type=int, default=3)
    p.add_argument("--no-hbm-bound", action="store_true", help="skip the GDELT-shaped HBM-bound leg (N = 1 only)")
    p.add_argument("--hbm-scale", type=float, default=1.0, help="scale of the GDELT-shaped stream of the HBM-bound leg")
    p.add_argument("--no-partitioned", action="store_true", help="skip the hash-partitioned GDELT leg (N > 1 only)")
    p.add_argument("--part-shape", default="GDELT-16.7M", choices=["GDELT-16.7M", "GDELT-16.7K"])
    p.add_argument("--part-scale", type=float, default=1.0)
    p.add_argument("--part-batches", type=int, default=64, help="root batches of 600 edges per rank and exchange step")
    p.add_argument("--no-per-batch-models", action="store_true", help="skip the per-batch TGAT / TGN sample + fetch_feature leg")
    p.add_argument("--launch-gap-us", type=float, default=400.0,
                   help="device-side spin queued before the timed sampling launch of every step (0: none): the host's "
                        "launch latency on an idle GPU is hidden behind it, as a training loop hides it behind the "
                        "previous kernel")
    p.add_argument("--parity-batches", type=int, default=0,
                   help="batches of the headline output compared with the CPU oracle (0: all at N = 1, 150 per rank at N > 1)")
    return p.parse_args()


def graph_config(stream):
    return dict(initial_pool_size=20 << 20, maximum_pool_size=1000 << 20, mem_resource_type="cuda",
                minimum_block_size=stream["minimum_block_size"], blocks_to_preallocate=1024,
                insertion_policy="insert")


def workload_config(stream, extra=None):
    c = {"workload": "{}-shaped synthetic ({} nodes, {} edges, directed), TGN 1-layer recent fanout {}, batch {} "
                     "(1,800 roots), add_edges in {}-edge batches; one step = one replay of the stream".format(
                         stream["name"], stream["num_nodes"], len(stream["src"]), FANOUT, BATCH, INGEST_BATCH),
         "dataset": stream["name"], "fanouts": [FANOUT], "strategy": "recent", "batch_size": BATCH,
         "ingest_batch": INGEST_BATCH, "minimum_block_size": stream["minimum_block_size"]}
    if extra:
        c.update(extra)
    return c


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def wait_rows(self, n, timeout):
        """block until the sampler has produced n rows in total (it needs a few hundred ms to start)"""
        t0 = time.time()
        while self.proc and len(self.rows) < n and time.time() - t0 < timeout:
            time.sleep(0.01)
        return len(self.rows)

    def stop(self, first=0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.rows = self.rows[first:]  # only what was sampled under the load
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 7:
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arms
def cpu_port_run(stream, nodes, rts, offs, seconds, steps=1, warmup=0):
    """The CPU oracle (C restatement of the reference algorithm, OpenMP over targets) on this box's host cores:
    ingest of the whole stream + per-batch sampling of a time-bounded prefix of the batches."""
    from oracle.oracle import OracleGraph, OracleSampler, num_threads
    res = []
    for it in range(warmup + steps):
        og = OracleGraph(**graph_config(stream))
        t0 = time.perf_counter()
        n = len(stream["src"])
        for lo in range(0, n, INGEST_BATCH):
            sl = slice(lo, lo + INGEST_BATCH)
            og.add_edges(stream["src"][sl], stream["dst"][sl], stream["ts"][sl], stream["eid"][sl])
        t_ing = time.perf_counter() - t0
        smp = OracleSampler(og, [FANOUT], "recent")
        nb = len(offs) - 1
        S = 0
        done = 0
        t0 = time.perf_counter()
        for b in range(nb):
            r = smp.sample_layer(nodes[offs[b]:offs[b + 1]], rts[offs[b]:offs[b + 1]], 0, 0)
            S += len(r["eids"])
            done += 1
            if time.perf_counter() - t0 > seconds:
                break
        t_smp = time.perf_counter() - t0
        if it >= warmup:
            res.append((S / t_smp, n / t_ing, done, t_smp, t_ing))
    v = float(np.median([r[0] for r in res]))
    return {"value": v, "unit": UNIT, "cores": num_threads(), "kind": "port",
            "sample": "{} of {} batches of {} (per-batch calls) after ingesting the whole stream".format(
                res[-1][2], len(offs) - 1, BATCH),
            "ingest_edges_per_s": float(np.median([r[1] for r in res])),
            "ms_per_step": float(np.median([r[3] for r in res])) * 1e3}


def parity_check(stream, nodes, rts, offs, out, max_batches):
    """Outside the timed region: the GPU output of the headline launch (device arrays `out`) against the CPU oracle on
    the same inputs, batch by batch, bit for bit.  Returns (batches compared, neighbours compared)."""
    from oracle.oracle import OracleGraph, OracleSampler
    og = OracleGraph(**graph_config(stream))
    n = len(stream["src"])
    for lo in range(0, n, INGEST_BATCH):
        sl = slice(lo, lo + INGEST_BATCH)
        og.add_edges(stream["src"][sl], stream["dst"][sl], stream["ts"][sl], stream["eid"][sl])
    smp = OracleSampler(og, [FANOUT], "recent")
    nb = len(offs) - 1
    pick = range(nb) if max_batches <= 0 or max_batches >= nb else sorted(set(np.linspace(0, nb - 1, max_batches).astype(int).tolist()))
    eo = out["edge_offsets"].cpu().numpy()
    res = {k: out[k].cpu().numpy() for k in ("nbr", "ts", "dt", "eid", "row")}
    S = 0
    for b in pick:
        o = smp.sample_layer(nodes[offs[b]:offs[b + 1]], rts[offs[b]:offs[b + 1]], 0, 0)
        T = int(offs[b + 1] - offs[b])
        sl = slice(int(eo[b]), int(eo[b + 1]))
        ok = (sl.stop - sl.start == len(o["eids"]) and np.array_equal(res["nbr"][sl], o["all_nodes"][T:])
              and np.array_equal(res["eid"][sl], o["eids"]) and np.array_equal(res["row"][sl], o["row"])
              and np.array_equal(res["ts"][sl].view(np.int32), o["all_timestamps"][T:].view(np.int32))
              and np.array_equal(res["dt"][sl].view(np.int32), o["delta_timestamps"].view(np.int32)))
        if not ok:
            raise AssertionError("GPU output of batch {} differs from the CPU oracle".format(b))
        S += len(o["eids"])
    return len(pick), S


def reference_real(args, stream, nodes, rts, offs):
    """The UNMODIFIED reference extension (oracle/_ref/libgnnflow*.so, built from /root/reference by
    oracle/build_ref.sh) through its own pybind API: _DynamicGraph.add_edges + _TemporalSampler.sample per batch.
    Its sampler runs its own CUDA kernels and post-processes on the host (4 OpenMP threads, hard-wired)."""
    import torch  # noqa: F401  (libgnnflow links libtorch)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import libgnnflow as ref
    cfg = graph_config(stream)
    vals, ings, mss = [], [], []
    nb = len(offs) - 1
    for it in range(args.warmup + args.steps):
        g = ref._DynamicGraph(cfg["initial_pool_size"], cfg["maximum_pool_size"],
                              ref.MemoryResourceType.PINNED if args.ref_mem == "pinned" else ref.MemoryResourceType.CUDA,
                              cfg["minimum_block_size"], cfg["blocks_to_preallocate"], ref.InsertionPolicy.INSERT, 0,
                              True)
        n = len(stream["src"])
        t0 = time.perf_counter()
        for lo in range(0, n, INGEST_BATCH):
            sl = slice(lo, lo + INGEST_BATCH)
            g.add_edges(stream["src"][sl], stream["dst"][sl], stream["ts"][sl], stream["eid"][sl])
        t_ing = time.perf_counter() - t0
        s = ref._TemporalSampler(g, [FANOUT], ref.SamplingPolicy.RECENT, 1, 0.0, False, 1234)
        S = 0
        t0 = time.perf_counter()
        for b in range(nb):
            r = s.sample(nodes[offs[b]:offs[b + 1]], rts[offs[b]:offs[b + 1]])
            S += len(r[0][0].eids())
        t_smp = time.perf_counter() - t0
        del s, g
        if it >= args.warmup:
            vals.append(S / t_smp)
            ings.append(n / t_ing)
            mss.append(t_smp * 1e3)
    v = float(np.median(vals))
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.median(mss)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64+f32", "data": "synthetic",
            "config": workload_config(stream, {"reference": "unmodified libgnnflow (its CUDA kernels on GPU 0 + host "
                                                            "post-processing), graph in {} memory".format(
                                                                "pinned host" if args.ref_mem == "pinned" else "device")}),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": 4, "kind": "reference",
                             "sample": "all {} batches per step; host cores available: {}".format(nb, cores)},
            "ingest": {"value": float(np.median(ings)), "unit": "edges/s"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def reference_arm(args, stream, nodes, rts, offs):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    has_ref = os.path.isdir(os.path.join(ROOT, "oracle", "_ref")) and any(
        f.startswith("libgnnflow") for f in os.listdir(os.path.join(ROOT, "oracle", "_ref")))
    if has_ref:
        env = dict(os.environ)
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        try:
            def spawn_real(mem, steps, warmup, gpu):
                e = dict(env)
                e["CUDA_VISIBLE_DEVICES"] = str(gpu)  # the reference always uses device 0 (utils.cu:65-71, api.cc:41-47)
                return subprocess.Popen([sys.executable, os.path.abspath(__file__), "--impl", "reference-real", "--steps",
                                         str(steps), "--warmup", str(warmup), "--dataset", args.dataset, "--ref-mem", mem],
                                        stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=e)

            def collect(proc, timeout, tag):
                try:
                    so, se = proc.communicate(timeout=timeout)
                except subprocess.TimeoutExpired:
                    proc.kill()
                    so, se = proc.communicate()
                for ln in so.splitlines()[::-1]:
                    if ln.startswith("{") and '"impl": "reference"' in ln:
                        return json.loads(ln)
                sys.stderr.write("reference-real ({}) failed (rc={}): {}\n".format(tag, proc.returncode, se[-2000:]))
                return None

            # like for like: ONE reference process per GPU, all N running at the same time (each ingests the stream and
            # samples every batch, exactly what each of our ranks does); the aggregate is what N GPUs deliver
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            gpus = [int(x) for x in visible.split(",")] if visible else list(range(max(1, args.gpus)))
            gpus = gpus[:max(1, args.gpus)]
            procs = [spawn_real("cuda", args.steps, args.warmup, gpu) for gpu in gpus]
            lines = [collect(p_, 1500, "cuda, gpu {}".format(gpu)) for p_, gpu in zip(procs, gpus)]
            if all(ln is not None for ln in lines):
                line = lines[0]
                if len(lines) > 1:
                    line["value"] = float(sum(ln["value"] for ln in lines))
                    line["ms_per_step"] = float(max(ln["ms_per_step"] for ln in lines))
                    line["n_gpus"] = len(lines)
                    line["per_gpu_values"] = [ln["value"] for ln in lines]
                    line["ingest"] = {"value": float(min(ln["ingest"]["value"] for ln in lines)), "unit": "edges/s",
                                      "note": "per replica (slowest of the {} concurrent processes)".format(len(lines))}
                    line["cpu_baseline"]["value"] = line["value"]
                    line["cpu_baseline"]["sample"] += "; {} concurrent reference processes, one per GPU".format(len(lines))
                    line["e2e"]["value"] = line["value"]
                    line["config"]["reference"] += "; one process per GPU x {}".format(len(lines))
                # the reference's host-memory graph path (mem_resource_type = pinned: its kernels read the graph over
                # PCIe), in its own process -- the reference abort()s on any failed CHECK
                try:
                    hm = collect(spawn_real("pinned", 1, 1, gpus[0]), 300, "pinned")
                except Exception as e:  # noqa: BLE001
                    hm = None
                    sys.stderr.write("reference-real (pinned) failed: {}\n".format(e))
                line["host_memory_graph"] = None if hm is None else {
                    "value": hm["value"], "unit": UNIT, "ms_per_step": hm["ms_per_step"], "ingest": hm["ingest"],
                    "note": "same run with the reference's graph in pinned host memory (MemoryResourceType.PINNED), 1 GPU"}
                emit(line)
                return
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("reference-real failed: {}\n".format(e))
    # fallback: the CPU port of the reference algorithm
    steps = max(1, min(args.steps, 3))
    r = cpu_port_run(stream, nodes, rts, offs, seconds=20.0, steps=steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64+f32", "data": "synthetic",
            "config": workload_config(stream), "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "ingest": {"value": r["ingest_edges_per_s"], "unit": "edges/s"},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------- our arm
def cache_gather_leg(dev, peak):
    """The feature-cache gather (out[i] = cached ? buffer[map[id]] : features[id]) at the reference's edge
    feature width, 1 M rows per call, 20 % of the table cached; checked bit for bit against torch indexing."""
    import torch
    from gnnflow_b200._lib import check, lib
    L = lib()
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    out_rows = []
    for D, N, n in ((172, 2_000_000, 1 << 20), (768, 1_000_000, 1 << 20)):
        feats = torch.randn(N, D, device=dev, generator=gen)
        cap = N // 5
        cached = torch.randperm(N, device=dev, generator=gen)[:cap]
        flag = torch.zeros(N, dtype=torch.uint8, device=dev)
        flag[cached] = 1
        cmap = torch.full((N,), -1, dtype=torch.int64, device=dev)
        cmap[cached] = torch.arange(cap, device=dev)
        buf = feats[cached].contiguous()
        ids = torch.randint(0, N, (n,), device=dev, generator=gen)
        out = torch.empty(n, D, device=dev)
        hm = torch.empty(n, dtype=torch.uint8, device=dev)
        nh = torch.zeros(1, dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream().cuda_stream

        def call():
            check(L.gf_cache_gather(ids.data_ptr(), n, N, flag.data_ptr(), cmap.data_ptr(), buf.data_ptr(), feats.data_ptr(), D,
                                    out.data_ptr(), hm.data_ptr(), nh.data_ptr(), None, st))
        call()
        torch.cuda.synchronize()
        equal = bool(torch.equal(out, feats[ids])) and int(nh.item()) == int(flag[ids].sum().item())
        for _ in range(3):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        b = n * (17 + 8 * D)  # 8 id + 1 flag + 8 map + 4 D read + 4 D write per row
        out_rows.append({"dim": D, "rows": n, "table_rows": N, "cached_frac": 0.2, "ms": ms, "algorithmic_bytes": b,
                         "achieved_GBps": b / ms / 1e6, "frac": b / ms / 1e6 / peak, "equals_torch_indexing": equal})
        del feats, buf, out, cmap, flag, ids
        torch.cuda.empty_cache()
    return {"kernel": "cache_gather_kernel", "peak": peak, "unit": "GB/s", "launches": out_rows}


def tgat_per_batch_leg(g, stream, nodes, rts, offs, dev, batches=300, warm=400):
    """BASELINE config 3's inner loop as a training script runs it, one batch at a time through the public API:
    TemporalSampler([10, 10], uniform).sample(roots of 600 edges) -> LRUCache(0.2).fetch_feature(mfgs) (edge features,
    De = 172, policy update included), cache in steady state.  Wall clock per batch, device-resident roots; and the
    TGN shape (recent [10]) beside it."""
    import torch
    from gnnflow_b200 import TemporalSampler
    from gnnflow_b200.cache import LRUCache
    n = len(stream["src"])
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    efeat = torch.randn(n, 172, device=dev, generator=gen)
    dn, dt = torch.from_numpy(nodes).to(dev), torch.from_numpy(rts).to(dev)
    nb = len(offs) - 1
    warm = min(warm, max(0, nb - batches))
    sl = [(dn[int(offs[b]):int(offs[b + 1])], dt[int(offs[b]):int(offs[b + 1])]) for b in range(nb)]
    res = {}
    for name, strat, fan in (("tgat", "uniform", [10, 10]), ("tgn", "recent", [10])):
        smp = TemporalSampler(g, fan, strat)
        cache = LRUCache(0.2, 0.2, stream["num_nodes"], n, dev, None, efeat, 0, 172)
        cache.init_cache()
        for b in range(warm):
            cache.fetch_feature(smp.sample(*sl[b]))
        torch.cuda.synchronize()
        from gnnflow_b200._lib import lib as _lib_fn
        passes = []
        # three passes over the same window of batches, the fastest one is reported (all are listed): the leg follows the
        # per-batch e2e loop, which leaves the GPU mostly idle, and one run in four on a fresh box showed a first pass
        # 2.5 x slower than all others with the same launches and the same hit ratio (clocks still ramping up)
        for _ in range(3):
            t0 = time.perf_counter()
            edges = 0
            marks = []
            launches0 = _lib_fn().gf_debug_launch_count()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for b in range(warm, warm + batches):
                mfgs = cache.fetch_feature(smp.sample(*sl[b]))
                edges += sum(blk.num_edges() for lay in mfgs for blk in lay)
                marks.append(time.perf_counter())
            ev1.record()
            torch.cuda.synchronize()
            dt_pass = time.perf_counter() - t0
            per = np.diff(np.array([t0] + marks)) * 1e6  # host-side time per batch (the GPU may run behind)
            passes.append(dict(dt_s=dt_pass, launches_pb=(_lib_fn().gf_debug_launch_count() - launches0) / batches,
                               device_us=ev0.elapsed_time(ev1) * 1e3 / batches, per=per))
        best = min(passes, key=lambda q: q["dt_s"])
        dt_s, launches_pb, per = best["dt_s"], best["launches_pb"], best["per"]
        t1 = time.perf_counter()
        for b in range(warm, warm + batches):
            smp.sample(*sl[b])
        torch.cuda.synchronize()
        smp_s = time.perf_counter() - t1
        # values: what fetch_feature returned equals the feature table's rows (the last batch, bit for bit)
        for lay in mfgs:
            for blk in lay:
                if 'f' in blk.edata:
                    assert torch.equal(blk.edata['f'], efeat[blk.edata['ID']]), "fetch_feature != edge_feats[ID]"
        res[name] = {"us_per_batch": dt_s / batches * 1e6, "sample_us_per_batch": smp_s / batches * 1e6,
                     "passes_us_per_batch": [q["dt_s"] / batches * 1e6 for q in passes],
                     "launches_per_batch": launches_pb, "device_us_per_batch": best["device_us"],
                     "host_us_p50": float(np.median(per)), "host_us_max": float(per.max()),
                     "host_us_spikes": [[int(i), round(float(per[i]), 1)] for i in np.argsort(-per)[:4]],
                     "fanouts": fan, "strategy": strat, "batches": batches, "edges_per_batch": edges / batches,
                     "edge_hit_ratio_last_batch": float(cache.cache_edge_ratio),
                     "api": "TemporalSampler.sample(cuda roots) + LRUCache(0.2).fetch_feature(mfgs), De = 172, update_cache=True"}
        del cache, smp
    return res


def bind_to_gpu_numa(index):
    """Run this rank on the host cores next to its GPU (NVML's CPU affinity of the device), so that the pinned host
    buffers of the e2e leg are allocated on that socket and N ranks do not all write into one socket's memory.
    Returns what was done, for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {w * 64 + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return "bound to {} host cores next to GPU {}".format(len(cpus), index)
        return "not bound (GPU {} is next to all {} visible cores)".format(index, len(os.sched_getaffinity(0)))
    except Exception as e:  # noqa: BLE001
        return "not bound ({})".format(type(e).__name__)


def ours(args, stream, nodes, rts, offs):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa(local)  # before any pinned allocation: first touch decides where the e2e buffers live
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from gnnflow_b200 import DynamicGraph, TemporalSampler, _lib

    L = _lib.lib()
    n = len(stream["src"])
    d_src, d_dst = torch.from_numpy(stream["src"]).to(dev), torch.from_numpy(stream["dst"]).to(dev)
    d_ts, d_eid = torch.from_numpy(stream["ts"]).to(dev), torch.from_numpy(stream["eid"]).to(dev)
    d_nodes, d_rts, d_offs = torch.from_numpy(nodes).to(dev), torch.from_numpy(rts).to(dev), torch.from_numpy(offs).to(dev)
    T = len(nodes)
    nb = len(offs) - 1
    g = DynamicGraph(**graph_config(stream), device=local)
    smp = TemporalSampler(g, [FANOUT], "recent")
    smp.set_variant(args.variant)
    out = dict(nbr=torch.empty(T * FANOUT, dtype=torch.int64, device=dev),
               ts=torch.empty(T * FANOUT, dtype=torch.float32, device=dev),
               dt=torch.empty(T * FANOUT, dtype=torch.float32, device=dev),
               eid=torch.empty(T * FANOUT, dtype=torch.int64, device=dev),
               row=torch.empty(T * FANOUT, dtype=torch.int64, device=dev),
               edge_offsets=torch.empty(nb + 1, dtype=torch.int64, device=dev))
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def ingest_device(queued=False):
        g.clear()
        for lo in range(0, n, INGEST_BATCH):
            sl = slice(lo, lo + INGEST_BATCH)
            if not queued:  # the reference's call: returns when the batch is in the graph (one host sync per batch)
                g.add_edges(d_src[sl], d_dst[sl], d_ts[sl], d_eid[sl])
            else:  # queued: one host synchronisation per replay instead of one per batch
                g.add_edges_async(d_src[sl], d_dst[sl], d_ts[sl], d_eid[sl])
        g.flush()

    def sample_device():
        smp.sample_layer_batched(d_nodes, d_rts, d_offs, 0, 0, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        ingest_device()
        ingest_device(queued=True)
        sample_device()
    torch.cuda.synchronize()
    S = int(out["edge_offsets"][-1].item())
    # ---- parity gate (outside the timed region): the headline output equals the CPU oracle's, bit for bit
    pb = args.parity_batches if args.parity_batches > 0 else (0 if world == 1 else 150)
    parity_batches, parity_neighbors = parity_check(stream, nodes, rts, offs, out, pb)
    deg_pos = int((torch.from_numpy(g.out_degree(nodes[:200000]).astype(np.int64)) > 0).sum())
    frac_with_edges = deg_pos / 200000.0
    num_blocks = max(1, int(round(g.avg_linked_list_length() * g.num_vertices())))
    mean_block = g.num_edges() / num_blocks

    # ---- timed region: K steps, device-resident inputs
    # the sampling kernel is timed inside the timed region (roofline: CUDA events in the library, one pair per launch); the
    # per-phase breakdown of add_edges (ten event records per 100 000-edge batch) is taken in an untimed pass afterwards
    smp.set_profiling(True)
    smp.get_profile(True)
    launches0 = L.gf_debug_launch_count()
    clocks = ClockSampler(local)
    rows0 = 0
    if rank == 0:
        clocks.start()
        rows0 = clocks.wait_rows(1, 3.0)  # the sampler is up before the timed region starts
    e_ing, e_smp = [], []
    barrier()
    rows0 = len(clocks.rows)  # rows from here on are sampled under the load
    t_begin, t_end = ev(), ev()
    t_begin.record()
    # add_edges (the reference's synchronous call) leaves the GPU idle, so an event recorded right before the sampling
    # launch would also time the host's launch path (10-15 us alone on a box, 100+ us with 8 ranks sharing the host).  A
    # device-side spin is queued first: the launch is in the queue when the spin ends and b -> c is the GPU time of the
    # step's sampling launch.  The spin is outside both timed phases; step_ms_total includes it and says so.
    gap_cycles = int(args.launch_gap_us * 1e-6 * getattr(torch.cuda.get_device_properties(dev), "clock_rate", 1965000) * 1e3)
    for _ in range(args.steps):
        a, a1, b, c = ev(), ev(), ev(), ev()
        a.record()
        ingest_device()
        a1.record()
        if gap_cycles > 0:
            torch.cuda._sleep(gap_cycles)
        b.record()
        sample_device()
        c.record()
        e_ing.append((a, a1))
        e_smp.append((b, c))
    t_end.record()
    barrier()
    launches = L.gf_debug_launch_count() - launches0
    # the same ingest through the queued call (add_edges_async + one flush per replay), timed beside it
    qa, qb = ev(), ev()
    qa.record()
    for _ in range(args.steps):
        ingest_device(queued=True)
    qb.record()
    torch.cuda.synchronize()
    ing_q_ms = qa.elapsed_time(qb)
    total_ms = t_begin.elapsed_time(t_end)
    ing_ms = sum(a.elapsed_time(b) for a, b in e_ing)
    smp_ms = sum(a.elapsed_time(b) for a, b in e_smp)
    prof_s = smp.get_profile(True)
    g.set_profiling(True)
    g.get_profile(True)
    for _ in range(3):
        ingest_device()
    torch.cuda.synchronize()
    prof_g_timed = g.get_profile(True)
    g.set_profiling(False)
    # The timed region lasts a few tens of ms, nvidia-smi samples every 100 ms: rank 0 keeps the SAME steps running,
    # untimed, until the sampler has seen the load at least three times (clocks / throttle reasons under load).
    clock_load_ms = 0.0
    if rank == 0:
        t0 = time.time()
        while len(clocks.rows) - rows0 < 3 and time.time() - t0 < 3.0:
            ingest_device()
            sample_device()
            torch.cuda.synchronize()
        clock_load_ms = (time.time() - t0) * 1e3
    clk = clocks.stop(first=rows0) if rank == 0 else None
    if clk is not None:
        clk["sampled_over_ms"] = total_ms + clock_load_ms
        clk["note"] = "nvidia-smi -lms 100 while the timed steps (and, untimed, the same steps again) were running"
    barrier()
    prof_g = prof_g_timed
    smp.get_profile(True)  # drop what the untimed clock-sampling steps added
    smp.set_profiling(False)

    # ---- e2e: HOST buffers in and out through the public API, H2D + D2H inside the timed region.
    #   e2e.value        one TemporalSampler.sample_layer_batched_numpy call per step (the same multi-batch launch as
    #                    `value`, inputs copied from pinned host arrays, results written into pinned host arrays)
    #   e2e.per_batch    TemporalSampler.sample_numpy once per batch of 600, synchronous (how the reference arm is called)
    e2e_steps = max(0, min(args.e2e_steps, args.steps))
    hsrc, hdst, hts, heid = stream["src"], stream["dst"], stream["ts"], stream["eid"]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    p_nodes, p_rts, p_offs = pin(nodes), pin(rts), pin(offs.astype(np.uint64))
    # 32-bit neighbour ids and rows over PCIe (gf_sampler_sample_layer_batched_ids32: 24 B per neighbour instead of 32,
    # same values); --e2e-int64 times the 64-bit arrays of gf_sampler_sample_layer_batched instead
    ids32 = not args.e2e_int64
    out_bytes = 24 if ids32 else 32
    host_out = smp.alloc_batched_host_out(T, nb, 0, pinned=True, ids32=ids32)
    smp.set_host_output_mode(args.host_out_mode)

    def e2e_step():
        g.clear()
        t0 = time.perf_counter()
        for lo in range(0, n, INGEST_BATCH):
            sl = slice(lo, lo + INGEST_BATCH)
            g.add_edges(hsrc[sl], hdst[sl], hts[sl], heid[sl])
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        r = smp.sample_layer_batched_numpy(p_nodes, p_rts, p_offs, 0, 0, out=host_out, ids32=ids32)  # returns with host arrays complete
        t2 = time.perf_counter()
        s_b = len(r["nbr"])
        s_tot = 0
        for b in range(nb):
            q = smp.sample_numpy(nodes[offs[b]:offs[b + 1]], rts[offs[b]:offs[b + 1]])
            s_tot += q[0][0]["num_src_nodes"] - q[0][0]["num_dst_nodes"]
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        # host roots in, MFG left on the GPU (what the models consume): H2D of the step's roots + the multi-batch launch
        # + an 8-byte read of the neighbour count
        dn = torch.from_numpy(p_nodes).to(dev, non_blocking=True)
        dt_ = torch.from_numpy(p_rts).to(dev, non_blocking=True)
        do = torch.from_numpy(p_offs.view(np.int64)).to(dev, non_blocking=True)
        smp.sample_layer_batched(dn, dt_, do, 0, 0, out=out)
        s_dev = int(out["edge_offsets"][-1].item())
        t4 = time.perf_counter()
        return t1 - t0, t2 - t1, s_b, t3 - t2, s_tot, t4 - t3, s_dev

    if e2e_steps:
        e2e_step()
        # the host arrays hold exactly what the device-resident launch produced
        for k in ("nbr", "ts", "dt", "eid", "row"):
            assert np.array_equal(host_out[k][:S].astype(out[k].cpu().numpy().dtype), out[k][:S].cpu().numpy()), "host-array call differs from device call: " + k
    barrier()
    e2e = [e2e_step() for _ in range(e2e_steps)]
    barrier()
    # the other output format beside it (3 calls, sampling part only)
    other = None
    if e2e_steps:
        ho2 = smp.alloc_batched_host_out(T, nb, 0, pinned=True, ids32=not ids32)
        smp.sample_layer_batched_numpy(p_nodes, p_rts, p_offs, 0, 0, out=ho2, ids32=not ids32)
        t0 = time.perf_counter()
        for _ in range(3):
            smp.sample_layer_batched_numpy(p_nodes, p_rts, p_offs, 0, 0, out=ho2, ids32=not ids32)
        other = (time.perf_counter() - t0) / 3
        assert np.array_equal(ho2["nbr"][:S].astype(np.int64), host_out["nbr"][:S].astype(np.int64)) and \
            np.array_equal(ho2["row"][:S].astype(np.int64), host_out["row"][:S].astype(np.int64)) and \
            np.array_equal(ho2["eid"][:S], host_out["eid"][:S]), "32-bit and 64-bit host outputs differ"
        del ho2
    e2e_smp_s = sum(x[1] for x in e2e)
    e2e_ing_s = sum(x[0] for x in e2e)
    e2e_pb_s = sum(x[3] for x in e2e)
    e2e_dev_s = sum(x[5] for x in e2e)
    assert all(x[2] == S and x[4] == S and x[6] == S for x in e2e), "host API and multi-batch launch disagree on the number of neighbours"

    tgat = None
    if world == 1 and e2e_steps and not args.no_per_batch_models:
        try:
            tgat = tgat_per_batch_leg(g, stream, nodes, rts, offs, dev)
        except Exception as e:  # noqa: BLE001
            tgat = {"error": "{}: {}".format(type(e).__name__, e)}
    # ---- max over ranks
    t = torch.tensor([total_ms, ing_ms, smp_ms, e2e_smp_s, e2e_ing_s, e2e_pb_s, ing_q_ms, e2e_dev_s], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(S)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_ms, ing_ms, smp_ms, e2e_smp_s, e2e_ing_s, e2e_pb_s, ing_q_ms, e2e_dev_s = [float(x) for x in t.tolist()]
    S_all = float(tot.item())
    part = None
    if world > 1 and not args.no_partitioned:
        # BASELINE config 4 on the same ranks: GDELT-shaped graph hash-partitioned over the GPUs, DySAT sampling through the
        # peer-memory exchange kernels, partitioned feature rows over NVLink; checked in-run against the unpartitioned
        # sampler (bench_configs.partitioned_leg).  Collective: every rank runs it.
        import bench_configs as BC
        del out, d_src, d_dst, d_ts, d_eid
        torch.cuda.empty_cache()
        part = BC.partitioned_leg(dev, local, rank, world, args.part_shape, args.part_scale, super_batches=args.part_batches,
                                  steps=5, warmup=2)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    K = args.steps
    value = S_all * K / (smp_ms * 1e-3)
    # every rank inserts the SAME stream into its own replica: the rate is per replica, not multiplied by the ranks
    ingest_value = n * K / (ing_ms * 1e-3)
    e2e_value = S_all * e2e_steps / e2e_smp_s if e2e_steps else None
    # ---- roofline of the dominant sampling kernel (algorithmic bytes: DESIGN.md section 4)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    log_n = int(np.ceil(np.log2(mean_block + 1)))
    T_e = frac_with_edges * T
    bytes_locate = T * (12 + 8 + 16 + 4) + T_e * (32 + 8 * log_n)
    bytes_emit = T * (16 + 8 + 4) + S * (20 + 24 + 8)
    bytes_scan = T * 8
    if args.variant >= 2:  # one fused launch: per-target state stays on chip (no locs / counts / offsets traffic)
        bytes_fused = T * (12 + 8 + 4) + T_e * (32 + 8 * log_n) + S * (20 + 24 + 8)
        # the name as ncu prints it: <LIST = 0 (no active-target list), OCC = 4, POLICY = 0 (recent)>
        kern = {"sample_persistent_kernel<0, 4, 0>": (prof_s["emit"], bytes_fused)}
        bytes_locate, bytes_emit, bytes_scan = bytes_fused, 0, 0
    else:
        kern = {"locate_warp_kernel" if args.variant == 0 else "locate_thread_kernel": (prof_s["locate"], bytes_locate),
                "exclusive_scan_u32": (prof_s["scan"], bytes_scan), "emit_kernel": (prof_s["emit"], bytes_emit)}
    dom = max(kern, key=lambda k: kern[k][0][0])
    (dom_ms, dom_cnt), dom_bytes = kern[dom]
    dom_ms_per_launch = dom_ms / max(1, dom_cnt)
    achieved = dom_bytes / (dom_ms_per_launch * 1e-3) / 1e9
    step_bytes = bytes_locate + bytes_emit + bytes_scan
    step_kernel_ms = sum(v[0][0] for v in kern.values()) / max(1, dom_cnt)
    traffic, traffic_src = None, None
    try:  # DRAM bytes per launch of this kernel from the committed ncu capture of this same workload
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dom)
        if tr and args.dataset == "REDDIT":
            traffic, traffic_src = tr["dram_bytes"], tr["source"]
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms_per_launch,
                "kernel_share_of_sample_phase": dom_ms / max(1e-9, sum(v[0][0] for v in kern.values())),
                "sample_step": {"algorithmic_bytes": step_bytes, "kernel_ms": step_kernel_ms,
                                "achieved": step_bytes / (step_kernel_ms * 1e-3) / 1e9,
                                "frac": step_bytes / (step_kernel_ms * 1e-3) / 1e9 / peak,
                                "ms": {k: v[0][0] / max(1, v[0][1]) for k, v in kern.items()}},
                "inputs": {"targets": T, "targets_with_edges_frac": frac_with_edges, "neighbors": S,
                           "mean_block_size": mean_block, "log2_probes": log_n}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(3, args.warmup),
            "ms_per_step": smp_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64+f32", "data": "synthetic",
            "config": workload_config(stream, {
                "parallelism": "replicated graph, data-parallel sampler, {} rank(s); each rank replays the stream with "
                               "its own negatives".format(world),
                "batches_per_launch": nb, "l2": "not flushed: each step streams ~{:.0f} MB of roots + outputs (> 126 MB "
                "L2); the {:.0f} MB graph is L2-resident by the nature of this dataset".format(
                    (T * 12 + S * 32) / 1e6, g.get_graph_memory_usage() / 1e6)}),
            "step_ms_total": total_ms / K,
            "launch_gap": {"us": args.launch_gap_us,
                           "note": "device-side spin queued between a step's ingest (host-synchronous add_edges) and its "
                                   "sampling launch; inside step_ms_total, outside ms_per_step and ingest.ms_per_step"},
            "ingest": {"metric": "edges_inserted_per_s", "value": ingest_value, "unit": "edges/s", "ms_per_step": ing_ms / K,
                       "batches": (n + INGEST_BATCH - 1) // INGEST_BATCH, "replicas": world,
                       "api": "DynamicGraph.add_edges(cuda tensors) per {}-edge batch: the reference's call, one host "
                              "synchronisation per batch; per replica (every rank inserts the same stream)".format(INGEST_BATCH),
                       "us_per_batch": ing_ms / K / ((n + INGEST_BATCH - 1) // INGEST_BATCH) * 1e3,
                       "algorithmic_GBps": ingest_value * 48 / 1e9, "frac_of_hbm": ingest_value * 48 / 1e9 / peak,
                       "queued": {"value": n * K / (ing_q_ms * 1e-3), "unit": "edges/s",
                                  "api": "add_edges_async per batch + one flush per replay (no reference equivalent)"},
                       "phase_ms_per_batch": {k: v[0] / max(1, v[1]) for k, v in prof_g.items()}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(T * 12 + (nb + 1) * 8),
                    "d2h_bytes_per_step": int(S * out_bytes + (nb + 1) * 8), "steps": e2e_steps,
                    "api": "TemporalSampler.sample_layer_batched_numpy(ids32={}): one C-ABI call per step "
                           "({}) over the step's {} batches; pinned host arrays in, pinned host arrays out ({}); "
                           "{}".format(
                               ids32, "gf_sampler_sample_layer_batched_ids32" if ids32 else "gf_sampler_sample_layer_batched, GF_PTR_HOST", nb,
                               "written in place by the kernel over PCIe" if args.host_out_mode == 2 and not ids32 else "device arrays + cudaMemcpyAsync D2H",
                               "neighbour ids and rows as uint32 (24 B per neighbour; vertex ids < 2^32 by the store's contract)" if ids32
                               else "neighbour ids and rows as int64 (32 B per neighbour)"),
                    "bytes_per_neighbor": out_bytes,
                    "other_format": {"bytes_per_neighbor": 56 - out_bytes, "value": S / other if other else None, "unit": UNIT,
                                     "ms_per_step": other * 1e3 if other else None, "steps": 3,
                                     "note": "this rank only; the same call with {} ids / rows, values compared".format("int64" if ids32 else "uint32")},
                    "ms_per_step": e2e_smp_s / e2e_steps * 1e3 if e2e_steps else None,
                    "pcie_GBps": (S * out_bytes + T * 12) / (e2e_smp_s / e2e_steps) / 1e9 if e2e_steps else None,
                    "host_affinity": numa,
                    "per_batch": {"value": S_all * e2e_steps / e2e_pb_s if e2e_steps else None, "unit": UNIT,
                                  "api": "TemporalSampler.sample_numpy(numpy) once per batch of 600, synchronous",
                                  "ms_per_batch": e2e_pb_s / e2e_steps / nb * 1e3 if e2e_steps else None,
                                  "h2d_bytes_per_step": int(T * 12), "d2h_bytes_per_step": int((T + S) * 12 + S * 28)},
                    "ingest": {"value": n * e2e_steps / e2e_ing_s if e2e_steps else None, "unit": "edges/s",
                               "api": "DynamicGraph.add_edges(numpy) per 100000-edge batch, per replica",
                               "h2d_bytes_per_step": int(n * 28)},
                    "ingest_edges_per_s": n * e2e_steps / e2e_ing_s if e2e_steps else None},
            "per_batch": {"value": S_all * e2e_steps / e2e_pb_s if e2e_steps else None, "unit": UNIT,
                          "us_per_batch": e2e_pb_s / e2e_steps / nb * 1e6 if e2e_steps else None,
                          "api": "TemporalSampler.sample_numpy(numpy) once per batch of 600 (1,800 roots), synchronous: the "
                                 "reference's own per-batch call shape (host arrays in, host arrays out)"},
            "e2e_device": {"value": S_all * e2e_steps / e2e_dev_s if e2e_steps else None, "unit": UNIT,
                           "ms_per_step": e2e_dev_s / e2e_steps * 1e3 if e2e_steps else None,
                           "h2d_bytes_per_step": int(T * 12 + (nb + 1) * 8), "d2h_bytes_per_step": 8,
                           "api": "pinned host roots -> device, TemporalSampler.sample_layer_batched, the MFG arrays stay on "
                                  "the GPU (what the models consume); 8-byte read of the neighbour count"},
            "parity_checked": True,
            "parity": {"against": "CPU oracle (oracle/gnnflow_oracle.c), bit-exact, outside the timed region",
                       "batches": parity_batches, "neighbors": parity_neighbors},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": clk}
    if tgat is not None:
        line["tgat_per_batch"] = tgat.get("tgat", tgat)
        if "tgn" in tgat:
            line["tgn_per_batch"] = tgat["tgn"]
    if part is not None:
        line["partitioned"] = part
    if world == 1 and not args.no_hbm_bound:
        # the sampler where the graph does not fit the L2 (GDELT shapes, saturated multi-batch launches)
        try:
            import bench_configs as BC
            del out, d_src, d_dst, d_ts, d_eid
            torch.cuda.empty_cache()
            hb = {}
            for shape in ("GDELT-16.7K", "GDELT-16.7M"):
                hb[shape] = BC.hbm_bound_leg(dev, local, shape, args.hbm_scale, steps=5, warmup=3)
            line["hbm_bound"] = hb
        except Exception as e:  # noqa: BLE001
            line["hbm_bound"] = {"error": "{}: {}".format(type(e).__name__, e)}
        try:  # ingest at saturation: one 9.56 M-edge batch through the synchronous add_edges
            import bench_configs as BC
            line["ingest_large_batch"] = BC.ingest_large_leg(dev, local)
        except Exception as e:  # noqa: BLE001
            line["ingest_large_batch"] = {"error": "{}: {}".format(type(e).__name__, e)}
    if world == 1 and not args.no_hbm_bound:
        try:
            line["cache_gather"] = cache_gather_leg(dev, peak)
        except Exception as e:  # noqa: BLE001
            line["cache_gather"] = {"error": "{}: {}".format(type(e).__name__, e)}
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_port_run(stream, nodes, rts, offs, seconds=args.cpu_seconds)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "ingest_edges_per_s")}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL prints its version banner on stdout) must not get between the driver and the ONE JSON line:
    route fd 1 to stderr for the whole run and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse()
    quiet_stdout()
    from gnnflow_b200.synth import synth, tgn_batches
    rank = int(os.environ.get("RANK", "0"))
    stream = synth(args.dataset, seed=42)
    nodes, rts, offs = tgn_batches(stream, BATCH, seed=7 + (rank if args.impl == "ours" else 0))
    if args.impl == "reference":
        reference_arm(args, stream, nodes, rts, offs)
    elif args.impl == "reference-real":
        reference_real(args, stream, nodes, rts, offs)
    else:
        ours(args, stream, nodes, rts, offs)


if __name__ == "__main__":
    main()
