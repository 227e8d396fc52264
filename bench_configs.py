#!/usr/bin/env python
"""bench_configs.py -- the BASELINE.json configs that are not bench.py's headline line, measured the same way
(CUDA events, inputs resident in HBM, max over ranks) and printed as one JSON line per run.

  --config wiki          configs[0]: WIKI-shaped (9,227 nodes, 157,474 edges, undirected), TGN [10] recent, batch 600
  --config tgat          configs[2]: REDDIT-shaped, TGAT [10,10] uniform + LRUCache(0.2) edge-feature gather, data-parallel
  --config dysat         configs[3]: GDELT-shaped, DySAT [10,10] uniform, 3 snapshots, window 25, prop_time;
                                     --gpus N > 1: hash-partitioned over N ranks with NCCL all-to-all
  --config online        configs[4]: GDELT-shaped, 100k-edge add_edges interleaved with 2-layer sampling of those edges
  --config sweep         REDDIT TGN [10] recent: targets per launch swept 1.8k -> 2M (per-batch latency vs saturation)
  --config ingest_sweep  add_edges batch size swept 1k -> 16M edges (launch-bound -> bandwidth-bound)

GDELT-shaped streams are generated on the GPU (`--scale` shrinks the 191M edges; the default 0.1 keeps a run within
a couple of minutes; the vertex count is kept, so the table / random-access regime is the full one).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bench as B  # noqa: E402


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--config", required=True, choices=["wiki", "tgat", "dysat", "online", "sweep", "ingest_sweep", "two_layer_sat", "hbm_bound", "partitioned"])
    p.add_argument("--dataset", default="REDDIT", choices=["REDDIT", "WIKI"])
    p.add_argument("--strategy", default="uniform", choices=["uniform", "recent"])
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--variant", type=int, default=None, help="sampling kernel variant (two_layer_sat; default = library default)")
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--scale", type=float, default=0.1)
    p.add_argument("--shape", default="GDELT-16.7M", choices=["GDELT-16.7M", "GDELT-16.7K"])
    p.add_argument("--max-batches", type=int, default=400)
    p.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                   help="dysat, N > 1: kernels write into peer windows over NVLink (default) or NCCL all-to-all")
    return p.parse_args()


def dist_setup():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, local, dev


def peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def synth_gpu(shape, scale, dev, seed=42):
    """GDELT-shaped stream generated on the device (same recipe as gnnflow_b200.synth: Zipf(0.8) endpoints,
    sorted uniform timestamps with ties, eid = arange)."""
    import torch
    from gnnflow_b200.synth import SHAPES
    num_src, num_dst, num_edges, undirected, minblk = SHAPES[shape]
    n = int(num_edges * scale)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    w = 1.0 / torch.arange(1, num_src + 1, dtype=torch.float64, device=dev) ** 0.8
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()

    def zipf(k):
        out = torch.empty(k, dtype=torch.int64, device=dev)
        for lo in range(0, k, 1 << 24):
            hi = min(k, lo + (1 << 24))
            u = torch.rand(hi - lo, dtype=torch.float64, device=dev, generator=g)
            out[lo:hi] = torch.searchsorted(cdf, u, right=True).clamp_(0, num_src - 1)
        return out
    src, dst = zipf(n), zipf(n)
    t_max = 2.6e6
    ts = torch.sort(torch.rand(n, dtype=torch.float64, device=dev, generator=g) * t_max).values.to(torch.float32)
    eid = torch.arange(n, dtype=torch.int64, device=dev)
    return dict(src=src, dst=dst, ts=ts, eid=eid, num_nodes=num_src, minimum_block_size=minblk, name=shape, n=n)


def sampling_bytes(T, S, e_frac, log_n, mfg):
    """algorithmic bytes of one (layer, snapshot) launch, DESIGN.md section 3 / SURVEY 8d"""
    b = T * (12 + 8 + 4) + e_frac * T * (32 + 8 * log_n) + S * (20 + 24 + 8)
    if mfg:
        b += S * 8 + (T + S) * 12 - S * 12  # col, and roots echoed into all_nodes / all_ts (neighbour part counted above)
    return b


def emit(line):
    B.emit(line)


# ------------------------------------------------------------------------------------------------ configs
def run_wiki(args):
    """configs[0]: same machinery as bench.py's headline on the WIKI shape (undirected: add_reverse)"""
    import torch
    rank, world, local, dev = dist_setup()
    from gnnflow_b200 import DynamicGraph, TemporalSampler
    from gnnflow_b200.synth import synth, tgn_batches
    stream = synth("WIKI", seed=42)
    nodes, rts, offs = tgn_batches(stream, B.BATCH, seed=7 + rank)
    g = DynamicGraph(**B.graph_config(stream), device=local)
    d = {k: torch.from_numpy(stream[k]).to(dev) for k in ("src", "dst", "ts", "eid")}
    n = len(stream["src"])

    def ingest():
        g.clear()
        for lo in range(0, n, B.INGEST_BATCH):
            sl = slice(lo, lo + B.INGEST_BATCH)
            g.add_edges(d["src"][sl], d["dst"][sl], d["ts"][sl], d["eid"][sl], add_reverse=True)
    smp = TemporalSampler(g, [B.FANOUT], "recent")
    dn, dt, do = torch.from_numpy(nodes).to(dev), torch.from_numpy(rts).to(dev), torch.from_numpy(offs).to(dev)
    out = None
    for _ in range(max(3, args.warmup)):
        ingest()
        out = smp.sample_layer_batched(dn, dt, do, 0, 0, out=out)
    torch.cuda.synchronize()
    S = int(out["edge_offsets"][-1].item())
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    smp.set_profiling(True)
    smp.get_profile(True)
    a, b, c = ev(), ev(), ev()
    ing_ms = smp_ms = 0.0
    for _ in range(args.steps):
        a.record(); ingest(); b.record(); smp.sample_layer_batched(dn, dt, do, 0, 0, out=out); c.record()
        torch.cuda.synchronize()
        ing_ms += a.elapsed_time(b); smp_ms += b.elapsed_time(c)
    prof = smp.get_profile(True)["emit"]
    peak, src_ = peak_hbm()
    T = len(nodes)
    deg = g.out_degree(nodes[:200000])
    e_frac = float((deg > 0).mean())
    nblk = max(1, int(round(g.avg_linked_list_length() * g.num_vertices())))
    log_n = int(np.ceil(np.log2(2 * n / nblk + 1)))
    kb = sampling_bytes(T, S, e_frac, log_n, False)
    kms = prof[0] / max(1, prof[1])
    emit({"metric": B.METRIC, "value": S * args.steps / (smp_ms * 1e-3), "unit": B.UNIT, "n_gpus": world,
          "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": smp_ms / args.steps,
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64+f32", "data": "synthetic",
          "config": {"workload": "WIKI-shaped synthetic (9227 nodes, 157474 edges, undirected -> 314948 stored), TGN "
                                 "1-layer recent fanout 10, batch 600; one step = one replay", "targets": T, "neighbors": S},
          "ingest": {"value": 2 * n * args.steps / (ing_ms * 1e-3), "unit": "stored edges/s"},
          "roofline": {"bound": "hbm", "kernel": "sample_persistent_kernel", "achieved": kb / (kms * 1e-3) / 1e9,
                       "peak": peak, "unit": "GB/s", "frac": kb / (kms * 1e-3) / 1e9 / peak, "traffic": None,
                       "peak_source": src_, "ms_per_launch": kms, "algorithmic_bytes_per_launch": kb}})


def run_sweep(args):
    """per-batch latency vs saturation: G batches of 1,800 roots per launch, G = 1 .. 1121"""
    import torch
    rank, world, local, dev = dist_setup()
    from gnnflow_b200 import DynamicGraph, TemporalSampler
    from gnnflow_b200.synth import synth, tgn_batches
    stream = synth("REDDIT", seed=42)
    nodes, rts, offs = tgn_batches(stream, B.BATCH, seed=7)
    g = DynamicGraph(**B.graph_config(stream), device=local)
    n = len(stream["src"])
    for lo in range(0, n, B.INGEST_BATCH):
        sl = slice(lo, lo + B.INGEST_BATCH)
        g.add_edges(stream["src"][sl], stream["dst"][sl], stream["ts"][sl], stream["eid"][sl])
    smp = TemporalSampler(g, [B.FANOUT], "recent")
    dn, dt = torch.from_numpy(nodes).to(dev), torch.from_numpy(rts).to(dev)
    nb = len(offs) - 1
    peak, src_ = peak_hbm()
    rows = []
    for G in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, nb):
        groups = [(b0, min(nb, b0 + G)) for b0 in range(0, nb, G)][:max(1, 2048 // G)]
        args_ = []
        for b0, b1 in groups:
            lo, hi = int(offs[b0]), int(offs[b1])
            o = torch.from_numpy(offs[b0:b1 + 1] - offs[b0]).to(dev)
            args_.append((dn[lo:hi], dt[lo:hi], o))
        out = None
        Tmax = max(a[0].shape[0] for a in args_)
        out = dict(nbr=torch.empty(Tmax * 10, dtype=torch.int64, device=dev), ts=torch.empty(Tmax * 10, dtype=torch.float32, device=dev),
                   dt=torch.empty(Tmax * 10, dtype=torch.float32, device=dev), eid=torch.empty(Tmax * 10, dtype=torch.int64, device=dev),
                   row=torch.empty(Tmax * 10, dtype=torch.int64, device=dev),
                   edge_offsets=torch.empty(G + 1, dtype=torch.int64, device=dev))
        S = 0
        for a in args_:
            smp.sample_layer_batched(*a, 0, 0, out=out)
            S += int(out["edge_offsets"][a[2].shape[0] - 1].item())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(1, args.steps)
        e0.record()
        for _ in range(reps):
            for a in args_:
                smp.sample_layer_batched(*a, 0, 0, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        T = sum(a[0].shape[0] for a in args_)
        rows.append({"batches_per_launch": G, "targets_per_launch": T // len(args_), "launches": len(args_),
                     "us_per_launch": ms * 1e3 / len(args_), "neighbors_per_s": S / (ms * 1e-3),
                     "algorithmic_GBps": sampling_bytes(T, S, 0.637, 6, False) / (ms * 1e-3) / 1e9})
    emit({"metric": B.METRIC, "unit": B.UNIT, "n_gpus": 1, "data": "synthetic", "value": rows[-1]["neighbors_per_s"],
          "config": {"workload": "REDDIT-shaped TGN [10] recent: targets per launch swept (back-to-back launches on one "
                                 "stream, device-resident inputs)"}, "peak": peak, "peak_source": src_, "sweep": rows})


def run_ingest_sweep(args):
    import torch
    rank, world, local, dev = dist_setup()
    from gnnflow_b200 import DynamicGraph
    st = synth_gpu(args.shape, args.scale, dev)
    n = st["n"]
    peak, src_ = peak_hbm()
    cfg = dict(initial_pool_size=int(n * 40), maximum_pool_size=150 << 30, mem_resource_type="cuda",
               minimum_block_size=st["minimum_block_size"], blocks_to_preallocate=1024, insertion_policy="insert")
    g = DynamicGraph(**cfg, device=local)
    rows = []
    for bs in (1000, 10000, 100000, 1000000, 4000000, 16000000):
        if bs > n:
            break
        m = min(n, max(bs * 4, min(n, 20_000_000)))

        def run():
            g.clear()
            for lo in range(0, m, bs):
                sl = slice(lo, min(m, lo + bs))
                g.add_edges(st["src"][sl], st["dst"][sl], st["ts"][sl], st["eid"][sl])
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        g.set_profiling(True); g.get_profile(True); run(); torch.cuda.synchronize()   # phase split (events between phases)
        ph = {k: v[0] / max(1, v[1]) * 1e3 for k, v in g.get_profile(True).items()}
        g.set_profiling(False)
        rows.append({"batch_edges": bs, "edges": m, "edges_per_s": m / (ms * 1e-3), "us_per_batch": ms * 1e3 / ((m + bs - 1) // bs),
                     "algorithmic_GBps": m * 48 / (ms * 1e-3) / 1e9, "frac": m * 48 / (ms * 1e-3) / 1e9 / peak,
                     "phase_us_per_batch": ph})
    emit({"metric": "edges_inserted_per_s", "unit": "edges/s", "n_gpus": 1, "data": "synthetic", "value": rows[-1]["edges_per_s"],
          "config": {"workload": "{}-shaped stream (scale {}): add_edges batch size swept, device-resident inputs, "
                                 "48 algorithmic B/edge".format(args.shape, args.scale), "num_nodes": st["num_nodes"]},
          "peak": peak, "peak_source": src_, "sweep": rows})


def run_two_layer_sat(args):
    """2-layer [10,10] sampling at saturation: every batch of the replay in ONE launch per layer.  Layer 1 samples
    [roots_b || neighbours_b] of every batch b (what gf_sampler_sample chains per batch); the per-batch concatenation
    is prepared with torch outside the timed region, the timed region is the two kernel launches."""
    import torch
    rank, world, local, dev = dist_setup()
    from gnnflow_b200 import DynamicGraph, TemporalSampler
    from gnnflow_b200.synth import synth, tgn_batches
    stream = synth(args.dataset, seed=42)
    nodes, rts, offs = tgn_batches(stream, B.BATCH, seed=7)
    g = DynamicGraph(**B.graph_config(stream), device=local)
    n = len(stream["src"])
    rev = bool(stream["undirected"])
    for lo in range(0, n, B.INGEST_BATCH):
        sl = slice(lo, lo + B.INGEST_BATCH)
        g.add_edges(stream["src"][sl], stream["dst"][sl], stream["ts"][sl], stream["eid"][sl], add_reverse=rev)
    F = [10, 10]
    smp = TemporalSampler(g, F, args.strategy)
    if args.variant is not None:
        smp.set_variant(args.variant)
    dn0, dt0, do0 = torch.from_numpy(nodes).to(dev), torch.from_numpy(rts).to(dev), torch.from_numpy(offs).to(dev)
    T0, nb = dn0.shape[0], do0.shape[0] - 1
    out0 = smp.sample_layer_batched(dn0, dt0, do0, 0, 0)
    eo = out0["edge_offsets"].to(torch.int64)
    S0 = int(eo[-1].item())
    # layer-1 targets: per batch [roots || neighbours]
    broot = torch.bucketize(torch.arange(T0, device=dev), do0[1:], right=True)
    bedge = torch.bucketize(torch.arange(S0, device=dev), eo[1:], right=True)
    T1 = T0 + S0
    dn1 = torch.empty(T1, dtype=torch.int64, device=dev)
    dt1 = torch.empty(T1, dtype=torch.float32, device=dev)
    ri = torch.arange(T0, device=dev) + eo[broot]
    ei = torch.arange(S0, device=dev) + do0[bedge + 1]
    dn1[ri], dt1[ri] = dn0, dt0
    dn1[ei], dt1[ei] = out0["nbr"][:S0], out0["ts"][:S0]
    do1 = do0 + eo
    del broot, bedge, ri, ei
    out1 = smp.sample_layer_batched(dn1, dt1, do1, 1, 0)
    S1 = int(out1["edge_offsets"][-1].item())
    # the same chaining done by the library on the device (gf_sampler_chain_batched): this is what is timed
    chain = smp.chain_batched(dn0, dt0, do0, out0)
    assert torch.equal(chain[0][:T1], dn1) and torch.equal(chain[1][:T1], dt1) and torch.equal(chain[2], do1.to(torch.int64))
    cn1, ct1 = chain[0][:T1], chain[1][:T1]
    for _ in range(max(3, args.warmup)):
        smp.sample_layer_batched(dn0, dt0, do0, 0, 0, out=out0)
        smp.chain_batched(dn0, dt0, do0, out0, out=chain)
        smp.sample_layer_batched(cn1, ct1, chain[2], 1, 0, out=out1)
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    ms = [0.0, 0.0]
    chain_ms = 0.0
    for _ in range(args.steps):
        a, b, b2, c = ev(), ev(), ev(), ev()
        a.record(); smp.sample_layer_batched(dn0, dt0, do0, 0, 0, out=out0)
        b.record(); smp.chain_batched(dn0, dt0, do0, out0, out=chain)
        b2.record(); smp.sample_layer_batched(cn1, ct1, chain[2], 1, 0, out=out1)
        c.record(); torch.cuda.synchronize()
        ms[0] += a.elapsed_time(b) / args.steps; ms[1] += b2.elapsed_time(c) / args.steps
        chain_ms += b.elapsed_time(b2) / args.steps
    peak, src_ = peak_hbm()
    nblk = max(1, int(round(g.avg_linked_list_length() * g.num_vertices())))
    stored = n * (2 if rev else 1)
    log_n = int(np.ceil(np.log2(stored / nblk + 1)))
    layers = []
    for T, S, dn, m in ((T0, S0, dn0, ms[0]), (T1, S1, dn1, ms[1])):
        samp = dn[torch.randint(0, T, (200000,), device=dev)].cpu().numpy()
        e_frac = float((g.out_degree(samp) > 0).mean())
        kb = sampling_bytes(T, S, e_frac, log_n, False)
        layers.append({"targets": T, "neighbors": S, "targets_with_edges_frac": e_frac, "ms": m,
                       "algorithmic_bytes": kb, "achieved_GBps": kb / (m * 1e-3) / 1e9, "frac": kb / (m * 1e-3) / 1e9 / peak,
                       "neighbors_per_s": S / (m * 1e-3)})
    chain_b = T1 * 24.0  # the chaining pass: 12 B read + 12 B written per next-layer target
    tot_ms = ms[0] + chain_ms + ms[1]
    tot_b = layers[0]["algorithmic_bytes"] + layers[1]["algorithmic_bytes"] + chain_b
    emit({"metric": B.METRIC, "value": (S0 + S1) / (tot_ms * 1e-3), "unit": B.UNIT, "n_gpus": 1, "steps": args.steps,
          "warmup": max(3, args.warmup), "ms_per_step": tot_ms, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "int64+f32", "data": "synthetic",
          "config": {"workload": "{}-shaped synthetic, 2-layer {} [10,10], batch 600 (1,800 roots), all {} batches of the "
                                 "replay per launch (one launch per layer + the device-side chaining of layer 1's targets, all "
                                 "inside the timed region), device-resident".format(args.dataset, args.strategy, nb),
                     "mean_block_size": stored / nblk, "log2_probes": log_n},
          "layers": layers,
          "chain": {"kernel": "chain_batched_kernel", "ms": chain_ms, "algorithmic_bytes": chain_b,
                    "frac": chain_b / (chain_ms * 1e-3) / 1e9 / peak},
          "roofline": {"bound": "hbm", "kernel": "sample_persistent_kernel<0, 4, 0>", "achieved": tot_b / (tot_ms * 1e-3) / 1e9,
                       "peak": peak, "unit": "GB/s", "frac": tot_b / (tot_ms * 1e-3) / 1e9 / peak, "traffic": None,
                       "peak_source": src_}})


def run_tgat(args):
    """configs[2]: per batch of 600 edges: 2-layer uniform sampling [10,10] + LRUCache(0.2) edge-feature gather into
    every block (De = 172) through the public API; batches sharded b % world == rank (replicated graph)."""
    import torch
    import torch.distributed as dist
    rank, world, local, dev = dist_setup()
    from gnnflow_b200 import DynamicGraph, TemporalSampler
    from gnnflow_b200.cache import LRUCache
    from gnnflow_b200.distributed import shard_batch_indices
    from gnnflow_b200.synth import synth, tgn_batches
    stream = synth("REDDIT", seed=42)
    nodes, rts, offs = tgn_batches(stream, B.BATCH, seed=7)
    n = len(stream["src"])
    g = DynamicGraph(**B.graph_config(stream), device=local)
    for lo in range(0, n, B.INGEST_BATCH):
        sl = slice(lo, lo + B.INGEST_BATCH)
        g.add_edges(stream["src"][sl], stream["dst"][sl], stream["ts"][sl], stream["eid"][sl])
    De = 172
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    efeat = torch.randn(n, De, device=dev, generator=gen)
    cache = LRUCache(0.2, 0.2, stream["num_nodes"], n, dev, None, efeat, 0, De)
    cache.init_cache()
    smp = TemporalSampler(g, [10, 10], "uniform")
    dn, dt = torch.from_numpy(nodes).to(dev), torch.from_numpy(rts).to(dev)
    mine = shard_batch_indices(len(offs) - 1, rank, world)[:args.max_batches]
    eids = torch.from_numpy(stream["eid"]).to(dev)

    def step(count=False):
        S = rows = 0
        for b in mine:
            lo, hi = int(offs[b]), int(offs[b + 1])
            mfgs = smp.sample(dn[lo:hi], dt[lo:hi])
            cache.fetch_feature(mfgs, eid=eids[b * B.BATCH:(b + 1) * B.BATCH])
            if count:
                for lay in mfgs:
                    for blk in lay:
                        S += blk.num_edges(); rows += blk.num_edges()
        return S, rows
    for _ in range(max(1, args.warmup - 1)):
        step()
    S, rows = step(True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    smp.set_profiling(True)
    smp.get_profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    prof = smp.get_profile(True)["emit"]
    # the gather alone, saturated: all edge ids of the last block, repeated
    ids = torch.randint(0, n, (1 << 20,), device=dev, generator=gen)
    blk = type("Blk", (), {})()
    blk.srcdata, blk.edata = {"ID": ids[:1]}, {"ID": ids}
    for _ in range(3):
        cache.fetch_feature([[blk]], update_cache=False, target_edge_features=False)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(10):
        cache.fetch_feature([[blk]], update_cache=False, target_edge_features=False)
    g1.record()
    torch.cuda.synchronize()
    gms = g0.elapsed_time(g1) / 10
    gbytes = ids.shape[0] * (17 + 8 * De)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(S), float(len(mine))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    peak, src_ = peak_hbm()
    if rank == 0:
        emit({"metric": B.METRIC, "value": float(tot[0]) / (float(t[0]) * 1e-3), "unit": B.UNIT, "n_gpus": world,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(t[0]), "higher_is_better": True,
              "scaling": "weak" if args.max_batches * world <= len(offs) - 1 else "strong", "vs_baseline": None,
              "dtype": "int64+f32", "data": "synthetic",
              "config": {"workload": "REDDIT-shaped, TGAT 2-layer uniform [10,10] + LRUCache(0.2) edge-feature gather "
                                     "(De=172) per batch of 600 edges through the public API (sample + fetch_feature), "
                                     "replicated graph, batches sharded b % world == rank",
                         "batches_total": int(tot[1]), "batches_per_rank": len(mine)},
              "ms_per_batch": float(t[0]) / max(1, len(mine)),
              "sampler_kernel_us_per_launch": prof[0] / max(1, prof[1]) * 1e3,
              "cache": {"hit_ratio_edges": float(cache.cache_edge_ratio), "rows_per_step": rows,
                        "gather_1M_rows_ms": gms, "gather_GBps": gbytes / (gms * 1e-3) / 1e9,
                        "gather_frac_of_peak": gbytes / (gms * 1e-3) / 1e9 / peak},
              "roofline": {"bound": "hbm", "kernel": "cache_gather_kernel", "achieved": gbytes / (gms * 1e-3) / 1e9,
                           "peak": peak, "unit": "GB/s", "frac": gbytes / (gms * 1e-3) / 1e9 / peak, "traffic": None,
                           "peak_source": src_}})
    if world > 1:
        dist.destroy_process_group()


def _gdelt_graph(args, dev, local, rank, world, partitioned):
    import torch
    from gnnflow_b200 import DynamicGraph
    from gnnflow_b200.distributed import owner_of
    st = synth_gpu(args.shape, args.scale, dev)
    n = st["n"]
    cfg = dict(initial_pool_size=int(n * 30 / (world if partitioned else 1)) + (64 << 20), maximum_pool_size=150 << 30,
               mem_resource_type="cuda", minimum_block_size=st["minimum_block_size"], blocks_to_preallocate=1024,
               insertion_policy="insert")
    g = DynamicGraph(**cfg, device=local)
    keep = None
    if partitioned and world > 1:
        keep = owner_of(st["src"], world) == rank
    return st, g, keep


def run_dysat(args):
    """configs[3]: DySAT [10,10] uniform, 3 snapshots, window 25, prop_time on the GDELT shape.  N = 1: one GPU holds
    the graph.  N > 1: vertices hash-partitioned by source over the ranks, targets / neighbours exchanged with NCCL
    all-to-all (gnnflow_b200.distributed)."""
    import torch
    import torch.distributed as dist
    rank, world, local, dev = dist_setup()
    from gnnflow_b200 import TemporalSampler
    from gnnflow_b200.distributed import CudaEngine, DistributedTemporalSampler, PeerTemporalSampler
    st, g, keep = _gdelt_graph(args, dev, local, rank, world, True)
    n = st["n"]
    IB = 1_000_000
    t0 = time.perf_counter()
    for lo in range(0, n, IB):
        sl = slice(lo, min(n, lo + IB))
        if keep is None:
            g.add_edges(st["src"][sl], st["dst"][sl], st["ts"][sl], st["eid"][sl])
        else:
            k = keep[sl]
            if bool(k.any()):
                g.add_edges(st["src"][sl][k], st["dst"][sl][k], st["ts"][sl][k], st["eid"][sl][k])
    torch.cuda.synchronize()
    ingest_s = time.perf_counter() - t0
    window = 25.0 if args.shape == "GDELT-16.7K" else 25.0 * 1000  # keep ~the same edges per window at avg degree 11
    kw = dict(sample_strategy="uniform", num_snapshots=3, snapshot_time_window=window, prop_time=True)
    local_s = TemporalSampler(g, [10, 10], **kw)
    if world > 1 and args.exchange == "peer":
        smp = PeerTemporalSampler(local_s, max_targets=1800 * 11 + 64)
    elif world > 1:
        smp = DistributedTemporalSampler(CudaEngine(local_s), [10, 10], 3)
    else:
        smp = local_s
    gen = torch.Generator(device=dev)
    gen.manual_seed(100 + rank)
    nb = args.max_batches
    # every rank samples its own batches: the last `nb` chunks of 600 edges of the stream, + random negatives
    starts = (n - (torch.arange(nb, device=dev) * world + rank + 1) * B.BATCH).clamp_(min=0).tolist()

    def roots(s):
        sl = slice(s, s + B.BATCH)
        neg = torch.randint(0, st["num_nodes"], (B.BATCH,), device=dev, generator=gen)
        return torch.cat([st["src"][sl], st["dst"][sl], neg]), torch.cat([st["ts"][sl]] * 3)
    batches = [roots(s) for s in starts]

    xbytes = [0.0]

    def step(count=False):
        S = 0
        for nd, tt in batches:
            mfgs = smp.sample(nd, tt)
            if count:
                for lay in mfgs:
                    for blk in lay:
                        e = blk["num_src_nodes"] - blk["num_dst_nodes"] if isinstance(blk, dict) else blk.num_edges()
                        t_ = blk["num_dst_nodes"] if isinstance(blk, dict) else blk.num_dst_nodes()
                        S += e
                        # algorithmic NVLink bytes: 12 B / remote target out, 24 B / neighbour + 4 B / target back
                        xbytes[0] += (t_ * 16 + e * 24) * (world - 1) / world
        return S
    for _ in range(max(1, args.warmup - 1)):
        step()
    S = step(True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    local_s.set_profiling(True)
    local_s.get_profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    prof = local_s.get_profile(True)["emit"]
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(S), xbytes[0]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        emit({"metric": B.METRIC, "value": float(tot[0]) / (float(t[0]) * 1e-3), "unit": B.UNIT, "n_gpus": world,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(t[0]), "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "int64+f32", "data": "synthetic",
              "config": {"workload": "{}-shaped synthetic (scale {}: {} nodes, {} edges), DySAT 2-layer uniform [10,10], 3 "
                                     "snapshots, window {}, prop_time; {} batches of 1,800 roots per rank per step through "
                                     "the public API".format(args.shape, args.scale, st["num_nodes"], n, window, nb),
                         "parallelism": "hash-partitioned by source vertex over {} ranks, exchange = {}".format(
                             world, "kernels writing into peer windows over NVLink (gf_peer_*)" if args.exchange == "peer"
                             else "NCCL all-to-all")
                         if world > 1 else "single GPU"},
              "ms_per_batch": float(t[0]) / nb, "sampler_kernel_us_per_launch": prof[0] / max(1, prof[1]) * 1e3,
              "graph_GB": g.get_device_memory_usage() / 1e9, "ingest_edges_per_s": n / ingest_s,
              "exchange": {"bytes_per_step_all_ranks": float(tot[1]),
                           "GBps_per_gpu": float(tot[1]) / world / (float(t[0]) * 1e-3) / 1e9,
                           "nvlink_peak_GBps_per_direction": 900} if world > 1 else None})
    if world > 1:
        if hasattr(smp, "close"):
            smp.close()
        dist.destroy_process_group()


def run_online(args):
    """configs[4]: alternate add_edges(100k) with 2-layer recent sampling [10,10] of the batches those edges form
    (roots = src || dst || neg per 600 edges), replicated graph, each rank samples its shard of the batches."""
    import torch
    import torch.distributed as dist
    rank, world, local, dev = dist_setup()
    from gnnflow_b200 import TemporalSampler
    st, g, _ = _gdelt_graph(args, dev, local, rank, world, False)
    n = st["n"]
    IB = B.INGEST_BATCH
    warm = max(0, n - IB * (args.steps + args.warmup + 2))
    for lo in range(0, warm, 1_000_000):
        sl = slice(lo, min(warm, lo + 1_000_000))
        g.add_edges(st["src"][sl], st["dst"][sl], st["ts"][sl], st["eid"][sl])
    smp = TemporalSampler(g, [10, 10], "recent")
    gen = torch.Generator(device=dev)
    gen.manual_seed(5 + rank)
    per = IB // B.BATCH
    mine = [b for b in range(per) if b % world == rank][:args.max_batches]
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    ing, smpt, S_tot, edges = [], [], 0, 0
    pos = warm
    for it in range(args.warmup + args.steps):
        sl = slice(pos, pos + IB)
        a, b_, c = ev(), ev(), ev()
        a.record()
        g.add_edges(st["src"][sl], st["dst"][sl], st["ts"][sl], st["eid"][sl])
        b_.record()
        S = 0
        for bi in mine:
            s2 = slice(pos + bi * B.BATCH, pos + (bi + 1) * B.BATCH)
            neg = torch.randint(0, st["num_nodes"], (B.BATCH,), device=dev, generator=gen)
            mf = smp.sample(torch.cat([st["src"][s2], st["dst"][s2], neg]), torch.cat([st["ts"][s2]] * 3))
            S += sum(blk.num_edges() for lay in mf for blk in lay)
        c.record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            ing.append(a.elapsed_time(b_)); smpt.append(b_.elapsed_time(c)); S_tot += S; edges += IB
        pos += IB
    t = torch.tensor([sum(ing), sum(smpt)], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(S_tot)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        emit({"metric": B.METRIC, "value": float(tot[0]) / (float(t[1]) * 1e-3), "unit": B.UNIT, "n_gpus": world,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(t[0] + t[1]) / args.steps,
              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64+f32", "data": "synthetic",
              "config": {"workload": "{}-shaped (scale {}, {} nodes): online loop, add_edges(100000) then 2-layer recent "
                                     "[10,10] sampling of the {} batches of 600 those edges form (public API, per-batch "
                                     "calls); replicated graph, batches sharded over ranks".format(
                                         args.shape, args.scale, st["num_nodes"], per),
                         "edges_in_graph_before": warm},
              "ingest": {"value": edges / (float(t[0]) * 1e-3), "unit": "edges/s (every rank applies every batch)",
                         "ms_per_100k_batch": float(t[0]) / args.steps},
              "sample_ms_per_100k_edges": float(t[1]) / args.steps})
    if world > 1:
        dist.destroy_process_group()


def ingest_large_leg(dev, local, shapes=("GDELT-16.7K", "GDELT-16.7M"), scale=0.05, reps=5):
    """add_edges at saturation: the whole GDELT-shaped stream at `scale` (9.56 M edges) as ONE batch
    into an empty graph through the reference-shaped synchronous call, device-resident arrays, CUDA-event time around
    the call (best of `reps` after two warm-up replays).  Algorithmic bytes: 28 B read + 20 B written per edge."""
    import torch
    from gnnflow_b200 import DynamicGraph
    peak, src_ = peak_hbm()
    out = {"api": "DynamicGraph.add_edges(cuda tensors), one batch = the whole stream, graph cleared before each replay",
           "algorithmic_bytes_per_edge": 48, "peak": peak, "peak_source": src_, "unit": "GB/s", "shapes": {}}
    for shape in shapes:
        st = synth_gpu(shape, scale, dev)
        n = st["n"]
        g = DynamicGraph(initial_pool_size=256 << 20, maximum_pool_size=100 << 30, mem_resource_type="cuda",
                         minimum_block_size=st["minimum_block_size"], blocks_to_preallocate=1024, insertion_policy="insert",
                         device=local)

        def run():
            g.clear()
            g.add_edges(st["src"], st["dst"], st["ts"], st["eid"])
        run(); run()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g.clear()
            torch.cuda.synchronize()
            e0.record()
            g.add_edges(st["src"], st["dst"], st["ts"], st["eid"])
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        assert g.num_edges() == n
        out["shapes"][shape] = {"edges_per_batch": n, "num_nodes": st["num_nodes"], "ms_per_batch": best,
                                "edges_per_s": n / (best * 1e-3), "achieved_GBps": n * 48 / (best * 1e-3) / 1e9,
                                "frac_of_hbm": n * 48 / (best * 1e-3) / 1e9 / peak}
        del g, st
        torch.cuda.empty_cache()
    return out


def hbm_bound_leg(dev, local, shape="GDELT-16.7K", scale=1.0, targets=2_400_000, steps=5, warmup=3, strategies=("recent", "uniform"),
                  keep_graph=False):
    """The sampler where the graph does NOT fit the 126 MB L2: GDELT shape at full scale, one
    multi-batch launch of >= 2 M targets (TGN batches of the newest edges: src || dst || random negatives), fan-out 10,
    recent and uniform, layer 0 and -- chained on the device -- layer 1.  Returns a dict of per-launch records with
    algorithmic bytes (SURVEY 8d formula) / CUDA-event time against the measured HBM peak."""
    import torch
    from gnnflow_b200 import DynamicGraph, TemporalSampler
    st = synth_gpu(shape, scale, dev)
    n = st["n"]
    cfg = dict(initial_pool_size=(1 << 30), maximum_pool_size=160 << 30, mem_resource_type="cuda",
               minimum_block_size=st["minimum_block_size"], blocks_to_preallocate=1024, insertion_policy="insert")
    g = DynamicGraph(**cfg, device=local)
    IB = 4_000_000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for lo in range(0, n, IB):
        sl = slice(lo, min(n, lo + IB))
        g.add_edges(st["src"][sl], st["dst"][sl], st["ts"][sl], st["eid"][sl])
    e1.record()
    torch.cuda.synchronize()
    ingest_ms = e0.elapsed_time(e1)
    peak, src_ = peak_hbm()
    nedges_roots = min(n, targets // 3)
    lo = n - nedges_roots
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    nbatch = (nedges_roots + B.BATCH - 1) // B.BATCH
    neg = torch.randint(0, st["num_nodes"], (nedges_roots,), device=dev, generator=gen)
    # batch b = [src_b || dst_b || neg_b] of its 600 edges
    idx = torch.arange(nedges_roots, device=dev)
    b_of = idx // B.BATCH
    b_lo = b_of * B.BATCH
    b_sz = torch.minimum(torch.full_like(b_of, B.BATCH), nedges_roots - b_lo)
    pos = b_lo * 3 + (idx - b_lo)
    T = nedges_roots * 3
    nodes = torch.empty(T, dtype=torch.int64, device=dev)
    rts = torch.empty(T, dtype=torch.float32, device=dev)
    for k, arr in enumerate((st["src"][lo:], st["dst"][lo:], neg)):
        nodes[pos + k * b_sz] = arr
        rts[pos + k * b_sz] = st["ts"][lo:]
    offs = torch.clamp(torch.arange(nbatch + 1, device=dev, dtype=torch.int64) * (3 * B.BATCH), max=T)
    del idx, b_of, b_lo, b_sz, pos, neg
    nblk = max(1, int(round(g.avg_linked_list_length() * g.num_vertices())))
    mean_block = g.num_edges() / nblk
    log_n = int(np.ceil(np.log2(mean_block + 1)))
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    recs = []
    for strat in strategies:
        smp = TemporalSampler(g, [10, 10], strat)
        out0 = smp.sample_layer_batched(nodes, rts, offs, 0, 0)
        chain = smp.chain_batched(nodes, rts, offs, out0)
        T1 = int(chain[2][-1].item())
        cn1, ct1 = chain[0][:T1], chain[1][:T1]
        out1 = smp.sample_layer_batched(cn1, ct1, chain[2], 1, 0)
        S0, S1 = int(out0["edge_offsets"][-1].item()), int(out1["edge_offsets"][-1].item())
        for _ in range(max(3, warmup)):
            smp.sample_layer_batched(nodes, rts, offs, 0, 0, out=out0)
            smp.chain_batched(nodes, rts, offs, out0, out=chain)
            smp.sample_layer_batched(cn1, ct1, chain[2], 1, 0, out=out1)
        torch.cuda.synchronize()
        ms = [0.0, 0.0, 0.0]
        if os.environ.get("GF_NCU_RANGE"):  # ncu --profile-from-start off: only the timed launches are captured
            torch.cuda.profiler.start()
        for _ in range(steps):
            a, b, c, d = ev(), ev(), ev(), ev()
            a.record(); smp.sample_layer_batched(nodes, rts, offs, 0, 0, out=out0)
            b.record(); smp.chain_batched(nodes, rts, offs, out0, out=chain)
            c.record(); smp.sample_layer_batched(cn1, ct1, chain[2], 1, 0, out=out1)
            d.record(); torch.cuda.synchronize()
            ms[0] += a.elapsed_time(b) / steps; ms[1] += b.elapsed_time(c) / steps; ms[2] += c.elapsed_time(d) / steps
        if os.environ.get("GF_NCU_RANGE"):
            torch.cuda.profiler.stop()
        for layer, (Tl, Sl, dn, m) in enumerate(((T, S0, nodes, ms[0]), (T1, S1, cn1, ms[2]))):
            samp = dn[torch.randint(0, Tl, (200000,), device=dev)].cpu().numpy()
            e_frac = float((g.out_degree(samp) > 0).mean())
            kb = sampling_bytes(Tl, Sl, e_frac, log_n, False)
            recs.append({"strategy": strat, "layer": layer, "kernel": "sample_persistent_kernel", "targets": Tl, "neighbors": Sl,
                         "targets_with_edges_frac": e_frac, "ms_per_launch": m, "algorithmic_bytes": kb,
                         "achieved_GBps": kb / (m * 1e-3) / 1e9, "frac": kb / (m * 1e-3) / 1e9 / peak,
                         "neighbors_per_s": Sl / (m * 1e-3)})
        recs.append({"strategy": strat, "layer": "chain", "kernel": "chain_batched_kernel", "targets": T1, "ms_per_launch": ms[1],
                     "algorithmic_bytes": T1 * 24.0, "achieved_GBps": T1 * 24.0 / (ms[1] * 1e-3) / 1e9,
                     "frac": T1 * 24.0 / (ms[1] * 1e-3) / 1e9 / peak})
        del smp, out0, out1, chain, cn1, ct1
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("hbm_bound:" + shape)
    except Exception:  # noqa: BLE001
        pass
    res = {"shape": shape, "scale": scale, "num_nodes": st["num_nodes"], "edges": n, "graph_payload_bytes": int(g.get_graph_memory_usage()),
           "graph_device_bytes": int(g.get_device_memory_usage()), "mean_block_size": mean_block, "log2_probes": log_n,
           "ingest_edges_per_s": n / (ingest_ms * 1e-3), "ingest_batch": IB, "peak": peak, "peak_source": src_, "unit": "GB/s",
           "batches_per_launch": nbatch, "launches": recs, "dram_traffic": traffic,
           "l2": "graph ({:.1f} GB of payload) >> 126 MB L2; each launch streams > 1 GB of outputs".format(
               g.get_graph_memory_usage() / 1e9)}
    if keep_graph:
        res["_graph"], res["_stream"] = g, st
    else:
        del g, st
        torch.cuda.empty_cache()
    return res


def partitioned_leg(dev, local, rank, world, shape="GDELT-16.7M", scale=1.0, super_batches=64, steps=5, warmup=2,
                    feature_rows=2_000_000, feature_dim=186):
    """BASELINE config 4 on the GPUs of one box: the GDELT-shaped graph hash-partitioned by source vertex over the ranks
    (device-side edge dispatch, gf_dispatch_edges), DySAT sampling ([10,10] uniform, 3 snapshots, prop_time) of
    `super_batches` root batches of 600 edges per rank and exchange step through the peer-memory kernels (gf_peer_*),
    and the partitioned feature rows read over NVLink (gf_gather_rows_partitioned).  In-run checks: the partitioned
    result equals, bit for bit, what the local sampler returns on the UNPARTITIONED graph (built on every rank for that
    purpose); the fetched rows equal the rows.  Also times the same call on the unpartitioned graph of one GPU."""
    import torch
    import torch.distributed as dist
    from gnnflow_b200 import DynamicGraph, TemporalSampler
    from gnnflow_b200.distributed import PartitionedDynamicGraph, PeerFeatureStore, PeerTemporalSampler, owner_of
    st = synth_gpu(shape, scale, dev)
    n = st["n"]
    base = dict(maximum_pool_size=170 << 30, mem_resource_type="cuda", minimum_block_size=st["minimum_block_size"],
                blocks_to_preallocate=1024, insertion_policy="insert")
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    IB = 4_000_000
    gp = PartitionedDynamicGraph(DynamicGraph(initial_pool_size=int(n * 30 / world) + (64 << 20), **base, device=local), rank, world)
    a, b = ev(), ev()
    a.record()
    for lo in range(0, n, IB):
        sl = slice(lo, min(n, lo + IB))
        gp.add_edges(st["src"][sl], st["dst"][sl], st["ts"][sl], st["eid"][sl])
    b.record()
    torch.cuda.synchronize()
    ingest_part_ms = a.elapsed_time(b)
    gfull = DynamicGraph(initial_pool_size=int(n * 30) + (64 << 20), **base, device=local)
    for lo in range(0, n, IB):
        sl = slice(lo, min(n, lo + IB))
        gfull.add_edges(st["src"][sl], st["dst"][sl], st["ts"][sl], st["eid"][sl])
    window = 25.0 if shape == "GDELT-16.7K" else 25.0 * 1000  # ~ the same edges per window at average degree 11
    kw = dict(sample_strategy="uniform", num_snapshots=3, snapshot_time_window=window, prop_time=True)
    fan = [10, 10]
    local_p, local_f = TemporalSampler(gp.graph, fan, **kw), TemporalSampler(gfull, fan, **kw)
    T0 = super_batches * 3 * B.BATCH
    peer = PeerTemporalSampler(local_p, max_targets=T0 * 11 + 64)
    gen = torch.Generator(device=dev)
    gen.manual_seed(100 + rank)
    # every rank samples its own root batches: chunks of 600 edges from the end of the stream, + random negatives
    starts = (n - (torch.arange(super_batches, device=dev) * world + rank + 1) * B.BATCH).clamp_(min=0)
    idx = (starts[:, None] + torch.arange(B.BATCH, device=dev)[None, :])
    neg = torch.randint(0, st["num_nodes"], (super_batches, B.BATCH), device=dev, generator=gen)
    nodes = torch.cat([st["src"][idx], st["dst"][idx], neg], dim=1).reshape(-1).contiguous()
    rts = torch.cat([st["ts"][idx]] * 3, dim=1).reshape(-1).contiguous()
    # ---- in-run equality check (outside the timed region): same launch index -> same Philox draws
    local_p.set_launch_index(5000)
    local_f.set_launch_index(5000)
    got = peer.sample(nodes, rts)
    exp = local_f._sample_results(nodes, rts)
    exp.reverse()
    equal, S = True, 0
    for l in range(2):
        for k in range(3):
            e = exp[l][k].tensors()
            for key in ("all_nodes", "all_timestamps", "delta_timestamps", "eids", "row"):
                x, y = got[l][k][key], e[key]
                same = x.shape == y.shape and bool(torch.equal(x.view(torch.int32) if x.dtype == torch.float32 else x,
                                                               y.view(torch.int32) if y.dtype == torch.float32 else y))
                equal = equal and same
            S += int(got[l][k]["eids"].shape[0])
    T_all = sum(int(got[l][k]["num_dst_nodes"]) for l in range(2) for k in range(3))
    del exp

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        x, y = ev(), ev()
        x.record()
        for _ in range(steps):
            fn()
        y.record()
        torch.cuda.synchronize()
        return x.elapsed_time(y) / steps
    ms_part = timed(lambda: peer.sample(nodes, rts))
    local_p.set_profiling(True)  # a second, untimed pass with CUDA events around the three kernel groups of a step
    local_p.get_profile(True)
    timed(lambda: peer.sample(nodes, rts))
    prof = local_p.get_profile(True)
    local_p.set_profiling(False)
    phases = {k: prof[p_][0] / max(1, prof[p_][1]) for k, p_ in (("route", "locate"), ("wait_and_sample", "scan"), ("wait_and_merge", "emit"))}
    ms_one = timed(lambda: local_f._sample_results(nodes, rts))  # one GPU, whole graph, same roots, same call shape
    # algorithmic NVLink bytes: 16 B / remote target out (this repo's request record), 24 B / neighbour + 4 B /
    # target back; (world - 1) / world of the targets are remote under the hash partition
    xbytes = (T_all * (16 + 4) + S * 24) * (world - 1) / world
    # ---- partitioned feature rows over NVLink
    own = owner_of(torch.arange(feature_rows, device=dev), world).to(torch.int8)
    mine = (own == rank).nonzero().squeeze(1)
    fg = torch.Generator(device=dev)
    fg.manual_seed(3)  # the SAME table on every rank (only this rank's shard is kept)
    rows_all = torch.randn(feature_rows, feature_dim, device=dev, generator=fg)
    fs = PeerFeatureStore(rows_all[mine].contiguous(), own, local)
    ids = torch.randint(0, feature_rows, (1_000_000,), device=dev, generator=gen)
    fetched = fs.fetch(ids)
    feat_equal = bool(torch.equal(fetched, rows_all[ids]))
    del fetched
    ms_feat = timed(lambda: fs.fetch(ids))
    t = torch.tensor([ms_part, ms_one, ms_feat, ingest_part_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(S), float(xbytes), float(equal), float(feat_equal)], dtype=torch.float64, device=dev)
    mn = tot.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    peer.close()
    fs.close()
    res = {"shape": shape, "scale": scale, "edges": n, "num_nodes": st["num_nodes"], "ranks": world,
           "workload": "DySAT [10,10] uniform, 3 snapshots, window {}, prop_time; {} root batches of 600 edges (= {} roots) per "
                       "rank and exchange step".format(window, super_batches, T0),
           "value": float(tot[0]) / (float(t[0]) * 1e-3), "unit": B.UNIT, "ms_per_step": float(t[0]),
           "neighbors_per_step_all_ranks": float(tot[0]),
           "equals_unpartitioned_sampler": bool(mn[2] > 0.5),
           "phase_ms_per_layer_snapshot_step_rank0": phases,
           "one_gpu_unpartitioned_same_call": {"value": S / (float(t[1]) * 1e-3), "unit": B.UNIT, "ms_per_step": float(t[1])},
           "x_one_gpu": float(tot[0]) / (float(t[0]) * 1e-3) / (S / (float(t[1]) * 1e-3)),
           "exchange": {"algorithmic_bytes_per_step_all_ranks": float(tot[1]),
                        "GBps_per_gpu": float(tot[1]) / world / (float(t[0]) * 1e-3) / 1e9, "nvlink_peak_GBps_per_direction": 900,
                        "frac_of_nvlink": float(tot[1]) / world / (float(t[0]) * 1e-3) / 1e9 / 900},
           "ingest": {"edges_per_s_all_ranks": n / (float(t[3]) * 1e-3), "api": "PartitionedDynamicGraph.add_edges(cuda tensors): "
                      "gf_dispatch_edges + add_edges of the owned rows, {}-edge batches".format(IB)},
           "features": {"rows": int(ids.shape[0]), "dim": feature_dim, "ms": float(t[2]), "equals_table": bool(mn[3] > 0.5),
                        "GBps_per_gpu": ids.shape[0] * feature_dim * 4 / (float(t[2]) * 1e-3) / 1e9,
                        "remote_frac": (world - 1) / world},
           "graph_device_bytes": {"partition": int(gp.graph.get_device_memory_usage()), "unpartitioned": int(gfull.get_device_memory_usage())}}
    del gp, gfull, st, rows_all
    torch.cuda.empty_cache()
    return res


def run_partitioned(args):
    import torch.distributed as dist
    rank, world, local, dev = dist_setup()
    if world < 2:
        raise SystemExit("--config partitioned needs torchrun with >= 2 ranks")
    res = partitioned_leg(dev, local, rank, world, args.shape, args.scale, super_batches=args.max_batches if args.max_batches != 400 else 64,
                          steps=args.steps, warmup=args.warmup)
    if rank == 0:
        emit({"metric": B.METRIC, "unit": B.UNIT, "n_gpus": world, "data": "synthetic", "steps": args.steps, "value": res["value"],
              "config": {"workload": res["workload"], "parallelism": "hash-partitioned by source vertex over {} ranks".format(world)},
              "partitioned": res})
    dist.destroy_process_group()


def run_hbm_bound(args):
    import torch  # noqa: F401
    rank, world, local, dev = dist_setup()
    res = hbm_bound_leg(dev, local, args.shape, args.scale, steps=args.steps, warmup=args.warmup)
    best = max(r["frac"] for r in res["launches"] if r["layer"] == 0)
    emit({"metric": B.METRIC, "unit": B.UNIT, "n_gpus": 1, "data": "synthetic", "steps": args.steps,
          "value": max(r["neighbors_per_s"] for r in res["launches"] if r["layer"] == 0), "best_layer0_frac": best,
          "config": {"workload": "{}-shaped at scale {}: saturated multi-batch sampling launches, fan-out 10".format(args.shape, args.scale)},
          "hbm_bound": res})


if __name__ == "__main__":
    a = parse()
    B.quiet_stdout()
    {"wiki": run_wiki, "tgat": run_tgat, "dysat": run_dysat, "online": run_online, "sweep": run_sweep,
     "ingest_sweep": run_ingest_sweep, "two_layer_sat": run_two_layer_sat, "hbm_bound": run_hbm_bound,
     "partitioned": run_partitioned}[a.config](a)
