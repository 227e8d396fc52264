"""Synthetic temporal interaction streams in the shapes BASELINE.json names (no datasets on disk, no network).
Generator spec: SURVEY.md section 8d."""
import numpy as np

# name -> (num_src, num_dst, num_edges, undirected, minimum_block_size)   [reference gnnflow/config.py:109-167]
SHAPES = {
    "WIKI": (8227, 1000, 157474, True, 18),
    "REDDIT": (10000, 984, 672447, False, 62),
    "GDELT-16.7K": (16682, 0, 191290882, False, 123),
    "GDELT-16.7M": (16700000, 0, 191000000, False, 123),
}


def synth_stream(num_src, num_dst, num_edges, seed=42, t_max=2.6e6, zipf=0.8):
    """sources ~ Zipf(0.8) over [0, num_src), destinations ~ Zipf(0.8) over [num_src, num_src + num_dst) (bipartite;
    num_dst == 0: destinations drawn from the source range), timestamps = sorted U(0, t_max) as f32 (ties occur),
    eid = arange."""
    rng = np.random.default_rng(seed)

    def zipf_ids(n_ids, n):
        w = 1.0 / np.arange(1, n_ids + 1, dtype=np.float64) ** zipf
        cdf = np.cumsum(w)
        cdf /= cdf[-1]
        return np.searchsorted(cdf, rng.random(n), side="right").astype(np.int64).clip(0, n_ids - 1)

    src = zipf_ids(num_src, num_edges)
    dst = zipf_ids(num_dst, num_edges) + num_src if num_dst else zipf_ids(num_src, num_edges)
    ts = np.sort(rng.random(num_edges) * t_max).astype(np.float32)
    eid = np.arange(num_edges, dtype=np.int64)
    return src, dst, ts, eid


def synth(name, seed=42, scale=1.0):
    num_src, num_dst, num_edges, undirected, minblk = SHAPES[name]
    num_edges = int(num_edges * scale)
    src, dst, ts, eid = synth_stream(num_src, num_dst, num_edges, seed)
    return dict(src=src, dst=dst, ts=ts, eid=eid, num_nodes=num_src + num_dst, undirected=undirected,
                minimum_block_size=minblk, name=name)


def tgn_batches(stream, batch_size=600, seed=7):
    """chronological chunks of `batch_size` edges; roots = src || dst || random negatives, ts repeated
    (reference benchmarks/benchmark_sampler.py:70-80).  Returns (nodes, ts, batch_offsets)."""
    rng = np.random.default_rng(seed)
    n = len(stream["src"])
    neg = rng.integers(0, stream["num_nodes"], n).astype(np.int64)
    nodes, tss, offs = [], [], [0]
    for lo in range(0, n, batch_size):
        hi = min(n, lo + batch_size)
        nodes += [stream["src"][lo:hi], stream["dst"][lo:hi], neg[lo:hi]]
        tss += [stream["ts"][lo:hi]] * 3
        offs.append(offs[-1] + 3 * (hi - lo))
    return np.concatenate(nodes), np.concatenate(tss).astype(np.float32), np.asarray(offs, dtype=np.int64)
