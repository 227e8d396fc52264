"""Runs the REAL reference extension (oracle/_ref) in its own process -- it calls abort() on any CHECK failure --
and dumps what it computes to an .npz for tests/test_gpu_vs_reference.py.

usage: python ref_runner.py <spec.npz> <out.npz>
spec: src,dst,ts,eid, batch, minblk, adaptive; mode in {store, sample}; for sample: fanouts, policy (0 recent,
1 uniform), num_snapshots, window, prop_time, roots_<i>, rts_<i> (i = 0..nroots-1)"""
import os
import sys

import numpy as np
import torch  # noqa: F401  (libgnnflow links libtorch)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import libgnnflow as ref  # noqa: E402


def fix(a):
    """The reference binds its vectors with the vendored pybind11 (pre-NumPy-2 ABI): under NumPy 2 the arrays it
    returns have the right data pointer, length and dtype but a stride of 0.  Re-stride and copy."""
    a = np.asarray(a)
    if a.ndim == 1 and a.size > 1 and a.strides[0] != a.dtype.itemsize:
        a = np.lib.stride_tricks.as_strided(a, shape=a.shape, strides=(a.dtype.itemsize,))
    return np.array(a)


def main():
    spec = np.load(sys.argv[1])
    src, dst, ts, eid = spec["src"], spec["dst"], spec["ts"], spec["eid"]
    batch, minblk, adaptive = int(spec["batch"]), int(spec["minblk"]), bool(spec["adaptive"])
    mode = str(spec["mode"])
    # OffloadOldBlocks dereferences block payloads on the host and SaveToFile insists on it
    # (dynamic_graph.cu:393, temporal_block_allocator.cu:184-187): host-memory graph for that mode
    mem = ref.MemoryResourceType.PINNED if mode == "offload" else ref.MemoryResourceType.CUDA
    g = ref._DynamicGraph(64 << 20, 1 << 30, mem, minblk, 1024, ref.InsertionPolicy.INSERT, 0, adaptive)
    for i in range(0, len(src), batch):
        sl = slice(i, i + batch)
        g.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
    out = {}
    if mode == "offload":  # .bin files land in the current directory (temporal_block_allocator.cu:189-191)
        out["num_blocks"] = g.offload_old_blocks(float(spec["offload_ts"]), True)
        out["num_edges"] = g.num_edges()
        n = g.max_vertex_id() + 1
        out["out_degree"] = fix(g.out_degree(list(range(n))))
        # a second sweep after more edges: the per-vertex file ordinal keeps counting
        if "src2" in spec:
            g.add_edges(spec["src2"], spec["dst2"], spec["ts2"], spec["eid2"])
            out["num_blocks2"] = g.offload_old_blocks(float(spec["offload_ts2"]), True)
            out["num_edges2"] = g.num_edges()
    elif mode == "store":
        n = g.max_vertex_id() + 1
        out["num_edges"], out["num_vertices"] = g.num_edges(), g.num_vertices()
        out["num_source_vertices"], out["max_vertex_id"] = g.num_source_vertices(), g.max_vertex_id()
        out["out_degree"] = fix(g.out_degree(list(range(n))))
        out["nodes"], out["src_nodes"] = fix(g.nodes()), fix(g.src_nodes())
        out["avg_linked_list_length"] = g.avg_linked_list_length()
        out["graph_mem"], out["meta_mem"] = g.get_graph_memory_usage(), g.get_metadata_memory_usage()
        for v in range(0, n, 3):
            d, t, e = g.get_temporal_neighbors(v)
            out["nbr_dst_%d" % v], out["nbr_ts_%d" % v], out["nbr_eid_%d" % v] = fix(d), fix(t), fix(e)
    else:
        pol = ref.SamplingPolicy.UNIFORM if int(spec["policy"]) else ref.SamplingPolicy.RECENT
        s = ref._TemporalSampler(g, [int(f) for f in spec["fanouts"]], pol, int(spec["num_snapshots"]),
                                 float(spec["window"]), bool(spec["prop_time"]), 1234)
        for i in range(int(spec["nroots"])):
            rr = s.sample(spec["roots_%d" % i], spec["rts_%d" % i])
            for l, layer in enumerate(rr):
                for k, r in enumerate(layer):
                    p = "r%d_l%d_s%d_" % (i, l, k)
                    out[p + "all_nodes"], out[p + "all_ts"] = fix(r.all_nodes()), fix(r.all_timestamps())
                    out[p + "dt"], out[p + "eids"] = fix(r.delta_timestamps()), fix(r.eids())
                    out[p + "row"], out[p + "col"] = fix(r.row()), fix(r.col())
                    out[p + "num_src"], out[p + "num_dst"] = r.num_src_nodes(), r.num_dst_nodes()
    np.savez(sys.argv[2], **out)


if __name__ == "__main__":
    main()
