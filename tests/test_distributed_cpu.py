"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes.  The exchange/merge code of the partitioned
sampler is backend-agnostic; here its local sampling engine is the CPU oracle, and the merged result of every rank
must equal the oracle on the unpartitioned graph bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import assert_same, synth_stream


def test_owner_hash_numpy_equals_torch():
    from gnnflow_b200.distributed import owner_of, partition_table, shard_batch_indices
    v = np.concatenate([np.arange(0, 5000), np.random.default_rng(0).integers(0, 2 ** 40, 5000)]).astype(np.int64)
    for P in (2, 3, 8):
        a, b = owner_of(v, P), owner_of(torch.from_numpy(v), P).numpy()
        assert np.array_equal(a, b)
        assert a.min() >= 0 and a.max() < P
        assert abs(np.bincount(a, minlength=P) / len(v) - 1 / P).max() < 0.03
    assert partition_table(100, 4).dtype == torch.int8
    assert shard_batch_indices(10, 1, 4) == [1, 5, 9]


class OracleEngine:
    def __init__(self, osampler):
        self.s = osampler

    def sample_layer(self, nodes, ts, layer, snapshot):
        r = self.s.sample_layer(nodes.numpy(), ts.numpy(), layer, snapshot)
        T = len(nodes)
        return dict(nbr=torch.from_numpy(r["all_nodes"][T:]), ts=torch.from_numpy(r["all_timestamps"][T:]),
                    dt=torch.from_numpy(r["delta_timestamps"]), eid=torch.from_numpy(r["eids"]),
                    row=torch.from_numpy(r["row"]))


def _worker(rank, world, port, case, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gnnflow_b200.distributed import DistributedTemporalSampler, PartitionedDynamicGraph
    from oracle.oracle import OracleGraph, OracleSampler
    src, dst, ts, eid = synth_stream(150, 30, 12000, seed=5, t_max=1200.0)
    ts = np.floor(ts).astype(np.float32)
    cfg = dict(minimum_block_size=5)
    pg = PartitionedDynamicGraph(OracleGraph(**cfg), rank, world)
    full = OracleGraph(**cfg)
    for i in range(0, len(src), 2000):
        sl = slice(i, i + 2000)
        pg.add_edges(src[sl], dst[sl], ts[sl], eid[sl], add_reverse=True)
        full.add_edges(src[sl], dst[sl], ts[sl], eid[sl], add_reverse=True)
    ds = DistributedTemporalSampler(OracleEngine(OracleSampler(pg.graph, **case)), case["fanouts"],
                                    case.get("num_snapshots", 1), rank, world)
    ref = OracleSampler(full, **case)
    rng = np.random.default_rng(100 + rank)  # every rank samples its own roots
    ok = True
    for it in range(3):
        lo = int(rng.integers(0, 11000))
        roots = np.concatenate([src[lo:lo + 300], dst[lo:lo + 300], rng.integers(0, 400, 300)]).astype(np.int64)
        rts = np.concatenate([ts[lo:lo + 300]] * 3).astype(np.float32)
        got = ds.sample(torch.from_numpy(roots), torch.from_numpy(rts))
        exp = ref.sample(roots, rts)
        for l in range(len(exp)):
            for k in range(len(exp[l])):
                for key in ("all_nodes", "all_timestamps", "delta_timestamps", "eids", "row", "col"):
                    assert_same("rank%d.it%d.l%d.s%d.%s" % (rank, it, l, k, key), got[l][k][key].numpy(), exp[l][k][key])
    assert ds.bytes_sent > 0
    # ---- merge_order="reference": the edge order of the reference's _merge_sampling_results (dist_sampler.py:276-299),
    # restated literally: per partition, in partition order, the owner's result for the targets it owns
    from gnnflow_b200.distributed import owner_of
    dr = DistributedTemporalSampler(OracleEngine(OracleSampler(pg.graph, **case)), case["fanouts"],
                                    case.get("num_snapshots", 1), rank, world, merge_order="reference")
    lo = 5000 + 100 * rank
    roots = np.concatenate([src[lo:lo + 300], dst[lo:lo + 300], rng.integers(0, 400, 300)]).astype(np.int64)
    rts = np.concatenate([ts[lo:lo + 300]] * 3).astype(np.float32)
    for k in range(case.get("num_snapshots", 1)):
        got = dr.sample_layer(torch.from_numpy(roots), torch.from_numpy(rts), 0, k)
        own = owner_of(roots, world)
        e_row, e_nbr, e_ts, e_dt, e_eid = [], [], [], [], []
        for p in range(world):
            idx = np.nonzero(own == p)[0]
            r = ref.sample_layer(roots[idx], rts[idx], 0, k)  # what partition p answers: it holds every out-edge of its vertices
            Tp = len(idx)
            e_row.append(idx[r["row"]]); e_nbr.append(r["all_nodes"][Tp:]); e_ts.append(r["all_timestamps"][Tp:])
            e_dt.append(r["delta_timestamps"]); e_eid.append(r["eids"])
        T = len(roots)
        assert_same("ref_order.row", got["row"].numpy(), np.concatenate(e_row))
        assert_same("ref_order.nbr", got["all_nodes"][T:].numpy(), np.concatenate(e_nbr))
        assert_same("ref_order.nts", got["all_timestamps"][T:].numpy(), np.concatenate(e_ts))
        assert_same("ref_order.dt", got["delta_timestamps"].numpy(), np.concatenate(e_dt))
        assert_same("ref_order.eid", got["eids"].numpy(), np.concatenate(e_eid))
        assert_same("ref_order.dst", got["all_nodes"][:T].numpy(), roots)
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok" if ok else "fail")
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("case", [dict(fanouts=[5, 3], sample_strategy="recent"),
                                  dict(fanouts=[4], sample_strategy="recent", num_snapshots=2, snapshot_time_window=100.0)],
                         ids=["2layer", "snapshots"])
def test_partitioned_sampler_gloo_world2(case, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / ("ok%d" % r)) for r in range(world))
