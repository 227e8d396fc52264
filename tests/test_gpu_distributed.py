"""Partitioned sampling over NCCL on 2 GPUs of one box (skipped on a single-GPU box): CUDA engine per rank, merged
result bit-identical to the CPU oracle on the unpartitioned graph."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import assert_same, synth_stream

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gnnflow_b200 import DynamicGraph, TemporalSampler
    from gnnflow_b200.distributed import CudaEngine, DistributedTemporalSampler, PartitionedDynamicGraph
    from oracle.oracle import OracleGraph, OracleSampler
    src, dst, ts, eid = synth_stream(500, 80, 40000, seed=5, t_max=4000.0)
    ts = np.floor(ts).astype(np.float32)
    cfg = dict(initial_pool_size=32 << 20, maximum_pool_size=1 << 30, mem_resource_type="cuda", minimum_block_size=8,
               blocks_to_preallocate=1024, insertion_policy="insert")
    pg = PartitionedDynamicGraph(DynamicGraph(**cfg, device=rank), rank, world)
    full = OracleGraph(**cfg)
    for i in range(0, len(src), 5000):
        sl = slice(i, i + 5000)
        pg.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
        full.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
    case = dict(fanouts=[10, 5], sample_strategy="recent")
    ds = DistributedTemporalSampler(CudaEngine(TemporalSampler(pg.graph, **case)), case["fanouts"], 1, rank, world)
    ref = OracleSampler(full, **case)
    rng = np.random.default_rng(100 + rank)
    dev = torch.device("cuda", rank)
    for it in range(3):
        lo = int(rng.integers(0, 39000))
        roots = np.concatenate([src[lo:lo + 600], dst[lo:lo + 600], rng.integers(0, 580, 600)]).astype(np.int64)
        rts = np.concatenate([ts[lo:lo + 600]] * 3).astype(np.float32)
        got = ds.sample(torch.from_numpy(roots).to(dev), torch.from_numpy(rts).to(dev))
        exp = ref.sample(roots, rts)
        for l in range(2):
            for key in ("all_nodes", "all_timestamps", "delta_timestamps", "eids", "row", "col"):
                assert_same("rank%d.it%d.l%d.%s" % (rank, it, l, key), got[l][0][key].cpu().numpy(), exp[l][0][key])
    # ---- the same exchange done by the kernels themselves over NVLink peer memory (gf_peer_*): recent and uniform,
    # snapshots, a partition table with unassigned vertices; bit-identical to the unpartitioned oracle
    from gnnflow_b200.distributed import PeerTemporalSampler, partition_table
    tab = partition_table(580, world)
    for case in (dict(fanouts=[10, 5], sample_strategy="recent"), dict(fanouts=[4, 3], sample_strategy="uniform", seed=7),
                 dict(fanouts=[3, 2], sample_strategy="uniform", num_snapshots=2, snapshot_time_window=300.0, prop_time=True)):
        for use_table in (False, True):
            local = TemporalSampler(pg.graph, **case)
            ps = PeerTemporalSampler(local, max_targets=1800 * 11 + 64, table=tab if use_table else None)
            ref = OracleSampler(full, **case)
            for it in range(3):
                lo = int(rng.integers(0, 39000))
                nr = 600 if (it + rank) % 3 else 200  # ranks ask for different numbers of targets
                roots = np.concatenate([src[lo:lo + nr], dst[lo:lo + nr], rng.integers(0, 580, nr)]).astype(np.int64)
                rts = np.concatenate([ts[lo:lo + nr]] * 3).astype(np.float32)
                got = ps.sample(torch.from_numpy(roots).to(dev), torch.from_numpy(rts).to(dev))
                exp = ref.sample(roots, rts)
                for l in range(2):
                    for k in range(case.get("num_snapshots", 1)):
                        for key in ("all_nodes", "all_timestamps", "delta_timestamps", "eids", "row", "col"):
                            assert_same("peer.%s.tab%d.rank%d.it%d.l%d.s%d.%s" % (case["sample_strategy"], use_table, rank, it, l, k, key),
                                        got[l][k][key].cpu().numpy(), exp[l][k][key])
            ps.close()
    # ---- merge_order="reference": edges in the order of the reference's _merge_sampling_results (dist_sampler.py:276-299)
    from gnnflow_b200.distributed import owner_of
    case = dict(fanouts=[6], sample_strategy="recent")
    ps = PeerTemporalSampler(TemporalSampler(pg.graph, **case), max_targets=4096, merge_order="reference")
    ref = OracleSampler(full, **case)
    lo = 7000 + 500 * rank
    roots = np.concatenate([src[lo:lo + 600], dst[lo:lo + 600], rng.integers(0, 580, 600)]).astype(np.int64)
    rts = np.concatenate([ts[lo:lo + 600]] * 3).astype(np.float32)
    got = ps.sample_layer(torch.from_numpy(roots).to(dev), torch.from_numpy(rts).to(dev), 0, 0)
    own = owner_of(roots, world)
    e_row, e_nbr, e_eid = [], [], []
    for p in range(world):
        idx = np.nonzero(own == p)[0]
        r = ref.sample_layer(roots[idx], rts[idx], 0, 0)
        e_row.append(idx[r["row"]]); e_nbr.append(r["all_nodes"][len(idx):]); e_eid.append(r["eids"])
    assert_same("peer.ref_order.row", got["row"].cpu().numpy(), np.concatenate(e_row))
    assert_same("peer.ref_order.nbr", got["all_nodes"][len(roots):].cpu().numpy(), np.concatenate(e_nbr))
    assert_same("peer.ref_order.eid", got["eids"].cpu().numpy(), np.concatenate(e_eid))
    ps.close()
    # ---- feature rows partitioned over the ranks, remote rows read over NVLink by the gather kernel
    from gnnflow_b200.distributed import PeerFeatureStore
    for N, D in ((5000, 172), (777, 7)):
        feats = np.random.default_rng(9).standard_normal((N, D)).astype(np.float32)  # same on every rank
        own = partition_table(N, world).clone()
        own[::11] = -1  # rows nobody holds come back as zeros
        shard = torch.from_numpy(feats[(own == rank).numpy()])
        fs = PeerFeatureStore(shard, own, rank)
        ids = torch.from_numpy(rng.integers(0, N, 4001)).to(dev)
        got = fs.fetch(ids).cpu().numpy()
        exp = feats[ids.cpu().numpy()].copy()
        exp[(own[ids.cpu()] < 0).numpy()] = 0
        assert_same("peer_features.%d.rank%d" % (D, rank), got.ravel(), exp.ravel())
        assert fs.fetch(ids[:0]).shape == (0, D)
        fs.close()
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_partitioned_sampler_nccl_world2(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / ("ok%d" % r)) for r in range(2))
