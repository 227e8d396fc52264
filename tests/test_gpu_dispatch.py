"""Device-side edge dispatch of the partitioned store (gf_dispatch_edges; replaces the host grouping + RPC fan-out of
gnnflow/distributed/dispatcher.py:41-100): each rank keeps, in order, the rows whose source vertex it owns.  Runs on ONE
GPU: the ranks are simulated one after the other, the union of what they keep must be the batch."""
import numpy as np
import pytest
import torch

from helpers import compare_graphs, synth_stream
from oracle.oracle import OracleGraph

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("use_table", [False, True])
def test_dispatch_equals_host_mask(world, use_table):
    from gnnflow_b200 import DynamicGraph
    from gnnflow_b200.distributed import PartitionedDynamicGraph, owner_of, partition_table
    src, dst, ts, eid = synth_stream(700, 300, 50000, seed=23, t_max=5000.0)
    table = None
    if use_table:
        table = partition_table(1000, world)
        table[::7] = -1  # unassigned vertices: nobody stores their edges (dist_sampler.py:223-236)
    cfg = dict(initial_pool_size=4 << 20, maximum_pool_size=256 << 20, mem_resource_type="cuda", minimum_block_size=8,
               blocks_to_preallocate=64, insertion_policy="insert")
    kept_total = 0
    for rank in range(world):
        gd = PartitionedDynamicGraph(DynamicGraph(**cfg), rank, world, table)  # device arrays -> gf_dispatch_edges
        gh = PartitionedDynamicGraph(DynamicGraph(**cfg), rank, world, table)  # host arrays -> numpy mask
        og = OracleGraph(**cfg)
        own = (table.numpy()[src] if use_table else owner_of(src, world)) == rank
        for lo in range(0, len(src), 7000):
            sl = slice(lo, lo + 7000)
            k = gd.add_edges(*[torch.from_numpy(x[sl]).cuda() for x in (src, dst, ts, eid)])
            gh.add_edges(src[sl], dst[sl], ts[sl], eid[sl])
            m = own[sl]
            assert k == int(m.sum())
            if m.any():
                og.add_edges(src[sl][m], dst[sl][m], ts[sl][m], eid[sl][m])
        kept_total += gd.num_edges()
        verts = np.arange(0, 1000, 37)
        compare_graphs(gd.graph, og, verts)
        compare_graphs(gh.graph, og, verts)
    expect = int((table.numpy()[src] >= 0).sum()) if use_table else len(src)
    assert kept_total == expect


def test_dispatch_add_reverse_and_empty():
    from gnnflow_b200 import DynamicGraph
    from gnnflow_b200.distributed import PartitionedDynamicGraph, owner_of
    cfg = dict(initial_pool_size=1 << 20, maximum_pool_size=64 << 20, mem_resource_type="cuda", minimum_block_size=4,
               blocks_to_preallocate=16, insertion_policy="insert")
    src = torch.tensor([1, 2, 3, 4, 5, 6], device="cuda")
    dst = torch.tensor([7, 8, 9, 10, 11, 12], device="cuda")
    ts = torch.arange(6, dtype=torch.float32, device="cuda")
    eid = torch.arange(6, device="cuda")
    total = 0
    for rank in range(3):
        g = PartitionedDynamicGraph(DynamicGraph(**cfg), rank, 3)
        k = g.add_edges(src, dst, ts, eid, add_reverse=True)
        both = np.concatenate([src.cpu().numpy(), dst.cpu().numpy()])
        assert k == int((owner_of(both, 3) == rank).sum())
        total += k
        with pytest.raises(ValueError):
            g.add_edges(src, dst, ts)  # the global edge-id counter is not replicated
    assert total == 12
