"""gf_unique_inverse / the device-resident Memory.prepare_input hand-off against torch.unique on the host, which is
what the reference runs (gnnflow/models/modules/memory.py:170-190)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(ids_np, num_items):
    from gnnflow_b200 import unique_inverse
    ids = torch.from_numpy(ids_np).cuda()
    u, inv = unique_inverse(ids, num_items)
    ru, rinv = torch.unique(torch.from_numpy(ids_np), return_inverse=True)  # the reference's host call
    assert np.array_equal(u.cpu().numpy(), ru.numpy())
    assert np.array_equal(inv.cpu().numpy(), rinv.numpy())
    assert np.array_equal(u[inv].cpu().numpy(), ids_np)


@pytest.mark.parametrize("n,num_items", [(1, 1), (1, 1000), (7, 3), (1800, 10984), (19800, 10984), (200000, 16_700_000),
                                          (1_000_000, 1 << 20), (300, 257), (5000, 256)])
def test_unique_inverse_matches_torch(n, num_items):
    rng = np.random.default_rng(n * 31 + num_items)
    _check(rng.integers(0, num_items, n).astype(np.int64), num_items)


def test_unique_inverse_edge_cases():
    from gnnflow_b200 import unique_inverse
    u, inv = unique_inverse(torch.empty(0, dtype=torch.int64, device="cuda"), 100)
    assert u.numel() == 0 and inv.numel() == 0
    _check(np.zeros(1000, np.int64), 5)                       # one distinct id
    _check(np.arange(4096, dtype=np.int64)[::-1].copy(), 4096)  # all distinct, descending input
    _check(np.array([255, 256, 257, 0, 31, 32, 255], np.int64), 300)  # word / chunk boundaries
    ids = torch.tensor([3, 1, 3], device="cuda")
    u, inv = unique_inverse(ids)                              # id space inferred from the ids
    assert u.tolist() == [1, 3] and inv.tolist() == [1, 0, 1]
    with pytest.raises(ValueError):
        unique_inverse(torch.tensor([1, 2]), 10)              # CPU tensor: no fallback
    # repeated calls reuse the scratch (stale bits would corrupt the ranks)
    for seed in range(5):
        rng = np.random.default_rng(seed)
        _check(rng.integers(0, 5000, 3000).astype(np.int64), 5000)


def test_prepare_memory_input_matches_reference_flow():
    from gnnflow_b200 import DynamicGraph, TemporalSampler, prepare_memory_input
    rng = np.random.default_rng(3)
    n, N = 20000, 600
    src = rng.integers(0, 500, n).astype(np.int64)
    dst = rng.integers(500, N, n).astype(np.int64)
    ts = np.sort(rng.uniform(0, 1000, n)).astype(np.float32)
    g = DynamicGraph(initial_pool_size=16 << 20, maximum_pool_size=1 << 30, mem_resource_type="cuda",
                     minimum_block_size=16, blocks_to_preallocate=1024, insertion_policy="insert")
    g.add_edges(src, dst, ts)
    b = TemporalSampler(g, [10], "recent").sample(np.concatenate([src[-600:], dst[-600:]]),
                                                  np.concatenate([ts[-600:]] * 2))[0][0]
    gen = torch.Generator().manual_seed(0)
    mem, mem_ts = torch.randn(N, 100, generator=gen), torch.rand(N, generator=gen)
    mail, mail_ts = torch.randn(N, 1, 372, generator=gen), torch.rand(N, 1, generator=gen)
    prepare_memory_input(b, mem.cuda(), mem_ts.cuda(), mail.cuda(), mail_ts.cuda())
    # the reference's flow on the host (memory.py:170-190)
    all_nodes = b.srcdata['ID'].cpu()
    uniq, inv = torch.unique(all_nodes, return_inverse=True)
    assert torch.equal(b.srcdata['mem'].cpu(), mem[uniq][inv])
    assert torch.equal(b.srcdata['mem_ts'].cpu(), mem_ts[uniq][inv])
    assert torch.equal(b.srcdata['mail_ts'].cpu(), mail_ts[uniq][inv])
    assert torch.equal(b.srcdata['mem_input'].cpu(), mail[uniq][inv])
