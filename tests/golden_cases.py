"""
Golden vectors that the reference's own tests pin for the hot path, restated as backend-agnostic checks.

  store   <- reference tests/test_dynamic_graph.py:26-573
  sampler <- reference tests/test_temporal_sampler.py:27-682

Every function takes factories so that the same vectors run against the CPU oracle (not gpu) and against the
CUDA path through the C-ABI (gpu).  `make_graph(**config)` returns an object with the reference's DynamicGraph
API, `make_sampler(graph, fanouts, **kw)` one with the TemporalSampler API, and `to_dict(block)` converts one
returned block to a dict(ID, ts, dt, eid, num_src, num_dst, src, dst) of python lists / ints.
"""
import numpy as np

MB = 1 << 20
GB = 1 << 30

graph_config = {
    "initial_pool_size": 1 * MB,
    "maximum_pool_size": 2 * MB,
    "mem_resource_type": "cuda",
    "minimum_block_size": 64,
    "blocks_to_preallocate": 128,
    "insertion_policy": "insert",
}

SRC9 = np.array([0, 0, 0, 1, 1, 1, 2, 2, 2])
DST9 = np.array([1, 2, 3, 1, 2, 3, 1, 2, 3])
TS_A = np.array([0, 1, 2, 0, 1, 2, 0, 1, 2])
TS_B = np.array([3, 4, 5, 3, 4, 5, 3, 4, 5])


def _nbrs(g, v):
    d, t, e = g.get_temporal_neighbors(v)
    return list(np.asarray(d).tolist()), list(np.asarray(t).tolist()), list(np.asarray(e).tolist())


def _expect_nbrs(g, v, d, t, e):
    gd, gt, ge = _nbrs(g, v)
    assert gd == d, (v, gd, d)
    assert gt == t, (v, gt, t)
    assert ge == e, (v, ge, e)


def _counts(g, ne, nv, deg):
    assert g.num_edges() == ne
    assert g.num_vertices() == nv
    assert np.asarray(g.out_degree(np.array([0, 1, 2, 3]))).tolist() == deg


# ---------------------------------------------------------------------------------------------- store
def store_sorted(make_graph, **over):  # test_dynamic_graph.py:26-70
    g = make_graph(**{**graph_config, **over})
    g.add_edges(SRC9, DST9, TS_A, add_reverse=False)
    _counts(g, 9, 4, [3, 3, 3, 0])
    _expect_nbrs(g, 0, [3, 2, 1], [2, 1, 0], [2, 1, 0])
    _expect_nbrs(g, 1, [3, 2, 1], [2, 1, 0], [5, 4, 3])
    _expect_nbrs(g, 2, [3, 2, 1], [2, 1, 0], [8, 7, 6])
    _expect_nbrs(g, 3, [], [], [])


def store_add_reverse(make_graph, **over):  # test_dynamic_graph.py:74-118 (stable tie order)
    g = make_graph(**{**graph_config, **over})
    g.add_edges(SRC9, DST9, TS_A, add_reverse=True)
    _counts(g, 9, 4, [3, 6, 6, 3])
    _expect_nbrs(g, 0, [3, 2, 1], [2, 1, 0], [2, 1, 0])
    _expect_nbrs(g, 1, [3, 2, 2, 1, 0, 1], [2, 1, 0, 0, 0, 0], [5, 4, 6, 3, 0, 3])
    _expect_nbrs(g, 2, [3, 2, 1, 0, 2, 1], [2, 1, 1, 1, 1, 0], [8, 7, 4, 1, 7, 6])
    _expect_nbrs(g, 3, [2, 1, 0], [2, 2, 2], [8, 5, 2])


def store_unsorted(make_graph, **over):  # test_dynamic_graph.py:122-163
    g = make_graph(**{**graph_config, **over})
    g.add_edges(SRC9, DST9, np.array([2, 1, 0, 2, 1, 0, 2, 1, 0]), add_reverse=False)
    _counts(g, 9, 4, [3, 3, 3, 0])
    _expect_nbrs(g, 0, [1, 2, 3], [2, 1, 0], [0, 1, 2])
    _expect_nbrs(g, 1, [1, 2, 3], [2, 1, 0], [3, 4, 5])
    _expect_nbrs(g, 2, [1, 2, 3], [2, 1, 0], [6, 7, 8])
    _expect_nbrs(g, 3, [], [], [])


def _two_batches(g, eids_a=None, eids_b=None):
    g.add_edges(SRC9, DST9, TS_A, eids_a, add_reverse=False)
    g.add_edges(SRC9, DST9, TS_B, eids_b, add_reverse=False)


def store_multiple_times(make_graph, policy, **over):  # test_dynamic_graph.py:167-325, 350-403
    g = make_graph(**{**graph_config, "minimum_block_size": 4, "insertion_policy": policy, **over})
    g.add_edges(SRC9, DST9, TS_A, add_reverse=False)
    _counts(g, 9, 4, [3, 3, 3, 0])
    _expect_nbrs(g, 0, [3, 2, 1], [2, 1, 0], [2, 1, 0])
    _expect_nbrs(g, 1, [3, 2, 1], [2, 1, 0], [5, 4, 3])
    _expect_nbrs(g, 2, [3, 2, 1], [2, 1, 0], [8, 7, 6])
    g.add_edges(SRC9, DST9, TS_B, add_reverse=False)
    _counts(g, 18, 4, [6, 6, 6, 0])
    _expect_nbrs(g, 0, [3, 2, 1, 3, 2, 1], [5, 4, 3, 2, 1, 0], [11, 10, 9, 2, 1, 0])
    _expect_nbrs(g, 1, [3, 2, 1, 3, 2, 1], [5, 4, 3, 2, 1, 0], [14, 13, 12, 5, 4, 3])
    _expect_nbrs(g, 2, [3, 2, 1, 3, 2, 1], [5, 4, 3, 2, 1, 0], [17, 16, 15, 8, 7, 6])
    _expect_nbrs(g, 3, [], [], [])
    return g


def store_with_eids(make_graph, **over):  # test_dynamic_graph.py:405-459
    g = make_graph(**{**graph_config, "minimum_block_size": 4, **over})
    _two_batches(g, np.arange(0, 9), np.arange(9, 18))
    _counts(g, 18, 4, [6, 6, 6, 0])
    _expect_nbrs(g, 0, [3, 2, 1, 3, 2, 1], [5, 4, 3, 2, 1, 0], [11, 10, 9, 2, 1, 0])
    _expect_nbrs(g, 1, [3, 2, 1, 3, 2, 1], [5, 4, 3, 2, 1, 0], [14, 13, 12, 5, 4, 3])
    _expect_nbrs(g, 2, [3, 2, 1, 3, 2, 1], [5, 4, 3, 2, 1, 0], [17, 16, 15, 8, 7, 6])


EIDS_NC_A = np.array([0, 2, 4, 6, 8, 10, 12, 14, 16])
EIDS_NC_B = np.array([17, 19, 21, 23, 25, 27, 29, 31, 33])


def store_noncontiguous_eids(make_graph, **over):  # test_dynamic_graph.py:461-515
    g = make_graph(**{**graph_config, "minimum_block_size": 4, **over})
    _two_batches(g, EIDS_NC_A, EIDS_NC_B)
    _counts(g, 18, 4, [6, 6, 6, 0])
    _expect_nbrs(g, 0, [3, 2, 1, 3, 2, 1], [5, 4, 3, 2, 1, 0], [21, 19, 17, 4, 2, 0])
    _expect_nbrs(g, 1, [3, 2, 1, 3, 2, 1], [5, 4, 3, 2, 1, 0], [27, 25, 23, 10, 8, 6])
    _expect_nbrs(g, 2, [3, 2, 1, 3, 2, 1], [5, 4, 3, 2, 1, 0], [33, 31, 29, 16, 14, 12])
    _expect_nbrs(g, 3, [], [], [])


def store_offload(make_graph, **over):  # test_dynamic_graph.py:517-573
    g = make_graph(**{**graph_config, "minimum_block_size": 4, "mem_resource_type": "pinned", **over})
    _two_batches(g, EIDS_NC_A, EIDS_NC_B)
    nb = g.offload_old_blocks(3.5, False)
    assert nb == 3
    assert g.num_edges() == 6
    assert g.num_vertices() == 4
    _expect_nbrs(g, 0, [3, 2], [5, 4], [21, 19])
    _expect_nbrs(g, 1, [3, 2], [5, 4], [27, 25])
    _expect_nbrs(g, 2, [3, 2], [5, 4], [33, 31])
    _expect_nbrs(g, 3, [], [], [])


ALL_STORE = [
    ("sorted", lambda mg: store_sorted(mg)),
    ("add_reverse", lambda mg: store_add_reverse(mg)),
    ("unsorted", lambda mg: store_unsorted(mg)),
    ("multi_insert", lambda mg: store_multiple_times(mg, "insert")),
    ("multi_replace", lambda mg: store_multiple_times(mg, "replace")),
    ("with_eids", lambda mg: store_with_eids(mg)),
    ("noncontiguous_eids", lambda mg: store_noncontiguous_eids(mg)),
    ("offload", lambda mg: store_offload(mg)),
]

# -------------------------------------------------------------------------------------------- sampler
sampler_graph_config = {**graph_config, "initial_pool_size": 1 * GB, "maximum_pool_size": 5 * GB}

SRC18 = np.array([0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2])
DST18 = np.array([1, 2, 3, 4, 5, 6, 1, 2, 3, 4, 5, 6, 1, 2, 3, 4, 5, 6])
TS18 = np.array([0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5])

L1_1P5 = dict(ID=[0, 1, 2, 2, 1, 2, 1, 2, 1], ts=[1.5, 1.5, 1.5, 1, 0, 1, 0, 1, 0], dt=[0.5, 1.5, 0.5, 1.5, 0.5, 1.5],
              eid=[1, 0, 4, 3, 7, 6], num_src=9, num_dst=3, src=[3, 4, 5, 6, 7, 8], dst=[0, 0, 1, 1, 2, 2])
L2_1P5 = dict(ID=[0, 1, 2, 2, 1, 2, 1, 2, 1, 2, 1, 2, 1, 2, 1, 1, 1, 1],
              ts=[1.5, 1.5, 1.5, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0],
              dt=[0.5, 1.5, 0.5, 1.5, 0.5, 1.5, 1, 1, 1], eid=[1, 0, 4, 3, 7, 6, 6, 6, 6], num_src=18, num_dst=9,
              src=[9, 10, 11, 12, 13, 14, 15, 16, 17], dst=[0, 0, 1, 1, 2, 2, 3, 5, 7])
SNAP_45 = dict(ID=[0, 1, 2, 5, 5, 5], ts=[5, 5, 5, 4, 4, 4], dt=[1, 1, 1], eid=[4, 10, 16], num_src=6, num_dst=3,
               src=[3, 4, 5], dst=[0, 1, 2])
SNAP_34 = dict(ID=[0, 1, 2, 4, 4, 4], ts=[5, 5, 5, 3, 3, 3], dt=[2, 2, 2], eid=[3, 9, 15], num_src=6, num_dst=3,
               src=[3, 4, 5], dst=[0, 1, 2])
SNAP2_45 = dict(ID=[0, 1, 2, 5, 5, 5, 5, 5, 5], ts=[5, 5, 5, 4, 4, 4, 4, 4, 4], dt=[1, 1, 1], eid=[4, 10, 16],
                num_src=9, num_dst=6, src=[6, 7, 8], dst=[0, 1, 2])
SNAP2_34 = dict(ID=[0, 1, 2, 4, 4, 4, 4, 4, 4], ts=[5, 5, 5, 3, 3, 3, 3, 3, 3], dt=[2, 2, 2], eid=[3, 9, 15],
                num_src=9, num_dst=6, src=[6, 7, 8], dst=[0, 1, 2])


def _check(block_dict, exp):
    for k, v in exp.items():
        got = block_dict[k]
        assert got == v, (k, got, v)


def _toy(make_graph, **over):
    g = make_graph(**{**sampler_graph_config, **over})
    g.add_edges(SRC9, DST9, TS_A, add_reverse=False)
    return g


def sampler_layer(make_graph, make_sampler, to_dict):  # test_temporal_sampler.py:27-78
    g = _toy(make_graph)
    s = make_sampler(g, [2])
    roots = np.array([0, 1, 2])
    blocks = s.sample(roots, np.array([1.5, 1.5, 1.5]))
    _check(to_dict(blocks[0][0]), L1_1P5)
    _check(to_dict(s.sample_layer(roots, np.array([1.5, 1.5, 1.5]), 0, 0)), L1_1P5)


def sampler_uniform_counts(make_graph, make_sampler, to_dict):  # test_temporal_sampler.py:82-110
    g = _toy(make_graph)
    s = make_sampler(g, [2], sample_strategy="uniform")
    roots = np.array([0, 1, 2])
    b = to_dict(s.sample(roots, np.array([3, 3, 3]))[0][0])
    assert b["num_src"] == 9 and b["num_dst"] == 3
    b = to_dict(s.sample_layer(roots, np.array([3, 3, 3]), 0, 0))
    assert b["num_src"] == 9 and b["num_dst"] == 3


def sampler_multi_blocks(make_graph, make_sampler, to_dict):  # test_temporal_sampler.py:114-172
    g = make_graph(**{**sampler_graph_config, "minimum_block_size": 4})
    g.add_edges(SRC9, DST9, TS_A, add_reverse=False)
    g.add_edges(SRC9, DST9, TS_B, add_reverse=False)
    s = make_sampler(g, [2])
    roots = np.array([0, 1, 2])
    _check(to_dict(s.sample(roots, np.array([1.5, 1.5, 1.5]))[0][0]), L1_1P5)
    _check(to_dict(s.sample_layer(roots, np.array([1.5, 1.5, 1.5]), 0, 0)), L1_1P5)


def sampler_offload(make_graph, make_sampler, to_dict):  # test_temporal_sampler.py:176-238
    g = make_graph(**{**sampler_graph_config, "minimum_block_size": 4, "mem_resource_type": "pinned"})
    g.add_edges(SRC9, DST9, TS_A, add_reverse=False)
    g.add_edges(SRC9, DST9, TS_B, add_reverse=False)
    g.offload_old_blocks(3.5, False)
    s = make_sampler(g, [2])
    roots = np.array([0, 1, 2])
    _check(to_dict(s.sample(roots, np.array([1.5, 1.5, 1.5]))[0][0]),
           dict(ID=[0, 1, 2], ts=[1.5, 1.5, 1.5], dt=[], eid=[], num_src=3, num_dst=3, src=[], dst=[]))
    _check(to_dict(s.sample(roots, np.array([4.5, 4.5, 4.5]))[0][0]),
           dict(ID=[0, 1, 2, 2, 2, 2], ts=[4.5, 4.5, 4.5, 4, 4, 4], dt=[0.5, 0.5, 0.5], eid=[10, 13, 16],
                num_src=6, num_dst=3, src=[3, 4, 5], dst=[0, 1, 2]))


def sampler_duplicates(make_graph, make_sampler, to_dict):  # test_temporal_sampler.py:242-293
    g = _toy(make_graph)
    s = make_sampler(g, [2])
    roots = np.array([0, 1, 2, 0])
    exp = dict(ID=[0, 1, 2, 0, 2, 1, 2, 1, 2, 1, 2, 1], ts=[1.5, 1.5, 1.5, 1.5, 1, 0, 1, 0, 1, 0, 1, 0],
               dt=[0.5, 1.5, 0.5, 1.5, 0.5, 1.5, 0.5, 1.5], eid=[1, 0, 4, 3, 7, 6, 1, 0], num_src=12, num_dst=4,
               src=[4, 5, 6, 7, 8, 9, 10, 11], dst=[0, 0, 1, 1, 2, 2, 3, 3])
    _check(to_dict(s.sample(roots, np.full(4, 1.5))[0][0]), exp)
    _check(to_dict(s.sample_layer(roots, np.full(4, 1.5), 0, 0)), exp)


def sampler_multi_layers(make_graph, make_sampler, to_dict):  # test_temporal_sampler.py:297-386
    g = _toy(make_graph)
    s = make_sampler(g, [2, 2])
    roots = np.array([0, 1, 2])
    blocks = s.sample(roots, np.full(3, 1.5))
    _check(to_dict(blocks[1][0]), L1_1P5)
    _check(to_dict(blocks[0][0]), L2_1P5)
    b = to_dict(s.sample_layer(roots, np.full(3, 1.5), 0, 0))
    _check(b, L1_1P5)
    b2 = to_dict(s.sample_layer(np.array(b["ID"]), np.array(b["ts"], dtype=np.float32), 1, 0))
    _check(b2, L2_1P5)


def _toy18(make_graph):
    g = make_graph(**sampler_graph_config)
    g.add_edges(SRC18, DST18, TS18, add_reverse=False)
    return g


def sampler_multi_snapshots(make_graph, make_sampler, to_dict):  # test_temporal_sampler.py:390-489
    g = _toy18(make_graph)
    s = make_sampler(g, [2], num_snapshots=2, snapshot_time_window=1)
    roots = np.array([0, 1, 2])
    blocks = s.sample(roots, np.array([5, 5, 5]))[0]
    _check(to_dict(blocks[1]), SNAP_45)
    _check(to_dict(blocks[0]), SNAP_34)
    _check(to_dict(s.sample_layer(roots, np.array([5, 5, 5]), 0, 1)), SNAP_45)
    _check(to_dict(s.sample_layer(roots, np.array([5, 5, 5]), 0, 0)), SNAP_34)


def sampler_multi_layers_multi_snapshots(make_graph, make_sampler, to_dict):  # test_temporal_sampler.py:493-656
    g = _toy18(make_graph)
    s = make_sampler(g, [2, 2], num_snapshots=2, snapshot_time_window=1)
    roots = np.array([0, 1, 2])
    blocks = s.sample(roots, np.array([5, 5, 5]))
    _check(to_dict(blocks[1][1]), SNAP_45)
    _check(to_dict(blocks[1][0]), SNAP_34)
    _check(to_dict(blocks[0][1]), SNAP2_45)
    _check(to_dict(blocks[0][0]), SNAP2_34)
    b = to_dict(s.sample_layer(roots, np.array([5, 5, 5]), 0, 1))
    _check(b, SNAP_45)
    _check(to_dict(s.sample_layer(np.array(b["ID"]), np.array(b["ts"], dtype=np.float32), 1, 1)), SNAP2_45)
    b = to_dict(s.sample_layer(roots, np.array([5, 5, 5]), 0, 0))
    _check(b, SNAP_34)
    _check(to_dict(s.sample_layer(np.array(b["ID"]), np.array(b["ts"], dtype=np.float32), 1, 0)), SNAP2_34)


def sampler_batch_sizes(make_graph, make_sampler, to_dict):  # test_temporal_sampler.py:660-682 (incl. empty)
    g = _toy(make_graph)
    s = make_sampler(g, [2])
    rng = np.random.default_rng(0)
    for bs in range(0, 100, 10):
        roots = rng.integers(0, 3, bs)
        ts = rng.integers(0, 3, bs)
        b = to_dict(s.sample(roots, ts)[0][0])
        assert b["num_dst"] == bs and b["ID"][:bs] == roots.tolist()
        b = to_dict(s.sample_layer(roots, ts, 0, 0))
        assert b["num_dst"] == bs


ALL_SAMPLER = [
    ("layer", sampler_layer),
    ("uniform_counts", sampler_uniform_counts),
    ("multi_blocks", sampler_multi_blocks),
    ("offload", sampler_offload),
    ("duplicates", sampler_duplicates),
    ("multi_layers", sampler_multi_layers),
    ("multi_snapshots", sampler_multi_snapshots),
    ("multi_layers_multi_snapshots", sampler_multi_layers_multi_snapshots),
    ("batch_sizes", sampler_batch_sizes),
]
