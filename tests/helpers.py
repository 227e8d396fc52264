"""Shared helpers for the parity tests: synthetic temporal graphs and comparison utilities."""
import numpy as np


def synth_stream(num_src, num_dst, num_edges, seed=42, t_max=2.6e6, zipf=0.8, bipartite=True):
    """Synthetic temporal interaction stream in the shape of SURVEY section 8d: sources ~ Zipf(0.8) over source ids,
    destinations ~ Zipf(0.8) over destination ids (bipartite id ranges), sorted f32 timestamps (ties occur)."""
    rng = np.random.default_rng(seed)

    def zipf_ids(n_ids, n):
        w = 1.0 / np.arange(1, n_ids + 1) ** zipf
        w /= w.sum()
        return rng.choice(n_ids, size=n, p=w).astype(np.int64)

    src = zipf_ids(num_src, num_edges)
    dst = zipf_ids(num_dst, num_edges) + (num_src if bipartite else 0)
    ts = np.sort(rng.uniform(0, t_max, num_edges)).astype(np.float32)
    eid = np.arange(num_edges, dtype=np.int64)
    return src, dst, ts, eid


def bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.int32)
    return a


def assert_same(name, got, exp):
    got, exp = np.asarray(got), np.asarray(exp)
    assert got.shape == exp.shape, "{}: shape {} vs {}".format(name, got.shape, exp.shape)
    if got.dtype != exp.dtype:
        exp = exp.astype(got.dtype)
    if not np.array_equal(bits(got), bits(exp)):
        bad = np.nonzero(bits(got) != bits(exp))[0]
        raise AssertionError("{}: {} of {} entries differ, first at {}: got {} expected {}".format(
            name, len(bad), len(got), bad[0], got[bad[0]], exp[bad[0]]))


def block_to_dict(b):
    """product Block (or DGL block) -> plain python lists, for golden_cases"""
    src, dst = b.edges()
    return dict(ID=b.srcdata['ID'].tolist(), ts=b.srcdata['ts'].tolist(), dt=b.edata['dt'].tolist(),
                eid=b.edata['ID'].tolist(), num_src=b.num_src_nodes(), num_dst=b.num_dst_nodes(),
                src=src.tolist(), dst=dst.tolist())


def compare_block(name, b, o):
    """product Block vs oracle dict, bit-exact"""
    src, dst = b.edges()
    assert b.num_dst_nodes() == o["num_dst_nodes"], name
    assert b.num_src_nodes() == o["num_src_nodes"], "{}: num_src {} vs {}".format(name, b.num_src_nodes(), o["num_src_nodes"])
    assert_same(name + ".all_nodes", b.srcdata['ID'].cpu().numpy(), o["all_nodes"])
    assert_same(name + ".all_ts", b.srcdata['ts'].cpu().numpy(), o["all_timestamps"])
    assert_same(name + ".dt", b.edata['dt'].cpu().numpy(), o["delta_timestamps"])
    assert_same(name + ".eids", b.edata['ID'].cpu().numpy(), o["eids"])
    assert_same(name + ".row", dst.cpu().numpy(), o["row"])
    assert_same(name + ".col", src.cpu().numpy(), o["col"])


def compare_graphs(g, og, vertices):
    assert g.num_edges() == og.num_edges()
    assert g.num_vertices() == og.num_vertices()
    assert g.num_source_vertices() == og.num_source_vertices()
    assert g.max_vertex_id() == og.max_vertex_id()
    assert_same("out_degree", g.out_degree(vertices), og.out_degree(vertices))
    assert_same("nodes", g.nodes(), og.nodes())
    assert_same("src_nodes", g.src_nodes(), og.src_nodes())
    assert g.get_graph_memory_usage() == og.get_graph_memory_usage()
    assert g.get_metadata_memory_usage() == og.get_metadata_memory_usage()
    assert abs(g.avg_linked_list_length() - og.avg_linked_list_length()) < 1e-6
    for v in vertices:
        d, t, e = g.get_temporal_neighbors(int(v))
        od, ot, oe = og.get_temporal_neighbors(int(v))
        assert_same("nbr[%d].dst" % v, d, od)
        assert_same("nbr[%d].ts" % v, t, ot)
        assert_same("nbr[%d].eid" % v, e, oe)
        s, c, a, b = g.block_shapes(int(v))
        os_, oc, oa, ob = og.block_shapes(int(v))
        assert_same("blk[%d].size" % v, s, os_)
        assert_same("blk[%d].cap" % v, c, oc)
        assert_same("blk[%d].start" % v, a, oa)
        assert_same("blk[%d].end" % v, b, ob)
