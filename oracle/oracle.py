"""
TEST INFRASTRUCTURE -- ctypes binding of the CPU oracle (oracle/gnnflow_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
The product package gnnflow_b200 never does.

The two classes restate the reference's Python wrappers on top of the C restatement:
  OracleGraph    <- gnnflow/dynamic_graph.py:8-204
  OracleSampler  <- gnnflow/temporal_sampler.py:14-177 (results as plain numpy dicts instead of DGL blocks)
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgnnflow_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gnnflow_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libgnnflow_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None

_i64p = C.POINTER(C.c_int64)
_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    L.og_graph_create.restype = C.c_void_p
    L.og_graph_create.argtypes = [C.c_uint64, C.c_int, C.c_int]
    L.og_graph_destroy.argtypes = [C.c_void_p]
    L.og_graph_add_edges.restype = C.c_int
    L.og_graph_add_edges.argtypes = [C.c_void_p, _i64p, _i64p, _f32p, _i64p, C.c_size_t]
    L.og_graph_offload_old_blocks.restype = C.c_size_t
    L.og_graph_offload_old_blocks.argtypes = [C.c_void_p, C.c_float]
    for n in ("num_nodes", "num_src_nodes", "num_edges"):
        getattr(L, "og_graph_" + n).restype = C.c_size_t
        getattr(L, "og_graph_" + n).argtypes = [C.c_void_p]
    L.og_graph_max_node_id.restype = C.c_int64
    L.og_graph_max_node_id.argtypes = [C.c_void_p]
    L.og_graph_out_degree.argtypes = [C.c_void_p, _i64p, C.c_size_t, _u64p]
    L.og_graph_get_temporal_neighbors.restype = C.c_size_t
    L.og_graph_get_temporal_neighbors.argtypes = [C.c_void_p, C.c_int64, _i64p, _f32p, _i64p, C.c_size_t]
    for n in ("nodes", "src_nodes", "edges"):
        getattr(L, "og_graph_" + n).restype = C.c_size_t
        getattr(L, "og_graph_" + n).argtypes = [C.c_void_p, _i64p, C.c_size_t]
    for n in ("avg_linked_list_length", "mem_usage", "metadata_mem_usage"):
        getattr(L, "og_graph_" + n).restype = C.c_float
        getattr(L, "og_graph_" + n).argtypes = [C.c_void_p]
    L.og_graph_block_shapes.restype = C.c_size_t
    L.og_graph_block_shapes.argtypes = [C.c_void_p, C.c_int64, _u64p, _u64p, _f32p, _f32p, C.c_size_t]
    L.og_sampler_create.restype = C.c_void_p
    L.og_sampler_create.argtypes = [C.c_void_p, _u32p, C.c_uint32, C.c_int, C.c_uint32, C.c_float, C.c_int,
                                    C.c_uint64]
    L.og_sampler_destroy.argtypes = [C.c_void_p]
    L.og_sampler_launch_index.restype = C.c_uint64
    L.og_sampler_launch_index.argtypes = [C.c_void_p]
    L.og_sampler_set_launch_index.argtypes = [C.c_void_p, C.c_uint64]
    L.og_sampler_sample_layer.restype = C.c_size_t
    L.og_sampler_sample_layer.argtypes = [C.c_void_p, _i64p, _f32p, C.c_size_t, C.c_uint32, C.c_uint32,
                                          _i64p, _i64p, _f32p, _f32p, _i64p, _u32p]
    L.og_philox_u32.restype = C.c_uint32
    L.og_philox_u32.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
    L.og_num_threads.restype = C.c_int
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(t)


def _i64(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.int64)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


class OracleGraph:
    """gnnflow/dynamic_graph.py:8-204 over the C restatement of gnnflow/csrc/dynamic_graph.cu."""

    def __init__(self, initial_pool_size=0, maximum_pool_size=0, mem_resource_type="cuda",
                 minimum_block_size=64, blocks_to_preallocate=0, insertion_policy="insert",
                 source_vertices=None, target_vertices=None, timestamps=None, eids=None,
                 add_reverse=False, device=0, adaptive_block_size=True):
        if mem_resource_type.lower() not in ("cuda", "unified", "pinned", "shared"):
            raise ValueError("Invalid memory resource type: {}".format(mem_resource_type))
        pol = insertion_policy.lower()
        if pol not in ("insert", "replace"):
            raise ValueError("Invalid insertion policy: {}".format(insertion_policy))
        self._L = lib()
        self._h = self._L.og_graph_create(int(minimum_block_size), 0 if pol == "insert" else 1,
                                          1 if adaptive_block_size else 0)
        if source_vertices is not None and target_vertices is not None and timestamps is not None:
            self.add_edges(source_vertices, target_vertices, timestamps, eids, add_reverse)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.og_graph_destroy(self._h)
            self._h = None

    def add_edges(self, source_vertices, target_vertices, timestamps, eids=None, add_reverse=False):
        src, dst, ts = _i64(source_vertices), _i64(target_vertices), _f32(timestamps)
        assert src.ndim == 1 and dst.ndim == 1 and ts.ndim == 1, "Edges must be 1D tensors"
        assert src.shape[0] == dst.shape[0] == ts.shape[0]
        if eids is None:
            n0 = self.num_edges()
            eids = np.arange(n0, n0 + len(src))
        eids = _i64(eids)
        if add_reverse:
            src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
            ts = np.concatenate([ts, ts])
            eids = np.concatenate([eids, eids])
        rc = self._L.og_graph_add_edges(self._h, _p(src, _i64p), _p(dst, _i64p), _p(ts, _f32p),
                                        _p(eids, _i64p), len(src))
        if rc == -2:
            raise ValueError("timestamps are older than the existing edges in the graph")
        if rc != 0:
            raise RuntimeError("og_graph_add_edges failed: {}".format(rc))

    def offload_old_blocks(self, timestamp, to_file=False):
        return self._L.og_graph_offload_old_blocks(self._h, float(timestamp))

    def num_vertices(self):
        return self._L.og_graph_num_nodes(self._h)

    def num_source_vertices(self):
        return self._L.og_graph_num_src_nodes(self._h)

    def num_edges(self):
        return self._L.og_graph_num_edges(self._h)

    def max_vertex_id(self):
        return self._L.og_graph_max_node_id(self._h)

    def out_degree(self, vertexs):
        v = _i64(vertexs)
        out = np.zeros(len(v), dtype=np.uint64)
        self._L.og_graph_out_degree(self._h, _p(v, _i64p), len(v), _p(out, _u64p))
        return out

    def _list(self, fn):
        n = fn(self._h, None, 0)
        out = np.zeros(n, dtype=np.int64)
        fn(self._h, _p(out, _i64p), n)
        return out

    def nodes(self):
        return self._list(self._L.og_graph_nodes)

    def src_nodes(self):
        return self._list(self._L.og_graph_src_nodes)

    def edges(self):
        return self._list(self._L.og_graph_edges)

    def get_temporal_neighbors(self, vertex):
        n = self._L.og_graph_get_temporal_neighbors(self._h, int(vertex), None, None, None, 0)
        d, t, e = np.zeros(n, np.int64), np.zeros(n, np.float32), np.zeros(n, np.int64)
        self._L.og_graph_get_temporal_neighbors(self._h, int(vertex), _p(d, _i64p), _p(t, _f32p), _p(e, _i64p), n)
        return d, t, e

    def avg_linked_list_length(self):
        return self._L.og_graph_avg_linked_list_length(self._h)

    def get_graph_memory_usage(self):
        return self._L.og_graph_mem_usage(self._h)

    def get_metadata_memory_usage(self):
        return self._L.og_graph_metadata_mem_usage(self._h)

    def block_shapes(self, vertex):
        n = self._L.og_graph_block_shapes(self._h, int(vertex), None, None, None, None, 0)
        s, c = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        a, b = np.zeros(n, np.float32), np.zeros(n, np.float32)
        self._L.og_graph_block_shapes(self._h, int(vertex), _p(s, _u64p), _p(c, _u64p), _p(a, _f32p), _p(b, _f32p), n)
        return s, c, a, b


class OracleSampler:
    """gnnflow/temporal_sampler.py:14-177; a 'block' is a dict with the MFG fields (row, col, all_nodes, ...)."""

    def __init__(self, graph, fanouts, sample_strategy="recent", num_snapshots=1,
                 snapshot_time_window=0.0, prop_time=False, seed=1234, *args, **kwargs):
        strat = sample_strategy.lower()
        if strat not in ("recent", "uniform"):
            raise ValueError("strategy must be 'recent' or 'uniform'")
        self._L = lib()
        self._graph = graph
        self._fanouts = [int(f) for f in fanouts]
        fo = np.asarray(self._fanouts, dtype=np.uint32)
        self._h = self._L.og_sampler_create(graph._h, _p(fo, _u32p), len(fo), 0 if strat == "recent" else 1,
                                            int(num_snapshots), float(snapshot_time_window),
                                            1 if prop_time else 0, int(seed))
        self._num_layers = len(fo)
        self._num_snapshots = int(num_snapshots)
        self._is_static = bool(kwargs.get("is_static", False))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.og_sampler_destroy(self._h)
            self._h = None

    def launch_index(self):
        return int(self._L.og_sampler_launch_index(self._h))

    def set_launch_index(self, v):
        """position in the shared counter-based RNG stream (uniform policy)"""
        self._L.og_sampler_set_launch_index(self._h, int(v))

    def sample_layer(self, target_vertices, timestamps, layer, snapshot):
        nodes, ts = _i64(target_vertices), _f32(timestamps)
        T = len(nodes)
        cap = max(1, T * self._fanouts[layer])
        nbr, eid = np.empty(cap, np.int64), np.empty(cap, np.int64)
        ots, dt = np.empty(cap, np.float32), np.empty(cap, np.float32)
        row, cnt = np.empty(cap, np.int64), np.zeros(max(1, T), np.uint32)
        S = self._L.og_sampler_sample_layer(self._h, _p(nodes, _i64p), _p(ts, _f32p), T, layer, snapshot,
                                            _p(nbr, _i64p), _p(eid, _i64p), _p(ots, _f32p), _p(dt, _f32p),
                                            _p(row, _i64p), _p(cnt, _u32p))
        if S == C.c_size_t(-1).value:
            raise ValueError("bad layer/snapshot")
        return {
            "row": row[:S].copy(), "col": np.arange(T, T + S, dtype=np.int64),
            "all_nodes": np.concatenate([nodes, nbr[:S]]), "all_timestamps": np.concatenate([ts, ots[:S]]),
            "delta_timestamps": dt[:S].copy(), "eids": eid[:S].copy(),
            "num_src_nodes": T + S, "num_dst_nodes": T, "num_sampled": cnt[:T].copy(),
        }

    def sample(self, target_vertices, timestamps):
        """temporal_sampler.cu:279-305 + temporal_sampler.py:149-165: [layer][snapshot], layers reversed."""
        nodes, ts = _i64(target_vertices), _f32(timestamps)
        if self._is_static:
            ts = np.full(nodes.shape, np.finfo(np.float32).max, dtype=np.float32)
        results = []
        for layer in range(self._num_layers):
            layer_results = []
            for snapshot in range(self._num_snapshots):
                if layer == 0:
                    n_in, t_in = nodes, ts
                else:
                    prev = results[-1][snapshot]
                    n_in, t_in = prev["all_nodes"], prev["all_timestamps"]
                layer_results.append(self.sample_layer(n_in, t_in, layer, snapshot))
            results.append(layer_results)
        results.reverse()
        return results


def num_threads():
    return lib().og_num_threads()
