"""DynamicGraph with the reference's Python API (gnnflow/dynamic_graph.py:8-204) over the C ABI."""
import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import GF_PTR_DEVICE, GF_PTR_HOST, GraphConfig, check


def _stream_ptr(device: int):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _as_1d(x, dtype_np, dtype_t, device: int, name: str):
    """numpy / list -> contiguous host array; torch -> contiguous tensor (host or this device).
    Returns (keepalive, pointer, kind)."""
    if isinstance(x, torch.Tensor):
        assert x.dim() == 1, "Edges must be 1D tensors"
        if x.is_cuda:
            if x.device.index != device:
                raise ValueError("{} lives on cuda:{}, the graph on cuda:{}".format(name, x.device.index, device))
            t = x if x.dtype is dtype_t and x.is_contiguous() else x.to(dtype_t).contiguous()
            return t, t.data_ptr(), GF_PTR_DEVICE
        x = x.numpy()
    a = np.ascontiguousarray(np.asarray(x), dtype=dtype_np)
    assert a.ndim == 1, "Edges must be 1D tensors"
    return a, C.c_void_p(a.ctypes.data), GF_PTR_HOST


class DynamicGraph:
    """
    A dynamic graph that can be updated at runtime: a vertex table whose entries are time-ordered lists of
    temporal blocks, resident in B200 HBM.  Same constructor and methods as reference
    gnnflow/dynamic_graph.py:17-204.  Arrays may be numpy arrays or torch tensors (CPU or CUDA).
    """

    def __init__(
            self, initial_pool_size: int,
            maximum_pool_size: int,
            mem_resource_type: str,
            minimum_block_size: int,
            blocks_to_preallocate: int,
            insertion_policy: str,
            source_vertices: Optional[np.ndarray] = None,
            target_vertices: Optional[np.ndarray] = None,
            timestamps: Optional[np.ndarray] = None,
            eids: Optional[np.ndarray] = None,
            add_reverse: bool = False,
            device: int = 0,
            adaptive_block_size: bool = True):
        mem = mem_resource_type.lower()
        if mem not in _lib.MEM:
            raise ValueError("Invalid memory resource type: {}".format(mem))
        pol = insertion_policy.lower()
        if pol not in _lib.INSERTION:
            raise ValueError("Invalid insertion policy: {}".format(pol))
        self._L = _lib.lib()
        self._device = int(device)
        cfg = GraphConfig(int(initial_pool_size), int(maximum_pool_size), _lib.MEM[mem], int(minimum_block_size),
                          int(blocks_to_preallocate), _lib.INSERTION[pol], self._device,
                          1 if adaptive_block_size else 0)
        h = C.c_void_p()
        check(self._L.gf_graph_create(C.byref(cfg), C.byref(h)))
        self._h = h
        if source_vertices is not None and target_vertices is not None and timestamps is not None:
            self.add_edges(source_vertices, target_vertices, timestamps, eids, add_reverse)

    def save(self, path: str):
        """Write a checkpoint of the whole graph (not in the reference API)."""
        check(self._L.gf_graph_save(self._h, str(path).encode()))

    @classmethod
    def load(cls, path: str, device: int = 0) -> "DynamicGraph":
        """A new graph from a checkpoint written by `save`: same content, same block structure, same allocator
        state -- it answers every query and continues every stream exactly as the saved graph would have."""
        self = cls.__new__(cls)
        self._L = _lib.lib()
        self._device = int(device)
        h = C.c_void_p()
        check(self._L.gf_graph_load(str(path).encode(), self._device, C.byref(h)))
        self._h = h
        return self

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._L.gf_graph_destroy(h)
            except Exception:
                pass
            self._h = None

    @property
    def device(self) -> int:
        return self._device

    def add_edges(self, source_vertices, target_vertices, timestamps, eids=None, add_reverse: bool = False):
        """gnnflow/dynamic_graph.py:85-126.  Raises ValueError if the timestamps are older than the existing
        edges of a source vertex (the behaviour the reference documents, dynamic_graph.py:99-101)."""
        n = len(source_vertices)
        assert len(source_vertices.shape) == 1 and len(target_vertices.shape) == 1 and \
            len(timestamps.shape) == 1, "Edges must be 1D tensors"
        assert source_vertices.shape[0] == target_vertices.shape[0] == timestamps.shape[0], \
            "The number of source vertices, target vertices, timestamps, and edge ids must be the same."
        on_gpu = isinstance(source_vertices, torch.Tensor) and source_vertices.is_cuda
        if eids is None:
            num_edges = self.num_edges()
            if on_gpu:
                eids = torch.arange(num_edges, num_edges + n, dtype=torch.int64, device=source_vertices.device)
            else:
                eids = np.arange(num_edges, num_edges + n, dtype=np.int64)
        if add_reverse:
            if on_gpu:
                cat = torch.cat
                source_vertices, target_vertices = cat([source_vertices, target_vertices]), \
                    cat([target_vertices, source_vertices])
                timestamps, eids = cat([timestamps, timestamps]), cat([torch.as_tensor(eids, device=timestamps.device)] * 2)
            else:
                s, d = np.asarray(source_vertices), np.asarray(target_vertices)
                source_vertices, target_vertices = np.concatenate([s, d]), np.concatenate([d, s])
                timestamps = np.concatenate([np.asarray(timestamps)] * 2)
                eids = np.concatenate([np.asarray(eids)] * 2)
        ks, ps, kind_s = _as_1d(source_vertices, np.int64, torch.int64, self._device, "source_vertices")
        kd, pd, kind_d = _as_1d(target_vertices, np.int64, torch.int64, self._device, "target_vertices")
        kt, pt, kind_t = _as_1d(timestamps, np.float32, torch.float32, self._device, "timestamps")
        ke, pe, kind_e = _as_1d(eids, np.int64, torch.int64, self._device, "eids")
        if not (kind_s == kind_d == kind_t == kind_e):
            raise ValueError("add_edges: all arrays must be host arrays or all CUDA tensors")
        check(self._L.gf_graph_add_edges(self._h, ps, pd, pt, pe, len(ks), kind_s, _stream_ptr(self._device)))

    def add_edges_async(self, source_vertices: torch.Tensor, target_vertices: torch.Tensor, timestamps: torch.Tensor,
                        eids: torch.Tensor):
        """add_edges without the per-batch host synchronisation (no reference equivalent: DynamicGraph::AddEdges ends in
        cudaStreamSynchronize, dynamic_graph.cu:135-137).  CUDA int64 / float32 tensors only, explicit eids; the tensors
        are kept alive here until the queue is settled -- by `flush()`, or implicitly by the next call that reads or
        changes the graph (sampling included).  Errors (ValueError for out-of-order batches) surface there."""
        ks, ps, kind_s = _as_1d(source_vertices, np.int64, torch.int64, self._device, "source_vertices")
        kd, pd, kind_d = _as_1d(target_vertices, np.int64, torch.int64, self._device, "target_vertices")
        kt, pt, kind_t = _as_1d(timestamps, np.float32, torch.float32, self._device, "timestamps")
        ke, pe, kind_e = _as_1d(eids, np.int64, torch.int64, self._device, "eids")
        if not (kind_s == kind_d == kind_t == kind_e == GF_PTR_DEVICE):
            raise ValueError("add_edges_async: CUDA tensors only")
        if not (len(ks) == len(kd) == len(kt) == len(ke)):
            raise ValueError("add_edges_async: arrays of different lengths")
        if not hasattr(self, "_queued"):
            self._queued = []
        self._queued.append((ks, kd, kt, ke))
        if len(self._queued) > 64:  # the library settles its queue every <= 14 batches; older entries are done
            del self._queued[:-16]
        check(self._L.gf_graph_add_edges_async(self._h, ps, pd, pt, pe, len(ks), _stream_ptr(self._device)))

    def flush(self):
        """Settle the batches queued by add_edges_async (one host synchronisation); raises what add_edges would have."""
        try:
            check(self._L.gf_graph_flush(self._h))
        finally:
            self._queued = []

    def offload_old_blocks(self, timestamp: float, to_file: bool = False):
        out = C.c_uint64()
        check(self._L.gf_graph_offload_old_blocks(self._h, float(timestamp), 1 if to_file else 0, C.byref(out),
                                                  _stream_ptr(self._device)))
        return out.value

    def _u64(self, fn) -> int:
        out = C.c_uint64()
        check(fn(self._h, C.byref(out)))
        return out.value

    def _f32(self, fn) -> float:
        out = C.c_float()
        check(fn(self._h, C.byref(out)))
        return out.value

    def num_vertices(self) -> int:
        return self._u64(self._L.gf_graph_num_vertices)

    def num_source_vertices(self) -> int:
        return self._u64(self._L.gf_graph_num_source_vertices)

    def max_vertex_id(self) -> int:
        out = C.c_int64()
        check(self._L.gf_graph_max_vertex_id(self._h, C.byref(out)))
        return out.value

    def num_edges(self) -> int:
        return self._u64(self._L.gf_graph_num_edges)

    def out_degree(self, vertexs) -> np.ndarray:
        v = np.ascontiguousarray(np.asarray(vertexs), dtype=np.int64)
        out = np.zeros(len(v), dtype=np.uint64)
        check(self._L.gf_graph_out_degree(self._h, C.c_void_p(v.ctypes.data), len(v), C.c_void_p(out.ctypes.data)))
        return out

    def _list(self, fn) -> np.ndarray:
        n = C.c_uint64()
        check(fn(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.int64)
        check(fn(self._h, C.c_void_p(out.ctypes.data), n.value, C.byref(n)))
        return out

    def nodes(self) -> np.ndarray:
        return self._list(self._L.gf_graph_nodes)

    def src_nodes(self) -> np.ndarray:
        return self._list(self._L.gf_graph_src_nodes)

    def edges(self) -> np.ndarray:
        return self._list(self._L.gf_graph_edges)

    def get_temporal_neighbors(self, vertex: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(target_vertices, timestamps, edge_ids) of one vertex, newest first (debug path)."""
        n = C.c_uint64()
        check(self._L.gf_graph_get_temporal_neighbors(self._h, int(vertex), None, None, None, 0, C.byref(n)))
        d, t, e = np.zeros(n.value, np.int64), np.zeros(n.value, np.float32), np.zeros(n.value, np.int64)
        check(self._L.gf_graph_get_temporal_neighbors(self._h, int(vertex), C.c_void_p(d.ctypes.data),
                                                      C.c_void_p(t.ctypes.data), C.c_void_p(e.ctypes.data), n.value,
                                                      C.byref(n)))
        return d, t, e

    def block_shapes(self, vertex: int):
        """(sizes, capacities, start_ts, end_ts) of the vertex's blocks, oldest first (not in the reference API)."""
        n = C.c_uint64()
        check(self._L.gf_graph_block_shapes(self._h, int(vertex), None, None, None, None, 0, C.byref(n)))
        s, c = np.zeros(n.value, np.uint64), np.zeros(n.value, np.uint64)
        a, b = np.zeros(n.value, np.float32), np.zeros(n.value, np.float32)
        check(self._L.gf_graph_block_shapes(self._h, int(vertex), C.c_void_p(s.ctypes.data), C.c_void_p(c.ctypes.data),
                                            C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), n.value, C.byref(n)))
        return s, c, a, b

    def avg_linked_list_length(self) -> float:
        return self._f32(self._L.gf_graph_avg_linked_list_length)

    def get_graph_memory_usage(self) -> int:
        return self._f32(self._L.gf_graph_memory_usage)

    def get_metadata_memory_usage(self) -> int:
        return self._f32(self._L.gf_graph_metadata_memory_usage)

    def clear(self):
        """Empty the graph, keeping its device memory for reuse (not in the reference API)."""
        check(self._L.gf_graph_clear(self._h, _stream_ptr(self._device)))

    def set_profiling(self, on: bool):
        check(self._L.gf_graph_set_profiling(self._h, 1 if on else 0))

    def get_profile(self, reset: bool = True):
        """{phase: (total ms, count)} of add_edges (CUDA events inside the library)"""
        ms = (C.c_double * 5)()
        cnt = (C.c_uint64 * 5)()
        check(self._L.gf_graph_get_profile(self._h, ms, cnt, 1 if reset else 0))
        return {n: (ms[i], cnt[i]) for i, n in enumerate(("prep", "sort", "plan", "realloc_copy", "apply"))}

    def get_device_memory_usage(self) -> int:
        return self._u64(self._L.gf_graph_device_bytes)

    def get_memory_breakdown(self) -> dict:
        """what the store holds in HBM, itemised (not in the reference API)"""
        out = (C.c_uint64 * 8)()
        check(self._L.gf_graph_memory_breakdown(self._h, out))
        return dict(zip(("pool", "bump_used", "free", "vertex_table", "eid_refcounts", "scratch", "allocator_books",
                         "free_blocks"), [int(x) for x in out]))
