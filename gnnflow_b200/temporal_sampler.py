"""TemporalSampler with the reference's Python API (gnnflow/temporal_sampler.py:14-177) over the C ABI.

The message-flow graphs are assembled on the GPU and stay there: `mfgs_to_cuda` (reference utils.py:477-481)
becomes a no-op because `Block.to(device)` returns self when already on that device.  When `dgl` is importable
a real `dgl.create_block` is built from the CUDA tensors; otherwise the duck-typed `Block` below carries the same
fields the reference models read (srcdata['ID','ts','h'], edata['dt','ID','f'], edges(), num_*_nodes())."""
import ctypes as C
import os
import time
from typing import List, Union

import numpy as np
import torch

from . import _lib
from ._lib import GF_PTR_DEVICE, GF_PTR_HOST, SamplingResultC, check
from .dynamic_graph import DynamicGraph, _stream_ptr

try:  # pragma: no cover - dgl is not installed in the build image
    if os.environ.get("GNNFLOW_B200_MFG", "").lower() == "native":
        raise ImportError
    import dgl
    _HAS_DGL = True
except Exception:  # noqa: BLE001
    dgl = None
    _HAS_DGL = False


class Block:
    """Minimal stand-in for dgl.heterograph.DGLBlock (bipartite src -> dst message-flow graph)."""

    def __init__(self, col: torch.Tensor, row: torch.Tensor, num_src_nodes: int, num_dst_nodes: int):
        self._col, self._row = col, row
        self._num_src, self._num_dst = int(num_src_nodes), int(num_dst_nodes)
        self.srcdata, self.dstdata, self.edata = {}, {}, {}

    def num_src_nodes(self) -> int:
        return self._num_src

    def num_dst_nodes(self) -> int:
        return self._num_dst

    def num_edges(self) -> int:
        return int(self._row.shape[0])

    def edges(self):
        """(source ids, destination ids) == (col, row) of dgl.create_block((col, row))."""
        return self._col, self._row

    @property
    def device(self):
        return self._row.device

    def to(self, device, **kwargs):
        device = torch.device(device)
        if device == self.device or (device.type == "cuda" and device.index is None and self.device.type == "cuda"):
            return self
        b = Block(self._col.to(device), self._row.to(device), self._num_src, self._num_dst)
        for name in ("srcdata", "dstdata", "edata"):
            getattr(b, name).update({k: v.to(device) for k, v in getattr(self, name).items()})
        return b


class SamplingResult:
    """The reference's pybind SamplingResult (gnnflow/csrc/api.cc:87-109): getters return host numpy arrays."""

    def __init__(self, all_nodes, all_ts, dt, eids, row, col, num_dst):
        self._t = dict(all_nodes=all_nodes, all_timestamps=all_ts, delta_timestamps=dt, eids=eids, row=row, col=col)
        self._num_dst = int(num_dst)

    def _np(self, k):
        return self._t[k].detach().cpu().numpy()

    def row(self): return self._np("row")
    def col(self): return self._np("col")
    def all_nodes(self): return self._np("all_nodes")
    def all_timestamps(self): return self._np("all_timestamps")
    def delta_timestamps(self): return self._np("delta_timestamps")
    def eids(self): return self._np("eids")
    def num_dst_nodes(self): return self._num_dst
    def num_src_nodes(self): return int(self._t["all_nodes"].shape[0])

    def tensors(self):
        """device tensors (not in the reference API)"""
        return self._t


class TemporalSampler:
    """Samples k-hop multi-snapshot temporal neighbours of given vertices (gnnflow/temporal_sampler.py:14-60)."""

    def __init__(self, graph: DynamicGraph, fanouts: List[int], sample_strategy: str = "recent",
                 num_snapshots: int = 1, snapshot_time_window: float = 0.0, prop_time: bool = False,
                 seed: int = 1234, *args, **kwargs):
        sample_strategy = sample_strategy.lower()
        if sample_strategy not in ["recent", "uniform"]:
            raise ValueError("strategy must be 'recent' or 'uniform'")
        self._L = _lib.lib()
        self._graph = graph  # the graph must outlive the sampler (temporal_sampler.h:62)
        self._device = graph.device
        self._fanouts = [int(f) for f in fanouts]
        self._num_layers = len(self._fanouts)
        self._num_snapshots = int(num_snapshots)
        fo = (C.c_uint32 * self._num_layers)(*self._fanouts)
        h = C.c_void_p()
        check(self._L.gf_sampler_create(graph._h, fo, self._num_layers, _lib.SAMPLING[sample_strategy],
                                        self._num_snapshots, float(snapshot_time_window), 1 if prop_time else 0,
                                        int(seed), C.byref(h)))
        self._h = h
        self._is_static = 'is_static' in kwargs and kwargs['is_static'] is True
        self._use_dgl = _HAS_DGL

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._L.gf_sampler_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------------------------------ plumbing
    def _inputs(self, target_vertices, timestamps):
        """-> (keepalive objects, nodes ptr, ts ptr, T, kind)"""
        dev = self._device
        if isinstance(target_vertices, torch.Tensor) and target_vertices.is_cuda:
            n = self._dev(target_vertices, torch.int64, "target_vertices")
            if isinstance(timestamps, torch.Tensor) and timestamps.is_cuda:
                t = self._dev(timestamps, torch.float32, "timestamps")
            else:
                t = torch.as_tensor(timestamps, device=n.device).to(torch.float32).contiguous()
            if self._is_static:
                t = torch.full_like(t, float(np.finfo(np.float32).max))
            return (n, t), C.c_void_p(n.data_ptr()), C.c_void_p(t.data_ptr()), n.shape[0], GF_PTR_DEVICE
        if isinstance(target_vertices, torch.Tensor):
            target_vertices = target_vertices.numpy()
        if isinstance(timestamps, torch.Tensor):
            timestamps = timestamps.cpu().numpy()
        n = np.ascontiguousarray(np.asarray(target_vertices), dtype=np.int64)
        if self._is_static:  # temporal_sampler.py:72-76
            t = np.full(n.shape, np.finfo(np.float32).max, dtype=np.float32)
        else:
            t = np.ascontiguousarray(np.asarray(timestamps), dtype=np.float32)
        assert n.ndim == 1 and t.shape == n.shape
        del dev
        return (n, t), C.c_void_p(n.ctypes.data), C.c_void_p(t.ctypes.data), n.shape[0], GF_PTR_HOST

    def _dev(self, t, dtype, name):
        """CUDA tensor on the graph's device, of `dtype`, contiguous (a wrong dtype or a tensor on another GPU would
        otherwise be silently misread through its raw pointer)"""
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise ValueError("{} must be a CUDA tensor".format(name))
        if t.device.index != self._device:
            raise ValueError("{} lives on cuda:{}, the graph on cuda:{}".format(name, t.device.index, self._device))
        return t if t.dtype is dtype and t.is_contiguous() else t.to(dtype).contiguous()

    def _alloc_steps(self, caps):
        """Output arrays of several (layer, snapshot) steps carved out of ONE device allocation (a torch.empty per
        array costs more than the kernel at the reference's batch sizes).  caps: [(cap_dst, fanout)] per step.
        -> (pool, typed views of the pool, per-step element offsets, ctypes array of gf_sampling_result); the layout
        and the ctypes array are kept for the next call with the same capacities."""
        key = tuple(caps)
        plan = self._plan if getattr(self, "_plan", None) is not None and self._plan[0] == key else None
        if plan is None:
            A = 256
            offs, total = [], 0
            for cap_dst, fanout in caps:
                cap_e = cap_dst * fanout
                sizes = ((cap_dst + cap_e) * 8, (cap_dst + cap_e) * 4, cap_e * 4, cap_e * 8, cap_e * 8, cap_e * 8)
                o = []
                for sz in sizes:
                    o.append(total)
                    total += (sz + A - 1) // A * A
                offs.append(o)
            plan = self._plan = (key, offs, max(total, A), (SamplingResultC * len(caps))())
        _, offs, total, arr = plan
        pool = torch.empty(total, dtype=torch.uint8, device=torch.device("cuda", self._device))
        base = pool.data_ptr()
        for i, ((cap_dst, _f), o) in enumerate(zip(caps, offs)):
            r = arr[i]
            r.all_nodes, r.all_timestamps, r.delta_timestamps = base + o[0], base + o[1], base + o[2]
            r.eids, r.row, r.col = base + o[3], base + o[4], base + o[5]
            r.capacity_dst, r.num_dst, r.num_edges = cap_dst, 0, 0
        return pool, offs, arr

    @staticmethod
    def _views(p64, p32, o, T, S):
        """the six arrays of one step as views of the pool (p64 / p32: the pool seen as int64 / float32)"""
        a, b, c, d, e, f = o[0] >> 3, o[1] >> 2, o[2] >> 2, o[3] >> 3, o[4] >> 3, o[5] >> 3
        return (p64[a:a + T + S], p32[b:b + T + S], p32[c:c + S], p64[d:d + S], p64[e:e + S], p64[f:f + S])

    def _sample_steps(self, target_vertices, timestamps):
        """-> [layer][snapshot] of (all_nodes, all_ts, dt, eids, row, col, num_dst): device views, one C call"""
        keep, pn, pt, T, kind = self._inputs(target_vertices, timestamps)
        caps, cap = [], T
        for layer in range(self._num_layers):
            caps += [(cap, self._fanouts[layer])] * self._num_snapshots
            cap = cap * (1 + self._fanouts[layer])
        pool, offs, arr = self._alloc_steps(caps)
        check(self._L.gf_sampler_sample(self._h, pn, pt, T, arr, kind, GF_PTR_DEVICE, _stream_ptr(self._device)))
        del keep
        p64, p32 = pool.view(torch.int64), pool.view(torch.float32)
        out, i = [], 0
        for layer in range(self._num_layers):
            lay = []
            for _s in range(self._num_snapshots):
                r = arr[i]
                Td = r.num_dst
                lay.append(self._views(p64, p32, offs[i], Td, r.num_edges) + (Td,))
                i += 1
            out.append(lay)
        return out

    def _sample_results(self, target_vertices, timestamps) -> List[List[SamplingResult]]:
        return [[SamplingResult(*v) for v in lay] for lay in self._sample_steps(target_vertices, timestamps)]

    def _sample_layer_result(self, target_vertices, timestamps, layer: int, snapshot: int) -> SamplingResult:
        if not 0 <= layer < self._num_layers:
            raise ValueError("layer out of range")
        keep, pn, pt, T, kind = self._inputs(target_vertices, timestamps)
        pool, offs, arr = self._alloc_steps([(T, self._fanouts[layer])])
        check(self._L.gf_sampler_sample_layer(self._h, pn, pt, T, layer, snapshot, arr, kind, GF_PTR_DEVICE,
                                              _stream_ptr(self._device)))
        del keep
        r = arr[0]
        return SamplingResult(*(self._views(pool.view(torch.int64), pool.view(torch.float32), offs[0], r.num_dst,
                                            r.num_edges) + (r.num_dst,)))

    # ------------------------------------------------------------------------------------------ public API
    def sample(self, target_vertices: np.ndarray, timestamps: np.ndarray) -> List[List[Block]]:
        """Sample k-hop neighbours; returns [layer][snapshot] blocks with mfgs[0] the outermost hop
        (gnnflow/temporal_sampler.py:60-80,149-165)."""
        steps = self._sample_steps(target_vertices, timestamps)
        if self._use_dgl:  # pragma: no cover
            return self._to_dgl_block([[SamplingResult(*v) for v in lay] for lay in steps])
        mfgs = [[self._block(*v) for v in lay] for lay in steps]
        mfgs.reverse()  # temporal_sampler.py:163-164
        return mfgs

    @staticmethod
    def _block(all_nodes, all_ts, dt, eids, row, col, num_dst) -> Block:
        b = Block(col, row, all_nodes.shape[0], num_dst)
        b.srcdata['ID'] = all_nodes
        b.edata['dt'] = dt
        b.srcdata['ts'] = all_ts
        b.edata['ID'] = eids
        return b

    def _sample(self, target_vertices: np.ndarray, timestamps: np.ndarray, sort: bool = False):
        """debug only (gnnflow/temporal_sampler.py:82-125, used by benchmarks/benchmark_sampler.py:83)"""
        sampling_results = []
        sort_time = 0
        for layer in range(self._num_layers):
            layer_results = []
            for snapshot in range(self._num_snapshots):
                if layer == 0:
                    input_nodes, input_ts = target_vertices, timestamps
                else:
                    t = sampling_results[layer - 1][snapshot].tensors()
                    input_nodes, input_ts = t["all_nodes"], t["all_timestamps"]
                if sort:
                    sort_start = time.time()
                    if isinstance(input_nodes, torch.Tensor):
                        idx = torch.argsort(input_nodes, stable=True)
                    else:
                        idx = np.argsort(input_nodes, kind="stable")
                    input_nodes, input_ts = input_nodes[idx], input_ts[idx]
                    sort_time += time.time() - sort_start
                layer_results.append(self._sample_layer_result(input_nodes, input_ts, layer, snapshot))
            sampling_results.append(layer_results)
        return sampling_results, sort_time

    def sample_layer(self, target_vertices: np.ndarray, timestamps: np.ndarray, layer: int, snapshot: int,
                     to_dgl_block: bool = True) -> Union[Block, SamplingResult]:
        """gnnflow/temporal_sampler.py:127-147"""
        r = self._sample_layer_result(target_vertices, timestamps, layer, snapshot)
        if to_dgl_block:
            return self._to_dgl_block_layer_snapshot(r)
        return r

    def _to_dgl_block(self, sampling_results) -> List[List[Block]]:
        mfgs = [[self._to_dgl_block_layer_snapshot(r) for r in layer] for layer in sampling_results]
        mfgs.reverse()  # temporal_sampler.py:163-164
        return mfgs

    def _to_dgl_block_layer_snapshot(self, r: SamplingResult):
        t = r.tensors()
        if self._use_dgl:  # pragma: no cover
            b = dgl.create_block((t["col"], t["row"]), num_src_nodes=r.num_src_nodes(),
                                 num_dst_nodes=r.num_dst_nodes())
        else:
            b = Block(t["col"], t["row"], r.num_src_nodes(), r.num_dst_nodes())
        b.srcdata['ID'] = t["all_nodes"]
        b.edata['dt'] = t["delta_timestamps"]
        b.srcdata['ts'] = t["all_timestamps"]
        b.edata['ID'] = t["eids"]
        return b

    # ------------------------------------------------------------------------------------------ extras
    def sample_layer_batched(self, nodes: torch.Tensor, timestamps: torch.Tensor, batch_offsets: torch.Tensor,
                             layer: int = 0, snapshot: int = 0, out=None):
        """Many independent root batches in one launch (include/gnnflow_b200.h: gf_sampler_sample_layer_batched).
        All tensors on the GPU; returns dict(nbr, ts, dt, eid, row, edge_offsets)."""
        nodes = self._dev(nodes, torch.int64, "nodes")
        timestamps = self._dev(timestamps, torch.float32, "timestamps")
        dev = nodes.device
        T = nodes.shape[0]
        F = self._fanouts[layer]
        nb = batch_offsets.shape[0] - 1
        if timestamps.shape[0] != T:
            raise ValueError("nodes and timestamps differ in length")
        if out is None:
            out = dict(nbr=torch.empty(T * F, dtype=torch.int64, device=dev),
                       ts=torch.empty(T * F, dtype=torch.float32, device=dev),
                       dt=torch.empty(T * F, dtype=torch.float32, device=dev),
                       eid=torch.empty(T * F, dtype=torch.int64, device=dev),
                       row=torch.empty(T * F, dtype=torch.int64, device=dev),
                       edge_offsets=torch.empty(nb + 1, dtype=torch.int64, device=dev))
        bo = self._dev(batch_offsets, torch.int64, "batch_offsets")
        check(self._L.gf_sampler_sample_layer_batched(
            self._h, nodes.data_ptr(), timestamps.data_ptr(), T, bo.data_ptr(), nb, layer, snapshot,
            out["nbr"].data_ptr(), out["ts"].data_ptr(), out["dt"].data_ptr(), out["eid"].data_ptr(),
            out["row"].data_ptr(), out["edge_offsets"].data_ptr(), GF_PTR_DEVICE, _stream_ptr(self._device)))
        return out

    def chain_batched(self, nodes: torch.Tensor, timestamps: torch.Tensor, batch_offsets: torch.Tensor, sampled: dict,
                      out=None):
        """Targets of the next layer of a multi-batch launch, built on the device (gf_sampler_chain_batched): per batch
        [roots || neighbours] with the neighbours' timestamps.  `sampled` is what `sample_layer_batched` returned for
        (nodes, timestamps, batch_offsets).  Returns (nodes_next, timestamps_next, batch_offsets_next); the first
        batch_offsets_next[-1] entries of the two arrays are valid (no host synchronisation here)."""
        nodes = self._dev(nodes, torch.int64, "nodes")
        timestamps = self._dev(timestamps, torch.float32, "timestamps")
        dev = nodes.device
        T, nb = nodes.shape[0], batch_offsets.shape[0] - 1
        cap = sampled["nbr"].shape[0]
        if out is None:
            out = (torch.empty(T + cap, dtype=torch.int64, device=dev), torch.empty(T + cap, dtype=torch.float32, device=dev),
                   torch.empty(nb + 1, dtype=torch.int64, device=dev))
        bo = self._dev(batch_offsets, torch.int64, "batch_offsets")
        check(self._L.gf_sampler_chain_batched(
            nodes.data_ptr(), timestamps.data_ptr(), T, bo.data_ptr(), nb, sampled["nbr"].data_ptr(),
            sampled["ts"].data_ptr(), sampled["edge_offsets"].data_ptr(), cap, out[0].data_ptr(), out[1].data_ptr(),
            out[2].data_ptr(), _stream_ptr(self._device)))
        return out

    def sample_batched(self, nodes: torch.Tensor, timestamps: torch.Tensor, batch_offsets: torch.Tensor):
        """Every layer of many independent root batches (1 snapshot): one sampling launch + one chaining launch per
        layer, one host synchronisation per layer boundary to size the next layer.  Returns a list (layer 0 first) of
        dict(nodes, timestamps, batch_offsets, nbr, ts, dt, eid, row, edge_offsets); batch b of a layer owns targets
        [batch_offsets[b], batch_offsets[b+1]) and edges [edge_offsets[b], edge_offsets[b+1]), and its result equals
        what `sample` returns for that batch alone."""
        layers = []
        cur = (self._dev(nodes, torch.int64, "nodes"), self._dev(timestamps, torch.float32, "timestamps"),
               self._dev(batch_offsets, torch.int64, "batch_offsets"))
        for layer in range(len(self._fanouts)):
            smp = self.sample_layer_batched(cur[0], cur[1], cur[2], layer, 0)
            layers.append(dict(nodes=cur[0], timestamps=cur[1], batch_offsets=cur[2], **smp))
            if layer + 1 < len(self._fanouts):
                nxt = self.chain_batched(cur[0], cur[1], cur[2], smp)
                total = int(nxt[2][-1].item())
                cur = (nxt[0][:total], nxt[1][:total], nxt[2])
        return layers

    def sample_layer_batched_numpy(self, nodes: np.ndarray, timestamps: np.ndarray, batch_offsets: np.ndarray,
                                   layer: int = 0, snapshot: int = 0, out=None, ids32: bool = False):
        """Host in, host out version of `sample_layer_batched`: one C-ABI call samples every batch of a replay
        (gf_sampler_sample_layer_batched with GF_PTR_HOST).  `out` (optional) is a dict of host arrays as returned
        by `alloc_batched_host_out`; pinned arrays are written in place by the kernel.  Returns
        dict(nbr, ts, dt, eid, row, edge_offsets) of numpy arrays; batch b owns [edge_offsets[b], edge_offsets[b+1]).
        ids32=True: `nbr` and `row` come back as uint32 (gf_sampler_sample_layer_batched_ids32: 24 instead of 32
        bytes per neighbour over PCIe, the bound of this call); same values."""
        n = np.ascontiguousarray(nodes, dtype=np.int64)
        t = np.ascontiguousarray(timestamps, dtype=np.float32)
        bo = np.ascontiguousarray(batch_offsets, dtype=np.uint64)
        T, nb = n.shape[0], bo.shape[0] - 1
        assert t.shape == n.shape and nb >= 1
        if out is None:
            out = self.alloc_batched_host_out(T, nb, layer, ids32=ids32)
        F = self._fanouts[layer]
        idt = np.uint32 if ids32 else np.int64
        for k, dt_, cnt in (("nbr", idt, T * F), ("ts", np.float32, T * F), ("dt", np.float32, T * F),
                            ("eid", np.int64, T * F), ("row", idt, T * F), ("edge_offsets", np.uint64, nb + 1)):
            a = out[k]
            if a.dtype != dt_ or a.shape[0] < cnt or not a.flags.c_contiguous:
                raise ValueError("out[%r] must be a contiguous %s array with >= %d elements" % (k, np.dtype(dt_), cnt))
        if ids32:
            check(self._L.gf_sampler_sample_layer_batched_ids32(
                self._h, n.ctypes.data, t.ctypes.data, T, bo.ctypes.data, nb, layer, snapshot,
                out["nbr"].ctypes.data, out["ts"].ctypes.data, out["dt"].ctypes.data, out["eid"].ctypes.data,
                out["row"].ctypes.data, out["edge_offsets"].ctypes.data, _stream_ptr(self._device)))
        else:
            check(self._L.gf_sampler_sample_layer_batched(
                self._h, n.ctypes.data, t.ctypes.data, T, bo.ctypes.data, nb, layer, snapshot,
                out["nbr"].ctypes.data, out["ts"].ctypes.data, out["dt"].ctypes.data, out["eid"].ctypes.data,
                out["row"].ctypes.data, out["edge_offsets"].ctypes.data, GF_PTR_HOST, _stream_ptr(self._device)))
        S = int(out["edge_offsets"][nb])
        return dict(nbr=out["nbr"][:S], ts=out["ts"][:S], dt=out["dt"][:S], eid=out["eid"][:S], row=out["row"][:S],
                    edge_offsets=out["edge_offsets"][:nb + 1])

    def alloc_batched_host_out(self, num_targets: int, num_batches: int, layer: int = 0, pinned: bool = True,
                               ids32: bool = False):
        """Host output arrays for `sample_layer_batched_numpy` (pinned: written in place by the kernel; ids32: uint32
        `nbr` / `row`)."""
        ce = max(1, int(num_targets) * self._fanouts[layer])

        def mk(n, dt_):
            t = torch.empty(n, dtype=dt_)
            return (t.pin_memory() if pinned else t).numpy()
        if ids32:  # torch has no pinned uint32 tensors everywhere: int32 storage viewed as uint32
            out = dict(nbr=mk(ce, torch.int32).view(np.uint32), ts=mk(ce, torch.float32), dt=mk(ce, torch.float32),
                       eid=mk(ce, torch.int64), row=mk(ce, torch.int32).view(np.uint32))
        else:
            out = dict(nbr=mk(ce, torch.int64), ts=mk(ce, torch.float32), dt=mk(ce, torch.float32), eid=mk(ce, torch.int64),
                       row=mk(ce, torch.int64))
        out["edge_offsets"] = np.zeros(int(num_batches) + 1, dtype=np.uint64)
        return out

    def set_host_output_mode(self, mode: int):
        """0: auto (default); 1: device mirror + D2H copies; 2: pinned host outputs written in place by the kernel"""
        check(self._L.gf_sampler_set_host_output_mode(self._h, int(mode)))

    def sample_numpy(self, target_vertices: np.ndarray, timestamps: np.ndarray):
        """Host in, host out: what the reference's `_TemporalSampler.sample` returns (csrc/api.cc:116-118), i.e.
        [layer][snapshot] results as numpy arrays (NOT reversed over layers).  The arrays are views into pinned
        buffers owned by the sampler (the kernel writes them in place over PCIe) and are overwritten by the next
        call."""
        keep, pn, pt, T, kind = self._inputs(target_vertices, timestamps)
        nsnap = self._num_snapshots
        cache = getattr(self, "_pinned", None)
        if cache is None or T > cache[0]:
            # ONE pinned allocation for every array of every step, declared to the library once (bind_host_outputs)
            cap0 = int(T * 1.5) + 64
            A = 256
            layout, total, cap = [], 0, cap0
            for layer in range(self._num_layers):
                for sn in range(nsnap):
                    ce = max(cap * self._fanouts[layer], 1)
                    o = []
                    for sz in ((cap + ce) * 8, (cap + ce) * 4, ce * 4, ce * 8, ce * 8, ce * 8):
                        o.append(total)
                        total += (sz + A - 1) // A * A
                    layout.append((o, cap))
                cap = cap * (1 + self._fanouts[layer])
            arena = torch.empty(total, dtype=torch.uint8).pin_memory()
            base = arena.data_ptr()
            a8 = arena.numpy()
            a64, a32 = a8.view(np.int64), a8.view(np.float32)
            arr = (SamplingResultC * len(layout))()
            for i, (o, c) in enumerate(layout):
                arr[i] = SamplingResultC(base + o[0], base + o[1], base + o[2], base + o[3], base + o[4], base + o[5], c, 0, 0)
            try:
                check(self._L.gf_sampler_bind_host_outputs(self._h, base, total))
            except NotImplementedError:  # very large area: the library checks the arrays on every call instead
                pass
            cache = (cap0, (arena, a64, a32, [o for o, _ in layout]), arr)
            self._pinned = cache
        _, (_arena, a64, a32, offs), arr = cache
        check(self._L.gf_sampler_sample(self._h, pn, pt, T, arr, kind, GF_PTR_HOST, _stream_ptr(self._device)))
        del keep
        out, i = [], 0
        for layer in range(self._num_layers):
            lay = []
            for sn in range(nsnap):
                r, o = arr[i], offs[i]
                Td, S = r.num_dst, r.num_edges
                a, b_, c, d, e, f = o[0] >> 3, o[1] >> 2, o[2] >> 2, o[3] >> 3, o[4] >> 3, o[5] >> 3
                lay.append(dict(all_nodes=a64[a:a + Td + S], all_timestamps=a32[b_:b_ + Td + S],
                                delta_timestamps=a32[c:c + S], eids=a64[d:d + S], row=a64[e:e + S], col=a64[f:f + S],
                                num_dst_nodes=Td, num_src_nodes=Td + S))
                i += 1
            out.append(lay)
        return out

    def set_profiling(self, on: bool):
        check(self._L.gf_sampler_set_profiling(self._h, 1 if on else 0))

    def get_profile(self, reset: bool = True):
        """{phase: (total ms, launches)} for phases locate / scan / emit (CUDA events inside the library)"""
        ms = (C.c_double * 3)()
        cnt = (C.c_uint64 * 3)()
        check(self._L.gf_sampler_get_profile(self._h, ms, cnt, 1 if reset else 0))
        return {n: (ms[i], cnt[i]) for i, n in enumerate(("locate", "scan", "emit"))}

    def launch_index(self) -> int:
        v = C.c_uint64()
        check(self._L.gf_sampler_get_launch_index(self._h, C.byref(v)))
        return v.value

    def set_launch_index(self, v: int):
        check(self._L.gf_sampler_set_launch_index(self._h, int(v)))

    def set_variant(self, variant: int):
        check(self._L.gf_sampler_set_variant(self._h, int(variant)))
