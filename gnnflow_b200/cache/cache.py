"""Feature cache on the GPU (reference gnnflow/cache/cache.py:10-413) with the gather and the policy updates as
CUDA kernels behind the C ABI (gf_cache_gather / gf_cache_update_*), instead of ~12 torch index kernels +
torch.unique + host index_select + H2D per MFG block.

Same constructor, attributes and methods as the reference `Cache`.  Differences, all invisible in the values:
  * misses are read by the gather kernel straight from the feature table -- a CUDA tensor, or a pinned /
    host-registered CPU tensor reached zero-copy through UVA -- so there is no host round trip per block;
  * hit ratios stay device tensors (no synchronisation inside fetch_feature);
  * the distributed (KVStore) miss path is not provided (out of scope)."""
import ctypes as C
import os
from typing import List, Optional, Union

import numpy as np
import torch

from .. import _lib
from .._lib import CacheStateC, check


_CHECK_IDS = os.environ.get("GNNFLOW_B200_CHECK_IDS", "0") not in ("", "0")
_REGISTER_IN_PLACE_BYTES = 64 << 20  # smaller pageable tables are simply copied into pinned memory


class _HostRegistration:
    """Keeps a pageable CPU tensor registered with CUDA (zero-copy reads by the gather kernel) and unregisters
    it when the last cache using it goes away -- a stale registration would poison later copies from re-used
    host addresses."""

    def __init__(self, L, tensor: torch.Tensor):
        self._L, self.tensor, self._owned = L, tensor, False
        owned = C.c_int(0)
        check(L.gf_host_register(C.c_void_p(tensor.data_ptr()), tensor.numel() * tensor.element_size(), C.byref(owned)))
        self._owned = bool(owned.value)

    def __del__(self):
        if getattr(self, "_owned", False):
            try:
                self._L.gf_host_unregister(C.c_void_p(self.tensor.data_ptr()))
            except Exception:  # noqa: BLE001
                pass
            self._owned = False


def _device_visible(feats: torch.Tensor, device: torch.device, L=None):
    """-> (float32 contiguous tensor whose data_ptr() a kernel on `device` may dereference, keepalive object)"""
    if feats.dtype != torch.float32:
        feats = feats.to(torch.float32)
    feats = feats.contiguous()
    if feats.is_cuda:
        return (feats if feats.device == device else feats.to(device)), None
    if feats.is_pinned():
        return feats, None
    if L is not None and feats.numel() * 4 >= _REGISTER_IN_PLACE_BYTES:
        try:  # register the existing pages in place (no copy); falls back to a pinned copy
            return feats, _HostRegistration(L, feats)
        except Exception:  # noqa: BLE001
            pass
    return feats.pin_memory(), None


class Cache:
    """Feature cache on GPU"""

    def __init__(self, edge_cache_ratio: int, node_cache_ratio: int,
                 num_nodes: int, num_edges: int,
                 device: Union[str, torch.device],
                 node_feats: Optional[torch.Tensor] = None,
                 edge_feats: Optional[torch.Tensor] = None,
                 dim_node_feat: Optional[int] = 0,
                 dim_edge_feat: Optional[int] = 0,
                 pinned_nfeat_buffs: Optional[torch.Tensor] = None,
                 pinned_efeat_buffs: Optional[torch.Tensor] = None,
                 kvstore_client=None,
                 distributed: Optional[bool] = False,
                 neg_sample_ratio: Optional[int] = 1):
        if device == 'cpu' or device == torch.device('cpu'):
            raise ValueError('Cache must be on GPU')
        if node_feats is None and edge_feats is None and not distributed:
            raise ValueError('At least one of node_feats and edge_feats must be provided')
        if node_feats is not None and node_feats.shape[0] != num_nodes:
            raise ValueError('The number of nodes in node_feats {} does not match num_nodes {}'.format(
                node_feats.shape[0], num_nodes))
        if edge_feats is not None and edge_feats.shape[0] != num_edges:
            raise ValueError('The number of edges in edge_feats {} does not match num_edges {}'.format(
                edge_feats.shape[0], num_edges))
        if distributed:
            raise NotImplementedError('the KVStore-backed distributed miss path is out of scope (SURVEY.md section 2); '
                                      'remote rows are fetched with NCCL all-to-all in gnnflow_b200.distributed')
        assert edge_cache_ratio >= 0 and edge_cache_ratio <= 1, 'edge_cache_ratio must be in [0, 1]'
        assert node_cache_ratio >= 0 and node_cache_ratio <= 1, 'node_cache_ratio must be in [0, 1]'
        self._L = _lib.lib()
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.edge_cache_ratio = edge_cache_ratio
        self.node_cache_ratio = node_cache_ratio
        self.node_capacity = int(node_cache_ratio * num_nodes)
        self.edge_capacity = int(edge_cache_ratio * num_edges)
        self.num_nodes = num_nodes
        self.num_edges = num_edges
        self.node_feats, self._node_reg = _device_visible(node_feats, self.device, self._L) \
            if node_feats is not None else (None, None)
        self.edge_feats, self._edge_reg = _device_visible(edge_feats, self.device, self._L) \
            if edge_feats is not None else (None, None)
        self.dim_node_feat = dim_node_feat if node_feats is not None else 0
        self.dim_edge_feat = dim_edge_feat if edge_feats is not None else 0
        if self.node_feats is not None:
            assert self.node_feats.shape[1] == dim_node_feat
        if self.edge_feats is not None:
            assert self.edge_feats.shape[1] == dim_edge_feat
        self.pinned_nfeat_buffs = pinned_nfeat_buffs  # accepted for API compatibility; unused (zero-copy misses)
        self.pinned_efeat_buffs = pinned_efeat_buffs
        # hit statistics: gf_cache_fetch stores each block's hit count into the next slot of a small device ring; the
        # ratios of the last fetch_feature are computed from it when somebody reads them
        self._ratio_override = {"node": 0, "edge": 0}
        self._last_fetch = {"node": [], "edge": []}
        self._hit_slot = 0
        self.kvstore_client = kvstore_client
        self.distributed = distributed
        self.target_edge_features = None
        self.neg_sample_ratio = neg_sample_ratio
        self._scratch = None
        dev = self.device
        self._bad_ids = torch.zeros(1, dtype=torch.int32, device=dev)  # fetches that saw an id outside the table
        self._hit_stats = torch.zeros(256, dtype=torch.int64, device=dev)
        if self.dim_node_feat != 0:
            self.cache_node_buffer = torch.zeros(self.node_capacity, self.dim_node_feat, dtype=torch.float32, device=dev)
            self.cache_node_flag = torch.zeros(num_nodes, dtype=torch.bool, device=dev)
            self.cache_node_map = torch.zeros(num_nodes, dtype=torch.int64, device=dev) - 1
            self.cache_index_to_node_id = torch.zeros(self.node_capacity, dtype=torch.int64, device=dev) - 1
        if self.dim_edge_feat != 0:
            self.cache_edge_buffer = torch.zeros(self.edge_capacity, self.dim_edge_feat, dtype=torch.float32, device=dev)
            self.cache_edge_flag = torch.zeros(num_edges, dtype=torch.bool, device=dev)
            self.cache_edge_map = torch.zeros(num_edges, dtype=torch.int64, device=dev) - 1
            self.cache_index_to_edge_id = torch.zeros(self.edge_capacity, dtype=torch.int64, device=dev) - 1

    # ------------------------------------------------------------------------------- reference API
    def get_mem_size(self) -> int:
        mem_size = 0
        for kind in ("node", "edge"):
            if getattr(self, "dim_%s_feat" % kind) != 0:
                for t in (getattr(self, "cache_%s_buffer" % kind), getattr(self, "cache_%s_flag" % kind),
                          getattr(self, "cache_%s_map" % kind), getattr(self, "cache_index_to_%s_id" % kind)):
                    mem_size += t.element_size() * t.nelement()
        return mem_size

    def init_cache(self, *args, **kwargs):
        """Init the cache with the first `capacity` rows (cache.py:175-195)"""
        for kind, feats in (("node", self.node_feats), ("edge", self.edge_feats)):
            if getattr(self, "dim_%s_feat" % kind) == 0:
                continue
            cap = getattr(self, "%s_capacity" % kind)
            ids = torch.arange(cap, dtype=torch.int64, device=self.device)
            getattr(self, "cache_%s_buffer" % kind)[ids] = feats[:cap].to(self.device, non_blocking=True)
            getattr(self, "cache_%s_flag" % kind)[ids] = True
            setattr(self, "cache_index_to_%s_id" % kind, ids)
            getattr(self, "cache_%s_map" % kind)[ids] = ids

    @staticmethod
    def _grown(t: torch.Tensor, n: int, fill) -> torch.Tensor:
        """`t` extended to n rows; the new rows hold `fill` (never uninitialised memory: the gather kernel
        dereferences flag / map / index_to_id of every id it is given)"""
        if n <= t.shape[0]:
            return t
        out = torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        out[:t.shape[0]] = t
        out[t.shape[0]:] = fill
        return out

    def resize(self, new_num_nodes: int, new_num_edges: int):
        """Grow the cache for a graph that gained vertices / edges (cache.py:197-221).  What is cached stays cached;
        the new ids start uncached (flag False, map -1), the new slots empty (index_to_id -1, zero rows).  The caller
        re-binds `node_feats` / `edge_feats` to tables covering the new ids, as after any growth of the dataset
        (`set_feats`)."""
        for kind, new_n in (("node", new_num_nodes), ("edge", new_num_edges)):
            if getattr(self, "dim_%s_feat" % kind) == 0 or new_n <= getattr(self, "num_%ss" % kind):
                continue
            setattr(self, "num_%ss" % kind, new_n)
            cap = int(getattr(self, "%s_cache_ratio" % kind) * new_n)
            setattr(self, "%s_capacity" % kind, cap)
            for name, n, fill in (("cache_%s_buffer", cap, 0.0), ("cache_%s_flag", new_n, False),
                                  ("cache_%s_map", new_n, -1), ("cache_index_to_%s_id", cap, -1)):
                setattr(self, name % kind, self._grown(getattr(self, name % kind), n, fill))

    def set_feats(self, node_feats: Optional[torch.Tensor] = None, edge_feats: Optional[torch.Tensor] = None):
        """Re-bind the feature tables the misses are read from (after the dataset grew; see `resize`)."""
        if node_feats is not None:
            if node_feats.shape[0] < self.num_nodes or node_feats.shape[1] != self.dim_node_feat:
                raise ValueError('node_feats must be [>= {}, {}]'.format(self.num_nodes, self.dim_node_feat))
            self.node_feats, self._node_reg = _device_visible(node_feats, self.device, self._L)
        if edge_feats is not None:
            if edge_feats.shape[0] < self.num_edges or edge_feats.shape[1] != self.dim_edge_feat:
                raise ValueError('edge_feats must be [>= {}, {}]'.format(self.num_edges, self.dim_edge_feat))
            self.edge_feats, self._edge_reg = _device_visible(edge_feats, self.device, self._L)

    def reset(self):
        raise NotImplementedError

    def update_node_cache(self, ids: torch.Tensor, hit_mask: torch.Tensor):
        """Policy update for one fetch of node ids (hit_mask as produced by the gather).  The reference's signature
        (cached_node_index, uncached_node_id, uncached_node_feature; cache.py:228-240) carries the same information
        after torch.unique; here de-duplication happens inside the kernel pipeline."""
        raise NotImplementedError

    def update_edge_cache(self, ids: torch.Tensor, hit_mask: torch.Tensor):
        raise NotImplementedError

    # ------------------------------------------------------------------------------------- plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _state(self, kind: str) -> CacheStateC:
        count = getattr(self, "cache_%s_count" % kind, None)
        return CacheStateC(getattr(self, "cache_%s_buffer" % kind).data_ptr(), getattr(self, "cache_%s_flag" % kind).data_ptr(),
                           getattr(self, "cache_%s_map" % kind).data_ptr(),
                           getattr(self, "cache_index_to_%s_id" % kind).data_ptr(),
                           count.data_ptr() if count is not None else None,
                           getattr(self, "%s_capacity" % kind), getattr(self, "num_%ss" % kind),
                           getattr(self, "dim_%s_feat" % kind))

    def _get_scratch(self, n: int, capacity: int, num_items: int):
        return self._scratch_of(int(self._L.gf_cache_update_scratch_bytes(n, capacity, num_items)))

    def _scratch_of(self, need: int):
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=self.device)
        return self._scratch

    def _count_bound(self, kind: str) -> int:
        """every entry of cache_<kind>_count lies in [-bound, bound]: one step per update plus the initial 1 of LFU"""
        b = getattr(self, "_updates", None)
        if b is None:
            b = self._updates = {"node": 0, "edge": 0}
        b[kind] += 1
        return b[kind] + 1

    def _gather(self, kind: str, ids: torch.Tensor):
        """-> (ids, features [n, D] f32, hit_mask [n] uint8, hits (device int64 [1])): the stand-alone gather
        (gf_cache_gather); `fetch_feature` uses the fused call below"""
        feats = getattr(self, "%s_feats" % kind)
        dim = getattr(self, "dim_%s_feat" % kind)
        ids = ids.to(torch.int64).contiguous()
        n = ids.shape[0]
        limit = min(getattr(self, "num_%ss" % kind), feats.shape[0])
        out = torch.empty(n, dim, dtype=torch.float32, device=self.device)
        hit = torch.empty(n, dtype=torch.uint8, device=self.device)
        nhits = torch.zeros(1, dtype=torch.int64, device=self.device)
        cap = getattr(self, "%s_capacity" % kind)
        flag = getattr(self, "cache_%s_flag" % kind).data_ptr() if cap > 0 else None
        if flag is None:
            hit.zero_()
        check(self._L.gf_cache_gather(ids.data_ptr(), n, limit, flag, getattr(self, "cache_%s_map" % kind).data_ptr(),
                                      getattr(self, "cache_%s_buffer" % kind).data_ptr(), feats.data_ptr(), dim,
                                      out.data_ptr(), hit.data_ptr(), nhits.data_ptr(), self._bad_ids.data_ptr(),
                                      self._stream()))
        if _CHECK_IDS:
            self.check_ids()
        return ids, out, hit, nhits

    # policy hooks of the fused fetch (gf_cache_fetch): subclasses set `_policy` and, where they have them, the FIFO ring
    # pointer / the bound of their counts
    _policy = 3  # GF_CACHE_STATIC: never updated

    def _fifo_ptr(self, kind: str):
        return None

    def _count_floor(self, kind: str):
        """LRU: device int32[1], a lower bound of every water level (kept by gf_cache_fetch)"""
        return None

    def _fetch(self, kind: str, ids: torch.Tensor, update: bool, stream) -> torch.Tensor:
        """One block of `fetch_feature` in one C call: rows through the cache, hit count into the next statistics slot,
        policy update.  -> features [n, D] f32"""
        feats = getattr(self, "%s_feats" % kind)
        dim = getattr(self, "dim_%s_feat" % kind)
        if ids.dtype is not torch.int64 or not ids.is_contiguous():
            ids = ids.to(torch.int64).contiguous()
        n = ids.shape[0]
        out = torch.empty((n, dim), dtype=torch.float32, device=self.device)
        st = self._state(kind)
        update = bool(update) and st.capacity > 0 and self._policy != 3
        scratch = self._get_scratch(n, st.capacity, st.num_items) if st.capacity > 0 else self._scratch_of(256)
        stats = self._hit_stats
        slot = self._hit_slot
        self._hit_slot = (slot + 1) % stats.shape[0]
        ptr = self._fifo_ptr(kind) if update else None
        bound = self._count_bound(kind) if update and self._policy != 1 else 0
        floor = self._count_floor(kind) if update else None
        check(self._L.gf_cache_fetch(st, ids.data_ptr(), n, feats.data_ptr(), feats.shape[0], self._policy,
                                     ptr.data_ptr() if ptr is not None else None, bound,
                                     floor.data_ptr() if floor is not None else None, 1 if update else 0,
                                     out.data_ptr(), stats.data_ptr() + 8 * slot, self._bad_ids.data_ptr(),
                                     scratch.data_ptr(), scratch.numel(), stream))
        self._last_fetch[kind].append((slot, n))
        return out

    def _ratio(self, kind: str):
        """mean over the blocks of the last fetch_feature of hits / rows (cache.py:330,411), computed when it is read: a
        0-dim device tensor (no synchronisation unless the caller converts it)"""
        blocks = self._last_fetch[kind]
        live = [(s, n) for s, n in blocks if n > 0]
        if not live:
            return 0
        if len(live) > self._hit_stats.shape[0]:
            raise RuntimeError("hit statistics of more than {} blocks are not kept".format(self._hit_stats.shape[0]))
        slots = torch.tensor([s for s, _ in live], dtype=torch.int64, device=self.device)
        rows = torch.tensor([float(n) for _, n in live], dtype=torch.float64, device=self.device)
        return ((self._hit_stats[slots].to(torch.float64) / rows).sum() / len(blocks)).to(torch.float32)

    @property
    def cache_node_ratio(self):
        return self._ratio("node") if self._ratio_override["node"] is None else self._ratio_override["node"]

    @cache_node_ratio.setter
    def cache_node_ratio(self, v):
        self._ratio_override["node"] = v

    @property
    def cache_edge_ratio(self):
        return self._ratio("edge") if self._ratio_override["edge"] is None else self._ratio_override["edge"]

    @cache_edge_ratio.setter
    def cache_edge_ratio(self, v):
        self._ratio_override["edge"] = v

    def check_ids(self):
        """Raise IndexError if a fetch since the last check was given an id outside its feature table (the reference's
        torch indexing raises at the fetch itself, cache.py:283; here such rows are zero-filled and counted on the
        device, and this is the one host synchronisation that looks at the counter).  GNNFLOW_B200_CHECK_IDS=1 calls
        it after every gather."""
        bad = int(self._bad_ids.item())
        if bad:
            self._bad_ids.zero_()
            raise IndexError("{} fetch(es) contained ids outside the feature table".format(bad))

    def fetch_feature(self, mfgs: List[List], eid: Optional[np.ndarray] = None, update_cache: bool = True,
                      target_edge_features: bool = True):
        """Fetch node features into b.srcdata['h'] for the blocks of mfgs[0] and edge features into b.edata['f'] for
        every block (cache.py:255-413).  Values equal node_feats[ID] / edge_feats[ID] bit for bit.  One C call
        (gf_cache_fetch) per block: gather, hit statistics and policy update; nothing here synchronises with the GPU."""
        stream = self._stream()
        self._ratio_override["node"] = self._ratio_override["edge"] = None
        self._last_fetch = {"node": [], "edge": []}
        if self.dim_node_feat != 0:
            for b in mfgs[0]:
                nodes = b.srcdata['ID']
                assert isinstance(nodes, torch.Tensor)
                if len(nodes) == 0:
                    b.srcdata['h'] = torch.empty((0, self.dim_node_feat), dtype=torch.float32, device=self.device)
                    self._last_fetch["node"].append((-1, 0))
                    continue
                b.srcdata['h'] = self._fetch("node", nodes, update_cache, stream)
        if self.dim_edge_feat != 0:
            for mfg in mfgs:
                for b in mfg:
                    edges = b.edata['ID']
                    assert isinstance(edges, torch.Tensor)
                    if len(edges) == 0:
                        continue
                    b.edata['f'] = self._fetch("edge", edges, update_cache, stream)
            if target_edge_features and eid is not None:
                e = torch.as_tensor(eid).to(self.device, torch.int64).contiguous()
                out = torch.empty(e.shape[0], self.dim_edge_feat, dtype=torch.float32, device=self.device)
                check(self._L.gf_gather_rows(e.data_ptr(), e.shape[0], self.edge_feats.shape[0], self.edge_feats.data_ptr(),
                                             self.dim_edge_feat, out.data_ptr(), self._bad_ids.data_ptr(), stream))
                self.target_edge_features = out
        if _CHECK_IDS:
            self.check_ids()
        return mfgs
