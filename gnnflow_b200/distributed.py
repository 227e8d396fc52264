"""Multi-GPU operation of the store + sampler on the GPUs of one box (one process per GPU, torch.distributed).

Two modes:

  replicated   every rank holds the whole graph and samples its own shard of the batches
               (`shard_batch_indices`, the rule of the reference's DistributedBatchSampler, gnnflow/data.py:157-159).
               No collective on the data path.

  partitioned  vertices are hash-partitioned by SOURCE vertex (reference gnnflow/distributed/partition.py:312-325,
               whose `hash(str(v)) % P` is salted per process; here owner(v) = splitmix64(v) % P, and the partition
               table stays the contract).  Each rank stores the out-edges of the vertices it owns.  A sampling step
               routes every target to its owner, samples there, and routes the neighbours back -- one all-to-all
               each way over NCCL/NVLink -- replacing the reference's per-partition rpc_async fan-out
               (gnnflow/distributed/dist_sampler.py:159-314).  The merged result is in the single-GPU order
               (target-major), i.e. bit-identical to sampling the unpartitioned graph for the recent policy.

The exchange logic is backend-agnostic (NCCL on GPUs, gloo on CPU for the tests); the local sampling engine is
injected (`CudaEngine` in production)."""
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

_M64 = (1 << 64) - 1


def splitmix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (deterministic across processes, unlike Python's hash(str))."""
    z = (np.asarray(x).astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _splitmix64_torch(x: torch.Tensor) -> torch.Tensor:
    # int64 arithmetic wraps like uint64; logical right shifts are emulated by masking the sign extension
    def shr(v, k):
        return (v >> k) & ((1 << (64 - k)) - 1)
    z = x + (-0x61C8864680B583EB)  # 0x9E3779B97F4A7C15 as int64
    z = (z ^ shr(z, 30)) * (-0x40A7B892E31B1A47)  # 0xBF58476D1CE4E5B9
    z = (z ^ shr(z, 27)) * (-0x6B2FB644ECCEEE15)  # 0x94D049BB133111EB
    return z ^ shr(z, 31)


def owner_of(vertices, num_partitions: int):
    """Partition id of each vertex: splitmix64(v) % P.  numpy in -> numpy out, torch in -> torch out."""
    if isinstance(vertices, torch.Tensor):
        h = _splitmix64_torch(vertices.to(torch.int64))
        # unsigned modulo of a value held in int64
        lo = h & 0x7FFFFFFFFFFFFFFF
        r = lo % num_partitions
        top = ((h >> 63) & 1) * ((1 << 63) % num_partitions)
        return (r + top) % num_partitions
    return (splitmix64(np.asarray(vertices)) % np.uint64(num_partitions)).astype(np.int64)


def partition_table(num_vertices: int, num_partitions: int) -> torch.Tensor:
    """The reference's contract object: int8 owner per vertex id (-1 = unassigned), dist_sampler.py:174-236."""
    return torch.from_numpy(owner_of(np.arange(num_vertices), num_partitions).astype(np.int8))


def shard_batch_indices(num_batches: int, rank: int, world_size: int) -> List[int]:
    """Replicated mode: batch b belongs to rank b % world_size (gnnflow/data.py:157-159)."""
    return [b for b in range(num_batches) if b % world_size == rank]


class PartitionedDynamicGraph:
    """Keeps the out-edges of the vertices this rank owns.  Every rank is handed the same (replicated) edge batch;
    the rows whose source it owns are appended to the local store."""

    def __init__(self, local_graph, rank: int, world_size: int, table: Optional[torch.Tensor] = None):
        self.graph, self.rank, self.world_size, self.table = local_graph, rank, world_size, table

    def owner(self, vertices):
        if self.table is not None:
            if isinstance(vertices, torch.Tensor):
                return self.table.to(vertices.device)[vertices].to(torch.int64)
            return self.table.numpy()[np.asarray(vertices)].astype(np.int64)
        return owner_of(vertices, self.world_size)

    def add_edges(self, source_vertices, target_vertices, timestamps, eids=None, add_reverse: bool = False):
        if eids is None:
            raise ValueError("partitioned ingest needs explicit edge ids (the global counter is not replicated)")
        if isinstance(source_vertices, torch.Tensor) and source_vertices.is_cuda:
            return self._add_edges_device(source_vertices, target_vertices, timestamps, eids, add_reverse)
        src, dst = np.asarray(source_vertices), np.asarray(target_vertices)
        ts = np.asarray(timestamps)
        eids = np.asarray(eids)
        if add_reverse:
            src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
            ts, eids = np.concatenate([ts, ts]), np.concatenate([eids, eids])
        keep = self.owner(src) == self.rank
        if keep.any():
            self.graph.add_edges(src[keep], dst[keep], ts[keep], eids[keep])

    def _add_edges_device(self, src, dst, ts, eids, add_reverse):
        """CUDA tensors: the rows this rank owns are picked out on the device (gf_dispatch_edges, one launch) -- the
        replacement of the reference's dispatcher (gnnflow/distributed/dispatcher.py:41-100: host grouping by partition +
        one RPC per partition)"""
        import ctypes as C
        from . import _lib
        dev = src.device
        src, dst = src.to(torch.int64).contiguous(), dst.to(dev, torch.int64).contiguous()
        ts, eids = ts.to(dev, torch.float32).contiguous(), eids.to(dev, torch.int64).contiguous()
        if add_reverse:
            src, dst = torch.cat([src, dst]), torch.cat([dst, src])
            ts, eids = torch.cat([ts, ts]), torch.cat([eids, eids])
        n = src.shape[0]
        o = (torch.empty_like(src), torch.empty_like(dst), torch.empty_like(ts), torch.empty_like(eids))
        tab = None if self.table is None else self.table.to(dev, torch.int8).contiguous()
        cnt = C.c_uint64()
        _lib.check(_lib.lib().gf_dispatch_edges(
            src.data_ptr(), dst.data_ptr(), ts.data_ptr(), eids.data_ptr(), n, tab.data_ptr() if tab is not None else None,
            tab.shape[0] if tab is not None else 0, self.rank, self.world_size, o[0].data_ptr(), o[1].data_ptr(),
            o[2].data_ptr(), o[3].data_ptr(), C.byref(cnt), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        k = int(cnt.value)
        if k:
            self.graph.add_edges(o[0][:k], o[1][:k], o[2][:k], o[3][:k])
        return k

    def __getattr__(self, name):
        return getattr(self.graph, name)


def to_reference_order(res: dict, owner: torch.Tensor) -> dict:
    """Re-order one merged (layer, snapshot) result from the single-GPU edge order (target-major) into the order the
    reference's distributed sampler produces (gnnflow/distributed/dist_sampler.py:276-299): its merge concatenates the
    per-partition results in partition order, so edges are sorted by (partition of the target, target, slot) while `row`
    keeps pointing at the target's position in the caller's order; the destination part of all_nodes / all_timestamps is
    unchanged.  `owner` = partition id of every target (-1: unassigned, no edges).  A stable sort of the edges by the
    partition of their target; the edge SET is the same, so every downstream aggregation gives the same numbers."""
    T = int(res["num_dst_nodes"])
    row = res["row"]
    perm = torch.argsort(owner.to(row.device)[row], stable=True)
    out = dict(res)
    out["all_nodes"] = torch.cat([res["all_nodes"][:T], res["all_nodes"][T:][perm]])
    out["all_timestamps"] = torch.cat([res["all_timestamps"][:T], res["all_timestamps"][T:][perm]])
    for k in ("delta_timestamps", "eids", "row"):
        out[k] = res[k][perm]
    return out


class CudaEngine:
    """Local sampling engine over a gnnflow_b200.TemporalSampler: device tensors in, device tensors out."""

    def __init__(self, sampler):
        self.sampler = sampler

    def sample_layer(self, nodes: torch.Tensor, ts: torch.Tensor, layer: int, snapshot: int):
        r = self.sampler.sample_layer(nodes, ts, layer, snapshot, to_dgl_block=False).tensors()
        T = nodes.shape[0]
        return dict(nbr=r["all_nodes"][T:], ts=r["all_timestamps"][T:], dt=r["delta_timestamps"], eid=r["eids"],
                    row=r["row"])


def _all_to_all(out_list_sizes, inp: torch.Tensor, in_sizes, group):
    out = inp.new_empty((int(sum(out_list_sizes)),) + tuple(inp.shape[1:]))
    dist.all_to_all_single(out, inp.contiguous(), output_split_sizes=[int(x) for x in out_list_sizes],
                           input_split_sizes=[int(x) for x in in_sizes], group=group)
    return out


class DistributedTemporalSampler:
    """Hash-partitioned sampling with one exchange each way per (layer, snapshot) step
    (replaces gnnflow/distributed/dist_sampler.py:129-314; same `sample` / `sample_layer` surface)."""

    def __init__(self, engine, fanouts: List[int], num_snapshots: int = 1, rank: Optional[int] = None,
                 world_size: Optional[int] = None, table: Optional[torch.Tensor] = None, group=None,
                 merge_order: str = "target"):
        if merge_order not in ("target", "reference"):
            raise ValueError("merge_order must be 'target' (single-GPU order) or 'reference' (dist_sampler.py:276-299)")
        self.merge_order = merge_order
        self.engine = engine
        self.fanouts = [int(f) for f in fanouts]
        self.num_layers, self.num_snapshots = len(self.fanouts), int(num_snapshots)
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world_size = dist.get_world_size(group) if world_size is None else world_size
        self.table = table
        self.bytes_sent = 0  # payload bytes this rank pushed through the exchanges (for the NVLink roofline)

    def _owner(self, nodes: torch.Tensor) -> torch.Tensor:
        if self.table is not None:
            n = self.table.shape[0]
            inside = (nodes >= 0) & (nodes < n)
            own = torch.full_like(nodes, -1)
            own[inside] = self.table.to(nodes.device)[nodes[inside]].to(torch.int64)
            return own
        return owner_of(nodes, self.world_size)

    def sample_layer(self, nodes: torch.Tensor, ts: torch.Tensor, layer: int, snapshot: int):
        """-> dict(all_nodes, all_timestamps, delta_timestamps, eids, row, col, num_dst_nodes, num_src_nodes)"""
        P, dev = self.world_size, nodes.device
        T = nodes.shape[0]
        own = self._owner(nodes)
        routed = own >= 0  # -1 = unpartitioned vertex: no neighbours (dist_sampler.py:223-236)
        key = torch.where(routed, own, torch.full_like(own, P))
        order = torch.argsort(key, stable=True)
        send_counts = torch.bincount(key, minlength=P + 1)[:P]
        n_routed = int(send_counts.sum())
        order_r = order[:n_routed]
        # ---- exchange 1: targets to their owners
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        sc, rc = send_counts.tolist(), recv_counts.tolist()
        r_nodes = _all_to_all(rc, nodes[order_r], sc, self.group)
        r_ts = _all_to_all(rc, ts[order_r], sc, self.group)
        self.bytes_sent += n_routed * 12
        # ---- local sampling of everything this rank owns
        loc = self.engine.sample_layer(r_nodes, r_ts, layer, snapshot)
        R = r_nodes.shape[0]
        per_target = torch.bincount(loc["row"], minlength=R) if R else torch.zeros(0, dtype=torch.int64, device=dev)
        # ---- exchange 2: per-target counts and the neighbour records back to the requesters
        back_counts = _all_to_all(sc, per_target, rc, self.group)  # counts of MY targets, in `order_r` order
        bounds = torch.cumsum(torch.tensor([0] + rc, device=dev), 0)
        cum = torch.cat([per_target.new_zeros(1), torch.cumsum(per_target, 0)])
        nbr_send = (cum[bounds[1:]] - cum[bounds[:-1]]).tolist()  # neighbours I return to each requester
        bc = torch.cat([back_counts.new_zeros(1), torch.cumsum(back_counts, 0)])
        sb = torch.cumsum(torch.tensor([0] + sc, device=dev), 0)
        nbr_recv = (bc[sb[1:]] - bc[sb[:-1]]).tolist()
        got = {k: _all_to_all(nbr_recv, loc[k], nbr_send, self.group) for k in ("nbr", "ts", "dt", "eid")}
        self.bytes_sent += int(sum(nbr_send)) * 24 + R * 8
        # ---- merge into the single-GPU order: target-major, the per-target order the owner produced
        counts = torch.zeros(T, dtype=torch.int64, device=dev)
        counts[order_r] = back_counts
        S = int(counts.sum())
        out_off = torch.cumsum(counts, 0) - counts  # first output slot of every original target
        # received neighbour j belongs to routed target t(j) = order_r[repeat(arange, back_counts)]
        tgt_sorted = torch.repeat_interleave(torch.arange(n_routed, device=dev), back_counts)
        within = torch.arange(S, device=dev) - (bc[:-1])[tgt_sorted]
        dest = out_off[order_r[tgt_sorted]] + within
        res = {}
        for k, name in (("nbr", "nbr"), ("ts", "nts"), ("dt", "delta_timestamps"), ("eid", "eids")):
            buf = torch.empty_like(got[k])
            buf[dest] = got[k]
            res[name] = buf
        row = torch.repeat_interleave(torch.arange(T, device=dev), counts)
        out = dict(all_nodes=torch.cat([nodes, res["nbr"]]), all_timestamps=torch.cat([ts, res["nts"]]),
                   delta_timestamps=res["delta_timestamps"], eids=res["eids"], row=row,
                   col=torch.arange(T, T + S, device=dev), num_dst_nodes=T, num_src_nodes=T + S)
        return to_reference_order(out, own) if self.merge_order == "reference" else out

    def sample(self, nodes: torch.Tensor, ts: torch.Tensor):
        """[layer][snapshot] results, layers reversed like TemporalSampler.sample (temporal_sampler.py:163-164).
        Every rank must call this the same number of times (collective)."""
        results = []
        for layer in range(self.num_layers):
            lay = []
            for s in range(self.num_snapshots):
                if layer == 0:
                    n_in, t_in = nodes, ts
                else:
                    n_in, t_in = results[-1][s]["all_nodes"], results[-1][s]["all_timestamps"]
                lay.append(self.sample_layer(n_in, t_in, layer, s))
            results.append(lay)
        results.reverse()
        return results


class PeerTemporalSampler:
    """Hash-/table-partitioned sampling over NVLink peer memory (C ABI: gf_peer_*, gf_sampler_sample_layer_partitioned).

    Same surface as `DistributedTemporalSampler` / the reference's `DistributedTemporalSampler.sample`
    (gnnflow/distributed/dist_sampler.py:129-157), but the exchange is done by the kernels themselves: requests and
    neighbours are written straight into the peers' exchange windows, so one (layer, snapshot) step costs seven small
    launches and one host synchronisation, with no collective call.  torch.distributed is used once, to all-gather
    the window handles.  All ranks of the group must live on one box (CUDA IPC) and call `sample` in lockstep.
    The result equals sampling the unpartitioned graph bit for bit, for the recent AND the uniform policy."""

    def __init__(self, sampler, max_targets: int, group=None, table: Optional[torch.Tensor] = None,
                 merge_order: str = "target"):
        import ctypes as C
        if merge_order not in ("target", "reference"):
            raise ValueError("merge_order must be 'target' (single-GPU order) or 'reference' (dist_sampler.py:276-299)")
        self.merge_order = merge_order
        from . import _lib
        self._C, self._lib = C, _lib
        self._L = _lib.lib()
        self.sampler = sampler
        self.group = group
        self.rank, self.world_size = dist.get_rank(group), dist.get_world_size(group)
        self.fanouts, self.num_layers = sampler._fanouts, sampler._num_layers
        self.num_snapshots = sampler._num_snapshots
        self.device = torch.device("cuda", sampler._device)
        self.table = None if table is None else table.to(self.device, torch.int8).contiguous()
        h = C.c_void_p()
        _lib.check(self._L.gf_peer_create(sampler._device, self.rank, self.world_size, int(max_targets),
                                          max(self.fanouts), C.byref(h)))
        self._h = h
        mine = (C.c_char * 64)()
        _lib.check(self._L.gf_peer_export(self._h, mine))
        handles = [None] * self.world_size
        dist.all_gather_object(handles, bytes(mine), group=group)
        blob = b"".join(handles)
        _lib.check(self._L.gf_peer_connect(self._h, blob))
        dist.barrier(group=group)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)  # nobody may still be writing into this rank's window
            self._L.gf_peer_destroy(self._h)
            self._h = None

    def sample_layer(self, nodes: torch.Tensor, ts: torch.Tensor, layer: int, snapshot: int):
        """-> dict(all_nodes, all_timestamps, delta_timestamps, eids, row, col, num_dst_nodes, num_src_nodes)"""
        C, s = self._C, self.sampler
        nodes = nodes.to(self.device, torch.int64).contiguous()
        ts = ts.to(self.device, torch.float32).contiguous()
        T = nodes.shape[0]
        pool, offs, arr = s._alloc_steps([(T, self.fanouts[layer])])
        r = arr[0]
        tab = self.table
        self._lib.check(self._L.gf_sampler_sample_layer_partitioned(
            s._h, self._h, nodes.data_ptr(), ts.data_ptr(), T, tab.data_ptr() if tab is not None else None,
            tab.shape[0] if tab is not None else 0, layer, snapshot, C.byref(r),
            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        S = int(r.num_edges)
        v = s._views(pool.view(torch.int64), pool.view(torch.float32), offs[0], T, S)
        out = dict(all_nodes=v[0], all_timestamps=v[1], delta_timestamps=v[2], eids=v[3], row=v[4], col=v[5],
                   num_dst_nodes=T, num_src_nodes=T + S)
        if self.merge_order == "reference":
            if tab is not None:
                n = tab.shape[0]
                inside = (nodes >= 0) & (nodes < n)
                own = torch.full_like(nodes, -1)
                own[inside] = tab[nodes[inside]].to(torch.int64)
            else:
                own = owner_of(nodes, self.world_size)
            out = to_reference_order(out, own)
        return out

    def sample(self, nodes: torch.Tensor, ts: torch.Tensor):
        """[layer][snapshot] results, layers reversed like TemporalSampler.sample (temporal_sampler.py:163-164)."""
        results = []
        for layer in range(self.num_layers):
            lay = []
            for sn in range(self.num_snapshots):
                if layer == 0:
                    n_in, t_in = nodes, ts
                else:
                    n_in, t_in = results[-1][sn]["all_nodes"], results[-1][sn]["all_timestamps"]
                lay.append(self.sample_layer(n_in, t_in, layer, sn))
            results.append(lay)
        results.reverse()
        return results


class PeerFeatureStore:
    """Feature rows partitioned over the GPUs of one box; `fetch(ids)` returns feats[ids] with the rows of other ranks
    read over NVLink by the gather kernel itself (C ABI: gf_shared_*, gf_gather_rows_partitioned).  Replaces the
    RPC pull of the reference's KVStoreClient (gnnflow/distributed/kvstore.py:251-394).

    owner: int8[num_items], rank that holds each row (-1: nobody, fetched as zeros) -- the partition table for node
    features; for edge features the reference keys an edge by its source vertex's partition (kvstore.py:300-308).
    local_rows: float32[(owner == rank).sum(), dim], this rank's rows in ascending id order."""

    def __init__(self, local_rows: torch.Tensor, owner: torch.Tensor, device: int, group=None):
        import ctypes as C
        from . import _lib
        self._C, self._lib, self._L = C, _lib, _lib.lib()
        self.group = group
        self.rank, self.world_size = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", device)
        self.dim = int(local_rows.shape[1])
        owner = owner.to(self.device, torch.int8).contiguous()
        self.owner = owner
        self.num_items = int(owner.shape[0])
        # position of every id among the ids of its owner (the same on every rank)
        li = torch.zeros(self.num_items, dtype=torch.int32, device=self.device)
        for r in range(self.world_size):
            m = owner == r
            li[m] = torch.arange(int(m.sum()), dtype=torch.int32, device=self.device)
        self.local_index = li
        n_local = int((owner == self.rank).sum())
        if local_rows.shape[0] != n_local:
            raise ValueError("local_rows has {} rows, this rank owns {}".format(local_rows.shape[0], n_local))
        nbytes = max(256, n_local * self.dim * 4)
        p = C.c_void_p()
        _lib.check(self._L.gf_shared_alloc(device, nbytes, C.byref(p)))
        self._own = p
        if n_local:
            rows = local_rows.to(self.device, torch.float32).contiguous()
            ar = torch.arange(n_local, dtype=torch.int64, device=self.device)
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self._L.gf_gather_rows(ar.data_ptr(), n_local, n_local, rows.data_ptr(), self.dim, p, None, stream))  # D2D copy
            torch.cuda.current_stream(self.device).synchronize()
        mine = (C.c_char * 64)()
        _lib.check(self._L.gf_shared_export(self._own, mine))
        handles = [None] * self.world_size
        dist.all_gather_object(handles, bytes(mine), group=group)
        self._peers, ptrs = [], []
        for r in range(self.world_size):
            if r == self.rank:
                ptrs.append(self._own.value)
                continue
            q = C.c_void_p()
            _lib.check(self._L.gf_shared_open(device, handles[r], C.byref(q)))
            self._peers.append(q)
            ptrs.append(q.value)
        self._shards = torch.tensor(ptrs, dtype=torch.int64, device=self.device)  # device array of row-table pointers
        dist.barrier(group=group)

    def fetch(self, ids: torch.Tensor) -> torch.Tensor:
        ids = ids.to(self.device, torch.int64).contiguous()
        out = torch.empty(ids.shape[0], self.dim, dtype=torch.float32, device=self.device)
        self._lib.check(self._L.gf_gather_rows_partitioned(
            ids.data_ptr(), ids.shape[0], self.owner.data_ptr(), self.local_index.data_ptr(), self.num_items,
            self._shards.data_ptr(), self.world_size, self.dim, out.data_ptr(),
            self._C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return out

    def close(self):
        if getattr(self, "_own", None) is not None:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)  # nobody may still be reading this rank's shard
            for q in self._peers:
                self._L.gf_shared_close(q)
            self._L.gf_shared_free(self._own)
            self._own, self._peers = None, []
