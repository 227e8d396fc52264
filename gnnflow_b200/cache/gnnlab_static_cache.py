"""GNNLab static cache (reference gnnflow/cache/gnnlab_static_cache.py:13-182; paper doi 10.1145/3492321.3519557):
a few pre-sampling rounds over the training batches count how often every vertex / edge is sampled, the most
frequently sampled rows are loaded once, and the cache is never updated afterwards.

The statistics stay on the device: `gf_cache_count_distinct` per sampled block (the reference's non-accumulating
`count[ids] += 1` on CPU tensors, :104-111) and `gf_cache_fill_topk` (its torch.topk + index_put, :130-168; ties
-> lowest id, which torch.topk leaves unspecified)."""
from typing import Optional, Union

import numpy as np
import torch

from .._lib import CacheStateC, check
from .cache import Cache


def get_batch_no_neg(df, batch_size: int):
    """gnnflow/utils.py:398-410 (positives only: src || dst, their timestamps, the batch's edge ids)"""
    indices = np.array(df.index // batch_size)
    for _, rows in df.groupby(indices):
        target_nodes = np.concatenate([rows.src.values, rows.dst.values]).astype(np.int64)
        ts = np.concatenate([rows.time.values, rows.time.values]).astype(np.float32)
        yield target_nodes, ts, rows['eid'].values


class GNNLabStaticCache(Cache):
    def __init__(self, cache_ratio: int, num_nodes: int, num_edges: int,
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
        # the reference forwards its arguments to Cache.__init__ shifted by one (cache_ratio lands in
        # edge_cache_ratio and num_nodes in node_cache_ratio, :51-58) and cannot be constructed; the evident intent
        # -- one ratio for both caches -- is what is implemented
        super(GNNLabStaticCache, self).__init__(cache_ratio, cache_ratio, num_nodes, num_edges, device, node_feats,
                                                edge_feats, dim_node_feat, dim_edge_feat, pinned_nfeat_buffs,
                                                pinned_efeat_buffs, kvstore_client, distributed, neg_sample_ratio)
        self.name = 'gnnlab'
        self.cache_index_to_node_id = None
        self.cache_index_to_edge_id = None
        self.node_sampled_count = None
        self.edge_sampled_count = None

    def reset(self):
        """do nothing (:66-69)"""
        return

    def get_mem_size(self) -> int:
        mem_size = 0
        for kind in ("node", "edge"):
            if getattr(self, "dim_%s_feat" % kind) != 0:
                for t in (getattr(self, "cache_%s_buffer" % kind), getattr(self, "cache_%s_flag" % kind),
                          getattr(self, "cache_%s_map" % kind)):
                    mem_size += t.element_size() * t.nelement()
        return mem_size

    def _state(self, kind: str) -> CacheStateC:
        return CacheStateC(getattr(self, "cache_%s_buffer" % kind).data_ptr(), getattr(self, "cache_%s_flag" % kind).data_ptr(),
                           getattr(self, "cache_%s_map" % kind).data_ptr(), None, None,
                           getattr(self, "%s_capacity" % kind), getattr(self, "num_%ss" % kind),
                           getattr(self, "dim_%s_feat" % kind))

    def _count(self, counts: torch.Tensor, ids: torch.Tensor):
        ids = ids.to(self.device, torch.int64).contiguous()
        check(self._L.gf_cache_count_distinct(ids.data_ptr(), ids.shape[0], counts.data_ptr(), counts.shape[0], self._stream()))

    def presample(self, mfgs):
        """accumulate the sampling statistics of one sampled batch (:100-111)"""
        if self.dim_node_feat != 0:
            for b in mfgs[0]:
                self._count(self.node_sampled_count, b.srcdata['ID'])
        if self.dim_edge_feat != 0:
            for mfg in mfgs:
                for b in mfg:
                    if b.num_src_nodes() > b.num_dst_nodes():
                        self._count(self.edge_sampled_count, b.edata['ID'])

    def fill(self):
        """load the most frequently sampled rows (:113-168)"""
        for kind in ("node", "edge"):
            if getattr(self, "dim_%s_feat" % kind) == 0:
                continue
            st = self._state(kind)
            scratch = self._scratch_of(int(self._L.gf_cache_fill_scratch_bytes(st.num_items)))
            check(self._L.gf_cache_fill_topk(st, getattr(self, "%s_sampled_count" % kind).data_ptr(),
                                             getattr(self, "%s_feats" % kind).data_ptr(), scratch.data_ptr(),
                                             scratch.numel(), self._stream()))

    def init_cache(self, *args, **kwargs):
        """kwargs: sampler, train_df (columns src, dst, time, eid), pre_sampling_rounds=2, batch_size=600 (:87-98)"""
        self.node_sampled_count = torch.zeros(self.num_nodes, dtype=torch.int32, device=self.device)
        self.edge_sampled_count = torch.zeros(self.num_edges, dtype=torch.int32, device=self.device)
        sampler = kwargs['sampler']
        train_df = kwargs['train_df']
        pre_sampling_rounds = kwargs.get('pre_sampling_rounds', 2)
        batch_size = kwargs.get('batch_size', 600)
        for _ in range(pre_sampling_rounds):
            for target_nodes, ts, _ in get_batch_no_neg(train_df, batch_size):
                self.presample(sampler.sample(target_nodes, ts))
        self.fill()

    def fetch_feature(self, mfgs, eid: Optional[np.ndarray] = None, update_cache: bool = True,
                      target_edge_features: bool = True):
        """:170-182: the cache is static"""
        return super(GNNLabStaticCache, self).fetch_feature(mfgs, eid=eid, update_cache=False,
                                                            target_edge_features=target_edge_features)
