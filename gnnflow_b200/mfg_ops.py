"""Device-resident hand-off of a sampled MFG into the models.

The reference's `Memory.prepare_input` (gnnflow/models/modules/memory.py:156-190) copies `b.srcdata['ID']` to the
host, runs `torch.unique(return_inverse=True)` there, indexes the memory / mailbox tables with the unique ids and
expands them again with the inverse map.  Here the ids never leave the GPU:

  * `unique_inverse(ids, num_items)`  -- sorted distinct ids + inverse map by bitmap ranking (gf_unique_inverse): what
    the partitioned path needs to pull every remote row once;
  * `prepare_memory_input(b, ...)`    -- the non-partitioned branch: mem[inv] of mem = table[unique] is table[ids],
    so the four tables are gathered by `ids` directly (gf_gather_rows for the 2-D float tables).

Values are bit-identical to the reference's (`tests/test_gpu_mfg_ops.py`)."""
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_scratch = {}


def unique_inverse(ids: torch.Tensor, num_items: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """torch.unique(ids, return_inverse=True) for non-negative int64 ids < num_items on a CUDA device.
    Returns (unique ascending int64[k], inverse int64[n]) with unique[inverse] == ids."""
    if not ids.is_cuda:
        raise ValueError("unique_inverse: ids must be a CUDA tensor (there is no CPU path)")
    L = _lib.lib()
    ids = ids.to(torch.int64).contiguous().view(-1)
    n = ids.shape[0]
    dev = ids.device
    if n == 0:
        return ids.new_empty(0), ids.new_empty(0)
    if num_items is None:
        num_items = int(ids.max().item()) + 1
    need = int(L.gf_unique_scratch_bytes(num_items))
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < need:
        buf = _scratch[key] = torch.empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=dev)
    uniq = torch.empty(min(n, num_items), dtype=torch.int64, device=dev)
    inv = torch.empty(n, dtype=torch.int64, device=dev)
    cnt = torch.empty(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(L.gf_unique_inverse(ids.data_ptr(), n, num_items, uniq.data_ptr(), inv.data_ptr(), cnt.data_ptr(),
                                  buf.data_ptr(), buf.numel(), _stream(dev)))
    return uniq[:int(cnt.item())], inv


def gather_rows(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """table[ids] for a CUDA float32 [N, D] table through the vectorised gather kernel; any other table falls back to
    torch indexing on the device (1-D timestamp tables)."""
    if table.is_cuda and table.dtype == torch.float32 and table.dim() == 2 and table.is_contiguous():
        ids = ids.to(torch.int64).contiguous()
        out = torch.empty(ids.shape[0], table.shape[1], dtype=torch.float32, device=table.device)
        if ids.shape[0]:
            with torch.cuda.device(table.device):
                check(_lib.lib().gf_gather_rows(ids.data_ptr(), ids.shape[0], table.shape[0], table.data_ptr(),
                                                table.shape[1], out.data_ptr(), None, _stream(table.device)))
        return out
    return table[ids]


def prepare_memory_input(b, node_memory: torch.Tensor, node_memory_ts: torch.Tensor, mailbox: torch.Tensor,
                         mailbox_ts: torch.Tensor):
    """Memory.prepare_input (memory.py:156-190, non-partitioned branch) without the host round trip: fills
    b.srcdata['mem'], ['mem_ts'], ['mail_ts'], ['mem_input'] for the block's source nodes."""
    ids = b.srcdata['ID']
    mb = mailbox.reshape(mailbox.shape[0], -1) if mailbox.dim() > 2 else mailbox
    b.srcdata['mem'] = gather_rows(node_memory, ids)
    b.srcdata['mem_ts'] = gather_rows(node_memory_ts, ids)
    b.srcdata['mail_ts'] = gather_rows(mailbox_ts, ids)
    mail = gather_rows(mb, ids)
    b.srcdata['mem_input'] = mail.reshape((ids.shape[0],) + tuple(mailbox.shape[1:])) if mailbox.dim() > 2 else mail
    return b
