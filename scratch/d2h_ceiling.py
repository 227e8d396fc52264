"""Platform ceiling of the end-to-end leg at N GPUs: every rank copies what one bench step ships --
24 MB host->device, 355 MB device->host -- between PINNED host memory and its GPU with plain cudaMemcpyAsync, all ranks at
once.  Run under torchrun; rank 0 prints one JSON line (per-GPU and aggregate GB/s, max over ranks of the time)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
H2D, D2H = 24_217_068, 355_496_592  # bytes per step and rank of bench.py's e2e leg (REDDIT replay)
h_in, h_out = torch.empty(H2D, dtype=torch.uint8).pin_memory(), torch.empty(D2H, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(H2D, dtype=torch.uint8, device=dev), torch.empty(D2H, dtype=torch.uint8, device=dev)


def step():
    d_in.copy_(h_in, non_blocking=True)
    h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()


for _ in range(3):
    step()
res = {}
for name, fn in (("h2d+d2h", step), ("d2h_only", lambda: (h_out.copy_(d_out, non_blocking=True), torch.cuda.synchronize()))):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    dt = (time.perf_counter() - t0) / 10
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    b = (H2D + D2H) if name == "h2d+d2h" else D2H
    res[name] = {"ms_per_step_max_over_ranks": float(t) * 1e3, "GBps_per_gpu": b / float(t) / 1e9, "GBps_aggregate": b * world / float(t) / 1e9}
if rank == 0:
    print(json.dumps({"n_gpus": world, "h2d_bytes": H2D, "d2h_bytes": D2H, "host_cores": os.cpu_count(), **res}))
if world > 1:
    dist.destroy_process_group()
