"""gf_unique_inverse vs torch.unique (device) vs the reference's flow (ids.cpu() + torch.unique on the host)."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from gnnflow_b200 import unique_inverse
rows = []
for n, N in ((19800, 10984), (217800, 10984), (217800, 16_700_000), (4_000_000, 16_700_000)):
    ids = torch.randint(0, N, (n,), device="cuda")
    def t_dev(fn, reps=20):
        for _ in range(3): fn()
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps * 1e3
    ours = t_dev(lambda: unique_inverse(ids, N))
    tdev = t_dev(lambda: torch.unique(ids, return_inverse=True))
    t0 = time.perf_counter()
    for _ in range(5): torch.unique(ids.cpu(), return_inverse=True)
    host = (time.perf_counter() - t0) / 5 * 1e6
    rows.append({"n": n, "num_items": N, "gf_unique_inverse_us": ours, "torch_unique_cuda_us": tdev, "reference_flow_host_us": host})
print(json.dumps({"experiment": "sorted unique + inverse map of MFG source ids", "rows": rows}))
