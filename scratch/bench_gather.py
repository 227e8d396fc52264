"""gather kernel sweep: rows x dim x hit ratio -> algorithmic GB/s (8 id + 1 flag + 8 map + 4D read + 4D write per row) and
parity against torch indexing (bit-exact)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnnflow_b200._lib import lib, check  # noqa: E402
L = lib()
dev = torch.device("cuda")
peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]) if os.path.exists("MEASURED_PEAKS.json") else 6550.0
g = torch.Generator(device=dev); g.manual_seed(0)
rows = []
for D in (172, 186, 413, 768):
    N = 2_000_000 if D < 500 else 1_000_000
    feats = torch.randn(N, D, device=dev, generator=g)
    cap = N // 5
    cached = torch.randperm(N, device=dev, generator=g)[:cap]
    flag = torch.zeros(N, dtype=torch.uint8, device=dev); flag[cached] = 1
    cmap = torch.full((N,), -1, dtype=torch.int64, device=dev); cmap[cached] = torch.arange(cap, device=dev)
    buf = feats[cached].contiguous()
    for n in (18_000, 200_000, 1 << 20, 1 << 22):
        if n * D * 4 > 6e9: continue
        ids = torch.randint(0, N, (n,), device=dev, generator=g)
        out = torch.empty(n, D, device=dev)
        hm = torch.empty(n, dtype=torch.uint8, device=dev)
        nh = torch.zeros(1, dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        call = lambda: check(L.gf_cache_gather(ids.data_ptr(), n, feats.shape[0], flag.data_ptr(), cmap.data_ptr(), buf.data_ptr(), feats.data_ptr(), D, out.data_ptr(), hm.data_ptr(), nh.data_ptr(), None, st))
        call(); torch.cuda.synchronize()
        assert torch.equal(out, feats[ids]) and torch.equal(hm.bool(), flag[ids].bool()) and int(nh.item()) == int(flag[ids].sum().item())
        for _ in range(3): call()
        reps = 20 if n >= (1 << 20) else 200
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        b = n * (17 + 8 * D)
        rows.append({"dim": D, "rows": n, "us": ms * 1e3, "GBps": b / ms / 1e6, "frac": b / ms / 1e6 / peak})
        print(rows[-1], file=sys.stderr)
    del feats, buf
print(json.dumps({"kernel": "cache_gather_kernel", "peak": peak, "hit_ratio": 0.2, "sweep": rows}))
