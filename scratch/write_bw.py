import torch, time
dev = torch.device("cuda")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
x = torch.empty(400_000_000 // 8, dtype=torch.int64, device=dev)
y = torch.empty_like(x)
ms = t(lambda: x.fill_(7)); print("fill 400MB: %.3f ms -> %.1f GB/s write" % (ms, 0.4 / ms * 1e3))
ms = t(lambda: y.copy_(x)); print("copy 400MB: %.3f ms -> %.1f GB/s r+w" % (ms, 0.8 / ms * 1e3))
ms = t(lambda: x.sum()); print("sum 400MB: %.3f ms -> %.1f GB/s read" % (ms, 0.4 / ms * 1e3))
z = torch.empty(1 << 28, dtype=torch.int64, device=dev)
ms = t(lambda: z.fill_(7)); print("fill 2GiB: %.3f ms -> %.1f GB/s write" % (ms, z.numel() * 8 / ms / 1e6))
