import torch
x = torch.empty(4 * 1024**3 // 4, dtype=torch.float32, device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.zero_()); print(f"write-only (zero_ 4 GiB): {4.295 / ms * 1e3:.0f} GB/s")
ms = t(lambda: x.sum()); print(f"read-only (sum 4 GiB): {4.295 / ms * 1e3:.0f} GB/s")
y = torch.empty_like(x)
ms = t(lambda: y.copy_(x)); print(f"copy (4 GiB -> 4 GiB): {2 * 4.295 / ms * 1e3:.0f} GB/s read+write")
