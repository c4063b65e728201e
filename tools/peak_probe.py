"""Probe pure-write, pure-read and copy bandwidth on the box (context for write-bound / read-bound kernels)."""
import torch

n = 1 << 30  # 4 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


w = t(lambda: a.fill_(1.0))
c = t(lambda: b.copy_(a))
r = t(lambda: a.sum())
print(f"write (fill_)  {4 * n / w / 1e6:8.1f} GB/s")
print(f"copy  (copy_)  {8 * n / c / 1e6:8.1f} GB/s (read+write)")
print(f"read  (sum)    {4 * n / r / 1e6:8.1f} GB/s")
