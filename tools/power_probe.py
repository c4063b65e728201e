"""How the 1 kW power cap treats a sustained memory-bound load on this GPU: back-to-back STREAM copies for ~0.4 s
vs the conservative kernel (skipna / zero-fill) on the same stack, with per-launch times."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from xarray_regrid_b200 import engine, synthetic
dev = torch.device("cuda", 0)
def series(fn, n):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    torch.cuda.synchronize(); ev[0].record()
    for i in range(n):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev); b = torch.empty_like(a)
t = series(lambda: b.copy_(a), 600)
gb = 2 * a.numel() * 2 / 1e9
print("copy: first 10 %.0f GB/s, launches 100-200 %.0f, 400-600 %.0f (total %.2f s)" % (
    gb / (np.mean(t[:10]) * 1e-3), gb / (np.mean(t[100:200]) * 1e-3), gb / (np.mean(t[400:]) * 1e-3), sum(t) / 1e3))
del a, b
lat, lon = synthetic.grid_025(); tlat, tlon = synthetic.grid_1()
plan = engine.ConservativePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"))
n = 4096
x = synthetic.conservative_stack(n, lat, lon, nan_fraction=0.1)
y = torch.empty((n, 181, 360), dtype=torch.float64, device=dev)
gbk = n * (721 * 1440 + 181 * 360) * 8 / 1e9
for name, kw in (("skipna", dict(skipna=True, nan_threshold=0.5)), ("zero-fill", dict(skipna=False))):
    t = series(lambda: plan(x, out=y, **kw), 60)
    print("%-9s: first 3 %.0f GB/s, launches 10-30 %.0f, 40-60 %.0f (total %.2f s)" % (
        name, gbk / (np.mean(t[:3]) * 1e-3), gbk / (np.mean(t[10:30]) * 1e-3), gbk / (np.mean(t[40:]) * 1e-3), sum(t) / 1e3))
