"""Conservative regridding on grid pairs other than the headline one: which kernel the library picks (lanes /
two-phase ring / generic) and how it compares with the generic kernel.  Device-resident, CUDA events.

    python tools/bench_shapes.py
"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from xarray_regrid_b200 import engine
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    e[0].record()
    for i in range(reps):
        fn(); e[i + 1].record()
    torch.cuda.synchronize()
    return sorted(e[i].elapsed_time(e[i + 1]) for i in range(reps))[reps // 2]
cases = {
 "0.25->1 (C2)": (np.arange(-90, 90.1, 0.25), np.arange(0, 360, 0.25), np.arange(-90.0, 91, 1.0), np.arange(0.0, 360, 1.0), 256),
 "0.25->1.5 (7 taps)": (np.arange(-90, 90.1, 0.25), np.arange(0, 360, 0.25), np.arange(-90.0, 90.1, 1.5), np.arange(0.0, 360, 1.5), 256),
 "0.1->1 (11 taps, W=3600)": (np.round(np.arange(-90, 90.05, 0.1), 10), np.round(np.arange(0, 359.95, 0.1), 10), np.arange(-90.0, 91, 1.0), np.arange(0.0, 360, 1.0), 48),
 "0.1->0.25 (3-4 taps, W=3600)": (np.round(np.arange(-90, 90.05, 0.1), 10), np.round(np.arange(0, 359.95, 0.1), 10), np.arange(-90.0, 90.1, 0.25), np.arange(0.0, 360, 0.25), 48),
 "0.25->0.5 (3 taps)": (np.arange(-90, 90.1, 0.25), np.arange(0, 360, 0.25), np.arange(-90.0, 90.1, 0.5), np.arange(0.0, 360, 0.5), 256),
 "0.5->1 (3 taps, W=720)": (np.arange(-90, 90.1, 0.5), np.arange(0, 360, 0.5), np.arange(-90.0, 91, 1.0), np.arange(0.0, 360, 1.0), 1024),
 "1->2.5 (non-integer)": (np.arange(-90, 90.1, 1.0), np.arange(0, 360, 1.0), np.arange(-90.0, 90.1, 2.5), np.arange(0.0, 360, 2.5), 2048),
}
for name, (lat, lon, tlat, tlon, B) in cases.items():
    plan = engine.ConservativePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"))
    x = torch.rand((B, lat.size, lon.size), device="cuda", dtype=torch.float64)
    x[x < 0.05] = float("nan")
    gb = B * (lat.size * lon.size + tlat.size * tlon.size) * 8 / 1e9
    ring = plan._ring(B, lat.size, lon.size, torch.float64)
    info = "generic" if ring is None else f"ring warps={ring.host.n_warps} strips={ring.host.n_strips} lanes={ring.host.lane_cols} live={ring.host.max_live}"
    ms = t(lambda: plan(x, skipna=True, nan_threshold=0.5))
    plan.use_ring = False
    ms_g = t(lambda: plan(x, skipna=True, nan_threshold=0.5))
    print(f"{name:32s} ring {ms:8.3f} ms {gb/ms*1e3:8.0f} GB/s | generic {ms_g:8.3f} ms {gb/ms_g*1e3:8.0f} GB/s  rows k={plan.rows_band.k_max} cols k={plan.cols_band.k_max}  {info}", flush=True)
    del x
