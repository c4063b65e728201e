"""Conservative regridding on the ratio 2 - 4 grids, fp64 and fp32, one process, one GPU (`variants`: (TASKS_PER_SM,
MIN_RUN_ROWS) settings of engine.BandOp._ring to A/B; the launch-shape A/Bs of profiles/r2_conservative_other_shapes.md
were run by patching tables.lane_blocks_layout's keyword defaults the same way): python tools/bench_blocks.py"""
import sys
from functools import partial
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from xarray_regrid_b200 import engine, tables

orig = tables.lane_blocks_layout


def t(fn, reps=7):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    e[0].record()
    for i in range(reps):
        fn(); e[i + 1].record()
    torch.cuda.synchronize()
    return sorted(e[i].elapsed_time(e[i + 1]) for i in range(reps))[reps // 2]


r = lambda a: np.round(a, 10)
cases = {
    "0.25->0.5": (r(np.arange(-90, 90.1, 0.25)), r(np.arange(0, 360, 0.25)), np.arange(-90.0, 90.1, 0.5), np.arange(0.0, 360, 0.5), 256),
    "0.5->1": (np.arange(-90, 90.1, 0.5), np.arange(0, 360, 0.5), np.arange(-90.0, 91, 1.0), np.arange(0.0, 360, 1.0), 1024),
    "1->2.5": (np.arange(-90, 90.1, 1.0), np.arange(0, 360, 1.0), np.arange(-90.0, 90.1, 2.5), np.arange(0.0, 360, 2.5), 2048),
    "0.1->0.25": (r(np.arange(-90, 90.05, 0.1)), r(np.arange(0, 359.95, 0.1)), np.arange(-90.0, 90.1, 0.25), np.arange(0.0, 360, 0.25), 48),
}
cases["1->2"] = (np.arange(-90, 90.1, 1.0), np.arange(0, 360, 1.0), np.arange(-90.0, 90.1, 2.0), np.arange(0.0, 360, 2.0), 2048)
variants = {"default": (32, 16)}
cases["0.25->1"] = (r(np.arange(-90, 90.1, 0.25)), r(np.arange(0, 360, 0.25)), np.arange(-90.0, 90.1, 1.0), np.arange(0.0, 360, 1.0), 256)

for name, (lat, lon, tlat, tlon, B) in cases.items():
    x = torch.rand((B, lat.size, lon.size), device="cuda", dtype=torch.float64)
    x[x < 0.05] = float("nan")
    gb = B * (lat.size * lon.size + tlat.size * tlon.size) * 8 / 1e9
    for dt in (torch.float64, torch.float32):
        xx = x if dt == torch.float64 else x.float()
        for vname, kw in variants.items():
            engine.TASKS_PER_SM, engine.MIN_RUN_ROWS = kw
            plan = engine.ConservativePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"))
            ring = plan._ring(B, lat.size, lon.size, dt)
            h = ring.host
            ms = t(lambda: plan(xx, skipna=True, nan_threshold=0.5))
            g = gb * (1 if dt == torch.float64 else 0.5)
            print(f"{name:10s} {str(dt)[6:]:8s} {vname:28s} warps={h.n_warps:2d} panels={h.n_panels} lanes={h.lane_cols} "
                  f"{ms:7.3f} ms {g / ms * 1e3:6.0f} GB/s", flush=True)
    del x
