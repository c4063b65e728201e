"""Time the stages of the cubic regrid separately (C3 shape, 64 slabs, fp64)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from xarray_regrid_b200 import engine

lat1, lon1 = np.arange(-90.0, 91.0, 1.0), np.arange(0.0, 360.0, 1.0)
lat01, lon01 = np.round(np.arange(-90.0, 90.05, 0.1), 10), np.round(np.arange(0.0, 359.95, 0.1), 10)
plan = engine.InterpPlan(engine.AxisSpec(lat1, lat01, "lat"), engine.AxisSpec(lon1, lon01, "lon"), "cubic")
x = 250 + 50 * torch.rand((64, 181, 360), device="cuda", dtype=torch.float32)
x64 = x.double()
print("prefilter taps", plan.pre.rows_band.k_max, plan.pre.cols_band.k_max, "eval taps", plan.op.rows_band.k_max)

def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    e[0].record()
    for i in range(reps):
        fn(); e[i + 1].record()
    torch.cuda.synchronize()
    return sorted(e[i].elapsed_time(e[i + 1]) for i in range(reps))[reps // 2]

c = plan.pre.run(x64, False, 0.0)
y = torch.empty((64, 1801, 3600), dtype=torch.float64, device="cuda")
print("cast fp32->fp64   %.3f ms" % t(lambda: x.double()))
print("prefilter         %.3f ms" % t(lambda: plan.pre.run(x64, False, 0.0)))
print("evaluate          %.3f ms" % t(lambda: plan.op.run(c, False, 0.0, out=y)))
print("whole plan        %.3f ms" % t(lambda: plan(x)))
