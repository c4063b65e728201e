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
print("prefilter (band)  %.3f ms" % t(lambda: plan.pre.run(x64, False, 0.0)))
Hf, Wf = plan.pre._out_hw(360, 181)
print("prefilter (LU)    %.3f ms" % t(lambda: plan._coefficients_lu(x64, Hf, Wf)))
from xarray_regrid_b200._lib import lib
cc = plan._coefficients_lu(x64, Hf, Wf)
st = torch.cuda.current_stream().cuda_stream
for axis, name in ((0, "rows"), (1, "cols")):
    e = plan._lu[name]
    print("  solve axis %d      %.3f ms" % (axis, t(lambda: lib.rg_spline_solve_f64(cc.data_ptr(), 64, Hf * Wf, Wf, Hf, Wf, axis, e[1].data_ptr(), e[2].data_ptr(), st))))
flag = torch.zeros(1, dtype=torch.int32, device="cuda")
print("nan flag          %.3f ms" % t(lambda: lib.rg_nan_flag_f64(x64.data_ptr(), 64, 181, 360, 181 * 360, 360, flag.data_ptr(), st)))
print("poison            %.3f ms" % t(lambda: lib.rg_poison_if_f64(y.data_ptr(), y.numel(), flag.data_ptr(), st)))
print("evaluate          %.3f ms" % t(lambda: plan.op.run(c, False, 0.0, out=y)))
print("whole plan        %.3f ms" % t(lambda: plan(x)))
