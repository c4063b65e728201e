"""Where the pageable host path loses against the pinned one: staging piece size and memcpy threads."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from xarray_regrid_b200 import engine, synthetic
lat, lon = synthetic.grid_025(); tlat, tlon = synthetic.grid_1()
plan = engine.ConservativePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"))
n = 512
x = synthetic.conservative_stack(n, lat, lon, nan_fraction=0.1).cpu()
xp = x.pin_memory(); xn = x.numpy()
out = torch.empty((n, 181, 360), dtype=torch.float64).pin_memory()
gb = n * (721 * 1440 + 181 * 360) * 8 / 1e9
def t(fn, reps=3):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps
print("pinned in, pinned out: %.1f GB/s" % (gb / t(lambda: plan.apply_host(xp, out_host=out, nan_threshold=0.5))))
for stage_mb in (8, 32, 128):
    for th in (4, 8, 14):
        pipe = engine.HostPipeline(slots=3, copy_threads=th); pipe.stage_bytes = 4 * (stage_mb << 20)
        dt = t(lambda: plan.apply_host(xn, out_host=out, nan_threshold=0.5, pipeline=pipe))
        print(f"pageable in, pinned out, pieces {stage_mb} MiB, {th} threads: {gb / dt:.1f} GB/s", flush=True)
        pipe.close()
