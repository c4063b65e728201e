"""Device-resident timing of the non-headline ops at their BASELINE config shapes (C3, C4, C5).

    python tools/bench_ops.py [--only linear,cubic,nearest,mode,stat] [--json out.json]

CUDA events on the launch stream, 2 warm-up + 5 timed launches; GB/s = algorithmic bytes / time.
"""
import argparse
import json
import sys
import warnings
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from xarray_regrid_b200 import engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="linear,cubic,nearest,mode,stat")
ap.add_argument("--json", default=None)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
only = set(a.only.split(","))
PEAK = 6557.1
try:
    PEAK = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass
results = {}


def timeit(name, fn, algo_bytes, cells, note=""):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
    ev[0].record()
    for i in range(a.reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(a.reps))
    med = ms[len(ms) // 2]
    gbs = algo_bytes / (med * 1e-3) / 1e9
    results[name] = {"ms": med, "GB/s": gbs, "frac_of_measured_hbm_peak": gbs / PEAK,
                     "Gcells/s": cells / (med * 1e-3) / 1e9, "algorithmic_bytes": algo_bytes, "note": note}
    print(f"{name:34s} {med:9.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / PEAK:5.1f}% of peak  "
          f"{cells / (med * 1e-3) / 1e9:7.1f} Gcells/s  {note}", flush=True)


lat1, lon1 = np.arange(-90.0, 91.0, 1.0), np.arange(0.0, 360.0, 1.0)
lat01, lon01 = np.round(np.arange(-90.0, 90.05, 0.1), 10), np.round(np.arange(0.0, 359.95, 0.1), 10)
lat025, lon025 = np.arange(-90.0, 90.25, 0.25), np.arange(0.0, 360.0, 0.25)
g = torch.Generator(device="cuda").manual_seed(7)

if "linear" in only or "cubic" in only:
    B = 128
    x = (250 + 50 * torch.rand((B, 181, 360), device="cuda", generator=g, dtype=torch.float32))
    per_slab_f32 = 181 * 360 * 4 + 1801 * 3600 * 4
    if "linear" in only:
        plan = engine.InterpPlan(engine.AxisSpec(lat1, lat01, "lat"), engine.AxisSpec(lon1, lon01, "lon"), "linear")
        timeit("C3 linear fp32->fp32 (128 slabs)", lambda: plan(x, out_dtype="same"), B * per_slab_f32,
               B * 1801 * 3600, "target cells/s")
        timeit("C3 linear fp32->fp64 (reference dtype)", lambda: plan(x[:64]), 64 * (181 * 360 * 4 + 1801 * 3600 * 8),
               64 * 1801 * 3600, "target cells/s")
    if "cubic" in only:
        plan = engine.InterpPlan(engine.AxisSpec(lat1, lat01, "lat"), engine.AxisSpec(lon1, lon01, "lon"), "cubic")
        timeit("C3 cubic fp32->fp64 (64 slabs)", lambda: plan(x[:64]), 64 * (181 * 360 * 4 + 1801 * 3600 * 8),
               64 * 1801 * 3600, "target cells/s")
    del x

if "nearest" in only:
    B = 73
    x = torch.randint(0, 32, (B, 1801, 3600), device="cuda", generator=g, dtype=torch.int32)
    plan = engine.InterpPlan(engine.AxisSpec(lat01, lat025, "lat"), engine.AxisSpec(lon01, lon025, "lon"), "nearest")
    timeit("C4 nearest int32->int32 (73 steps)", lambda: plan(x, out_dtype="same"), B * 721 * 1440 * 8,
           B * 1801 * 3600, "source cells/s; bytes = out cells x 8")
    timeit("C4 nearest int32->fp64 (reference dtype)", lambda: plan(x), B * 721 * 1440 * 12, B * 1801 * 3600,
           "source cells/s; bytes = out cells x 12")
    del x

if "mode" in only or "stat" in only:
    lat, lon = np.round(np.arange(-89.995, 90, 0.01), 3), np.round(np.arange(0.005, 360, 0.01), 3)
    tlat, tlon = np.arange(-89.875, 90, 0.25), np.arange(0.125, 360, 0.25)
    plan = engine.ReducePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"))
    if "mode" in only:
        coarse = torch.randint(0, 32, (18000 // 64 + 1, 36000 // 64 + 1), device="cuda", generator=g)
        x = coarse.repeat_interleave(64, 0).repeat_interleave(64, 1)[:18000, :36000].to(torch.uint8).contiguous()
        noise = torch.rand((18000, 36000), device="cuda", generator=g) < 0.05
        x[noise] = torch.randint(0, 32, (int(noise.sum()),), device="cuda", generator=g, dtype=torch.uint8)
        del noise, coarse
        vals = np.arange(32)
        timeit("C5 most_common uint8 18000x36000", lambda: plan.mode(x, vals), 18000 * 36000 + 720 * 1440,
               18000 * 36000, "source cells/s")
        del x
    if "stat" in only:
        x = torch.randn((18000, 36000), device="cuda", generator=g, dtype=torch.float32)
        x[torch.rand((18000, 36000), device="cuda", generator=g) < 0.02] = float("nan")
        for method in ("mean", "var", "median"):
            timeit(f"C5 stat {method} fp32 skipna", lambda m=method: plan.stat(x, m, skipna=True),
                   18000 * 36000 * 4 + 720 * 1440 * 4, 18000 * 36000, "source cells/s")
        del x

if a.json:
    Path(a.json).write_text(json.dumps(results, indent=1))
