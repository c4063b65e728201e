"""Small driver for ncu: a few launches of the conservative kernel on a C2-shaped stack.

    python tools/prof_conservative.py [--steps 256] [--no-ring] [--no-lanes] [--launches 3] [--dtype f64]
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

from xarray_regrid_b200 import engine, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=256)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--no-ring", action="store_true")
ap.add_argument("--no-lanes", action="store_true")
ap.add_argument("--no-skipna", action="store_true")
ap.add_argument("--dtype", default="f64")
ap.add_argument("--nan", type=float, default=0.10)
ap.add_argument("--src", type=float, default=0.25, help="source resolution in degrees")
ap.add_argument("--tgt", type=float, default=1.0, help="target resolution in degrees")
a = ap.parse_args()

import numpy as np  # noqa: E402

lat, lon = np.round(np.arange(-90.0, 90.0 + a.src / 2, a.src), 10), np.round(np.arange(0.0, 360.0 - a.src / 2, a.src), 10)
tlat, tlon = np.round(np.arange(-90.0, 90.0 + a.tgt / 2, a.tgt), 10), np.round(np.arange(0.0, 360.0 - a.tgt / 2, a.tgt), 10)
plan = engine.ConservativePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"))
plan.use_ring = not a.no_ring
plan.use_lanes = not a.no_lanes
dtype = torch.float64 if a.dtype == "f64" else torch.float32
x = synthetic.conservative_stack(a.steps, lat, lon, dtype=dtype, nan_fraction=a.nan)
y = torch.empty((a.steps, tlat.size, tlon.size), dtype=dtype, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.launches + 1)]
ev[0].record()
for i in range(a.launches):
    plan(x, skipna=not a.no_skipna, nan_threshold=0.5, out=y)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.launches)]
gb = a.steps * (lat.size * lon.size + tlat.size * tlon.size) * x.element_size() / 1e9
print("ms per launch:", [round(m, 3) for m in ms], "GB/s:", [round(gb / (m * 1e-3)) for m in ms])
