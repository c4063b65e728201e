"""Scratch: where does the conservative kernel differ from the oracle on a grid pair (NaN placement / values)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from xarray_regrid_b200 import engine
from oracle import conservative as oc
lat01, lon01 = np.round(np.arange(-90.0, 90.05, 0.1), 10), np.round(np.arange(0.0, 359.95, 0.1), 10)
lat025, lon025 = np.arange(-90.0, 90.25, 0.25), np.arange(0.0, 360.0, 0.25)
cases = {"0.25->0.5": (lat025, lon025, np.arange(-90.0, 90.1, 0.5), np.arange(0.0, 360, 0.5)),
         "0.1->0.25": (lat01, lon01, lat025, lon025)}
g = torch.Generator(device="cuda").manual_seed(7)
for name, (la, lo, tla, tlo) in cases.items():
    x = torch.rand((2, la.size, lo.size), device="cuda", dtype=torch.float64, generator=g) + 0.5
    x[torch.rand(x.shape, device="cuda", generator=g) < 0.05] = float("nan")
    plan = engine.ConservativePlan(engine.AxisSpec(la, tla, "lat"), engine.AxisSpec(lo, tlo, "lon"))
    for thr in (0.5, 0.3):
        y = plan(x, skipna=True, nan_threshold=thr).cpu().numpy()
        want = oc.regrid_latlon(x.cpu().numpy(), la, lo, tla, tlo, skipna=True, nan_threshold=thr)
        nan_diff = np.isnan(y) != np.isnan(want)
        both = ~np.isnan(y) & ~np.isnan(want)
        rel = np.abs(y[both] - want[both]) / np.abs(want[both])
        print(name, "thr", thr, "nan mismatches", int(nan_diff.sum()), "of", y.size, "max rel", rel.max(),
              "rows of mismatches", np.unique(np.nonzero(nan_diff)[1])[:10], flush=True)
        if nan_diff.any():
            b, k, l = [int(v[0]) for v in np.nonzero(nan_diff)]
            # valid mass of that cell, recomputed in numpy from the band
            print("  first mismatch", (b, k, l), "engine", y[b, k, l], "oracle", want[b, k, l])
