"""Seeded synthetic inputs for the benchmark configs, generated on device.

Used by ``bench.py``, ``__graft_entry__.smoke()`` and the GPU tests so that all of them see
the same data.  Grids follow BASELINE.json's configs: 0.25 degree = 721 x 1440
(lat -90..90, lon 0..359.75), 1 degree = 181 x 360 (lat -90..90, lon 0..359).
"""

from __future__ import annotations

import numpy as np
import torch


def grid_025():
    return np.arange(-90.0, 90.0 + 0.25, 0.25), np.arange(0.0, 360.0, 0.25)


def grid_1():
    return np.arange(-90.0, 91.0, 1.0), np.arange(0.0, 360.0, 1.0)


def continent_mask(lat, lon, coverage=0.09):
    """Static "continent" mask: sin(3 phi) cos(2 lambda) above the quantile giving ``coverage``."""
    phi = np.radians(lat)[:, None]
    lam = np.radians(lon)[None, :]
    field = np.sin(3 * phi) * np.cos(2 * lam)
    q = np.quantile(field, 1.0 - coverage)
    return field > q


def conservative_stack(steps, lat, lon, *, dtype=torch.float64, nan_fraction=0.10, seed=1234,
                       device="cuda", chunk=64):
    """``0.5 + U[0,1)`` stack ``[steps, lat, lon]``; NaNs = static continent mask (~9 %) plus
    1 % iid speckle per cell when ``nan_fraction`` > 0 (so the valid mass spans (0, 1))."""
    gen = torch.Generator(device=device).manual_seed(seed)
    x = torch.empty((steps, lat.size, lon.size), dtype=dtype, device=device)
    static = None
    if nan_fraction > 0:
        static = torch.from_numpy(continent_mask(lat, lon, max(nan_fraction - 0.01, 0.0))).to(device)
    for s in range(0, steps, chunk):
        e = min(s + chunk, steps)
        blk = x[s:e]
        blk.uniform_(0.0, 1.0, generator=gen).add_(0.5)
        if nan_fraction > 0:
            speckle = torch.rand(blk.shape, device=device, generator=gen, dtype=torch.float32) < 0.01
            blk.masked_fill_(speckle | static, float("nan"))
    return x
