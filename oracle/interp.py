"""Oracle: linear / nearest / cubic regridding (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference does ``data.interp(coords=..., method=...)`` (``methods/interp.py:43-46``).
xarray decomposes an interpolation onto orthogonal 1-D target coordinates into
one ``scipy.interpolate.interp1d(x, y, kind=method, axis=..., bounds_error=False,
fill_value=nan, assume_sorted=True)`` call per axis, applied sequentially.  xarray
is not installed here; scipy (1.18.1) IS, so the oracle calls the same scipy
routine directly -- the arithmetic that defines parity is scipy's:

* nearest: ``searchsorted(x[1:]/2 + x[:-1]/2, t, side="left")`` -> ties go to the LOWER
  neighbour; integer input comes back float64 (``_interpolate.py:336-338, 358-365, 520-535``)
* linear:  ``w_hi*y_hi + w_lo*y_lo`` with float64 weights, NaN taps poison the result even
  at weight 0 (``_interpolate.py:491-518``)
* cubic:   global not-a-knot interpolating spline; ANY NaN in the slab makes the WHOLE
  result NaN (``_interpolate.py:409-437, 555-558``)
* all:     targets outside ``[x[0], x[-1]]`` are NaN (``_interpolate.py:561-593``)

Parity note ("unpinned"): the reference's only tests of these numerics compare against
CDO output files whose inputs are missing (``tests/test_regrid.py:75-96``,
``.MISSING_LARGE_BLOBS``).  The axis order xarray uses is the iteration order of a
Python ``set`` (``interp.py:39``); the oracle fixes it to the order of ``axes``.
"""

from __future__ import annotations

import numpy as np
from scipy.interpolate import interp1d

from . import formatting


def interp_axis(data, src, tgt, axis, method):
    """One orthogonal 1-D interpolation, as xarray's ScipyInterpolator does it."""
    f = interp1d(
        src, data, kind=method, axis=axis, bounds_error=False,
        fill_value=np.nan, assume_sorted=True, copy=False,
    )
    return f(tgt)


def regrid(data, src_coords, tgt_coords, axes, method):
    """``interp_regrid`` (interp.py:24-52) over ``axes`` in the given order.

    xarray sorts the source along each interpolated coordinate first
    (``Dataset.interp`` -> ``sortby``); the target keeps its original order.
    """
    data = np.asarray(data)
    for src, tgt, ax in zip(src_coords, tgt_coords, axes):
        data, src = formatting.ensure_monotonic(data, src, ax)
        data = interp_axis(data, np.asarray(src, dtype=np.float64), np.asarray(tgt), ax, method)
    return data


def regrid_latlon(data, src_lat, src_lon, tgt_lat, tgt_lon, method):
    """``da.regrid.<method>(target)`` for ``[..., lat, lon]`` data with coordinates named
    latitude / longitude: ``format_for_regrid`` (regrid.py:44/63/82) then interp."""
    data, lat, lon = formatting.format_for_regrid(
        np.asarray(data), src_lat, src_lon, tgt_lat, tgt_lon
    )
    return regrid(data, [lat, lon], [tgt_lat, tgt_lon], [-2, -1], method)
