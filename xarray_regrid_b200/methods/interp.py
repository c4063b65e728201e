"""Mirror of src/xarray_regrid/methods/interp.py:24-52 (``interp_regrid``)."""

from ..regrid import Regridder


def interp_regrid(data, target_ds, method):
    if method not in ("linear", "nearest", "cubic"):
        raise ValueError(f"unknown interpolation method {method!r}")
    return getattr(Regridder(data), method)(target_ds, time_dim=None)
