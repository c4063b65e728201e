"""Method-level entry points with the reference's names (src/xarray_regrid/methods/):
``conservative.conservative_regrid``, ``interp.interp_regrid``, ``flox_reduce.compute_mode`` and
``flox_reduce.statistic_reduce``.  They expect data that ``format_for_regrid`` would have produced in the
reference; here pre-formatting is folded into the engine's tables, so they simply forward to the accessor."""

from . import conservative, flox_reduce, interp  # noqa: F401
