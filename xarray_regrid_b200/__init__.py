"""xarray_regrid_b200 -- B200-native (sm_100a) regridding engine behind xarray-regrid's API.

Drop-in surface mirrors the reference package (``src/xarray_regrid/__init__.py:1-12``):
``Grid``, ``Regridder``, ``create_regridding_dataset``, ``methods`` and ``__version__``.
Importing the package loads ``csrc/libregrid_b200.so`` and fails loudly when it is missing:
there is no CPU fallback.
"""

from . import _lib, tables  # noqa: F401  (loads the CUDA library; raises if it is not built)
from . import methods  # noqa: E402
from .grid import Grid, InvalidBoundsError, create_regridding_dataset  # noqa: E402
from .labelled import DataArray, Dataset  # noqa: E402
from .regrid import Regridder  # noqa: E402

__all__ = [
    "Grid",
    "Regridder",
    "create_regridding_dataset",
    "methods",
    "DataArray",
    "Dataset",
    "InvalidBoundsError",
]

__version__ = "0.4.2+b200.r1"
