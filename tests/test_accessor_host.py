"""Host-side behaviour of the accessor mirror that needs no GPU: Grid validation, target validation
errors, layout decisions."""

import numpy as np
import pytest

import xarray_regrid_b200 as xrb
from xarray_regrid_b200 import regrid
from xarray_regrid_b200.labelled import DataArray, Dataset


def test_exports_match_reference_package():
    """reference src/xarray_regrid/__init__.py:1-12."""
    for name in ("Grid", "Regridder", "create_regridding_dataset", "methods"):
        assert hasattr(xrb, name)
    assert xrb.__version__.startswith("0.4.2")
    assert hasattr(xrb.methods, "conservative") and hasattr(xrb.methods, "interp") and hasattr(xrb.methods, "flox_reduce")


def test_grid_bounds_validation_and_coords(golden_helpers):
    """reference utils.py:29-44 (InvalidBoundsError) and :66-89 (coords; pinned to the reference's output)."""
    with pytest.raises(xrb.InvalidBoundsError):
        xrb.Grid(north=0, east=10, south=10, west=0, resolution_lat=1, resolution_lon=1)
    with pytest.raises(xrb.InvalidBoundsError):
        xrb.Grid(north=10, east=0, south=0, west=10, resolution_lat=1, resolution_lon=1)
    g = golden_helpers
    for i, p in enumerate(g["grid/params"]):
        ds = xrb.create_regridding_dataset(xrb.Grid(*p))
        np.testing.assert_array_equal(ds.coords["latitude"], g[f"grid/{i}/lat"])
        np.testing.assert_array_equal(ds.coords["longitude"], g[f"grid/{i}/lon"])
    ds = xrb.Grid(90, 180, 45, 90, 0.17, 0.17).create_regridding_dataset(lat_name="lat", lon_name="lon")
    assert set(ds.coords) == {"lat", "lon"} and ds.coord_attrs["lat"] == {"units": "degrees_north"}


def _da():
    return DataArray(np.zeros((3, 4, 5)), ("time", "latitude", "longitude"),
                     {"time": np.arange(3), "latitude": np.arange(4.0), "longitude": np.arange(5.0)})


def test_validate_input_errors():
    """reference regrid.py:289-315."""
    da = _da()
    with pytest.raises(ValueError, match="None of the target dims"):
        regrid.validate_input(da, Dataset({}, {"x": np.arange(3.0)}), "time")
    # the time coordinate of the target is dropped first
    tgt = Dataset({}, {"time": np.arange(2), "latitude": np.arange(2.0), "longitude": np.arange(2.0)})
    out = regrid.validate_input(da, tgt, "time")
    assert "time" not in out.coords and set(out.coords) == {"latitude", "longitude"}
    with pytest.raises(ValueError, match="None of the target"):
        regrid.validate_input(da, Dataset({}, {"time": np.arange(2)}), "time")


def test_argument_errors_before_any_gpu_work():
    da = _da()
    tgt = Dataset({}, {"latitude": np.arange(2.0), "longitude": np.arange(2.0)})
    with pytest.raises(ValueError, match="nan_threshold"):
        da.regrid.conservative(tgt, nan_threshold=1.5)  # regrid.py:118-120
    with pytest.raises(ValueError, match="Invalid method"):
        da.regrid.stat(tgt, "mode")  # flox_reduce.py:70-73
    ds = Dataset({"a": da})
    with pytest.raises(ValueError, match="not implemented"):
        ds.regrid.most_common(tgt, values=np.arange(3))  # regrid.py:165-172
    with pytest.raises(ValueError, match="not implemented"):
        ds.regrid.least_common(tgt, values=np.arange(3))


def test_layout_rules():
    da = DataArray(np.zeros((2, 3, 4)), ("longitude", "time", "latitude"),
                   {"longitude": np.arange(2.0), "time": np.arange(3), "latitude": np.arange(4.0)})
    assert regrid._layout(da, ["longitude", "latitude"]) == ("latitude", "longitude")
    assert regrid._layout(da, ["longitude"]) == (None, "longitude")
    assert regrid._layout(da, ["latitude"]) == ("latitude", None)
    xy = DataArray(np.zeros((2, 3)), ("x", "y"), {"x": np.arange(2.0), "y": np.arange(3.0)})
    assert regrid._layout(xy, ["y", "x"]) == ("x", "y")
    assert regrid._layout(xy, ["y", "x"], prefer_rows="y") == ("y", "x")
    assert regrid._layout(xy, ["time"]) is None
