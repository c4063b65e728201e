"""The accessor mirror end to end on the GPU: the reference's accessor-level tests, transcribed
(tests/test_regrid.py, tests/test_reduce.py), plus Dataset / attrs / coordinate-order behaviour."""

import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")

import xarray_regrid_b200 as xrb  # noqa: E402
from oracle import conservative as oc  # noqa: E402
from oracle import interp as oi  # noqa: E402
from xarray_regrid_b200.labelled import DataArray, Dataset  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _target(**coords):
    return Dataset({}, {k: np.asarray(v, dtype=float) for k, v in coords.items()})


def test_conservative_nan_aggregation_over_dims():
    """reference tests/test_regrid.py:172-185."""
    data = np.broadcast_to(np.array([[1, np.nan], [2, 2]]), (3, 2, 2)).copy()
    data[2] = np.nan
    da = DataArray(data, ("time", "x", "y"), {"time": [0, 1, 2], "x": [-1, 1], "y": [-1, 1]})
    result = da.regrid.conservative(_target(x=[0], y=[0]), skipna=True, nan_threshold=1)
    assert result.dims == ("time", "x", "y") and result.shape == (3, 1, 1)
    assert isinstance(result.data, np.ndarray)
    np.testing.assert_allclose(result.data[0].mean(), np.nanmean(data[0]))


@pytest.mark.parametrize("nan_threshold", [0, 1])
def test_conservative_nan_thresholds_against_coarsen(nan_threshold):
    """reference tests/test_regrid.py:188-208."""
    nan = np.nan
    arr = np.array([[1, nan, nan, nan], [2, 2, nan, nan], [3, 3, 3, nan], [4, 4, 4, 4.0]])
    da = DataArray(arr, ("x", "y"), {"x": [0, 1, 2, 3], "y": [0, 1, 2, 3]}, name="foo")
    blocks = arr.reshape(2, 2, 2, 2).transpose(0, 2, 1, 3).reshape(2, 2, 4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        expected = np.nanmean(blocks, axis=-1) if nan_threshold else blocks.mean(axis=-1)
    out = da.regrid.conservative(_target(x=[0.5, 2.5], y=[0.5, 2.5]), skipna=True, nan_threshold=nan_threshold)
    assert out.name == "foo"
    np.testing.assert_allclose(out.data, expected)


LC = np.array(
    [
        [2, 2, 2, 2, 0, 0, 0, 0, 0, 0, 0],
        [2, 2, 0, 2, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [3, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1],
        [3, 3, 3, 3, 0, 0, 0, 0, 1, 1, 1],
        [3, 3, 0, 3, 0, 0, 0, 0, 1, 1, 1],
    ]
)


def _lc(dtype="uint8"):
    c = np.linspace(0, 40, num=11)
    return DataArray(LC.astype(dtype), ("longitude", "latitude"), {"longitude": c, "latitude": c},
                     attrs={"test": "testing"},
                     coord_attrs={"longitude": {"units": "degrees_east"}, "latitude": {"units": "degrees_north"}},
                     name="lc")


def _grid8(lo, hi):
    return xrb.Grid(north=hi, east=hi, south=lo, west=lo, resolution_lat=8, resolution_lon=8).create_regridding_dataset()


def test_most_common_dims_lon_lat():
    """reference tests/test_reduce.py:87-110: data laid out (longitude, latitude), uint8 preserved."""
    expected = np.array([[2, 2, 0, 0, 0, 0], [0] * 6, [0] * 6, [0] * 6, [0] * 6, [3, 3, 0, 0, 0, 1]], dtype="uint8")
    out = _lc().regrid.most_common(_grid8(0, 40), values=np.array([0, 1, 2, 3]))
    assert out.dims == ("longitude", "latitude") and out.dtype == np.uint8
    np.testing.assert_array_equal(out.data, expected)
    np.testing.assert_array_equal(out.coords["latitude"], np.linspace(0, 40, 6))
    # attrs: data's own attrs, the TARGET's coordinate attrs (tests/test_reduce.py:156-164)
    assert out.attrs == {"test": "testing"}
    assert out.coord_attrs["longitude"] == {"units": "degrees_east"}
    _lc().regrid.least_common(_grid8(0, 40), values=np.array([0, 1, 2, 3]))  # tests/test_reduce.py:113-118


def test_oversized_most_common_and_reversed_target():
    """reference tests/test_reduce.py:121-153 and :189-199."""
    with pytest.warns(UserWarning, match="fill_value"):
        out = _lc("int64").regrid.most_common(_grid8(-8, 48), values=np.array([0, 1, 2, 3]))
    assert out.data.dtype == np.float64 and np.isnan(out.data[0]).all() and np.isnan(out.data[:, -1]).all()
    assert out.data[1, 1] == 2 and out.data[6, 6] == 1
    tgt = _grid8(0, 40)
    fwd = _lc().regrid.most_common(tgt, values=np.array([0, 1, 2, 3]))
    tgt.coords["latitude"] = tgt.coords["latitude"][::-1].copy()
    rev = _lc().regrid.most_common(tgt, values=np.array([0, 1, 2, 3]))
    np.testing.assert_array_equal(rev.coords["latitude"], tgt.coords["latitude"])
    np.testing.assert_array_equal(rev.data, fwd.data[:, ::-1])


def test_stat_var_and_min():
    """reference tests/test_reduce.py:202-235."""
    var = np.array([[0, 0, 1.0, 0, 0, 0], [1.0, 0.75, 0.75, 0, 0, 0], [0] * 6, [0] * 6,
                    [2.25, 0, 0, 0, 0, 0.25], [0, 1.6875, 2.25, 0, 0.25, 0]])
    out = _lc(float).regrid.stat(_grid8(0, 40), "var")
    np.testing.assert_array_equal(out.data, var)
    out = _lc(float).regrid.stat(_grid8(0, 40), "min")
    assert out.data[0, 0] == 2.0 and out.data[5, 0] == 3.0 and out.data[5, 5] == 1.0


@pytest.mark.parametrize("method", ["linear", "nearest", "cubic", "conservative"])
def test_dataset_attrs_and_coord_order(method):
    """reference tests/test_regrid.py:139-169 (attrs) and :257-300 (coords equal the target's, also reversed).
    ERA5-like: latitude stored descending, time leading, a variable without spatial dims passes through."""
    rng = np.random.default_rng(3)
    lat, lon = np.arange(90.0, 44.9, -0.25), np.arange(90.0, 180.1, 0.25)
    d2m = DataArray(280 + rng.random((2, lat.size, lon.size)), ("time", "latitude", "longitude"),
                    {"time": np.arange(2), "latitude": lat, "longitude": lon}, attrs={"units": "K"},
                    coord_attrs={"longitude": {"long_name": "longitude"}}, name="d2m")
    ds = Dataset({"d2m": d2m, "scalar_series": DataArray(np.arange(2.0), ("time",), {"time": np.arange(2)})},
                 attrs={"Conventions": "CF-1.6"})
    tgt = xrb.Grid(north=90, east=180, south=45, west=90, resolution_lat=0.17, resolution_lon=0.17).create_regridding_dataset()
    for reverse in (False, True):
        if reverse:
            tgt.coords["latitude"] = tgt.coords["latitude"][::-1].copy()
        fn = getattr(ds.regrid, method)
        out = fn(tgt, latitude_coord="latitude") if method == "conservative" else fn(tgt)
        assert out.attrs == ds.attrs and out["d2m"].attrs == {"units": "K"}
        assert out.coord_attrs["longitude"] == {"long_name": "longitude"}
        np.testing.assert_array_equal(out["d2m"].coords["latitude"], tgt.coords["latitude"])
        np.testing.assert_array_equal(out["d2m"].coords["longitude"], tgt.coords["longitude"])
        assert out["d2m"].dims == ("time", "latitude", "longitude")
        np.testing.assert_array_equal(out["scalar_series"].data, np.arange(2.0))
        if method == "conservative":
            want = oc.regrid_latlon(d2m.data, lat, lon, tgt.coords["latitude"], tgt.coords["longitude"])
        else:
            want = oi.regrid_latlon(d2m.data, lat, lon, tgt.coords["latitude"], tgt.coords["longitude"], method)
        got = out["d2m"].data
        np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
        np.testing.assert_allclose(got, want, rtol=1e-11)


def test_single_coordinate_and_cuda_input():
    """Only one shared coordinate; CUDA tensors stay on the device."""
    rng = np.random.default_rng(4)
    x = rng.random((5, 30, 7))
    da = DataArray(torch.from_numpy(x).cuda(), ("time", "y", "member"),
                   {"time": np.arange(5), "y": np.arange(30.0), "member": np.arange(7)})
    tgt = _target(y=np.arange(1.0, 28.0, 2.0))
    out = da.regrid.linear(tgt)
    assert out.data.is_cuda and out.shape == (5, 14, 7)
    want = oi.regrid(x, [np.arange(30.0)], [tgt.coords["y"]], [1], "linear")
    np.testing.assert_allclose(out.data.cpu().numpy(), want, rtol=1e-13)
    out = da.regrid.conservative(tgt)
    want = oc.regrid(x, [np.arange(30.0)], [tgt.coords["y"]], [1])
    np.testing.assert_allclose(out.data.cpu().numpy(), want, rtol=1e-12)
    lon_only = DataArray(x[0, :, 0], ("longitude",), {"longitude": np.arange(0.0, 360.0, 12.0)})
    out = lon_only.regrid.nearest(_target(longitude=np.arange(0.0, 360.0, 5.0)))
    want = oi.regrid_latlon(x[0, :, 0][None, :], None, np.arange(0.0, 360.0, 12.0), None,
                            np.arange(0.0, 360.0, 5.0), "nearest") if False else None
    assert out.shape == (72,) and not np.isnan(out.data).any()  # wrap padding makes 355 deg reachable
