"""The oracle against every synthetic known-answer test the reference holds for the
hot path.  Each test cites the reference test it transcribes (inputs and expected
values are re-typed from the test's literals; no reference code is executed here)."""

import warnings

import numpy as np
import pytest

from oracle import conservative, formatting, grid, interp, reduce


# ------------------------------------------------------------------ tests/test_regrid.py
def test_conservative_nan_aggregation_over_dims():
    """reference tests/test_regrid.py:172-185 -- joint (2-D) valid mass gives 1.6667,
    not 1.5 / 1.75; singleton target; an all-NaN time slice stays NaN."""
    data = np.array([[1, np.nan], [2, 2]])
    data = np.broadcast_to(data, (3, 2, 2)).copy()
    data[2] = np.nan
    out = conservative.regrid(
        data, [np.array([-1.0, 1.0])] * 2, [np.array([0.0])] * 2, [-2, -1],
        skipna=True, nan_threshold=1,
    )
    assert out.shape == (3, 1, 1)
    np.testing.assert_allclose(out[0].mean(), np.nanmean(data[0]))
    np.testing.assert_allclose(out[0, 0, 0], 5.0 / 3.0)
    assert np.isnan(out[2]).all()


@pytest.mark.parametrize("nan_threshold", [0, 1])
def test_conservative_nan_thresholds_against_coarsen(nan_threshold):
    """reference tests/test_regrid.py:188-208."""
    nan = np.nan
    da = np.array([[1, nan, nan, nan], [2, 2, nan, nan], [3, 3, 3, nan], [4, 4, 4, 4.0]])
    blocks = da.reshape(2, 2, 2, 2).transpose(0, 2, 1, 3).reshape(2, 2, 4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        expected = np.nanmean(blocks, axis=-1) if nan_threshold else blocks.mean(axis=-1)
    src = np.array([0.0, 1, 2, 3])
    tgt = np.array([0.5, 2.5])
    out = conservative.regrid(
        da, [src, src], [tgt, tgt], [-2, -1], skipna=True, nan_threshold=nan_threshold
    )
    np.testing.assert_allclose(out, expected)


# ------------------------------------------------------------------ tests/test_reduce.py
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
LC_COORD = np.linspace(0, 40, num=11)  # dims are (longitude, latitude) in the reference
LABELS = np.array([0, 1, 2, 3])


def _target(lo, hi):
    lat, lon = grid.lat_lon_coords(hi, hi, lo, lo, 8, 8)
    return lon, lat


def test_most_common():
    """reference tests/test_reduce.py:87-110 (two tie cells resolve to label 0; uint8 kept)."""
    expected = np.array(
        [
            [2, 2, 0, 0, 0, 0],
            [0, 0, 0, 0, 0, 0],
            [0, 0, 0, 0, 0, 0],
            [0, 0, 0, 0, 0, 0],
            [0, 0, 0, 0, 0, 0],
            [3, 3, 0, 0, 0, 1],
        ],
        dtype="uint8",
    )
    tlon, tlat = _target(0, 40)
    # stats=True formatting touches only longitude (axis 0 here) and this regional grid is not padded
    data, _, lon = formatting.format_for_regrid(
        LC.astype("uint8"), None, LC_COORD, None, tlon, lat_axis=1, lon_axis=0, stats=True
    )
    np.testing.assert_array_equal(lon, LC_COORD)
    out = reduce.mode(data, [lon, LC_COORD], [tlon, tlat], [0, 1], LABELS)
    np.testing.assert_array_equal(out, expected)
    assert out.dtype == np.uint8


def test_least_common_runs():
    """reference tests/test_reduce.py:113-118 (only checks that it runs)."""
    tlon, tlat = _target(0, 40)
    out = reduce.mode(LC, [LC_COORD, LC_COORD], [tlon, tlat], [0, 1], LABELS, anti_mode=True)
    assert out.shape == (6, 6)


def test_oversized_most_common():
    """reference tests/test_reduce.py:121-153 (NaN ring, float result, warning)."""
    nan = np.nan
    expected = np.array(
        [
            [nan, nan, nan, nan, nan, nan, nan, nan],
            [nan, 2, 2, 0, 0, 0, 0, nan],
            [nan, 0, 0, 0, 0, 0, 0, nan],
            [nan, 0, 0, 0, 0, 0, 0, nan],
            [nan, 0, 0, 0, 0, 0, 0, nan],
            [nan, 0, 0, 0, 0, 0, 0, nan],
            [nan, 3, 3, 0, 0, 0, 1, nan],
            [nan, nan, nan, nan, nan, nan, nan, nan],
        ]
    )
    tlon, tlat = _target(-8, 48)
    with pytest.warns(UserWarning, match="fill_value"):
        out = reduce.mode(LC, [LC_COORD, LC_COORD], [tlon, tlat], [0, 1], LABELS)
    np.testing.assert_array_equal(out, expected)


def test_coord_order_reversed_target():
    """reference tests/test_reduce.py:189-199 -- result follows the target's own order."""
    tlon, tlat = _target(0, 40)
    fwd = reduce.mode(LC, [LC_COORD, LC_COORD], [tlon, tlat], [0, 1], LABELS)
    rev = reduce.mode(LC, [LC_COORD, LC_COORD], [tlon[::-1], tlat], [0, 1], LABELS)
    np.testing.assert_array_equal(rev, fwd[::-1])


def test_min():
    """reference tests/test_reduce.py:202-217."""
    expected = np.array(
        [
            [2.0, 2.0, 0.0, 0.0, 0.0, 0.0],
            [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
            [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
            [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
            [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
            [3.0, 0.0, 0.0, 0.0, 0.0, 1.0],
        ]
    )
    tlon, tlat = _target(0, 40)
    out = reduce.stat(LC.astype(float), [LC_COORD, LC_COORD], [tlon, tlat], [0, 1], "min")
    np.testing.assert_array_equal(out, expected)


VAR_EXPECTED = np.array(
    [
        [0.0, 0.0, 1.0, 0.0, 0.0, 0.0],
        [1.0, 0.75, 0.75, 0.0, 0.0, 0.0],
        [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
        [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
        [2.25, 0.0, 0.0, 0.0, 0.0, 0.25],
        [0.0, 1.6875, 2.25, 0.0, 0.25, 0.0],
    ]
)


def test_var():
    """reference tests/test_reduce.py:220-235 (population variance, ddof=0)."""
    tlon, tlat = _target(0, 40)
    out = reduce.stat(LC.astype(float), [LC_COORD, LC_COORD], [tlon, tlat], [0, 1], "var")
    np.testing.assert_array_equal(out, VAR_EXPECTED)


def test_unsorted_coords():
    """reference tests/test_reduce.py:238-260 -- source rows shuffled along latitude."""
    order = np.argsort(np.array([1, 3, 7, 0, 2, 8, 9, -1, 5, 11, 12]), kind="stable")
    shuffled = LC.astype(float)[:, order]
    tlon, tlat = _target(0, 40)
    out = reduce.stat(shuffled, [LC_COORD, LC_COORD[order]], [tlon, tlat], [0, 1], "var")
    np.testing.assert_array_equal(out, VAR_EXPECTED)


def test_invalid_stat_method():
    """reference methods/flox_reduce.py:70-73."""
    tlon, tlat = _target(0, 40)
    with pytest.raises(ValueError):
        reduce.stat(LC.astype(float), [LC_COORD, LC_COORD], [tlon, tlat], [0, 1], "mode")


def test_mode_requires_integers():
    """reference methods/flox_reduce.py:156-162."""
    tlon, tlat = _target(0, 40)
    with pytest.raises(ValueError):
        reduce.mode(LC.astype(float), [LC_COORD, LC_COORD], [tlon, tlat], [0, 1], LABELS)


# ------------------------------------------------------------------ tests/test_format.py
def _grid(n, e, s, w, d):
    lat, lon = grid.lat_lon_coords(n, e, s, w, d, d)
    return lat, lon


def _fmt(src, tgt, stats=False, dtype=float):
    slat, slon = src
    tlat, tlon = tgt
    data = np.zeros((slat.size, slon.size), dtype=dtype)
    return formatting.format_for_regrid(data, slat, slon, tlat, tlon, stats=stats)


def test_format_covered():
    """reference tests/test_format.py:8-32 -- nothing changes."""
    src = _grid(90, 360, -90, 0, 2)
    data, lat, lon = _fmt(src, _grid(80, 350, -80, 10, 1))
    np.testing.assert_array_equal(lat, src[0])
    np.testing.assert_array_equal(lon, src[1])
    assert data.shape == (src[0].size, src[1].size)


def test_format_no_edges():
    """reference tests/test_format.py:35-63 -- pole rows and wrap columns are added."""
    data, lat, lon = _fmt(_grid(89, 359, -89, 1, 2), _grid(90, 360, -90, 0, 1))
    assert lat[0] == -90 and lat[-1] == 90
    assert lon[0] == -1 and lon[-1] == 361
    assert (np.diff(lon) == 2).all()
    assert data.shape == (lat.size, lon.size)


def test_format_360_to_180():
    """reference tests/test_format.py:66-92."""
    _, _, lon = _fmt(_grid(90, 360, -90, 0, 2), _grid(90, 180, -90, -180, 1))
    assert lon[0] == -182 and lon[-1] == 182 and (np.diff(lon) == 2).all()


def test_format_180_to_360():
    """reference tests/test_format.py:95-121."""
    _, _, lon = _fmt(_grid(90, 180, -90, -180, 2), _grid(90, 360, -90, 0, 1))
    assert lon[0] == -2 and lon[-1] == 362 and (np.diff(lon) == 2).all()


def test_format_0_to_360():
    """reference tests/test_format.py:124-150."""
    _, _, lon = _fmt(_grid(90, 0, -90, -360, 2), _grid(90, 360, -90, 0, 1))
    assert lon[0] == -2 and lon[-1] == 362 and (np.diff(lon) == 2).all()


def test_format_global_to_local_shift():
    """reference tests/test_format.py:153-179."""
    _, _, lon = _fmt(_grid(90, 180, -90, -180, 2), _grid(90, 300, -90, 270, 1))
    assert lon.min() <= 270 and lon.max() >= 300 and (np.diff(lon) == 2).all()


def test_format_stats():
    """reference tests/test_format.py:182-218 -- no pole rows, wrap columns kept, ints stay ints."""
    src = _grid(89.5, 359.5, -89.5, 0.5, 1)
    data, lat, lon = _fmt(src, _grid(90, 360, -90, 0, 2), stats=True, dtype=np.int64)
    np.testing.assert_array_equal(lat, src[0])
    assert lon[0] == -1.5 and lon[-1] == 361.5 and (np.diff(lon) == 1).all()
    assert data.dtype == np.int64


# ------------------------------------------------------------------- tests/test_utils.py
def test_format_lat_pole_values():
    """reference tests/test_utils.py:7-29 -- pole rows hold the longitude-mean of the edge rows."""
    lat = np.arange(-89.5, 89.5 + 1, 1)
    lon = np.arange(-179.5, 179.5 + 1, 1)
    x = np.broadcast_to(lat, (lon.size, lat.size))  # dims (lon, lat)
    out, new_lat = formatting.format_lat(x, lat, lat_axis=1, lon_axis=0)
    assert new_lat[0] == -90 and new_lat[-1] == 90
    assert (out[:, 0] == -89.5).all() and (out[:, -1] == 89.5).all()
    assert out.shape == (lon.size, lat.size + 2)


# ------------------------------------------------------- config shapes quoted in SURVEY 8(a3)
def test_config_padding_shapes():
    lat025, lon025 = grid.lat_lon_coords(90, 359.75, -90, 0, 0.25, 0.25)
    lat1, lon1 = grid.lat_lon_coords(90, 359, -90, 0, 1, 1)
    assert (lat025.size, lon025.size, lat1.size, lon1.size) == (721, 1440, 181, 360)
    data = np.arange(lon025.size, dtype=float)[None, :].repeat(2, 0)
    out, _, lon = formatting.format_for_regrid(data, None, lon025, None, lon1)
    # 359.75 is shifted to -0.25 (wrap point 359.5), then one wrapped column is added on the left
    assert lon.size == 1441 and lon[0] == -0.5 and lon[-1] == 359.5
    np.testing.assert_array_equal(out[0], np.concatenate([[1438, 1439], np.arange(1439)]))


def test_interp_semantics_pinned_to_scipy():
    """ties to the lower neighbour, ints -> float64, NaN outside, NaN taps poison."""
    x = np.arange(5.0)
    y = np.array([10, 20, 30, 40, 50], dtype=np.int32)
    out = interp.regrid(y, [x], [np.array([0.5, 2.5, -1.0, 4.0, 4.5])], [0], "nearest")
    assert out.dtype == np.float64
    np.testing.assert_array_equal(out, [10, 30, np.nan, 50, np.nan])
    yf = np.array([0, 1, 2, np.nan, 4], dtype=np.float32)
    out = interp.regrid(yf, [x], [np.array([1.5, 2.0, 2.5, 4.0])], [0], "linear")
    assert out.dtype == np.float64
    np.testing.assert_array_equal(out, [1.5, 2.0, np.nan, np.nan])
    out = interp.regrid(yf, [x], [np.array([0.5])], [0], "cubic")
    assert np.isnan(out).all()
