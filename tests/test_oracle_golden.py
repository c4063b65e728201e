"""The oracle's helpers against the REFERENCE's own helpers (fixtures made by
oracle/gen_golden.py, which executes /root/reference code in the build container)."""

import numpy as np
import pytest

from oracle import conservative, grid, reduce

AXIS_CASES = [
    "c1_lat", "c1_lon", "era5_to_2p2_lat", "cellcentred_lat", "upsample",
    "irregular", "single_target", "tgt_outside",
]


@pytest.mark.parametrize("name", AXIS_CASES)
def test_weights_bit_exact(golden_helpers, name):
    g = golden_helpers
    src, tgt = g[f"axis/{name}/src"], g[f"axis/{name}/tgt"]
    np.testing.assert_array_equal(conservative.cell_edges(src), g[f"axis/{name}/src_edges"])
    np.testing.assert_array_equal(conservative.get_weights(src, tgt), g[f"axis/{name}/weights"])


@pytest.mark.parametrize("name", AXIS_CASES)
def test_spherical_correction_bit_exact(golden_helpers, name):
    g = golden_helpers
    key = f"axis/{name}/weights_spherical"
    if key not in g:
        pytest.skip("not a latitude-like axis")
    src, tgt = g[f"axis/{name}/src"], g[f"axis/{name}/tgt"]
    res = np.median(np.diff(src, 1))
    np.testing.assert_array_equal(conservative.lat_weight(src, res), g[f"axis/{name}/lat_weight"])
    w = conservative.spherical_correction(conservative.get_weights(src, tgt), src)
    np.testing.assert_array_equal(w, g[key])


@pytest.mark.parametrize("name", AXIS_CASES)
def test_bin_breaks_bit_exact(golden_helpers, name):
    g = golden_helpers
    key = f"axis/{name}/bin_breaks"
    if key not in g:
        pytest.skip("single-point target")
    np.testing.assert_array_equal(reduce.bin_breaks(g[f"axis/{name}/tgt"]), g[key])


def test_valid_threshold_bit_exact(golden_helpers):
    g = golden_helpers
    got = np.array([conservative.valid_threshold(t) for t in g["threshold/in"]])
    np.testing.assert_array_equal(got, g["threshold/out"])


def test_grid_coords_bit_exact(golden_helpers):
    g = golden_helpers
    for i, p in enumerate(g["grid/params"]):
        lat, lon = grid.lat_lon_coords(*p)
        np.testing.assert_array_equal(lat, g[f"grid/{i}/lat"])
        np.testing.assert_array_equal(lon, g[f"grid/{i}/lon"])


def test_pole_rows_have_exactly_zero_area_weight(golden_helpers):
    """lat_weight is exactly 0.0 at +-90 on the 0.25 degree grid, so the
    pole rows drop out of the band (matters for 0*NaN with skipna=False)."""
    lw = golden_helpers["axis/c1_lat/lat_weight"]
    assert lw[0] == 0.0 and lw[-1] == 0.0
    assert (lw[1:-1] > 0).all()
