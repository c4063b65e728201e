"""Parity of linear / nearest / cubic (CUDA, through the C ABI) against the oracle, which calls the
same scipy ``interp1d`` the reference reaches through ``xr.interp`` (methods/interp.py:43-46).

Tolerances (BASELINE.json): nearest bit-exact; linear / cubic 1e-12 relative in fp64, 1e-5 in fp32;
NaN placement identical."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import interp as oi  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xarray_regrid_b200 import engine

    return engine


def _check(got, want, rtol, atol=0.0):
    got = got.cpu().numpy() if hasattr(got, "cpu") else got
    assert got.shape == want.shape
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want), err_msg="NaN placement differs")
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol)


def _plan(engine, lat, lon, tlat, tlon, method):
    return engine.InterpPlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"), method)


GRID_1 = (np.arange(-90.0, 91.0, 1.0), np.arange(0.0, 360.0, 1.0))
GRID_01 = (np.round(np.arange(-90.0, 90.05, 0.1), 10), np.round(np.arange(0.0, 359.95, 0.1), 10))
GRID_025 = (np.arange(-90.0, 90.25, 0.25), np.arange(0.0, 360.0, 0.25))


@pytest.mark.parametrize("ring", [True, False])
def test_c3_linear_upsampling(engine, ring):
    """BASELINE config 3 shape: 1 deg -> 0.1 deg (1801 x 3600), fp32 data."""
    rng = np.random.default_rng(7)
    (lat, lon), (tlat, tlon) = GRID_1, GRID_01
    assert (tlat.size, tlon.size) == (1801, 3600)
    x = (250 + 50 * rng.random((2, lat.size, lon.size))).astype(np.float32)
    plan = _plan(engine, lat, lon, tlat, tlon, "linear")
    plan.op.use_ring = ring
    want = oi.regrid_latlon(x, lat, lon, tlat, tlon, "linear")
    assert want.dtype == np.float64
    y = plan(torch.from_numpy(x).cuda())  # reference dtype rule: float64 out
    assert y.dtype == torch.float64
    _check(y, want, 1e-12)
    y32 = plan(torch.from_numpy(x).cuda(), out_dtype="same")  # engine option: stay in fp32
    assert y32.dtype == torch.float32
    _check(y32, want.astype(np.float32), 1e-5)


def test_c3_cubic_upsampling(engine):
    rng = np.random.default_rng(8)
    (lat, lon), (tlat, tlon) = GRID_1, GRID_01
    x = (250 + 50 * rng.random((2, lat.size, lon.size))).astype(np.float32)
    plan = _plan(engine, lat, lon, tlat, tlon, "cubic")
    y = plan(torch.from_numpy(x).cuda())
    want = oi.regrid_latlon(x, lat, lon, tlat, tlon, "cubic")
    assert y.dtype == torch.float64 and want.dtype == np.float64
    _check(y, want, 1e-11)  # spec: 1e-5 for fp32 input; the fp64 pipeline is ~1e-13


@pytest.mark.parametrize("case", ["global", "irregular", "rows_only", "descending", "irregular_long",
                                  "regular_long_regional"])
def test_cubic_lu_prefilter_matches_band_prefilter_and_oracle(engine, case):
    """The O(n) banded LU substitution (rg_spline_solve_f64) and the truncated-inverse band pass are two
    routes to scipy's spline coefficients: both must match the oracle (scipy itself) to 1e-11.  Lines of 200+
    elements along the contiguous axis are solved by four threads each, every segment warmed up from a zero state
    over a contractive window (global, regular_long_regional); an irregular axis of that length is not contractive
    and must take the one-thread-per-line walk (irregular_long)."""
    rng = np.random.default_rng(21)
    if case == "global":
        (lat, lon), (tlat, tlon) = GRID_1, (np.arange(-89.9, 90, 0.7), np.arange(0.05, 359.9, 0.9))
    elif case == "irregular":
        lat, lon = np.sort(rng.uniform(-60, 60, 75)), np.sort(rng.uniform(10, 200, 130))
        tlat, tlon = np.linspace(lat[0], lat[-1], 140), np.linspace(lon[0], lon[-1], 260)
    elif case == "irregular_long":
        lat, lon = np.sort(rng.uniform(-60, 60, 45)), np.sort(rng.uniform(10, 200, 300))
        lon = lon[np.concatenate([[True], np.diff(lon) > 0.05])]  # (near-coincident knots: ill-conditioned for scipy too)
        tlat, tlon = np.linspace(lat[0], lat[-1], 60), np.linspace(lon[0], lon[-1], 500)
    elif case == "regular_long_regional":
        lat, lon = np.arange(-20.0, 21.0, 1.0), np.arange(30.0, 93.0, 0.25)  # 252 columns, 41 x 3 = 123 lines
        tlat, tlon = np.arange(-19.5, 20.0, 0.4), np.arange(30.1, 92.0, 0.1)
    elif case == "rows_only":
        lat, lon = np.arange(0.0, 40.0, 1.0), np.arange(100.0, 133.0, 1.0)
        tlat, tlon = np.arange(0.5, 38.0, 0.25), None
    else:
        lat, lon = np.arange(50.0, 10.0, -1.0), np.arange(100.0, 133.0, 1.0)
        tlat, tlon = np.arange(11.0, 49.0, 0.5), np.arange(132.0, 100.0, -0.3)
    x = rng.standard_normal((3, lat.size, lon.size))
    if tlon is None:
        plan = engine.InterpPlan(engine.AxisSpec(lat, tlat, "lat"), None, "cubic")
        want = oi.regrid(x, [lat], [tlat], [-2], "cubic")
    else:
        plan = _plan(engine, lat, lon, tlat, tlon, "cubic")
        want = oi.regrid_latlon(x, lat, lon, tlat, tlon, "cubic")
    assert plan._lu is not None, "grid should qualify for the LU prefilter"
    xt = torch.from_numpy(x).cuda()
    y_lu = plan(xt)
    plan.use_lu = False
    y_band = plan(xt)
    _check(y_lu, want, 1e-11, 1e-13)
    _check(y_band, want, 1e-11, 1e-13)
    assert torch.equal(xt.cpu(), torch.from_numpy(x)), "the in-place solve must not touch the input"


def test_cubic_any_nan_makes_everything_nan(engine):
    """scipy interp1d(kind='cubic'): one NaN anywhere -> the whole result is NaN."""
    rng = np.random.default_rng(9)
    lat, lon = np.arange(0.0, 20.0, 1.0), np.arange(100.0, 130.0, 1.0)
    tlat, tlon = np.arange(1.0, 18.0, 0.5), np.arange(101.0, 128.0, 0.5)
    x = rng.random((3, lat.size, lon.size))
    x[1, 4, 7] = np.nan
    plan = _plan(engine, lat, lon, tlat, tlon, "cubic")
    y = plan(torch.from_numpy(x).cuda())
    want = oi.regrid_latlon(x, lat, lon, tlat, tlon, "cubic")
    assert np.isnan(want).all()
    assert torch.isnan(y).all()
    x[1, 4, 7] = 0.5
    _check(plan(torch.from_numpy(x).cuda()), oi.regrid_latlon(x, lat, lon, tlat, tlon, "cubic"), 1e-11, 1e-13)


def test_c4_nearest_int32_bit_exact(engine):
    """BASELINE config 4 shape: 0.1 deg -> 0.25 deg, int32 categorical, bit-exact incl. ties."""
    rng = np.random.default_rng(11)
    (lat, lon), (tlat, tlon) = GRID_01, GRID_025
    x = rng.integers(0, 32, size=(2, lat.size, lon.size), dtype=np.int32)
    plan = _plan(engine, lat, lon, tlat, tlon, "nearest")
    want = oi.regrid_latlon(x, lat, lon, tlat, tlon, "nearest")
    assert want.dtype == np.float64 and not np.isnan(want).any()
    y = plan(torch.from_numpy(x).cuda())
    assert y.dtype == torch.float64
    np.testing.assert_array_equal(y.cpu().numpy(), want)
    y_same = plan(torch.from_numpy(x).cuda(), out_dtype="same")
    assert y_same.dtype == torch.int32
    np.testing.assert_array_equal(y_same.cpu().numpy(), want.astype(np.int32))


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.int64, np.float32, np.float64])
def test_nearest_dtypes_ties_and_out_of_range(engine, dtype):
    """Exact midpoints go to the LOWER neighbour; targets outside the source are NaN; ints -> float64."""
    rng = np.random.default_rng(12)
    lat, lon = np.arange(0.0, 10.0, 1.0), np.arange(0.0, 12.0, 2.0)
    tlat, tlon = np.array([-0.5, 0.0, 0.5, 1.5, 2.49, 2.5, 9.0, 9.5]), np.array([-1.0, 1.0, 3.0, 9.9, 10.0, 11.0])
    x = rng.integers(1, 100, size=(3, lat.size, lon.size)).astype(dtype)
    if np.issubdtype(dtype, np.floating):
        x[0, 2, 1] = np.nan
    plan = engine.InterpPlan(engine.AxisSpec(lat, tlat), engine.AxisSpec(lon, tlon), "nearest")
    y = plan(torch.from_numpy(x).cuda())
    want = oi.regrid(x, [lat, lon], [tlat, tlon], [-2, -1], "nearest")
    assert str(y.dtype).endswith(str(want.dtype))
    np.testing.assert_array_equal(y.cpu().numpy(), want)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n_tlat", [36, 33, 68, 32, 5])
def test_upsampling_band_prefetch_transitions(engine, dtype, n_tlat):
    """The upsampling kernel prefetches the NEXT band's row taps and source rows while it produces the current one,
    and falls back to a synchronous prologue after a band of <= 4 rows.  Many slabs on a small grid make every
    persistent CTA walk through long / short band sequences (32 + 4, 32 + 1, 32 + 32 + 4, a single band, one short
    band only); every slab is checked, so a row staged into the wrong window buffer cannot hide."""
    rng = np.random.default_rng(31)
    lat, lon = np.linspace(-30.0, 30.0, 12), np.linspace(10.0, 100.0, 16)
    tlat, tlon = np.linspace(-30.0, 30.0, n_tlat), np.linspace(10.0, 100.0, 64)
    batch = 700  # x bands > resident CTAs: several tasks per CTA
    x = rng.normal(size=(batch, lat.size, lon.size)).astype(dtype)
    x[rng.random(x.shape) < 0.01] = np.nan
    x += np.arange(batch, dtype=dtype)[:, None, None]  # slabs differ by construction
    plan = engine.InterpPlan(engine.AxisSpec(lat, tlat), engine.AxisSpec(lon, tlon), "linear")
    y = plan(torch.from_numpy(x).cuda(), out_dtype="same")
    assert y.dtype == (torch.float32 if dtype == np.float32 else torch.float64)
    want = oi.regrid(x, [lat, lon], [tlat, tlon], [-2, -1], "linear").astype(dtype)
    _check(y, want, 1e-5 if dtype == np.float32 else 1e-12, atol=1e-4 if dtype == np.float32 else 1e-10)


def test_linear_nan_taps_poison_even_at_zero_weight(engine):
    lat, lon = np.arange(0.0, 6.0, 1.0), np.arange(0.0, 8.0, 1.0)
    tlat, tlon = np.array([0.0, 1.0, 1.5, 2.0, 3.0, 5.0]), np.array([0.0, 2.0, 2.5, 3.0, 4.0, 7.0])
    x = np.arange(48, dtype=np.float64).reshape(1, 6, 8) + 1
    x[0, 2, 3] = np.nan
    plan = engine.InterpPlan(engine.AxisSpec(lat, tlat), engine.AxisSpec(lon, tlon), "linear")
    want = oi.regrid(x, [lat, lon], [tlat, tlon], [-2, -1], "linear")
    assert np.isnan(want).sum() >= 6  # the NaN node poisons (x[m-1], x[m]] on both axes
    _check(plan(torch.from_numpy(x).cuda()), want, 1e-14)


@pytest.mark.parametrize("method", ["linear", "nearest", "cubic"])
def test_cell_centred_grid_pole_rows_and_wrap(engine, method):
    """Source -89..89 / 1..359 (2 deg) -> pole-centred 1 deg target: format_for_regrid adds pole rows
    (lon-mean) and wrap columns before xr.interp (regrid.py:44/63/82; tests/test_format.py:35-63)."""
    rng = np.random.default_rng(5)
    lat, lon = np.arange(-89.0, 90, 2), np.arange(1.0, 360, 2)
    tlat, tlon = np.arange(-90.0, 91, 1), np.arange(0.0, 361, 1)
    x = 0.5 + rng.random((3, lat.size, lon.size))
    if method != "cubic":
        x[1, 10:12, 20:22] = np.nan
    plan = _plan(engine, lat, lon, tlat, tlon, method)
    y = plan(torch.from_numpy(x).cuda())
    want = oi.regrid_latlon(x, lat, lon, tlat, tlon, method)
    _check(y, want, 1e-11 if method == "cubic" else 1e-13)


@pytest.mark.parametrize("method", ["linear", "nearest", "cubic"])
def test_regional_descending_source_reversed_target(engine, method):
    """ERA5-style descending latitude, a target that pokes outside the source (-> NaN) and is reversed."""
    rng = np.random.default_rng(6)
    lat, lon = np.arange(60.0, 29.9, -0.25), np.arange(-10.0, 30.25, 0.25)
    tlat, tlon = np.arange(28.0, 63.0, 0.7)[::-1].copy(), np.arange(-12.0, 33.0, 0.9)
    x = 280 + rng.random((2, lat.size, lon.size))
    plan = _plan(engine, lat, lon, tlat, tlon, method)
    y = plan(torch.from_numpy(x).cuda())
    want = oi.regrid_latlon(x, lat, lon, tlat, tlon, method)
    if method == "cubic":
        # rows outside the source become NaN in the first pass and scipy's cubic then blanks everything
        assert np.isnan(want).all()
    else:
        assert np.isnan(want).any() and not np.isnan(want).all()
    _check(y, want, 1e-11 if method == "cubic" else 1e-13)
    if method == "cubic":  # same source, targets inside on the row axis: finite results
        inside = tlat[(tlat >= lat.min()) & (tlat <= lat.max())]
        plan = _plan(engine, lat, lon, inside, tlon, method)
        want = oi.regrid_latlon(x, lat, lon, inside, tlon, method)
        assert np.isnan(want).any() and not np.isnan(want).all()
        _check(plan(torch.from_numpy(x).cuda()), want, 1e-11)


def test_single_axis_and_leading_dims(engine):
    rng = np.random.default_rng(4)
    src, tgt = np.arange(0.0, 30.0, 1.0), np.arange(0.5, 29.0, 0.25)
    x = rng.random((2, 3, src.size, 9))
    for method in ("linear", "cubic", "nearest"):
        plan = engine.InterpPlan(engine.AxisSpec(src, tgt), None, method)
        y = plan(torch.from_numpy(x).cuda())
        want = oi.regrid(x, [src], [tgt], [-2], method)
        assert y.shape == (2, 3, tgt.size, 9)
        _check(y, want, 1e-11, 1e-13)


def test_empty_batch_and_ragged_shapes(engine):
    """Empty stacks, a 1-row / 1-column source axis for nearest, and widths that defeat 16-byte alignment."""
    lat, lon = np.arange(0.0, 7.0, 1.0), np.arange(0.0, 9.0, 1.0)
    tlat, tlon = np.arange(0.5, 6.0, 0.5), np.arange(0.25, 8.0, 0.75)
    for method in ("linear", "nearest", "cubic"):
        plan = engine.InterpPlan(engine.AxisSpec(lat, tlat), engine.AxisSpec(lon, tlon), method)
        y = plan(torch.empty((0, 7, 9), dtype=torch.float64, device="cuda"))
        assert y.shape == (0, tlat.size, tlon.size)
        rng = np.random.default_rng(21)
        x = rng.random((5, 7, 9))
        _check(plan(torch.from_numpy(x).cuda()), oi.regrid(x, [lat, lon], [tlat, tlon], [-2, -1], method),
               1e-11, 1e-13)
    with pytest.raises(ValueError):
        engine.InterpPlan(engine.AxisSpec(lat, tlat), None, "quadratic")
    with pytest.raises(ValueError):
        engine.InterpPlan(engine.AxisSpec(np.arange(3.0), tlat), None, "cubic")  # needs >= 4 points (scipy)
