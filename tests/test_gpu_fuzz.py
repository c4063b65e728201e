"""Randomised parity of every method on irregular grids (CUDA, through the C ABI) against the oracle.

The hand-written cases use the reference's regular lat / lon grids; the kernels, however, take arbitrary
band / bin tables.  Here the source coordinates are irregular (clustered), the targets are shuffled,
partly outside the source range, finer or coarser than the source, with odd sizes, so that every table
contract between the host and the kernels (window sizes, strip widths, alignment fallbacks) is exercised.
Tolerances as everywhere: bit-exact for nearest / modes / min / max, 1e-12 relative (fp64) otherwise,
identical NaN placement.
"""

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import conservative as oc  # noqa: E402
from oracle import interp as oi  # noqa: E402
from oracle import reduce as orr  # noqa: E402

pytestmark = pytest.mark.gpu

SEEDS = list(range(10))


@pytest.fixture(scope="module")
def engine():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xarray_regrid_b200 import engine

    return engine


def _irregular(rng, n, lo, hi):
    """Sorted coordinates with clusters and gaps (distinct values)."""
    x = np.sort(np.r_[rng.uniform(lo, hi, n - n // 3), rng.normal(rng.uniform(lo, hi), (hi - lo) / 40, n // 3)])
    x = np.unique(np.round(x, 6))
    return x


def _targets(rng, src, n, shuffle, overshoot):
    span = src[-1] - src[0]
    lo = src[0] - (overshoot * span if rng.random() < 0.5 else -0.02 * span)
    hi = src[-1] + (overshoot * span if rng.random() < 0.5 else -0.02 * span)
    t = np.linspace(lo, hi, n)
    return rng.permutation(t) if shuffle else t


def _check(got, want, rtol, atol=0.0):
    got = got.cpu().numpy()
    assert got.shape == want.shape
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want), err_msg="NaN placement differs")
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol)


@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("method", ["linear", "nearest", "cubic"])
def test_interp_irregular(engine, seed, method):
    rng = np.random.default_rng(1000 + seed)
    lat, lon = _irregular(rng, int(rng.integers(8, 70)), -50, 60), _irregular(rng, int(rng.integers(8, 90)), 0, 300)
    tlat = _targets(rng, lat, int(rng.integers(3, 160)), shuffle=seed % 2 == 1, overshoot=0.05 if method != "cubic" else 0.0)
    tlon = _targets(rng, lon, int(rng.integers(3, 200)), shuffle=seed % 3 == 1, overshoot=0.05)
    if method == "cubic":  # a row target outside the source blanks everything (second pass sees NaN): keep rows inside
        tlat = np.clip(tlat, lat[0], lat[-1])
    x = rng.standard_normal((int(rng.integers(1, 4)), lat.size, lon.size))
    if method == "linear" and seed % 2 == 0:
        x[rng.random(x.shape) < 0.05] = np.nan
    plan = engine.InterpPlan(engine.AxisSpec(lat, tlat), engine.AxisSpec(lon, tlon), method)
    y = plan(torch.from_numpy(x).cuda())
    want = oi.regrid(x, [lat, lon], [tlat, tlon], [-2, -1], method)
    if method == "nearest":
        _check(y, want, 0.0)
    else:
        scale = np.nanmax(np.abs(want)) if np.isfinite(want).any() else 1.0
        _check(y, want, 1e-10 if method == "cubic" else 1e-12, atol=(1e-10 if method == "cubic" else 1e-13) * scale)


@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_conservative_irregular(engine, seed, dtype):
    rng = np.random.default_rng(2000 + seed)
    lat, lon = _irregular(rng, int(rng.integers(8, 70)), -50, 60), _irregular(rng, int(rng.integers(8, 120)), 0, 300)
    tlat = _targets(rng, lat, int(rng.integers(3, 90)), shuffle=seed % 2 == 1, overshoot=0.1)
    tlon = _targets(rng, lon, int(rng.integers(3, 130)), shuffle=seed % 3 == 1, overshoot=0.1)
    x = (0.5 + rng.random((int(rng.integers(1, 5)), lat.size, lon.size))).astype(dtype)
    skipna = seed % 2 == 0
    x[rng.random(x.shape) < 0.1] = np.nan
    thr = float(rng.choice([0.0, 0.3, 0.5, 1.0]))
    plan = engine.ConservativePlan(engine.AxisSpec(lat, tlat), engine.AxisSpec(lon, tlon))
    want = oc.regrid(x, [lat, lon], [tlat, tlon], [-2, -1], skipna=skipna, nan_threshold=thr)
    rtol = 1e-12 if dtype == np.float64 else 1e-5
    for ring in (True, False):
        plan.use_ring = ring
        _check(plan(torch.from_numpy(x).cuda(), skipna=skipna, nan_threshold=thr), want, rtol, atol=rtol * 1e-2)


@pytest.mark.parametrize("seed", SEEDS)
def test_reduce_irregular_source(engine, seed):
    rng = np.random.default_rng(3000 + seed)
    lat, lon = _irregular(rng, int(rng.integers(20, 120)), -50, 60), _irregular(rng, int(rng.integers(20, 160)), 0, 300)
    # the reference infers the bin width from the target spacing: uniform targets (any order)
    tlat = np.linspace(lat[0] - 3, lat[-1] + 3, int(rng.integers(3, 30)))
    tlon = np.linspace(lon[0] + 5, lon[-1] - 5, int(rng.integers(3, 40)))
    if seed % 2:
        tlat = tlat[::-1].copy()
    plan = engine.ReducePlan(engine.AxisSpec(lat, tlat), engine.AxisSpec(lon, tlon))
    labels = rng.integers(0, 7, (2, lat.size, lon.size)).astype(np.uint8 if seed % 2 else np.int32)
    values = np.arange(7)
    for anti in (False, True):
        y = plan.mode(torch.from_numpy(labels).cuda(), values, fill_value=255 if seed % 2 else -9, anti_mode=anti)
        want = orr.mode(labels, [lat, lon], [tlat, tlon], [-2, -1], values, fill_value=255 if seed % 2 else -9,
                        anti_mode=anti)
        np.testing.assert_array_equal(y.cpu().numpy(), want)
    x = rng.standard_normal((2, lat.size, lon.size))
    x[rng.random(x.shape) < 0.05] = np.nan
    for method in ("sum", "mean", "var", "std", "min", "max", "median"):
        for skipna in (False, True):
            y = plan.stat(torch.from_numpy(x).cuda(), method, skipna=skipna)
            want = orr.stat(x, [lat, lon], [tlat, tlon], [-2, -1], method, skipna=skipna)
            if method in ("min", "max", "median"):
                np.testing.assert_array_equal(y.cpu().numpy(), want, err_msg=method)
            else:
                _check(y, want, 1e-11, atol=1e-12)


@pytest.mark.parametrize("seed", list(range(16)))
def test_conservative_regular_ratio_2_to_3(engine, seed):
    """Regular source and target grids at a random coarsening ratio in [2, 3] with random offsets, global (periodic
    longitude, with or without wrap padding) or regional, even / odd sizes, fp64 / fp32: the block layout of the
    register-resident kernel (one or two targets per lane, halo widths 1 - 2, panels on wide rows) -- or whatever the
    host picks when a grid does not qualify -- against the oracle, all three NaN modes."""
    rng = np.random.default_rng(1000 + seed)
    dtype = np.float64 if seed % 2 == 0 else np.float32
    res = float(rng.choice([0.25, 0.5, 1.0, 0.2]))
    ratio = float(rng.choice([2.0, 2.5, 3.0, 2.25, 2.8]))
    if seed % 3 == 0:  # global longitude
        lon = np.round(np.arange(0.0, 360.0, res) + rng.choice([0.0, res / 2]), 10)
        tlon = np.round(np.arange(0.0, 360.0, res * ratio) + rng.choice([0.0, res * ratio / 2, 0.1]), 10)
    else:
        n = int(rng.integers(60, 700))
        lon0 = float(rng.uniform(-170, 20))
        lon = np.round(lon0 + res * np.arange(n), 10)
        tlon = np.round(np.arange(lon[0] + rng.choice([0.0, res / 2, 3 * res]), lon[-1] - res, res * ratio), 10)
    nlat = int(rng.integers(30, 90))
    lat = np.round(-20.0 + res * np.arange(nlat), 10)
    tlat = np.round(np.arange(lat[0] + res, lat[-1] - res, res * ratio), 10)
    if seed % 4 == 1:
        lat = lat[::-1].copy()
    B = int(rng.integers(1, 5))
    x = (0.5 + rng.random((B, lat.size, lon.size))).astype(dtype)
    x[rng.random(x.shape) < 0.08] = np.nan
    plan = engine.ConservativePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"))
    xt = torch.from_numpy(x).cuda()
    rtol = 1e-12 if dtype == np.float64 else 2e-5
    for skipna, thr in [(True, 0.4), (True, 1.0), (False, 1.0)]:
        want, valid = oc.regrid_latlon(x, lat, lon, tlat, tlon, skipna=skipna, nan_threshold=thr, return_valid=True)
        ties = np.abs(valid - oc.valid_threshold(thr)) <= (1e-13 if dtype == np.float64 else 1e-6)
        y = plan(xt, skipna=skipna, nan_threshold=thr).cpu().numpy()
        np.testing.assert_array_equal(np.isnan(y)[~ties], np.isnan(want)[~ties])
        both = ~np.isnan(y) & ~np.isnan(want)
        np.testing.assert_allclose(y[both], want[both], rtol=rtol, atol=rtol)
