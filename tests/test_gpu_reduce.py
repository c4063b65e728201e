"""Parity of most_common / least_common / stat (CUDA, through the C ABI) against the oracle and the
reference's known answers (tests/test_reduce.py).  Bit-exact for the label modes and min / max;
1e-12 (fp64) / 1e-5 (fp32) relative for the floating-point statistics."""

import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import reduce as orr  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xarray_regrid_b200 import engine

    return engine


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
C11 = np.linspace(0, 40, num=11)
T6 = np.arange(0.0, 41.0, 8.0)
LABELS = np.array([0, 1, 2, 3])


def _plan(engine, r_src, c_src, r_tgt, c_tgt, kinds=("other", "other")):
    return engine.ReducePlan(engine.AxisSpec(r_src, r_tgt, kinds[0]), engine.AxisSpec(c_src, c_tgt, kinds[1]))


def test_reference_known_answer_most_common(engine):
    """reference tests/test_reduce.py:87-110 (ties -> label 0, uint8 preserved) through the CUDA path."""
    expected = np.array(
        [[2, 2, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0],
         [0, 0, 0, 0, 0, 0], [3, 3, 0, 0, 0, 1]], dtype="uint8")
    plan = _plan(engine, C11, C11, T6, T6)
    y = plan.mode(torch.from_numpy(LC.astype("uint8")).cuda(), LABELS)
    assert y.dtype == torch.uint8
    np.testing.assert_array_equal(y.cpu().numpy(), expected)
    yl = plan.mode(torch.from_numpy(LC.astype("uint8")).cuda(), LABELS, anti_mode=True)
    np.testing.assert_array_equal(
        yl.cpu().numpy(), orr.mode(LC.astype("uint8"), [C11, C11], [T6, T6], [0, 1], LABELS, anti_mode=True))


def test_reference_known_answer_oversized_target(engine):
    """reference tests/test_reduce.py:121-153: NaN ring, float result, warning."""
    t8 = np.arange(-8.0, 49.0, 8.0)
    plan = _plan(engine, C11, C11, t8, t8)
    with pytest.warns(UserWarning, match="fill_value"):
        y = plan.mode(torch.from_numpy(LC).cuda(), LABELS)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = orr.mode(LC, [C11, C11], [t8, t8], [0, 1], LABELS)
    assert y.dtype == torch.float64 and np.isnan(want[0]).all()
    np.testing.assert_array_equal(y.cpu().numpy(), want)
    y = plan.mode(torch.from_numpy(LC).cuda(), LABELS, fill_value=-9)
    assert y.dtype == torch.int64 and (y[0] == -9).all()


@pytest.mark.parametrize("method,expected", [
    ("min", [[2.0, 2.0, 0, 0, 0, 0], [0] * 6, [0] * 6, [0] * 6, [0] * 6, [3.0, 0, 0, 0, 0, 1.0]]),
    ("var", [[0, 0, 1.0, 0, 0, 0], [1.0, 0.75, 0.75, 0, 0, 0], [0] * 6, [0] * 6,
             [2.25, 0, 0, 0, 0, 0.25], [0, 1.6875, 2.25, 0, 0.25, 0]]),
])
def test_reference_known_answer_stats(engine, method, expected):
    """reference tests/test_reduce.py:202-235 (min; population variance) and :238-260 (unsorted source)."""
    plan = _plan(engine, C11, C11, T6, T6)
    y = plan.stat(torch.from_numpy(LC.astype(float)).cuda(), method)
    np.testing.assert_array_equal(y.cpu().numpy(), np.array(expected, dtype=float))
    order = np.argsort(np.array([1, 3, 7, 0, 2, 8, 9, -1, 5, 11, 12]), kind="stable")
    plan = _plan(engine, C11, C11[order], T6, T6)
    y = plan.stat(torch.from_numpy(LC.astype(float)[:, order].copy()).cuda(), method)
    np.testing.assert_array_equal(y.cpu().numpy(), np.array(expected, dtype=float))


def _landcover(rng, shape, n_classes, block):
    coarse = rng.integers(0, n_classes, size=(shape[0] // block + 1, shape[1] // block + 1))
    x = np.kron(coarse, np.ones((block, block), dtype=np.int64))[: shape[0], : shape[1]]
    noise = rng.random(shape) < 0.05
    x[noise] = rng.integers(0, n_classes, size=int(noise.sum()))
    return x


@pytest.mark.parametrize("aligned", [True, False])
def test_c5_most_common_uint8_32_classes(engine, aligned):
    """BASELINE config 5 shape at 1/10 scale in longitude: 0.01 deg uint8 land cover, 32 classes -> 0.25 deg.
    aligned=True: cell-aligned target (25 x 25 cells per bin); False: pole-centred target whose bin edges
    coincide with source centres (membership decided by fp64 rounding -> host arithmetic must match)."""
    rng = np.random.default_rng(13)
    lat = np.round(np.arange(-9.995, 10, 0.01), 3)
    lon = np.round(np.arange(0.005, 360, 0.01), 3)[:3600]
    if aligned:
        tlat, tlon = np.arange(-9.875, 10, 0.25), np.arange(0.125, 36, 0.25)
    else:
        tlat, tlon = np.arange(-10.0, 10.01, 0.25), np.arange(0.0, 36.01, 0.25)
    x = _landcover(rng, (lat.size, lon.size), 32, 64).astype(np.uint8)
    plan = _plan(engine, lat, lon, tlat, tlon, ("lat", "lon"))
    values = np.arange(32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y = plan.mode(torch.from_numpy(x).cuda(), values)
        want = orr.regrid_latlon(x, lat, lon, tlat, tlon, "most_common", values=values)
    np.testing.assert_array_equal(y.cpu().numpy(), want)
    assert y.dtype == (torch.uint8 if aligned else torch.float64)


def test_mode_sparse_labels_batch_and_ignored_values(engine):
    """Non-dense `values` in a non-sorted order (ties go to the first listed), labels not listed are
    ignored, leading batch dims, int32 data, reversed target."""
    rng = np.random.default_rng(14)
    lat, lon = np.arange(0.5, 40, 1.0), np.arange(0.5, 60, 1.0)
    tlat, tlon = np.arange(2.0, 40, 4.0), np.arange(2.5, 60, 5.0)[::-1].copy()
    labels = np.array([40, 10, 30, 20])
    x = rng.choice(np.array([10, 20, 30, 40, 99]), size=(2, 3, lat.size, lon.size)).astype(np.int32)
    plan = _plan(engine, lat, lon, tlat, tlon)
    for anti in (False, True):
        y = plan.mode(torch.from_numpy(x).cuda(), labels, anti_mode=anti)
        want = orr.mode(x, [lat, lon], [tlat, tlon], [-2, -1], labels, anti_mode=anti)
        assert y.shape == (2, 3, tlat.size, tlon.size)
        np.testing.assert_array_equal(y.cpu().numpy(), want)


@pytest.mark.parametrize("dtype,rtol", [(np.float64, 1e-12), (np.float32, 1e-5)])
@pytest.mark.parametrize("skipna", [False, True])
def test_stats_all_methods(engine, dtype, rtol, skipna):
    """Every statistic on a float field with 2 % NaN (SURVEY 8d, config 5's stat variant, scaled)."""
    rng = np.random.default_rng(17)
    lat, lon = np.round(np.arange(-4.995, 5, 0.01), 3), np.round(np.arange(0.005, 20, 0.01), 3)
    tlat, tlon = np.arange(-4.875, 5, 0.25), np.arange(0.125, 20, 0.25)
    x = rng.normal(size=(2, lat.size, lon.size)).astype(dtype)
    x[rng.random(x.shape) < 0.02] = np.nan
    x[0, :30, :30] = np.nan  # a few all-NaN cells
    plan = _plan(engine, lat, lon, tlat, tlon, ("lat", "lon"))
    xt = torch.from_numpy(x).cuda()
    for method in ["sum", "mean", "var", "std", "min", "max", "median"]:
        y = plan.stat(xt, method, skipna=skipna).cpu().numpy()
        want = orr.regrid_latlon(x, lat, lon, tlat, tlon, "stat", method=method, skipna=skipna)
        assert y.dtype == dtype and want.dtype == dtype
        np.testing.assert_array_equal(np.isnan(y), np.isnan(want), err_msg=method)
        if method in ("min", "max", "median"):
            np.testing.assert_array_equal(y, want, err_msg=method)
        else:
            np.testing.assert_allclose(y, want, rtol=rtol, atol=rtol * 10, err_msg=method)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("cells", [(25, 25), (4, 3), (1, 1), (32, 32)])
def test_median_adversarial_values(engine, dtype, cells):
    """The register-resident select (sentinel keys, threshold counting, smallest / largest-of-range exits) on values
    that stress the key order: heavy ties, +-inf, signed zeros, denormals, the type's extremes, constant cells,
    cells whose valid count is odd / even / 1, for every cell shape class of the kernel's templates."""
    rng = np.random.default_rng(99)
    nr, nc = cells
    n_lat, n_lon = 6, 9
    lat = (np.arange(n_lat * nr) + 0.5) / nr
    lon = (np.arange(n_lon * nc) + 0.5) / nc
    tlat, tlon = np.arange(n_lat) + 0.5, np.arange(n_lon) + 0.5
    fi = np.finfo(dtype)
    big = fi.max / 4  # (big + big) / 2 stays finite in the data's own type (np.median averages in that type)
    pool = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, big, -big, fi.tiny, -fi.tiny, fi.tiny / 4,
                     -fi.tiny / 4, 2.0, 2.0 + 2 * fi.eps, 0.5, 3.0], dtype=dtype)
    fields = [
        pool[rng.integers(0, pool.size, (n_lat * nr, n_lon * nc))],             # everything, heavy ties
        np.round(rng.normal(size=(n_lat * nr, n_lon * nc)) * 2).astype(dtype),  # small integers: long runs of ties
        np.full((n_lat * nr, n_lon * nc), 7.25, dtype=dtype),                   # constant
        (250 + 50 * rng.random((n_lat * nr, n_lon * nc))).astype(dtype),        # one binade or two, no sign change
        np.where(rng.random((n_lat * nr, n_lon * nc)) < 0.5, big, -big).astype(dtype),
    ]
    x = np.stack(fields)
    x[:, rng.random(x.shape[1:]) < 0.15] = np.nan      # odd and even valid counts
    x[1, :nr, :nc] = np.nan                            # an all-NaN cell
    x[2, :nr, :nc] = np.nan
    x[2, 0, 0] = 5.0                                   # a single valid value
    plan = _plan(engine, lat, lon, tlat, tlon)
    for skipna in (True, False):
        y = plan.stat(torch.from_numpy(x).cuda(), "median", skipna=skipna).cpu().numpy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            want = orr.stat(x, [lat, lon], [tlat, tlon], [-2, -1], "median", skipna=skipna)
        assert y.dtype == dtype
        np.testing.assert_array_equal(np.isnan(y), np.isnan(want))
        np.testing.assert_array_equal(y, want)  # equal values (signed zeros compare equal)


@pytest.mark.parametrize("n_lon", [1001, 1003, 1006])
@pytest.mark.parametrize("descending", [False, True])
def test_staging_phase_carried_across_rows(engine, n_lon, descending):
    """The lane kernels (most_common planes, stat) carry the 16-byte phase of a staged row segment from row to row
    ((phase + row stride) mod 16).  Row pitches that are not multiples of 16 bytes (odd widths; uint8, fp32 and fp64
    elements) and a descending source latitude (negative row step) exercise every residue of that update."""
    rng = np.random.default_rng(41 + n_lon)
    lat = (np.arange(60) + 0.5) * 0.1
    lon = (np.arange(n_lon) + 0.5) * 0.1
    tlat = np.arange(0.5, 6.0, 1.0)                 # 10 source rows per cell
    tlon = np.arange(0.5, n_lon * 0.1 - 0.5, 1.0)  # 10 source columns per cell
    if descending:
        lat = lat[::-1].copy()
    plan = _plan(engine, lat, lon, tlat, tlon)
    labels = rng.integers(0, 9, (3, lat.size, lon.size)).astype(np.uint8)
    values = np.arange(9)
    y = plan.mode(torch.from_numpy(labels).cuda(), values).cpu().numpy()
    np.testing.assert_array_equal(y, orr.mode(labels, [lat, lon], [tlat, tlon], [-2, -1], values))
    for dtype, rtol in ((np.float32, 1e-5), (np.float64, 1e-12)):
        x = rng.normal(size=(3, lat.size, lon.size)).astype(dtype)
        x[rng.random(x.shape) < 0.03] = np.nan
        for method in ("mean", "var", "max"):
            got = plan.stat(torch.from_numpy(x).cuda(), method, skipna=True).cpu().numpy()
            want = orr.stat(x, [lat, lon], [tlat, tlon], [-2, -1], method, skipna=True)
            np.testing.assert_array_equal(np.isnan(got), np.isnan(want), err_msg=method)
            np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * 10, err_msg=method)


def test_var_fp32_large_offset(engine):
    """fp32 var / std use one pass with fp64 moments shifted by the cell's first valid value: a field with
    mean / spread = 1e4 (temperatures in K with 0.03 K noise) must still match np.var within fp32 tolerance."""
    rng = np.random.default_rng(23)
    lat, lon = np.round(np.arange(-4.995, 5, 0.01), 3), np.round(np.arange(0.005, 20, 0.01), 3)
    tlat, tlon = np.arange(-4.875, 5, 0.25), np.arange(0.125, 20, 0.25)
    x = (300.0 + 0.03 * rng.normal(size=(1, lat.size, lon.size))).astype(np.float32)
    x[rng.random(x.shape) < 0.05] = np.nan
    x[0, :, ::25] = np.nan  # the first column of every cell is missing: the shift comes from a later sample
    plan = _plan(engine, lat, lon, tlat, tlon, ("lat", "lon"))
    for method in ("var", "std", "mean"):
        y = plan.stat(torch.from_numpy(x).cuda(), method, skipna=True).cpu().numpy()
        want = orr.regrid_latlon(x, lat, lon, tlat, tlon, "stat", method=method, skipna=True)
        np.testing.assert_array_equal(np.isnan(y), np.isnan(want), err_msg=method)
        np.testing.assert_allclose(y, want, rtol=1e-5, err_msg=method)


def test_stat_global_wrap_and_uncovered(engine):
    """Global 1 deg cell-centred source -> 2 deg pole-centred target: longitude wrap padding is applied
    (stats=True skips the pole rows, tests/test_format.py:182-218) and edge bins are half populated."""
    rng = np.random.default_rng(18)
    lat, lon = np.arange(-89.5, 90, 1.0), np.arange(0.5, 360, 1.0)
    tlat, tlon = np.arange(-90.0, 91, 2.0), np.arange(0.0, 361, 2.0)
    x = rng.random((3, lat.size, lon.size))
    plan = _plan(engine, lat, lon, tlat, tlon, ("lat", "lon"))
    for method in ("mean", "max", "sum"):
        y = plan.stat(torch.from_numpy(x).cuda(), method).cpu().numpy()
        want = orr.regrid_latlon(x, lat, lon, tlat, tlon, "stat", method=method)
        np.testing.assert_array_equal(np.isnan(y), np.isnan(want))
        np.testing.assert_allclose(y, want, rtol=1e-12)


def test_errors(engine):
    plan = _plan(engine, C11, C11, T6, T6)
    with pytest.raises(ValueError):
        plan.stat(torch.zeros((11, 11), device="cuda"), "mode")  # flox_reduce.py:70-73
    with pytest.raises(ValueError):
        plan.mode(torch.zeros((11, 11), device="cuda"), LABELS)  # flox_reduce.py:156-162


def test_empty_batch_large_bins_and_median_paths(engine):
    """Empty stack; bins wider than 32 columns (bitonic median path); many labels (tile + match.any path)."""
    rng = np.random.default_rng(22)
    lat, lon = np.arange(0.5, 40, 1.0), np.arange(0.5, 200, 1.0)
    tlat, tlon = np.arange(5.0, 40, 10.0), np.arange(25.0, 200, 50.0)  # 10 x 50 cells per bin
    plan = _plan(engine, lat, lon, tlat, tlon)
    assert plan.stat(torch.empty((0, lat.size, lon.size), device="cuda"), "mean").shape == (0, 4, 4)
    x = rng.normal(size=(2, lat.size, lon.size))
    x[rng.random(x.shape) < 0.05] = np.nan
    for skipna in (False, True):
        y = plan.stat(torch.from_numpy(x).cuda(), "median", skipna=skipna).cpu().numpy()
        want = orr.stat(x, [lat, lon], [tlat, tlon], [-2, -1], "median", skipna=skipna)
        np.testing.assert_array_equal(y, want)
    labels = np.arange(100)  # > 64 labels: the warp-per-cell kernel
    xi = rng.integers(0, 100, size=(2, lat.size, lon.size)).astype(np.int16)
    y = plan.mode(torch.from_numpy(xi).cuda(), labels).cpu().numpy()
    np.testing.assert_array_equal(y, orr.mode(xi, [lat, lon], [tlat, tlon], [-2, -1], labels))


def test_latitude_band_sharding_matches_full_slab(engine):
    """sharding.latitude_band: a single slab split by target rows over 3 'ranks' (run one after the other on this
    GPU) reproduces the full most_common / mean result exactly, with no exchange between the bands."""
    from xarray_regrid_b200 import sharding

    rng = np.random.default_rng(29)
    lat, lon = np.round(np.arange(-19.95, 20, 0.1), 6), np.round(np.arange(0.05, 30, 0.1), 6)
    tlat, tlon = np.arange(-18.0, 19.0, 1.0), np.arange(1.0, 29.0, 1.0)
    labels = torch.from_numpy(rng.integers(0, 12, (lat.size, lon.size)).astype(np.uint8)).cuda()
    field = torch.from_numpy(rng.standard_normal((lat.size, lon.size)).astype(np.float32)).cuda()
    full = _plan(engine, lat, lon, tlat, tlon, ("lat", "lon"))
    want_mode = full.mode(labels, np.arange(12), fill_value=255).cpu().numpy()
    want_mean = full.stat(field, "mean", skipna=True).cpu().numpy()
    got_mode, got_mean = np.full_like(want_mode, 254), np.full_like(want_mean, np.nan)
    for rank in range(3):
        t_idx, s_rows = sharding.latitude_band(lat, tlat, 3, rank)
        band = _plan(engine, lat[s_rows], lon, tlat[t_idx], tlon, ("lat", "lon"))
        got_mode[t_idx] = band.mode(labels[s_rows], np.arange(12), fill_value=255).cpu().numpy()
        got_mean[t_idx] = band.stat(field[s_rows], "mean", skipna=True).cpu().numpy()
    np.testing.assert_array_equal(got_mode, want_mode)
    np.testing.assert_array_equal(got_mean, want_mean)


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.int16, np.int32, np.int64])
@pytest.mark.parametrize("labels", ["dense12", "dense32", "dense40", "dense100", "sparse"])
def test_mode_every_kernel_and_label_dtype(engine, dtype, labels):
    """All three mode kernels (bit-sliced registers <= 32 dense labels, lane counters <= 64, tile + match.any
    otherwise / sparse label sets) for every integer dtype, with values that are NOT in the label list mixed in
    (the reference ignores them) and both most_common and least_common."""
    rng = np.random.default_rng(41)
    lat, lon = np.round(np.arange(-9.95, 10, 0.1), 6), np.round(np.arange(0.05, 16, 0.1), 6)
    tlat, tlon = np.arange(-9.0, 10.0, 1.0), np.arange(1.0, 15.0, 1.0)
    if labels == "sparse":
        values = np.array([3, 7, 19, 40, 77, 100, 120], dtype=dtype)
    else:
        values = np.arange(int(labels[5:]), dtype=dtype)
    pool = np.concatenate([values, np.array([values.max() + 1, 126], dtype=dtype)])  # two foreign values
    if np.issubdtype(dtype, np.signedinteger):
        pool = np.concatenate([pool, np.array([-1], dtype=dtype)])
    x = pool[rng.integers(0, pool.size, (2, lat.size, lon.size))]
    plan = _plan(engine, lat, lon, tlat, tlon, ("lat", "lon"))
    for anti in (False, True):
        y = plan.mode(torch.from_numpy(x).cuda(), values, fill_value=int(values.max()), anti_mode=anti)
        want = orr.regrid_latlon(x, lat, lon, tlat, tlon, "least_common" if anti else "most_common", values=values,
                                 fill_value=int(values.max()))
        assert y.cpu().numpy().dtype == want.dtype == dtype
        np.testing.assert_array_equal(y.cpu().numpy(), want)
