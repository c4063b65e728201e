"""Parity of the CUDA conservative path (through the C ABI) against the oracle.

Tolerances are BASELINE.json's: 1e-12 relative for fp64, 1e-5 relative for fp32, NaN
placement identical.  Inputs are positive (0.5 + U[0,1)) so relative error is meaningful
("parity metric").
"""

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import conservative as oc  # noqa: E402

pytestmark = pytest.mark.gpu

RTOL = {np.float64: 1e-12, np.float32: 1e-5}


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xarray_regrid_b200 import engine, synthetic

    return engine, synthetic


def _check(got, want, rtol):
    got = got.cpu().numpy() if hasattr(got, "cpu") else got
    assert got.shape == want.shape
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want), err_msg="NaN placement differs")
    np.testing.assert_allclose(got, want, rtol=rtol, atol=0)


def _latlon_plan(engine, lat, lon, tlat, tlon):
    return engine.ConservativePlan(
        engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon")
    )


def _pick(plan, kernel):
    """kernel: "lanes" (register-resident ring variant), "ring" (two-phase ring kernel) or "generic"."""
    plan.use_ring = kernel != "generic"
    plan.use_lanes = kernel == "lanes"


KERNELS = ["lanes", "ring", "generic"]


@pytest.mark.parametrize("skipna", [False, True])
@pytest.mark.parametrize("ring", KERNELS)
def test_c1_fp64(eng, skipna, ring):
    """BASELINE config 1: one 721x1440 fp64 step -> 181x360, no NaNs (every kernel variant)."""
    engine, syn = eng
    lat, lon = syn.grid_025()
    tlat, tlon = syn.grid_1()
    x = syn.conservative_stack(1, lat, lon, nan_fraction=0.0)
    plan = _latlon_plan(engine, lat, lon, tlat, tlon)
    _pick(plan, ring)
    y = plan(x, skipna=skipna)
    want = oc.regrid_latlon(x.cpu().numpy(), lat, lon, tlat, tlon, skipna=skipna)
    assert y.shape == (1, 181, 360) and y.dtype == torch.float64
    _check(y, want, 1e-12)


@pytest.mark.parametrize("nan_threshold", [0.0, 0.25, 0.5, 1.0])
@pytest.mark.parametrize("ring", KERNELS)
def test_c2_nan_semantics_fp64(eng, nan_threshold, ring):
    """BASELINE config 2 semantics on 3 steps: 10 % NaN, skipna=True, joint 2-D valid mass."""
    engine, syn = eng
    lat, lon = syn.grid_025()
    tlat, tlon = syn.grid_1()
    x = syn.conservative_stack(3, lat, lon, nan_fraction=0.10, chunk=2)
    x[2, 100:140, 200:300] = float("nan")  # a block that wipes out whole target cells
    plan = _latlon_plan(engine, lat, lon, tlat, tlon)
    _pick(plan, ring)
    y = plan(x, skipna=True, nan_threshold=nan_threshold)
    want = oc.regrid_latlon(x.cpu().numpy(), lat, lon, tlat, tlon, skipna=True,
                            nan_threshold=nan_threshold)
    assert 0 < np.isnan(want).sum() < want.size
    _check(y, want, 1e-12)


@pytest.mark.parametrize("ring", KERNELS)
def test_fp32(eng, ring):
    engine, syn = eng
    lat, lon = syn.grid_025()
    tlat, tlon = syn.grid_1()
    x = syn.conservative_stack(2, lat, lon, dtype=torch.float32, nan_fraction=0.10)
    plan = _latlon_plan(engine, lat, lon, tlat, tlon)
    _pick(plan, ring)
    y = plan(x, skipna=True, nan_threshold=0.5)
    assert y.dtype == torch.float32
    want = oc.regrid_latlon(x.cpu().numpy(), lat, lon, tlat, tlon, skipna=True, nan_threshold=0.5)
    assert want.dtype == np.float32
    _check(y, want, 1e-5)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("skipna", [False, True])
@pytest.mark.parametrize("aligned_edges", [False, True])
def test_lane_owned_columns_variant(eng, dtype, skipna, aligned_edges):
    """The register-resident kernel on small global 4:1 grids: target cells centred on source centres
    (5 taps, half weights at both ends, wrap at the date line) and target edges on source edges (4 taps:
    the 5th, zero-weight column must not leak its NaN).  Checked against the oracle AND the generic kernel."""
    engine, _ = eng
    from xarray_regrid_b200 import tables

    if aligned_edges:
        lat, lon = np.arange(-89.5, 90, 1.0), np.arange(0.5, 360, 1.0)
        tlat, tlon = np.arange(-88.0, 90, 4.0), np.arange(2.0, 360, 4.0)
    else:
        lat, lon = np.arange(-90.0, 90.5, 1.0), np.arange(0.0, 360, 1.0)
        tlat, tlon = np.arange(-88.0, 89, 4.0), np.arange(0.0, 360, 4.0)
    rng = np.random.default_rng(31)
    x = (0.5 + rng.random((5, lat.size, lon.size))).astype(dtype)
    x[rng.random(x.shape) < 0.08] = np.nan
    x[1, :, 40:48] = np.nan      # whole target columns missing
    x[2, 60:70, :] = np.nan      # whole target rows missing
    x[3] = 1.25                  # NaN-free slab: the fast path only
    plan = _latlon_plan(engine, lat, lon, tlat, tlon)
    assert not plan.pole_rows
    xt = torch.from_numpy(x).cuda()
    ring = plan._ring(5, lat.size, lon.size, xt.dtype)
    assert ring is not None and ring.host.lane_cols == 4, "grid should qualify for the lanes variant"
    assert int(ring.host.n_strips) <= int(ring.host.n_warps) <= 12
    y = plan(xt, skipna=skipna, nan_threshold=0.5)
    # skipna=False: a NaN counts as 0 (the reference's unconditional fillna(0))
    want = oc.regrid_latlon(x, lat, lon, tlat, tlon, skipna=skipna, nan_threshold=0.5)
    _check(y, want, RTOL[dtype])
    plan.use_ring = False
    _check(plan(xt, skipna=skipna, nan_threshold=0.5), want, RTOL[dtype])
    assert tables.lane_owned_columns(plan.cols_band, ring.host.strip_l0, ring.host.n_warps) == 4


def test_skipna_false_counts_nan_as_zero(eng):
    """apply_weights contracts da.fillna(0) whether or not skipna is set (reference
    methods/conservative.py:196-198): with skipna=False a NaN counts as 0 and every target stays finite,
    on all three kernels and through the host pipeline."""
    engine, syn = eng
    lat, lon = syn.grid_025()
    tlat, tlon = syn.grid_1()
    x = syn.conservative_stack(2, lat, lon, nan_fraction=0.0)
    x[0, 360, 720] = float("nan")
    x[0, 0, 5] = float("nan")
    x[1, 100:140, 300:420] = float("nan")  # whole target cells missing -> exactly 0
    plan = _latlon_plan(engine, lat, lon, tlat, tlon)
    want = oc.regrid_latlon(x.cpu().numpy(), lat, lon, tlat, tlon, skipna=False)
    assert np.isfinite(want).all() and (want == 0).any()
    for kernel in KERNELS:
        _pick(plan, kernel)
        _check(plan(x, skipna=False), want, 1e-12)
    _pick(plan, "lanes")
    _check(plan.apply_host(x.cpu().pin_memory(), skipna=False, chunk_batch=1), want, 1e-12)
    _check(plan.apply_host(x.cpu().numpy(), skipna=False, chunk_batch=1), want, 1e-12)


def test_reference_known_answer_joint_valid_mass(eng):
    """reference tests/test_regrid.py:172-185 through the CUDA path."""
    engine, _ = eng
    data = np.broadcast_to(np.array([[1, np.nan], [2, 2]]), (3, 2, 2)).copy()
    data[2] = np.nan
    c = np.array([-1.0, 1.0])
    plan = engine.ConservativePlan(engine.AxisSpec(c, np.array([0.0])), engine.AxisSpec(c, np.array([0.0])))
    y = plan(torch.from_numpy(data).cuda(), skipna=True, nan_threshold=1).cpu().numpy()
    assert y.shape == (3, 1, 1)
    np.testing.assert_allclose(y[0, 0, 0], 5.0 / 3.0, rtol=1e-15)
    np.testing.assert_allclose(y[1, 0, 0], 5.0 / 3.0, rtol=1e-15)
    assert np.isnan(y[2, 0, 0])


@pytest.mark.parametrize("nan_threshold", [0, 1])
def test_reference_known_answer_coarsen(eng, nan_threshold):
    """reference tests/test_regrid.py:188-208 through the CUDA path."""
    engine, _ = eng
    nan = np.nan
    da = np.array([[1, nan, nan, nan], [2, 2, nan, nan], [3, 3, 3, nan], [4, 4, 4, 4.0]])
    src, tgt = np.array([0.0, 1, 2, 3]), np.array([0.5, 2.5])
    plan = engine.ConservativePlan(engine.AxisSpec(src, tgt), engine.AxisSpec(src, tgt))
    y = plan(torch.from_numpy(da).cuda(), skipna=True, nan_threshold=nan_threshold).cpu().numpy()
    want = oc.regrid(da, [src, src], [tgt, tgt], [-2, -1], skipna=True, nan_threshold=nan_threshold)
    np.testing.assert_array_equal(np.isnan(y), np.isnan(want))
    np.testing.assert_allclose(y, want, rtol=1e-15)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_cell_centred_grid_with_pole_rows_and_wrap(eng, dtype):
    """2 degree cell-centred source (-89..89, 1..359) -> 1 degree pole-centred target: the
    reference adds pole rows (lon-mean of the edge rows) and wrap columns (tests/test_format.py:35-63)."""
    engine, _ = eng
    rng = np.random.default_rng(5)
    lat, lon = np.arange(-89.0, 90, 2), np.arange(1.0, 360, 2)
    tlat, tlon = np.arange(-90.0, 91, 1), np.arange(0.0, 361, 1)
    x = (0.5 + rng.random((4, lat.size, lon.size))).astype(dtype)
    x[1, 0, :5] = np.nan
    x[2, -1, :] = np.nan  # all-NaN edge row -> NaN pole value
    x[3, 40:44, 100:104] = np.nan
    plan = _latlon_plan(engine, lat, lon, tlat, tlon)
    assert plan.rows_band.meta["pole_rows"]
    for kernel in ("ring", "generic"):  # the pole rows are ordinary rows of the TMA ring
        _pick(plan, kernel)
        if kernel == "ring":
            assert plan._ring(4, lat.size, lon.size, torch.from_numpy(x).dtype) is not None
        for skipna, thr in [(True, 1.0), (True, 0.3), (False, 1.0)]:
            y = plan(torch.from_numpy(x).cuda(), skipna=skipna, nan_threshold=thr)
            want = oc.regrid_latlon(x, lat, lon, tlat, tlon, skipna=skipna, nan_threshold=thr)
            _check(y, want, RTOL[dtype])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_cell_centred_quarter_degree_grid_all_kernels(eng, dtype):
    """Cell-centred 0.25 deg global grid (-89.875 .. 89.875, 0.125 .. 359.875) -> 1 deg: format_lat adds pole
    rows (zero area under the spherical correction), target edges fall on source edges (4 taps), the columns
    qualify for the register-resident kernel after re-stripping; checked against the oracle for every kernel
    and through the host-buffer pipeline."""
    engine, _ = eng
    rng = np.random.default_rng(6)
    lat, lon = np.arange(-89.875, 90, 0.25), np.arange(0.125, 360, 0.25)
    tlat, tlon = np.arange(-90.0, 91, 1.0), np.arange(0.0, 360, 1.0)
    x = (0.5 + rng.random((3, lat.size, lon.size))).astype(dtype)
    x[rng.random(x.shape) < 0.05] = np.nan
    x[1, 0, :] = np.nan  # all-NaN edge row -> NaN pole row
    plan = _latlon_plan(engine, lat, lon, tlat, tlon)
    assert plan.rows_band.meta["pole_rows"]
    xt = torch.from_numpy(x).cuda()
    ring = plan._ring(3, lat.size, lon.size, xt.dtype)
    assert ring is not None and ring.host.lane_cols == 4
    want = oc.regrid_latlon(x, lat, lon, tlat, tlon, skipna=True, nan_threshold=0.5)
    for kernel in KERNELS:
        _pick(plan, kernel)
        _check(plan(xt, skipna=True, nan_threshold=0.5), want, RTOL[dtype])
    _pick(plan, "lanes")
    y_host = plan.apply_host(torch.from_numpy(x).pin_memory(), skipna=True, nan_threshold=0.5, chunk_batch=2)
    _check(y_host, want, RTOL[dtype])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_pole_rows_with_weight_through_ring_and_host_pipeline(eng, dtype):
    """Without the spherical correction the synthetic pole rows keep a non-zero weight: they are fetched by
    the TMA producer from the [batch, 2, W] pole buffer (ring kernels), read as scalars by the generic kernel,
    and produced per chunk inside the host-buffer pipeline."""
    engine, _ = eng
    rng = np.random.default_rng(8)
    lat, lon = np.arange(-89.0, 90, 2), np.arange(1.0, 360, 2)
    tlat, tlon = np.arange(-90.0, 91, 1), np.arange(0.0, 361, 1)
    x = (0.5 + rng.random((5, lat.size, lon.size))).astype(dtype)
    x[1, 0, :7] = np.nan
    x[2, -1, :] = np.nan  # all-NaN edge row -> NaN pole row
    plan = engine.ConservativePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"),
                                   spherical_rows=False)
    assert plan.pole_rows, "the pole rows must be referenced by the rows band"
    xt = torch.from_numpy(x).cuda()
    for skipna, thr in [(True, 0.5), (False, 1.0)]:
        want = oc.regrid_latlon(x, lat, lon, tlat, tlon, skipna=skipna, nan_threshold=thr,
                                latitude_weighting=False)
        for kernel in ("ring", "generic"):
            _pick(plan, kernel)
            if kernel == "ring":
                ring = plan._ring(5, lat.size, lon.size, xt.dtype)
                assert ring is not None and (ring.host.sched_row >= lat.size).any()
            _check(plan(xt, skipna=skipna, nan_threshold=thr), want, RTOL[dtype])
        _pick(plan, "ring")
        y_host = plan.apply_host(torch.from_numpy(x).pin_memory(), skipna=skipna, nan_threshold=thr, chunk_batch=2)
        _check(y_host, want, RTOL[dtype])


def test_descending_source_and_reversed_target(eng):
    """ERA5 stores latitude 90 -> -90; targets may be reversed (tests/test_regrid.py:289-300)."""
    engine, syn = eng
    lat, lon = syn.grid_025()
    tlat, tlon = syn.grid_1()
    x = syn.conservative_stack(2, lat, lon, nan_fraction=0.10)
    plan = _latlon_plan(engine, lat[::-1].copy(), lon, tlat[::-1].copy(), tlon[::-1].copy())
    y = plan(torch.flip(x, dims=[1]), skipna=True, nan_threshold=0.5)
    want = oc.regrid_latlon(x.cpu().numpy()[:, ::-1], lat[::-1], lon, tlat[::-1], tlon[::-1],
                            skipna=True, nan_threshold=0.5)
    _check(y, want, 1e-12)
    base = _latlon_plan(engine, lat, lon, tlat, tlon)(x, skipna=True, nan_threshold=0.5)
    _check(torch.flip(base, dims=[1, 2]), y.cpu().numpy(), 1e-13)


def test_regional_target_and_uncovered_cells(eng):
    """Target partly outside a regional source: uncovered centres -> NaN, covered edge cells
    are renormalised (methods/conservative.py:123-125, 161-164; utils.py:165-169)."""
    engine, _ = eng
    rng = np.random.default_rng(9)
    lat, lon = np.arange(30.0, 60.25, 0.25), np.arange(-10.0, 30.25, 0.25)
    tlat, tlon = np.arange(25.0, 66, 1.0), np.arange(-15.0, 36, 1.0)
    x = 0.5 + rng.random((2, lat.size, lon.size))
    plan = _latlon_plan(engine, lat, lon, tlat, tlon)
    y = plan(torch.from_numpy(x).cuda(), skipna=True)
    want = oc.regrid_latlon(x, lat, lon, tlat, tlon, skipna=True)
    assert np.isnan(want).any() and not np.isnan(want).all()
    _check(y, want, 1e-12)


def test_upsampling_and_odd_unaligned_width(eng):
    """Fine target (taps <= 2), odd source width (scalar-load path) and a strided view."""
    engine, _ = eng
    rng = np.random.default_rng(11)
    lat, lon = np.arange(-10.0, 10.5, 1.0), np.arange(100.0, 133.0, 1.0)  # 21 x 33
    tlat, tlon = np.arange(-9.0, 9.1, 0.3), np.arange(101.0, 131.0, 0.4)
    big = 0.5 + rng.random((3, lat.size, lon.size + 7))
    big[1, 5, 5] = np.nan
    xt = torch.from_numpy(big).cuda()[:, :, 3:3 + lon.size]  # row stride 40, offset 3: unaligned
    assert not xt.is_contiguous()
    plan = _latlon_plan(engine, lat, lon, tlat, tlon)
    y = plan(xt, skipna=True, nan_threshold=0.9)
    want = oc.regrid_latlon(big[:, :, 3:3 + lon.size], lat, lon, tlat, tlon, skipna=True,
                            nan_threshold=0.9)
    _check(y, want, 1e-12)


def test_single_axis_and_integer_input(eng):
    """Only one shared coordinate (rows); int32 data -> float64 like np.result_type(float32, int32)."""
    engine, _ = eng
    rng = np.random.default_rng(13)
    src, tgt = np.arange(0.0, 40.0, 1.0), np.arange(1.5, 38.0, 4.0)
    x = rng.integers(1, 100, size=(5, src.size, 17)).astype(np.int32)
    plan = engine.ConservativePlan(engine.AxisSpec(src, tgt), None)
    y = plan(torch.from_numpy(x).cuda(), skipna=False)
    want = oc.regrid(x, [src], [tgt], [-2], skipna=False)
    assert y.dtype == torch.float64 and want.dtype == np.float64
    _check(y, want, 1e-12)


def test_empty_batch_and_leading_dims(eng):
    engine, syn = eng
    lat, lon = syn.grid_025()
    tlat, tlon = syn.grid_1()
    plan = _latlon_plan(engine, lat, lon, tlat, tlon)
    y = plan(torch.empty((0, 721, 1440), dtype=torch.float64, device="cuda"))
    assert y.shape == (0, 181, 360)
    x = syn.conservative_stack(6, lat, lon, nan_fraction=0.1).view(2, 3, 721, 1440)
    y = plan(x, skipna=True, nan_threshold=0.5)
    assert y.shape == (2, 3, 181, 360)
    flat = plan(x.view(6, 721, 1440), skipna=True, nan_threshold=0.5)
    assert torch.equal(torch.nan_to_num(y.view(6, 181, 360)), torch.nan_to_num(flat))


def test_host_buffer_pipeline_matches_device_path(eng):
    """rg_conservative_host_*: chunked H2D/kernel/D2H must give bit-identical results."""
    engine, syn = eng
    lat, lon = syn.grid_025()
    tlat, tlon = syn.grid_1()
    x = syn.conservative_stack(11, lat, lon, nan_fraction=0.10)
    plan = _latlon_plan(engine, lat, lon, tlat, tlon)
    want = plan(x, skipna=True, nan_threshold=0.5).cpu()
    xh = x.cpu().pin_memory()
    got = plan.apply_host(xh, skipna=True, nan_threshold=0.5, chunk_batch=4)
    assert torch.equal(torch.nan_to_num(got, nan=-1.0), torch.nan_to_num(want, nan=-1.0))


def test_errors(eng):
    engine, syn = eng
    lat, lon = syn.grid_025()
    tlat, tlon = syn.grid_1()
    plan = _latlon_plan(engine, lat, lon, tlat, tlon)
    with pytest.raises(ValueError):
        plan(torch.zeros((1, 721, 1440), device="cuda"), nan_threshold=1.5)  # regrid.py:118-120
    with pytest.raises(ValueError):
        plan(torch.zeros((1, 720, 1440), device="cuda"))


def test_properties_at_scale(eng):
    """Size-independent properties on a 512-step stack (3.9 GB): constants are preserved,
    the map is linear, and sampled slabs match the oracle."""
    engine, syn = eng
    lat, lon = syn.grid_025()
    tlat, tlon = syn.grid_1()
    plan = _latlon_plan(engine, lat, lon, tlat, tlon)
    steps = 512
    x = syn.conservative_stack(steps, lat, lon, nan_fraction=0.10, seed=99)
    y = plan(x, skipna=True, nan_threshold=0.5)
    for s in (0, 1, steps // 2, steps - 1):
        want = oc.regrid_latlon(x[s].cpu().numpy(), lat, lon, tlat, tlon, skipna=True, nan_threshold=0.5)
        _check(y[s], want, 1e-12)
    # the generic (non-TMA) kernel must agree with the ring kernel to the last few bits
    plan.use_ring = False
    yg = plan(x[:64], skipna=True, nan_threshold=0.5)
    plan.use_ring = True
    assert torch.equal(torch.isnan(yg), torch.isnan(y[:64]))
    assert torch.allclose(torch.nan_to_num(yg), torch.nan_to_num(y[:64]), rtol=1e-14, atol=0)
    # constant field -> the same constant wherever defined (weights are normalised)
    c = torch.full((4, 721, 1440), 3.25, dtype=torch.float64, device="cuda")
    yc = plan(c, skipna=True)
    assert torch.allclose(yc, torch.full_like(yc, 3.25), rtol=1e-14, atol=0)
    # linearity in the data for a fixed mask: R(2a + 3) == 2 R(a) + 3
    y2 = plan(x[:8] * 2 + 3, skipna=True, nan_threshold=0.5)
    ref = y[:8] * 2 + 3
    assert torch.equal(torch.isnan(y2), torch.isnan(ref))
    assert torch.allclose(torch.nan_to_num(y2), torch.nan_to_num(ref), rtol=1e-13, atol=0)
