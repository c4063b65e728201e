"""Round-2 parity cases: pole rows on descending / shuffled latitudes, latitude-only pole rows (plain copies),
the streaming host paths of every plan (page-locked and pageable memory) and the accessor's host routing."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import conservative as oc  # noqa: E402
from oracle import formatting as of  # noqa: E402
from oracle import interp as oi  # noqa: E402
from oracle import reduce as orr  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xarray_regrid_b200 import engine as e

    return e


def _check(got, want, rtol, atol=0.0):
    got = got.cpu().numpy() if hasattr(got, "cpu") else np.asarray(got)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol)


def _lat_orders():
    asc = np.arange(-89.0, 90.0, 2.0)
    rng = np.random.default_rng(5)
    return {"descending": asc[::-1].copy(), "shuffled": rng.permutation(asc)}


# ------------------------------------------------------------------ pole rows of a re-ordered latitude
@pytest.mark.parametrize("order", ["descending", "shuffled"])
@pytest.mark.parametrize("method", ["linear", "nearest", "cubic"])
def test_pole_rows_descending_and_shuffled_latitude_interp(engine, order, method):
    """The reference sorts the latitude first (ensure_monotonic, utils.py:277) and takes the pole means from
    the first / last SORTED row (utils.py:320-333): the south pole is the mean of the southernmost row
    wherever that row sits in the caller's array."""
    lat = _lat_orders()[order]
    lon = np.arange(1.0, 360.0, 2.0)
    tlat, tlon = np.arange(-90.0, 90.5, 1.5), np.arange(0.0, 360.0, 3.0)
    rng = np.random.default_rng(17)
    # a strong latitude gradient: the two pole means differ by ~178, so a swapped pole row cannot hide
    x = lat[None, :, None] + rng.random((3, lat.size, lon.size))
    plan = engine.InterpPlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"), method)
    assert plan.rows_axis.has_pole_rows
    y = plan(torch.from_numpy(x).cuda())
    want = oi.regrid_latlon(x, lat, lon, tlat, tlon, method)
    scale = np.nanmax(np.abs(want))
    # (nearest is exact except for the pole MEANS, whose summation order differs from numpy's by an ulp)
    _check(y, want, 1e-14 if method == "nearest" else 1e-11, atol=0.0 if method == "nearest" else 1e-11 * scale)
    # the pole targets themselves
    south = np.nanmean(x[:, np.argmin(lat), :], axis=-1)
    if method != "cubic":
        np.testing.assert_allclose(y[:, 0, :].cpu().numpy(), np.broadcast_to(south[:, None], (3, tlon.size)),
                                   rtol=1e-12)


@pytest.mark.parametrize("order", ["descending", "shuffled"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_pole_rows_descending_latitude_conservative_plain_rows(engine, order, dtype):
    """Non-spherical conservative keeps a non-zero weight on the pole rows: all kernels and the host pipeline."""
    lat = _lat_orders()[order]
    lon = np.arange(1.0, 360.0, 2.0)
    tlat, tlon = np.arange(-90.0, 91.0, 1.0), np.arange(0.0, 361.0, 1.0)
    rng = np.random.default_rng(23)
    x = (lat[None, :, None] + rng.random((4, lat.size, lon.size))).astype(dtype)
    x[1, np.argmin(lat), :9] = np.nan
    x[2, np.argmax(lat), :] = np.nan
    plan = engine.ConservativePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"),
                                   spherical_rows=False)
    assert plan.pole_rows and plan.pole_src == (int(np.argmin(lat)), int(np.argmax(lat)))
    rtol = 1e-12 if dtype == np.float64 else 2e-5
    for skipna, thr in [(True, 0.5), (False, 1.0)]:
        want = oc.regrid_latlon(x, lat, lon, tlat, tlon, skipna=skipna, nan_threshold=thr,
                                latitude_weighting=False)
        for use_ring in (True, False):
            plan.use_ring = use_ring
            _check(plan(torch.from_numpy(x).cuda(), skipna=skipna, nan_threshold=thr), want, rtol, atol=rtol)
        plan.use_ring = True
        _check(plan.apply_host(torch.from_numpy(x).pin_memory(), skipna=skipna, nan_threshold=thr, chunk_batch=3),
               want, rtol, atol=rtol)
        _check(plan.apply_host(x, skipna=skipna, nan_threshold=thr, chunk_batch=3), want, rtol, atol=rtol)


@pytest.mark.parametrize("method", ["linear", "nearest", "cubic", "conservative"])
def test_latitude_only_pole_rows_are_copies(method):
    """Without a lon / longitude coordinate format_lat pads the poles with a plain COPY of the edge row
    (utils.py:322-323, 330-331 only average when lon exists): a (time, lat) array must not be averaged over
    time, and (lat, x) must not be averaged over x."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xarray_regrid_b200.labelled import DataArray, Dataset

    lat = np.arange(89.5, -90.0, -1.0)  # descending, cell-centred, global
    tlat = np.arange(-90.0, 90.5, 2.5)
    rng = np.random.default_rng(3)
    for dims, shape in ((("time", "lat"), (5, lat.size)), (("lat", "x"), (lat.size, 7))):
        x = rng.random(shape) + 10.0 * np.arange(shape[0])[:, None]
        coords = {d: (lat if d == "lat" else np.arange(float(n))) for d, n in zip(dims, shape)}
        da = DataArray(x, dims, coords)
        tgt = Dataset({}, {"lat": tlat})
        fn = getattr(da.regrid, method)
        out = fn(tgt, skipna=False) if method == "conservative" else fn(tgt)
        ax = dims.index("lat")
        data, flat, _ = of.format_for_regrid(x, lat, None, tlat, None, lat_axis=ax - 2, lon_axis=None)
        if method == "conservative":
            want = oc.regrid(data, [flat], [tlat], [ax], latitude_index=0, skipna=False)
        else:
            want = oi.regrid(data, [flat], [tlat], [ax], method)
        assert out.dims == dims
        np.testing.assert_array_equal(np.isnan(out.data), np.isnan(want))
        np.testing.assert_allclose(out.data, want, rtol=1e-10, atol=1e-10)


# ------------------------------------------------------------------------------- streaming host paths
@pytest.mark.parametrize("method", ["linear", "nearest", "cubic"])
@pytest.mark.parametrize("pinned", [True, False])
def test_interp_apply_host_matches_device_path(engine, method, pinned):
    """InterpPlan.apply_host streams chunks through the persistent pipeline; the result is bit-identical to
    the device path and the device only ever holds `slots` chunks."""
    lat, lon = np.arange(-90.0, 91.0, 3.0), np.arange(0.0, 360.0, 3.0)
    tlat, tlon = np.arange(-90.0, 90.1, 0.75), np.arange(0.0, 359.9, 0.75)
    rng = np.random.default_rng(11)
    if method == "nearest":
        x = rng.integers(0, 50, (11, lat.size, lon.size)).astype(np.int32)
    else:
        x = (250 + 50 * rng.random((11, lat.size, lon.size))).astype(np.float32)
    plan = engine.InterpPlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"), method)
    ref = plan(torch.from_numpy(x).cuda()).cpu()
    xh = torch.from_numpy(x).pin_memory() if pinned else x
    pipe = engine.HostPipeline(slots=3)
    got = plan.apply_host(xh, chunk_batch=4, pipeline=pipe)
    assert got.dtype == ref.dtype and got.shape == ref.shape
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref, nan=-7.0))
    want = oi.regrid_latlon(x, lat, lon, tlat, tlon, method)
    _check(got, want, 0.0 if method == "nearest" else 1e-10, atol=0.0 if method == "nearest" else 1e-9)
    # a second call reuses the pipeline, its buffers and (for "same") a caller-provided result buffer
    out = torch.empty(ref.shape, dtype=ref.dtype)
    again = plan.apply_host(xh, out_host=out, chunk_batch=5, pipeline=pipe)
    assert again.data_ptr() == out.data_ptr()
    assert torch.equal(torch.nan_to_num(again, nan=-7.0), torch.nan_to_num(ref, nan=-7.0))
    pipe.close()


def test_cubic_host_path_nan_anywhere_poisons_everything(engine):
    """scipy's cubic: one NaN anywhere in the array makes the whole result NaN -- also when the NaN sits in a
    later chunk than the ones already copied out."""
    lat, lon = np.arange(-60.0, 61.0, 4.0), np.arange(0.0, 100.0, 4.0)
    tlat, tlon = np.arange(-58.0, 58.0, 1.0), np.arange(2.0, 90.0, 1.0)
    x = np.random.default_rng(2).random((6, lat.size, lon.size))
    x[5, 3, 3] = np.nan
    plan = engine.InterpPlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"), "cubic")
    got = plan.apply_host(x, chunk_batch=2)
    assert torch.isnan(got).all()


@pytest.mark.parametrize("pinned", [True, False])
def test_reduce_host_paths_match_device_path(engine, pinned):
    lat, lon = np.arange(-29.95, 30, 0.1), np.arange(0.05, 90, 0.1)
    tlat, tlon = np.arange(-29.5, 30, 1.0), np.arange(0.5, 90, 1.0)
    rng = np.random.default_rng(9)
    lab = rng.integers(0, 12, (3, lat.size, lon.size)).astype(np.uint8)
    fld = rng.standard_normal((3, lat.size, lon.size)).astype(np.float32)
    fld[rng.random(fld.shape) < 0.02] = np.nan
    plan = engine.ReducePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"))
    wrap = (lambda a: torch.from_numpy(a).pin_memory()) if pinned else (lambda a: a)
    values = np.arange(12)
    ref = plan.mode(torch.from_numpy(lab).cuda(), values).cpu()
    got = plan.mode_host(wrap(lab), values, chunk_batch=2)
    assert got.dtype == torch.uint8 and torch.equal(got, ref)
    want = orr.regrid_latlon(lab, lat, lon, tlat, tlon, "most_common", values=values)
    np.testing.assert_array_equal(got.numpy(), want)
    for method in ("mean", "var", "median", "max"):
        ref = plan.stat(torch.from_numpy(fld).cuda(), method, skipna=True).cpu()
        got = plan.stat_host(wrap(fld), method, skipna=True, chunk_batch=2)
        assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref, nan=-7.0))


def test_apply_host_validates_out_host(engine):
    lat, lon = np.arange(-90.0, 91.0, 1.0), np.arange(0.0, 360.0, 1.0)
    tlat, tlon = np.arange(-88.0, 89, 4.0), np.arange(0.0, 360, 4.0)
    plan = engine.ConservativePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"))
    x = torch.rand(4, lat.size, lon.size, dtype=torch.float64).pin_memory()
    good = (4, tlat.size, tlon.size)
    with pytest.raises(ValueError, match="dtype"):
        plan.apply_host(x, out_host=torch.empty(good, dtype=torch.float32))
    with pytest.raises(ValueError, match="elements"):
        plan.apply_host(x, out_host=torch.empty((3,) + good[1:], dtype=torch.float64))
    with pytest.raises(ValueError, match="contiguous"):
        plan.apply_host(x, out_host=torch.empty(good[::-1], dtype=torch.float64).permute(2, 1, 0))
    with pytest.raises(ValueError, match="CPU"):
        plan.apply_host(x, out_host=torch.empty(good, dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError, match="plan was built"):
        plan.apply_host(torch.rand(2, 10, 10, dtype=torch.float64))


def test_deferred_drain_on_a_persistent_pipeline(engine):
    """Several host calls queued on one pipeline without draining in between, one drain at the end."""
    lat, lon = np.arange(-90.0, 90.25, 0.5), np.arange(0.0, 360.0, 0.5)
    tlat, tlon = np.arange(-90.0, 91, 2.0), np.arange(0.0, 360, 2.0)
    plan = engine.ConservativePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"))
    rng = np.random.default_rng(4)
    xs = [torch.from_numpy(0.5 + rng.random((7, lat.size, lon.size))).pin_memory() for _ in range(3)]
    outs = [torch.empty((7, tlat.size, tlon.size), dtype=torch.float64).pin_memory() for _ in range(3)]
    pipe = engine.HostPipeline(slots=2)
    for x, o in zip(xs, outs):
        plan.apply_host(x, out_host=o, chunk_batch=3, pipeline=pipe, drain=False)
    pipe.drain()
    for x, o in zip(xs, outs):
        want = oc.regrid_latlon(x.numpy(), lat, lon, tlat, tlon)
        _check(o, want, 1e-12)
    pipe.close()


def test_accessor_numpy_route_equals_cuda_route():
    """The accessor streams numpy input through the host pipeline; same bits as CUDA-tensor input."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xarray_regrid_b200.labelled import DataArray, Dataset

    lat, lon = np.arange(-89.5, 90, 1.0), np.arange(0.5, 360, 1.0)
    tgt = Dataset({}, {"lat": np.arange(-88.0, 90, 4.0), "lon": np.arange(2.0, 360, 4.0)})
    rng = np.random.default_rng(8)
    x = 0.5 + rng.random((9, lat.size, lon.size))
    x[rng.random(x.shape) < 0.1] = np.nan
    coords = {"time": np.arange(9), "lat": lat, "lon": lon}
    host = DataArray(x, ("time", "lat", "lon"), coords)
    dev = DataArray(torch.from_numpy(x).cuda(), ("time", "lat", "lon"), coords)
    for call in (lambda d: d.regrid.conservative(tgt, nan_threshold=0.5), lambda d: d.regrid.linear(tgt),
                 lambda d: d.regrid.cubic(tgt), lambda d: d.regrid.stat(tgt, "mean", skipna=True)):
        a, b = call(host), call(dev)
        assert isinstance(a.data, np.ndarray) and b.data.is_cuda
        np.testing.assert_array_equal(a.data, b.data.cpu().numpy())
    lab = rng.integers(0, 5, x.shape).astype(np.int32)
    a = DataArray(lab, ("time", "lat", "lon"), coords).regrid.most_common(tgt, values=np.arange(5))
    b = DataArray(torch.from_numpy(lab).cuda(), ("time", "lat", "lon"), coords).regrid.most_common(
        tgt, values=np.arange(5))
    np.testing.assert_array_equal(a.data, b.data.cpu().numpy())


def test_c5_most_common_full_scale_18000x36000(engine):
    """BASELINE config 5 at its full size: 18000 x 36000 uint8 land cover, 32 classes -> 0.25 deg (720 x 1440).
    The oracle runs on 4 latitude bands of 10 target rows (incl. both polar edge bands); sharding.latitude_band
    gives the source rows a band needs."""
    from xarray_regrid_b200 import sharding

    lat, lon = np.round(np.arange(-89.995, 90, 0.01), 3), np.round(np.arange(0.005, 360, 0.01), 3)
    tlat, tlon = np.arange(-89.875, 90, 0.25), np.arange(0.125, 360, 0.25)
    g = torch.Generator(device="cuda").manual_seed(13)
    coarse = torch.randint(0, 32, (18000 // 64 + 1, 36000 // 64 + 1), device="cuda", generator=g)
    x = coarse.repeat_interleave(64, 0).repeat_interleave(64, 1)[:18000, :36000].to(torch.uint8).contiguous()
    noise = torch.rand((18000, 36000), device="cuda", generator=g) < 0.05
    x[noise] = torch.randint(0, 32, (int(noise.sum()),), device="cuda", generator=g, dtype=torch.uint8)
    del noise, coarse
    plan = engine.ReducePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"))
    values = np.arange(32)
    y = plan.mode(x, values)
    assert y.shape == (720, 1440) and y.dtype == torch.uint8
    for r in (0, 23, 48, 71):
        tgt_idx, src = sharding.latitude_band(lat, tlat, 72, r)
        want = orr.regrid_latlon(x[src].cpu().numpy(), lat[src], lon, tlat[tgt_idx], tlon, "most_common", values=values)
        np.testing.assert_array_equal(y[tgt_idx.min():tgt_idx.max() + 1].cpu().numpy(), want)
    # size-independent property: regridding a constant field returns the constant, and permuting the label
    # values permutes the result (ties aside: use least_common's complement only where counts are distinct)
    const = torch.full_like(x, 7)
    assert bool((plan.mode(const, values) == 7).all())


def test_threshold_ties_are_the_only_nan_placement_differences(engine):
    """On 2:1 coarsening with nan_threshold=0.5 the valid mass of a cell can equal the threshold EXACTLY in
    real arithmetic (e.g. the centre column, weight 1/2, missing in every row).  Which side of the comparison
    such a cell lands on depends on the summation order -- in the reference on the BLAS behind einsum -- so
    NaN placement is compared outside those ties, and the ties must be rare."""
    lat, lon = np.arange(-90.0, 90.25, 0.25), np.arange(0.0, 360.0, 0.25)
    tlat, tlon = np.arange(-90.0, 90.1, 0.5), np.arange(0.0, 360, 0.5)
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand((2, lat.size, lon.size), device="cuda", dtype=torch.float64, generator=g) + 0.5
    x[torch.rand(x.shape, device="cuda", generator=g) < 0.05] = float("nan")
    plan = engine.ConservativePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"))
    y = plan(x, skipna=True, nan_threshold=0.5).cpu().numpy()
    want, valid = oc.regrid_latlon(x.cpu().numpy(), lat, lon, tlat, tlon, skipna=True, nan_threshold=0.5,
                                   return_valid=True)
    ties = np.abs(valid - 0.5) <= 1e-13
    assert ties.sum() < 1e-3 * ties.size
    np.testing.assert_array_equal(np.isnan(y)[~ties], np.isnan(want)[~ties])
    both = ~np.isnan(y) & ~np.isnan(want)
    np.testing.assert_allclose(y[both], want[both], rtol=1e-12)


# ------------------------------------------------------------------- column-panel layouts of the two-phase kernel
PANEL_CASES = {
    # name: (src lat, src lon, tgt lat, tgt lon)
    "ratio2_centred": (np.arange(-90.0, 90.25, 0.5), np.arange(0.0, 360.0, 0.5),
                       np.arange(-90.0, 90.5, 1.0), np.arange(0.0, 360.0, 1.0)),
    "ratio2_edge_aligned_poles": (np.arange(-89.75, 90, 0.5), np.arange(0.25, 360, 0.5),
                                  np.arange(-90.0, 90.5, 1.0), np.arange(0.0, 360.0, 1.0)),
    "ratio2p5": (np.arange(-90.0, 90.5, 1.0), np.arange(0.0, 360.0, 1.0),
                 np.arange(-90.0, 90.1, 2.5), np.arange(0.0, 360.0, 2.5)),
    "ratio2_wide_two_panels": (np.arange(-45.0, 45.1, 0.25), np.arange(0.0, 360.0, 0.25),
                               np.arange(-45.0, 45.1, 0.5), np.arange(0.0, 360.0, 0.5)),
    "ragged_last_panel": (np.arange(-30.0, 30.1, 0.5), np.arange(0.0, 190.0, 0.5),   # regional: 7 strips -> 4 + 3
                          np.arange(-30.0, 30.1, 1.0), np.arange(0.0, 189.5, 1.0)),
    "descending_lat_shifted_lon": (np.arange(60.0, -60.1, -0.5), np.arange(-180.0, 180.0, 0.5),
                                   np.arange(-60.0, 60.1, 1.0), np.arange(0.0, 360.0, 1.0)),
}


PANEL_CASES["ratio2_odd_targets_regional"] = (np.arange(-30.0, 30.1, 0.5), np.arange(10.0, 130.0, 0.5),
                                              np.arange(-30.0, 30.1, 1.0), np.arange(10.0, 128.5, 1.0))  # 240 -> 119
_LON01 = np.round(np.arange(0, 359.95, 0.1), 10)
PANEL_CASES["ratio2p5_five_block_panels"] = (np.round(np.arange(-6.0, 6.05, 0.1), 10), _LON01,   # W = 3600
                                             np.arange(-6.0, 6.1, 0.25), np.arange(0.0, 360.0, 0.25))
PANEL_CASES["ratio3"] = (np.arange(-60.0, 60.1, 0.5), np.arange(0.0, 360.0, 0.5),
                         np.arange(-60.0, 60.1, 1.5), np.arange(0.0, 360.0, 1.5))
PANEL_CASES["ratio2p5_offset_regional"] = (np.arange(-30.0, 30.1, 1.0), np.arange(-20.0, 100.0, 1.0),
                                           np.arange(-28.0, 28.1, 2.5), np.arange(-15.0, 95.0, 2.5))


@pytest.mark.parametrize("case", sorted(PANEL_CASES))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("layout", ["default", "panels"])
def test_panel_layouts_match_oracle_and_generic_kernel(engine, case, dtype, layout):
    """Coarsening ratios of 2 - 2.5.  ``panels``: the two-phase ring kernel in its column-panel layout (several
    CTAs per SM, strips of up to 64 targets, panels that wrap the date line, panels with fewer strips than
    consumer warps, synthetic pole rows with weight).  ``default``: ratios 2 - 3 run the register-resident kernel in
    its block layout (a lane owns 4 source columns and up to two targets; halo lanes at both ends of a warp, one- and
    two-column halos, odd target counts, regional axes, the date line inside a warp, column panels of blocks for
    W = 3600).  Every NaN mode against the oracle and against the generic kernel."""
    lat, lon, tlat, tlon = PANEL_CASES[case]
    rng = np.random.default_rng(77)
    x = (0.5 + rng.random((5, lat.size, lon.size))).astype(dtype)
    x[rng.random(x.shape) < 0.06] = np.nan
    x[1, 10:14, :] = np.nan        # whole target rows missing
    x[2, :, 37:45] = np.nan        # whole target columns missing
    x[3] = 2.5                     # NaN-free slab
    spherical = case != "ratio2_edge_aligned_poles"  # keep a weight on the synthetic pole rows in that case
    plan = engine.ConservativePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"),
                                   spherical_rows=spherical)
    plan.use_lanes = layout == "default"
    xt = torch.from_numpy(x).cuda()
    ring = plan._ring(5, lat.size, lon.size, xt.dtype)
    assert ring is not None
    if layout == "default":
        assert ring.host.lane_cols == 2 and ring.host.n_warps <= 12, "block layout expected"
        assert int(np.diff(ring.host.strip_l0).max()) <= 60 and int(np.diff(ring.host.strip_b0).max()) <= 30
        assert (ring.host.n_panels > 1) == (case == "ratio2p5_five_block_panels")
        if case.startswith("ratio2p5") or case == "ratio3":
            assert max(ring.host.halo_l, ring.host.halo_r) == 2 and np.diff(ring.host.block_l0).min() < 2
        # the interpolation (NaN-propagating) mode never takes that layout
        assert plan._ring(5, lat.size, lon.size, xt.dtype, pairs=False).host.lane_cols == 0
    else:
        assert ring.host.lane_cols == 0 and ring.host.n_warps <= 6, "panel layout expected"
        assert int(np.diff(ring.host.strip_l0).max()) <= 64
        if case == "ragged_last_panel" and dtype == np.float64:  # (the fp32 strip decomposition differs)
            assert np.diff(ring.host.panel_strip0).min() < ring.host.n_warps
    if case == "ratio2_edge_aligned_poles":
        assert plan.pole_rows
    rtol = 1e-12 if dtype == np.float64 else 2e-5
    for skipna, thr in [(True, 0.3), (True, 1.0), (False, 1.0)]:
        want, valid = oc.regrid_latlon(x, lat, lon, tlat, tlon, skipna=skipna, nan_threshold=thr,
                                       latitude_weighting=spherical, return_valid=True)
        ties = np.abs(valid - oc.valid_threshold(thr)) <= (1e-13 if dtype == np.float64 else 1e-6)
        got = {}
        for use_ring in (True, False):
            plan.use_ring = use_ring
            got[use_ring] = plan(xt, skipna=skipna, nan_threshold=thr).cpu().numpy()
            y = got[use_ring]
            np.testing.assert_array_equal(np.isnan(y)[~ties], np.isnan(want)[~ties])
            both = ~np.isnan(y) & ~np.isnan(want)
            np.testing.assert_allclose(y[both], want[both], rtol=rtol, atol=rtol)
        plan.use_ring = True
    # the host pipeline drives the same kernels chunk by chunk
    y_host = plan.apply_host(torch.from_numpy(x).pin_memory(), skipna=True, nan_threshold=0.3, chunk_batch=2)
    np.testing.assert_array_equal(y_host.numpy(), plan(xt, skipna=True, nan_threshold=0.3).cpu().numpy())


def test_interp_on_a_ratio2_grid_keeps_nan_propagation(engine):
    """Linear interpolation onto a grid twice as coarse shares the bands' shape with the conservative case but must
    not take the two-targets-per-lane kernel (its zero-weight taps would spread a NaN to the neighbouring target)."""
    lat, lon = np.arange(-30.0, 30.1, 0.5), np.arange(0.0, 360.0, 0.5)
    tlat, tlon = np.arange(-29.75, 30.0, 1.0), np.arange(0.25, 359.0, 1.0)
    rng = np.random.default_rng(5)
    x = rng.random((3, lat.size, lon.size))
    x[rng.random(x.shape) < 0.02] = np.nan
    plan = engine.InterpPlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"), "linear")
    y = plan(torch.from_numpy(x).cuda()).cpu().numpy()
    want = oi.regrid_latlon(x, lat, lon, tlat, tlon, "linear")
    assert np.isnan(want).any()
    np.testing.assert_array_equal(np.isnan(y), np.isnan(want))
    np.testing.assert_allclose(y[~np.isnan(y)], want[~np.isnan(want)], rtol=1e-12)


def test_c2_full_stack_8760_steps(engine):
    """BASELINE config 2 at its full size on one GPU: 8760 x 721 x 1440 fp64 (72.8 GB), 10 % NaN, skipna,
    nan_threshold 0.5.  SURVEY 8(d): the oracle checks steps {0, 1, 4379, 8759}; a size-independent property covers
    the rest: the regrid of a time-permuted stack is the permuted regrid (every slab is independent)."""
    from xarray_regrid_b200 import synthetic

    free, _ = torch.cuda.mem_get_info()
    if free < 90 * 2**30:
        pytest.skip("needs ~80 GB of free HBM")
    lat, lon = synthetic.grid_025()
    tlat, tlon = synthetic.grid_1()
    plan = engine.ConservativePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"))
    x = synthetic.conservative_stack(8760, lat, lon, nan_fraction=0.10, seed=1234)
    y = plan(x, skipna=True, nan_threshold=0.5)
    assert y.shape == (8760, 181, 360)
    steps = [0, 1, 4379, 8759]
    idx = torch.tensor(steps, device="cuda")
    want, valid = oc.regrid_latlon(x.index_select(0, idx).cpu().numpy(), lat, lon, tlat, tlon, skipna=True,
                                   nan_threshold=0.5, return_valid=True)
    got = y.index_select(0, idx).cpu().numpy()
    ties = np.abs(valid - 0.5) <= 1e-13
    assert ties.sum() < 1e-3 * ties.size
    np.testing.assert_array_equal(np.isnan(got)[~ties], np.isnan(want)[~ties])
    both = ~np.isnan(got) & ~np.isnan(want)
    np.testing.assert_allclose(got[both], want[both], rtol=1e-12)
    # every slab is independent: slabs regridded alone (other batch size, other task decomposition) give the same bits
    probe = torch.tensor([7, 2048, 6001, 8758], device="cuda")
    alone = plan(x.index_select(0, probe), skipna=True, nan_threshold=0.5)
    assert torch.equal(torch.nan_to_num(alone, nan=-1.0), torch.nan_to_num(y.index_select(0, probe), nan=-1.0))
    # NaN fraction of the result is plausible for a 10 % masked input with a 0.5 threshold
    frac = float(torch.isnan(y).float().mean())
    assert 0.02 < frac < 0.2


# --------------------------------------------------------------- more than two regridded coordinates at once
def test_conservative_over_three_coordinates_joint_valid_mass():
    """The reference contracts ALL shared coordinates in one `xr.dot` (methods/conservative.py:188-198), so the
    valid mass is joint over (level, latitude, longitude).  The engine folds level x latitude into the kernels'
    rows axis (tables.kron_band)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xarray_regrid_b200.labelled import DataArray, Dataset

    rng = np.random.default_rng(12)
    lev, tlev = np.arange(0.0, 10.0), np.array([0.5, 2.5, 4.5, 6.5, 8.5])
    lat, tlat = np.arange(-30.0, 30.1, 1.0), np.arange(-28.0, 28.1, 4.0)
    lon, tlon = np.arange(0.0, 360.0, 2.0), np.arange(0.0, 360.0, 8.0)
    x = 0.5 + rng.random((3, lev.size, lat.size, lon.size))
    x[rng.random(x.shape) < 0.15] = np.nan
    x[1, 2:4, 10:20, 30:60] = np.nan
    da = DataArray(x, ("time", "level", "latitude", "longitude"),
                   {"time": np.arange(3), "level": lev, "latitude": lat, "longitude": lon})
    tgt = Dataset({}, {"level": tlev, "latitude": tlat, "longitude": tlon})
    xf, latf, lonf = of.format_for_regrid(x, lat, lon, tlat, tlon)
    for skipna, thr in [(True, 0.4), (True, 1.0), (False, 1.0)]:
        out = da.regrid.conservative(tgt, skipna=skipna, nan_threshold=thr)
        assert out.dims == da.dims and out.shape == (3, tlev.size, tlat.size, tlon.size)
        want, valid = oc.regrid(xf, [lev, latf, lonf], [tlev, tlat, tlon], [-3, -2, -1], latitude_index=1,
                                skipna=skipna, nan_threshold=thr, return_valid=True)
        ties = np.abs(valid - oc.valid_threshold(thr)) <= 1e-13
        np.testing.assert_array_equal(np.isnan(out.data)[~ties], np.isnan(want)[~ties])
        both = ~np.isnan(out.data) & ~np.isnan(want)
        np.testing.assert_allclose(out.data[both], want[both], rtol=1e-12)
    # dims in another order, CUDA input
    xt = torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 3, 1))).cuda()
    da2 = DataArray(xt, ("latitude", "time", "longitude", "level"), dict(da.coords))
    out2 = da2.regrid.conservative(tgt, skipna=True, nan_threshold=0.4)
    ref = da.regrid.conservative(tgt, skipna=True, nan_threshold=0.4).data
    got2 = out2.data.cpu().numpy()  # (the folding order follows the dim order: same taps, another summation order)
    np.testing.assert_array_equal(np.isnan(got2), np.isnan(ref.transpose(2, 0, 3, 1)))
    np.testing.assert_allclose(got2, ref.transpose(2, 0, 3, 1), rtol=1e-13)


@pytest.mark.parametrize("method", ["linear", "nearest", "cubic"])
def test_interp_over_three_coordinates(method):
    """xr.interp interpolates coordinate after coordinate: three shared coordinates = the fused (lat, lon) pass,
    then the third axis."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xarray_regrid_b200.labelled import DataArray, Dataset

    rng = np.random.default_rng(13)
    lev, tlev = np.arange(0.0, 12.0), np.linspace(0.5, 10.5, 9)
    lat, tlat = np.arange(-30.0, 30.1, 2.0), np.arange(-29.0, 29.1, 0.5)
    lon, tlon = np.arange(0.0, 360.0, 4.0), np.arange(0.0, 358.0, 1.5)
    x = rng.random((2, lev.size, lat.size, lon.size))
    da = DataArray(x, ("time", "level", "latitude", "longitude"),
                   {"time": np.arange(2), "level": lev, "latitude": lat, "longitude": lon})
    tgt = Dataset({}, {"level": tlev, "latitude": tlat, "longitude": tlon})
    out = getattr(da.regrid, method)(tgt)
    xf, latf, lonf = of.format_for_regrid(x, lat, lon, tlat, tlon)
    want = oi.regrid(xf, [latf, lonf, lev], [tlat, tlon, tlev], [-2, -1, -3], method)
    assert out.shape == want.shape
    np.testing.assert_array_equal(np.isnan(out.data), np.isnan(want))
    np.testing.assert_allclose(out.data, want, rtol=0 if method == "nearest" else 1e-10, atol=1e-12)


def test_group_by_over_three_coordinates():
    """flox bins any number of coordinates (methods/flox_reduce.py:89-96, 174-183).  The engine walks the bins of the
    coordinate that is neither latitude nor longitude on the host and merges its slices into the kernels' rows axis,
    so a (level, lat, lon) bin is reduced as one 2-D bin: against the oracle's N-D group-by, dims in two orders,
    ascending and descending third coordinate, numpy and CUDA input."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xarray_regrid_b200.labelled import DataArray, Dataset

    rng = np.random.default_rng(31)
    lev, tlev = np.arange(0.0, 12.0, 1.0), np.arange(1.0, 12.0, 3.0)            # bins of 3 levels
    lat, tlat = np.arange(-10.0, 10.25, 0.25), np.arange(-9.5, 10.0, 1.0)
    lon, tlon = np.arange(0.0, 360.0, 1.0), np.arange(0.0, 360.0, 4.0)
    x = rng.standard_normal((2, lev.size, lat.size, lon.size))
    x[rng.random(x.shape) < 0.03] = np.nan
    da = DataArray(x, ("time", "level", "latitude", "longitude"),
                   {"time": np.arange(2), "level": lev, "latitude": lat, "longitude": lon})
    tgt = Dataset({}, {"level": tlev, "latitude": tlat, "longitude": tlon})
    xf, latf, lonf = of.format_for_regrid(x, lat, lon, tlat, tlon, stats=True)
    for method, skipna in [("mean", True), ("var", True), ("max", False), ("median", True), ("sum", True)]:
        out = da.regrid.stat(tgt, method, skipna=skipna)
        want = orr.stat(xf, [lev, latf, lonf], [tlev, tlat, tlon], [-3, -2, -1], method, skipna=skipna)
        assert out.dims == da.dims and out.shape == want.shape
        np.testing.assert_array_equal(np.isnan(out.data), np.isnan(want))
        np.testing.assert_allclose(out.data, want, rtol=1e-12, atol=1e-13)
    # labels; descending third coordinate; dims in another order; CUDA input
    labels = rng.integers(0, 6, (lat.size, lev.size, 2, lon.size)).astype(np.int32)
    lev_d = lev[::-1].copy()
    dal = DataArray(torch.from_numpy(labels).cuda(), ("latitude", "level", "time", "longitude"),
                    {"time": np.arange(2), "level": lev_d, "latitude": lat, "longitude": lon})
    lf, latf2, lonf2 = of.format_for_regrid(labels.transpose(2, 1, 0, 3), lat, lon, tlat, tlon, stats=True)
    for op in ("most_common", "least_common"):
        out = getattr(dal.regrid, op)(tgt, values=np.arange(6))
        want = orr.mode(lf, [lev_d, latf2, lonf2], [tlev, tlat, tlon], [-3, -2, -1], values=np.arange(6),
                        anti_mode=(op == "least_common"))
        got = out.data.cpu().numpy().transpose(2, 1, 0, 3)
        assert out.dims == dal.dims and got.shape == want.shape
        np.testing.assert_array_equal(got, want)
    # a third coordinate that leaves a target bin empty is refused, not guessed
    with pytest.raises(NotImplementedError):
        da.regrid.stat(Dataset({}, {"level": np.arange(-20.0, 12.0, 3.0), "latitude": tlat, "longitude": tlon}), "mean")
