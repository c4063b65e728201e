"""Host-side tables (product code) against the oracle / the reference's golden weights."""

import numpy as np
import pytest

from oracle import conservative as oc
from oracle import formatting as of
from xarray_regrid_b200 import tables

AXIS_CASES = [
    "c1_lat", "c1_lon", "era5_to_2p2_lat", "cellcentred_lat", "upsample",
    "irregular", "single_target", "tgt_outside",
]


@pytest.mark.parametrize("name", AXIS_CASES)
@pytest.mark.parametrize("spherical", [False, True])
def test_band_equals_reference_dense_weights(golden_helpers, name, spherical):
    g = golden_helpers
    key = f"axis/{name}/weights_spherical" if spherical else f"axis/{name}/weights"
    if key not in g:
        pytest.skip("no spherical variant")
    src, tgt = g[f"axis/{name}/src"], g[f"axis/{name}/tgt"]
    band = tables.conservative_band(tables.plain_axis(src), tgt, spherical=spherical)
    dense = band.dense()
    ref = g[key]
    # entry for entry; column sums are taken over the band only, so allow the last bit
    np.testing.assert_allclose(dense, ref, rtol=4e-16, atol=0)
    assert ((dense != 0) == (ref != 0)).all(), "band must hold exactly the non-zero taps"
    assert band.k_max == max(int((ref != 0).sum(axis=0).max()), 1)
    np.testing.assert_array_equal(band.cover.astype(bool), oc.covered_mask(src, tgt))


def test_c1_band_shape():
    lat = np.arange(-90, 90.25, 0.25)
    lon = np.arange(0, 360, 0.25)
    tlat = np.arange(-90.0, 91.0, 1.0)
    tlon = np.arange(0.0, 360.0, 1.0)
    rows = tables.conservative_band(tables.format_lat_axis(lat), tlat, spherical=True)
    cols = tables.conservative_band(tables.format_lon_axis(lon, tlon), tlon)
    assert (rows.n_out, rows.k_max, cols.n_out, cols.k_max) == (181, 5, 360, 5)
    assert rows.cnt[0] == 2 and rows.cnt[-1] == 2  # zero-area pole rows dropped
    assert cols.idx.min() == 0 and cols.idx.max() == 1439  # wrap folded into the indices
    np.testing.assert_array_equal(cols.idx[: cols.cnt[0], 0], [1438, 1439, 0, 1, 2])
    assert not rows.meta["pole_rows"]


@pytest.mark.parametrize(
    "src,tgt",
    [
        (np.arange(1.0, 360, 2), np.arange(0.0, 361, 1)),       # test_no_edges
        (np.arange(0.0, 361, 2), np.arange(-180.0, 181, 1)),    # 360 -> 180
        (np.arange(-180.0, 181, 2), np.arange(0.0, 361, 1)),    # 180 -> 360
        (np.arange(-360.0, 1, 2), np.arange(0.0, 361, 1)),      # 0 -> 360
        (np.arange(-180.0, 181, 2), np.arange(270.0, 301, 1)),  # global -> local
        (np.arange(0.0, 360, 0.25), np.arange(0.0, 360, 1.0)),  # C1
        (np.arange(359.75, -0.1, -0.25), np.arange(0.0, 360, 1.0)),  # descending source
    ],
)
def test_lon_formatting_matches_oracle(src, tgt):
    ax = tables.format_lon_axis(src, tgt)
    data = np.arange(src.size, dtype=np.float64)[None, :]
    out, _, lon = of.format_for_regrid(data, None, src, None, tgt)
    np.testing.assert_array_equal(ax.values, lon)
    np.testing.assert_array_equal(ax.source, out[0].astype(int))


def test_lat_formatting_matches_oracle():
    lat = np.arange(-89.5, 90, 1.0)
    ax = tables.format_lat_axis(lat)
    assert ax.values[0] == -90 and ax.values[-1] == 90 and ax.has_pole_rows
    np.testing.assert_array_equal(ax.source, np.concatenate([[180], np.arange(180), [181]]))
    ax = tables.format_lat_axis(np.arange(-90.0, 91, 1.0))
    assert ax.is_identity
    ax = tables.format_lat_axis(np.arange(-88.0, 89, 1.0))
    assert ax.is_identity
    ax = tables.format_lat_axis(np.arange(90.0, -91, -1.0))  # descending: sorted via indices
    np.testing.assert_array_equal(ax.source, np.arange(180, -1, -1))


def test_reversed_target_is_folded_into_band():
    src = np.arange(0.0, 10.0, 1.0)
    tgt = np.arange(0.5, 9.0, 2.0)
    fwd = tables.conservative_band(tables.plain_axis(src), tgt)
    rev = tables.conservative_band(tables.plain_axis(src), tgt[::-1])
    np.testing.assert_array_equal(rev.dense(), fwd.dense()[:, ::-1])


def test_valid_threshold(golden_helpers):
    g = golden_helpers
    got = np.array([tables.valid_threshold(t) for t in g["threshold/in"]])
    np.testing.assert_array_equal(got, g["threshold/out"])


@pytest.mark.parametrize("n_runs", [1, 3, 50, 181])
@pytest.mark.parametrize("descending", [False, True])
def test_row_ring_schedule_invariants(n_runs, descending):
    """What the TMA row-ring kernel relies on: every tap of an output is in the ring (loaded,
    not yet released) when the output is computed, rows are released in FIFO order, and the
    number of simultaneously live rows never exceeds max_live."""
    lat = np.arange(-90, 90.25, 0.25)
    tlat = np.arange(-90.0, 91.0, 1.0)
    if descending:
        lat, tlat = lat[::-1].copy(), tlat[::-1].copy()
    rows = tables.conservative_band(tables.format_lat_axis(lat), tlat, spherical=True)
    ring = tables.row_ring_schedule(rows, n_runs)
    assert ring.n_runs == min(n_runs, 181)
    assert ring.run_first[0] == 0 and ring.run_first[-1] == rows.n_out
    assert ring.run_sched[-1] == ring.sched_row.size
    for r in range(ring.n_runs):
        k0, k1 = ring.run_first[r], ring.run_first[r + 1]
        s0, s1 = ring.run_sched[r], ring.run_sched[r + 1]
        prev_release = 0
        for k in range(k0, k1):
            assert prev_release <= ring.release[k] <= ring.need[k] <= s1 - s0
            assert ring.need[k] - prev_release <= ring.max_live
            assert ring.kcnt[k] == rows.cnt[k]
            for t in range(rows.cnt[k]):
                assert 0 <= ring.tap_back[t, k] < ring.max_live
                p = ring.need[k] - 1 - ring.tap_back[t, k]
                assert prev_release <= p < ring.need[k]
                assert ring.sched_row[s0 + p] == rows.idx[t, k]
            prev_release = ring.release[k]
        assert prev_release == s1 - s0, "every row must be released at the end of its run"
    if n_runs == 1:
        assert ring.sched_row.size == 719  # every source row with non-zero weight, exactly once


def test_row_ring_schedule_schedules_pole_rows_like_any_row():
    """The synthetic pole rows (indices n_src, n_src + 1) are loaded from the [batch, 2, W] pole buffer."""
    # (with the spherical correction the pole rows have zero area and are dropped from conservative bands;
    # they carry weight for the interpolation methods and for conservative regridding of a plain axis)
    rows = tables.linear_band(tables.format_lat_axis(np.arange(-89.0, 90, 2)), np.arange(-90.0, 91, 1))
    assert (rows.idx >= rows.n_src).any()
    ring = tables.row_ring_schedule(rows, 3)
    assert ring is not None
    virtual = ring.sched_row[ring.sched_row >= rows.n_src]
    assert set(virtual.tolist()) <= {rows.n_src, rows.n_src + 1} and virtual.size >= 1
    assert ring.sched_row.max() <= rows.n_src + 1


# ------------------------------------------------------------------ interpolation / bin tables (CPU)
from oracle import interp as oi  # noqa: E402
from oracle import reduce as orr  # noqa: E402


def _apply_band(band, y, n_src=None):
    """Dense application of a band along the last axis, NaN poisoning through every tap like the kernels."""
    n_src = band.n_src if n_src is None else n_src
    out = np.zeros(y.shape[:-1] + (band.n_out,))
    for o in range(band.n_out):
        acc = np.zeros(y.shape[:-1])
        for t in range(band.cnt[o]):
            acc = acc + band.w[t, o] * y[..., band.idx[t, o]]
        out[..., o] = acc if band.cover[o] else np.nan
    return out


@pytest.mark.parametrize("method", ["linear", "nearest", "cubic"])
@pytest.mark.parametrize(
    "src,tgt",
    [
        (np.arange(0.0, 40.0, 1.0), np.linspace(-1, 41, 97)),          # upsampling, targets outside
        (np.arange(0.0, 40.0, 0.25), np.arange(0.5, 39.0, 1.7)),        # downsampling
        (np.sort(np.random.default_rng(3).uniform(0, 50, 31)), np.linspace(1, 49, 40)),  # irregular source
        (np.arange(10.0, 0.0, -0.5), np.arange(9.5, 0.4, -0.75)),       # descending source and target
    ],
)
def test_interp_bands_match_scipy(method, src, tgt):
    rng = np.random.default_rng(0)
    y = rng.random((3, src.size))
    if method == "linear":
        y[1, 4] = np.nan
    ax = tables.plain_axis(src)
    want = oi.regrid(y, [src], [tgt], [-1], method)
    if method == "cubic":
        pre, ev = tables.cubic_bands(ax, tgt)
        got = _apply_band(ev, _apply_band(pre, y), n_src=ax.values.size)
    else:
        band = tables.linear_band(ax, tgt) if method == "linear" else tables.nearest_band(ax, tgt)
        got = _apply_band(band, y)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)


def test_nearest_ties_go_to_lower_neighbour():
    ax = tables.plain_axis(np.arange(5.0))
    idx, cover = tables.nearest_index(ax, np.array([0.5, 1.5, 2.5, 3.5, -0.1, 4.0, 4.1]))
    np.testing.assert_array_equal(idx, [0, 1, 2, 3, 0, 4, 4])
    np.testing.assert_array_equal(cover, [1, 1, 1, 1, 0, 1, 0])


@pytest.mark.parametrize(
    "src,tgt",
    [
        (np.linspace(0, 40, 11), np.arange(0.0, 41.0, 8.0)),            # reference tests/test_reduce.py grids
        (np.linspace(0, 40, 11), np.arange(-8.0, 49.0, 8.0)),           # oversized target
        (np.round(np.arange(-9.995, 10, 0.01), 3), np.arange(-10.0, 10.01, 0.25)),  # edges ON source centres
        (np.arange(20.0, 0.0, -0.5), np.arange(18.0, 1.0, -2.0)),       # descending source and target
    ],
)
def test_bins_match_oracle_membership(src, tgt):
    """Bin tables against the oracle's digitize-based membership, coverage and target order."""
    bins = tables.bins_axis(tables.plain_axis(src), tgt)
    data = np.arange(src.size, dtype=np.float64)  # value = source index
    want = orr.stat(data, [src], [tgt], [0], "sum")
    want_n = orr.stat(np.ones(src.size), [src], [tgt], [0], "sum")
    got = np.full(tgt.size, np.nan)
    got_n = np.full(tgt.size, np.nan)
    for o in range(bins.n_out):
        pos = bins.start[o] + np.arange(bins.count[o])
        idx = (bins.src_offset + bins.src_step * pos) % bins.n_src
        if bins.cover[o]:
            got[bins.out_index[o]] = data[idx].sum()
            got_n[bins.out_index[o]] = idx.size
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_array_equal(np.nan_to_num(got), np.nan_to_num(want))
    np.testing.assert_array_equal(np.nan_to_num(got_n), np.nan_to_num(want_n))


def test_bins_with_longitude_wrap_padding():
    """stats=True formatting keeps the wrap columns: a pole-centred 2 degree target over a cell-centred
    1 degree source gets the wrapped column in its first bin (tests/test_format.py:182-218)."""
    lon = np.arange(0.5, 360, 1.0)
    tlon = np.arange(0.0, 361, 2.0)
    ax = tables.format_lon_axis(lon, tlon)
    assert ax.values[0] == -1.5 and ax.values[-1] == 361.5
    bins = tables.bins_axis(ax, tlon)
    assert bins.gather is None and bins.src_step == 1
    first = (bins.src_offset + bins.start[0] + np.arange(bins.count[0])) % bins.n_src
    np.testing.assert_array_equal(first, [359, 0])  # [-1, 1): the wrapped -0.5 and 0.5
    assert bins.cover.all()


def test_plan_chunks_respects_shared_memory_budget():
    rows = tables.bins_axis(tables.plain_axis(np.round(np.arange(-89.995, 90, 0.01), 3)), np.arange(-89.875, 90, 0.25))
    cols = tables.bins_axis(tables.plain_axis(np.round(np.arange(0.005, 360, 0.01), 3)), np.arange(0.125, 360, 0.25))
    for esize in (1, 4, 8):
        cpc, tile = tables.plan_chunks(rows, cols, esize)
        assert tile * esize <= 72 * 1024 and tile >= 25 * 25 * min(cpc, 1)
    with pytest.raises(NotImplementedError):
        big = tables.Bins(np.array([0], np.int32), np.array([2000], np.int32), np.ones(1, np.uint8),
                          np.zeros(1, np.int32), 2000, 0, 1)
        tables.plan_chunks(big, big, 8)


def test_lane_owned_columns_eligibility():
    """rg_row_ring_t.lane_cols: only regular 4:1 coarsening with full coverage qualifies."""
    from xarray_regrid_b200 import engine

    def ring_for(lat, lon, tlat, tlon, vec=2):
        cols = tables.conservative_band(engine._axis(engine.AxisSpec(lon, tlon, "lon"), False), tlon)
        rows = tables.conservative_band(engine._axis(engine.AxisSpec(lat, tlat, "lat"), True), tlat,
                                        spherical=True)
        return tables.row_ring_schedule(rows, 4, cols, vec), cols

    lat025, lon025 = np.arange(-90, 90.25, 0.25), np.arange(0, 360, 0.25)
    lat1, lon1 = np.arange(-90, 91, 1.0), np.arange(0, 360, 1.0)
    for vec in (2, 4):  # fp64 / fp32 strip alignment
        ring, cols = ring_for(lat025, lon025, lat1, lon1, vec)
        assert ring.lane_cols == 4 and ring.n_strips <= ring.n_warps <= 12
        assert all(b - a <= 31 for a, b in zip(ring.strip_l0[:-1], ring.strip_l0[1:]))
        first = cols.idx[0].astype(np.int64)
        assert (first % 2 == 0).all() and ((np.diff(first) - 4) % cols.n_src == 0).all()
    # coarsening by 2 - 3 takes the block layout: a lane owns 4 source columns and <= 2 consecutive targets whose
    # taps all lie within halo_l / halo_r <= 2 columns of the block; <= 30 blocks per warp; panels of blocks for
    # rows wider than 12 warps
    lat01, lon01 = np.round(np.arange(-9, 9.05, 0.1), 10), np.round(np.arange(0, 359.95, 0.1), 10)
    block_cases = [(lat1, lon1, np.arange(-90, 91, 2.0), t) for t in
                   (np.arange(0, 360, 2.0), np.arange(0.5, 360, 2.0), np.arange(0, 357, 2.0), np.arange(0, 360, 2.5),
                    np.arange(0, 360, 3.0), np.arange(7.0, 300, 2.5))]
    block_cases.append((lat01, lon01, np.arange(-9, 9.1, 0.25), np.arange(0, 360, 0.25)))
    for vec in (2, 4):
        for la, lo, tla, tlo in block_cases:
            ring, cols = ring_for(la, lo, tla, tlo, vec)
            W = cols.n_src
            assert ring.lane_cols == 2 and ring.n_warps <= 12
            assert 1 <= ring.halo_l <= 2 and 1 <= ring.halo_r <= 2
            assert (ring.strip_c0 % (4 if vec == 4 and ring.n_panels > 1 else 2) == 0).all()
            assert (np.diff(ring.strip_l0) <= 60).all() and (np.diff(ring.strip_b0) <= 30).all()
            assert ring.strip_l0[0] == 0 and ring.strip_l0[-1] == cols.n_out == ring.block_l0[-1]
            assert (np.diff(ring.block_l0) >= 0).all() and (np.diff(ring.block_l0) <= 2).all()
            assert (np.diff(ring.panel_strip0) <= ring.n_warps).all() and ring.panel_strip0[-1] == ring.n_strips
            assert (ring.n_panels > 1) == (W == 3600)
            for p in range(ring.n_panels):
                for s in range(ring.panel_strip0[p], ring.panel_strip0[p + 1]):
                    for j in range(ring.strip_b0[s + 1] - ring.strip_b0[s]):
                        c0 = int(ring.strip_c0[s]) + 4 * j
                        if ring.n_panels > 1:  # the block and both halo blocks lie inside the panel's ring row
                            rel = (c0 - int(ring.panel_c0[p])) % W
                            assert rel >= 4 and rel + 8 <= ring.panel_nc[p] and ring.panel_nc[p] % 4 == 0
                        b = ring.strip_b0[s] + j
                        for l in range(ring.block_l0[b], ring.block_l0[b + 1]):
                            q = (cols.idx[: cols.cnt[l], l].astype(np.int64) - (c0 - ring.halo_l)) % W
                            assert (q < 4 + ring.halo_l + ring.halo_r).all(), (s, l, q)
            # ... but never for the NaN-propagating interpolation mode
            rows = tables.conservative_band(engine._axis(engine.AxisSpec(la, tla, "lat"), True), tla, spherical=True)
            assert tables.row_ring_schedule(rows, 4, cols, vec, allow_pairs=False).lane_cols == 0
    # a regional target with uncovered edge cells and a non-integer ratio do not qualify for either
    ring, _ = ring_for(lat025[200:400], lon025[100:700], lat1[40:110], lon1[20:180])
    assert ring is None or ring.lane_cols == 0
    ring, _ = ring_for(lat025, lon025, np.arange(-90, 90.1, 0.7), np.arange(0, 360, 0.7))
    assert ring.lane_cols == 2  # (ratio 2.8: blocks)
    ring, _ = ring_for(lat025, lon025, np.arange(-90, 90.1, 0.4), np.arange(0, 360, 0.4))
    assert ring is None or ring.lane_cols == 0  # ratio 1.6: more than two targets per block of 4 columns


def test_band_window_rows_is_exact():
    """rg_band_t.window_rows: regular upsampling gives 32, ratio ~2 a smaller multiple of 4, clustered
    source points or shuffled targets 0 (the upsampling kernel then stays out of the way)."""
    ax = tables.plain_axis(np.arange(-90.0, 91.0, 1.0))
    assert tables.band_window_rows(tables.linear_band(ax, np.round(np.arange(-90, 90.05, 0.1), 10))) == 32
    _, ev = tables.cubic_bands(ax, np.round(np.arange(-90, 90.05, 0.1), 10))
    assert tables.band_window_rows(ev) == 32
    half = tables.band_window_rows(tables.linear_band(ax, np.arange(-90, 90.01, 0.5)))
    assert 0 < half < 32 and half % 4 == 0
    rng = np.random.default_rng(3)
    tgt = np.round(np.arange(-90, 90.05, 0.1), 10)
    assert tables.band_window_rows(tables.linear_band(ax, rng.permutation(tgt))) == 0
    clustered = tables.plain_axis(np.sort(np.r_[np.linspace(0, 1, 60), np.linspace(1.001, 100, 40)]))
    band = tables.linear_band(clustered, np.linspace(0, 100, 400))
    rows = tables.band_window_rows(band)
    # the definition, checked by brute force
    for cand in range(32, 0, -4):
        fits = all(band.idx[:, k0:k0 + cand].max() - band.idx[:, k0:k0 + cand].min() + 1 <= 16
                   for k0 in range(0, band.n_out, cand))
        if fits:
            break
    else:
        cand = 0
    assert rows == cand


def test_band_group_span():
    """rg_band_t.group_span: contiguous-tap bands report the widest 4-output window, others 0."""
    ax = tables.plain_axis(np.arange(0.0, 360.0, 1.0))
    fine = np.round(np.arange(0, 358.95, 0.1), 10)
    _, ev = tables.cubic_bands(ax, fine)
    assert tables.band_group_span(ev) == 5          # 4 taps, 4 outputs 0.1 apart cross at most one node
    assert tables.band_group_span(tables.linear_band(ax, fine)) == 3
    _, ev2 = tables.cubic_bands(ax, np.arange(1.0, 350.0, 0.5))
    assert tables.band_group_span(ev2) == 5         # ratio 2: 4 outputs 0.5 apart still cross one node only
    cons = tables.conservative_band(tables.format_lon_axis(np.arange(0, 360, 0.25), np.arange(0, 360, 1.0)),
                                    np.arange(0, 360, 1.0))
    assert tables.band_group_span(cons) == 0        # wrap-around taps are not index-contiguous
    # brute force on an irregular axis
    rng = np.random.default_rng(4)
    irr = tables.plain_axis(np.sort(rng.uniform(0, 100, 80)))
    _, ev3 = tables.cubic_bands(irr, np.linspace(irr.values[0], irr.values[-1], 333))
    want = max(ev3.idx[:, a:a + 4].max() - ev3.idx[:, a:a + 4].min() + 1 for a in range(0, ev3.n_out, 4))
    assert tables.band_group_span(ev3) == want


@pytest.mark.parametrize("order", ["ascending", "descending", "shuffled"])
@pytest.mark.parametrize("lon_mean", [True, False])
def test_pole_rows_come_from_the_sorted_edge_rows(order, lon_mean):
    """format_lat pads the poles AFTER sorting (utils.py:277, 313-336): the south pole row is the longitude
    mean (or, without a lon coordinate, a plain copy) of the SOUTHERNMOST row wherever it sits in the
    caller's array.  Emulates the kernel's read of the tables in numpy against the oracle."""
    lat = np.arange(-89.0, 90.0, 2.0)
    if order == "descending":
        lat = lat[::-1].copy()
    elif order == "shuffled":
        lat = np.random.default_rng(5).permutation(lat)
    tlat = np.arange(-90.0, 90.5, 1.5)
    rng = np.random.default_rng(1)
    x = lat[:, None] + rng.random((lat.size, 12))  # [lat, lon]
    ax = tables.format_lat_axis(lat, lon_mean=lon_mean)
    assert ax.pole_src == (int(np.argmin(lat)), int(np.argmax(lat)))
    assert ax.has_pole_rows == lon_mean
    band = tables.linear_band(ax, tlat)
    # what the kernels read: original rows, plus (lon_mean) the two pole rows appended at n_src, n_src + 1
    rows = x
    if lon_mean:
        pole = np.stack([np.broadcast_to(x[ax.pole_src[0]].mean(), (12,)), np.broadcast_to(x[ax.pole_src[1]].mean(), (12,))])
        rows = np.concatenate([x, pole])
    got = _apply_band(band, rows.T, n_src=rows.shape[0]).T
    data, flat, _ = of.format_for_regrid(x, lat, np.arange(12.0) if lon_mean else None, tlat,
                                         np.arange(12.0) if lon_mean else None)
    if lon_mean:  # undo the (irrelevant here) longitude formatting: compare on the latitude axis only
        data, flat = of.format_lat(*of.ensure_monotonic(x, lat, 0), 0, 1)
    want = oi.regrid(data, [flat], [tlat], [0], "linear")
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)


def test_kron_band_is_the_kronecker_product_of_the_dense_weights():
    """tables.kron_band folds two regridded axes into one (flattened) axis: its dense matrix is the Kronecker
    product of the two axes' weight matrices, its coverage the AND of theirs."""
    z, tz = np.arange(0.0, 12.0), np.array([8.5, 0.5, 2.5, 12.5])  # unsorted target, one target outside
    lat, tlat = np.arange(-10.0, 10.1, 1.0), np.arange(-9.0, 9.1, 3.0)
    bo = tables.conservative_band(tables.plain_axis(z), tz)
    bi = tables.conservative_band(tables.format_lat_axis(lat), tlat, spherical=True)
    k = tables.kron_band(bo, bi)
    want = np.einsum("ik,jl->ijkl", bo.dense(), bi.dense()).reshape(z.size * lat.size, tz.size * tlat.size)
    np.testing.assert_array_equal(k.dense(), want)
    np.testing.assert_array_equal(k.cover.reshape(tz.size, tlat.size), bo.cover[:, None] & bi.cover[None, :])
    assert k.n_src == z.size * lat.size and k.k_max == bo.k_max * bi.k_max
    with pytest.raises(NotImplementedError):  # pole rows keep a weight without the spherical correction
        tables.kron_band(bo, tables.conservative_band(tables.format_lat_axis(np.arange(-89.5, 90, 1.0)),
                                                      np.arange(-90.0, 91, 5.0)))


def test_lane_blocks_layout_invariants_random_grids():
    """rg_row_ring_t block layout (lane_cols == 2): whatever regular grid pair the host accepts, every target belongs to
    exactly one block, a block has at most two targets, every tap lies within halo_l / halo_r columns of its block, a
    strip has at most 30 blocks, panels hold whole strips and their ring rows contain the blocks plus one halo block either
    side, and bulk-copy alignment holds (even first columns; multiples of 4 for fp32 panels)."""
    from xarray_regrid_b200 import engine

    rng = np.random.default_rng(2024)
    accepted = 0
    for trial in range(120):
        res = float(rng.choice([0.1, 0.125, 0.2, 0.25, 0.5, 1.0]))
        ratio = float(rng.choice([2.0, 2.25, 2.5, 2.8, 3.0, 1.6, 3.5]))
        vec = int(rng.choice([2, 4]))
        if rng.random() < 0.5:
            lon = np.round(np.arange(0.0, 360.0, res) + rng.choice([0.0, res / 2]), 10)
            tlon = np.round(np.arange(0.0, 360.0, res * ratio) + rng.choice([0.0, res * ratio / 2]), 10)
        else:
            n = int(rng.integers(40, 900))
            lon = np.round(-100.0 + res * np.arange(n), 10)
            tlon = np.round(np.arange(lon[0] + rng.choice([0.0, res / 2, 2 * res]), lon[-1] - res, res * ratio), 10)
        if tlon.size < 2:
            continue
        cols = tables.conservative_band(engine._axis(engine.AxisSpec(lon, tlon, "lon"), False), tlon)
        lay = tables.lane_blocks_layout(cols, vec)
        if lay is None:
            continue
        accepted += 1
        W, n_out = cols.n_src, cols.n_out
        bl0, sb0, sl0, sc0 = lay["block_l0"], lay["strip_b0"], lay["strip_l0"], lay["strip_c0"]
        hl, hr = lay["halo_l"], lay["halo_r"]
        assert bl0[0] == 0 and bl0[-1] == n_out and (np.diff(bl0) >= 0).all() and (np.diff(bl0) <= 2).all()
        assert sb0[0] == 0 and sb0[-1] == bl0.size - 1 and (np.diff(sb0) >= 1).all() and (np.diff(sb0) <= 30).all()
        assert (sl0 == bl0[sb0]).all() and 1 <= hl <= 2 and 1 <= hr <= 2 and lay["n_warps"] <= 12
        ps0, pc0, pnc = lay["panel_strip0"], lay["panel_c0"], lay["panel_nc"]
        assert ps0[0] == 0 and ps0[-1] == sc0.size and (np.diff(ps0) <= lay["n_warps"]).all() and (np.diff(ps0) >= 1).all()
        n_panels = pc0.size
        assert (sc0 % 2 == 0).all()
        if n_panels > 1:
            assert W % 4 == 0 and (pnc <= W).all() and (pnc % 4 == 0).all() and (pc0 % (4 if vec == 4 else 2) == 0).all()
        else:
            assert pc0[0] == 0 and pnc[0] == W
        for p in range(n_panels):
            for s in range(ps0[p], ps0[p + 1]):
                for j in range(sb0[s + 1] - sb0[s]):
                    c0 = int(sc0[s]) + 4 * j
                    if n_panels > 1:
                        rel = (c0 - int(pc0[p])) % W
                        assert 4 <= rel and rel + 8 <= pnc[p]
                    b = sb0[s] + j
                    for l in range(bl0[b], bl0[b + 1]):
                        q = (cols.idx[: cols.cnt[l], l].astype(np.int64) - (c0 - hl)) % W
                        assert (q < 4 + hl + hr).all(), (trial, s, l, q)
    assert accepted >= 40
