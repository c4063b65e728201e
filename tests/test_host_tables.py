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


def test_row_ring_schedule_rejects_pole_rows():
    rows = tables.conservative_band(tables.format_lat_axis(np.arange(-89.0, 90, 2)),
                                    np.arange(-90.0, 91, 1), spherical=True)
    if (rows.idx >= rows.n_src).any():
        assert tables.row_ring_schedule(rows, 1) is None
