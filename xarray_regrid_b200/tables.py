"""Host-side table construction for the B200 regrid kernels (numpy only, no CUDA).

Everything the reference computes per AXIS -- a few thousand numbers -- is computed here
with the reference's own floating-point expressions, then shipped to the GPU as small
index / weight tables.  Everything the reference does per CELL happens in the kernels.

What is folded into tables instead of touching the data (reference paths relative to
``src/xarray_regrid/``):

* ``ensure_monotonic``  (utils.py:411-421)   -> a permutation composed into the indices
* ``format_lon``        (utils.py:341-390)   -> the +-360 shift, the re-sort and the wrap
  padding become a gather map ``formatted column -> original column``
* ``format_lat``        (utils.py:289-338)   -> pole rows become the virtual row indices
  ``n_src`` (south) and ``n_src + 1`` (north), whose value is the longitude-mean of the
  edge row (computed on device by ``rg_row_nanmean_*``)
* ``get_weights`` / ``apply_spherical_correction`` (methods/conservative.py:219-260) ->
  banded (ELL) weights holding only the structurally non-zero taps, built in
  O(n_src + n_tgt) with ``searchsorted`` instead of the dense [n_src, n_tgt] overlap matrix
* ``reindex_like(target)`` (conservative.py:99) -> the output permutation ``restore``
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

LAT_NAMES = ("lat", "latitude")
LON_NAMES = ("lon", "longitude")


# ------------------------------------------------------------------------------ sorting
def monotonic_order(coord):
    """Index array that sorts ``coord`` ascending (stable) and drops duplicate labels,
    or None when the coordinate is already strictly usable as is (utils.py:411-421)."""
    coord = np.asarray(coord)
    order = None
    if coord.size > 1 and np.any(coord[1:] < coord[:-1]):
        order = np.argsort(coord, kind="stable")
        coord = coord[order]
    if coord.size > 1 and np.any(coord[1:] == coord[:-1]):
        _, first = np.unique(coord, return_index=True)
        first = np.sort(first)
        order = first if order is None else order[first]
    return order


@dataclass
class FormattedAxis:
    """A source axis after the reference's pre-formatting.

    ``values[i]`` is the coordinate label of formatted position ``i`` and ``source[i]`` the
    position in the ORIGINAL array it reads from (``n_src`` / ``n_src + 1`` = south / north
    pole row).
    """

    values: np.ndarray
    source: np.ndarray
    n_src: int
    # ORIGINAL positions of the southernmost / northernmost source row (after sorting): the rows whose
    # longitude-mean fills the synthetic pole rows ``n_src`` / ``n_src + 1`` (utils.py:320-333)
    pole_src: tuple = (0, 0)

    @property
    def has_pole_rows(self) -> bool:
        return bool(self.source.size and self.source.max() >= self.n_src)

    @property
    def is_identity(self) -> bool:
        return self.source.size == self.n_src and bool(
            np.array_equal(self.source, np.arange(self.n_src))
        )


def plain_axis(coord) -> FormattedAxis:
    """Sort / de-duplicate only (coordinates that are not named lat / lon)."""
    coord = np.asarray(coord)
    order = monotonic_order(coord)
    src = np.arange(coord.size, dtype=np.int64) if order is None else order.astype(np.int64)
    return FormattedAxis(coord[src], src, coord.size)


def format_lat_axis(lat, lon_mean: bool = True) -> FormattedAxis:
    """utils.py:313-336: add -90 / +90 rows when the edge row is within one (maximum)
    spacing of the pole but not on it.

    The reference sorts first (``ensure_monotonic``, utils.py:277) and then takes the FIRST / LAST row of the
    sorted data, so for a descending or shuffled latitude the south pole row comes from the original row
    ``source[0]`` (not row 0).  ``lon_mean``: the pole row is the longitude-mean of that edge row only when
    the object has a lon / longitude coordinate (utils.py:322-323, 330-331); without one it is a plain copy
    of the edge row, which needs no virtual row -- the pole position simply reads the edge row itself."""
    ax = plain_axis(lat)
    vals, src = ax.values, ax.source
    pole_src = (0, 0)
    if vals.size > 1:
        dy = np.diff(vals).max()
        south = dy - 90 >= vals[0] > -90
        north = 90 - dy <= vals[-1] < 90
        pole_src = (int(src[0]), int(src[-1]))
        if south:
            vals = np.concatenate([[-90], vals])
            src = np.concatenate([[ax.n_src if lon_mean else pole_src[0]], src])
        if north:
            vals = np.concatenate([vals, [90]])
            src = np.concatenate([src, [ax.n_src + 1 if lon_mean else pole_src[1]]])
    return FormattedAxis(vals, src, ax.n_src, pole_src)


def format_lon_axis(lon, tgt_lon) -> FormattedAxis:
    """utils.py:358-388 on indices: shift by +-360 about the wrap point, sort, wrap-pad."""
    lon = np.asarray(lon)
    tgt = np.asarray(tgt_lon)
    t_order = monotonic_order(tgt)
    if t_order is not None:
        tgt = tgt[t_order]
    ax = plain_axis(lon)
    vals, src = ax.values, ax.source

    wrap_point = (tgt[-1] + tgt[0] + 360) / 2
    vals = np.where(vals < wrap_point - 360, vals + 360, vals)
    vals = np.where(vals > wrap_point, vals - 360, vals)
    order = monotonic_order(vals)
    if order is not None:
        vals, src = vals[order], src[order]

    if vals.size > 1 and tgt.size > 1:
        dx_s = np.diff(vals).max()
        dx_t = np.diff(tgt).max()
        if vals.max() - vals.min() >= 360 - dx_s:
            left = (vals[0] - tgt[0] + dx_t / 2) / dx_s
            right = (tgt[-1] - vals[-1] + dx_t / 2) / dx_s
            left = int(np.ceil(max(left, 0)))
            right = int(np.ceil(max(right, 0)))
            n = vals.size
            gather = np.concatenate([np.arange(n - left, n), np.arange(n), np.arange(0, right)])
            new_vals = vals[gather].astype(np.result_type(vals.dtype, np.float64), copy=True)
            if left:
                new_vals[:left] = vals[n - left:] - 360
            if right:
                new_vals[new_vals.size - right:] = vals[:right] + 360
            vals, src = new_vals, src[gather]
            order = monotonic_order(vals)
            if order is not None:
                vals, src = vals[order], src[order]
    return FormattedAxis(vals, src, ax.n_src)


# ------------------------------------------------------------------- conservative bands
def cell_edges(coords):
    """utils.py:128-140: midpoints inside, mirrored outer edges, one point spans all space."""
    coords = np.asarray(coords)
    if coords.size <= 1:
        return np.array([-np.inf, np.inf])
    mid = (coords[:-1] + coords[1:]) / 2  # in the coordinate's own dtype, like the reference
    first = 2 * coords[0] - mid[0]
    last = 2 * coords[-1] - mid[-1]
    return np.concatenate([[first], mid, [last]]).astype(np.float64)


def lat_weight(lat_deg, res_deg):
    """methods/conservative.py:247-260."""
    dlat = np.radians(res_deg)
    lat = np.radians(lat_deg)
    h = np.sin(lat + dlat / 2) - np.sin(lat - dlat / 2)
    return h * dlat / (np.pi * 4)


def valid_threshold(nan_threshold: float) -> float:
    """methods/conservative.py:211-216."""
    return float(1 - np.clip(nan_threshold, 1e-6, 1.0 - 1e-6))


@dataclass
class Band:
    """Tap-major ELL band: ``idx[t, o]`` / ``w[t, o]`` for tap ``t < cnt[o]`` of output ``o``."""

    idx: np.ndarray  # int32 [k_max, n_out]
    w: np.ndarray  # float64 [k_max, n_out]
    cnt: np.ndarray  # int32 [n_out]
    cover: np.ndarray  # uint8 [n_out]
    n_src: int
    restore: np.ndarray | None = None  # output permutation back to the target's own order
    meta: dict = field(default_factory=dict)

    @property
    def n_out(self) -> int:
        return int(self.cnt.size)

    @property
    def k_max(self) -> int:
        return int(self.idx.shape[0])

    def dense(self, n_src_total=None) -> np.ndarray:
        """[n_src, n_out] matrix (tests / debugging only)."""
        n = self.n_src if n_src_total is None else n_src_total
        out = np.zeros((n, self.n_out))
        for t in range(self.k_max):
            use = self.cnt > t
            np.add.at(out, (self.idx[t, use], np.nonzero(use)[0]), self.w[t, use])
        return out


def _pack(first, count, weights_fn):
    """Evaluate ``weights_fn(src_index[k, n_out], valid[k, n_out])`` on the candidate window
    ``first[o] .. first[o] + count[o] - 1`` of every output."""
    k = int(max(count.max(), 1)) if count.size else 1
    taps = np.arange(k)[:, None]
    idx = first[None, :] + taps
    valid = taps < count[None, :]
    return idx, valid, weights_fn(np.where(valid, idx, 0), valid)


def _normalize(w):
    total = w.sum(axis=0)
    total[total == 0] = 1e-12  # utils.py:168
    return w / total


def _compress(idx, w):
    """Keep only the non-zero taps of every column, in their original order."""
    keep = w != 0
    cnt = keep.sum(axis=0).astype(np.int32)
    k = int(max(cnt.max(), 1)) if cnt.size else 1
    order = np.argsort(~keep, axis=0, kind="stable")[:k]
    idx_c = np.take_along_axis(idx, order, axis=0)
    w_c = np.take_along_axis(w, order, axis=0)
    dead = np.arange(k)[:, None] >= cnt[None, :]
    idx_c = np.where(dead, 0, idx_c)
    w_c = np.where(dead, 0.0, w_c)
    return idx_c.astype(np.int32), w_c, cnt


def conservative_band(axis: FormattedAxis, tgt, spherical: bool = False) -> Band:
    """Banded equivalent of ``get_weights`` (+ ``apply_spherical_correction``) for one axis.

    The weights equal the reference's dense matrix entry for entry (the same ``min/max``
    overlap expression, the same two normalisations); exact zeros are dropped, which is
    what the reference's ``sparse.COO`` weights do (conservative.py:297-300).
    """
    tgt = np.asarray(tgt)
    t_order = monotonic_order(tgt)
    tgt_sorted = tgt if t_order is None else tgt[t_order]
    restore = None
    if t_order is not None or tgt_sorted.size != tgt.size:
        restore = np.searchsorted(tgt_sorted, tgt).astype(np.int64)  # reindex_like(target)

    src = axis.values
    se = cell_edges(src)
    te = cell_edges(tgt_sorted)
    s_lo, s_hi = se[:-1], se[1:]
    t_lo, t_hi = te[:-1], te[1:]
    # candidate window: source cells with s_hi > t_lo and s_lo < t_hi
    first = np.searchsorted(s_hi, t_lo, side="right")
    last = np.searchsorted(s_lo, t_hi, side="left") - 1
    count = np.maximum(last - first + 1, 0)

    def overlap(i, valid):
        ov = np.minimum(s_hi[i], t_hi[None, :]) - np.maximum(s_lo[i], t_lo[None, :])
        return np.where(valid, np.maximum(ov, 0), 0.0)

    idx, valid, w = _pack(first, count, overlap)
    w = _normalize(w)
    if spherical:
        res = np.median(np.diff(src, 1))
        lw = lat_weight(src, res)
        w = _normalize(w * np.where(valid, lw[np.where(valid, idx, 0)], 0.0))
    idx_c, w_c, cnt = _compress(np.where(valid, idx, 0), w)

    cover = ((tgt_sorted <= src.max()) & (tgt_sorted >= src.min())).astype(np.uint8)
    # formatted position -> original position (sort, lon shift / wrap, pole rows)
    idx_src = axis.source[idx_c].astype(np.int32)
    idx_src = np.where(np.arange(idx_c.shape[0])[:, None] < cnt[None, :], idx_src, 0).astype(np.int32)
    if restore is not None:
        # emit the outputs directly in the target's own order: reindex_like costs nothing
        idx_src, w_c, cnt, cover = idx_src[:, restore], w_c[:, restore], cnt[restore], cover[restore]
    return Band(np.ascontiguousarray(idx_src), np.ascontiguousarray(w_c), np.ascontiguousarray(cnt),
                np.ascontiguousarray(cover), axis.n_src, restore,
                {"formatted_size": int(src.size), "pole_rows": axis.has_pole_rows, "pole_src": axis.pole_src})


def kron_band(outer: Band, inner: Band) -> Band:
    """Band of the FLATTENED axis ``outer x inner`` (C order: index = i_outer * n_inner + i_inner) whose weights
    are the products of the two bands' weights: contracting it is contracting both axes at once.

    This is how more than two coordinates are regridded conservatively in one pass with the reference's JOINT
    valid mass (``xr.dot`` over all regridded dims, methods/conservative.py:188-198): every axis but the last is
    folded into the kernels' rows axis.  Taps multiply (k_outer * k_inner per output); synthetic pole rows have no
    place in a flattened axis (under the spherical correction they carry zero weight and are dropped anyway)."""
    if (outer.idx >= outer.n_src).any() or (inner.idx >= inner.n_src).any():
        raise NotImplementedError("synthetic pole rows cannot be folded into a flattened axis")
    ko, ki = outer.k_max, inner.k_max
    no, ni = outer.n_out, inner.n_out
    live_o = np.arange(ko)[:, None] < outer.cnt[None, :]
    live_i = np.arange(ki)[:, None] < inner.cnt[None, :]
    # [ko, ki, no, ni] -> taps of output (o, i) ordered (to, ti); dead taps are compacted away below
    idx = outer.idx[:, None, :, None].astype(np.int64) * inner.n_src + inner.idx[None, :, None, :]
    w = outer.w[:, None, :, None] * inner.w[None, :, None, :]
    live = live_o[:, None, :, None] & live_i[None, :, None, :]
    idx = np.where(live, idx, 0).reshape(ko * ki, no * ni)
    w = np.where(live, w, 0.0).reshape(ko * ki, no * ni)
    live = live.reshape(ko * ki, no * ni)
    order = np.argsort(~live, axis=0, kind="stable")
    idx = np.take_along_axis(idx, order, axis=0)
    w = np.take_along_axis(w, order, axis=0)
    cnt = live.sum(axis=0).astype(np.int32)
    k = int(max(cnt.max(), 1))
    cover = (outer.cover[:, None].astype(bool) & inner.cover[None, :].astype(bool)).reshape(-1)
    return Band(np.ascontiguousarray(idx[:k].astype(np.int32)), np.ascontiguousarray(w[:k]), cnt,
                cover.astype(np.uint8), outer.n_src * inner.n_src, None,
                {"pole_rows": False, "pole_src": (0, 0), "kron": (no, ni)})


def identity_band(n: int) -> Band:
    """Pass-through axis (used when only one spatial coordinate is regridded)."""
    return Band(
        np.arange(n, dtype=np.int32)[None, :], np.ones((1, n)), np.ones(n, np.int32),
        np.ones(n, np.uint8), n,
    )


# ------------------------------------------------------------- TMA row-ring schedule
@dataclass
class RowRing:
    """Load order / liveness schedule of the rows band for the TMA row-ring kernel
    (``rg_row_ring_t`` in include/regrid_b200.h)."""

    sched_row: np.ndarray  # int32 [n_sched]
    tap_back: np.ndarray  # int32 [k_max, n_out]: rows back from the newest landed row
    kcnt: np.ndarray  # int32 [n_out]: taps per output, -1 = not covered
    need: np.ndarray  # int32 [n_out]
    release: np.ndarray  # int32 [n_out]
    run_first: np.ndarray  # int32 [n_runs + 1]
    run_sched: np.ndarray  # int32 [n_runs + 1]
    max_live: int
    pad_shift: int
    strip_l0: np.ndarray = None  # int32 [n_strips + 1]
    strip_c0: np.ndarray = None  # int32 [n_strips]
    strip_nc: np.ndarray = None  # int32 [n_strips]
    n_warps: int = 0
    lane_cols: int = 0  # 4: lane-owned columns structure (see ``lane_owned_columns``), else 0
    # column panels (two-phase ring kernel): panel p owns strips panel_strip0[p] .. panel_strip0[p+1]-1 and the
    # source columns panel_c0[p] .. + panel_nc[p] - 1 (modulo n_src); strip_c0 is RELATIVE to the panel's first
    # column.  One panel covering the whole row = the classic layout.
    panel_strip0: np.ndarray = None  # int32 [n_panels + 1]
    panel_c0: np.ndarray = None  # int32 [n_panels]
    panel_nc: np.ndarray = None  # int32 [n_panels]
    # block layout of the register-resident kernel (lane_cols == 2, see ``lane_blocks_layout``); strip_c0 is then
    # the ABSOLUTE source column of the strip's first own block
    strip_b0: np.ndarray = None  # int32 [n_strips + 1]: first block of every strip
    block_l0: np.ndarray = None  # int32 [n_blocks + 1]: first target of every block
    halo_l: int = 0
    halo_r: int = 0

    @property
    def n_strips(self) -> int:
        return int(self.strip_c0.size)

    def krec(self) -> np.ndarray:
        """int32 [n_out, 4] = {need, release, kcnt, 8 x 4-bit tap_back} (the kernel's 16-byte record)."""
        packed = np.zeros(self.need.size, np.uint32)
        for t in range(min(self.tap_back.shape[0], 8)):
            packed |= (self.tap_back[t].astype(np.uint32) & 15) << np.uint32(4 * t)
        return np.stack([self.need, self.release, self.kcnt, packed.view(np.int32)], axis=1).astype(np.int32)

    def wrec(self, rows: "Band", dtype) -> np.ndarray:
        """[n_out, 8] tap weights of the rows band, output-major, zero padded."""
        out = np.zeros((rows.n_out, 8), dtype)
        k = min(rows.k_max, 8)
        live = np.arange(k)[:, None] < rows.cnt[None, :]
        out[:, :k] = np.where(live, rows.w[:k], 0.0).T.astype(dtype)
        return out

    @property
    def strip_cols_max(self) -> int:
        return int(self.strip_nc.max())

    @property
    def n_panels(self) -> int:
        return 0 if self.panel_c0 is None else int(self.panel_c0.size)

    @property
    def panel_cols_max(self) -> int:
        return int(self.panel_nc.max()) if self.n_panels else 0

    def single_panel(self, n_src: int) -> None:
        """The whole row as one panel (strip_c0 relative to column 0 = absolute)."""
        self.panel_strip0 = np.asarray([0, self.n_strips], np.int32)
        self.panel_c0 = np.asarray([0], np.int32)
        self.panel_nc = np.asarray([n_src], np.int32)

    @property
    def n_runs(self) -> int:
        return int(self.run_first.size - 1)


def column_strips(cols: Band, vec: int, max_warps: int = 12, n_warps: int | None = None):
    """Cut the target columns into strips of <= 32 outputs for the consumer warps of the ring
    kernel and find, per strip, the (circular) range of source columns that covers all its taps.

    Returns ``(strip_l0, strip_c0, strip_nc, n_warps)``.  The warp count minimises a simple
    instruction-count model of the kernel (per-warp fixed cost + 32-lane steps of the latitude
    pass), which favours strips whose source range is just under a multiple of 32 lane steps.
    """
    n_out, n_src = cols.n_out, cols.n_src

    def build(tps):
        l0 = np.arange(0, n_out, tps)
        l0 = np.append(l0, n_out)
        c0s, ncs = [], []
        for a, b in zip(l0[:-1], l0[1:]):
            taps = [cols.idx[: cols.cnt[l], l] for l in range(a, b) if cols.cnt[l] > 0]
            if not taps:
                c0s.append(0)
                ncs.append(vec)
                continue
            s = np.unique(np.concatenate(taps)).astype(np.int64)
            gaps = np.diff(np.append(s, s[0] + n_src))  # circular gaps
            g = int(np.argmax(gaps))
            start, end = int(s[(g + 1) % s.size]), int(s[g])
            length = (end - start) % n_src + 1
            c0 = start - start % vec
            length += start - c0
            length = -(-length // vec) * vec
            if length >= n_src:
                c0, length = 0, n_src
            c0s.append(c0)
            ncs.append(length)
        return l0.astype(np.int32), np.asarray(c0s, np.int32), np.asarray(ncs, np.int32)

    if n_warps is not None:  # fixed decomposition: ceil(n_out / n_warps) targets per strip
        l0, c0, nc = build(-(-n_out // n_warps))
        return l0, c0, nc, min(n_warps, c0.size)
    best = None
    for n_warps in range(1, max_warps + 1):
        tps = min(32, -(-n_out // n_warps))
        l0, c0, nc = build(tps)
        steps = -(-(nc // vec) // 32)
        per_strip = steps * 63 + 75
        per_warp = np.zeros(n_warps)
        for i, c in enumerate(per_strip):
            per_warp[i % n_warps] += c
        # the kernel is bound by the serial per-target-row path of its busiest warp (every warp visits every
        # target row), so the critical path dominates; total issue work only breaks ties
        cost = 8.0 * per_warp.max() + 0.25 * (per_strip.sum() + 75 * min(n_warps, c0.size))
        if best is None or cost < best[0] - 1e-9:
            best = (cost, l0, c0, nc, min(n_warps, c0.size))
    return best[1:]


def column_panels(cols: Band, vec: int, warps: int = 6, table_bytes: int = 0, min_slots: int = 6,
                  cta_budget: int = 74 * 1024):
    """Column panels for the two-phase ring kernel when the target row is too wide for one CTA's consumer warps
    (or so narrow that one CTA per SM would leave the SM idle): the target columns are cut into strips of equal
    size (<= 32, chosen with the same instruction-count model as ``column_strips``), ``warps`` consecutive strips
    form a panel, and a CTA works on ONE panel of a (slab, run of target rows) task: its ring rows hold only the
    panel's source columns (short rows: several CTAs fit an SM, i.e. several independent row pipelines) and every
    consumer warp owns exactly one strip.  The strip size is the cheapest (per target) one whose CTA -- a ring of
    ``min_slots`` panel rows, the warps' strip buffers and ``table_bytes`` of per-row tables -- still fits
    ``cta_budget`` bytes of shared memory, i.e. three CTAs per SM.

    Returns ``(strip_l0, strip_c0_relative, strip_nc, panel_strip0, panel_c0, panel_nc)`` or None when a panel's
    source range cannot be expressed (it would cover the whole circle)."""
    n_out, n_src = cols.n_out, cols.n_src
    if n_out == 0 or n_src % vec:
        return None

    def span_of(a, b):
        """Circular source range (start, length) covering every tap of targets a .. b-1, vec-aligned."""
        taps = [cols.idx[: cols.cnt[l], l] for l in range(a, b) if cols.cnt[l] > 0 and cols.cover[l]]
        if not taps:
            return 0, vec
        u = np.unique(np.concatenate(taps)).astype(np.int64)
        gaps = np.diff(np.append(u, u[0] + n_src))
        g = int(np.argmax(gaps))
        start, end = int(u[(g + 1) % u.size]), int(u[g])
        length = (end - start) % n_src + 1
        c0 = start - start % vec
        length = -(-(length + start - c0) // vec) * vec
        if length >= n_src:
            return 0, n_src
        return c0, length

    esz = 16 // vec
    ratio = n_src / n_out
    r = int(round(ratio))
    padded = r >= 2 and abs(ratio - r) < 0.26 and (r & (r - 1)) == 0  # pad_shift of row_ring_schedule

    def layout(tps):
        l0 = np.append(np.arange(0, n_out, tps), n_out)
        spans = [span_of(a, b) for a, b in zip(l0[:-1], l0[1:])]
        n_strips = len(spans)
        n_pan = -(-n_strips // warps)
        per = -(-n_strips // n_pan)
        p_strip0 = np.append(np.arange(0, n_strips, per), n_strips)
        p_c0, p_nc, s_c0 = [], [], []
        for a, b in zip(p_strip0[:-1], p_strip0[1:]):
            c0, nc = span_of(int(l0[a]), int(l0[b]))
            if nc >= n_src and n_pan > 1:
                return None
            p_c0.append(c0)
            p_nc.append(nc)
            for sidx in range(a, b):
                rel = (spans[sidx][0] - c0) % n_src
                if rel + spans[sidx][1] > nc and nc < n_src:  # (a full-circle panel wraps its strips like the classic layout)
                    return None
                s_c0.append(rel)
        return (l0.astype(np.int32), np.asarray(s_c0, np.int32), np.asarray([nc for _, nc in spans], np.int32),
                p_strip0.astype(np.int32), np.asarray(p_c0, np.int32), np.asarray(p_nc, np.int32))

    cands = []
    for tps in sorted({-(-n_out // -(-n_out // t)) for t in range(64, 7, -1)}, reverse=True):  # equal strips
        lay = layout(tps)
        if lay is None:
            continue
        strip_nc_max, panel_nc_max = int(lay[2].max()), int(lay[5].max())
        steps = -(-(strip_nc_max // vec) // 32)
        cost = (110 + steps * 55 + -(-tps // 32) * 48) / tps  # per target: fixed + latitude steps + passes
        row_stride = -(-(panel_nc_max * esz) // 128) * 128
        strip_slots = strip_nc_max + (strip_nc_max // r if padded else 0) + 1
        strip_bytes = -(-(strip_slots * 2 * esz) // 128) * 128
        need = min_slots * row_stride + int(np.diff(lay[3]).max()) * strip_bytes + 512 + table_bytes
        cands.append((need > cta_budget, len(lay[1]) < warps, cost, tps, lay))
    if not cands:
        return None
    # fitting three CTAs per SM first, a strip for every consumer warp next, then the cheapest per target
    cands.sort(key=lambda c: (c[0], c[1], c[2]))
    return cands[0][4]


def row_ring_schedule(rows: Band, n_runs: int, cols: Band | None = None, vec: int = 2,
                      allow_pairs: bool = True) -> RowRing | None:
    """Schedule for streaming the source rows of ``rows`` through a shared-memory ring.

    Outputs are visited in order and split into ``n_runs`` contiguous runs.  Inside a run a
    source row is loaded when first needed and stays in the ring until its last use, so rows
    shared by neighbouring outputs are read once.  The synthetic pole rows of ``format_lat`` (indices
    ``n_src`` and ``n_src + 1``) are scheduled like any other row: the kernels fetch them from the
    ``[batch, 2, W]`` pole buffer.
    """
    n_out, k_max = rows.n_out, rows.k_max
    cnt = np.where(rows.cover > 0, rows.cnt, 0)
    if n_out == 0 or (rows.idx > rows.n_src + 1).any():
        return None
    if k_max > 8:
        return None
    n_runs = int(max(1, min(n_runs, n_out)))
    run_first = np.unique(np.round(np.linspace(0, n_out, n_runs + 1)).astype(np.int64))
    tap_pos = np.zeros((k_max, n_out), np.int32)
    need = np.zeros(n_out, np.int32)
    release = np.zeros(n_out, np.int32)
    sched: list[int] = []
    run_sched = [0]
    max_live = 1
    for r in range(run_first.size - 1):
        k0, k1 = int(run_first[r]), int(run_first[r + 1])
        last_use: dict[int, int] = {}
        for k in range(k0, k1):
            for t in range(int(cnt[k])):
                last_use[int(rows.idx[t, k])] = k
        pos_of: dict[int, int] = {}
        order: list[int] = []
        freed = 0
        for k in range(k0, k1):
            for t in range(int(cnt[k])):
                row = int(rows.idx[t, k])
                if row not in pos_of:
                    pos_of[row] = len(order)
                    order.append(row)
                tap_pos[t, k] = pos_of[row]
            need[k] = len(order)
            max_live = max(max_live, len(order) - freed)
            while freed < len(order) and last_use[order[freed]] <= k:
                freed += 1
            release[k] = freed
        assert freed == len(order)
        sched.extend(order)
        run_sched.append(len(sched))

    # shared-memory padding of the t / v column arrays: with an integer power-of-two ratio r
    # of source to target columns the longitude pass reads with stride r; one pad word every
    # r makes the stride odd and the accesses bank-conflict free
    pad_shift = 31
    if cols is not None and cols.n_out > 0:
        ratio = cols.n_src / cols.n_out
        r = int(round(ratio))
        if r >= 2 and abs(ratio - r) < 0.26 and (r & (r - 1)) == 0:
            pad_shift = r.bit_length() - 1
    live = np.arange(k_max)[:, None] < cnt[None, :]
    tap_back = np.where(live, need[None, :] - 1 - tap_pos, 0).astype(np.int32)
    kcnt = np.where(rows.cover > 0, rows.cnt, -1).astype(np.int32)
    if max_live > 15:
        return None
    ring = RowRing(
        np.asarray(sched, np.int32), tap_back, kcnt, need, release,
        run_first.astype(np.int32), np.asarray(run_sched, np.int32), int(max_live), pad_shift,
    )
    if cols is not None:
        if cols.n_src % vec != 0:
            return None
        if k_max >= 6 and cols.k_max > 5:
            # many row taps and no register-resident columns: the generic kernel's plain streaming loads win
            # (measured: 0.25 -> 1.5 deg, 7 x 7 taps: 5.2 vs 4.6 TB/s for the two-phase ring kernel)
            return None
        ring.strip_l0, ring.strip_c0, ring.strip_nc, ring.n_warps = column_strips(cols, vec)
        ring.single_panel(cols.n_src)
        too_wide = ring.strip_c0.size > ring.n_warps
        if not too_wide:
            ring.lane_cols = lane_owned_columns(cols, ring.strip_l0, ring.n_warps)
            if not ring.lane_cols and cols.n_out <= 31 * 12:
                # the register-resident kernel wants one strip of <= 31 targets per warp: try that decomposition
                n_warps = int(min(12, max(1, -(-cols.n_out // 8))))
                while -(-cols.n_out // n_warps) > 31:
                    n_warps += 1
                l0, c0, nc, _ = column_strips(cols, vec, n_warps=n_warps)
                if l0.size - 1 <= n_warps and lane_owned_columns(cols, l0, n_warps):
                    ring.strip_l0, ring.strip_c0, ring.strip_nc, ring.n_warps = l0, c0, nc, n_warps
                    ring.lane_cols = lane_owned_columns(cols, l0, n_warps)
                    ring.single_panel(cols.n_src)
        blocks = lane_blocks_layout(cols, vec) if (allow_pairs and not ring.lane_cols and rows.k_max <= 5) else None
        if blocks is not None:
            # coarsening by 2 - 3: the register-resident kernel with up to two targets per lane (conservative
            # modes only: its zero-weight taps would let a NaN through in the NaN-propagating interpolation mode)
            for name, val in blocks.items():
                setattr(ring, name, val)
            ring.strip_nc = np.full(ring.strip_c0.size, 4 * 32, np.int32)
            ring.lane_cols = 2
        elif too_wide or (not ring.lane_cols and (ring.n_warps < 8 or rows.k_max <= 4)):
            # more than 12 x 32 target columns (a warp would own several strips and reload its column taps for
            # every target row), few short strips (one CTA per SM leaves it idle), or few row taps (coarsening
            # ratios of 2 - 2.5: two new source rows per target row leave no time for one CTA's per-row dependency
            # chain): column panels -- every CTA works on one panel of 6 strips, short ring rows, several CTAs
            # (independent row pipelines) per SM
            esz = 16 // vec
            w_row = 4 if rows.k_max <= 4 else 8
            pan = column_panels(cols, vec, table_bytes=n_out * (16 + w_row * esz) + 4 * len(sched),
                                min_slots=int(max_live) + 3)
            if pan is None:
                return None
            (ring.strip_l0, ring.strip_c0, ring.strip_nc, ring.panel_strip0, ring.panel_c0, ring.panel_nc) = pan
            ring.n_warps = int(np.diff(ring.panel_strip0).max())
            ring.lane_cols = 0
    return ring


def lane_owned_columns(cols: Band, strip_l0: np.ndarray, n_warps: int, step: int = 4) -> int:
    """``step`` if the columns band has the regular coarsening structure the register-resident kernel
    needs (``rg_row_ring_t.lane_cols``), else 0: every target covered with 1..step+1 contiguous taps
    (modulo n_src) starting at an even source column, neighbouring targets of a strip exactly ``step``
    source columns apart, strips of <= 31 targets, one strip per consumer warp, at most 12 consumer warps."""
    n_src, n_out = cols.n_src, cols.n_out
    if n_out == 0 or n_src % 2 or cols.k_max > step + 1 or n_warps > 12 or strip_l0.size - 1 > n_warps:
        return 0
    if not (cols.cover > 0).all() or (cols.cnt < 1).any() or (cols.cnt > step + 1).any():
        return 0
    first = cols.idx[0].astype(np.int64)
    if (first % 2).any():
        return 0
    for i in range(1, cols.k_max):
        live = cols.cnt > i
        if ((cols.idx[i].astype(np.int64) - first - i) % n_src)[live].any():
            return 0
    for a, b in zip(strip_l0[:-1], strip_l0[1:]):
        if b - a > 31:
            return 0
        if ((first[a + 1:b] - first[a:b - 1] - step) % n_src).any():
            return 0
    return step


def lane_blocks_layout(cols: Band, vec: int = 2, max_warps: int = 12, panel_warps: int = 12,
                       blocks_per_warp: int = 30, warp_choices=()):
    """Layout of the register-resident kernel's BLOCK variant (``rg_row_ring_t.lane_cols == 2``), or None.

    Coarsening by 2 - 3: a lane owns a block of 4 source columns ``[c00 + 4b, c00 + 4b + 3]`` and produces the (at
    most two, consecutive) targets assigned to its block; every tap of such a target lies in the block or within
    two columns of it (``halo_l`` / ``halo_r`` <= 2: fetched from the neighbouring lanes with shuffles).  Needs all
    targets covered with 1..6 contiguous taps (modulo n_src).  Lane 0 of a warp and the lane after its last block
    only produce halos, so a warp (= strip) holds at most ``blocks_per_warp`` blocks.  More than ``max_warps``
    strips are split into the fewest column panels of at most ``panel_warps`` strips (wide CTAs beat
    several narrow ones: 0.1 -> 0.25 deg 3 panels of 10 warps 5.06 TB/s, 5 of 6 warps 4.5 TB/s), each panel a ring row of its own blocks
    plus one halo block either side.

    Returns a dict: ``block_l0`` (CSR: first target of every block), ``strip_b0``, ``strip_l0``, ``strip_c0``
    (absolute source column of a strip's first own block), ``n_warps``, ``halo_l``, ``halo_r``, ``panel_strip0``,
    ``panel_c0``, ``panel_nc``."""
    n_src, n_out = cols.n_src, cols.n_out
    if n_out < 2 or n_src % 2 or n_src < 16 or cols.k_max > 6:
        return None
    if not (cols.cover > 0).all() or (cols.cnt < 1).any():
        return None
    first = cols.idx[0].astype(np.int64)
    for i in range(1, cols.k_max):
        live = cols.cnt > i
        if ((cols.idx[i].astype(np.int64) - first - i) % n_src)[live].any():
            return None
    # unwrapped first taps (forward steps of a few columns, the date line crossed at most once)
    step = (first[1:] - first[:-1]) % n_src
    if (step > 8).any():
        return None
    first_u = first[0] + np.concatenate([[0], np.cumsum(step)])
    last_u = first_u + cols.cnt - 1
    if last_u[-1] - first_u[0] > n_src + 4:
        return None
    best = None
    align = 4 if vec == 4 else 2  # (panel rows start at c00 - 4 + 4k: 16-byte aligned bulk copies)
    for c00 in range(int(first_u[0]) - 4, int(first_u[0]) + 3):
        if c00 % align:
            continue
        blk = np.empty(n_out, np.int64)
        cur, fill, ok = -1, 0, True
        for l in range(n_out):
            lo_b = -((-(int(last_u[l]) - c00 - 5)) // 4)  # ceil
            hi_b = (int(first_u[l]) - c00 + 2) // 4
            b = max(lo_b, 0, cur if fill < 2 else cur + 1)
            if b > hi_b:
                ok = False
                break
            cur, fill = (cur, fill + 1) if b == cur else (b, 1)
            blk[l] = b
        if not ok:
            continue
        own0 = c00 + 4 * blk
        halo_l = int(max(0, (own0 - first_u).max()))
        halo_r = int(max(0, (last_u - own0 - 3).max()))
        score = (int(blk[-1]) + 1, halo_l + halo_r, halo_l)  # (the kernel has a (1, 2) halo instance, no (2, 1))
        if best is None or score < best[0]:
            best = (score, c00, blk, halo_l, halo_r)
    if best is None:
        return None
    (n_blocks, _, _), c00, blk, halo_l, halo_r = best
    if 4 * n_blocks > n_src + 4:
        return None
    block_l0 = np.searchsorted(blk, np.arange(n_blocks + 1)).astype(np.int32)  # CSR over blocks
    total_warps = -(-n_blocks // blocks_per_warp)
    if total_warps <= max_warps:
        # (as few warps as the blocks need: full warps beat more, emptier ones -- 1 -> 2.5 deg: 3 warps of 30 blocks
        # 4.2 TB/s, 4 warps of 23 blocks 2.7 TB/s; ``warp_choices`` is the experiment's knob)
        n_panels = 1
        n_warps = next((w for w in warp_choices if w >= total_warps), total_warps)
        n_warps = max(total_warps, min(n_warps, n_blocks // 8))
    else:
        n_panels = -(-total_warps // panel_warps)
        n_warps = -(-total_warps // n_panels)
    n_strips = n_panels * n_warps
    per = -(-n_blocks // n_strips)
    if per > blocks_per_warp:
        return None
    strip_b0 = np.minimum(np.arange(n_strips + 1) * per, n_blocks).astype(np.int32)
    strip_b0 = strip_b0[: int(np.searchsorted(strip_b0, n_blocks)) + 1]  # drop empty strips at the end
    n_strips = strip_b0.size - 1
    panel_strip0 = np.minimum(np.arange(n_panels + 1) * n_warps, n_strips).astype(np.int32)
    panel_strip0 = panel_strip0[: int(np.searchsorted(panel_strip0, n_strips)) + 1]
    n_panels = panel_strip0.size - 1
    if n_panels == 1:
        panel_c0, panel_nc = np.zeros(1, np.int32), np.asarray([n_src], np.int32)
    else:
        pb0, pb1 = strip_b0[panel_strip0[:-1]], strip_b0[panel_strip0[1:]]
        panel_c0 = ((c00 + 4 * pb0.astype(np.int64) - 4) % n_src).astype(np.int32)
        panel_nc = (4 * (pb1 - pb0 + 2)).astype(np.int32)
        if n_src % 4 or (panel_nc > n_src).any():
            return None
    return dict(block_l0=block_l0, strip_b0=strip_b0, strip_l0=block_l0[strip_b0].astype(np.int32),
                strip_c0=((c00 + 4 * strip_b0[:-1].astype(np.int64)) % n_src).astype(np.int32),
                n_warps=int(min(n_warps, n_strips)), halo_l=max(halo_l, 1), halo_r=max(halo_r, 1),
                panel_strip0=panel_strip0, panel_c0=panel_c0, panel_nc=panel_nc)


def band_window_rows(band: "Band", cap: int = 16, max_rows: int = 32, step: int = 4) -> int:
    """``rg_band_t.window_rows``: the largest multiple of ``step`` (<= ``max_rows``) such that every aligned
    group of that many consecutive outputs takes its taps from at most ``cap`` consecutive source indices
    (0 if even groups of ``step`` do not).  Exact, from the band's own indices: irregular source spacing or
    unsorted targets simply yield a small value (or 0) instead of breaking the upsampling kernel's window."""
    live = (np.arange(band.k_max)[:, None] < band.cnt[None, :]) & (band.cover[None, :] > 0)
    real = live & (band.idx < band.n_src)
    if band.n_out == 0 or not real.any():
        return 0
    hi = np.where(real, band.idx, -1).max(axis=0).astype(np.int64)
    lo = np.where(real, band.idx, np.iinfo(np.int32).max).min(axis=0).astype(np.int64)
    for rows in range(max_rows, 0, -step):
        starts = np.arange(0, band.n_out, rows)
        g_hi = np.maximum.reduceat(hi, starts)
        g_lo = np.minimum.reduceat(lo, starts)
        if np.all((g_hi < 0) | (g_hi - g_lo + 1 <= cap)):
            return rows
    return 0


def band_group_span(band: "Band", group: int = 4) -> int:
    """``rg_band_t.group_span``: for a band whose taps are contiguous (``idx[t] == idx[0] + t``), the largest
    number of source indices spanned by the taps of an aligned group of ``group`` consecutive covered outputs;
    0 when the taps are not contiguous (or nothing is covered)."""
    live = (np.arange(band.k_max)[:, None] < band.cnt[None, :]) & (band.cover[None, :] > 0)
    if band.n_out == 0 or not live.any() or (band.idx[live] >= band.n_src).any():
        return 0
    idx = band.idx.astype(np.int64)
    if ((idx - idx[:1] - np.arange(band.k_max)[:, None]) != 0)[live].any():
        return 0
    has = live.any(axis=0)
    hi = np.where(live, idx, -1).max(axis=0)
    lo = np.where(live, idx, np.iinfo(np.int32).max).min(axis=0)
    starts = np.arange(0, band.n_out, group)
    g_hi = np.maximum.reduceat(np.where(has, hi, -1), starts)
    g_lo = np.minimum.reduceat(np.where(has, lo, np.iinfo(np.int32).max), starts)
    spans = np.where(g_hi >= 0, g_hi - g_lo + 1, 0)
    return int(spans.max())


# ------------------------------------------------------------ interpolation bands (xr.interp)
def _restore_not_needed(n_out):
    return None


def linear_band(axis: FormattedAxis, tgt) -> Band:
    """2-tap band equal to scipy ``interp1d(kind="linear")`` as xarray calls it
    (scipy/interpolate/_interpolate.py ``_call_linear``; reference methods/interp.py:43-46).

    ``m = searchsorted(x, t).clip(1, n-1)``; weights ``(t - x_lo)/(x_hi - x_lo)`` and
    ``(x_hi - t)/(x_hi - x_lo)``; targets outside ``[x[0], x[-1]]`` are NaN.  Both taps are ALWAYS
    kept (even at weight 0) because the reference's ``w_hi*y_hi + w_lo*y_lo`` lets a NaN tap poison the
    result at weight 0.  Targets keep their own order (no sort needed for interpolation).
    """
    x = np.asarray(axis.values, dtype=np.float64)
    t = np.asarray(tgt)
    n = x.size
    if n < 2:
        raise ValueError("x and y arrays must have at least 2 entries")
    m = np.searchsorted(x, t).clip(1, n - 1).astype(int)
    lo, hi = m - 1, m
    x_lo, x_hi = x[lo], x[hi]
    w_hi = (t - x_lo) / (x_hi - x_lo)
    w_lo = (x_hi - t) / (x_hi - x_lo)
    cover = ~((t < x[0]) | (t > x[-1]))
    idx = np.stack([axis.source[lo], axis.source[hi]]).astype(np.int32)
    w = np.stack([w_lo, w_hi]).astype(np.float64)
    return Band(idx, w, np.full(t.size, 2, np.int32), cover.astype(np.uint8), axis.n_src, None,
                {"pole_rows": axis.has_pole_rows, "pole_src": axis.pole_src, "method": "linear"})


def nearest_index(axis: FormattedAxis, tgt):
    """Source index and in-range mask of scipy ``interp1d(kind="nearest")``
    (``_interpolate.py`` ``_call_nearest``): ``searchsorted(x[1:]/2 + x[:-1]/2, t, side="left")``,
    i.e. exact midpoints go to the LOWER neighbour; outside ``[x[0], x[-1]]`` -> NaN."""
    x = np.asarray(axis.values)
    t = np.asarray(tgt)
    if x.size < 1:
        raise ValueError("x and y arrays must have at least 1 entries")
    x_bds = x / 2.0
    x_bds = x_bds[1:] + x_bds[:-1]
    i = np.searchsorted(x_bds, t, side="left").clip(0, x.size - 1).astype(np.intp)
    cover = ~((t < x[0]) | (t > x[-1]))
    return axis.source[i].astype(np.int32), cover.astype(np.uint8)


def nearest_band(axis: FormattedAxis, tgt) -> Band:
    idx, cover = nearest_index(axis, tgt)
    return Band(idx[None, :], np.ones((1, idx.size)), np.ones(idx.size, np.int32), cover, axis.n_src,
                None, {"pole_rows": axis.has_pole_rows, "pole_src": axis.pole_src, "method": "nearest"})


def cubic_bands(axis: FormattedAxis, tgt, rel_cut: float = 1e-18):
    """Prefilter + evaluation bands equal to scipy ``interp1d(kind="cubic")`` =
    ``make_interp_spline(x, y, k=3)`` (global not-a-knot spline) followed by B-spline evaluation.

    Both steps are linear in ``y``: ``c = A^-1 y`` (A = collocation matrix incl. the not-a-knot rows) and
    ``out = E c`` (E = 4 non-zero B-spline basis values per target).  ``A^-1`` is computed with scipy's
    own routine applied to the identity and cut where its entries fall below ``rel_cut`` of the row
    maximum (they decay by ~0.27 per node, so ~ +-31 taps survive at 1e-18 -- far below fp64 rounding).

    Returns ``(prefilter, evaluate)``: ``prefilter`` maps the ORIGINAL source axis to the formatted axis
    (n_fmt outputs), ``evaluate`` maps the formatted axis to the targets.
    """
    from scipy.interpolate import BSpline, make_interp_spline

    x = np.asarray(axis.values, dtype=np.float64)
    t = np.asarray(tgt, dtype=np.float64)
    n = x.size
    if n < 4:
        raise ValueError("x and y arrays must have at least 4 entries")
    spl = make_interp_spline(x, np.eye(n), k=3, check_finite=False)
    minv = np.asarray(spl.c)  # [n coefficients, n data points]
    keep = np.abs(minv) >= rel_cut * np.abs(minv).max(axis=1, keepdims=True)
    cnt = keep.sum(axis=1).astype(np.int32)
    k_max = int(cnt.max())
    order = np.argsort(~keep, axis=1, kind="stable")[:, :k_max]  # kept columns first, ascending
    w = np.take_along_axis(minv, order, axis=1)
    dead = np.arange(k_max)[None, :] >= cnt[:, None]
    w = np.where(dead, 0.0, w)
    src = axis.source[np.where(dead, order[:, :1], order)]
    pre = Band(np.ascontiguousarray(src.T.astype(np.int32)), np.ascontiguousarray(w.T),
               cnt, np.ones(n, np.uint8), axis.n_src, None,
               {"pole_rows": axis.has_pole_rows, "pole_src": axis.pole_src, "method": "cubic-prefilter"})

    cover = ~((t < x[0]) | (t > x[-1]))
    tc = np.clip(t, x[0], x[-1])
    dm = BSpline.design_matrix(tc, spl.t, 3).tocsr()
    e_idx = np.zeros((4, t.size), np.int32)
    e_w = np.zeros((4, t.size))
    for o in range(t.size):
        a, b = dm.indptr[o], dm.indptr[o + 1]
        cols, vals = dm.indices[a:b], dm.data[a:b]
        k = cols.size
        e_idx[:k, o] = cols
        e_w[:k, o] = vals
        e_idx[k:, o] = cols[0]
    ev = Band(e_idx, e_w, np.full(t.size, 4, np.int32), cover.astype(np.uint8), n, None,
              {"pole_rows": False, "method": "cubic-evaluate"})
    return pre, ev


def cubic_lu(axis: FormattedAxis):
    """LU factors (no pivoting) of the not-a-knot collocation matrix of ``make_interp_spline(x, y, k=3)`` on
    the formatted axis, for ``rg_spline_solve_f64``: ``(lower [2, n], upper [3, n])`` with
    ``lower = [L[i,i-1], L[i,i-2]]`` and ``upper = [1/U[i,i], U[i,i+1], U[i,i+2]]``, or ``None`` when the
    matrix is not penta-diagonal, a pivot is tiny, or the substitution does not reproduce scipy's own
    coefficients to 1e-13 (then the truncated-inverse band of ``cubic_bands`` is used).
    """
    from scipy.interpolate import BSpline, make_interp_spline

    x = np.asarray(axis.values, dtype=np.float64)
    n = x.size
    if n < 6:
        return None
    probe = np.random.default_rng(12345).standard_normal((n, 3))
    spl = make_interp_spline(x, probe, k=3, check_finite=False)
    a = BSpline.design_matrix(x, spl.t, 3).todia()
    diag = {d: np.zeros(n) for d in range(-2, 3)}  # diag[d][i] = A[i, i + d]
    for off, data in zip(a.offsets, a.data):
        cols = np.arange(max(off, 0), min(n, n + off))
        if abs(int(off)) > 2:
            if np.any(data[cols] != 0.0):
                return None  # not penta-diagonal
            continue
        diag[int(off)][cols - off] = data[cols]
    d0, u1, u2 = diag[0].copy(), diag[1].copy(), diag[2].copy()
    s1, s2 = diag[-1].copy(), diag[-2].copy()  # A[i, i-1], A[i, i-2]
    l1, l2 = np.zeros(n), np.zeros(n)
    for i in range(n):
        # eliminate A[i, i-2] with row i-2, then A[i, i-1] with row i-1 (rows above are already U rows)
        if i >= 2:
            l2[i] = s2[i] / d0[i - 2]
            s1[i] -= l2[i] * u1[i - 2]
            d0[i] -= l2[i] * u2[i - 2]
        if i >= 1:
            l1[i] = s1[i] / d0[i - 1]
            d0[i] -= l1[i] * u1[i - 1]
            u1[i] -= l1[i] * u2[i - 1]
        if not np.isfinite(d0[i]) or abs(d0[i]) < 1e-10 * max(1.0, abs(diag[0]).max()):
            return None
    dinv = 1.0 / d0
    # the device's recurrences, checked against scipy's banded LAPACK solve
    z = probe.copy()
    for i in range(n):
        if i >= 1:
            z[i] -= l1[i] * z[i - 1]
        if i >= 2:
            z[i] -= l2[i] * z[i - 2]
    c = z
    for i in range(n - 1, -1, -1):
        acc = c[i].copy()
        if i + 1 < n:
            acc -= u1[i] * c[i + 1]
        if i + 2 < n:
            acc -= u2[i] * c[i + 2]
        c[i] = acc * dinv[i]
    ref = np.asarray(spl.c)
    if not np.all(np.abs(c - ref) <= 1e-13 * np.abs(ref).max()):
        return None
    return np.stack([l1, l2]), np.stack([dinv, u1, u2])


# ------------------------------------------------------------- group-by bins (most_common / stat)
@dataclass
class Bins:
    """Closed-left target bins over one (formatted, ascending) source axis: ``rg_bins_t``."""

    start: np.ndarray  # int32 [n_out] first formatted position of the bin (sorted-target order)
    count: np.ndarray  # int32 [n_out]
    cover: np.ndarray  # uint8 [n_out]
    out_index: np.ndarray  # int32 [n_out] index of the bin's target in the target's own order
    n_src: int
    src_offset: int
    src_step: int
    gather: np.ndarray | None = None  # set when the source order is not affine: materialise x[gather] first

    @property
    def n_out(self) -> int:
        return int(self.start.size)


def bin_breaks(tgt_sorted):
    """methods/_shared.py:12-18: ``append(t, t[-1] + step) - step / 2`` with the MEDIAN step."""
    step = np.median(np.diff(tgt_sorted, n=1))
    return np.append(tgt_sorted, tgt_sorted[-1] + step) - step / 2


def bins_axis(axis: FormattedAxis, tgt) -> Bins:
    """Bin table of one axis, following ``construct_intervals`` (_shared.py:12-18),
    ``reduce_data_to_new_domain`` (_shared.py:92-108) and the coverage rule of
    ``restore_properties`` (_shared.py:56-59)."""
    tgt = np.asarray(tgt)
    if tgt.size < 2:
        raise ValueError("the group-by regridders need at least two target points per axis")
    t_order = monotonic_order(tgt)
    tgt_sorted = tgt if t_order is None else tgt[t_order]
    if tgt_sorted.size != tgt.size:
        raise NotImplementedError("duplicate target coordinates")
    out_index = np.arange(tgt.size) if t_order is None else t_order

    src = np.asarray(axis.values)
    res = np.median(np.diff(tgt_sorted, 1))
    keep = (src >= tgt_sorted.min() - res) & (src <= tgt_sorted.max() + res)
    breaks = bin_breaks(tgt_sorted)
    n_out = tgt_sorted.size
    code = np.digitize(src, breaks, right=False) - 1  # closed-left membership
    code = np.where(src < breaks[0], -1, np.where(~(src < breaks[-1]), n_out, code))
    code = np.where(keep | (code < 0) | (code >= n_out), code, np.where(src < tgt_sorted.min(), -1, n_out))
    start = np.searchsorted(code, np.arange(n_out), side="left")
    stop = np.searchsorted(code, np.arange(n_out), side="right")
    kept = src[keep]
    if kept.size:
        cover = (tgt_sorted <= kept.max()) & (tgt_sorted >= kept.min())
    else:
        cover = np.zeros(n_out, bool)

    # source order as an affine map of the formatted position (ascending, descending, rotated/wrapped)
    source = axis.source.astype(np.int64)
    n_src = axis.n_src
    gather = None
    offset, step = 0, 1
    if source.size:
        ok = False
        for s in (1, -1):
            if np.array_equal(source, (source[0] + s * np.arange(source.size)) % n_src):
                offset, step, ok = int(source[0]), s, True
                break
        if not ok:
            gather, n_src = source, int(source.size)
    return Bins(start.astype(np.int32), (stop - start).astype(np.int32), cover.astype(np.uint8),
                out_index.astype(np.int32), n_src, offset, step, gather)


def identity_bins(n: int) -> Bins:
    """Every source index is its own bin (axis that is not regridded)."""
    return Bins(np.arange(n, dtype=np.int32), np.ones(n, np.int32), np.ones(n, np.uint8),
                np.arange(n, dtype=np.int32), n, 0, 1)


def plan_chunks(rows: Bins, cols: Bins, elem_size: int, extra_bytes_per_cta: int = 0,
                budget_bytes: int = 72 * 1024, hard_limit: int = 200 * 1024):
    """Target columns per CTA and the tile capacity (elements) the stat / mode kernels need."""
    max_rows = int(rows.count[rows.cover > 0].max()) if (rows.cover > 0).any() else 1
    max_rows = max(max_rows, 1)
    ends = cols.start.astype(np.int64) + cols.count
    best = None
    for c in (64, 48, 32, 24, 16, 12, 8, 6, 4, 3, 2, 1):
        l0 = np.arange(0, cols.n_out, c)
        l1 = np.minimum(l0 + c, cols.n_out) - 1
        span = int((ends[l1] - cols.start[l0]).max())
        tile = max(max_rows * max(span, 1), 1)
        total = tile * elem_size + extra_bytes_per_cta
        if total <= budget_bytes or (c == 1 and total <= hard_limit):
            best = (c, tile)
            break
    if best is None:
        raise NotImplementedError(
            "a single target cell does not fit in shared memory (bins of more than ~200 KB of source data)"
        )
    return best
