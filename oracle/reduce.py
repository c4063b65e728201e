"""Oracle: most_common / least_common / stat (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, on plain numpy arrays, the reference's group-by regridders:

* bin edges            ``methods/_shared.py:12-18``   (``construct_intervals``, closed LEFT)
* source trimming      ``methods/_shared.py:92-108``  (``reduce_data_to_new_domain``)
* coverage / fill      ``methods/_shared.py:41-73``   (``restore_properties``)
* mode                 ``methods/flox_reduce.py:122-187`` (count per label, ``idxmax``/``idxmin``)
* statistics           ``methods/flox_reduce.py:40-100``

flox itself is not installed.  What flox does for these calls is restated as:
``np.digitize(coord, breaks, right=False) - 1`` per axis (closed-left membership),
a flat group code, and a per-group aggregation.  Pinned by the reference's known
answers (tests/test_reduce.py:87-110, 121-153, 202-260): counts + idxmax incl. ties,
``min``, ``var`` (ddof=0), unsorted source.  NOT pinned by any reference test
("parity unpinned"): ``sum/mean/std/max/median``, ``skipna=True``, ``least_common``
results, covered-but-empty bins (flox ``fill_value`` defaults are restated from
memory: count -> -1 as passed by the reference, sum -> 0, others -> NaN).
"""

from __future__ import annotations

import warnings

import numpy as np

from . import formatting

VALID_METHODS = ["sum", "mean", "var", "std", "median", "max", "min"]


def bin_breaks(tgt):
    """_shared.py:12-18: ``append(t, t[-1] + step) - step / 2`` with the MEDIAN step."""
    tgt = np.asarray(tgt)
    step = np.median(np.diff(tgt, n=1))
    return np.append(tgt, tgt[-1] + step) - step / 2


def bin_index(coord, breaks):
    """Closed-left membership ``breaks[i] <= c < breaks[i+1]``; -1 outside every bin."""
    idx = np.digitize(coord, breaks, right=False) - 1
    idx[(idx < 0) | (idx >= breaks.size - 1)] = -1
    return idx


def trim_to_target(data, coord, tgt_sorted, axis):
    """_shared.py:92-108: label slice ``[tmin - res, tmax + res]`` (both ends inclusive)."""
    res = np.median(np.diff(tgt_sorted, 1))
    keep = (coord >= tgt_sorted.min() - res) & (coord <= tgt_sorted.max() + res)
    idx = np.nonzero(keep)[0]
    return np.take(data, idx, axis=axis), coord[idx]


def _prepare(data, src_coords, tgt_coords, axes):
    data = np.asarray(data)
    codes = []
    nbins = []
    cover = []
    restore = []
    for src, tgt, ax in zip(src_coords, tgt_coords, axes):
        ax = ax - data.ndim if ax >= 0 else ax
        data, src = formatting.ensure_monotonic(data, np.asarray(src), ax)
        tgt = np.asarray(tgt)
        _, tgt_sorted = formatting.ensure_monotonic(None, tgt, 0)
        data, src = trim_to_target(data, src, tgt_sorted, ax)
        breaks = bin_breaks(tgt_sorted)
        codes.append(bin_index(src, breaks))
        nbins.append(tgt_sorted.size)
        if src.size:
            cover.append((tgt_sorted <= src.max()) & (tgt_sorted >= src.min()))
        else:
            cover.append(np.zeros(tgt_sorted.size, bool))
        restore.append(np.searchsorted(tgt_sorted, tgt))
    return data, codes, nbins, cover, restore


def _move_axes_last(data, axes):
    axes = [a % data.ndim for a in axes]
    return np.moveaxis(data, axes, list(range(data.ndim - len(axes), data.ndim)))


def _group_code(codes, nbins, shape):
    """Flat group id per (regridded) cell, -1 where any axis falls outside the bins."""
    grids = np.meshgrid(*codes, indexing="ij")
    flat = np.zeros(shape, dtype=np.int64)
    bad = np.zeros(shape, dtype=bool)
    for g, n in zip(grids, nbins):
        bad |= g < 0
        flat = flat * n + np.maximum(g, 0)
    flat[bad] = -1
    return flat.ravel()


def _finish(out, cover, restore, fill_value, n_axes, was_integer):
    """restore_properties (_shared.py:41-73) + reindex_like(target)."""
    for i, (mask, idx) in enumerate(zip(cover, restore)):
        ax = out.ndim - n_axes + i
        if (~mask).any():
            shape = [1] * out.ndim
            shape[ax] = mask.size
            m = mask.reshape(shape)
            if fill_value is None:
                if np.issubdtype(out.dtype, np.integer):
                    warnings.warn(
                        "No fill_value is provided; data will be cast to floating point "
                        "dtype to be able to use NaN for missing values.",
                        stacklevel=1,
                    )
                    out = out.astype(np.float64)
                out = np.where(m, out, np.nan)
            else:
                out = np.where(m, out, fill_value)
        out = np.take(out, idx, axis=ax)
    return out


def mode(data, src_coords, tgt_coords, axes, values, fill_value=None, anti_mode=False):
    """``compute_mode`` (flox_reduce.py:122-187).  Regridded axes come back LAST-positioned
    in their original places (layout preserved)."""
    data = np.asarray(data)
    if not np.issubdtype(data.dtype, np.integer):
        raise ValueError("Your input data has to be of an integer datatype for this method.")
    orig_axes = [a % data.ndim for a in axes]
    data, codes, nbins, cover, restore = _prepare(data, src_coords, tgt_coords, axes)
    work = _move_axes_last(data, orig_axes)
    batch_shape = work.shape[: work.ndim - len(axes)]
    cell_shape = work.shape[work.ndim - len(axes):]
    group = _group_code(codes, nbins, cell_shape)
    ngroups = int(np.prod(nbins))
    labels = np.asarray(values).astype(data.dtype)
    work = work.reshape((-1, group.size))
    out = np.empty((work.shape[0], ngroups), dtype=labels.dtype)
    inside = group >= 0
    for b in range(work.shape[0]):
        counts = np.empty((labels.size, ngroups), dtype=np.int64)
        for li, lab in enumerate(labels):
            sel = inside & (work[b] == lab)
            c = np.bincount(group[sel], minlength=ngroups)
            counts[li] = np.where(c > 0, c, -1)  # flox fill_value=-1 for empty groups
        pick = counts.argmin(axis=0) if anti_mode else counts.argmax(axis=0)
        out[b] = labels[pick]
    out = out.reshape(batch_shape + tuple(nbins))
    out = _finish(out, cover, restore, fill_value, len(axes), True)
    return np.moveaxis(out, list(range(out.ndim - len(axes), out.ndim)), orig_axes)


def _aggregate(x, group, ngroups, method, skipna):
    """Per-group statistic of the 1-D float array ``x`` (group == -1 is dropped)."""
    keep = group >= 0
    x = x[keep]
    g = group[keep]
    isnan = np.isnan(x)
    if skipna:
        x = x[~isnan]
        g = g[~isnan]
        poisoned = np.zeros(ngroups, bool)
    else:
        poisoned = np.bincount(g[isnan], minlength=ngroups) > 0
        x = x[~isnan]
        g = g[~isnan]
    n = np.bincount(g, minlength=ngroups)
    xd = x.astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        if method == "sum":
            out = np.bincount(g, weights=xd, minlength=ngroups)
        elif method in ("mean", "var", "std"):
            mean = np.bincount(g, weights=xd, minlength=ngroups) / n
            if method == "mean":
                out = mean
            else:
                dev = xd - mean[g]
                out = np.bincount(g, weights=dev * dev, minlength=ngroups) / n
                if method == "std":
                    out = np.sqrt(out)
        elif method in ("min", "max", "median"):
            order = np.lexsort((xd, g))
            xs = xd[order]
            start = np.concatenate([[0], np.cumsum(n)[:-1]])
            out = np.full(ngroups, np.nan)
            has = n > 0
            if method == "min":
                out[has] = xs[start[has]]
            elif method == "max":
                out[has] = xs[start[has] + n[has] - 1]
            else:
                lo = start[has] + (n[has] - 1) // 2
                hi = start[has] + n[has] // 2
                out[has] = (xs[lo] + xs[hi]) / 2
        else:
            raise ValueError(f"Invalid method. Please choose from '{VALID_METHODS}'.")
    if method != "sum":
        out = np.where(n > 0, out, np.nan)
    out = np.where(poisoned, np.nan, out)
    return out


def stat(data, src_coords, tgt_coords, axes, method, skipna=False, fill_value=None):
    """``statistic_reduce`` (flox_reduce.py:40-100)."""
    if method not in VALID_METHODS:
        raise ValueError(f"Invalid method. Please choose from '{VALID_METHODS}'.")
    data = np.asarray(data)
    orig_axes = [a % data.ndim for a in axes]
    data, codes, nbins, cover, restore = _prepare(data, src_coords, tgt_coords, axes)
    work = _move_axes_last(data, orig_axes)
    batch_shape = work.shape[: work.ndim - len(axes)]
    cell_shape = work.shape[work.ndim - len(axes):]
    group = _group_code(codes, nbins, cell_shape)
    ngroups = int(np.prod(nbins))
    work = work.reshape((-1, group.size))
    out_dtype = work.dtype if np.issubdtype(work.dtype, np.floating) else np.float64
    out = np.empty((work.shape[0], ngroups), dtype=np.float64)
    for b in range(work.shape[0]):
        out[b] = _aggregate(work[b].astype(np.float64), group, ngroups, method, skipna)
    out = out.astype(out_dtype).reshape(batch_shape + tuple(nbins))
    out = _finish(out, cover, restore, fill_value, len(axes), False)
    return np.moveaxis(out, list(range(out.ndim - len(axes), out.ndim)), orig_axes)


def regrid_latlon(data, src_lat, src_lon, tgt_lat, tgt_lon, op, **kw):
    """``da.regrid.most_common / least_common / stat`` for ``[..., lat, lon]`` data with
    coordinates named latitude / longitude: ``format_for_regrid(stats=True)``
    (regrid.py:174, 226, 266 -- lon wrap padding only, no pole rows) then the reduction."""
    data, lat, lon = formatting.format_for_regrid(
        np.asarray(data), src_lat, src_lon, tgt_lat, tgt_lon, stats=True
    )
    if op == "stat":
        return stat(data, [lat, lon], [tgt_lat, tgt_lon], [-2, -1], **kw)
    return mode(
        data, [lat, lon], [tgt_lat, tgt_lon], [-2, -1],
        anti_mode=(op == "least_common"), **kw,
    )
