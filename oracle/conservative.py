"""Oracle: conservative regridding (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, on plain numpy arrays, the reference's conservative path:

* cell edges           ``utils.py:117-142``  (``to_intervalindex``)
* pairwise overlap     ``utils.py:145-162``  (``overlap``)
* column normalisation ``utils.py:165-169``  (``normalize_overlap``)
* weights              ``methods/conservative.py:219-233`` (``get_weights``)
* spherical correction ``methods/conservative.py:236-260``
* valid threshold      ``methods/conservative.py:211-216``
* contraction + NaN    ``methods/conservative.py:181-208`` (``apply_weights``)
* coverage mask        ``methods/conservative.py:123-125, 161-164``
* sort / reindex       ``methods/conservative.py:82-99``

The contraction is the reference's "no extras" install: DENSE weight matrices
and ``numpy.einsum(..., optimize=True)`` (what ``xr.dot`` reduces to when neither
``opt_einsum`` nor ``sparse`` is importable).
"""

from __future__ import annotations

import numpy as np

from . import formatting


# ----------------------------------------------------------------------------- weights
def cell_edges(coords):
    """Edges of the cells centred on ``coords`` (utils.py:128-140).

    Interior edges are midpoints; the two outer edges mirror the first / last
    midpoint about the first / last centre.  A single point spans all of space.
    """
    coords = np.asarray(coords)
    if coords.size <= 1:
        return np.array([-np.inf, np.inf])
    # arithmetic in the coordinate's own dtype (float32 coords -> float32 midpoints), then
    # float64: pandas' IntervalIndex.from_breaks stores float64 edges
    mid = (coords[:-1] + coords[1:]) / 2
    first = 2 * coords[0] - mid[0]
    last = 2 * coords[-1] - mid[-1]
    return np.concatenate([[first], mid, [last]]).astype(np.float64)


def overlap_matrix(src_edges, tgt_edges):
    """Length of the intersection of every (source, target) cell pair -> [n_src, n_tgt]
    (utils.py:159-162)."""
    s_lo, s_hi = src_edges[:-1], src_edges[1:]
    t_lo, t_hi = tgt_edges[:-1], tgt_edges[1:]
    # Built target-major and returned as a transposed VIEW, like the reference: the
    # memory order decides whether numpy's column sums run pairwise or sequentially,
    # i.e. the last bit of every normalised weight.
    upper = np.minimum(s_hi[None, :], t_hi[:, None])
    lower = np.maximum(s_lo[None, :], t_lo[:, None])
    return np.maximum(upper - lower, 0).T


def normalize_columns(w):
    """Each column sums to one; empty columns are divided by 1e-12 (utils.py:165-169)."""
    total = w.sum(axis=0)
    total[total == 0] = 1e-12
    return w / total


def get_weights(src, tgt):
    """conservative.py:219-233 -> dense [n_src, n_tgt] column-normalised overlaps."""
    return normalize_columns(overlap_matrix(cell_edges(src), cell_edges(tgt)))


def lat_weight(lat_deg, res_deg):
    """conservative.py:247-260: band area fraction ``(sin(top) - sin(bottom)) * dlat / 4pi``."""
    dlat = np.radians(res_deg)
    lat = np.radians(lat_deg)
    h = np.sin(lat + dlat / 2) - np.sin(lat - dlat / 2)
    return h * dlat / (np.pi * 4)


def spherical_correction(w, src_lat):
    """conservative.py:236-244: scale rows by the band area (median source spacing), renormalise."""
    res = np.median(np.diff(np.asarray(src_lat), 1))
    return normalize_columns(w * lat_weight(np.asarray(src_lat), res)[:, None])


def valid_threshold(nan_threshold):
    """conservative.py:211-216."""
    return 1 - np.clip(nan_threshold, 1e-6, 1.0 - 1e-6)


def covered_mask(src, tgt):
    """conservative.py:123-125: target CENTRES inside [min, max] of the source centres."""
    src = np.asarray(src)
    tgt = np.asarray(tgt)
    return (tgt <= src.max()) & (tgt >= src.min())


# ------------------------------------------------------------------------- contraction
def apply_weights(data, w_by_axis, skipna, nan_threshold, return_valid=False):
    """conservative.py:181-208 for 1 or 2 regridded axes.

    ``return_valid`` (checker convenience, not part of the reference): also return the valid mass
    ``valid_frac`` so that a comparison can tell apart the cells whose mass equals the threshold to
    within rounding -- there the reference's own NaN placement depends on the summation order of the
    BLAS behind ``einsum``.

    ``w_by_axis`` maps a (negative) axis of ``data`` to its dense [n_src, n_tgt]
    weights, already cast to ``result_type(float32, data.dtype)``
    (conservative.py:281).

    ``skipna=False``: the reference still contracts ``da.fillna(0)`` (conservative.py:196-198
    is outside the ``if skipna`` branches); only the division by the valid mass and the
    threshold mask are conditional (conservative.py:200-202).  A NaN therefore counts as 0
    and the result is finite, with dense and with ``sparse`` weights alike.
    """
    axes = sorted(w_by_axis)  # e.g. [-2, -1]
    nd = data.ndim
    letters = "abcdefghijklmnopqrstuvw"
    in_sub = list(letters[:nd])
    out_sub = list(in_sub)
    operands = []
    subs = []
    for n, ax in enumerate(axes):
        tgt_letter = "xyz"[n]
        subs.append(in_sub[ax] + tgt_letter)
        out_sub[ax] = tgt_letter
        operands.append(w_by_axis[ax])
    spec = "".join(in_sub) + "," + ",".join(subs) + "->" + "".join(out_sub)

    wdtype = operands[0].dtype
    notnull = ~np.isnan(data)
    filled = np.where(notnull, data, 0).astype(np.result_type(data.dtype, wdtype))

    if not skipna:
        out = np.einsum(spec, filled, *operands, optimize=True)
        return (out, np.ones_like(out)) if return_valid else out

    valid = np.einsum(spec, notnull, *operands, optimize=True)
    out = np.einsum(spec, filled, *operands, optimize=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        out = out / valid
    out = np.where(valid >= valid_threshold(nan_threshold), out, np.nan)
    return (out, valid) if return_valid else out


# ---------------------------------------------------------------------------- pipeline
def regrid(
    data,
    src_coords,
    tgt_coords,
    axes,
    latitude_index=None,
    skipna=True,
    nan_threshold=1.0,
    return_valid=False,
):
    """Everything ``conservative_regrid`` does after ``format_for_regrid``
    (conservative.py:39-178) for the regridded ``axes`` of ``data``.

    ``src_coords[i]`` / ``tgt_coords[i]`` are the 1-D coordinates for ``axes[i]``;
    ``latitude_index`` selects which of them gets the spherical correction (None: none).
    The result is labelled by ``tgt_coords`` in their ORIGINAL order.
    """
    data = np.asarray(data)
    w_by_axis = {}
    cover = {}
    restore = {}
    wdtype = np.result_type(np.float32, data.dtype)
    for i, ax in enumerate(axes):
        ax = ax - data.ndim if ax >= 0 else ax
        data, src = formatting.ensure_monotonic(data, src_coords[i], ax)
        tgt = np.asarray(tgt_coords[i])
        _, tgt_sorted = formatting.ensure_monotonic(None, tgt, 0)
        w = get_weights(src, tgt_sorted)
        if latitude_index is not None and i == latitude_index:
            w = spherical_correction(w, src)
        w_by_axis[ax] = w.astype(wdtype)
        cover[ax] = covered_mask(src, tgt_sorted)
        # reindex_like(target): pick, for every original target label, its sorted position
        restore[ax] = np.searchsorted(tgt_sorted, tgt)

    out, valid = apply_weights(data, w_by_axis, skipna, nan_threshold, return_valid=True)
    for ax, mask in cover.items():
        shape = [1] * out.ndim
        shape[ax] = mask.size
        out = np.where(mask.reshape(shape), out, np.nan)
    for ax, idx in restore.items():
        out = np.take(out, idx, axis=ax)
        valid = np.take(valid, idx, axis=ax)
    return (out, valid) if return_valid else out


def regrid_latlon(
    data, src_lat, src_lon, tgt_lat, tgt_lon, skipna=True, nan_threshold=1.0,
    latitude_weighting=True, return_valid=False,
):
    """``da.regrid.conservative(target)`` for data laid out ``[..., lat, lon]`` with
    coordinates NAMED latitude/longitude: ``format_for_regrid`` (regrid.py:123) then
    ``conservative_regrid`` (regrid.py:124-131)."""
    data, lat, lon = formatting.format_for_regrid(
        np.asarray(data), src_lat, src_lon, tgt_lat, tgt_lon
    )
    return regrid(
        data, [lat, lon], [tgt_lat, tgt_lon], [-2, -1],
        latitude_index=0 if latitude_weighting else None,
        skipna=skipna, nan_threshold=nan_threshold, return_valid=return_valid,
    )
