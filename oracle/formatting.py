"""Oracle: pre-formatting of the source data (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, on plain numpy arrays, reference ``src/xarray_regrid/utils.py``:

* ``ensure_monotonic``  utils.py:411-421
* ``format_lat``        utils.py:289-338   (pole rows)
* ``format_lon``        utils.py:341-390   (+-360 shift, sort, wrap padding)
* ``format_for_regrid`` utils.py:243-286   (lat handler first, then lon; ``stats=True`` skips lat)

Unlike the product code (which folds all of this into index tables and never
copies the data), the oracle materialises the padded array exactly like the
reference does -- that is the point of an oracle.
"""

from __future__ import annotations

import numpy as np


def ensure_monotonic(data, coord, axis):
    """Sort ``data`` along ``axis`` by ``coord`` if needed, then drop duplicate labels.

    utils.py:411-421.  pandas' ``is_monotonic_increasing`` is non-strict; xarray's
    ``sortby`` is a stable lexsort; ``drop_duplicates`` keeps the first occurrence.
    """
    coord = np.asarray(coord)
    if coord.size > 1 and np.any(coord[1:] < coord[:-1]):
        order = np.argsort(coord, kind="stable")
        coord = coord[order]
        if data is not None:
            data = np.take(data, order, axis=axis)
    if coord.size > 1 and np.any(coord[1:] == coord[:-1]):
        _, first = np.unique(coord, return_index=True)
        first = np.sort(first)
        coord = coord[first]
        if data is not None:
            data = np.take(data, first, axis=axis)
    return data, coord


def _lon_mean(row, lon_axis_in_row):
    """xarray ``.mean(lon)``: NaN-skipping for floats, float64 result for ints."""
    if lon_axis_in_row is None:
        return row
    if np.issubdtype(row.dtype, np.floating):
        with np.errstate(invalid="ignore"):
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                return np.nanmean(row, axis=lon_axis_in_row, keepdims=True)
    return row.mean(axis=lon_axis_in_row, keepdims=True)


def format_lat(data, lat, lat_axis, lon_axis):
    """Add a row at -90 / +90 holding the longitude-mean of the edge row.

    utils.py:313-336: a pole row is added only when the edge latitude lies within
    one (maximum) grid spacing of the pole and is not the pole itself.
    """
    lat = np.asarray(lat)
    dy = np.diff(lat).max() if lat.size > 1 else np.nan
    lat_axis = lat_axis % data.ndim
    lon_axis = None if lon_axis is None else lon_axis % data.ndim
    new_lat = lat

    def pole_row(index):
        row = np.take(data, [index], axis=lat_axis)
        row = _lon_mean(row, lon_axis)
        if lon_axis is not None:
            reps = [1] * data.ndim
            reps[lon_axis] = data.shape[lon_axis]
            row = np.tile(row, reps)
        return row

    south = dy - 90 >= lat[0] > -90
    north = 90 - dy <= lat[-1] < 90
    pieces = [data]
    if south:
        pieces.insert(0, pole_row(0))
        new_lat = np.concatenate([[-90], new_lat])
    if north:
        pieces.append(pole_row(-1))
        new_lat = np.concatenate([new_lat, [90]])
    if len(pieces) > 1:
        data = np.concatenate(pieces, axis=lat_axis)  # promotes int + float64 -> float64
    return data, new_lat


def format_lon(data, lon, tgt_lon, lon_axis):
    """Shift the source longitudes next to the target's, sort, and wrap-pad a global axis.

    utils.py:358-388.
    """
    lon = np.asarray(lon)
    tgt_lon = np.asarray(tgt_lon)
    wrap_point = (tgt_lon[-1] + tgt_lon[0] + 360) / 2
    lon = np.where(lon < wrap_point - 360, lon + 360, lon)
    lon = np.where(lon > wrap_point, lon - 360, lon)
    data, lon = ensure_monotonic(data, lon, lon_axis)

    dx_s = np.diff(lon).max()
    dx_t = np.diff(tgt_lon).max()
    if lon.max() - lon.min() >= 360 - dx_s:
        left = (lon[0] - tgt_lon[0] + dx_t / 2) / dx_s
        right = (tgt_lon[-1] - lon[-1] + dx_t / 2) / dx_s
        left = int(np.ceil(max(left, 0)))
        right = int(np.ceil(max(right, 0)))
        n = lon.size
        idx = np.concatenate(
            [np.arange(n - left, n), np.arange(n), np.arange(0, right)]
        ).astype(np.intp)
        new_lon = lon[idx].astype(np.result_type(lon.dtype, np.float64), copy=True)
        if left:
            new_lon[:left] = lon[n - left :] - 360
        if right:
            new_lon[new_lon.size - right :] = lon[:right] + 360
        data = np.take(data, idx, axis=lon_axis)
        data, new_lon = ensure_monotonic(data, new_lon, lon_axis)
        lon = new_lon
    return data, lon


def format_for_regrid(
    data, lat, lon, tgt_lat, tgt_lon, lat_axis=-2, lon_axis=-1, stats=False
):
    """utils.py:243-286.  ``lat``/``lon`` are None when the data has no coordinate with
    that (case-insensitive) name, in which case the handler is skipped.

    Returns ``(data, lat, lon)`` after formatting.
    """
    if lat is not None and not stats:
        data, lat = ensure_monotonic(data, lat, lat_axis)
        data, lat = format_lat(data, lat, lat_axis, lon_axis if lon is not None else None)
    if lon is not None:
        data, lon = ensure_monotonic(data, lon, lon_axis)
        _, tgt_lon_sorted = ensure_monotonic(None, tgt_lon, 0)
        data, lon = format_lon(data, lon, tgt_lon_sorted, lon_axis)
    return data, lat, lon
