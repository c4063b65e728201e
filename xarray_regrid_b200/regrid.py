"""The ``.regrid`` accessor: same methods, arguments and errors as the reference's ``Regridder``
(src/xarray_regrid/regrid.py:11-315), executed by the CUDA engine.

Works on ``labelled.DataArray`` / ``labelled.Dataset`` everywhere, and -- when xarray is importable --
is registered as the ``regrid`` accessor of ``xarray.DataArray`` / ``xarray.Dataset`` exactly like the
reference does (regrid.py:11-12), converting through ``labelled.from_xarray`` / ``to_xarray``.

What stays on the host is label bookkeeping only: which coordinates are regridded, their names
(lat / lon formatting), dimension order, attributes.  Arrays that arrive in host memory (numpy, CPU
tensors) are streamed through the engine's persistent host pipeline in chunks along the non-regridded dims
(``engine.HostPipeline``: page-locked staging, H2D / kernels / D2H overlapped -- the stand-in for the
reference's dask chunk scheduling) and come back as numpy / CPU tensors; CUDA tensors stay on the device.
"""

from __future__ import annotations

import importlib.util
from collections import OrderedDict

import numpy as np
import torch

from . import engine, tables
from .labelled import DataArray, Dataset

_PLAN_CACHE: "OrderedDict[tuple, object]" = OrderedDict()
_PLAN_CACHE_SIZE = 16


def _kind(name) -> str:
    n = str(name).lower()
    if n in tables.LAT_NAMES:
        return "lat"
    if n in tables.LON_NAMES:
        return "lon"
    return "other"


def _key(*parts) -> tuple:
    out = []
    for p in parts:
        if isinstance(p, np.ndarray):
            out.append((p.dtype.str, p.shape, p.tobytes()))
        else:
            out.append(p)
    return tuple(out)


def _cached(key, build):
    if key in _PLAN_CACHE:
        _PLAN_CACHE.move_to_end(key)
        return _PLAN_CACHE[key]
    plan = build()
    _PLAN_CACHE[key] = plan
    while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
        _PLAN_CACHE.popitem(last=False)
    return plan


def validate_input(data, ds_target_grid: Dataset, time_dim):
    """regrid.py:289-315: drop the time coordinate from the target; require shared dims and coords."""
    if time_dim is not None and time_dim in ds_target_grid.coords:
        coords = {k: v for k, v in ds_target_grid.coords.items() if k != time_dim}
        ds_target_grid = Dataset({}, coords, dict(ds_target_grid.attrs),
                                 {k: v for k, v in ds_target_grid.coord_attrs.items() if k != time_dim})
    data_dims = set(data.dims)
    if len(data_dims.intersection(set(ds_target_grid.dims))) == 0:
        raise ValueError(
            "None of the target dims are in the data:\n"
            " regridding is not possible.\n"
            f"Target dims: {list(ds_target_grid.dims)}\n"
            f"Source dims: {list(data.dims)}"
        )
    if len(set(data.coords).intersection(set(ds_target_grid.coords))) == 0:
        raise ValueError(
            "None of the target coords are in the data:\n"
            " regridding is not possible.\n"
            f"Target coords: {list(ds_target_grid.coords)}\n"
            f"Dataset coords: {list(data.coords)}"
        )
    return ds_target_grid


def _layout(da: DataArray, names, prefer_rows=None):
    """Pick (rows, cols) among the regridded coordinate names.  A latitude-named coordinate is always
    rows and a longitude-named one always cols (the kernels' pole / wrap handling is tied to that
    layout); other coordinates fill the free slots in the data's dim order (``prefer_rows`` first)."""
    names = [n for n in names if n in da.dims]
    if not names:
        return None
    if len(names) > 2:
        raise NotImplementedError("this method regrids at most two dimensions at once")
    rows = next((n for n in names if _kind(n) == "lat"), None)
    cols = next((n for n in names if _kind(n) == "lon"), None)
    rest = sorted((n for n in names if n not in (rows, cols)),
                  key=lambda n: (n != prefer_rows, da.dims.index(n)))
    for n in rest:
        if rows is None:
            rows = n
        elif cols is None:
            cols = n
    return rows, cols


def _interp_groups(da: DataArray, names):
    """Regridded coordinate names of ``da`` in groups of at most two: lat + lon together, the rest alone."""
    names = [n for n in names if n in da.dims]
    if len(names) <= 2:
        return [names]
    pair = [n for n in names if _kind(n) in ("lat", "lon")][:2]
    rest = [n for n in names if n not in pair]
    groups = [pair] if pair else []
    return groups + [[n] for n in rest]


def _to_device(data):
    """numpy / CPU tensor -> CUDA tensor; remembers how to go back."""
    if isinstance(data, torch.Tensor):
        if data.is_cuda:
            return data, "cuda"
        return data.cuda(), "cpu-tensor"
    arr = np.asarray(data)
    if arr.dtype.byteorder not in ("=", "|"):
        arr = arr.astype(arr.dtype.newbyteorder("="))
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda(), "numpy"


def _from_device(y: torch.Tensor, how: str):
    if how == "cuda":
        return y
    if how == "cpu-tensor":
        return y.cpu()
    return y.cpu().numpy()


def _device_of(x: torch.Tensor) -> torch.device:
    return x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())


def _is_host(data) -> bool:
    return not (isinstance(data, torch.Tensor) and data.is_cuda)


def _apply(da: DataArray, target: Dataset, names, run, coord_attrs_from: str, prefer_rows=None) -> DataArray:
    """Shared driver: permute to ``[..., rows, cols]``, call ``run(x, rows_name, rows_spec, cols_spec, host)``,
    permute back, relabel.  ``run`` receives the two ``engine.AxisSpec`` (or None) and either CUDA data
    (``host=False``) or, when the data lives in host memory and already has the kernels' layout, a CPU tensor
    sharing the caller's memory (``host=True``: the plan's streaming host path)."""
    lay = _layout(da, names, prefer_rows)
    if lay is None:
        return da
    rows, cols = lay
    batch = [d for d in da.dims if d not in (rows, cols)]
    pad = None
    if cols is None and batch:  # one regridded axis (rows): let the last batch dim ride along as columns
        order = batch[:-1] + [rows, batch[-1]]
    elif cols is None:  # 1-D data: add a singleton column axis
        order, pad = [rows], -1
    elif rows is None:  # one regridded axis (cols): singleton row axis, every slab is one row
        order, pad = batch + [cols], -2
    else:
        order = batch + [rows, cols]
    perm = [da.dims.index(d) for d in order]
    spec = {n: engine.AxisSpec(np.asarray(da.coords[n]), np.asarray(target.coords[n]), _kind(n))
            for n in (rows, cols) if n is not None}
    if _is_host(da.data) and perm == list(range(len(perm))):
        # host memory in the kernels' own layout: stream chunks of the leading dims, no whole-array copy
        x = engine._host_tensor(da.data)
        how = "cpu-tensor" if isinstance(da.data, torch.Tensor) else "numpy"
        if pad is not None:
            x = x.unsqueeze(pad)
        y = run(x, rows, spec.get(rows), spec.get(cols), True)
    else:
        x, how = _to_device(da.data)
        x = x.permute(perm)
        if pad is not None:
            x = x.unsqueeze(pad)
        y = run(x, rows, spec.get(rows), spec.get(cols), False)
    if pad is not None:
        y = y.squeeze(pad)
    y = y.permute([order.index(d) for d in da.dims])
    coords = dict(da.coords)
    cattrs = dict(da.coord_attrs)
    for n in (rows, cols):
        if n is not None:
            coords[n] = np.asarray(target.coords[n])
            if coord_attrs_from == "target":
                cattrs[n] = dict(target.coord_attrs.get(n, {}))
    return DataArray(_from_device(y.contiguous(), how), da.dims, coords, dict(da.attrs), cattrs, da.name)


def _reduce_three(da: DataArray, target: Dataset, names, call2d) -> DataArray:
    """most_common / least_common / stat over THREE coordinates at once (flox bins any number of them,
    methods/flox_reduce.py:89-96, 174-183): the bins of the coordinate that is neither latitude nor longitude
    are walked on the host; for each of them the selected slices are merged into the kernels' rows axis (row i,
    slice j -> merged row i * n + j, labelled ``row_i + j * eps`` with ``eps`` far below the distance of any row
    label to the next bin edge, so that bin membership, trimming and coverage of the rows axis are unchanged) and the
    ordinary two-coordinate kernels reduce a 3-D bin as ONE 2-D bin.  Needs every bin of the third coordinate
    covered and non-empty."""
    extra = next((n for n in sorted(names, key=da.dims.index) if _kind(n) == "other"), None)
    if extra is None or len(names) != 3:
        raise NotImplementedError("the group-by methods regrid at most three dimensions, one of them neither "
                                  "latitude nor longitude")
    rows, cols = _layout(da, [n for n in names if n != extra])
    if rows is None or cols is None:
        raise NotImplementedError("three regridded coordinates need a rows and a columns coordinate")
    src_a, tgt_a = np.asarray(da.coords[extra]), np.asarray(target.coords[extra])
    axis_a = tables.plain_axis(src_a)
    bins = tables.bins_axis(axis_a, tgt_a)
    if not bins.cover.all() or (bins.count == 0).any():
        raise NotImplementedError(f"every target bin along {extra!r} must be covered by the source and non-empty")
    lab = np.asarray(da.coords[rows], dtype=np.float64)
    breaks = tables.bin_breaks(np.sort(np.asarray(target.coords[rows], dtype=np.float64)))
    gaps = breaks[None, :] - lab[:, None]
    gaps = gaps[gaps > 0]
    n_max = int(bins.count.max())
    eps = 0.25 * float(gaps.min()) / n_max if gaps.size else 1.0
    if n_max > 1 and not np.all(lab + eps != lab):
        raise NotImplementedError(f"{rows!r} labels lie too close to a bin edge to merge {extra!r} into them")
    x, how = _to_device(da.data)
    ia, ir = da.dims.index(extra), da.dims.index(rows)
    keep = [i for i in range(x.dim()) if i != ia]
    pos_r = keep.index(ir)
    perm = keep[:pos_r + 1] + [ia] + keep[pos_r + 1:]
    sub_dims = tuple(d for d in da.dims if d != extra)
    sub_target = Dataset({}, {n: np.asarray(target.coords[n]) for n in (rows, cols)}, dict(target.attrs),
                         {n: dict(target.coord_attrs.get(n, {})) for n in (rows, cols)})
    base_coords = {k: v for k, v in da.coords.items() if k != extra}
    base_cattrs = {k: v for k, v in da.coord_attrs.items() if k != extra}
    outs = [None] * bins.n_out
    for g in range(bins.n_out):
        cnt, p0 = int(bins.count[g]), int(bins.start[g])
        idx = torch.as_tensor(np.ascontiguousarray(axis_a.source[p0:p0 + cnt]).astype(np.int64), device=x.device)
        xm = x.index_select(ia, idx).permute(perm)
        shape = list(xm.shape)
        xm = xm.reshape(shape[:pos_r] + [shape[pos_r] * cnt] + shape[pos_r + 2:])
        coords = dict(base_coords)
        coords[rows] = (lab[:, None] + eps * np.arange(cnt)[None, :]).reshape(-1)
        sub = DataArray(xm, sub_dims, coords, dict(da.attrs), dict(base_cattrs), da.name)
        outs[int(bins.out_index[g])] = call2d(sub, sub_target)
    y = torch.stack([_to_device(o.data)[0] for o in outs], dim=ia)
    coords = dict(outs[0].coords)
    coords[extra] = tgt_a
    cattrs = dict(outs[0].coord_attrs)
    cattrs[extra] = dict(target.coord_attrs.get(extra, {}))
    return DataArray(_from_device(y.contiguous(), how), da.dims, coords, dict(da.attrs), cattrs, da.name)


def _conservative_nd(da: DataArray, target: Dataset, names, latitude_coord, skipna, nan_threshold) -> DataArray:
    """Conservative regridding over MORE than two coordinates at once, with the reference's joint valid mass
    (``xr.dot`` contracts all regridded dims together, methods/conservative.py:188-198): every coordinate but the
    last is folded into the kernels' rows axis through a Kronecker band (``tables.kron_band``), then the usual
    fused kernels run on ``[batch, rows..., cols]`` viewed as ``[batch, prod(rows), cols]``."""
    cols = next((n for n in names if _kind(n) == "lon"), None) or max(names, key=da.dims.index)
    folded = sorted((n for n in names if n != cols), key=da.dims.index)
    batch = [d for d in da.dims if d not in names]
    order = batch + folded + [cols]
    x, how = _to_device(da.data)
    x = x.permute([da.dims.index(d) for d in order]).contiguous()
    has_lon = cols is not None and _kind(cols) == "lon"

    def axis_band(n):
        spec = engine.AxisSpec(np.asarray(da.coords[n]), np.asarray(target.coords[n]), _kind(n))
        if _kind(n) == "lon":
            ax = tables.format_lon_axis(spec.src, spec.tgt)
        elif _kind(n) == "lat":
            ax = tables.format_lat_axis(spec.src, lon_mean=has_lon)
        else:
            ax = tables.plain_axis(spec.src)
        return tables.conservative_band(ax, spec.tgt, spherical=(n == latitude_coord))

    key = _key("conservative-nd", tuple(names), latitude_coord,
               *[np.asarray(da.coords[n]) for n in folded + [cols]],
               *[np.asarray(target.coords[n]) for n in folded + [cols]], str(x.device))

    def build():
        rows_band = axis_band(folded[0])
        for n in folded[1:]:
            rows_band = tables.kron_band(rows_band, axis_band(n))
        return engine.BandOp(rows_band, axis_band(cols), x.device)

    op = _cached(key, build)
    lead = x.shape[:len(batch)]
    n_rows = int(np.prod([x.shape[len(batch) + i] for i in range(len(folded))]))
    dtype = engine._result_dtype(x)
    x3 = x.reshape((-1, n_rows, x.shape[-1])).to(dtype)
    y = op.run(x3, engine.NAN_SKIP if skipna else engine.NAN_ZERO, tables.valid_threshold(nan_threshold))
    out_shape = tuple(lead) + tuple(len(target.coords[n]) for n in folded) + (len(target.coords[cols]),)
    y = y.reshape(out_shape).permute([order.index(d) for d in da.dims])
    coords = dict(da.coords)
    for n in names:
        coords[n] = np.asarray(target.coords[n])
    return DataArray(_from_device(y.contiguous(), how), da.dims, coords, dict(da.attrs), dict(da.coord_attrs), da.name)


def _map(obj, fn):
    """Apply ``fn`` to a DataArray or to every data variable of a Dataset (utils.py:201-224)."""
    if isinstance(obj, DataArray):
        return fn(obj)
    out = {k: fn(v) for k, v in obj.data_vars.items()}
    coords = dict(obj.coords)
    cattrs = dict(obj.coord_attrs)
    for v in out.values():
        coords.update(v.coords)
        for k, a in v.coord_attrs.items():
            cattrs[k] = a
    return Dataset(out, coords, dict(obj.attrs), cattrs)


class Regridder:
    """Regridding labelled datasets and dataarrays (mirror of regrid.py:13-270).

    Available methods: linear, nearest, cubic, conservative, most_common, least_common, stat.
    """

    def __init__(self, obj):
        self._obj = obj

    # ---------------------------------------------------------------- interpolation (regrid.py:28-83)
    def _interp(self, ds_target_grid, time_dim, method):
        target = validate_input(self._obj, ds_target_grid, time_dim)
        names = [c for c in target.coords if c in self._obj.coords]

        def one(da):
            def run(x, rows_name, rs, cs, host):
                dev = _device_of(x)
                key = _key("interp", method, None if rs is None else rs.src, None if rs is None else rs.tgt,
                           None if rs is None else rs.kind, None if cs is None else cs.src,
                           None if cs is None else cs.tgt, None if cs is None else cs.kind, str(dev))
                plan = _cached(key, lambda: engine.InterpPlan(rs, cs, method, device=dev))
                return plan.apply_host(x) if host else plan(x)

            # xr.interp interpolates one coordinate after the other (interp.py:39-46), so more than two shared
            # coordinates are a chain: the (lat, lon) pair in one fused pass, the others one by one
            out = da
            for group in _interp_groups(da, names):
                out = _apply(out, target, group, run, "data")
            return out

        return _map(self._obj, one)

    def linear(self, ds_target_grid, time_dim="time"):
        """Regrid to the coords of the target dataset with linear interpolation."""
        return self._interp(ds_target_grid, time_dim, "linear")

    def nearest(self, ds_target_grid, time_dim="time"):
        """Regrid to the coords of the target with nearest-neighbor interpolation."""
        return self._interp(ds_target_grid, time_dim, "nearest")

    def cubic(self, ds_target_grid, time_dim="time"):
        """Regrid to the coords of the target dataset with cubic interpolation."""
        return self._interp(ds_target_grid, time_dim, "cubic")

    # ------------------------------------------------------------------ conservative (regrid.py:85-131)
    def conservative(self, ds_target_grid, latitude_coord=None, time_dim="time", skipna=True,
                     nan_threshold=1.0, output_chunks=None):
        """Regrid to the coords of the target dataset with a conservative scheme.

        ``output_chunks`` is accepted for signature compatibility; the engine has no dask graph."""
        if not 0.0 <= nan_threshold <= 1.0:
            raise ValueError("nan_threshold must be between [0, 1]]")
        target = validate_input(self._obj, ds_target_grid, time_dim)
        if latitude_coord is None:  # conservative.py:75-79
            for c in self._obj.coords:
                if str(c).lower() in tables.LAT_NAMES:
                    latitude_coord = c
                    break
        names = [c for c in target.coords if c in self._obj.coords]

        def one(da):
            def run(x, rows_name, rs, cs, host):
                spherical = rows_name is not None and rows_name == latitude_coord
                if latitude_coord in names and latitude_coord in da.dims and not spherical:
                    raise NotImplementedError("the latitude coordinate must be the row axis of the kernel")
                dev = _device_of(x)
                key = _key("conservative", spherical, None if rs is None else rs.src,
                           None if rs is None else rs.tgt, None if rs is None else rs.kind,
                           None if cs is None else cs.src, None if cs is None else cs.tgt,
                           None if cs is None else cs.kind, str(dev))
                plan = _cached(key, lambda: engine.ConservativePlan(rs, cs, spherical_rows=spherical,
                                                                    device=dev))
                if host:
                    return plan.apply_host(x, skipna=skipna, nan_threshold=nan_threshold)
                return plan(x, skipna=skipna, nan_threshold=nan_threshold)

            if len([n for n in names if n in da.dims]) > 2:
                return _conservative_nd(da, target, [n for n in names if n in da.dims], latitude_coord, skipna,
                                        nan_threshold)
            return _apply(da, target, names, run, "data", prefer_rows=latitude_coord)

        return _map(self._obj, one)

    # ------------------------------------------------------- most / least common (regrid.py:133-235)
    def _mode(self, ds_target_grid, values, time_dim, fill_value, anti_mode, what):
        target = validate_input(self._obj, ds_target_grid, time_dim)
        if isinstance(self._obj, Dataset):
            raise ValueError(
                f"The '{what} common value' regridder is not implemented for\n"
                "xarray.Dataset, as it requires specifying the expected labels.\n"
                "Please select only a single variable (as DataArray),\n"
                " and regrid it separately."
            )
        names = [c for c in target.coords if c in self._obj.coords and c != time_dim]

        def run(x, rows_name, rs, cs, host):
            dev = _device_of(x)
            key = _key("reduce", None if rs is None else rs.src, None if rs is None else rs.tgt,
                       None if rs is None else rs.kind, None if cs is None else cs.src,
                       None if cs is None else cs.tgt, None if cs is None else cs.kind, str(dev))
            plan = _cached(key, lambda: engine.ReducePlan(rs, cs, device=dev))
            if host:
                return plan.mode_host(x, values, fill_value=fill_value, anti_mode=anti_mode)
            return plan.mode(x, values, fill_value=fill_value, anti_mode=anti_mode)

        if len([n for n in names if n in self._obj.dims]) == 3:
            return _reduce_three(self._obj, target, [n for n in names if n in self._obj.dims],
                                 lambda sub, tgt: Regridder(sub)._mode(tgt, values, time_dim, fill_value, anti_mode, what))
        return _apply(self._obj, target, names, run, "target")

    def most_common(self, ds_target_grid, values, time_dim="time", fill_value=None):
        """Regrid by taking the most common value within the new grid cells (ties: first of ``values``)."""
        return self._mode(ds_target_grid, values, time_dim, fill_value, False, "most")

    def least_common(self, ds_target_grid, values, time_dim="time", fill_value=None):
        """Regrid by taking the least common value within the new grid cells."""
        return self._mode(ds_target_grid, values, time_dim, fill_value, True, "least")

    # ----------------------------------------------------------------------- stat (regrid.py:237-270)
    def stat(self, ds_target_grid, method, time_dim="time", skipna=False, fill_value=None):
        """Zonal statistics: "sum", "mean", "var", "std", "median", "min" or "max" per target cell."""
        target = validate_input(self._obj, ds_target_grid, time_dim)
        if method not in engine.VALID_STAT_METHODS:
            raise ValueError(f"Invalid method. Please choose from '{engine.VALID_STAT_METHODS}'.")
        names = [c for c in target.coords if c in self._obj.coords and c != time_dim]

        def one(da):
            def run(x, rows_name, rs, cs, host):
                dev = _device_of(x)
                key = _key("reduce", None if rs is None else rs.src, None if rs is None else rs.tgt,
                           None if rs is None else rs.kind, None if cs is None else cs.src,
                           None if cs is None else cs.tgt, None if cs is None else cs.kind, str(dev))
                plan = _cached(key, lambda: engine.ReducePlan(rs, cs, device=dev))
                if host:
                    return plan.stat_host(x, method, skipna=skipna, fill_value=fill_value)
                return plan.stat(x, method, skipna=skipna, fill_value=fill_value)

            if len([n for n in names if n in da.dims]) == 3:
                return _reduce_three(da, target, [n for n in names if n in da.dims],
                                     lambda sub, tgt: Regridder(sub).stat(tgt, method, time_dim, skipna, fill_value))
            return _apply(da, target, names, run, "target")

        return _map(self._obj, one)


# ---------------------------------------------------------------------- xarray accessor registration
if importlib.util.find_spec("xarray") is not None:  # pragma: no cover - xarray is absent in this image
    import xarray as xr

    from .labelled import from_xarray, to_xarray

    @xr.register_dataarray_accessor("regrid")
    @xr.register_dataset_accessor("regrid")
    class _XarrayRegridder:
        __doc__ = Regridder.__doc__

        def __init__(self, xarray_obj):
            self._x = xarray_obj

        def __getattr__(self, name):
            method = getattr(Regridder(from_xarray(self._x)), name)

            def call(ds_target_grid, *args, **kwargs):
                return to_xarray(method(from_xarray(ds_target_grid), *args, **kwargs))

            return call
