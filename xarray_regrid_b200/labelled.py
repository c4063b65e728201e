"""A minimal labelled-array pair (DataArray / Dataset) so the accessor API works without xarray.

xarray is not installed in the build or GPU images, so the accessor mirror in ``regrid.py`` is written
against these two small classes; ``from_xarray`` / ``to_xarray`` adapt real xarray objects 1:1 when
xarray is importable (the registered ``.regrid`` accessor uses them).

Only what the regridding path needs is modelled: named dims, 1-D index coordinates, attrs on the
array / dataset / coordinates, and a name.  ``data`` is a numpy array or a torch tensor (CPU or CUDA).
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class DataArray:
    data: object
    dims: tuple
    coords: dict = field(default_factory=dict)  # name -> 1-D numpy array (index coordinate of dim `name`)
    attrs: dict = field(default_factory=dict)
    coord_attrs: dict = field(default_factory=dict)  # name -> attrs of that coordinate
    name: str | None = None

    def __post_init__(self):
        self.dims = tuple(self.dims)
        if len(self.dims) != len(self.data.shape):
            raise ValueError(f"{len(self.dims)} dims for data of shape {tuple(self.data.shape)}")
        for d, n in zip(self.dims, self.data.shape):
            if d in self.coords and len(self.coords[d]) != n:
                raise ValueError(f"coordinate {d!r} has {len(self.coords[d])} labels for a dim of size {n}")
        self.coords = {k: np.asarray(v) for k, v in self.coords.items()}

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def regrid(self):
        from .regrid import Regridder

        return Regridder(self)

    def to_dataset(self, name: str | None = None) -> "Dataset":
        nm = name or self.name or "_UNNAMED_ARRAY"
        return Dataset({nm: self}, dict(self.coords), dict(self.attrs), dict(self.coord_attrs))


@dataclass
class Dataset:
    data_vars: dict = field(default_factory=dict)  # name -> DataArray
    coords: dict = field(default_factory=dict)
    attrs: dict = field(default_factory=dict)
    coord_attrs: dict = field(default_factory=dict)

    def __post_init__(self):
        self.coords = {k: np.asarray(v) for k, v in self.coords.items()}
        for da in self.data_vars.values():
            for k, v in da.coords.items():
                self.coords.setdefault(k, v)
            for k, v in da.coord_attrs.items():
                self.coord_attrs.setdefault(k, v)

    @property
    def dims(self):
        names = list(self.coords)
        for da in self.data_vars.values():
            names += [d for d in da.dims if d not in names]
        return tuple(names)

    def __getitem__(self, key):
        if key in self.data_vars:
            return self.data_vars[key]
        if key in self.coords:
            return self.coords[key]
        raise KeyError(key)

    @property
    def regrid(self):
        from .regrid import Regridder

        return Regridder(self)


# ------------------------------------------------------------------------------- xarray adapters
def from_xarray(obj):
    """xarray.DataArray / Dataset -> labelled.DataArray / Dataset (index coordinates only)."""
    import xarray as xr

    def conv(da):
        coords = {str(d): np.asarray(da[d].values) for d in da.dims if d in da.coords}
        cattrs = {str(d): dict(da[d].attrs) for d in da.dims if d in da.coords}
        return DataArray(da.data, tuple(str(d) for d in da.dims), coords, dict(da.attrs), cattrs,
                         None if da.name is None else str(da.name))

    if isinstance(obj, xr.DataArray):
        return conv(obj)
    coords = {str(k): np.asarray(v.values) for k, v in obj.coords.items() if v.ndim == 1 and k in obj.dims}
    cattrs = {str(k): dict(obj[k].attrs) for k in coords}
    return Dataset({str(k): conv(v) for k, v in obj.data_vars.items()}, coords, dict(obj.attrs), cattrs)


def to_xarray(obj):
    """labelled.DataArray / Dataset -> xarray (data is brought to numpy)."""
    import xarray as xr

    def np_data(d):
        return d.detach().cpu().numpy() if hasattr(d, "detach") else np.asarray(d)

    def conv(da):
        out = xr.DataArray(np_data(da.data), dims=da.dims,
                           coords={k: v for k, v in da.coords.items() if k in da.dims},
                           attrs=dict(da.attrs), name=da.name)
        for k, a in da.coord_attrs.items():
            if k in out.coords:
                out[k].attrs = dict(a)
        return out

    if isinstance(obj, DataArray):
        return conv(obj)
    ds = xr.Dataset({k: conv(v) for k, v in obj.data_vars.items()},
                    coords={k: (k, v) for k, v in obj.coords.items()}, attrs=dict(obj.attrs))
    for k, a in obj.coord_attrs.items():
        if k in ds.coords:
            ds[k].attrs = dict(a)
    return ds
