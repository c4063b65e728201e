"""The accessor registration path of regrid.py (only taken when xarray is importable) exercised against a
minimal stand-in for xarray: importing the package must register ``.regrid`` on DataArray / Dataset exactly
like the reference (src/xarray_regrid/regrid.py:11-12) and the labelled <-> xarray converters must round-trip.
No CUDA work happens here (argument errors are raised before any kernel is built)."""

import subprocess
import sys
import textwrap
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

STUB = textwrap.dedent(
    '''
    import numpy as np

    _ACCESSORS = {"DataArray": {}, "Dataset": {}}

    def _register(kind):
        def outer(name):
            def deco(cls):
                _ACCESSORS[kind][name] = cls
                return cls
            return deco
        return outer

    register_dataarray_accessor = _register("DataArray")
    register_dataset_accessor = _register("Dataset")

    class _Coord:
        def __init__(self, values, attrs=None):
            self.values = np.asarray(values); self.attrs = dict(attrs or {}); self.ndim = self.values.ndim

    class _Base:
        def __getattr__(self, name):
            acc = _ACCESSORS[type(self).__name__].get(name)
            if acc is None:
                raise AttributeError(name)
            return acc(self)

    class DataArray(_Base):
        def __init__(self, data, dims=(), coords=None, attrs=None, name=None):
            self.data = data; self.dims = tuple(dims); self.attrs = dict(attrs or {}); self.name = name
            self.coords = {k: (v if isinstance(v, _Coord) else _Coord(v)) for k, v in (coords or {}).items()}
        def __getitem__(self, key):
            return self.coords[key]

    class Dataset(_Base):
        def __init__(self, data_vars=None, coords=None, attrs=None):
            self.data_vars = dict(data_vars or {}); self.attrs = dict(attrs or {})
            self.coords = {k: (v if isinstance(v, _Coord) else _Coord(v)) for k, v in (coords or {}).items()}
            self.dims = tuple(self.coords)
        def __getitem__(self, key):
            return self.coords[key] if key in self.coords else self.data_vars[key]
    '''
)

SCRIPT = textwrap.dedent(
    '''
    import sys, types, importlib.machinery
    import numpy as np
    stub = types.ModuleType("xarray")
    stub.__spec__ = importlib.machinery.ModuleSpec("xarray", loader=None)
    exec(STUB_SOURCE, stub.__dict__)
    sys.modules["xarray"] = stub
    sys.path.insert(0, ROOT_DIR)
    import xarray_regrid_b200 as pkg
    from xarray_regrid_b200 import labelled
    import xarray as xr
    assert "regrid" in xr._ACCESSORS["DataArray"] and "regrid" in xr._ACCESSORS["Dataset"]
    lat, lon = np.arange(-2.0, 3.0), np.arange(0.0, 6.0)
    da = xr.DataArray(np.arange(30.0).reshape(5, 6), dims=("latitude", "longitude"),
                      coords={"latitude": xr._Coord(lat, {"units": "degrees_north"}), "longitude": lon},
                      attrs={"long_name": "t"}, name="t2m")
    lab = labelled.from_xarray(da)
    assert lab.dims == ("latitude", "longitude") and lab.attrs == {"long_name": "t"} and lab.name == "t2m"
    assert np.array_equal(lab.coords["latitude"], lat) and lab.coord_attrs["latitude"] == {"units": "degrees_north"}
    back = labelled.to_xarray(lab)
    assert back.dims == da.dims and np.array_equal(back.data, da.data) and back.name == "t2m"
    ds = xr.Dataset({"t2m": da}, coords={"latitude": lat, "longitude": lon}, attrs={"title": "x"})
    lab_ds = labelled.from_xarray(ds)
    assert list(lab_ds.data_vars) == ["t2m"] and lab_ds.attrs == {"title": "x"}
    # the accessor forwards to Regridder: an argument error of the reference surfaces before any GPU work
    target = xr.Dataset({}, coords={"latitude": np.arange(-2.0, 3.0, 2.0), "longitude": np.arange(0.0, 6.0, 2.0)})
    try:
        da.regrid.conservative(target, nan_threshold=2.0)
    except ValueError as e:
        assert "nan_threshold" in str(e)
    else:
        raise SystemExit("expected a ValueError")
    print("registration ok")
    '''
)


def test_accessor_registers_and_converters_round_trip():
    code = f"STUB_SOURCE = {STUB!r}\nROOT_DIR = {str(ROOT)!r}\n" + SCRIPT
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "registration ok" in res.stdout
