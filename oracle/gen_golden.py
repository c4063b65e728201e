"""Generate tests/golden/ref_helpers.npz by running the REFERENCE's own helpers.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Runs only in the build container,
where the reference is mounted read-only at /root/reference.  The reference
package imports ``xarray`` and ``flox`` at module level; neither is installed, but
the numeric helpers on the hot path never touch them, so they execute unchanged
behind empty stub modules.  Nothing is copied from the reference: this script
imports it, calls it on seeded inputs and stores inputs + outputs as fixtures that
travel to the GPU box (where /root/reference does not exist).

    python oracle/gen_golden.py            # rewrites tests/golden/ref_helpers.npz

Functions executed (reference file:line):
    utils.create_lat_lon_coords           src/xarray_regrid/utils.py:66-89
    utils.to_intervalindex                utils.py:117-142
    utils.overlap / normalize_overlap     utils.py:145-169
    conservative.get_weights              methods/conservative.py:219-233
    conservative.lat_weight               methods/conservative.py:247-260
    conservative.get_valid_threshold      methods/conservative.py:211-216
    _shared.construct_intervals           methods/_shared.py:12-18
"""

from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np

REF_SRC = "/root/reference/src"
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "ref_helpers.npz"


def _import_reference():
    def _decorator(_name):
        return lambda cls: cls

    xr = types.ModuleType("xarray")
    xr.DataArray = type("DataArray", (), {})
    xr.Dataset = type("Dataset", (), {})
    xr.register_dataarray_accessor = _decorator
    xr.register_dataset_accessor = _decorator
    flox = types.ModuleType("flox")
    flox_xarray = types.ModuleType("flox.xarray")
    flox.xarray = flox_xarray
    sys.modules.setdefault("xarray", xr)
    sys.modules.setdefault("flox", flox)
    sys.modules.setdefault("flox.xarray", flox_xarray)
    sys.path.insert(0, REF_SRC)
    from xarray_regrid import utils  # noqa: PLC0415
    from xarray_regrid.methods import _shared, conservative  # noqa: PLC0415

    return utils, conservative, _shared


# Coordinate pairs exercised (name, source centres, target centres).
def axis_cases():
    rng = np.random.default_rng(20261019)
    lat025 = np.arange(-90, 90.25, 0.25)
    lon025_padded = np.arange(-0.5, 359.75, 0.25)  # what format_lon yields at C1/C2
    cases = {
        "c1_lat": (lat025, np.arange(-90.0, 91.0, 1.0)),
        "c1_lon": (lon025_padded, np.arange(0.0, 360.0, 1.0)),
        "era5_to_2p2_lat": (lat025, np.arange(-90, 90.0, 2.2)),
        "cellcentred_lat": (np.arange(-89.5, 90, 1.0), np.arange(-88.0, 89.0, 2.0)),
        "upsample": (np.arange(0.0, 10.0, 2.0), np.arange(0.0, 8.5, 0.5)),
        "irregular": (
            np.sort(rng.uniform(-50, 50, 37)),
            np.sort(rng.uniform(-60, 60, 11)),
        ),
        "single_target": (np.array([-1.0, 1.0]), np.array([0.0])),
        "tgt_outside": (np.arange(0.0, 10.0, 1.0), np.arange(-5.0, 16.0, 2.5)),
    }
    return cases


def main():
    utils, conservative, shared = _import_reference()
    out = {}

    for name, (src, tgt) in axis_cases().items():
        out[f"axis/{name}/src"] = src
        out[f"axis/{name}/tgt"] = tgt
        w = conservative.get_weights(src, tgt)
        out[f"axis/{name}/weights"] = w
        ii = utils.to_intervalindex(src)
        out[f"axis/{name}/src_edges"] = np.append(ii.left.to_numpy(), ii.right.to_numpy()[-1])
        # spherical correction arithmetic of apply_spherical_correction (conservative.py:241-243)
        if np.all(np.abs(src) <= 90.5) and src.size > 1:
            res = np.median(np.diff(src, 1))
            lw = conservative.lat_weight(src, res)
            out[f"axis/{name}/lat_weight"] = lw
            out[f"axis/{name}/weights_spherical"] = utils.normalize_overlap(w * lw[:, np.newaxis])
        if tgt.size > 1:
            iv = shared.construct_intervals(tgt)
            out[f"axis/{name}/bin_breaks"] = np.append(iv.left.to_numpy(), iv.right.to_numpy()[-1])
            assert iv.closed == "left"

    thr_in = np.array([0.0, 1e-9, 0.25, 0.5, 0.75, 1.0 - 1e-9, 1.0])
    out["threshold/in"] = thr_in
    out["threshold/out"] = np.array([conservative.get_valid_threshold(t) for t in thr_in])

    grids = np.array(
        [
            # north, east, south, west, res_lat, res_lon
            [90, 360, -90, 0, 1.0, 1.0],
            [90, 359, -90, 0, 1.0, 1.0],
            [90, 359.75, -90, 0, 0.25, 0.25],
            [90, 180, 45, 90, 0.17, 0.17],
            [90, 360, -90, 0, 2.2, 2.2],
            [40, 40, 0, 0, 8, 8],
            [48, 48, -8, -8, 8, 8],
            [89, 359, -89, 1, 2.0, 2.0],
            [90, 360, -90, 0, 2.0, 1.0],
        ]
    )
    out["grid/params"] = grids
    for i, g in enumerate(grids):
        lat, lon = utils.create_lat_lon_coords(utils.Grid(*g))
        out[f"grid/{i}/lat"] = lat
        out[f"grid/{i}/lon"] = lon

    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes, {len(out)} arrays)")


if __name__ == "__main__":
    main()
