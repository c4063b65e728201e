"""ctypes binding of the engine's C ABI (``include/regrid_b200.h``).

There is no CPU fallback: if ``csrc/libregrid_b200.so`` is missing this module raises at
import time with the command that builds it.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_CSRC = Path(__file__).resolve().parent / "csrc"
import os as _os

# (RG_LIB_PATH: another build of the same library, for A/B measurements of kernel variants)
LIB_PATH = Path(_os.environ.get("RG_LIB_PATH") or (_CSRC / "libregrid_b200.so"))

RG_ABI_VERSION = 4


class RegridB200Error(RuntimeError):
    """A C-ABI call returned a non-zero code (argument error or CUDA error)."""


class rg_band_t(C.Structure):  # noqa: N801 - mirrors the C name
    _fields_ = [
        ("idx", C.c_void_p),
        ("w", C.c_void_p),
        ("cnt", C.c_void_p),
        ("cover", C.c_void_p),
        ("n_out", C.c_int32),
        ("k_max", C.c_int32),
        ("n_src", C.c_int32),
        ("span", C.c_int32),
        ("window_rows", C.c_int32),
        ("group_span", C.c_int32),
    ]


class rg_row_ring_t(C.Structure):  # noqa: N801 - mirrors the C name
    _fields_ = [
        ("sched_row", C.c_void_p),
        ("krec", C.c_void_p),
        ("wrec", C.c_void_p),
        ("run_first", C.c_void_p),
        ("run_sched", C.c_void_p),
        ("strip_l0", C.c_void_p),
        ("strip_c0", C.c_void_p),
        ("strip_nc", C.c_void_p),
        ("n_runs", C.c_int32),
        ("n_sched", C.c_int32),
        ("max_live", C.c_int32),
        ("pad_shift", C.c_int32),
        ("n_strips", C.c_int32),
        ("strip_cols_max", C.c_int32),
        ("n_warps", C.c_int32),
        ("lane_cols", C.c_int32),
        ("panel_strip0", C.c_void_p),
        ("panel_c0", C.c_void_p),
        ("panel_nc", C.c_void_p),
        ("n_panels", C.c_int32),
        ("panel_cols_max", C.c_int32),
        ("strip_targets_max", C.c_int32),
        ("strip_b0", C.c_void_p),
        ("block_l0", C.c_void_p),
        ("halo_l", C.c_int32),
        ("halo_r", C.c_int32),
    ]


class rg_bins_t(C.Structure):  # noqa: N801 - mirrors the C name
    _fields_ = [
        ("start", C.c_void_p),
        ("count", C.c_void_p),
        ("cover", C.c_void_p),
        ("out_index", C.c_void_p),
        ("n_out", C.c_int32),
        ("n_src", C.c_int32),
        ("src_offset", C.c_int32),
        ("src_step", C.c_int32),
        ("uniform_count", C.c_int32),
    ]


RG_OK = 0
RG_ERR_UNSUPPORTED = -4  # include/regrid_b200.h
RG_HOST_NO_DRAIN = 1


def _load():
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA engine is not built and there is no CPU fallback. "
            f"Build it with `make -C {_CSRC}` (or `python -c 'import __graft_entry__ as g; g.build()'`)."
        )
    lib = C.CDLL(str(LIB_PATH))
    p, i32, i64, f32, f64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_size_t
    band = C.POINTER(rg_band_t)
    ring = C.POINTER(rg_row_ring_t)
    bins = C.POINTER(rg_bins_t)
    sig = {
        "rg_abi_version": (C.c_int, []),
        "rg_error_string": (C.c_char_p, [C.c_int]),
        "rg_device_sm_count": (C.c_int, []),
        "rg_conservative_f64": (C.c_int, [p, p, i64, i64, i64, i64, band, band, i32, f64, p, ring, p]),
        "rg_conservative_f32": (C.c_int, [p, p, i64, i64, i64, i64, band, band, i32, f32, p, ring, p]),
        "rg_row_nanmean_f64": (C.c_int, [p, p, i64, i32, i32, i64, i64, i32, i32, p]),
        "rg_row_nanmean_f32": (C.c_int, [p, p, i64, i32, i32, i64, i64, i32, i32, p]),
        "rg_separable_f64": (C.c_int, [p, p, i64, i64, i64, i64, band, band, p, ring, p]),
        "rg_separable_f32": (C.c_int, [p, p, i64, i64, i64, i64, band, band, p, ring, p]),
        "rg_nearest": (C.c_int, [p, p, i32, i32, i64, i64, i64, i64, i32, i32, p, p, p, p, f64, p]),
        "rg_spline_solve_f64": (C.c_int, [p, i64, i64, i64, i32, i32, i32, p, p, p]),
        "rg_nan_flag_f64": (C.c_int, [p, i64, i32, i32, i64, i64, p, p]),
        "rg_nan_flag_f32": (C.c_int, [p, i64, i32, i32, i64, i64, p, p]),
        "rg_poison_if_f64": (C.c_int, [p, i64, p, p]),
        "rg_poison_if_f32": (C.c_int, [p, i64, p, p]),
        "rg_mode": (C.c_int, [p, p, i32, i32, i64, i64, i64, i64, bins, bins, i32, i32, i32, i32, i32, p, p, p,
                              i32, i32, i32, f64, p]),
        "rg_stat_f64": (C.c_int, [p, p, i64, i64, i64, i64, bins, bins, i32, i32, i32, i32, i32, i32, i32, i32, f64, p]),
        "rg_stat_f32": (C.c_int, [p, p, i64, i64, i64, i64, bins, bins, i32, i32, i32, i32, i32, i32, i32, i32, f64, p]),
        "rg_conservative_host_workspace": (sz, [i64, i32, i32, i32, i32, i32, i32]),
        "rg_conservative_host_f64": (C.c_int, [p, p, i64, i64, band, band, i32, f64, i32, i32, i32, ring, p, sz,
                                               p, i32]),
        "rg_conservative_host_f32": (C.c_int, [p, p, i64, i64, band, band, i32, f32, i32, i32, i32, ring, p, sz,
                                               p, i32]),
        "rg_pipeline_create": (C.c_int, [C.POINTER(C.c_void_p), i32]),
        "rg_pipeline_destroy": (C.c_int, [p]),
        "rg_pipeline_slots": (i32, [p]),
        "rg_pipeline_stream": (C.c_void_p, [p, i32]),
        "rg_pipeline_h2d": (C.c_int, [p, i32, p, p, sz]),
        "rg_pipeline_compute_begin": (C.c_int, [p, i32]),
        "rg_pipeline_compute_end": (C.c_int, [p, i32]),
        "rg_pipeline_d2h": (C.c_int, [p, i32, p, p, sz]),
        "rg_pipeline_h2d_pageable": (C.c_int, [p, i32, p, p, sz, p, sz, i32]),
        "rg_pipeline_d2h_pageable": (C.c_int, [p, i32, p, p, sz, p, sz, i32]),
        "rg_pipeline_sync": (C.c_int, [p, i32, i32]),
        "rg_pipeline_wait_stream": (C.c_int, [p, p]),
        "rg_pipeline_drain": (C.c_int, [p]),
        "rg_host_copy": (C.c_int, [p, p, sz, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.rg_abi_version() != RG_ABI_VERSION:
        raise ImportError(
            f"{LIB_PATH} has ABI version {lib.rg_abi_version()}, expected {RG_ABI_VERSION}: rebuild it"
        )
    return lib, sig


lib, SIGNATURES = _load()


def check(code: int, what: str) -> None:
    if code != 0:
        text = lib.rg_error_string(code).decode()
        raise RegridB200Error(f"{what} failed with code {code}: {text}")
