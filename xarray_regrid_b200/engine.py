"""Array-level engine: plans (device-resident tables) + launches through the C ABI.

PyTorch is used for plumbing only -- device allocations, the current stream and host<->device
copies of the KB-sized tables.  All per-cell arithmetic happens in the hand-written sm_100a
kernels behind ``include/regrid_b200.h``; there is no CPU or eager-torch fallback.

Layout convention: data is ``[..., rows, cols]`` (rows = latitude-like, cols = longitude-like),
all leading dims are flattened into the kernels' ``batch``.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, tables
from ._lib import check, lib, rg_band_t, rg_row_ring_t


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "xarray_regrid_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


_TORCH_FLOAT = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}


class DeviceBand:
    """A ``tables.Band`` uploaded to the GPU, plus the ``rg_band_t`` that points at it."""

    def __init__(self, band: tables.Band, dtype: torch.dtype, device: torch.device):
        self.host = band
        np_dtype = np.float32 if dtype == torch.float32 else np.float64
        self.idx = torch.from_numpy(np.ascontiguousarray(band.idx, dtype=np.int32)).to(device)
        # weights are cast to the data's float type like format_weights (conservative.py:281-282)
        self.w = torch.from_numpy(np.ascontiguousarray(band.w.astype(np_dtype))).to(device)
        self.cnt = torch.from_numpy(np.ascontiguousarray(band.cnt, dtype=np.int32)).to(device)
        self.cover = torch.from_numpy(np.ascontiguousarray(band.cover, dtype=np.uint8)).to(device)
        self.c = rg_band_t(
            self.idx.data_ptr(), self.w.data_ptr(), self.cnt.data_ptr(), self.cover.data_ptr(),
            band.n_out, band.k_max, band.n_src,
        )

    @property
    def ref(self):
        return C.byref(self.c)


class DeviceRing:
    """A ``tables.RowRing`` uploaded to the GPU, plus the ``rg_row_ring_t`` that points at it."""

    def __init__(self, ring: tables.RowRing, rows: tables.Band, dtype: torch.dtype, device: torch.device):
        self.host = ring

        def up(a):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(device)

        np_dtype = np.float32 if dtype == torch.float32 else np.float64
        self.sched_row, self.krec = up(ring.sched_row), up(ring.krec())
        self.wrec = torch.from_numpy(np.ascontiguousarray(ring.wrec(rows, np_dtype))).to(device)
        self.run_first, self.run_sched = up(ring.run_first), up(ring.run_sched)
        self.strip_l0, self.strip_c0, self.strip_nc = up(ring.strip_l0), up(ring.strip_c0), up(ring.strip_nc)
        self.c = rg_row_ring_t(
            self.sched_row.data_ptr(), self.krec.data_ptr(), self.wrec.data_ptr(),
            self.run_first.data_ptr(), self.run_sched.data_ptr(), self.strip_l0.data_ptr(),
            self.strip_c0.data_ptr(), self.strip_nc.data_ptr(),
            ring.n_runs, int(ring.sched_row.size), ring.max_live, ring.pad_shift,
            ring.n_strips, ring.strip_cols_max, ring.n_warps,
        )

    @property
    def ref(self):
        return C.byref(self.c)


@dataclass
class AxisSpec:
    """One regridded axis: source / target coordinate values and how the reference would
    pre-format it (``kind``: "lat", "lon" or "other", from the coordinate's NAME)."""

    src: np.ndarray
    tgt: np.ndarray
    kind: str = "other"


def _result_dtype(x: torch.Tensor) -> torch.dtype:
    """``np.result_type(float32, data.dtype)`` (conservative.py:281)."""
    if x.dtype == torch.float32 or x.dtype == torch.float16 or x.dtype == torch.bfloat16:
        return torch.float32
    if x.dtype in (torch.int8, torch.uint8, torch.int16, torch.bool):
        return torch.float32
    return torch.float64


def _as_batch(x: torch.Tensor, dtype: torch.dtype):
    """View ``[..., H, W]`` as ``[B, H, W]`` with unit column stride (copy only if needed)."""
    if x.dim() < 2:
        raise ValueError("data must have at least the two regridded dims [rows, cols]")
    if x.dtype != dtype:
        x = x.to(dtype)
    lead = x.shape[:-2]
    x3 = x.reshape((-1,) + tuple(x.shape[-2:]))
    if x3.stride(2) != 1 or (x3.shape[0] > 1 and x3.stride(0) < 0) or x3.stride(1) < x3.shape[2]:
        x3 = x3.contiguous()
    return x3, lead


class ConservativePlan:
    """Device-resident tables for ``ds.regrid.conservative`` on a fixed pair of grids.

    Mirrors what ``conservative_regrid_dataset`` prepares before ``apply_weights``
    (methods/conservative.py:119-137) and what ``format_for_regrid`` (utils.py:243-286)
    would do to the data, as index tables.  Build once, apply to any number of stacks.
    """

    def __init__(self, rows: AxisSpec, cols: AxisSpec | None, *, spherical_rows: bool | None = None,
                 device=None):
        self.device = _require_cuda(device)
        self.rows_spec, self.cols_spec = rows, cols
        if rows.kind == "lon" or (cols is not None and cols.kind == "lat"):
            raise ValueError("layout must be [..., lat, lon]: transpose the data first")
        rows_axis = tables.format_lat_axis(rows.src) if rows.kind == "lat" else tables.plain_axis(rows.src)
        if spherical_rows is None:
            spherical_rows = rows.kind == "lat"
        self.rows_band = tables.conservative_band(rows_axis, rows.tgt, spherical=spherical_rows)
        if cols is None:
            self.cols_band = None
        else:
            cols_axis = (
                tables.format_lon_axis(cols.src, cols.tgt) if cols.kind == "lon"
                else tables.plain_axis(cols.src)
            )
            self.cols_band = tables.conservative_band(cols_axis, cols.tgt, spherical=False)
        self.pole_rows = bool(self.rows_band.meta.get("pole_rows")) and bool(
            (self.rows_band.idx >= self.rows_band.n_src).any()
        )
        self._dev: dict = {}
        self._rings: dict = {}
        self.use_ring = True  # tests flip this to exercise the generic kernel

    # ------------------------------------------------------------------------------
    def _ring(self, batch: int, n_cols: int, dtype: torch.dtype):
        """Row-ring schedule sized so that (slabs x runs) fills the SMs about twice over."""
        if not self.use_ring or self.pole_rows:
            return None
        sms = max(sm_count(), 1)
        n_runs = int(min(self.rows_band.n_out, max(1, -(-2 * sms // max(batch, 1)))))
        vec = 2 if dtype == torch.float64 else 4
        key = (n_runs, n_cols, vec)
        if key not in self._rings:
            cols = self.cols_band if self.cols_band is not None else tables.identity_band(n_cols)
            host = tables.row_ring_schedule(self.rows_band, n_runs, cols, vec)
            self._rings[key] = None if host is None else DeviceRing(host, self.rows_band, dtype, self.device)
        return self._rings[key]

    def _bands(self, dtype: torch.dtype, n_cols: int):
        key = (dtype, n_cols)
        if key not in self._dev:
            cols = self.cols_band if self.cols_band is not None else tables.identity_band(n_cols)
            self._dev[key] = (DeviceBand(self.rows_band, dtype, self.device),
                              DeviceBand(cols, dtype, self.device))
        return self._dev[key]

    def _out_hw(self, n_cols: int):
        return (self.rows_band.n_out, n_cols if self.cols_band is None else self.cols_band.n_out)

    def __call__(self, x: torch.Tensor, skipna: bool = True, nan_threshold: float = 1.0,
                 out: torch.Tensor | None = None) -> torch.Tensor:
        if not 0.0 <= nan_threshold <= 1.0:
            raise ValueError("nan_threshold must be between [0, 1]]")
        if not x.is_cuda:
            raise ValueError("device path takes CUDA tensors; use apply_host() for host buffers")
        dtype = _result_dtype(x)
        x3, lead = _as_batch(x, dtype)
        B, H, W = x3.shape
        if B == 0:
            return torch.empty(tuple(x.shape[:-2]) + self._out_hw(W), dtype=dtype, device=x.device)
        rows, cols = self._bands(dtype, W)
        if H != rows.host.n_src or W != cols.host.n_src:
            raise ValueError(
                f"data is [{H}, {W}] but the plan was built for [{rows.host.n_src}, {cols.host.n_src}]"
            )
        Ho, Wo = rows.host.n_out, cols.host.n_out
        if out is None:
            y = torch.empty((B, Ho, Wo), dtype=dtype, device=x3.device)
        else:
            y = out.view(B, Ho, Wo)
            if y.dtype != dtype or not y.is_contiguous():
                raise ValueError("out must be a contiguous tensor of the result dtype")
        stream = _stream_ptr(x3.device)
        pole = None
        if self.pole_rows:
            pole = torch.empty((B, 2), dtype=dtype, device=x3.device)
            fn = lib.rg_row_nanmean_f64 if dtype == torch.float64 else lib.rg_row_nanmean_f32
            check(fn(x3.data_ptr(), pole.data_ptr(), B, H, W, x3.stride(0), x3.stride(1), stream),
                  "rg_row_nanmean")
        fn = lib.rg_conservative_f64 if dtype == torch.float64 else lib.rg_conservative_f32
        thr = tables.valid_threshold(nan_threshold)
        ring = self._ring(B, W, dtype)
        check(
            fn(x3.data_ptr(), y.data_ptr(), B, x3.stride(0) if B > 1 else H * x3.stride(1),
               x3.stride(1), Ho * Wo, rows.ref, cols.ref, int(bool(skipna)), thr,
               None if pole is None else pole.data_ptr(), None if ring is None else ring.ref, stream),
            "rg_conservative",
        )
        return y.view(tuple(lead) + (Ho, Wo))

    # ------------------------------------------------------------------------------
    def apply_host(self, x_host: torch.Tensor, skipna: bool = True, nan_threshold: float = 1.0,
                   out_host: torch.Tensor | None = None, chunk_batch: int = 32,
                   workspace: torch.Tensor | None = None) -> torch.Tensor:
        """Host-buffer path (the reference-facing call for data that lives in host memory):
        chunked H2D -> kernel -> D2H on three streams inside the C library."""
        if x_host.is_cuda:
            raise ValueError("apply_host takes host tensors")
        if x_host.dtype not in (torch.float32, torch.float64):
            x_host = x_host.to(_result_dtype(x_host))
        dtype = x_host.dtype
        x3 = x_host.reshape((-1,) + tuple(x_host.shape[-2:])).contiguous()
        lead = x_host.shape[:-2]
        B, H, W = x3.shape
        with torch.cuda.device(self.device):
            rows, cols = self._bands(dtype, W)
            Ho, Wo = rows.host.n_out, cols.host.n_out
            if out_host is None:
                out_host = torch.empty((B, Ho, Wo), dtype=dtype, pin_memory=x3.is_pinned())
            need = lib.rg_conservative_host_workspace(chunk_batch, H, W, Ho, Wo, x3.element_size())
            if workspace is None or workspace.numel() * workspace.element_size() < need:
                workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
            fn = lib.rg_conservative_host_f64 if dtype == torch.float64 else lib.rg_conservative_host_f32
            ring = self._ring(min(chunk_batch, B), W, dtype)
            check(
                fn(x3.data_ptr(), out_host.data_ptr(), B, chunk_batch, rows.ref, cols.ref,
                   int(bool(skipna)), tables.valid_threshold(nan_threshold), int(self.pole_rows),
                   None if ring is None else ring.ref,
                   workspace.data_ptr(), workspace.numel() * workspace.element_size()),
                "rg_conservative_host",
            )
        return out_host.view(tuple(lead) + (Ho, Wo))


def sm_count() -> int:
    return int(_lib.lib.rg_device_sm_count())
