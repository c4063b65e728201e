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
from ._lib import check, lib, rg_band_t, rg_bins_t, rg_row_ring_t


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
        live = np.arange(band.k_max)[:, None] < band.cnt[None, :]
        real = live & (band.idx < band.n_src)
        if real.any():
            hi = np.where(real, band.idx, -1).max(axis=0)
            lo = np.where(real, band.idx, band.n_src).min(axis=0)
            span = int(np.where(real.any(axis=0), hi - lo + 1, 0).max())
        else:
            span = 0
        self.c = rg_band_t(
            self.idx.data_ptr(), self.w.data_ptr(), self.cnt.data_ptr(), self.cover.data_ptr(),
            band.n_out, band.k_max, band.n_src, span, tables.band_window_rows(band),
            tables.band_group_span(band),
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
        self.panel_strip0, self.panel_c0, self.panel_nc = up(ring.panel_strip0), up(ring.panel_c0), up(ring.panel_nc)
        blocks = ring.block_l0 is not None
        self.strip_b0 = up(ring.strip_b0) if blocks else None
        self.block_l0 = up(ring.block_l0) if blocks else None
        self.c = rg_row_ring_t(
            self.sched_row.data_ptr(), self.krec.data_ptr(), self.wrec.data_ptr(),
            self.run_first.data_ptr(), self.run_sched.data_ptr(), self.strip_l0.data_ptr(),
            self.strip_c0.data_ptr(), self.strip_nc.data_ptr(),
            ring.n_runs, int(ring.sched_row.size), ring.max_live, ring.pad_shift,
            ring.n_strips, ring.strip_cols_max, ring.n_warps, ring.lane_cols,
            self.panel_strip0.data_ptr(), self.panel_c0.data_ptr(), self.panel_nc.data_ptr(),
            ring.n_panels, ring.panel_cols_max, int(np.diff(ring.strip_l0).max()),
            self.strip_b0.data_ptr() if blocks else None, self.block_l0.data_ptr() if blocks else None,
            int(ring.halo_l), int(ring.halo_r),
        )

    @property
    def ref(self):
        return C.byref(self.c)


# ------------------------------------------------------------------ host <-> device streaming
class HostPipeline:
    """A persistent ``rg_pipeline_t`` (three streams + per-slot events, created once) with the per-slot device
    buffers and page-locked staging buffers that the host-buffer paths of the plans reuse from call to call.

    It stands in for the reference's dask chunk scheduling (methods/conservative.py:263-302): a stack that
    lives in host memory is cut into chunks along the flattened batch dimension; chunk ``i`` uses slot
    ``i % slots``, so the H2D copy of chunk ``i + 1``, the kernels of chunk ``i`` and the D2H copy of chunk
    ``i - 1`` overlap.  Page-locked sources / results are copied directly (one ``cudaMemcpyAsync`` per chunk
    and direction); pageable ones (numpy arrays) go through page-locked staging chunks filled / emptied by
    the library's multi-threaded ``rg_host_copy``.
    """

    def __init__(self, device=None, slots: int = 3, copy_threads: int | None = None):
        self.device = _require_cuda(device)
        self.slots = int(slots)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.rg_pipeline_create(C.byref(handle), self.slots), "rg_pipeline_create")
        self.handle = handle
        import os

        # staging memcpy threads: most of the host's cores (the copies are what bounds the pageable path)
        self.copy_threads = int(copy_threads or max(1, min(24, (os.cpu_count() or 2) - 2)))
        self.stage_bytes = 4 * (32 << 20)  # staging ring per direction: 4 pieces of 32 MiB (measured best: 8 MiB
        # pieces leave the per-piece event / launch latency exposed, 128 MiB ones overlap less)
        self._dev: dict = {}     # name -> per-slot device buffers (uint8)
        self._pinned: dict = {}  # direction -> page-locked staging ring (uint8)
        self.compute_stream = torch.cuda.ExternalStream(lib.rg_pipeline_stream(handle, 1), device=self.device)

    # -- buffers (grown on demand, reused afterwards)
    def device_buffers(self, name: str, nbytes: int, count: int | None = None):
        """Per-slot device buffers of at least ``nbytes`` (only the first ``count`` slots are needed)."""
        count = self.slots if count is None else max(1, min(count, self.slots))
        cur = self._dev.get(name)
        if cur is None or len(cur) < count or cur[0].numel() < nbytes:
            size = max(nbytes, 1 if cur is None else cur[0].numel())
            n = count if cur is None else max(count, len(cur))
            self._dev.pop(name, None)
            cur = [torch.empty(size, dtype=torch.uint8, device=self.device) for _ in range(n)]
            self._dev[name] = cur
        return cur

    def device_block(self, name: str, nbytes: int) -> torch.Tensor:
        """One device buffer (not per slot), grown on demand."""
        cur = self._dev.get(("block", name))
        if cur is None or cur.numel() < nbytes:
            self._dev.pop(("block", name), None)
            cur = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.device)
            self._dev[("block", name)] = cur
        return cur

    def staging(self, name: str) -> torch.Tensor:
        """Page-locked staging ring of one direction (4 pieces, see rg_pipeline_h2d_pageable)."""
        if name not in self._pinned:
            self._pinned[name] = torch.empty(self.stage_bytes, dtype=torch.uint8, pin_memory=True)
        return self._pinned[name]

    def wait_current_stream(self):
        """Order the pipeline after the work queued so far on torch's current stream."""
        check(lib.rg_pipeline_wait_stream(self.handle, _stream_ptr(self.device)), "rg_pipeline_wait_stream")

    def drain(self):
        check(lib.rg_pipeline_drain(self.handle), "rg_pipeline_drain")

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            lib.rg_pipeline_destroy(self.handle)
            self.handle = C.c_void_p()
        self._dev.clear()
        self._pinned.clear()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # -- the chunk loop shared by every host-buffer path
    def stream(self, x_host: torch.Tensor, out_host: torch.Tensor, chunk: int, compute, drain: bool = True):
        """``x_host`` ``[B, ...]`` and ``out_host`` ``[B, ...]`` are contiguous CPU tensors; for every chunk of
        ``chunk`` leading items ``compute(x_dev, y_dev)`` is called on the pipeline's compute stream with the
        slot's device views and must queue kernels that fill ``y_dev``.

        Page-locked tensors are copied directly, one ``cudaMemcpyAsync`` per chunk; pageable ones pass through
        the staging rings.  A pageable result makes the host wait for chunk ``i - 1`` AFTER chunk ``i`` has been
        queued, so the device never idles on the host's copy-out."""
        B = x_host.shape[0]
        if B == 0:
            return out_host
        chunk = int(max(1, min(chunk, B)))
        n_chunks = -(-B // chunk)
        in_item = x_host[0].numel() * x_host.element_size()
        out_item = out_host[0].numel() * out_host.element_size()
        d_in = self.device_buffers("in", chunk * in_item, n_chunks)
        d_out = self.device_buffers("out", chunk * out_item, n_chunks)
        stage_in = None if x_host.is_pinned() else self.staging("in")
        stage_out = None if out_host.is_pinned() else self.staging("out")
        h = self.handle
        nt = self.copy_threads

        def copy_out(slot, b0, nb):
            dst = out_host[b0:b0 + nb]
            if stage_out is None:
                check(lib.rg_pipeline_d2h(h, slot, dst.data_ptr(), d_out[slot].data_ptr(), nb * out_item),
                      "rg_pipeline_d2h")
            else:
                check(lib.rg_pipeline_d2h_pageable(h, slot, dst.data_ptr(), d_out[slot].data_ptr(), nb * out_item,
                                                   stage_out.data_ptr(), stage_out.numel(), nt),
                      "rg_pipeline_d2h_pageable")

        with torch.cuda.device(self.device):
            self.wait_current_stream()
            try:
                previous = None
                for i, b0 in enumerate(range(0, B, chunk)):
                    s = i % self.slots
                    nb = min(chunk, B - b0)
                    src = x_host[b0:b0 + nb]
                    if stage_in is None:
                        check(lib.rg_pipeline_h2d(h, s, d_in[s].data_ptr(), src.data_ptr(), nb * in_item),
                              "rg_pipeline_h2d")
                    else:
                        check(lib.rg_pipeline_h2d_pageable(h, s, d_in[s].data_ptr(), src.data_ptr(), nb * in_item,
                                                           stage_in.data_ptr(), stage_in.numel(), nt),
                              "rg_pipeline_h2d_pageable")
                    check(lib.rg_pipeline_compute_begin(h, s), "rg_pipeline_compute_begin")
                    x_dev = d_in[s][:nb * in_item].view(x_host.dtype).view((nb,) + tuple(x_host.shape[1:]))
                    y_dev = d_out[s][:nb * out_item].view(out_host.dtype).view((nb,) + tuple(out_host.shape[1:]))
                    with torch.cuda.stream(self.compute_stream):
                        compute(x_dev, y_dev)
                    check(lib.rg_pipeline_compute_end(h, s), "rg_pipeline_compute_end")
                    if stage_out is None:
                        copy_out(s, b0, nb)
                    else:
                        if previous is not None:
                            copy_out(*previous)
                        previous = (s, b0, nb)
                if previous is not None:
                    copy_out(*previous)
            except BaseException:
                lib.rg_pipeline_drain(h)  # nothing may stay in flight into the caller's buffers
                raise
            if drain:
                self.drain()
        return out_host


# Results of the host paths are allocated page-locked even when the input is pageable (numpy): the D2H copies
# then need no staging, and a fresh pageable result would pay a first-touch page fault per 4 KB while it is
# filled (measured: 77 ms for a 267 MB result, more than the regrid itself).  torch's host allocator caches the
# blocks, so only the first result of a size pays for the page-locking.
PINNED_RESULTS = True

_PIPELINES: dict = {}


def default_pipeline(device) -> HostPipeline:
    """One shared pipeline per device for callers that do not manage their own."""
    device = _require_cuda(device)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _PIPELINES:
        _PIPELINES[key] = HostPipeline(device)
    return _PIPELINES[key]


def _host_tensor(x) -> torch.Tensor:
    """numpy array / CPU tensor -> contiguous CPU tensor sharing memory where possible."""
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            raise ValueError("host path takes host arrays; call the plan directly for CUDA tensors")
        return x if x.is_contiguous() else x.contiguous()
    arr = np.asarray(x)
    if arr.dtype.byteorder not in ("=", "|"):
        arr = arr.astype(arr.dtype.newbyteorder("="))
    return torch.from_numpy(np.ascontiguousarray(arr))


def _check_out_host(out_host: torch.Tensor, shape: tuple, dtype: torch.dtype) -> torch.Tensor:
    if not isinstance(out_host, torch.Tensor) or out_host.is_cuda:
        raise ValueError("out_host must be a CPU tensor")
    if out_host.dtype != dtype:
        raise ValueError(f"out_host has dtype {out_host.dtype}, the result is {dtype}")
    if not out_host.is_contiguous():
        raise ValueError("out_host must be contiguous")
    n = 1
    for d in shape:
        n *= int(d)
    if out_host.numel() != n:
        raise ValueError(f"out_host has {out_host.numel()} elements, the result has {n} ({tuple(shape)})")
    return out_host.view(shape)


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


TASKS_PER_SM = 32  # (slab, run, panel) tasks of the row-ring kernels per SM, see BandOp._ring
MIN_RUN_ROWS = 16

# What a NaN in the source does in ``BandOp.run``:
NAN_PROPAGATE = 0  # poisons every output it touches (interpolation; rg_separable_*)
NAN_SKIP = 1       # apply_weights(skipna=True): valid-mass normalisation + threshold
NAN_ZERO = 2       # apply_weights(skipna=False): the reference contracts da.fillna(0) (conservative.py:196-198)


class BandOp:
    """Two per-axis bands resident on the device + the launch of the separable band kernels.

    ``run`` computes ``y[b,k,l] = sum_t sum_c wr[t,k] wc[c,l] x[b, ri[t,k], ci[c,l]]`` with the
    conservative NaN bookkeeping (``NAN_SKIP`` / ``NAN_ZERO``) or plain NaN propagation
    (``NAN_PROPAGATE``); it is the common back end of the conservative, linear and cubic regridders.
    """

    def __init__(self, rows_band: tables.Band | None, cols_band: tables.Band | None, device):
        self.device = _require_cuda(device)
        self.rows_band, self.cols_band = rows_band, cols_band  # None = axis is not regridded (identity)
        self.pole_rows = rows_band is not None and bool(rows_band.meta.get("pole_rows")) and bool(
            (rows_band.idx >= rows_band.n_src).any()
        )
        # original rows whose longitude-mean fills the south / north pole row (sorted-first / sorted-last)
        self.pole_src = (0, 0) if rows_band is None else tuple(rows_band.meta.get("pole_src", (0, 0)))
        self._dev: dict = {}
        self._rings: dict = {}
        self.use_ring = True  # tests flip this to exercise the generic kernel
        self.use_lanes = True  # ... and this to exercise the two-phase ring kernel where the lanes variant applies

    def _host_bands(self, n_rows: int, n_cols: int):
        rows = self.rows_band if self.rows_band is not None else tables.identity_band(n_rows)
        cols = self.cols_band if self.cols_band is not None else tables.identity_band(n_cols)
        return rows, cols

    def _ring(self, batch: int, n_rows: int, n_cols: int, dtype: torch.dtype, pairs: bool = True):
        """Row-ring schedule sized so that (slabs x runs x panels) gives every SM about TASKS_PER_SM tasks.  ``pairs``: allow the
        two-targets-per-lane layout of the register-resident kernel (conservative NaN modes only)."""
        if not self.use_ring:
            return None
        rows, cols = self._host_bands(n_rows, n_cols)
        sms = max(sm_count(), 1)
        # (slab, run, panel) tasks are dealt round-robin to 1 - 4 persistent CTAs per SM: enough of them that the
        # last, partial round does not matter (256 slabs x 2 runs on 148 CTAs = 3.46 rounds ran at 86 %), but runs of
        # at least MIN_RUN_ROWS target rows (a run re-reads the k_max - 1 source rows it shares with its neighbour)
        def runs_for(n_panels):
            want = -(-TASKS_PER_SM * sms // max(batch * n_panels, 1))
            return int(min(max(1, rows.n_out // MIN_RUN_ROWS), max(1, want)))

        n_runs = runs_for(1)
        vec = 2 if dtype == torch.float64 else 4
        key = (n_runs, n_rows, n_cols, vec, self.use_lanes, pairs)
        if key not in self._rings:
            host = tables.row_ring_schedule(rows, n_runs, cols, vec, allow_pairs=pairs and self.use_lanes)
            if host is not None and host.n_panels > 1 and runs_for(host.n_panels) != n_runs:
                host = tables.row_ring_schedule(rows, runs_for(host.n_panels), cols, vec,
                                                allow_pairs=pairs and self.use_lanes)
            if host is not None and not self.use_lanes:
                host.lane_cols = 0
            self._rings[key] = None if host is None else DeviceRing(host, rows, dtype, self.device)
        return self._rings[key]

    def _bands(self, dtype: torch.dtype, n_rows: int, n_cols: int):
        key = (dtype, n_rows, n_cols)
        if key not in self._dev:
            rows, cols = self._host_bands(n_rows, n_cols)
            self._dev[key] = (DeviceBand(rows, dtype, self.device), DeviceBand(cols, dtype, self.device))
        return self._dev[key]

    def _out_hw(self, n_cols: int, n_rows: int | None = None):
        return (n_rows if self.rows_band is None else self.rows_band.n_out,
                n_cols if self.cols_band is None else self.cols_band.n_out)

    def run(self, x3: torch.Tensor, nan_mode: int, thr: float, out: torch.Tensor | None = None,
            pole: torch.Tensor | None = None) -> torch.Tensor:
        """``x3``: CUDA ``[B, H, W]`` float32/float64 with unit column stride -> ``[B, Ho, Wo]``.
        ``pole``: optional caller-owned ``[B, 2, W]`` scratch for the synthetic pole rows."""
        dtype = x3.dtype
        B, H, W = x3.shape
        rows, cols = self._bands(dtype, H, W)
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
        if B == 0:
            return y
        stream = _stream_ptr(x3.device)
        if not self.pole_rows:
            pole = None
        else:
            if pole is None:
                pole = torch.empty((B, 2, W), dtype=dtype, device=x3.device)  # full rows: rg_row_nanmean_*
            fn = lib.rg_row_nanmean_f64 if dtype == torch.float64 else lib.rg_row_nanmean_f32
            check(fn(x3.data_ptr(), pole.data_ptr(), B, H, W, x3.stride(0), x3.stride(1),
                     int(self.pole_src[0]), int(self.pole_src[1]), stream), "rg_row_nanmean")
        ring = self._ring(B, H, W, dtype, pairs=nan_mode != NAN_PROPAGATE)
        xbs = x3.stride(0) if B > 1 else H * x3.stride(1)
        pole_ptr = None if pole is None else pole.data_ptr()
        ring_ref = None if ring is None else ring.ref
        if nan_mode == NAN_PROPAGATE:
            fn = lib.rg_separable_f64 if dtype == torch.float64 else lib.rg_separable_f32
            check(fn(x3.data_ptr(), y.data_ptr(), B, xbs, x3.stride(1), Ho * Wo, rows.ref, cols.ref,
                     pole_ptr, ring_ref, stream), "rg_separable")
        else:
            fn = lib.rg_conservative_f64 if dtype == torch.float64 else lib.rg_conservative_f32
            check(fn(x3.data_ptr(), y.data_ptr(), B, xbs, x3.stride(1), Ho * Wo, rows.ref, cols.ref,
                     int(nan_mode == NAN_SKIP), thr, pole_ptr, ring_ref, stream), "rg_conservative")
        return y


def _axis(spec: AxisSpec | None, is_rows: bool, lon_mean: bool = True):
    """``lon_mean``: whether the data has a lon / longitude coordinate, i.e. whether the synthetic pole rows
    are the longitude-mean of the edge rows (utils.py:322-323) or plain copies of them."""
    if spec is None:
        return None
    if is_rows and spec.kind == "lon" or (not is_rows and spec.kind == "lat"):
        raise ValueError("layout must be [..., lat, lon]: transpose the data first")
    if spec.kind == "lat":
        return tables.format_lat_axis(spec.src, lon_mean=lon_mean)
    if spec.kind == "lon":
        return tables.format_lon_axis(spec.src, spec.tgt)
    return tables.plain_axis(spec.src)


class ConservativePlan(BandOp):
    """Device-resident tables for ``ds.regrid.conservative`` on a fixed pair of grids.

    Mirrors what ``conservative_regrid_dataset`` prepares before ``apply_weights``
    (methods/conservative.py:119-137) and what ``format_for_regrid`` (utils.py:243-286)
    would do to the data, as index tables.  Build once, apply to any number of stacks.
    """

    def __init__(self, rows: AxisSpec | None, cols: AxisSpec | None, *, spherical_rows: bool | None = None,
                 device=None):
        self.rows_spec, self.cols_spec = rows, cols
        if rows is None and cols is None:
            raise ValueError("at least one axis must be regridded")
        if spherical_rows is None:
            spherical_rows = rows is not None and rows.kind == "lat"
        lon_mean = cols is not None and cols.kind == "lon"
        rows_band = None if rows is None else tables.conservative_band(_axis(rows, True, lon_mean), rows.tgt,
                                                                      spherical=spherical_rows)
        cols_band = None if cols is None else tables.conservative_band(_axis(cols, False), cols.tgt)
        super().__init__(rows_band, cols_band, device)

    def __call__(self, x: torch.Tensor, skipna: bool = True, nan_threshold: float = 1.0,
                 out: torch.Tensor | None = None) -> torch.Tensor:
        if not 0.0 <= nan_threshold <= 1.0:
            raise ValueError("nan_threshold must be between [0, 1]]")
        if not x.is_cuda:
            raise ValueError("device path takes CUDA tensors; use apply_host() for host buffers")
        x3, lead = _as_batch(x, _result_dtype(x))
        y = self.run(x3, NAN_SKIP if skipna else NAN_ZERO, tables.valid_threshold(nan_threshold), out)
        return y.view(tuple(lead) + tuple(y.shape[-2:]))

    # ------------------------------------------------------------------------------
    def apply_host(self, x_host, skipna: bool = True, nan_threshold: float = 1.0,
                   out_host: torch.Tensor | None = None, chunk_batch: int = 32,
                   workspace: torch.Tensor | None = None, pipeline: HostPipeline | None = None,
                   drain: bool = True) -> torch.Tensor:
        """Host-buffer path (the reference-facing call for data that lives in host memory): chunked
        H2D -> kernel -> D2H on the three streams of a persistent pipeline.

        Page-locked source AND result: one C call, ``rg_conservative_host_*``.  Pageable memory (numpy):
        the same chunk loop driven through ``HostPipeline.stream`` with page-locked staging chunks.
        ``drain=False`` (page-locked buffers, explicit ``pipeline``) returns with copies still in flight;
        finish with ``pipeline.drain()``."""
        if not 0.0 <= nan_threshold <= 1.0:
            raise ValueError("nan_threshold must be between [0, 1]]")
        x_host = _host_tensor(x_host)
        if x_host.dtype not in (torch.float32, torch.float64):
            x_host = x_host.to(_result_dtype(x_host))
        dtype = x_host.dtype
        if x_host.dim() < 2:
            raise ValueError("data must have at least the two regridded dims [rows, cols]")
        x3 = x_host.reshape((-1,) + tuple(x_host.shape[-2:]))
        lead = x_host.shape[:-2]
        B, H, W = x3.shape
        chunk_batch = int(max(1, chunk_batch))
        thr = tables.valid_threshold(nan_threshold)
        with torch.cuda.device(self.device):
            rows, cols = self._bands(dtype, H, W)
            if H != rows.host.n_src or W != cols.host.n_src:
                raise ValueError(
                    f"data is [{H}, {W}] but the plan was built for [{rows.host.n_src}, {cols.host.n_src}]")
            Ho, Wo = rows.host.n_out, cols.host.n_out
            if out_host is None:
                out_host = torch.empty((B, Ho, Wo), dtype=dtype, pin_memory=x3.is_pinned() or PINNED_RESULTS)
            else:
                out_host = _check_out_host(out_host, (B, Ho, Wo), dtype)
            if B == 0:
                return out_host.view(tuple(lead) + (Ho, Wo))
            pipe = pipeline if pipeline is not None else default_pipeline(self.device)
            if x3.is_pinned() and out_host.is_pinned():
                need = lib.rg_conservative_host_workspace(chunk_batch, H, W, Ho, Wo, x3.element_size(), pipe.slots)
                if workspace is None:
                    workspace = pipe.device_block("conservative", need)  # one block holding every slot
                if workspace.is_cuda is False or workspace.numel() * workspace.element_size() < need:
                    raise ValueError(f"workspace must be a CUDA buffer of at least {need} bytes")
                fn = lib.rg_conservative_host_f64 if dtype == torch.float64 else lib.rg_conservative_host_f32
                ring = self._ring(min(chunk_batch, B), H, W, dtype)
                # the workspace and the host buffers may have been produced on torch's current stream
                pipe.wait_current_stream()
                check(
                    fn(x3.data_ptr(), out_host.data_ptr(), B, chunk_batch, rows.ref, cols.ref,
                       int(bool(skipna)), thr, int(self.pole_rows), int(self.pole_src[0]), int(self.pole_src[1]),
                       None if ring is None else ring.ref, workspace.data_ptr(),
                       workspace.numel() * workspace.element_size(), pipe.handle,
                       0 if drain else _lib.RG_HOST_NO_DRAIN),
                    "rg_conservative_host",
                )
            else:
                mode = NAN_SKIP if skipna else NAN_ZERO
                pole = (pipe.device_buffers("pole", min(chunk_batch, B) * 2 * W * x3.element_size(),
                                            -(-B // chunk_batch)) if self.pole_rows else None)
                state = {"i": 0}

                def compute(x_dev, y_dev):
                    pb = None
                    if pole is not None:
                        pb = pole[state["i"] % len(pole)][: x_dev.shape[0] * 2 * W * x3.element_size()]
                        pb = pb.view(dtype).view(x_dev.shape[0], 2, W)
                    state["i"] += 1
                    self.run(x_dev, mode, thr, out=y_dev, pole=pb)

                pipe.stream(x3, out_host, chunk_batch, compute)
        return out_host.view(tuple(lead) + (Ho, Wo))


def _out_or_new(out, shape: tuple, dtype: torch.dtype, device) -> torch.Tensor:
    """The caller's result buffer (validated) viewed as ``shape``, or a fresh tensor."""
    if out is None:
        return torch.empty(shape, dtype=dtype, device=device)
    n = 1
    for d in shape:
        n *= int(d)
    if out.dtype != dtype or not out.is_contiguous() or out.numel() != n or out.device != torch.device(device):
        raise ValueError(f"out must be a contiguous {dtype} tensor of {n} elements on {device}")
    return out.view(shape)


_TYPE_CODE = {
    torch.uint8: 0, torch.int8: 1, torch.int16: 2, torch.int32: 3, torch.int64: 4,
    torch.float32: 5, torch.float64: 6,
}


class InterpPlan:
    """Device-resident tables for ``ds.regrid.linear / nearest / cubic`` on a fixed pair of grids.

    The reference does ``data.interp(coords, method)`` (methods/interp.py:43-46), which xarray turns into
    one scipy ``interp1d`` per axis.  Here every method is a separable band:

    * linear  -> 2-tap bands, one fused pass;
    * nearest -> an index gather (integers stay exact; no arithmetic);
    * cubic   -> prefilter bands on the source grid (A^-1 of the not-a-knot collocation system), then 4-tap
      B-spline bands onto the target grid, plus scipy's "any NaN -> everything NaN" rule.

    Result dtype follows the reference by default (``out_dtype="reference"``): float64 for linear / cubic
    and for nearest on integer data, the input's dtype for nearest on floats.  ``out_dtype="same"`` keeps
    float32 for linear and the integer dtype for nearest (``fill`` marks cells outside the source range).
    """

    def __init__(self, rows: AxisSpec | None, cols: AxisSpec | None, method: str, *, device=None):
        if method not in ("linear", "nearest", "cubic"):
            raise ValueError(f"unknown interpolation method {method!r}")
        if rows is None and cols is None:
            raise ValueError("at least one axis must be regridded")
        self.method = method
        self.device = _require_cuda(device)
        rows_axis = _axis(rows, True, cols is not None and cols.kind == "lon")
        cols_axis = _axis(cols, False)
        self.rows_axis, self.cols_axis = rows_axis, cols_axis
        if method == "linear":
            self.op = BandOp(None if rows is None else tables.linear_band(rows_axis, rows.tgt),
                             None if cols is None else tables.linear_band(cols_axis, cols.tgt), self.device)
        elif method == "nearest":
            self.op = BandOp(None if rows is None else tables.nearest_band(rows_axis, rows.tgt),
                             None if cols is None else tables.nearest_band(cols_axis, cols.tgt), self.device)
            self._gather: dict = {}
        else:
            r_pre, r_ev = (None, None) if rows is None else tables.cubic_bands(rows_axis, rows.tgt)
            c_pre, c_ev = (None, None) if cols is None else tables.cubic_bands(cols_axis, cols.tgt)
            self.pre = BandOp(r_pre, c_pre, self.device)
            self.op = BandOp(r_ev, c_ev, self.device)
            self.chunk = 2048  # slabs per pass: bounds the fp64 coefficient buffer
            # xr.interp interpolates one axis after the other; a row target outside the source range
            # leaves NaN rows, and scipy's cubic turns ANY NaN in its input into an all-NaN result, so
            # the second (column) pass blanks everything.  (The reference's axis order is the iteration
            # order of a Python set, interp.py:39; the engine and the oracle fix it to rows, then cols.)
            self.all_nan = rows is not None and cols is not None and not bool(r_ev.cover.all())
            # O(n) prefilter: banded LU substitution per line (rg_spline_solve_f64) instead of the +-31-tap
            # truncated inverse, when both formatted axes are plain index maps of the source (no pole rows)
            self.use_lu = True  # tests flip this to exercise the band prefilter
            self._lu = {}
            for name, spec, axis in (("rows", rows, rows_axis), ("cols", cols, cols_axis)):
                if spec is None:
                    self._lu[name] = None
                    continue
                fac = None if axis.has_pole_rows else tables.cubic_lu(axis)
                if fac is None:
                    self._lu = None
                    break
                self._lu[name] = (
                    torch.from_numpy(np.ascontiguousarray(axis.source, dtype=np.int32)).to(self.device),
                    torch.from_numpy(np.ascontiguousarray(fac[0])).to(self.device),
                    torch.from_numpy(np.ascontiguousarray(fac[1])).to(self.device),
                )
            self._lu_maps: dict = {}

    def _coefficients_lu(self, xs: torch.Tensor, Hf: int, Wf: int):
        """Spline coefficients on the formatted grid: gather (sort / wrap padding) + in-place banded solves.
        Returns None when the library declines (factors larger than shared memory)."""
        B, H, W = xs.shape
        if (H, W) not in self._lu_maps:
            def index_map(entry, n):
                if entry is not None:
                    return entry[0]
                return torch.arange(n, dtype=torch.int32, device=self.device)
            ri, ci = index_map(self._lu["rows"], H), index_map(self._lu["cols"], W)
            self._lu_maps[(H, W)] = (ri, ci, torch.ones(ri.numel(), dtype=torch.uint8, device=self.device),
                                     torch.ones(ci.numel(), dtype=torch.uint8, device=self.device))
        ri, ci, rok, cok = self._lu_maps[(H, W)]
        if ri.numel() != Hf or ci.numel() != Wf:
            raise ValueError(f"data is [{H}, {W}] but the plan's formatted grid is [{Hf}, {Wf}]")
        if xs.stride(2) != 1:
            xs = xs.contiguous()
        stream = _stream_ptr(xs.device)
        coeff = torch.empty((B, Hf, Wf), dtype=torch.float64, device=xs.device)
        check(lib.rg_nearest(xs.data_ptr(), coeff.data_ptr(), _TYPE_CODE[xs.dtype], _TYPE_CODE[torch.float64], B,
                             xs.stride(0) if B > 1 else H * xs.stride(1), xs.stride(1), Hf * Wf, Hf, Wf,
                             ri.data_ptr(), rok.data_ptr(), ci.data_ptr(), cok.data_ptr(), float("nan"), stream),
              "rg_nearest")
        for axis, name in ((0, "rows"), (1, "cols")):
            entry = self._lu[name]
            if entry is None:
                continue
            rc = lib.rg_spline_solve_f64(coeff.data_ptr(), B, Hf * Wf, Wf, Hf, Wf, axis, entry[1].data_ptr(),
                                         entry[2].data_ptr(), stream)
            if rc == _lib.RG_ERR_UNSUPPORTED:
                return None
            check(rc, "rg_spline_solve")
        return coeff

    # nearest on integers (or out_dtype="same"): pure gather kernel
    def _gather_tables(self, n_rows: int, n_cols: int):
        if (n_rows, n_cols) not in self._gather:
            def up(a, dt):
                return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.device)
            rb, cb = self.op._host_bands(n_rows, n_cols)
            self._gather[(n_rows, n_cols)] = (up(rb.idx[0], np.int32), up(rb.cover, np.uint8),
                                              up(cb.idx[0], np.int32), up(cb.cover, np.uint8),
                                              rb.n_src, cb.n_src)
        return self._gather[(n_rows, n_cols)]

    def _nearest_gather(self, x: torch.Tensor, out_dtype: torch.dtype, fill: float,
                        out: torch.Tensor | None = None) -> torch.Tensor:
        lead = x.shape[:-2]
        x3 = x.reshape((-1,) + tuple(x.shape[-2:]))
        if x3.stride(2) != 1:
            x3 = x3.contiguous()
        B, H, W = x3.shape
        ri, rok, ci, cok, n_rows, n_cols = self._gather_tables(H, W)
        if H != n_rows or W != n_cols:
            raise ValueError(f"data is [{H}, {W}] but the plan was built for [{n_rows}, {n_cols}]")
        Ho, Wo = ri.numel(), ci.numel()
        y = _out_or_new(out, (B, Ho, Wo), out_dtype, x3.device)
        if B:
            check(
                lib.rg_nearest(x3.data_ptr(), y.data_ptr(), _TYPE_CODE[x3.dtype], _TYPE_CODE[out_dtype], B,
                               x3.stride(0) if B > 1 else H * x3.stride(1), x3.stride(1), Ho * Wo, Ho, Wo,
                               ri.data_ptr(), rok.data_ptr(), ci.data_ptr(), cok.data_ptr(), float(fill),
                               _stream_ptr(x3.device)),
                "rg_nearest",
            )
        return y.view(tuple(lead) + (Ho, Wo))

    def result_dtype(self, in_dtype: torch.dtype, out_dtype: str = "reference") -> torch.dtype:
        """Dtype of the result for data of ``in_dtype`` (the reference's rule, or ``out_dtype="same"``)."""
        if out_dtype not in ("reference", "same"):
            raise ValueError("out_dtype must be 'reference' or 'same'")
        if self.method == "nearest":
            if in_dtype not in _TYPE_CODE:
                in_dtype = torch.float32 if in_dtype in (torch.float16, torch.bfloat16) else torch.int32
            is_float = in_dtype in (torch.float32, torch.float64)
            if self.op.pole_rows:
                return in_dtype if is_float else torch.float64
            return in_dtype if (is_float or out_dtype == "same") else torch.float64
        if self.method == "linear" and out_dtype == "same" and in_dtype in (torch.float32, torch.float16,
                                                                           torch.bfloat16):
            return torch.float32
        return torch.float64

    def out_hw(self, n_rows: int, n_cols: int):
        """Shape of the regridded ``[rows, cols]`` for a source slab of ``[n_rows, n_cols]``."""
        if self.method == "cubic":
            Hf, Wf = self.pre._out_hw(n_cols, n_rows)
            return self.op._out_hw(Wf, Hf)
        return self.op._out_hw(n_cols, n_rows)

    def __call__(self, x: torch.Tensor, out_dtype: str = "reference", fill: float = float("nan"),
                 out: torch.Tensor | None = None, nan_flag: torch.Tensor | None = None) -> torch.Tensor:
        """``nan_flag`` (cubic, used by the chunked host path): a device int32 the NaN rule ORs into INSTEAD of
        poisoning this call's result; the caller applies scipy's "any NaN -> everything NaN" over all chunks."""
        if not x.is_cuda:
            raise ValueError("InterpPlan takes CUDA tensors; use apply_host() for host buffers")
        if out_dtype not in ("reference", "same"):
            raise ValueError("out_dtype must be 'reference' or 'same'")
        is_float = x.dtype in (torch.float32, torch.float64)
        if self.method == "nearest":
            if x.dtype not in _TYPE_CODE:
                x = x.to(torch.float32 if x.dtype in (torch.float16, torch.bfloat16) else torch.int32)
                is_float = x.dtype == torch.float32
            if self.op.pole_rows:  # a synthetic pole row can be the nearest neighbour: needs the mean
                x3, lead = _as_batch(x, x.dtype if is_float else torch.float64)
                y = self.op.run(x3, NAN_PROPAGATE, 0.0, out=out)
                return y.view(tuple(lead) + tuple(y.shape[-2:]))
            target = x.dtype if (is_float or out_dtype == "same") else torch.float64
            return self._nearest_gather(x, target, fill, out=out)

        work = torch.float64
        if self.method == "linear" and out_dtype == "same" and x.dtype in (torch.float32, torch.float16, torch.bfloat16):
            work = torch.float32
        # (cubic on fp32 data: the gather in front of the solves converts, no separate cast pass)
        keep32 = self.method == "cubic" and x.dtype == torch.float32
        x3, lead = _as_batch(x, torch.float32 if keep32 else work)
        if self.method == "linear":
            y = self.op.run(x3, NAN_PROPAGATE, 0.0, out=out)
            return y.view(tuple(lead) + tuple(y.shape[-2:]))

        # cubic: NaN flag -> prefilter (source grid) -> evaluate (target grid) -> poison
        B, H, W = x3.shape
        Hf, Wf = self.pre._out_hw(W, H)
        Ho, Wo = self.op._out_hw(Wf, Hf)
        y = _out_or_new(out, (B, Ho, Wo), work, x3.device)
        if self.all_nan:
            y.fill_(float("nan"))
            return y.view(tuple(lead) + (Ho, Wo))
        flag = nan_flag if nan_flag is not None else torch.zeros(1, dtype=torch.int32, device=x3.device)
        stream = _stream_ptr(x3.device)
        if B:
            flag_fn = lib.rg_nan_flag_f32 if keep32 else lib.rg_nan_flag_f64
            check(flag_fn(x3.data_ptr(), B, H, W, x3.stride(0) if B > 1 else H * x3.stride(1),
                          x3.stride(1), flag.data_ptr(), stream), "rg_nan_flag")
        for s0 in range(0, B, self.chunk):
            xs = x3[s0:s0 + self.chunk]
            coeff = self._coefficients_lu(xs, Hf, Wf) if (self._lu is not None and self.use_lu) else None
            if coeff is None:
                coeff = self.pre.run(xs.to(work), NAN_PROPAGATE, 0.0)
            self.op.run(coeff, NAN_PROPAGATE, 0.0, out=y[s0:s0 + self.chunk])
        if B and nan_flag is None:
            check(lib.rg_poison_if_f64(y.data_ptr(), y.numel(), flag.data_ptr(), stream), "rg_poison_if")
        return y.view(tuple(lead) + (Ho, Wo))

    def apply_host(self, x_host, out_host: torch.Tensor | None = None, out_dtype: str = "reference",
                   fill: float = float("nan"), chunk_batch: int | None = None,
                   pipeline: HostPipeline | None = None) -> torch.Tensor:
        """Host-buffer path: the stack is streamed in chunks of ``chunk_batch`` slabs through a persistent
        pipeline (H2D, kernels and D2H overlap; the device holds only ``slots`` chunks, so results far larger
        than HBM -- config 3 is 1.4 TB -- pass through a reused block).  Takes numpy arrays or CPU tensors
        (page-locked ones are copied without staging) and returns a CPU tensor."""
        x_host = _host_tensor(x_host)
        if self.method == "nearest" and x_host.dtype not in _TYPE_CODE:
            x_host = x_host.to(torch.float32 if x_host.dtype in (torch.float16, torch.bfloat16) else torch.int32)
        if self.method != "nearest" and x_host.dtype not in (torch.float32, torch.float64):
            x_host = x_host.to(torch.float64)  # scipy computes linear / cubic in float64
        if x_host.dim() < 2:
            raise ValueError("data must have at least the two regridded dims [rows, cols]")
        res_dtype = self.result_dtype(x_host.dtype, out_dtype)
        x3 = x_host.reshape((-1,) + tuple(x_host.shape[-2:]))
        lead = x_host.shape[:-2]
        B, H, W = x3.shape
        Ho, Wo = self.out_hw(H, W)
        if out_host is None:
            out_host = torch.empty((B, Ho, Wo), dtype=res_dtype, pin_memory=x3.is_pinned() or PINNED_RESULTS)
        else:
            out_host = _check_out_host(out_host, (B, Ho, Wo), res_dtype)
        if B == 0:
            return out_host.view(tuple(lead) + (Ho, Wo))
        if chunk_batch is None:  # ~256 MB of the larger side per chunk
            per_slab = max(H * W * x3.element_size(), Ho * Wo * out_host.element_size())
            chunk_batch = max(1, min(B, (256 << 20) // per_slab))
        pipe = pipeline if pipeline is not None else default_pipeline(self.device)
        with torch.cuda.device(self.device):
            flag = None
            if self.method == "cubic":
                flag = torch.zeros(1, dtype=torch.int32, device=self.device)
                torch.cuda.current_stream(self.device).synchronize()

            def compute(x_dev, y_dev):
                self(x_dev, out_dtype=out_dtype, fill=fill, out=y_dev, nan_flag=flag)

            pipe.stream(x3, out_host, chunk_batch, compute)
            if flag is not None and int(flag.item()) != 0:  # scipy: one NaN anywhere -> everything NaN
                out_host.fill_(float("nan"))
        return out_host.view(tuple(lead) + (Ho, Wo))


class DeviceBins:
    """A ``tables.Bins`` uploaded to the GPU, plus the ``rg_bins_t`` that points at it."""

    def __init__(self, bins: tables.Bins, device: torch.device):
        self.host = bins

        def up(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(device)

        self.start, self.count = up(bins.start, np.int32), up(bins.count, np.int32)
        self.cover, self.out_index = up(bins.cover, np.uint8), up(bins.out_index, np.int32)
        self.gather = None if bins.gather is None else up(bins.gather, np.int64)
        covered = bins.count[bins.cover > 0]
        uniform = int(covered[0]) if covered.size and bool((covered == covered[0]).all()) else 0
        self.c = rg_bins_t(self.start.data_ptr(), self.count.data_ptr(), self.cover.data_ptr(),
                           self.out_index.data_ptr(), bins.n_out, bins.n_src, bins.src_offset, bins.src_step,
                           uniform)

    @property
    def ref(self):
        return C.byref(self.c)


def _strip_elems(cols: tables.Bins) -> int:
    """Max number of source columns spanned by 32 consecutive target columns (strip kernels)."""
    ends = cols.start.astype(np.int64) + cols.count
    first = np.arange(0, cols.n_out, 32)
    return int((ends[np.minimum(first + 31, cols.n_out - 1)] - cols.start[first]).max())


_STAT_CODE = {"sum": 0, "mean": 1, "var": 2, "std": 3, "min": 4, "max": 5, "median": 6}
VALID_STAT_METHODS = ["sum", "mean", "var", "std", "median", "max", "min"]


class ReducePlan:
    """Device-resident bin tables for ``da.regrid.most_common / least_common / stat``.

    Mirrors what ``compute_mode`` / ``statistic_reduce`` (methods/flox_reduce.py:40-187) prepare before
    calling flox -- sorted targets, closed-left interval bins, the trimmed source domain -- after
    ``format_for_regrid(stats=True)`` (longitude shift / wrap padding only, no pole rows).
    """

    def __init__(self, rows: AxisSpec | None, cols: AxisSpec | None, *, device=None):
        self.device = _require_cuda(device)
        if rows is None and cols is None:
            raise ValueError("at least one axis must be regridded")
        if (rows is not None and rows.kind == "lon") or (cols is not None and cols.kind == "lat"):
            raise ValueError("layout must be [..., lat, lon]: transpose the data first")
        # stats=True formatting: no pole rows (utils.py:258-261)
        self.rows_bins = None if rows is None else tables.bins_axis(tables.plain_axis(rows.src), rows.tgt)
        if cols is None:
            self.cols_bins = None
        else:
            cols_axis = (tables.format_lon_axis(cols.src, cols.tgt) if cols.kind == "lon"
                         else tables.plain_axis(cols.src))
            self.cols_bins = tables.bins_axis(cols_axis, cols.tgt)
        self._dev: dict = {}
        self._labels: dict = {}   # (dtype, values) -> device label tables
        self._launch: dict = {}   # launch geometry per (shape, element size, scratch)

    def _label_tables(self, np_dtype, values):
        """Device copies of the label list (as given / sorted / rank), cached per plan."""
        vals = np.asarray(values).astype(np_dtype)
        key = (np_dtype.str, vals.tobytes())
        if key not in self._labels:
            order = np.argsort(vals, kind="stable")
            self._labels[key] = (
                vals,
                bool(np.array_equal(vals, np.arange(vals.size))),
                torch.from_numpy(np.ascontiguousarray(vals)).to(self.device),
                torch.from_numpy(np.ascontiguousarray(vals[order])).to(self.device),
                torch.from_numpy(order.astype(np.int32)).to(self.device),
            )
        return self._labels[key]

    def _geometry(self, rows, cols, elem_size: int, scratch: int):
        """(cols_per_chunk, tile_elems, strip_elems, max_bin_rows, max_bin_cols) of a table pair."""
        key = (id(rows), id(cols), elem_size, scratch)
        if key not in self._launch:
            cpc, tile = tables.plan_chunks(rows.host, cols.host, elem_size, scratch)
            self._launch[key] = (cpc, tile, _strip_elems(cols.host), int(rows.host.count.max()),
                                 int(cols.host.count.max()))
        return self._launch[key]

    def _bins(self, n_rows: int, n_cols: int):
        if (n_rows, n_cols) not in self._dev:
            rows = self.rows_bins if self.rows_bins is not None else tables.identity_bins(n_rows)
            cols = self.cols_bins if self.cols_bins is not None else tables.identity_bins(n_cols)
            self._dev[(n_rows, n_cols)] = (DeviceBins(rows, self.device), DeviceBins(cols, self.device))
        return self._dev[(n_rows, n_cols)]

    @property
    def fully_covered(self) -> bool:
        return (self.rows_bins is None or bool(self.rows_bins.cover.all())) and (
            self.cols_bins is None or bool(self.cols_bins.cover.all()))

    def _prepare(self, x: torch.Tensor):
        lead = x.shape[:-2]
        x3 = x.reshape((-1,) + tuple(x.shape[-2:]))
        rows, cols = self._bins(x3.shape[1], x3.shape[2])
        # a source axis whose order is neither ascending, descending nor a rotation is sorted first
        if rows.gather is not None:
            x3 = x3.index_select(1, rows.gather)
        if cols.gather is not None:
            x3 = x3.index_select(2, cols.gather)
        if x3.stride(2) != 1:
            x3 = x3.contiguous()
        if x3.shape[1] != rows.host.n_src or x3.shape[2] != cols.host.n_src:
            raise ValueError(
                f"data is {list(x3.shape[1:])} but the plan was built for [{rows.host.n_src}, {cols.host.n_src}]"
            )
        return x3, lead, rows, cols

    def out_hw(self, n_rows: int, n_cols: int):
        rows, cols = self._bins(n_rows, n_cols)
        return rows.host.n_out, cols.host.n_out

    @staticmethod
    def _check_mode_dtype(dtype):
        if dtype not in (torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64):
            raise ValueError(
                "Your input data has to be of an integer datatype for this method.\n"
                f"    instead, your data is of type '{dtype}'."
            )

    def mode_result(self, dtype: torch.dtype, fill_value, warn: bool = True):
        """(result dtype, fill) of ``mode`` for data of ``dtype`` (methods/_shared.py:56-72)."""
        if fill_value is None and not self.fully_covered:
            if warn:
                import warnings

                warnings.warn(
                    "No fill_value is provided; data will be cast to floating point dtype to be able to use "
                    "NaN for missing values.", stacklevel=1,
                )
            return torch.float64, float("nan")
        return dtype, 0.0 if fill_value is None else float(fill_value)

    def mode(self, x: torch.Tensor, values, fill_value=None, anti_mode: bool = False,
             out: torch.Tensor | None = None, _warn: bool = True) -> torch.Tensor:
        """``compute_mode`` (methods/flox_reduce.py:122-187)."""
        if not x.is_cuda:
            raise ValueError("ReducePlan takes CUDA tensors; use mode_host() for host buffers")
        self._check_mode_dtype(x.dtype)
        x3, lead, rows, cols = self._prepare(x)
        np_dtype = np.dtype(str(x.dtype).replace("torch.", ""))
        vals, dense, v_dev, vs_dev, rank_dev = self._label_tables(np_dtype, values)
        out_dtype, fill = self.mode_result(x.dtype, fill_value, _warn)
        B, Ho, Wo = x3.shape[0], rows.host.n_out, cols.host.n_out
        y = _out_or_new(out, (B, Ho, Wo), out_dtype, x3.device)
        cpc, tile, strip, max_rows, max_cols = self._geometry(rows, cols, x3.element_size(), 8 * vals.size * 4)
        if B:
            check(
                lib.rg_mode(x3.data_ptr(), y.data_ptr(), _TYPE_CODE[x3.dtype], _TYPE_CODE[out_dtype], B,
                            x3.stride(0) if B > 1 else x3.shape[1] * x3.stride(1), x3.stride(1), Ho * Wo,
                            rows.ref, cols.ref, cpc, tile, strip, max_rows, max_cols,
                            vs_dev.data_ptr(), rank_dev.data_ptr(),
                            v_dev.data_ptr(), int(vals.size), int(dense), int(bool(anti_mode)), fill,
                            _stream_ptr(x3.device)),
                "rg_mode",
            )
        return y.view(tuple(lead) + (Ho, Wo))

    def stat(self, x: torch.Tensor, method: str, skipna: bool = False, fill_value=None,
             out: torch.Tensor | None = None) -> torch.Tensor:
        """``statistic_reduce`` (methods/flox_reduce.py:40-100)."""
        if method not in VALID_STAT_METHODS:
            raise ValueError(f"Invalid method. Please choose from '{VALID_STAT_METHODS}'.")
        if not x.is_cuda:
            raise ValueError("ReducePlan takes CUDA tensors; use stat_host() for host buffers")
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float64)
        x3, lead, rows, cols = self._prepare(x)
        B, Ho, Wo = x3.shape[0], rows.host.n_out, cols.host.n_out
        y = _out_or_new(out, (B, Ho, Wo), x3.dtype, x3.device)
        sort_cap = 0
        if method == "median":
            _, _, _, max_rows, max_cols = self._geometry(rows, cols, x3.element_size(), 0)
            sort_cap = 1 << max(max_rows * max_cols - 1, 0).bit_length()
        cpc, tile, strip, max_rows, max_cols = self._geometry(rows, cols, x3.element_size(),
                                                             8 * sort_cap * x3.element_size())
        fn = lib.rg_stat_f64 if x3.dtype == torch.float64 else lib.rg_stat_f32
        fill = float("nan") if fill_value is None else float(fill_value)
        if B:
            check(
                fn(x3.data_ptr(), y.data_ptr(), B, x3.stride(0) if B > 1 else x3.shape[1] * x3.stride(1),
                   x3.stride(1), Ho * Wo, rows.ref, cols.ref, cpc, tile, _STAT_CODE[method],
                   int(bool(skipna)), sort_cap, strip, max_rows, max_cols, fill, _stream_ptr(x3.device)),
                "rg_stat",
            )
        return y.view(tuple(lead) + (Ho, Wo))

    # ------------------------------------------------------------------------ host-buffer paths
    def _host(self, x_host, res_dtype, out_host, chunk_batch, pipeline, compute):
        if x_host.dim() < 2:
            raise ValueError("data must have at least the two regridded dims [rows, cols]")
        x3 = x_host.reshape((-1,) + tuple(x_host.shape[-2:]))
        lead = x_host.shape[:-2]
        B, H, W = x3.shape
        Ho, Wo = self.out_hw(H, W)
        if out_host is None:
            out_host = torch.empty((B, Ho, Wo), dtype=res_dtype, pin_memory=x3.is_pinned() or PINNED_RESULTS)
        else:
            out_host = _check_out_host(out_host, (B, Ho, Wo), res_dtype)
        if B:
            if chunk_batch is None:  # ~256 MB of source per chunk
                chunk_batch = max(1, min(B, (256 << 20) // max(H * W * x3.element_size(), 1)))
            pipe = pipeline if pipeline is not None else default_pipeline(self.device)
            pipe.stream(x3, out_host, chunk_batch, compute)
        return out_host.view(tuple(lead) + (Ho, Wo))

    def mode_host(self, x_host, values, fill_value=None, anti_mode: bool = False,
                  out_host: torch.Tensor | None = None, chunk_batch: int | None = None,
                  pipeline: HostPipeline | None = None) -> torch.Tensor:
        """``mode`` for data in host memory (numpy / CPU tensor), streamed through a persistent pipeline."""
        x_host = _host_tensor(x_host)
        self._check_mode_dtype(x_host.dtype)
        res_dtype, _ = self.mode_result(x_host.dtype, fill_value)

        def compute(x_dev, y_dev):
            self.mode(x_dev, values, fill_value=fill_value, anti_mode=anti_mode, out=y_dev, _warn=False)

        return self._host(x_host, res_dtype, out_host, chunk_batch, pipeline, compute)

    def stat_host(self, x_host, method: str, skipna: bool = False, fill_value=None,
                  out_host: torch.Tensor | None = None, chunk_batch: int | None = None,
                  pipeline: HostPipeline | None = None) -> torch.Tensor:
        """``stat`` for data in host memory (numpy / CPU tensor), streamed through a persistent pipeline."""
        if method not in VALID_STAT_METHODS:
            raise ValueError(f"Invalid method. Please choose from '{VALID_STAT_METHODS}'.")
        x_host = _host_tensor(x_host)
        if x_host.dtype not in (torch.float32, torch.float64):
            x_host = x_host.to(torch.float64)

        def compute(x_dev, y_dev):
            self.stat(x_dev, method, skipna=skipna, fill_value=fill_value, out=y_dev)

        return self._host(x_host, x_host.dtype, out_host, chunk_batch, pipeline, compute)


def sm_count() -> int:
    return int(_lib.lib.rg_device_sm_count())
