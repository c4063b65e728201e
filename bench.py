#!/usr/bin/env python
"""Benchmark of the conservative 0.25 -> 1 degree regrid hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling strong|weak]

A "step" is one pass of the hot path over the job's synthetic input: BASELINE config 2, ONE
8760 x 721 x 1440 fp64 stack with 10 % NaN, skipna=True, nan_threshold=0.5, regridded to 181 x 360.
With N GPUs the stack is time-sharded (``sharding.shard_bounds``: 1095 steps per GPU at N = 8), no
collective on the data path -> ``"scaling": "strong"`` (``--scaling weak`` gives every rank its own
8760-step stack instead).  Rank 0 prints ONE JSON line.

* ``value``        whole-job GB/s of ALGORITHMIC bytes ((721*1440 + 181*360) * 8 B per time step),
                   inputs resident in HBM, CUDA events on the launch stream, max over ranks
* ``roofline``     the conservative kernel against the measured HBM copy peak (MEASURED_PEAKS.json)
* ``e2e``          the same metric through the host-buffer C-ABI call on a persistent pipeline (pinned host
                   memory -> H2D -> kernel -> D2H inside the timed region), next to a copy-only run of
                   the same chunk loop (``copy_roofline``: what PCIe / host memory allow)
* ``parity_checked`` sampled slabs of the TIMED output compared with the oracle after timing
* ``ops``          (N = 1) configs 3, 4, 5 and conservative off the headline shape, device-resident, each
                   with its own clock sample and a post-timing oracle check of the timed output
* ``cpu_baseline`` the oracle (numpy port of the reference's dense-einsum path) on this box's host
                   cores, on a bounded sample of the same workload (N = 1, rank 0 only)

``--impl reference`` times only that CPU port, for the driver's reference arm; it never imports the
engine package (no CUDA library is mapped into that process).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "conservative 0.25°→1° regrid GB/s & Gcells/s at 1/2/4/8 B200 vs host CPU"
H, W, HO, WO = 721, 1440, 181, 360
STEPS_PER_STACK = 8760
NAN_FRACTION = 0.10
NAN_THRESHOLD = 0.5
BYTES_PER_TIMESTEP = (H * W + HO * WO) * 8  # un-padded input + output, fp64
CELLS_PER_TIMESTEP = H * W
REFERENCE_SAMPLE_STEPS = 32


def workload_config(world: int, scaling: str, timesteps: int) -> dict:
    """The workload both arms run (identical dict in the engine and the reference arm)."""
    per_gpu = timesteps if scaling == "weak" else -(-timesteps // world)
    return {
        "workload": (
            f"C2: conservative 0.25deg (721x1440) -> 1deg (181x360), fp64, {timesteps} hourly steps"
            f"{' per GPU' if scaling == 'weak' else ''}, 10% NaN, skipna=True, nan_threshold=0.5"
        ),
        "timesteps": timesteps * (world if scaling == "weak" else 1),
        "timesteps_per_gpu": per_gpu,
        "scaling": scaling,
        "parallelism": f"time-sharded x{world}, no collective",
        "algorithmic_bytes_per_timestep": BYTES_PER_TIMESTEP,
        "l2": f"inputs larger than L2 ({per_gpu * H * W * 8 / 1e9:.1f} GB read per step per GPU, 126 MB L2): no flush",
    }


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------- clock sampling
class ClockSampler:
    FIELDS = (
        "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,clocks.mem,temperature.gpu"
    )
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, period_ms: int = 100):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.rows = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", str(period_ms), "-i", str(index)],
                stdout=self.file, stderr=subprocess.DEVNULL,
            )
        except OSError:
            self.proc = None

    def _finish(self):
        if self.rows is not None:
            return
        self.rows = []
        if self.proc is None:
            return
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        import datetime as dt

        for line in self.file.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                ts = dt.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                extra = []
                for v in parts[9:11]:
                    try:
                        extra.append(float(v))
                    except ValueError:
                        extra.append(None)
                try:
                    extra.append(float(parts[3]))
                except ValueError:
                    extra.append(None)
                self.rows.append((ts, float(parts[1]), float(parts[2]), parts[5:9], extra))
            except ValueError:
                continue
        try:
            os.unlink(self.file.name)
        except OSError:
            pass

    def window(self, t0: float, t1: float) -> dict:
        """Clock summary of the samples taken in [t0, t1] (several windows may be read from one sampler)."""
        self._finish()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm, mx, reasons = [], [], set()
        mem, temp, power = [], [], []
        for ts, clk, cmax, flags, extra in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            sm.append(clk)
            mx.append(cmax)
            for lst, v in zip((mem, temp, power), extra):
                if v is not None:
                    lst.append(v)
            for name, val in zip(self.NAMES, flags):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        if mem:
            out["mem_mhz"] = statistics.median(mem)
        if temp:
            out["temp_c"] = max(temp)
        if power:
            out["power_w"] = statistics.median(power)
        out["reasons"] = sorted(reasons)
        return out

    stop = window


# ---------------------------------------------------------------------------- CPU port leg
# numpy-only copies of xarray_regrid_b200.synthetic's grids / mask (tests/test_bench_contract.py keeps them
# equal): the reference arm must not import the engine package, whose __init__ maps the CUDA library.
def _grid_025():
    import numpy as np

    return np.arange(-90.0, 90.0 + 0.25, 0.25), np.arange(0.0, 360.0, 0.25)


def _grid_1():
    import numpy as np

    return np.arange(-90.0, 91.0, 1.0), np.arange(0.0, 360.0, 1.0)


def _continent_mask(lat, lon, coverage=0.09):
    import numpy as np

    phi = np.radians(lat)[:, None]
    lam = np.radians(lon)[None, :]
    field = np.sin(3 * phi) * np.cos(2 * lam)
    return field > np.quantile(field, 1.0 - coverage)


def _use_all_host_threads() -> int:
    """BLAS / OpenMP threads = all host cores, also under torchrun (which exports OMP_NUM_THREADS=1).
    Must run before numpy is imported; threadpoolctl covers an already-loaded BLAS."""
    n = os.cpu_count() or 1
    if "numpy" not in sys.modules:
        for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[var] = str(n)
    try:
        import numpy  # noqa: F401
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=n)
    except Exception:  # noqa: BLE001
        pass
    return n


def _cpu_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        n = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:  # noqa: BLE001
        return 1


def _cpu_stack(steps: int, seed: int = 1234):
    """Same distribution as synthetic.conservative_stack, generated with numpy on the host."""
    import numpy as np

    lat, lon = _grid_025()
    rng = np.random.default_rng(seed)
    x = 0.5 + rng.random((steps, lat.size, lon.size))
    mask = _continent_mask(lat, lon, NAN_FRACTION - 0.01)[None] | (
        rng.random(x.shape, dtype=np.float32) < 0.01
    )
    x[mask] = np.nan
    return x


def _cpu_port_once(x):
    """The reference's conservative path restated in numpy (oracle): format_for_regrid + dense
    weights + notnull/fillna + 2 einsums + divide + where (methods/conservative.py:104-208)."""
    from oracle import conservative as oc

    lat, lon = _grid_025()
    tlat, tlon = _grid_1()
    t0 = time.perf_counter()
    y = oc.regrid_latlon(x, lat, lon, tlat, tlon, skipna=True, nan_threshold=NAN_THRESHOLD)
    return time.perf_counter() - t0, y


def cpu_baseline(target_seconds: float = 12.0) -> dict:
    _use_all_host_threads()
    probe = _cpu_stack(8)
    _cpu_port_once(probe[:2])  # warm up BLAS threads / einsum path cache
    t, _ = _cpu_port_once(probe)
    per_step = t / probe.shape[0]
    steps = int(min(256, max(16, target_seconds / max(per_step, 1e-6))))
    x = _cpu_stack(steps)
    best = min(_cpu_port_once(x)[0] for _ in range(2))
    return {
        "value": steps * BYTES_PER_TIMESTEP / best / 1e9,
        "unit": "GB/s",
        "gcells_per_s": steps * CELLS_PER_TIMESTEP / best / 1e9,
        "cores": _cpu_threads(),
        "host_cpus": os.cpu_count(),
        "kind": "port",
        "sample": f"{steps} of 8760 time steps of the same workload, best of 2, "
                  f"{best / steps * 1e3:.2f} ms per time step (numpy einsum, dense weights)",
    }


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _use_all_host_threads()
    world = max(int(args.gpus), 1)
    sample_steps = REFERENCE_SAMPLE_STEPS
    x = _cpu_stack(sample_steps)
    for _ in range(max(args.warmup, 1)):
        _cpu_port_once(x[:8])
    times = [_cpu_port_once(x)[0] for _ in range(args.steps)]
    total = sum(times)
    value = args.steps * sample_steps * BYTES_PER_TIMESTEP / total / 1e9
    cores = _cpu_threads()
    loaded = sorted(m for m in sys.modules if m.startswith("xarray_regrid_b200"))
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": "GB/s",
        "gcells_per_s": args.steps * sample_steps * CELLS_PER_TIMESTEP / total / 1e9,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(world, args.scaling, args.timesteps),
        "sample": f"each step = {sample_steps} of the {args.timesteps} time steps on the host CPU "
                  f"({cores} BLAS threads of {os.cpu_count()} cores); value = GB/s of that sample",
        "cpu_baseline": {
            "value": value, "unit": "GB/s", "cores": cores, "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"{sample_steps} time steps per step x {args.steps} steps; oracle = numpy port of the "
                      "reference's dense-einsum path (the reference itself needs xarray, not installable here)",
        },
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "engine_modules_loaded": loaded,
    }
    emit(line)


# ----------------------------------------------------------------------------- parity checks
def _parity(got, want, rtol, atol=0.0, exact=False, ignore=None) -> bool:
    """NaN placement equal and values within tolerance; cells flagged in ``ignore`` are skipped."""
    import numpy as np

    got, want = np.asarray(got), np.asarray(want)
    if got.shape != want.shape:
        return False
    if ignore is not None:
        got, want = np.where(ignore, 0, got), np.where(ignore, 0, want)
    if not np.array_equal(np.isnan(got), np.isnan(want)):
        return False
    if exact:
        return bool(np.array_equal(got, want, equal_nan=True))
    return bool(np.allclose(got, want, rtol=rtol, atol=atol, equal_nan=True))


def check_conservative_slabs(x, y, steps, lat, lon, tlat, tlon, skipna=True, thr=NAN_THRESHOLD, rtol=1e-12):
    """Oracle vs the timed output on the given time steps of this rank's shard."""
    import torch

    from oracle import conservative as oc

    import numpy as np

    idx = torch.tensor(sorted(set(int(s) for s in steps)), device=x.device)
    xs = x.index_select(0, idx).cpu().numpy()
    ys = y.index_select(0, idx).cpu().numpy()
    want, valid = oc.regrid_latlon(xs, lat, lon, tlat, tlon, skipna=skipna, nan_threshold=thr, return_valid=True)
    # a valid mass that equals the threshold to within rounding is a coin toss in the reference itself (it
    # depends on the summation order of the BLAS behind einsum): such cells are not compared
    ties = np.abs(valid - oc.valid_threshold(thr)) <= 1e-13
    if ties.sum() > 1e-3 * ties.size:
        return False
    return _parity(ys, want, rtol, atol=rtol * 1e-3, ignore=ties)


# --------------------------------------------------------------------------------- ops block
def _time_launches(fn, device, min_seconds=0.45, max_reps=2000):
    """(ms per launch, wall t0, wall t1): CUDA events on the current stream around back-to-back launches."""
    import torch

    for _ in range(3):
        fn()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize(device)
    one = max(e0.elapsed_time(e1), 1e-3)
    reps = int(min(max_reps, max(5, min_seconds * 1e3 / one)))
    t0 = time.time()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(device)
    t1 = time.time()
    return e0.elapsed_time(e1) / reps, reps, t0, t1


def run_ops(device, peak: float, sampler_index) -> dict:
    """Configs 3, 4, 5 and conservative off the headline shape: device-resident timing (CUDA events), each with
    its own clock window and an oracle check of the timed output (a slab / a latitude band)."""
    import warnings

    import numpy as np
    import torch

    from oracle import conservative as oc
    from oracle import interp as oi
    from oracle import reduce as orr
    from xarray_regrid_b200 import engine, sharding

    ops: dict = {}
    windows: dict = {}
    sampler = ClockSampler(sampler_index, period_ms=50)
    g = torch.Generator(device=device).manual_seed(7)

    def record(name, fn, algo_bytes, cells, cells_kind, check, pinned="pinned", extra=None):
        """Time ``fn`` (which writes its result into a fixed buffer), THEN run ``check`` on that timed output."""
        ms, reps, t0, t1 = _time_launches(fn, device)
        gbs = algo_bytes / (ms * 1e-3) / 1e9
        ops[name] = {
            "ms": ms, "reps": reps, "algorithmic_bytes": int(algo_bytes), "gbs": gbs, "frac": gbs / peak,
            "gcells_per_s": cells / (ms * 1e-3) / 1e9, "cells": cells_kind,
            "parity_checked": bool(check()), "parity_pin": pinned,
        }
        if extra:
            ops[name].update(extra)
        windows[name] = (t0, t1)

    lat1, lon1 = np.arange(-90.0, 91.0, 1.0), np.arange(0.0, 360.0, 1.0)
    lat01, lon01 = np.round(np.arange(-90.0, 90.05, 0.1), 10), np.round(np.arange(0.0, 359.95, 0.1), 10)
    lat025, lon025 = np.arange(-90.0, 90.25, 0.25), np.arange(0.0, 360.0, 0.25)

    # ---- C3: linear / cubic upsampling 1 -> 0.1 deg, fp32 input (batches of the 54,020 slabs)
    B = 128
    x = 250 + 50 * torch.rand((B, 181, 360), device=device, generator=g, dtype=torch.float32)
    xs = x[:1].cpu().numpy()
    lin = engine.InterpPlan(engine.AxisSpec(lat1, lat01, "lat"), engine.AxisSpec(lon1, lon01, "lon"), "linear")
    cub = engine.InterpPlan(engine.AxisSpec(lat1, lat01, "lat"), engine.AxisSpec(lon1, lon01, "lon"), "cubic")
    want_lin = oi.regrid_latlon(xs, lat1, lon1, lat01, lon01, "linear")
    want_cub = oi.regrid_latlon(xs, lat1, lon1, lat01, lon01, "cubic")
    y32 = torch.empty((B, 1801, 3600), dtype=torch.float32, device=device)
    y64 = torch.empty((64, 1801, 3600), dtype=torch.float64, device=device)
    in_b, out_cells = 181 * 360 * 4, 1801 * 3600
    record("c3_linear_f32_to_f32", lambda: lin(x, out_dtype="same", out=y32), B * (in_b + out_cells * 4),
           B * out_cells, "target cells",
           lambda: _parity(y32[:1].cpu().numpy(), want_lin, 1e-5),
           pinned="unpinned by a runnable reference test (oracle = scipy interp1d, the routine xarray calls)",
           extra={"slabs": B, "l2": "write-bound: 3.3 GB written per launch >> 126 MB L2"})
    record("c3_linear_f32_to_f64", lambda: lin(x[:64], out=y64), 64 * (in_b + out_cells * 8), 64 * out_cells,
           "target cells", lambda: _parity(y64[:1].cpu().numpy(), want_lin, 1e-12),
           pinned="unpinned by a runnable reference test (oracle = scipy interp1d)",
           extra={"slabs": 64, "note": "the reference's result dtype (float64)"})
    record("c3_cubic_f32_to_f64", lambda: cub(x[:64], out=y64), 64 * (in_b + out_cells * 8), 64 * out_cells,
           "target cells", lambda: _parity(y64[:1].cpu().numpy(), want_cub, 1e-10, atol=1e-9),
           pinned="unpinned: the reference holds no numeric cubic test (oracle = scipy interp1d(kind='cubic'))",
           extra={"slabs": 64, "note": "NaN flag + gather + 2 spline-coefficient solves + 4-tap evaluation + poison"})
    # end to end through InterpPlan.apply_host on the persistent pipeline (pinned memory, D2H-bound)
    pipe = engine.default_pipeline(device)
    xh = x.cpu().pin_memory()
    yh = torch.empty((B, 1801, 3600), dtype=torch.float32).pin_memory()
    lin.apply_host(xh, out_host=yh, out_dtype="same", chunk_batch=8, pipeline=pipe)
    t0 = time.perf_counter()
    for _ in range(3):
        lin.apply_host(xh, out_host=yh, out_dtype="same", chunk_batch=8, pipeline=pipe)
    dt = (time.perf_counter() - t0) / 3
    ops["c3_linear_f32_to_f32"]["e2e"] = {
        "value": B * (in_b + out_cells * 4) / dt / 1e9, "unit": "GB/s", "h2d_bytes": B * in_b,
        "d2h_bytes": B * out_cells * 4, "path": "InterpPlan.apply_host, 8-slab chunks, pinned, persistent pipeline",
        "parity_checked": _parity(yh[:1].numpy(), want_lin, 1e-5)}
    del x, y32, y64, xh, yh, lin, cub
    torch.cuda.empty_cache()

    # ---- C4: nearest 0.1 -> 0.25 deg, int32 categorical
    B = 73
    x = torch.randint(0, 32, (B, 1801, 3600), device=device, generator=g, dtype=torch.int32)
    near = engine.InterpPlan(engine.AxisSpec(lat01, lat025, "lat"), engine.AxisSpec(lon01, lon025, "lon"), "nearest")
    want = oi.regrid_latlon(x[:1].cpu().numpy(), lat01, lon01, lat025, lon025, "nearest")
    yi = torch.empty((B, 721, 1440), dtype=torch.int32, device=device)
    yd = torch.empty((B, 721, 1440), dtype=torch.float64, device=device)
    sector_bytes = 721 * 3600 * 4 + 721 * 1440 * 4  # every 32-byte sector of a used source row + the output
    record("c4_nearest_i32_to_i32", lambda: near(x, out_dtype="same", out=yi), B * 721 * 1440 * 8,
           B * 1801 * 3600, "source cells",
           lambda: _parity(yi[:1].cpu().numpy().astype(np.float64), want, 0, exact=True),
           pinned="unpinned by a runnable reference test (oracle = scipy interp1d(kind='nearest'))",
           extra={"steps": B, "traffic_bytes": B * sector_bytes,
                  "note": "a 2.5-element stride touches every 32-byte sector of a used row: DRAM traffic is "
                          "1.75x the algorithmic bytes; traffic_bound_frac is the fraction of peak on the bytes moved"})
    ops["c4_nearest_i32_to_i32"]["traffic_bound_frac"] = (
        B * sector_bytes / (ops["c4_nearest_i32_to_i32"]["ms"] * 1e-3) / 1e9 / peak)
    record("c4_nearest_i32_to_f64", lambda: near(x, out=yd), B * 721 * 1440 * 12, B * 1801 * 3600, "source cells",
           lambda: _parity(yd[:1].cpu().numpy(), want, 0, exact=True),
           pinned="unpinned by a runnable reference test (oracle = scipy interp1d(kind='nearest'))",
           extra={"steps": B, "traffic_bytes": B * (721 * 3600 * 4 + 721 * 1440 * 8),
                  "note": "the reference's result dtype (float64)"})
    ops["c4_nearest_i32_to_f64"]["traffic_bound_frac"] = (
        B * (721 * 3600 * 4 + 721 * 1440 * 8) / (ops["c4_nearest_i32_to_f64"]["ms"] * 1e-3) / 1e9 / peak)
    xh = x.cpu().pin_memory()
    yh = torch.empty((B, 721, 1440), dtype=torch.int32).pin_memory()
    near.apply_host(xh, out_host=yh, out_dtype="same", chunk_batch=8, pipeline=pipe)
    t0 = time.perf_counter()
    for _ in range(3):
        near.apply_host(xh, out_host=yh, out_dtype="same", chunk_batch=8, pipeline=pipe)
    dt = (time.perf_counter() - t0) / 3
    ops["c4_nearest_i32_to_i32"]["e2e"] = {
        "value": B * 721 * 1440 * 8 / dt / 1e9, "unit": "GB/s (algorithmic: out cells x 8 B)",
        "h2d_bytes": B * 1801 * 3600 * 4, "d2h_bytes": B * 721 * 1440 * 4,
        "h2d_gbs": B * 1801 * 3600 * 4 / dt / 1e9,
        "path": "InterpPlan.apply_host, 8-step chunks, pinned, persistent pipeline (H2D-bound: the whole source "
                "crosses PCIe although the gather reads 16 % of it)",
        "parity_checked": _parity(yh[:1].numpy().astype(np.float64), want, 0, exact=True)}
    del x, yi, yd, xh, yh, near
    torch.cuda.empty_cache()

    # ---- C5: most_common / zonal statistics 0.01 -> 0.25 deg (single slab)
    lat, lon = np.round(np.arange(-89.995, 90, 0.01), 3), np.round(np.arange(0.005, 360, 0.01), 3)
    tlat, tlon = np.arange(-89.875, 90, 0.25), np.arange(0.125, 360, 0.25)
    red = engine.ReducePlan(engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"))
    coarse = torch.randint(0, 32, (18000 // 64 + 1, 36000 // 64 + 1), device=device, generator=g)
    x = coarse.repeat_interleave(64, 0).repeat_interleave(64, 1)[:18000, :36000].to(torch.uint8).contiguous()
    noise = torch.rand((18000, 36000), device=device, generator=g) < 0.05
    x[noise] = torch.randint(0, 32, (int(noise.sum()),), device=device, generator=g, dtype=torch.uint8)
    del noise, coarse
    vals = np.arange(32)
    bands = [sharding.latitude_band(lat, tlat, 72, r) for r in (0, 40)]  # 2 bands of 10 target rows

    def band_check(y, data, op, rtol=0.0, **kw):
        good = True
        for tgt_idx, src in bands:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want = orr.regrid_latlon(data[src].cpu().numpy(), lat[src], lon, tlat[tgt_idx], tlon, op, **kw)
            got = y[tgt_idx.min():tgt_idx.max() + 1].cpu().numpy()
            good &= _parity(got, want, rtol, atol=rtol, exact=(rtol == 0.0))
        return good

    ym = torch.empty((720, 1440), dtype=torch.uint8, device=device)
    record("c5_most_common_u8_32_classes", lambda: red.mode(x, vals, out=ym), 18000 * 36000 + 720 * 1440,
           18000 * 36000, "source cells", lambda: band_check(ym, x, "most_common", values=vals),
           pinned="pinned (count + idxmax incl. ties: reference tests/test_reduce.py:87-110)",
           extra={"check": "2 latitude bands of 10 target rows (20 x 1440 cells) of the timed output vs the "
                           "oracle, bit-exact"})
    xh = x.cpu().pin_memory()
    yh = torch.empty((720, 1440), dtype=torch.uint8).pin_memory()
    red.mode_host(xh, vals, out_host=yh, pipeline=pipe)
    t0 = time.perf_counter()
    for _ in range(3):
        red.mode_host(xh, vals, out_host=yh, pipeline=pipe)
    dt = (time.perf_counter() - t0) / 3
    ops["c5_most_common_u8_32_classes"]["e2e"] = {
        "value": (18000 * 36000 + 720 * 1440) / dt / 1e9, "unit": "GB/s", "h2d_bytes": 18000 * 36000,
        "d2h_bytes": 720 * 1440, "path": "ReducePlan.mode_host, single slab, pinned, persistent pipeline",
        "parity_checked": bool(torch.equal(yh, ym.cpu()))}
    del x, xh, yh, ym
    torch.cuda.empty_cache()
    x = torch.randn((18000, 36000), device=device, generator=g, dtype=torch.float32)
    x[torch.rand((18000, 36000), device=device, generator=g) < 0.02] = float("nan")
    ys = torch.empty((720, 1440), dtype=torch.float32, device=device)
    for method, pin in (("mean", "unpinned (flox absent; numpy nanmean per closed-left bin)"),
                        ("var", "var(ddof=0) pinned by reference tests/test_reduce.py:220-235; skipna=True unpinned"),
                        ("median", "unpinned (flox absent; numpy nanmedian per closed-left bin)")):
        record(f"c5_stat_{method}_f32_skipna", lambda m=method: red.stat(x, m, skipna=True, out=ys),
               18000 * 36000 * 4 + 720 * 1440 * 4, 18000 * 36000, "source cells",
               lambda m=method: band_check(ys, x, "stat", rtol=2e-5, method=m, skipna=True), pinned=pin)
    del x, ys, red
    torch.cuda.empty_cache()

    # ---- conservative off the headline shape (fp64, 5 % NaN, skipna) and the headline shape in fp32
    shapes = {
        "cons_0p25_to_0p5_f64": (lat025, lon025, np.arange(-90.0, 90.1, 0.5), np.arange(0.0, 360, 0.5), 256, np.float64),
        "cons_0p5_to_1_f64": (np.arange(-90.0, 90.1, 0.5), np.arange(0.0, 360, 0.5), lat1, lon1, 1024, np.float64),
        "cons_0p1_to_0p25_f64": (lat01, lon01, lat025, lon025, 48, np.float64),
        "cons_1_to_2p5_f64": (lat1, lon1, np.arange(-90.0, 90.1, 2.5), np.arange(0.0, 360, 2.5), 2048, np.float64),
        "cons_0p25_to_1_f32": (lat025, lon025, lat1, lon1, 512, np.float32),
    }
    for name, (la, lo, tla, tlo, Bc, dt_np) in shapes.items():
        tdt = torch.float64 if dt_np == np.float64 else torch.float32
        plan = engine.ConservativePlan(engine.AxisSpec(la, tla, "lat"), engine.AxisSpec(lo, tlo, "lon"), device=device)
        x = torch.rand((Bc, la.size, lo.size), device=device, dtype=tdt, generator=g) + 0.5
        x[torch.rand(x.shape, device=device, generator=g) < 0.05] = float("nan")
        y = torch.empty((Bc, tla.size, tlo.size), device=device, dtype=tdt)
        ring = plan._ring(Bc, la.size, lo.size, tdt)
        kern = ("generic" if ring is None else "lanes (one target per lane)" if ring.host.lane_cols == 4
                else f"lanes (blocks, {ring.host.n_panels} panel(s) x {ring.host.n_warps} warps)" if ring.host.lane_cols == 2
                else "two-phase ring")
        esz = 8 if tdt == torch.float64 else 4
        record(name, lambda: plan(x, skipna=True, nan_threshold=0.5, out=y),
               Bc * (la.size * lo.size + tla.size * tlo.size) * esz, Bc * la.size * lo.size, "source cells",
               lambda: check_conservative_slabs(x, y, [0, Bc - 1], la, lo, tla, tlo,
                                                rtol=1e-12 if tdt == torch.float64 else 1e-5),
               pinned="pinned (weights bit-exact vs the reference's helpers; joint valid mass: tests/test_regrid.py:172-208)",
               extra={"slabs": Bc, "kernel": kern})
        del x, y, plan
        torch.cuda.empty_cache()

    for name, (t0, t1) in windows.items():
        ops[name]["clocks"] = sampler.window(t0, t1)
    return ops


# --------------------------------------------------------------------------------- GPU arm
def run_engine_arm(args) -> None:
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the engine arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    from xarray_regrid_b200 import engine, sharding, synthetic

    sys.path.insert(0, str(ROOT / "tools"))
    from h2d_probe import copy_roofline

    # one process per GPU: keep the rank (and the pinned buffers it allocates) on the GPU's NUMA node
    numa_cores = sharding.bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1 and numa_cores is None:
        print(f"[bench] WARNING rank {rank}: could not bind to the GPU's NUMA node (NVML / affinity unavailable); "
              "pinned buffers may land on a remote node", file=sys.stderr, flush=True)

    lat, lon = synthetic.grid_025()
    tlat, tlon = synthetic.grid_1()
    plan = engine.ConservativePlan(
        engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"), device=device
    )
    if args.scaling == "strong":
        s0, s1 = sharding.shard_bounds(args.timesteps, world, rank)
        total_steps = args.timesteps
    else:
        s0, s1 = 0, args.timesteps
        total_steps = args.timesteps * world
    my_steps = s1 - s0
    x = synthetic.conservative_stack(my_steps, lat, lon, nan_fraction=NAN_FRACTION,
                                     seed=1234 + rank, device=device)
    y = torch.empty((my_steps, HO, WO), dtype=torch.float64, device=device)
    job_bytes = total_steps * BYTES_PER_TIMESTEP
    job_cells = total_steps * CELLS_PER_TIMESTEP
    my_bytes = my_steps * BYTES_PER_TIMESTEP

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(device)

    def one_step():
        plan(x, skipna=True, nan_threshold=NAN_THRESHOLD, out=y)

    gpu_id = f"GPU-{torch.cuda.get_device_properties(device).uuid}"
    sampler = ClockSampler(gpu_id) if rank == 0 else None
    for _ in range(args.warmup):
        one_step()
    barrier()
    stream = torch.cuda.current_stream(device)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    wall0 = time.time()
    ev[0].record(stream)
    for i in range(args.steps):
        one_step()
        ev[i + 1].record(stream)
    barrier()
    wall1 = time.time()
    launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    clocks = sampler.window(wall0, wall1) if sampler else None

    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_total_ms = float(t.item())
    ms_per_step = max_total_ms / args.steps
    value = job_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- the timed output against the oracle (first, second, middle and last step of the shard)
    parity_steps = sorted({0, min(1, my_steps - 1), my_steps // 2, my_steps - 1})
    ok = check_conservative_slabs(x, y, parity_steps, lat, lon, tlat, tlon)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    parity_checked = bool(flag.item())

    # ---- roofline of the dominant (only) kernel: one launch per step on this rank
    peak, peak_src = _peaks()
    avg_launch_ms = statistics.mean(launch_ms)
    achieved = my_bytes / (avg_launch_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": None, "kernel": "rg::conservative_lanes_kernel<double, skipna=true, 14 warps>",
        "peak_source": peak_src, "algorithmic_bytes_per_launch": my_bytes,
        "launch_ms_avg": avg_launch_ms, "launch_ms_min": min(launch_ms),
    }
    # the same STREAM-style copy MEASURED_PEAKS.json was taken with (torch b.copy_(a), 1 Gi bf16 elements, read +
    # write bytes, best of 10), on THIS GPU right after the timed region: boxes of the pool differ under the 1 kW
    # power cap, and this tells a slow kernel from a slow box.  Informational; `peak` stays the recorded figure.
    try:
        a_ = torch.empty(1 << 30, dtype=torch.bfloat16, device=device)
        b_ = torch.empty_like(a_)
        best = float("inf")
        for _ in range(12):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            b_.copy_(a_)
            c1.record()
            torch.cuda.synchronize(device)
            best = min(best, c0.elapsed_time(c1))
        roofline["copy_gbs_this_gpu_now"] = 2 * a_.numel() * 2 / (best * 1e-3) / 1e9
        roofline["frac_of_copy_this_gpu_now"] = achieved / roofline["copy_gbs_this_gpu_now"]
        # ... and a pure READ (torch sum over the same 2 GiB): the regrid is 94 % reads, and the read rate is what
        # differs most between the GPUs of the pool (6.1 - 7 TB/s)
        a32 = a_.view(torch.float32)
        best_r = float("inf")
        for _ in range(8):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            a32.sum()
            c1.record()
            torch.cuda.synchronize(device)
            best_r = min(best_r, c0.elapsed_time(c1))
        roofline["read_gbs_this_gpu_now"] = a_.numel() * 2 / (best_r * 1e-3) / 1e9
        del a_, b_, a32
    except Exception:  # noqa: BLE001 - out of memory next to the 72.8 GB stack: skip the extra figure
        pass
    traffic_file = ROOT / "profiles" / "traffic_conservative.json"
    if traffic_file.exists():
        try:
            per_step = json.loads(traffic_file.read_text()).get("dram_bytes_per_timestep")
            roofline["traffic"] = None if per_step is None else round(per_step * my_steps)
        except Exception:  # noqa: BLE001
            pass

    # ---- end to end through the host-buffer C-ABI call (pinned host memory both ways, persistent pipeline)
    e2e = None
    if not args.no_e2e:
        block = min(args.e2e_block, my_steps)
        xh = torch.empty((block, H, W), dtype=torch.float64).pin_memory()
        xh.copy_(x[:block])
        yh = torch.empty((block, HO, WO), dtype=torch.float64).pin_memory()  # result block, reused too
        chunk = args.e2e_chunk
        pipe = engine.HostPipeline(device, slots=3)
        need = engine.lib.rg_conservative_host_workspace(chunk, H, W, HO, WO, 8, pipe.slots)
        ws = torch.empty(need, dtype=torch.uint8, device=device)

        def e2e_step():
            # the pinned source / result blocks are cycled to cover the shard; the calls are queued on ONE
            # persistent pipeline without draining in between, one drain at the end of the step
            done = 0
            while done < my_steps:
                n = min(block, my_steps - done)
                plan.apply_host(xh[:n], skipna=True, nan_threshold=NAN_THRESHOLD, out_host=yh[:n],
                                chunk_batch=chunk, workspace=ws, pipeline=pipe, drain=False)
                done += n
            pipe.drain()

        e2e_step()  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt_s = (time.perf_counter() - t0) / args.e2e_steps
        # the result block of the timed run: every call regrids the head of the (cycled) pinned source block
        from oracle import conservative as oc

        e2e_ok = _parity(yh[:2].numpy(), oc.regrid_latlon(xh[:2].numpy(), lat, lon, tlat, tlon, skipna=True,
                                                          nan_threshold=NAN_THRESHOLD), 1e-12)
        # copy-only roofline: the same chunk loop without the kernel, all ranks at once
        reps = max(1, my_steps // block)
        copy_roofline(pipe, xh, yh, chunk, 1)
        barrier()
        ct_s = copy_roofline(pipe, xh, yh, chunk, reps) * (my_steps / (reps * block))
        barrier()
        tt = torch.tensor([dt_s, ct_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt_s, ct_s = float(tt[0].item()), float(tt[1].item())
        e2e = {
            "value": job_bytes / dt_s / 1e9, "unit": "GB/s",
            "h2d_bytes_per_step": total_steps * H * W * 8,
            "d2h_bytes_per_step": total_steps * HO * WO * 8,
            "ms_per_step": dt_s * 1e3, "steps": args.e2e_steps,
            "copy_roofline": {"value": job_bytes / ct_s / 1e9, "unit": "GB/s", "ms_per_step": ct_s * 1e3,
                              "what": "the same chunked pinned H2D + D2H loop on the same pipeline with the kernel "
                                      "left out, all ranks at once (tools/h2d_probe.py)"},
            "frac_of_copy_roofline": ct_s / dt_s,
            "parity_checked": bool(e2e_ok),
            "numa_bound": (None if world == 1 else numa_cores is not None),
            "path": "ConservativePlan.apply_host -> rg_conservative_host_f64 on a persistent rg_pipeline_t "
                    f"(3 streams, {chunk}-step chunks, one cudaMemcpyAsync per chunk and direction, one drain per "
                    f"step); pinned source / result blocks of {block} steps cycled over the shard "
                    "(bounded host memory: the full 72.8 GB stack is never resident on the host)",
        }
        pipe.close()
        del xh, yh, ws

    # ---- accessor-level end to end (numpy in -> numpy out through Regridder, pageable memory)
    accessor = None
    if rank == 0 and world == 1 and not args.no_e2e:
        from xarray_regrid_b200.labelled import DataArray, Dataset

        n_acc = min(512, my_steps)
        xa = x[:n_acc].cpu().numpy()
        da = DataArray(xa, ("time", "latitude", "longitude"),
                       {"time": __import__("numpy").arange(n_acc), "latitude": lat, "longitude": lon})
        tgt = Dataset({}, {"latitude": tlat, "longitude": tlon})
        out = da.regrid.conservative(tgt, skipna=True, nan_threshold=NAN_THRESHOLD)  # warm-up (plan, staging)
        out = None
        # (the previous result is dropped before every call in both arms: its page-locked block is then reused
        # instead of a second one being page-locked inside the timed region)
        t0 = time.perf_counter()
        for _ in range(3):
            out = None
            out = da.regrid.conservative(tgt, skipna=True, nan_threshold=NAN_THRESHOLD)
        dt_a = (time.perf_counter() - t0) / 3
        xp = torch.from_numpy(xa).pin_memory()
        ref = plan.apply_host(xp, skipna=True, nan_threshold=NAN_THRESHOLD)
        t0 = time.perf_counter()
        for _ in range(3):
            ref = None
            ref = plan.apply_host(xp, skipna=True, nan_threshold=NAN_THRESHOLD)
        dt_p = (time.perf_counter() - t0) / 3
        accessor = {
            "value": n_acc * BYTES_PER_TIMESTEP / dt_a / 1e9, "unit": "GB/s", "timesteps": n_acc,
            "apply_host_pinned_value": n_acc * BYTES_PER_TIMESTEP / dt_p / 1e9,
            "ratio_to_apply_host": dt_p / dt_a,
            "bit_identical_to_apply_host": bool(__import__("numpy").array_equal(out.data, ref.numpy(), equal_nan=True)),
            "path": "DataArray(numpy).regrid.conservative(target): pageable numpy staged through page-locked "
                    "pieces (rg_pipeline_h2d_pageable / _d2h_pageable), same kernels",
        }
        del xa, xp, da, out, ref

    del x, y
    torch.cuda.empty_cache()
    ops = None
    if rank == 0 and world == 1 and not args.no_ops:
        try:
            ops = run_ops(device, peak, gpu_id)
        except Exception as exc:  # noqa: BLE001 - the headline line must still be printed
            ops = {"error": f"{type(exc).__name__}: {exc}"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline()

    if rank == 0:
        config = workload_config(world, args.scaling, args.timesteps)
        line = {
            "metric": METRIC,
            "value": value,
            "unit": "GB/s",
            "gcells_per_s": job_cells / (ms_per_step * 1e-3) / 1e9,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": args.scaling,
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": config,
            "rank0_cpu_affinity": (f"{len(numa_cores)} cores local to the GPU (NVML)" if numa_cores else "unchanged"),
            "parity_checked": parity_checked,
            "parity": f"timed output vs the oracle on steps {parity_steps} of every rank's shard, rtol 1e-12, NaN placement equal",
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "accessor_e2e": accessor,
            "ops": ops,
            "gpu_launches": args.steps * world,
            "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout() -> None:
    """Route everything libraries print on fd 1 (e.g. NCCL's version banner) to stderr so that the
    ONE JSON line is the only thing on stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: ONE stack of --timesteps steps split over the GPUs (BASELINE config 2); "
                         "weak: every GPU owns its own stack of --timesteps steps")
    ap.add_argument("--timesteps", type=int, default=STEPS_PER_STACK,
                    help="time steps of the stack (default: the full 8760 of config 2)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-block", type=int, default=512)
    ap.add_argument("--e2e-chunk", type=int, default=32)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ops", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup

    if args.impl == "reference":
        _quiet_stdout()
        run_reference_arm(args)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", "29517",
               str(Path(__file__).resolve())] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    _quiet_stdout()
    run_engine_arm(args)


if __name__ == "__main__":
    main()
