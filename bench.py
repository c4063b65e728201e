#!/usr/bin/env python
"""Benchmark of the conservative 0.25 -> 1 degree regrid hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input: BASELINE config 2,
an 8760 x 721 x 1440 fp64 stack with 10 % NaN, skipna=True, nan_threshold=0.5, regridded to
181 x 360.  With N GPUs every rank owns one such stack (time-sharded job of N x 8760 steps,
no collective on the data path -> "scaling": "weak").  Rank 0 prints ONE JSON line.

* ``value``      whole-job GB/s of ALGORITHMIC bytes ((721*1440 + 181*360) * 8 B per time
                 step), inputs resident in HBM, CUDA events on the launch stream, max over ranks
* ``roofline``   the conservative kernel against the measured HBM copy peak (MEASURED_PEAKS.json)
* ``e2e``        the same metric through the host-buffer C-ABI call (pinned host memory ->
                 H2D -> kernel -> D2H inside the timed region)
* ``cpu_baseline`` the oracle (numpy port of the reference's dense-einsum path) on this box's
                 host cores, on a bounded sample of the same workload (N=1, rank 0 only)

``--impl reference`` times only that CPU port, for the driver's reference arm.
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
WORKLOAD = (
    "C2: conservative 0.25deg (721x1440) -> 1deg (181x360), fp64, 8760 hourly steps per GPU, "
    "10% NaN, skipna=True, nan_threshold=0.5"
)


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
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(index)],
                stdout=self.file, stderr=subprocess.DEVNULL,
            )
        except OSError:
            self.proc = None

    def stop(self, t0: float, t1: float) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        import datetime as dt

        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.file.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                ts = dt.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk, cmax = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            sm.append(clk)
            mx.append(cmax)
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.file.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------- CPU port leg
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

    from xarray_regrid_b200 import synthetic

    lat, lon = synthetic.grid_025()
    rng = np.random.default_rng(seed)
    x = 0.5 + rng.random((steps, lat.size, lon.size))
    mask = synthetic.continent_mask(lat, lon, NAN_FRACTION - 0.01)[None] | (
        rng.random(x.shape, dtype=np.float32) < 0.01
    )
    x[mask] = np.nan
    return x


def _cpu_port_once(x):
    """The reference's conservative path restated in numpy (oracle): format_for_regrid + dense
    weights + notnull/fillna + 2 einsums + divide + where (methods/conservative.py:104-208)."""
    from oracle import conservative as oc
    from xarray_regrid_b200 import synthetic

    lat, lon = synthetic.grid_025()
    tlat, tlon = synthetic.grid_1()
    t0 = time.perf_counter()
    y = oc.regrid_latlon(x, lat, lon, tlat, tlon, skipna=True, nan_threshold=NAN_THRESHOLD)
    return time.perf_counter() - t0, y


def cpu_baseline(target_seconds: float = 12.0) -> dict:
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
    sample_steps = 32
    x = _cpu_stack(sample_steps)
    for _ in range(max(args.warmup, 1)):
        _cpu_port_once(x[:8])
    times = [_cpu_port_once(x)[0] for _ in range(args.steps)]
    total = sum(times)
    value = args.steps * sample_steps * BYTES_PER_TIMESTEP / total / 1e9
    cores = _cpu_threads()
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
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "step": f"bounded sample: {sample_steps} of 8760 time steps per step on the host CPU"},
        "cpu_baseline": {
            "value": value, "unit": "GB/s", "cores": cores, "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"{sample_steps} time steps per step x {args.steps} steps; oracle = numpy port of the "
                      "reference's dense-einsum path (the reference itself needs xarray, not installable here)",
        },
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


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

    # one process per GPU: keep the rank (and the pinned buffers it allocates) on the GPU's NUMA node
    numa_cores = sharding.bind_to_gpu_numa_node(local_rank) if world > 1 else None

    lat, lon = synthetic.grid_025()
    tlat, tlon = synthetic.grid_1()
    plan = engine.ConservativePlan(
        engine.AxisSpec(lat, tlat, "lat"), engine.AxisSpec(lon, tlon, "lon"), device=device
    )
    steps_per_stack = args.timesteps
    x = synthetic.conservative_stack(steps_per_stack, lat, lon, nan_fraction=NAN_FRACTION,
                                     seed=1234 + rank, device=device)
    y = torch.empty((steps_per_stack, HO, WO), dtype=torch.float64, device=device)
    bytes_per_step = steps_per_stack * BYTES_PER_TIMESTEP
    cells_per_step = steps_per_stack * CELLS_PER_TIMESTEP

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(device)

    def one_step():
        plan(x, skipna=True, nan_threshold=NAN_THRESHOLD, out=y)

    sampler = ClockSampler(f"GPU-{torch.cuda.get_device_properties(device).uuid}") if rank == 0 else None
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
    clocks = sampler.stop(wall0, wall1) if sampler else None

    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_total_ms = float(t.item())
    ms_per_step = max_total_ms / args.steps
    value = world * bytes_per_step / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant (only) kernel: one launch per step on this rank
    peak, peak_src = _peaks()
    avg_launch_ms = statistics.mean(launch_ms)
    achieved = bytes_per_step / (avg_launch_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": None, "kernel": "rg::conservative_lanes_kernel<double, skipna=true, 14 warps>",
        "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_step,
        "launch_ms_avg": avg_launch_ms, "launch_ms_min": min(launch_ms),
    }
    traffic_file = ROOT / "profiles" / "traffic_conservative.json"
    if traffic_file.exists():
        try:
            per_step = json.loads(traffic_file.read_text()).get("dram_bytes_per_timestep")
            roofline["traffic"] = None if per_step is None else round(per_step * args.timesteps)
        except Exception:  # noqa: BLE001
            pass

    # ---- end to end through the host-buffer C-ABI call (pinned host memory both ways)
    e2e = None
    if not args.no_e2e:
        block = min(args.e2e_block, steps_per_stack)
        xh = torch.empty((block, H, W), dtype=torch.float64).pin_memory()
        xh.copy_(x[:block])
        yh = torch.empty((block, HO, WO), dtype=torch.float64).pin_memory()  # result block, reused too
        chunk = 32
        need = engine.lib.rg_conservative_host_workspace(chunk, H, W, HO, WO, 8)
        ws = torch.empty(need, dtype=torch.uint8, device=device)

        def e2e_step():
            done = 0
            while done < steps_per_stack:  # the pinned source block is cycled to cover the stack
                n = min(block, steps_per_stack - done)
                plan.apply_host(xh[:n], skipna=True, nan_threshold=NAN_THRESHOLD,
                                out_host=yh[:n], chunk_batch=chunk, workspace=ws)
                done += n

        e2e_step()  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt_s = (time.perf_counter() - t0) / args.e2e_steps
        tt = torch.tensor([dt_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt_s = float(tt.item())
        e2e = {
            "value": world * bytes_per_step / dt_s / 1e9, "unit": "GB/s",
            "h2d_bytes_per_step": steps_per_stack * H * W * 8,
            "d2h_bytes_per_step": steps_per_stack * HO * WO * 8,
            "ms_per_step": dt_s * 1e3, "steps": args.e2e_steps,
            "path": "ConservativePlan.apply_host -> rg_conservative_host_f64 (3-stream H2D/kernel/D2H, "
                    f"{chunk}-step chunks); pinned source / result blocks of {block} steps cycled over the stack "
                    "(bounded host memory: the full 72.8 GB stack is never resident on the host)",
        }
        del xh, yh, ws

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline()

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value,
            "unit": "GB/s",
            "gcells_per_s": world * cells_per_step / (ms_per_step * 1e-3) / 1e9,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD if steps_per_stack == STEPS_PER_STACK
                else WORKLOAD.replace("8760", str(steps_per_stack)),
                "timesteps_per_gpu": steps_per_stack,
                "bytes_per_step_per_gpu": bytes_per_step,
                "parallelism": f"time-sharded x{world}, no collective",
                "rank0_cpu_affinity": (f"{len(numa_cores)} cores local to the GPU (NVML)" if numa_cores else "unchanged"),
                "l2": f"inputs larger than L2 ({steps_per_stack * H * W * 8 / 1e9:.1f} GB read per step, 126 MB L2)",
            },
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": args.steps,
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
    ap.add_argument("--timesteps", type=int, default=STEPS_PER_STACK,
                    help="time steps per GPU stack (default: the full 8760 of config 2)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-block", type=int, default=512)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
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
