"""Copy-only roofline of the host-buffer pipeline: the SAME chunk loop as ``rg_conservative_host_*`` (one pinned
H2D copy per chunk into a slot, one pinned D2H copy per chunk out of it, three streams, the pipeline's own
dependencies) with the kernel left out.  What it measures is what bounds the end-to-end number: PCIe and host
memory, per GPU and -- under torchrun -- for all ranks at once.

    python tools/h2d_probe.py                          # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_probe.py --json out.json

``copy_roofline`` is also imported by ``bench.py`` (``e2e.copy_roofline``).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def copy_roofline(pipe, x_host, y_host, chunk_items: int, repeat: int = 1, h2d: bool = True, d2h: bool = True):
    """Seconds for ``repeat`` passes of chunked pinned H2D (+ D2H) copies of ``x_host`` / ``y_host`` through the
    slots of ``pipe`` (an ``engine.HostPipeline``), no kernel.  ``x_host`` ``[B, ...]`` and ``y_host`` ``[B, ...]``
    must be page-locked."""
    import torch

    from xarray_regrid_b200._lib import check, lib

    assert x_host.is_pinned() and y_host.is_pinned()
    B = x_host.shape[0]
    in_item = x_host[0].numel() * x_host.element_size()
    out_item = y_host[0].numel() * y_host.element_size()
    n_chunks = -(-B // chunk_items)
    d_in = pipe.device_buffers("probe_in", chunk_items * in_item, n_chunks)
    d_out = pipe.device_buffers("probe_out", chunk_items * out_item, n_chunks)
    h = pipe.handle
    with torch.cuda.device(pipe.device):
        pipe.drain()
        t0 = time.perf_counter()
        for _ in range(repeat):
            for i, b0 in enumerate(range(0, B, chunk_items)):
                s = i % pipe.slots
                nb = min(chunk_items, B - b0)
                if h2d:
                    check(lib.rg_pipeline_h2d(h, s, d_in[s].data_ptr(), x_host[b0:b0 + nb].data_ptr(), nb * in_item),
                          "rg_pipeline_h2d")
                check(lib.rg_pipeline_compute_begin(h, s), "rg_pipeline_compute_begin")
                check(lib.rg_pipeline_compute_end(h, s), "rg_pipeline_compute_end")
                if d2h:
                    check(lib.rg_pipeline_d2h(h, s, y_host[b0:b0 + nb].data_ptr(), d_out[s].data_ptr(), nb * out_item),
                          "rg_pipeline_d2h")
        pipe.drain()
        return time.perf_counter() - t0


def main() -> None:
    import torch
    import torch.distributed as dist

    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk", type=int, default=32, help="time steps per chunk (the pipeline's chunk size)")
    ap.add_argument("--block", type=int, default=512, help="time steps in the pinned source block")
    ap.add_argument("--repeat", type=int, default=4)
    ap.add_argument("--json", default=None)
    ap.add_argument("--no-bind", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from xarray_regrid_b200 import engine, sharding

    cores = None if a.no_bind else sharding.bind_to_gpu_numa_node(local)
    H, W, HO, WO = 721, 1440, 181, 360
    x = torch.empty((a.block, H, W), dtype=torch.float64).pin_memory()
    y = torch.empty((a.block, HO, WO), dtype=torch.float64).pin_memory()
    x.fill_(1.0)
    pipe = engine.HostPipeline(dev, slots=3)
    out = {"n_gpus": world, "chunk_steps": a.chunk, "block_steps": a.block,
           "cpu_affinity_cores": None if cores is None else len(cores), "host_cpus": os.cpu_count()}

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(device_ids=[local])

    for name, kw in (("h2d_only", {"d2h": False}), ("d2h_only", {"h2d": False}), ("h2d_and_d2h", {})):
        copy_roofline(pipe, x, y, a.chunk, 1, **kw)  # warm-up
        barrier()
        dt = copy_roofline(pipe, x, y, a.chunk, a.repeat, **kw)
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        nbytes = a.repeat * ((x.numel() * 8 if kw.get("h2d", True) else 0) + (y.numel() * 8 if kw.get("d2h", True) else 0))
        out[name] = {"seconds_max_over_ranks": dt, "GB/s_aggregate": world * nbytes / dt / 1e9,
                     "GB/s_per_gpu": nbytes / dt / 1e9}
        barrier()
    if rank == 0:
        text = json.dumps(out, indent=1)
        print(text)
        if a.json:
            Path(a.json).write_text(text)
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
