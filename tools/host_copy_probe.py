"""Throughput of rg_host_copy (the staging memcpy of the pageable host paths) by thread count and size."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from xarray_regrid_b200._lib import lib
src = np.ones(512 << 20, np.uint8)
dst_pin = torch.empty(512 << 20, dtype=torch.uint8).pin_memory()
dst_pag = np.empty(512 << 20, np.uint8); dst_pag[:] = 0
for name, dst_ptr in (("pageable->pinned", dst_pin.data_ptr()), ("pageable->pageable", dst_pag.ctypes.data)):
    for size in (32 << 20, 128 << 20, 512 << 20):
        row = []
        for th in (1, 2, 4, 8, 14, 16, 24):
            lib.rg_host_copy(dst_ptr, src.ctypes.data, size, th)
            t0 = time.perf_counter(); n = 0
            while time.perf_counter() - t0 < 0.3:
                lib.rg_host_copy(dst_ptr, src.ctypes.data, size, th); n += 1
            row.append(f"{th}t {n * size / (time.perf_counter() - t0) / 1e9:5.1f}")
        print(name, f"{size >> 20:4d} MiB:", "  ".join(row), "GB/s", flush=True)
