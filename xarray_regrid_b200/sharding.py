"""Multi-GPU partitioning: contiguous blocks of the flattened batch (time x level x member) per rank.

Every 2-D slab is regridded independently with replicated KB-sized tables, so the data path needs no
collective.  ``torch.distributed`` is only used by callers for barriers / timing
reductions and, optionally, to gather results on one rank.
"""

from __future__ import annotations


def shard_bounds(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """Half-open range ``[start, stop)`` of ``n_items`` owned by ``rank``: contiguous, balanced to
    within one item, earlier ranks take the remainder."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    if n_items < 0:
        raise ValueError("n_items must be non-negative")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world_size: int) -> list[int]:
    return [b - a for a, b in (shard_bounds(n_items, world_size, r) for r in range(world_size))]


def regrid_sharded(plan, x_full, rank: int, world_size: int, **kw):
    """Apply ``plan`` to this rank's contiguous block of the leading (batch) dimension of ``x_full``
    and return ``(start, stop, y_block)``.  ``x_full`` may be any indexable stack (a memory-mapped
    array, a host tensor, ...); only the rank's block is touched."""
    start, stop = shard_bounds(x_full.shape[0], world_size, rank)
    return start, stop, plan(x_full[start:stop], **kw)


def bind_to_gpu_numa_node(device_index: int) -> list[int] | None:
    """Pin the calling process to the CPU cores NVML reports as local to GPU ``device_index`` (same NUMA node
    / PCIe root).  Pinned host buffers allocated afterwards land in that node's memory, so the H2D / D2H
    streams of the ranks do not cross the inter-socket link.  Returns the cores, or None when NVML or the
    affinity call is unavailable (the process is then left alone)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        try:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            n_cpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        finally:
            pynvml.nvmlShutdown()
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cores = [c for c in cores if c in allowed]
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:  # noqa: BLE001 - best effort: no NVML, no permission, non-Linux
        return None


def latitude_band(src, tgt, world_size: int, rank: int):
    """Split a SINGLE-slab group-by regrid (most_common / least_common / stat) over ranks by target rows.

    Returns ``(tgt_index, src_slice)``: the positions (in ``tgt``'s own order) of the target rows owned by
    ``rank`` and the slice of source rows that rank needs.  Bins are closed-left intervals of width ``res``
    around the targets (``_shared.py:12-18``), so a band of targets needs exactly the source rows inside
    ``[first target - res, last target + res]`` -- the trimming rule of ``reduce_data_to_new_domain``
    (``_shared.py:92-108``) applied to the band: no halo exchange, no collective, and the per-band result equals
    the corresponding rows of the full result as long as the source is at least as fine as the target (the
    coarsening case these methods exist for), which is checked.  Every rank gets at least two targets (the bin
    width is inferred from the target spacing); surplus ranks get an empty band.
    """
    import numpy as np

    src, tgt = np.asarray(src, dtype=np.float64), np.asarray(tgt, dtype=np.float64)
    if tgt.size < 2:
        raise ValueError("need at least two target points")
    order = np.argsort(tgt, kind="stable")
    tgt_sorted = tgt[order]
    res = float(np.median(np.diff(tgt_sorted)))
    if src.size > 1 and float(np.max(np.abs(np.diff(src)))) > res * (1 + 1e-9):
        raise ValueError("latitude_band needs a source at least as fine as the target")
    usable = max(1, min(world_size, tgt.size // 2))  # ranks that get >= 2 targets
    if not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    if rank >= usable:
        return np.empty(0, np.int64), slice(0, 0)
    a, b = shard_bounds(tgt.size, usable, rank)
    lo, hi = tgt_sorted[a] - res, tgt_sorted[b - 1] + res
    inside = np.nonzero((src >= lo) & (src <= hi))[0]
    src_slice = slice(int(inside[0]), int(inside[-1]) + 1) if inside.size else slice(0, 0)
    return order[a:b], src_slice
