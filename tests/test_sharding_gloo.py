"""N>1 host logic on CPU: world-size-2 gloo processes shard a stack by time with no data-path
collective; the only communication is the final (optional) gather and the max-over-ranks timing."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xarray_regrid_b200 import sharding


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 8, 365, 8760, 54020):
        for world in (1, 2, 3, 4, 8):
            bounds = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(bounds[:-1], bounds[1:]))
            sizes = sharding.shard_sizes(n, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    assert sharding.shard_sizes(8760, 8) == [1095] * 8  # config 2 over 8 B200
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakePlan:
    """Stands in for the CUDA plan (no GPU here): a per-slab function, like the real regrid."""

    def __call__(self, x, scale=1.0):
        return x.reshape(x.shape[0], -1).sum(dim=1) * scale


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    x = torch.rand(11, 4, 6, dtype=torch.float64)  # every rank sees the same (virtual) stack
    start, stop, y = sharding.regrid_sharded(_FakePlan(), x, rank, world, scale=2.0)
    # timing convention of bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # gather the blocks on rank 0 (outside any timed region)
    parts = [None] * world  # ragged blocks (11 = 6 + 5): gather as objects
    dist.all_gather_object(parts, (start, stop, y))
    if rank == 0:
        assert [p[:2] for p in parts] == [sharding.shard_bounds(11, world, r) for r in range(world)]
        out["y"] = torch.cat([p[2] for p in parts]).numpy()
        out["tmax"] = float(t.item())
        out["want"] = (x.reshape(11, -1).sum(dim=1) * 2.0).numpy()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        np.testing.assert_array_equal(out["y"], out["want"])
        assert out["tmax"] == float(world)


def test_latitude_band_split_of_a_single_slab_is_exact():
    """Single-slab configs (C5) shard over ranks by target-row bands with no exchange: the per-band results of the
    oracle, put back in place, equal the full result bit for bit (most_common) / exactly (mean, max)."""
    import numpy as np

    from oracle import reduce as orr
    from xarray_regrid_b200 import sharding

    rng = np.random.default_rng(5)
    lat = np.round(np.arange(-29.95, 30, 0.1), 6)[::-1].copy()  # descending source, like ERA5
    lon = np.round(np.arange(0.05, 40, 0.1), 6)
    tlat, tlon = np.arange(-28.0, 29.0, 1.0), np.arange(1.0, 39.0, 1.0)
    labels = rng.integers(0, 9, (lat.size, lon.size)).astype(np.uint8)
    field = rng.standard_normal((lat.size, lon.size))
    full_mode = orr.mode(labels, [lat, lon], [tlat, tlon], [-2, -1], np.arange(9), fill_value=255)
    full_mean = orr.stat(field, [lat, lon], [tlat, tlon], [-2, -1], "mean", skipna=True)
    full_max = orr.stat(field, [lat, lon], [tlat, tlon], [-2, -1], "max")
    for world in (1, 2, 3, 8):
        got_mode = np.full_like(full_mode, 254)
        got_mean = np.full_like(full_mean, np.nan)
        got_max = np.full_like(full_max, np.nan)
        seen = np.zeros(tlat.size, int)
        for rank in range(world):
            t_idx, s_rows = sharding.latitude_band(lat, tlat, world, rank)
            if t_idx.size == 0:
                continue
            seen[t_idx] += 1
            sub_t = tlat[t_idx]
            got_mode[t_idx] = orr.mode(labels[s_rows], [lat[s_rows], lon], [sub_t, tlon], [-2, -1], np.arange(9),
                                       fill_value=255)
            got_mean[t_idx] = orr.stat(field[s_rows], [lat[s_rows], lon], [sub_t, tlon], [-2, -1], "mean", skipna=True)
            got_max[t_idx] = orr.stat(field[s_rows], [lat[s_rows], lon], [sub_t, tlon], [-2, -1], "max")
        assert (seen == 1).all(), "every target row is owned by exactly one rank"
        np.testing.assert_array_equal(got_mode, full_mode)
        np.testing.assert_array_equal(got_mean, full_mean)
        np.testing.assert_array_equal(got_max, full_max)
