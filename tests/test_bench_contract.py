"""bench.py's contract pieces that do not need a GPU: the reference arm (CPU port) runs with every host thread
also under torchrun's OMP_NUM_THREADS=1, never maps the engine package, and carries the same ``config`` as the
engine arm; bench.py's numpy-only grids equal the package's."""

import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def test_bench_grids_equal_the_package_grids():
    sys.path.insert(0, str(ROOT))
    import bench
    from xarray_regrid_b200 import synthetic

    for mine, theirs in ((bench._grid_025(), synthetic.grid_025()), (bench._grid_1(), synthetic.grid_1())):
        for a, b in zip(mine, theirs):
            np.testing.assert_array_equal(a, b)
    lat, lon = synthetic.grid_025()
    np.testing.assert_array_equal(bench._continent_mask(lat, lon, 0.09), synthetic.continent_mask(lat, lon, 0.09))


def test_reference_arm_line():
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")  # as under torchrun
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600, check=True)
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    sys.path.insert(0, str(ROOT))
    import bench

    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == "GB/s"
    assert line["config"] == bench.workload_config(2, "strong", bench.STEPS_PER_STACK)
    assert line["scaling"] == "strong" and line["n_gpus"] == 2
    assert line["engine_modules_loaded"] == []  # no CUDA library in the reference arm's process
    cores = os.cpu_count() or 1
    assert line["cpu_baseline"]["cores"] == cores or cores == 1, "the CPU port must use every host thread"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, env=env, timeout=120, check=True)
    assert out.stdout.strip() == ""
