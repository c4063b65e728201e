"""The C-ABI library loads without a GPU and exports every function include/regrid_b200.h declares."""

import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "regrid_b200.h"
LIB = ROOT / "xarray_regrid_b200" / "csrc" / "libregrid_b200.so"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert LIB.exists(), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    lib = ctypes.CDLL(str(LIB))
    names = declared_functions()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert missing == []


def test_python_binding_covers_the_header():
    from xarray_regrid_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_functions()
    assert _lib.lib.rg_abi_version() == _lib.RG_ABI_VERSION
    assert b"NULL" in _lib.lib.rg_error_string(-1)


def test_argument_errors_without_a_gpu():
    """Argument validation happens before any CUDA call, so it is testable on the CPU box."""
    from xarray_regrid_b200 import _lib

    assert _lib.lib.rg_conservative_f64(None, None, 1, 1, 1, 1, None, None, 1, 0.5, None, None, None) == -1
    assert _lib.lib.rg_conservative_host_workspace(0, 1, 1, 1, 1, 8, 3) == 0
    assert _lib.lib.rg_conservative_host_workspace(2, 721, 1440, 181, 360, 8, 3) >= 3 * 2 * (721 * 1440 + 181 * 360) * 8
    assert _lib.lib.rg_pipeline_h2d(None, 0, None, None, 0) == -1
    assert _lib.lib.rg_pipeline_slots(None) == 0
    assert _lib.lib.rg_row_nanmean_f64(None, None, 1, 4, 4, 16, 4, 0, 3, None) == -1


def test_host_copy_is_a_plain_memcpy():
    """rg_host_copy (the staging copy of the pageable host paths) needs no GPU: any size, any thread count."""
    import numpy as np

    from xarray_regrid_b200 import _lib

    rng = np.random.default_rng(0)
    for n, threads in [(0, 4), (1, 1), (4097, 3), (9 << 20, 1), (33 << 20, 7), ((40 << 20) + 13, 64)]:
        src = rng.integers(0, 256, n, dtype=np.uint8)
        dst = np.zeros(n + 8, dtype=np.uint8)
        assert _lib.lib.rg_host_copy(dst.ctypes.data, src.ctypes.data, n, threads) == 0
        assert np.array_equal(dst[:n], src) and not dst[n:].any()


def test_host_copy_pool_survives_fork():
    """The staging memcpy keeps worker threads alive between calls; a fork()ed child (multiprocessing workers)
    inherits the pool object but none of its threads and must start its own instead of waiting for them."""
    import os
    import warnings

    import numpy as np

    from xarray_regrid_b200 import _lib

    src = np.random.default_rng(1).integers(0, 256, 48 << 20, dtype=np.uint8)
    dst = np.zeros_like(src)
    assert _lib.lib.rg_host_copy(dst.ctypes.data, src.ctypes.data, src.size, 6) == 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        pid = os.fork()
    if pid == 0:
        ok = 1
        try:
            d2 = np.zeros_like(src)
            _lib.lib.rg_host_copy(d2.ctypes.data, src.ctypes.data, src.size, 6)
            ok = 0 if np.array_equal(d2, src) else 1
        finally:
            os._exit(ok)
    _, status = os.waitpid(pid, 0)
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0
    assert np.array_equal(dst, src)
