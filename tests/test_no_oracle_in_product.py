"""The oracle is test infrastructure: nothing in the product package may import or execute it,
and the product must fail loudly (not fall back) when the CUDA library is missing."""

import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "xarray_regrid_b200"


def test_product_never_imports_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|importlib\.import_module\(\s*['\"]oracle", re.M)
    offenders = [str(p) for p in PKG.rglob("*.py") if pat.search(p.read_text())]
    assert offenders == []
    for p in list(PKG.rglob("*.cu")) + list(PKG.rglob("*.cuh")) + [ROOT / "include" / "regrid_b200.h"]:
        assert "oracle" not in p.read_text().lower(), p


def test_bench_uses_oracle_only_as_cpu_leg_or_checker():
    """bench.py may execute the oracle in its CPU legs (cpu_baseline / --impl reference) and, in the engine
    arm, only as the CHECKER of an output that has already been timed -- never inside a timed region."""
    import ast

    src = (ROOT / "bench.py").read_text()
    tree = ast.parse(src)
    users = {}
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and (node.module or "").split(".")[0] == "oracle":
                users.setdefault(fn.name, []).append(node.lineno)
    allowed = {"_cpu_port_once", "check_conservative_slabs", "run_ops", "run_engine_arm"}
    assert set(users) <= allowed, users
    lines = src.splitlines()
    # engine arm: the only direct use is the check of the e2e result block, after its timed loop has ended
    timed_end = next(i for i, ln in enumerate(lines, 1) if "dt_s = (time.perf_counter() - t0) / args.e2e_steps" in ln)
    assert all(ln > timed_end for ln in users.get("run_engine_arm", []))
    # the timing helper of the ops block knows nothing about the oracle
    helper = src[src.index("def _time_launches"):src.index("def run_ops")]
    assert "oracle" not in helper


def test_missing_library_fails_loudly(tmp_path):
    """Import the package from a copy that has no built .so: must raise ImportError, not fall back."""
    code = (
        "import sys, importlib.util, pathlib\n"
        f"src = pathlib.Path({str(PKG)!r})\n"
        f"dst = pathlib.Path({str(tmp_path)!r}) / 'xarray_regrid_b200'\n"
        "dst.mkdir()\n"
        "[ (dst / p.name).write_text(p.read_text()) for p in src.glob('*.py') ]\n"
        f"sys.path.insert(0, {str(tmp_path)!r})\n"
        "try:\n"
        "    import xarray_regrid_b200\n"
        "except ImportError as e:\n"
        "    assert 'no CPU fallback' in str(e), e\n"
        "    print('OK')\n"
        "else:\n"
        "    raise SystemExit('import succeeded without the CUDA library')\n"
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=tmp_path)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr
