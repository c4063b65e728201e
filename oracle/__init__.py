"""CPU oracle for the regrid hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy (+ scipy for interp1d) restatement, pandas-free, of what the reference
(xarray-regrid v0.4.2, mounted at /root/reference in the build container)
computes on the path ``ds.regrid.{conservative, linear, nearest, cubic,
most_common, least_common, stat}``.  It exists so that the CUDA engine in
``xarray_regrid_b200`` can be checked for parity on machines where the
reference itself cannot be imported (``xarray``, ``flox``, ``dask``, ``sparse``
and ``opt_einsum`` are not installed here or on the GPU box).

Rules (enforced by ``tests/test_no_oracle_in_product.py``):

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` legs may import this package;
* nothing under ``xarray_regrid_b200/`` imports it -- the product path has no
  CPU fallback and fails loudly when the CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):

* the weight / threshold / interval / grid helpers are pinned bit-for-bit
  against the reference's OWN functions, executed in the build container
  behind stub ``xarray`` / ``flox`` modules by ``oracle/gen_golden.py``
  (fixtures: ``tests/golden/ref_helpers.npz``);
* the end-to-end paths are pinned against every synthetic known-answer test
  the reference holds for this path (``tests/test_oracle_known_answers.py``
  cites each reference test file:line);
* third-party arithmetic that the reference delegates to and that is NOT
  pinned by any runnable reference test is called out where it is restated:
  cubic (scipy ``interp1d(kind="cubic")`` is called directly -- scipy is on the
  box), flox ``median/std/sum/mean/max`` and ``skipna=True`` statistics
  ("parity unpinned": numpy semantics per closed-left bin are used).
"""

from . import conservative, formatting, grid, interp, reduce  # noqa: F401
