"""CPU oracle for the g3py exact-GP hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy fp64 restatement of the reference's algorithm
(griosd/g3py, Theano graph + SciPy LAPACK).  It exists to *check* the CUDA
path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it; nothing under
`g3py_b200/` does (tests/test_host.py::test_oracle_is_imported_only_where_allowed enforces that).

Parity status: PINNED against outputs of the reference itself, executed in the
build container through a stand-in for the Theano / PyMC3 API (fixtures and
generator under tests/golden/, checks in tests/test_reference_goldens.py), and
against the known-answer prints of notebooks/07-Student-t-Process.ipynb:206-218
(tests/test_oracle.py).  Limits of the stand-in: DESIGN.md section 2.
"""
from .g3_oracle import *  # noqa: F401,F403
