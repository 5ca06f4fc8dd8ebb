"""CPU oracle for the g3py exact-GP hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy fp64 restatement of the reference's algorithm
(griosd/g3py, Theano graph + SciPy LAPACK).  It exists to *check* the CUDA
path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it; nothing under
`g3py_b200/` does (tests/test_layout.py enforces that).

Parity status: PARTIALLY PINNED.  The reference (Theano/PyMC3) cannot run in
this image, and it ships no test-suite.  The only known-answer vectors in the
reference tree are the two N=2 Student-t term prints and the N=30 term-sum
identity stored in notebooks/07-Student-t-Process.ipynb:206-218,273-282;
`tests/test_oracle_kat.py` pins the oracle to those.  Everything else is
cross-checked three ways (finite differences, torch-CPU fp64 autograd of the
same forward, LU- vs Cholesky-based posterior) — see DESIGN.md §3.
"""
from .g3_oracle import *  # noqa: F401,F403
