"""CPU emulation of the error-free int8 slicing that experiments/i8gemm/ozaki_dgemm.cu runs on the tensor cores
(tools/ozaki_cholesky_study.py): the arithmetic claims the prototype and DESIGN.md section 8 rest on."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import ozaki_cholesky_study as oz  # noqa: E402


def _rows(rng, r, K):
    return (10.0 ** (-3 * np.abs(rng.uniform(-1, 1, (r, 1))))) * rng.uniform(-1, 1, (r, K))


def test_slicing_is_exact_and_fits_int8():
    rng = np.random.default_rng(0)
    A = _rows(rng, 40, 96)
    S = 9
    q, scale = oz.slices(A, S)
    assert all(np.all(np.abs(x) <= 127) and np.array_equal(x, np.trunc(x)) for x in q)
    rec = sum(x.astype(np.longdouble) * np.longdouble(2.0) ** (-7 * (t + 1)) for t, x in enumerate(q)) * scale[:, None]
    assert float(np.max(np.abs(rec - A) / scale[:, None])) <= 2.0 ** (-7 * S)
    # the level sums stay inside int32 for K <= 14 000:  (d+1) K 127^2 < 2^31
    assert S * 14000 * 127 ** 2 < 2 ** 31


def test_sliced_update_has_fp64_quality():
    rng = np.random.default_rng(1)
    A = _rows(rng, 64, 256)
    C0 = rng.uniform(-1, 1, (64, 64))
    f = A @ A.T
    C0[:, ::4] = f[:, ::4] + 1e-9 * rng.uniform(-1, 1, (64, 16))      # Schur-complement-like cancellation
    ref = C0.astype(np.longdouble) - A.astype(np.longdouble) @ A.T.astype(np.longdouble)
    mag = np.abs(C0) + np.abs(A) @ np.abs(A).T
    err = {}
    for S in (7, 8, 9):
        C = C0.copy()
        oz.ozaki_update(C, A, S)
        err[S] = float(np.max(np.abs(C - ref) / mag))
    err64 = float(np.max(np.abs((C0 - f) - ref) / mag))
    assert err[9] < 1e-15 and err[9] < 4 * max(err64, 1.2e-16)
    assert err[7] > err[8] > err[9]


def test_one_scale_per_row_of_the_factor_never_overflows():
    """|L_ij| <= sqrt(K_ii): the scale 2^ceil(log2 sqrt(K_ii)) is valid for every panel of row i (left-looking use)."""
    K, y = oz.gram(256)
    L = np.linalg.cholesky(K)
    bound = np.sqrt(np.diag(K)) * (1.0 + 2.0 ** -40)
    assert np.all(np.abs(L) <= bound[:, None])
    q, scale = oz.slices(L[128:, :128], 9, bound[128:])
    assert all(np.all(np.abs(x) <= 127) for x in q)
    # blocked factorisation with that scale: log-det at fp64 quality
    Lo = oz.blocked_cholesky(K, 64, lambda C, A, row0: oz.ozaki_update(C, A, 9, bound[row0:]))
    ld, ld0 = 2 * np.sum(np.log(np.diag(Lo))), 2 * np.sum(np.log(np.diag(L)))
    assert abs(ld - ld0) <= 1e-13 * abs(ld0)
