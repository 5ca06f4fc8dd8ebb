"""Feasibility study for the next step named in DESIGN.md: fp64-equivalent GEMM on the INT8 tensor cores
(Ozaki-type error-free slicing) for the trailing updates of the blocked Cholesky.  CPU / NumPy only - it emulates the
integer arithmetic exactly and answers one question: how many 7-bit slices (hence how many int8 GEMMs) do the operands
that actually occur in this pipeline (panels of a Cholesky factor of the benchmark's Gram matrix) need for an update
C -= A B^T that is as accurate as the DMMA one?

    python tools/ozaki_study.py

Scheme: every row of A and of B is scaled by a power of two so that |x| < 1, then cut into slices of 7 signed bits,
x = sum_t x_t 2^(-7(t+1)), x_t integer in [-64, 64].  A_t B_u^T is an exact integer GEMM (K <= 4096: |sum| <= 2^24,
int32 accumulators in TMEM are enough); the fp64 result is the sum over slice pairs with t + u < s, each scaled by
2^(-7(t+u+2)) and the two row scales.  Slices beyond s are dropped (truncation error ~2^(-7s) relative to the row
scales, not to the result - the cancellation inside a Cholesky update is what decides how many are needed).
"""
import os
import sys

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from g3py_b200 import workloads  # noqa: E402


def slices(M, s):
    """Row-scaled 7-bit slices of M: returns (list of int matrices, per-row exponent e with |M_row| < 2^e)."""
    amax = np.max(np.abs(M), axis=1)
    e = np.where(amax > 0, np.floor(np.log2(np.maximum(amax, 1e-300))) + 1, 0.0)
    R = M / np.exp2(e)[:, None]                      # |R| < 1, exact (power-of-two scaling)
    out = []
    for _ in range(s):
        R = R * 128.0                                # exact
        q = np.trunc(R)                              # 7 bits + sign, |q| <= 127 (<= 64 after the first if rounded)
        out.append(q)                                # small integers held in float64: products and sums stay exact
        R = R - q                                    # exact remainder, |R| < 1
    return out, e


def ozaki_gemm(A, B, s):
    """A B^T from s slices per operand, slice pairs with t + u < s.  Integer products are exact."""
    As, ea = slices(A, s)
    Bs, eb = slices(B, s)
    C = np.zeros((A.shape[0], B.shape[0]), dtype=np.longdouble)
    n_gemm = 0
    for t in range(s):
        for u in range(s - t):
            P = As[t] @ Bs[u].T                       # integer-valued, |sum| <= K * 127^2 < 2^31: exact in float64
            assert np.max(np.abs(P)) < 2 ** 31         # (BLAS); int32 accumulators suffice on the device
            C += P.astype(np.longdouble) * np.longdouble(2.0) ** (-7 * (t + u + 2))
            n_gemm += 1
    return (C * np.exp2(ea)[:, None].astype(np.longdouble) * np.exp2(eb)[None, :].astype(np.longdouble)), n_gemm


def main():
    N, nb = 2048, 1024
    X, y, Theta = workloads.c2_inputs(N, 1)
    th = np.exp(Theta[0])
    # K = SE + MAT52 + noise of the benchmark (natural-space hypers), as in bench.py's config 2
    d_se = ((X[:, None, :] - X[None, :, :]) ** 2 * (0.5 * th[2:5] ** 2)).sum(-1)
    d_m = ((X[:, None, :] - X[None, :, :]) ** 2 * (0.5 * th[6:9] ** 2)).sum(-1)
    s5 = np.sqrt(5 * d_m)
    K = th[1] * np.exp(-d_se) + th[5] * (1 + s5 + 5 * d_m / 3) * np.exp(-s5) + th[9] * np.eye(N)
    L = sla.cholesky(K, lower=True)
    # right-looking trailing update after the first nb columns:  C = K22 - L21 L21^T  (severe cancellation: the
    # Schur complement is much smaller than K22)
    L21 = L[nb:, :nb]
    K22 = K[nb:, nb:]
    ref = K22.astype(np.longdouble) - (L21.astype(np.longdouble) @ L21.T.astype(np.longdouble))
    f64 = K22 - L21 @ L21.T
    scale = np.max(np.abs(ref))
    print("N=%d, panel %d: |K22| ~ %.2e, |Schur complement| ~ %.2e, dynamic range of |L21| rows: %.1e .. %.1e"
          % (N, nb, np.max(np.abs(K22)), scale, np.min(np.max(np.abs(L21), axis=1)), np.max(np.abs(L21))))
    print("fp64 GEMM (what DMMA gives):          max abs err / max|C| = %.2e" % float(np.max(np.abs(f64 - ref)) / scale))
    for s in range(4, 11):
        P, n = ozaki_gemm(L21, L21, s)
        C = K22.astype(np.longdouble) - P
        err = float(np.max(np.abs(C - ref)) / scale)
        print("slices s=%2d: %3d int8 GEMMs (%2d with symmetry), max abs err / max|C| = %.2e,  int8 peak / GEMMs = %.0f TFLOP/s "
              "equivalent (4500 dense int8 TOP/s nominal)" % (s, n, (n + s) // 2, err, 4500.0 / n))
    # gradient path: K^-1 = U U^T with U = L^-T (lauum).  Rows of U span a wider range than rows of L.
    m = 768                                            # leading block (the extended-precision reference product is slow)
    U = sla.solve_triangular(L[:m, :m], np.eye(m), lower=True).T
    refK = U.astype(np.longdouble) @ U.T.astype(np.longdouble)
    f64K = U @ U.T
    sc = np.max(np.abs(refK))
    rmax = np.max(np.abs(U), axis=1)
    print("lauum K^-1 = U U^T: |K^-1| ~ %.2e, row maxima of U: %.1e .. %.1e; fp64 GEMM err %.2e"
          % (sc, rmax.min(), rmax.max(), float(np.max(np.abs(f64K - refK)) / sc)))
    for s in (7, 8, 9, 10):
        P, n = ozaki_gemm(U, U, s)
        print("  slices s=%2d: %3d int8 GEMMs, max abs err / max|K^-1| = %.2e" % (s, n, float(np.max(np.abs(P - refK)) / sc)))


if __name__ == "__main__":
    main()
