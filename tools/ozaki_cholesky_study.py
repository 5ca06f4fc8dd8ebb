"""End-to-end check for the step DESIGN.md section 8 plans: a blocked right-looking Cholesky whose trailing updates run
through the error-free int8 slicing of experiments/i8gemm/ozaki_dgemm.cu (emulated exactly on the CPU: integer-valued
float64 GEMMs are exact below 2^53), on the Gram matrix of the headline benchmark.  Question: do log|K| and
alpha = K^-1 y keep fp64 quality (the 1e-9 parity bar of the path) when ONLY the updates change arithmetic?

    python tools/ozaki_cholesky_study.py [N] [nb] [noise_scale]      (noise_scale < 1: worse conditioning)

Reference: the same factorisation in 80-bit long double.  CPU / NumPy only."""
import os
import sys

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from g3py_b200 import workloads  # noqa: E402


def gram(N, noise_scale=1.0):
    X, y, Theta = workloads.c2_inputs(N, 1)
    th = np.exp(Theta[0])
    th[9] *= noise_scale
    d2 = (X[:, None, :] - X[None, :, :]) ** 2
    d_se = (d2 * (0.5 * th[2:5] ** 2)).sum(-1)
    d_m = (d2 * (0.5 * th[6:9] ** 2)).sum(-1)
    s5 = np.sqrt(5 * d_m)
    return th[1] * np.exp(-d_se) + th[5] * (1 + s5 + 5 * d_m / 3) * np.exp(-s5) + th[9] * np.eye(N), y


def slices(M, s, bound=None):
    """bound: per-row upper bound of |M| to derive the power-of-two scale from (default: the row maximum itself)."""
    amax = np.max(np.abs(M), axis=1) if bound is None else bound
    _, e = np.frexp(amax)
    R = M * np.ldexp(1.0, -e)[:, None]
    out = []
    for _ in range(s):
        R = R * 128.0
        q = np.trunc(R)
        out.append(q)
        R = R - q
    return out, np.ldexp(1.0, e)


def ozaki_update(C, A, s, bound=None):
    """C -= A A^T the way ozaki_kernel does it: levels d = s-1 .. 0, each an exact integer sum, one fp64 RMW per level."""
    As, sc = slices(A, s, bound)
    for d in range(s - 1, -1, -1):
        P = sum(As[t] @ As[d - t].T for t in range(d + 1))
        C -= P * (sc[:, None] * (sc[None, :] * 2.0 ** (-7 * (d + 2))))


def blocked_cholesky(K, nb, update):
    A = K.copy()
    N = A.shape[0]
    for k in range(0, N, nb):
        e = min(k + nb, N)
        A[k:e, k:e] = np.linalg.cholesky(A[k:e, k:e])
        if e < N:
            A[e:, k:e] = sla.solve_triangular(A[k:e, k:e], A[e:, k:e].T, lower=True).T
            update(A[e:, e:], A[e:, k:e], e)
    return np.tril(A)


def longdouble_cholesky(K):
    A = K.astype(np.longdouble)
    N = A.shape[0]
    for j in range(N):
        A[j, j] = np.sqrt(A[j, j] - np.dot(A[j, :j], A[j, :j]))
        if j + 1 < N:
            A[j + 1:, j] = (A[j + 1:, j] - A[j + 1:, :j] @ A[j, :j]) / A[j, j]
    return np.tril(A)


def solve_ld(L, y):
    """alpha = L^-T L^-1 y in long double (so that only the FACTOR's error is graded)."""
    L = L.astype(np.longdouble)
    N = L.shape[0]
    z = np.zeros(N, dtype=np.longdouble)
    for i in range(N):
        z[i] = (y[i] - np.dot(L[i, :i], z[:i])) / L[i, i]
    a = np.zeros(N, dtype=np.longdouble)
    for i in range(N - 1, -1, -1):
        a[i] = (z[i] - np.dot(L[i + 1:, i], a[i + 1:])) / L[i, i]
    return a


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    K, y = gram(N, float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
    ev = np.linalg.eigvalsh(K)
    print("N=%d nb=%d cond(K)=%.2e" % (N, nb, ev[-1] / ev[0]))
    Lref = longdouble_cholesky(K)
    ld_ref = 2 * np.sum(np.log(np.diag(Lref)))
    a_ref = solve_ld(Lref, y.astype(np.longdouble))
    q_ref = np.dot(y.astype(np.longdouble), a_ref)

    def report(tag, L):
        ld = 2 * np.sum(np.log(np.diag(L).astype(np.longdouble)))
        a = solve_ld(L, y.astype(np.longdouble))
        q = np.dot(y.astype(np.longdouble), a)
        print("%-28s |dlogdet|/|logdet| = %.2e   |dq|/|q| = %.2e   max|dalpha|/max|alpha| = %.2e   max|dL|/max|L| = %.2e"
              % (tag, float(abs(ld - ld_ref) / abs(ld_ref)), float(abs(q - q_ref) / abs(q_ref)),
                 float(np.max(np.abs(a - a_ref)) / np.max(np.abs(a_ref))),
                 float(np.max(np.abs(L.astype(np.longdouble) - Lref)) / np.max(np.abs(Lref)))))

    def f64_update(C, A, row0):
        C -= A @ A.T
    report("fp64 updates (DMMA today)", blocked_cholesky(K, nb, f64_update))
    for s in (7, 8, 9, 10):
        report("int8 slices s=%d" % s, blocked_cholesky(K, nb, lambda C, A, row0, s=s: ozaki_update(C, A, s)))
    # ONE scale per row of L for the whole factorisation: |L_ij| <= sqrt(K_ii) (the row of L has norm sqrt(K_ii)), so
    # 2^ceil(log2 sqrt(K_ii)) is known before anything is factored and the slices of different panels of one row share
    # it -- what a left-looking schedule (K = all previous columns in ONE accumulation) needs
    rootd = np.sqrt(np.diag(K)) * (1.0 + 2.0 ** -40)
    for s in (8, 9, 10):
        report("s=%d, row scale sqrt(K_ii)" % s,
               blocked_cholesky(K, nb, lambda C, A, row0, s=s: ozaki_update(C, A, s, rootd[row0:])))


if __name__ == "__main__":
    main()
