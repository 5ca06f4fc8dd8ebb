"""Split-K rule A/B on single evaluations: 1 = at least 128 of the contraction per share, no triangular operand (default),
2 = shares down to 32, triangular solves included, 0 = off.  (profiles/r02t_splitk_ab.txt was taken with the meanings of
1 and 2 swapped.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads

def timeit(f, n):
    for _ in range(4): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n

for N in (200, 1024, 2048, 4096):
    x, y = workloads.c1_inputs() if N == 200 else workloads.c2_inputs(N, 1)[:2]
    gp = g3.GP(x, g3.Bias(), g3.SE(x)); gp.observed(x, y)
    th = gp.dict_to_array(gp.params_default)
    ref = None
    for mode in (1, 2, 0):
        gp.ctx.set_splitk(mode)
        n = 60 if N <= 2048 else 15
        tl = timeit(lambda: gp.logp(th, array=True), n)
        tg = timeit(lambda: gp.logp_dlogp(th), n)
        lp, g = gp.logp_dlogp(th)[:2]
        if ref is None: ref = (lp, g)
        print("N=%-5d splitk %d  logp %7.0f us  logp+grad %7.0f us   dlogp rel diff vs mode 1 %.1e  logp diff %.1e"
              % (N, mode, 1e6 * tl, 1e6 * tg, np.max(np.abs(g - ref[1])) / np.max(np.abs(ref[1])), abs(lp - ref[0]) / abs(ref[0])), flush=True)
    gp.ctx.set_splitk(1)
