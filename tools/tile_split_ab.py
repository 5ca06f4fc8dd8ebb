"""Column split of few-tile GEMM launches (g3_set_tile_split) A/B on single evaluations and on the 8 chains of config 3.
The split changes the CTA count a launch starts from, hence the split-K partition of the deep launches: results agree to
rounding (1e-16), not bitwise, between the two settings; each setting is bitwise reproducible."""
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

for N, B in ((200, 1), (1024, 1), (2048, 1), (4096, 1), (2048, 8), (8192, 1)):
    if N == 200:
        x, y = workloads.c1_inputs(); Th = None
    else:
        x, y, Th = workloads.c2_inputs(N, max(B, 2))
    gp = g3.GP(x, g3.Bias(), g3.SE(x)); gp.observed(x, y)
    th = gp.dict_to_array(gp.params_default)
    TH = np.tile(th, (B, 1)) + 0.01 * np.random.default_rng(0).standard_normal((B, gp.ndim))
    ref = None
    for mode in (0, 1):
        gp.ctx.set_tile_split(mode)
        n = 60 if N <= 2048 else (15 if N <= 4096 else 5)
        tl = timeit(lambda: gp.logp_batch(TH), n)
        tg = timeit(lambda: gp.logp_dlogp_batch(TH), n)
        lp, g = gp.logp_dlogp_batch(TH)[:2]
        if ref is None: ref = (lp, g)
        print("N=%-5d B=%d tile_split %d  logp %7.0f us  logp+grad %7.0f us   same bits as split 0: %s"
              % (N, B, mode, 1e6 * tl, 1e6 * tg, bool(np.array_equal(lp, ref[0]) and np.array_equal(g, ref[1]))), flush=True)
    gp.ctx.set_tile_split(1)
