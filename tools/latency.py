"""Call latency of the public API at small and mid sizes (BASELINE config 1, one MCMC chain of config 3, ...), with
the per-kernel-class device times (g3_prof_*) of one gradient call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads
for N in (200, 1024, 2048, 4096, 8192):
    if N == 200:
        x, y = workloads.c1_inputs()
    else:
        X3, y, _ = workloads.c2_inputs(N, 1); x = X3
    gp = g3.GP(x, g3.Bias(), g3.SE(x)); gp.observed(x, y)
    th = gp.dict_to_array(gp.params_default)
    for _ in range(5): gp.logp(th, array=True); gp.dlogp(th, array=True)
    n = 50 if N <= 2048 else 10
    t0 = time.perf_counter()
    for _ in range(n): gp.logp(th, array=True)
    t1 = time.perf_counter()
    for _ in range(n): gp.dlogp(th, array=True)
    t2 = time.perf_counter()
    l0 = gp.ctx.launch_count(); gp.dlogp(th, array=True); l1 = gp.ctx.launch_count()
    gp.ctx.prof_enable(True); gp.ctx.set_groups(1)
    gp.dlogp(th, array=True); gp.ctx.prof_read()
    gp.dlogp(th, array=True); pr = gp.ctx.prof_read()
    gp.ctx.prof_enable(False); gp.ctx.set_groups(4)
    tg = (t2 - t1) / n
    print("N=%d  logp %.0f us   logp+grad %.0f us = %.2f TFLOP/s  (%d launches per gradient call)  device ms by class: %s"
          % (N, 1e6 * (t1 - t0) / n, 1e6 * tg, N ** 3 / tg / 1e12, l1 - l0,
             {k: round(v["ms"], 3) for k, v in pr.items()}), flush=True)
