"""Call latency of the public API at small sizes (BASELINE config 1 and friends)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads
for N in (200, 1024, 2048):
    if N == 200:
        x, y = workloads.c1_inputs()
    else:
        X3, y, _ = workloads.c2_inputs(N, 1); x = X3
    gp = g3.GP(x, g3.Bias(), g3.SE(x)); gp.observed(x, y)
    th = gp.dict_to_array(gp.params_default)
    for _ in range(5): gp.logp(th, array=True); gp.dlogp(th, array=True)
    n = 50
    t0 = time.perf_counter()
    for _ in range(n): gp.logp(th, array=True)
    t1 = time.perf_counter()
    for _ in range(n): gp.dlogp(th, array=True)
    t2 = time.perf_counter()
    l0 = gp.ctx.launch_count(); gp.dlogp(th, array=True); l1 = gp.ctx.launch_count()
    print("N=%d  logp %.0f us   logp+grad %.0f us  (%d launches per gradient call)" % (N, 1e6 * (t1 - t0) / n, 1e6 * (t2 - t1) / n, l1 - l0), flush=True)
