"""Outer-block width sweep for single evaluations (g3_set_potrf_block)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads
for N in (1024, 2048, 4096, 8192):
    X3, y, _ = workloads.c2_inputs(N, 1)
    gp = g3.GP(X3, g3.Bias(), g3.SE(X3)); gp.observed(X3, y)
    th = gp.dict_to_array(gp.params_default)
    out = []
    for w in (0, 1, 2, 3, 4, 6, 8, 12, 16):
        gp.ctx.set_potrf_block(w)
        for _ in range(3): gp.dlogp(th, array=True)
        n = 20 if N <= 4096 else 8
        t0 = time.perf_counter()
        for _ in range(n): gp.dlogp(th, array=True)
        tg = (time.perf_counter() - t0) / n
        t0 = time.perf_counter()
        for _ in range(n): gp.logp(th, array=True)
        tl = (time.perf_counter() - t0) / n
        out.append("w=%d: %.2f/%.2f" % (w, tl * 1e3, tg * 1e3))
    gp.ctx.set_potrf_block(0)
    print("N=%d  logp/logp+grad ms  " % N + "  ".join(out), flush=True)
