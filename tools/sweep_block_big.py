"""Outer-block width sweep for one N=16384 evaluation."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads
X, y, Xs = workloads.c4_inputs(16384, 64)
gp = g3.TP(X, g3.Bias(), g3.SE(X)); gp.observed(X, y)
th = gp.dict_to_array(gp.params_default)
lay = [n for n, s, _ in gp.layout for _ in range(s)]
th[lay.index("TP_Freedom_degree")] = np.log(5.0); th[lay.index("TP_Noise_var")] = np.log(0.05)
for w in (4, 6, 8, 12, 16, 24):
    gp.ctx.set_potrf_block(w)
    gp.logp_dlogp(th)
    t0 = time.perf_counter(); gp.logp_dlogp(th); gp.logp_dlogp(th); tg = (time.perf_counter() - t0) / 2
    gp.logp(th, array=True)
    t0 = time.perf_counter(); gp.logp(th, array=True); gp.logp(th, array=True); tl = (time.perf_counter() - t0) / 2
    print("N=16384 w=%d logp %.1f ms  logp+grad %.1f ms" % (w, tl * 1e3, tg * 1e3), flush=True)
