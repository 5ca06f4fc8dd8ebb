"""Where a few-matrix evaluation spends its device time: per-kernel-class sums (single stream, no look-ahead overlap) next
to the wall time of the normal (overlapped, graph-replayed) call, for one matrix and for the 8 chains of config 3."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads

def run(N, B, kernel):
    X3, y, Th = workloads.c2_inputs(N, max(B, 2))
    Th = Th[:B]
    gp = g3.GP(X3, g3.Bias(), kernel(X3)); gp.observed(X3, y)
    if Th.shape[1] != gp.ndim:
        Th = np.tile(gp.dict_to_array(gp.params_default), (B, 1)) + 0.01 * np.random.default_rng(0).standard_normal((B, gp.ndim))
    for _ in range(3): gp.logp_dlogp_batch(Th)
    n = 10
    t0 = time.perf_counter()
    for _ in range(n): gp.logp_dlogp_batch(Th)
    wall = (time.perf_counter() - t0) / n * 1e3
    ctx = gp.ctx
    ctx.set_graphs(0); ctx.prof_enable(True)
    gp.logp_dlogp_batch(Th); ctx.prof_read()
    gp.logp_dlogp_batch(Th)
    pr = ctx.prof_read()
    ctx.prof_enable(False); ctx.set_graphs(1)
    tot = sum(v["ms"] for v in pr.values())
    print("N=%d B=%d  overlapped wall %.2f ms | serial class sums %.2f ms: " % (N, B, wall, tot) +
          "  ".join("%s %.2f (%d)" % (k, v["ms"], v["launches"]) for k, v in pr.items()), flush=True)

for N, B in ((2048, 1), (2048, 8), (4096, 1), (1024, 1)):
    run(N, B, lambda X: g3.SE(X))
