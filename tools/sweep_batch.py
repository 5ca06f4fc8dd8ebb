"""Left-looking (batch) vs blocked (w=1,4) schedules as a function of the batch size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads
for N in (1024, 2048, 4096):
    for B in (2, 4, 8, 16, 32):
        X3, y, Th = workloads.c2_inputs(N, B)
        gp = g3.GP(X3, g3.Bias(), g3.SE(X3) + g3.MAT52(X3)); gp.observed(X3, y)
        out = []
        for w in (1 << 20, 1, 4, 0):
            gp.ctx.set_potrf_block(w)
            for _ in range(2): gp.logp_dlogp_batch(Th)
            n = 6
            t0 = time.perf_counter()
            for _ in range(n): gp.logp_dlogp_batch(Th)
            out.append("%s %.2f" % ("left" if w > 100 else ("auto" if w == 0 else "w=%d" % w), (time.perf_counter() - t0) / n * 1e3))
        gp.ctx.set_potrf_block(0)
        print("N=%d B=%d (B*T=%d)  ms: %s" % (N, B, B * N // 128, "  ".join(out)), flush=True)
