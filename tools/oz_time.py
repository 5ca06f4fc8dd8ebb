"""Serial (one stream group) per-class device times of the bench step in DMMA and Ozaki modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads
X, y, Theta = workloads.c2_inputs(4096, 64)
gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X)); gp.observed(X, y)
ctx = gp.ctx
nat = gp.natural(Theta); delta, _, _, _ = gp._host_terms(nat, X, y, False); thk = gp._kernel_theta(nat)
for groups in (1, 4):
    for mode, mk in (("dmma", 0), ("ozaki", 512), ("ozaki", 1024), ("ozaki", 2048)):
        for grad in (False, True):
            ctx.set_gemm_mode(mode, mk); ctx.set_groups(groups)
            ctx.gp_upload(gp.desc, 0, np.array(delta), thk, want_grad=grad)
            ctx.gp_run(); ctx.sync()
            ctx.prof_enable(groups == 1)
            ctx.timer_begin()
            for _ in range(2): ctx.gp_run()
            ms = ctx.timer_end() / 2
            prof = ctx.prof_read() if groups == 1 else {}
            ctx.prof_enable(False)
            print("groups %d mode %-5s min_k %4d grad %d: %.2f ms  %s" % (groups, mode, mk, grad, ms, {k: round(v["ms"] / 2, 2) for k, v in prof.items()}), flush=True)
