import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
X, y, Theta = workloads.c2_inputs(N, B)
gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X)); gp.observed(X, y)
ctx = gp.ctx
nat = gp.natural(Theta)
delta, det_m, _, _ = gp._host_terms(nat, X, y, False)
thk = gp._kernel_theta(nat)
ctx.gp_upload(gp.desc, 0, delta, thk, want_grad=True)
ref = None
for it in range(8):
    prof = it >= 4
    ctx.prof_enable(prof)
    nrun = 1 if it % 2 == 0 else 3
    for _ in range(nrun):
        ctx.gp_run()
    r = ctx.gp_download()
    if ref is None: ref = r
    print(it, "prof", prof, "nrun", nrun, "status", np.unique(r["status"]), "dbeta", np.max(np.abs(r["beta"]-ref["beta"])),
          "dlogdet", np.max(np.abs(r["logdet"]-ref["logdet"])), "dgrad", np.max(np.abs(r["dtheta"]-ref["dtheta"])), flush=True)
    if prof: print({k: v for k, v in ctx.prof_read().items()})
