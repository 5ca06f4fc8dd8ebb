import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads
N = int(sys.argv[1]); B = int(sys.argv[2]); grad = int(sys.argv[3]); nb = min(B, 3)
X, y, Theta = workloads.c2_inputs(N, B)
gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X)); gp.observed(X, y)
ctx = gp.ctx
nat = gp.natural(Theta)
delta, det_m, _, _ = gp._host_terms(nat, X, y, False)
thk = gp._kernel_theta(nat)
Np = (N + 127) // 128 * 128
ctx.gp_upload(gp.desc, 0, delta, thk, want_grad=bool(grad))
prev = None
for it in range(4):
    ctx.gp_run()
    r = ctx.gp_download()
    A = ctx.debug_read("gp_A", (nb, Np, Np))
    Dinv = ctx.debug_read("gp_Dinv", (nb, Np // 128, 128, 128))
    d = ""
    if prev is not None:
        dA = np.abs(np.tril(A) - np.tril(prev[0]))
        bad = np.argwhere(dA.reshape(nb, Np // 128, 128, Np // 128, 128).max(axis=(2, 4)) > 0)
        d = " dA %.2e dDinv %.2e dbeta %.2e dlogdet %.2e badtiles(first 12 of %d) %s" % (dA.max(), np.max(np.abs(Dinv - prev[1])), np.max(np.abs(r["beta"] - prev[2])), np.max(np.abs(r["logdet"] - prev[3])), len(bad), bad[:12].tolist())
    print(N, B, grad, it, np.unique(r["status"]), d, flush=True)
    prev = (A, Dinv, r["beta"], r["logdet"])
