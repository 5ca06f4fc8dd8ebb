import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads
N = int(sys.argv[1]); B = int(sys.argv[2]); grad = int(sys.argv[3]); runs = int(sys.argv[4])
X, y, Theta = workloads.c2_inputs(N, B)
gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X)); gp.observed(X, y)
ctx = gp.ctx
nat = gp.natural(Theta)
delta, det_m, _, _ = gp._host_terms(nat, X, y, False)
thk = gp._kernel_theta(nat)
ctx.gp_upload(gp.desc, 0, delta, thk, want_grad=bool(grad))
res = []
for it in range(runs):
    ctx.gp_run()
    r = ctx.gp_download()
    res.append(np.concatenate([r["beta"], r["logdet"]] + ([r["dtheta"].ravel()] if grad else [])))
res = np.array(res)
med = np.median(res, axis=0)
bad = [(it, np.nonzero(np.abs(res[it] - med) > 1e-13 * np.abs(med))[0].tolist()[:6]) for it in range(runs)]
print("G3_DBG", os.environ.get("G3_DBG"), N, B, grad, "bad:", [(i, b) for i, b in bad if b], flush=True)
