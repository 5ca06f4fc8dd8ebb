import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads
N = int(sys.argv[1]); B = int(sys.argv[2]); grad = int(sys.argv[3])
X, y, Theta = workloads.c2_inputs(N, B)
gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X)); gp.observed(X, y)
ctx = gp.ctx
nat = gp.natural(Theta)
delta, det_m, _, _ = gp._host_terms(nat, X, y, False)
thk = gp._kernel_theta(nat)
single = []
for b in range(B):
    r = ctx.gp_logp_grad(gp.desc, 0, delta[b:b+1] if delta.ndim == 2 else delta, thk[b:b+1], want_grad=False)
    single.append((r["beta"][0], r["logdet"][0]))
single = np.array(single)
ctx.gp_upload(gp.desc, 0, delta, thk, want_grad=bool(grad))
for it in range(4):
    ctx.gp_run()
    r = ctx.gp_download()
    eb = np.abs(r["beta"] - single[:, 0]) / np.abs(single[:, 0]); el = np.abs(r["logdet"] - single[:, 1]) / np.abs(single[:, 1])
    print(N, B, grad, "run", it, "status", np.unique(r["status"]), "bad beta items", np.nonzero(eb > 1e-12)[0].tolist(), "bad logdet items", np.nonzero(el > 1e-12)[0].tolist(), "max", eb.max(), el.max(), flush=True)
