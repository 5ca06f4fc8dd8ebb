"""Small end-to-end case for compute-sanitizer (racecheck / memcheck): exercises the TMA ring of the fp64 GEMM, the
split-K path (single matrix), the batched left-looking schedule (B > 8), the jitter ladder, the fast2 Gram kernels,
the posterior and the block-cyclic single-GPU Cholesky + gradient.  Sizes are tiny: the tools slow kernels down 10-100x.

    compute-sanitizer --tool racecheck python tools/sanitize_case.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import g3py_b200 as g3  # noqa: E402
from g3py_b200 import workloads  # noqa: E402
from g3py_b200.dist import se_noise_desc  # noqa: E402

X, y, Theta = workloads.c2_inputs(600, 12)
gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X))
gp.observed(X, y)
lp, g, info = gp.logp_dlogp_batch(Theta)                 # batched, left-looking, 2 stream groups
lp1, g1, _ = gp.logp_dlogp_batch(Theta[:1])              # single matrix: blocked + look-ahead + split-K + pipelined trtri
assert np.all(np.isfinite(lp)) and np.all(np.isfinite(g)) and abs(lp1[0] - lp[0]) < 1e-9 * abs(lp[0])
pr = gp.predict(Theta[0], space=X[:200] + 0.01, array=True, var=True)
assert np.all(np.isfinite(pr["mean"]))
Xd = X.copy()
Xd[300:] = Xd[:300]                                      # singular without noise: ladder
gd = g3.GP(Xd, g3.Zero(), g3.SE(Xd), noisy=False)
gd.observed(Xd, y)
v = gd.logp(gd.dict_to_array(gd.params_default), array=True)
assert np.isfinite(v)
Xc, yc = workloads.c5_inputs(1024)
ctx = g3.Context(0)
ctx.set_data(Xc)
desc = se_noise_desc(Xc)
th = np.array([1.0, 1.0, 1.0, 1.0, 0.01])
f = ctx.dist_factor(desc, th, 256, 1, 1)
s = ctx.dist_solve(yc)
m, vv = ctx.dist_posterior(Xc[:130] + 0.05)
gr = ctx.dist_grad(5)
assert f["info"] == 0 and np.all(np.isfinite(gr["dtheta"]))
ctx.close()
print("SANITIZE_CASE_OK", float(lp[0]), float(v), f["logdet"], s["beta"])
