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
Np = (N + 127) // 128 * 128
ctx.gp_upload(gp.desc, 0, delta, thk, want_grad=bool(grad))
K, _ = ctx.gram(gp.desc, X, None, thk)
prev = None
for it in range(5):
    ctx.gp_run()
    r = ctx.gp_download()
    A = ctx.debug_read("gp_A", (B, Np, Np))
    Dinv = ctx.debug_read("gp_Dinv", (B, Np // 128, 128, 128))
    msg = []
    if not grad:
        for b in range(B):
            L = np.tril(A[b, :N, :N])
            msg.append("%.1e" % (np.max(np.abs(L @ L.T - K[b])) / np.max(np.abs(K[b]))))
    else:
        U = ctx.debug_read("gp_U", (B, Np, Np))
        for b in range(B):
            Ub = np.triu(U[b, :N, :N])
            Kinv = np.tril(A[b, :N, :N]); Kinv = Kinv + np.tril(Kinv, -1).T
            e1 = np.max(np.abs(Ub @ Ub.T @ K[b] - np.eye(N)))
            e2 = np.max(np.abs(Kinv @ K[b] - np.eye(N)))
            msg.append("U %.1e Kinv %.1e" % (e1, e2))
    d = "" if prev is None else " dA %.2e dDinv %.2e dbeta %.2e" % (np.max(np.abs(np.tril(A) - np.tril(prev[0]))), np.max(np.abs(Dinv - prev[1])), np.max(np.abs(r["beta"] - prev[2])))
    print(it, r["status"], msg, d, flush=True)
    prev = (A, Dinv, r["beta"])
