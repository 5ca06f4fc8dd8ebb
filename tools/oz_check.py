"""int8 tensor-core (Ozaki) panel updates against the DMMA path: factor element-wise, logp / gradient, then timing of the
bench configuration in both modes.   python tools/oz_check.py [--time]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import g3py_b200 as g3  # noqa: E402
from g3py_b200 import workloads  # noqa: E402


def factor(ctx, gp, Theta, mode, min_k):
    ctx.set_gemm_mode(mode, min_k)
    nat = gp.natural(Theta)
    delta, _, _, _ = gp._host_terms(nat, gp.inputs, gp.outputs, False)
    thk = gp._kernel_theta(nat)
    l0 = ctx.ozaki_launch_count()
    r = ctx.gp_logp_grad(gp.desc, 0, np.array(delta), thk, want_grad=False)
    N = len(gp.outputs)
    Np = (N + 127) // 128 * 128
    L = ctx.debug_read("gp_A", (len(Theta), Np, Np))
    return r, np.tril(L), ctx.ozaki_launch_count() - l0


for N, B in ((2048, 12), (1920, 9), (1100, 10)):
    X, y, Theta = workloads.c2_inputs(N, B)
    gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X))
    gp.observed(X, y)
    ctx = gp.ctx
    r0, L0, n0 = factor(ctx, gp, Theta, "dmma", 0)
    r1, L1, n1 = factor(ctx, gp, Theta, "ozaki", 256)
    err = float(np.max(np.abs(L1 - L0)) / np.max(np.abs(L0)))
    eb = float(np.max(np.abs(r1["beta"] - r0["beta"]) / np.abs(r0["beta"])))
    el = float(np.max(np.abs(r1["logdet"] - r0["logdet"]) / np.abs(r0["logdet"])))
    print("N=%d B=%d ozaki launches %d (dmma %d)  max|dL|/max|L| %.3e  beta %.3e  logdet %.3e  status %s" % (N, B, n1, n0, err, eb, el, r1["status"].tolist()), flush=True)
    assert n0 == 0 and n1 > 0 and err < 1e-13 and eb < 1e-12 and el < 1e-13, "OZAKI MISMATCH"
    lp0, g0, _ = (ctx.set_gemm_mode("dmma"), gp.logp_dlogp_batch(Theta))[1]
    lp1, g1, _ = (ctx.set_gemm_mode("ozaki", 256), gp.logp_dlogp_batch(Theta))[1]
    assert np.max(np.abs(lp1 - lp0) / np.abs(lp0)) < 1e-11 and np.max(np.abs(g1 - g0)) / np.max(np.abs(g0)) < 1e-9
print("OZ_CHECK_OK", flush=True)

if "--time" in sys.argv:
    X, y, Theta = workloads.c2_inputs(4096, 64)
    gp = g3.GP(X, g3.Bias(), g3.SE(X) + g3.MAT52(X))
    gp.observed(X, y)
    ctx = gp.ctx
    nat = gp.natural(Theta)
    delta, _, _, _ = gp._host_terms(nat, X, y, False)
    thk = gp._kernel_theta(nat)
    for mode, mk in (("dmma", 0), ("ozaki", 512), ("ozaki", 768), ("ozaki", 1024), ("ozaki", 1536), ("ozaki", 2048)):
        ctx.set_gemm_mode(mode, mk)
        ctx.gp_upload(gp.desc, 0, np.array(delta), thk, want_grad=True)
        for _ in range(2):
            ctx.gp_run()
        ctx.sync()
        ctx.timer_begin()
        for _ in range(3):
            ctx.gp_run()
        ms = ctx.timer_end() / 3
        res = ctx.gp_download()
        print("mode %s min_k %d: %.2f ms/step = %.1f evals/s  status ok %s" % (mode, mk, ms, 64 / ms * 1e3, bool(np.all(res["status"] == 0))), flush=True)
        if mode == "dmma":
            ref = res
        else:
            print("   max rel diff dtheta vs dmma %.3e, logdet %.3e" % (float(np.max(np.abs(res["dtheta"] - ref["dtheta"])) / np.max(np.abs(ref["dtheta"]))),
                                                                      float(np.max(np.abs(res["logdet"] - ref["logdet"]) / np.abs(ref["logdet"])))), flush=True)
