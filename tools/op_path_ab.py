"""Value-then-gradient through the Op boundary (GPLogpOp.perform, then GPLogpGradOp.perform on the same inputs) with and
without the speculative U = L^-T behind the logp-only factorisation (g3_set_speculate_grad)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import g3py_b200 as g3
from g3py_b200 import workloads
import bench

for N in (1024, 2048, 4096):
    X3, y, Th = workloads.c2_inputs(N, 8)
    gp = g3.GP(X3, g3.Bias(), g3.SE(X3) + g3.MAT52(X3)); gp.observed(X3, y)
    ctx = gp.ctx
    real = ctx.set_speculate_grad
    out = {}
    for mode in (0, 1):
        ctx.set_speculate_grad = real if mode else (lambda on: None)
        r = bench.op_path_rate(gp, Th, 4)
        out[mode] = r
    ctx.set_speculate_grad = real
    print("N=%d  op path: %.1f evals/s (%.2f ms) without, %.1f evals/s (%.2f ms) with speculation; resumed %d / %d"
          % (N, out[0]["value"], out[0]["ms_per_eval"], out[1]["value"], out[1]["ms_per_eval"],
             out[1]["resumed_from_resident_factor"], out[1]["evals"]), flush=True)
