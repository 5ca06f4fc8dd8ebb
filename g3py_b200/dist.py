"""Exact GP across the GPUs of one node: 2-D block-cyclic Cholesky + solve behind the C ABI (g3_dist_*, include/g3b.h).

Host side only: builds the kernel descriptor, creates the communicator (g3py_b200/comm.py) and reports the device
times the library measured (max over ranks).  No torch; NCCL is called from libg3b.so.
The reference has no counterpart."""
import math

import numpy as np

from . import comm


def se_noise_desc(X):
    """SE (ARD) + Noise on the columns of X: the config-5 kernel."""
    import g3py_b200 as g3
    k = g3.SE(X) + g3.KernelNoise(name="Noise")
    reg = g3.Registry()
    k.check_dims(X)
    k.check_hypers("", reg)
    b = g3.DescBuilder(X.shape[1])
    k.compile(b, process_noise=True)         # the Noise leaf is the process noise: noise=False selectors drop it (elliptical.py:73-74)
    return b.finish()


def run_dist_cholesky(ctx, N, nb=1024, grid=None, theta=None, lookahead=True, ring=2, verify=0, seed=1234, X=None, y=None):
    """One timed distributed factorisation + solve of the SE(+noise) Gram matrix of the config-5 inputs on the
    communicator of `ctx` (comm.init first).  Returns a dict on every rank; times are device times, max over ranks.
    verify = number of probe vectors of the on-hardware residual check (0: skip)."""
    from . import workloads
    if X is None:
        X, y = workloads.c5_inputs(N)
    world = ctx.comm_size()
    Pr, Pc = grid if grid is not None else comm.grid_for(world)
    desc = se_noise_desc(X)
    th = np.array([1.0] * (1 + X.shape[1]) + [0.01]) if theta is None else np.asarray(theta, dtype=np.float64)
    ctx.set_data(X)
    ctx._data_tag = None
    f = ctx.dist_factor(desc, th, nb, Pr, Pc, lookahead=lookahead, ring=ring)
    s = ctx.dist_solve(y, want_u=False)
    n = float(N)
    out = {"N": N, "nb": nb, "n_gpus": world, "grid": [Pr, Pc], "ms_gram": f["ms_gram"], "ms_potrf": f["ms_potrf"],
           "ms_solve": s["ms_solve"], "tflops": n ** 3 / 3.0 / (f["ms_potrf"] * 1e-3) / 1e12, "logdet": f["logdet"],
           "beta": s["beta"], "logp": -0.5 * n * math.log(2 * math.pi) - 0.5 * s["beta"] - f["logdet"],
           "info": f["info"], "local_gib": f["local_gib"], "lookahead": bool(lookahead), "ring": ring}
    if verify:
        r = ctx.dist_residual(verify, seed)
        out["residual"] = [float(v) for v in r]
    return out
