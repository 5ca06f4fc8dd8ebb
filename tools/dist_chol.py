"""Multi-GPU block-cyclic Cholesky driver (PyTorch-free: NCCL lives inside libg3b.so).

  python -m torch.distributed.run --nproc-per-node G tools/dist_chol.py N [nb] [--grid PRxPC] [--verify] [--check]
                                                                           [--no-lookahead] [--ring3] [--reps R] [--post M] [--grad]
  (any launcher that exports RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT works; G = 1: plain `python`)

--verify : on-hardware residual probe (4 vectors, L (L^T v) vs K v with K regenerated from X)
--check  : element-wise comparison of every local piece, u, beta, log-det with NumPy's Cholesky (N <= 8192)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from g3py_b200 import comm, workloads  # noqa: E402
from g3py_b200._cabi import Context  # noqa: E402
from g3py_b200.dist import run_dist_cholesky  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    N = int(args[0])
    nb = int(args[1]) if len(args) > 1 else 1024
    rank, world, local = comm.env_rank()
    grid = None
    reps = 2
    for i, a in enumerate(sys.argv):
        if a == "--grid":
            grid = tuple(int(v) for v in sys.argv[i + 1].lower().split("x"))
        if a == "--reps":
            reps = int(sys.argv[i + 1])
    ctx = Context(local)
    comm.init(ctx, rank, world)
    verify = 4 if "--verify" in sys.argv else 0
    r = None
    for rep in range(reps):                          # first pass warms allocations / NCCL channels
        r = run_dist_cholesky(ctx, N, nb=nb, grid=grid, lookahead="--no-lookahead" not in sys.argv,
                              ring=3 if "--ring3" in sys.argv else 2, verify=verify if rep == reps - 1 else 0)
    if "--check" in sys.argv:
        X, y = workloads.c5_inputs(N)
        d = ((X[:, None, :] - X[None, :, :]) ** 2 * 0.5).sum(-1)
        L = np.linalg.cholesky(np.exp(-d) + 0.01 * np.eye(N))
        Pr, Pc = r["grid"]
        p, q = rank % Pr, rank // Pr
        err = 0.0
        for J in range(q, N // nb, Pc):
            piece = ctx.dist_read_piece(J, nb)
            if piece is None:
                continue
            I0 = J + ((p - J % Pr) % Pr + Pr) % Pr
            for i in range(piece.shape[0] // nb):
                I = I0 + i * Pr
                got = piece[i * nb:(i + 1) * nb]
                if I == J:
                    got = np.tril(got)
                err = max(err, float(np.abs(got - L[I * nb:(I + 1) * nb, J * nb:(J + 1) * nb]).max()))
        s = ctx.dist_solve(y, want_u=True)
        uref = np.linalg.solve(L, y)
        err_u = float(np.abs(s["u"] - uref).max())
        ld = float(np.log(np.diag(L)).sum())
        assert err < 1e-10 and err_u < 1e-9, (rank, err, err_u)
        assert abs(s["beta"] - uref @ uref) < 1e-10 * (uref @ uref) and abs(r["logdet"] - ld) < 1e-10 * abs(ld), (r, ld)
        r["max_abs_err_vs_numpy"] = float(ctx.comm_allreduce([err], "max")[0])
        r["max_abs_err_u"] = float(ctx.comm_allreduce([err_u], "max")[0])
        if Pr == 1:                              # posterior and gradient from the distributed factor (1 x G grids)
            K = np.exp(-d) + 0.01 * np.eye(N)
            rng = np.random.default_rng(7)
            Xs = rng.uniform(0, N ** (1.0 / 3.0), size=(300, 3))
            ds = ((Xs[:, None, :] - X[None, :, :]) ** 2 * 0.5).sum(-1)
            Ks = np.exp(-ds)
            V = np.linalg.solve(L, Ks.T)
            m, v = ctx.dist_posterior(Xs, noise=False)
            e_m = float(np.abs(m - V.T @ uref).max() / np.abs(V.T @ uref).max())
            e_v = float(np.abs(v - np.maximum(1.0 - (V * V).sum(0), 0.0)).max())
            g = ctx.dist_grad(5, cfac=1.0)
            Kinv = np.linalg.inv(K)
            al = Kinv @ y
            W = 0.5 * (np.outer(al, al) - Kinv)
            E = np.exp(-d)
            want = [np.sum(W * E)] + [np.sum(W * E * (-(X[:, None, k] - X[None, :, k]) ** 2)) for k in range(3)] + [np.trace(W)]
            e_g = float(np.abs(g["dtheta"] - want).max() / np.abs(want).max())
            e_a = float(np.abs(g["ddelta"] + al).max() / np.abs(al).max())
            assert e_m < 1e-9 and e_v < 1e-9 and e_g < 1e-9 and e_a < 1e-9, (rank, e_m, e_v, e_g, e_a)
            r.update(post_mean_err=e_m, post_var_err=e_v, grad_err=e_g, alpha_err=e_a, ms_alpha=g["ms_alpha"],
                     ms_inverse=g["ms_inverse"], ms_contract=g["ms_contract"])
        print("rank", rank, "check ok", err, err_u, flush=True)
    for i, a in enumerate(sys.argv):             # --post M: distributed posterior moments at M test points (1 x G grid)
        if a == "--post":
            import time
            M = int(sys.argv[i + 1])
            Xs = np.random.default_rng(11).uniform(0, N ** (1.0 / 3.0), size=(M, 3))
            ctx.comm_barrier()
            t0 = time.perf_counter()
            m, v = ctx.dist_posterior(Xs, noise=False)
            ctx.comm_barrier()
            dt = time.perf_counter() - t0
            r.update(post_M=M, ms_posterior=1e3 * dt, post_tflops=float(N) ** 2 * M * 2 / 2 / dt / 1e12 * 1.0,
                     post_finite=bool(np.all(np.isfinite(m)) and np.all(v >= 0) and np.all(v <= 1.0 + 1e-9)))
    if "--grad" in sys.argv:                     # timing of the distributed gradient at full size (1 x G grid)
        g = ctx.dist_grad(5, cfac=1.0, want_ddelta=False)
        n = float(N)
        r.update(ms_alpha=g["ms_alpha"], ms_inverse=g["ms_inverse"], ms_contract=g["ms_contract"], dtheta=[float(v) for v in g["dtheta"]],
                 grad_tflops=2.0 * n ** 3 / 3.0 / ((g["ms_inverse"] + g["ms_contract"]) * 1e-3) / 1e12)
    if rank == 0:
        print(json.dumps(r), flush=True)
    ctx.dist_free()
    ctx.comm_destroy()


if __name__ == "__main__":
    main()
