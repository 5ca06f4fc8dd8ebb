"""Multi-GPU block-cyclic Cholesky driver (PyTorch-free: NCCL lives inside libg3b.so).

  python -m torch.distributed.run --nproc-per-node G tools/dist_chol.py N [nb] [--grid PRxPC] [--verify] [--check]
                                                                           [--no-lookahead] [--ring2] [--reps R]
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
                              ring=2 if "--ring2" in sys.argv else 3, verify=verify if rep == reps - 1 else 0)
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
        print("rank", rank, "check ok", err, err_u, flush=True)
    if rank == 0:
        print(json.dumps(r), flush=True)
    ctx.dist_free()
    ctx.comm_destroy()


if __name__ == "__main__":
    main()
