"""Exact GP too large / too slow for one GPU: block-cyclic Cholesky across the GPUs of one node.

Not in the reference (its only parallelism is `multiprocessing.Pool.map` over chains,
g3py/processes/stochastic.py:775-783); this is SURVEY §2.2 K17 / §8e.

Layout.  2-D block-cyclic with block nb (default 1024 = 8 tile columns) on a 1 x G process grid: panel J
(columns [J nb, (J+1) nb)) lives on rank J mod G.  Only the lower trapezoid of each panel is stored, as its
own contiguous tall matrix (rows J nb .. N, leading dimension nb), so (a) a rank holds N^2/(2G) doubles,
(b) every panel is already the K-contiguous operand the fp64 GEMM wants and is broadcast in place (no
packing), (c) each rank generates its own panels from the replicated X (MBs) - K never crosses a link.
NVSwitch gives every GPU full bandwidth to every peer, so the column broadcast of a 1 x G grid (N^2/2
doubles received per rank in total, ~0.2 s at N = 131072) hides completely under the trailing update; the
row broadcasts of a Pr > 1 grid would only add latency.

Schedule (right-looking between panels, left-looking inside one, with look-ahead):
    owner(J+1): update panel J+1 with panel J -> factor panel J+1 -> start its broadcast (async)
    everyone  : update the remaining local panels with panel J while panel J+1 is in flight
The panel primitives are libg3b.so kernels on device pointers (`g3_dev_*`); torch.distributed (NCCL) is
used only for the broadcast and torch only for device memory and streams.
"""
import math

import numpy as np


def panel_owner(J, world):
    return J % world


def local_panels(n_panels, rank, world):
    return list(range(rank, n_panels, world))


class LibBackend:
    """Panel primitives through the C ABI on torch-owned device memory."""

    def __init__(self, ctx, desc, theta, diag_shift, torch, stream):
        self.ctx, self.desc, self.theta, self.shift, self.torch = ctx, desc, np.asarray(theta, dtype=np.float64), diag_shift, torch
        self.stream = stream
        self.panel_stream = torch.cuda.Stream(priority=-1)     # critical path: next panel's update, factor, broadcast
        ctx.set_stream(stream.cuda_stream)

    # two-stream look-ahead support: `with be.on_panel_stream():` routes library launches AND torch ops there
    def on_panel_stream(self):
        be = self

        class _Ctx:
            def __enter__(self_):
                self_.cm = be.torch.cuda.stream(be.panel_stream)
                self_.cm.__enter__()
                be.ctx.set_stream(be.panel_stream.cuda_stream)

            def __exit__(self_, *a):
                be.ctx.set_stream(be.stream.cuda_stream)
                self_.cm.__exit__(*a)
        return _Ctx()

    def fence_main_to_panel(self):
        self.panel_stream.wait_stream(self.stream)

    def fence_panel_to_main(self):
        self.stream.wait_stream(self.panel_stream)

    def record_panel_event(self):
        return self.panel_stream.record_event()

    def main_wait_event(self, ev):
        self.stream.wait_event(ev)

    def alloc(self, n):
        return self.torch.empty(n, dtype=self.torch.float64, device="cuda")

    def scalars(self):
        return (self.torch.zeros(1, dtype=self.torch.float64, device="cuda"),
                self.torch.zeros(1, dtype=self.torch.int32, device="cuda"))

    def gram(self, out, row0, col0, rows, cols):
        self.ctx.dev_gram_block(self.desc, self.theta, row0, col0, rows, cols, self.shift, out.data_ptr(), cols)

    def factor(self, P, rows, nb, logdet, info, dinv=None):
        self.ctx.dev_potrf_panel(P.data_ptr(), rows, nb, logdet.data_ptr(), info.data_ptr(),
                                 dinv.data_ptr() if dinv is not None else 0)

    def trsv(self, P, rows, nb, dinv, r, u, beta):
        self.ctx.dev_trsv_panel(P.data_ptr(), rows, nb, dinv.data_ptr(), r.data_ptr(), u.data_ptr(), beta.data_ptr())

    def update(self, P, rows_p, nb, row_off, D, rows_d):
        self.ctx.dev_syrk_panel(P.data_ptr(), rows_p, nb, row_off, D.data_ptr(), rows_d)

    def close(self):
        self.ctx.set_stream(0)


class DistCholesky:
    """Block-cyclic lower Cholesky of K = cov(X) (+ shift on the diagonal) over `world` ranks.

    backend: object with alloc / scalars / gram / factor / update (LibBackend on GPUs; the CPU tests pass a
    NumPy double).  dist: torch.distributed module or None (single rank)."""

    def __init__(self, N, nb, rank, world, backend, dist=None, lookahead=True):
        if N % nb or nb % 128:
            raise ValueError("N must be a multiple of nb, nb a multiple of 128")
        self.N, self.nb, self.rank, self.world = N, nb, rank, world
        self.nP = N // nb
        self.be, self.dist, self.lookahead = backend, dist, lookahead
        self.mine = local_panels(self.nP, rank, world)
        self.rows = {J: N - J * nb for J in range(self.nP)}
        sizes = [self.rows[J] * nb for J in self.mine]
        self.store = backend.alloc(max(sum(sizes), 1))
        self.off = {}
        o = 0
        for J, sz in zip(self.mine, sizes):
            self.off[J] = o
            o += sz
        self.pbuf = [backend.alloc(N * nb), backend.alloc(N * nb)] if world > 1 else None
        self.logdet, self.info = backend.scalars()
        self.dinv = backend.alloc(max(len(self.mine), 1) * nb * 128)       # 128x128 block inverses of my panels
        self.dinv_of = {J: self.dinv[i * nb * 128:(i + 1) * nb * 128] for i, J in enumerate(self.mine)}

    def panel(self, J):
        return self.store[self.off[J]: self.off[J] + self.rows[J] * self.nb]

    def local_bytes(self):
        return 8 * sum(self.rows[J] * self.nb for J in self.mine)

    def build(self):
        """Each rank generates its own panels of K from the replicated X."""
        for J in self.mine:
            self.be.gram(self.panel(J), J * self.nb, J * self.nb, self.rows[J], self.nb)

    def _bcast(self, J):
        o = panel_owner(J, self.world)
        if self.world == 1:
            return self.panel(J), None
        t = self.panel(J) if self.rank == o else self.pbuf[J % 2][: self.rows[J] * self.nb]
        work = self.dist.broadcast(t, src=o, async_op=True)
        return t, work

    def _update(self, PJ, J, Jp):
        """local panel Jp (> J) -= PJ[rows of Jp] PJ[cols of Jp]^T"""
        self.be.update(PJ, self.rows[J], self.nb, (Jp - J) * self.nb, self.panel(Jp), self.rows[Jp])

    def factor(self):
        if self.lookahead and hasattr(self.be, "on_panel_stream"):
            return self._factor_two_streams()
        return self._factor_one_stream()

    def _factor_two_streams(self):
        """Look-ahead with the critical path on its own (high-priority) stream: while the main stream applies
        panel J to the rank's remaining panels, the panel stream updates, factors and broadcasts panel J+1, so
        neither the 1-CTA diagonal kernels of the panel factorisation nor the broadcast ever idle the GPU."""
        nb, nP, me, be = self.nb, self.nP, self.rank, self.be
        be.fence_main_to_panel()
        with be.on_panel_stream():
            if panel_owner(0, self.world) == me:
                be.factor(self.panel(0), self.rows[0], nb, self.logdet, self.info, self.dinv_of[0])
            cur = self._bcast(0)
            ev = be.record_panel_event()              # panel 0 factored (owner) / enqueued
        for J in range(nP):
            PJ, work = cur
            nxt = None
            done_early = None
            be.main_wait_event(ev)                    # main stream: panel J is factored (matters on its owner)
            with be.on_panel_stream():
                if work is not None:
                    work.wait()                       # panel stream waits for panel J
                be.fence_main_to_panel()              # ... and for the main stream's updates of iteration J-1
                if J + 1 < nP:
                    if panel_owner(J + 1, self.world) == me:
                        self._update(PJ, J, J + 1)
                        done_early = J + 1
                        be.factor(self.panel(J + 1), self.rows[J + 1], nb, self.logdet, self.info, self.dinv_of[J + 1])
                    nxt = self._bcast(J + 1)
                    ev = be.record_panel_event()
            if work is not None:
                work.wait()                           # main stream waits for the broadcast of panel J as well
            for Jp in self.mine:
                if Jp > J and Jp != done_early:
                    self._update(PJ, J, Jp)
            cur = nxt
        be.fence_panel_to_main()
        return self

    def _factor_one_stream(self):
        nb, nP, me = self.nb, self.nP, self.rank
        if panel_owner(0, self.world) == me:
            self.be.factor(self.panel(0), self.rows[0], nb, self.logdet, self.info, self.dinv_of[0])
        cur = self._bcast(0)
        for J in range(nP):
            PJ, work = cur
            if work is not None:
                work.wait()
            nxt = None
            done_early = None
            if J + 1 < nP:
                if panel_owner(J + 1, self.world) == me:
                    if self.lookahead or self.world == 1:
                        self._update(PJ, J, J + 1)
                        done_early = J + 1
                        self.be.factor(self.panel(J + 1), self.rows[J + 1], nb, self.logdet, self.info, self.dinv_of[J + 1])
                if self.lookahead:
                    nxt = self._bcast(J + 1)
            for Jp in self.mine:
                if Jp > J and Jp != done_early:
                    self._update(PJ, J, Jp)
            if J + 1 < nP and not self.lookahead:
                if panel_owner(J + 1, self.world) == me and self.world > 1:
                    self.be.factor(self.panel(J + 1), self.rows[J + 1], nb, self.logdet, self.info, self.dinv_of[J + 1])
                nxt = self._bcast(J + 1)
            cur = nxt
        return self

    def solve(self, delta, make_zero, all_reduce_sum):
        """u = L^-1 delta across the ranks; returns (u pieces {J: tensor[nb]}, beta tensor (local part)).

        Every rank carries a length-N residual `c` (rank 0 starts with delta, the others with 0); the true
        residual of panel J is the sum over ranks of c[J rows], obtained with ONE nb-sized all-reduce per panel;
        the owner then solves with its panel and pushes the update into its own `c` (rows below).  No panel
        data moves; L is read exactly once."""
        nb = self.nb
        c = delta.clone() if self.rank == 0 else make_zero(self.N)
        beta = make_zero(1)
        u = {}
        for J in range(self.nP):
            seg = c[J * nb:(J + 1) * nb]
            if self.world > 1:
                all_reduce_sum(seg)                      # everyone now holds the true residual of panel J
            if panel_owner(J, self.world) == self.rank:
                u[J] = make_zero(nb)
                self.be.trsv(self.panel(J), self.rows[J], nb, self.dinv_of[J], c[J * nb:], u[J], beta)
            else:
                seg.zero_()                              # rows of a finished panel are never read again
        return u, beta

    def flops(self):
        return float(self.N) ** 3 / 3.0


def run_dist_cholesky(N, D=3, nb=1024, theta=None, lookahead=True, seed=5, verify=False):
    """One timed distributed factorisation of the SE(+noise) Gram of the config-5 inputs.  Launch one
    process per GPU (torchrun); returns a dict on every rank (times are max over ranks)."""
    import os
    import torch
    import torch.distributed as dist
    import g3py_b200 as g3
    from g3py_b200 import workloads
    from g3py_b200.processes import get_context
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    X, y = workloads.c5_inputs(N)
    k = g3.SE(X) + g3.KernelNoise(name="Noise")
    reg = g3.Registry()
    k.check_dims(X)
    k.check_hypers("", reg)
    b = g3.DescBuilder(D)
    k.compile(b)
    desc = b.finish()
    th = np.array([1.0, 1.0, 1.0, 1.0, 0.01]) if theta is None else np.asarray(theta, dtype=np.float64)
    ctx = get_context(local)
    ctx.set_data(X)
    ctx._data_tag = None
    stream = torch.cuda.Stream()
    be = LibBackend(ctx, desc, th, 0.0, torch, stream)
    out = {}
    try:
        with torch.cuda.stream(stream):
            ch = DistCholesky(N, nb, rank, world, be, dist if world > 1 else None, lookahead=lookahead)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e[0].record()
            ch.build()
            e[1].record()
            ch.factor()
            e[2].record()
            delta = torch.from_numpy(np.ascontiguousarray(y)).to("cuda")
            zeros = lambda n: torch.zeros(n, dtype=torch.float64, device="cuda")
            upieces, beta = ch.solve(delta, zeros, (lambda t_: dist.all_reduce(t_)) if world > 1 else (lambda t_: None))
            e[3].record()
            torch.cuda.synchronize()
            t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(beta)
            ld = ch.logdet.clone()
            info = ch.info.clone().to(torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(ld, op=dist.ReduceOp.SUM)
                dist.all_reduce(info, op=dist.ReduceOp.MAX)
            ms_gram, ms_potrf, ms_solve = (float(v) for v in t.cpu())
            n = float(N)
            logdet, betav = float(ld.item()), float(beta.item())
            out = {"N": N, "nb": nb, "n_gpus": world, "ms_gram": ms_gram, "ms_potrf": ms_potrf, "ms_solve": ms_solve,
                   "tflops": ch.flops() / (ms_potrf * 1e-3) / 1e12, "logdet": logdet, "beta": betav,
                   "logp": -0.5 * n * math.log(2 * math.pi) - 0.5 * betav - logdet,      # exact-constant Gaussian logp, zero mean
                   "info": int(info.item()), "local_gib": ch.local_bytes() / 2 ** 30, "lookahead": bool(lookahead)}
            if verify:
                out["panels"] = {J: ch.panel(J).cpu().numpy().reshape(ch.rows[J], nb) for J in ch.mine}
                out["u"] = {J: v.cpu().numpy() for J, v in upieces.items()}
    finally:
        be.close()
    return out
