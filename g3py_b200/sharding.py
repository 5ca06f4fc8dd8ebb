"""theta-batch / chain / test-point sharding across GPUs.

The axes of the hot path that shard naturally at N=4096 are "independent theta evaluations" (emcee walkers,
fixed-chain rows, multi-start points, MCMC chains) and "independent test points": the reference fans them out with
`multiprocessing.Pool.map` over chain groups (g3py/processes/stochastic.py:775-783).  Here every rank (one process per
GPU) evaluates a contiguous slice on its own device; X, y are replicated (KBs..MBs).  No collective sits on the data
path; the results (8*B*(P+1) bytes) are gathered with ONE all-gather - through the NCCL communicator libg3b.so owns
(`LibGroup`, g3_comm_allgather: no torch), or through a `torch.distributed` group the caller already has (`TorchGroup`;
gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(B, rank, world):
    """Contiguous, balanced split of B rows: first (B % world) ranks get one extra row."""
    base, extra = divmod(int(B), int(world))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


class LibGroup:
    """The communicator created by g3py_b200.comm.init on a device context (NCCL inside libg3b.so)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.rank, self.world = ctx.comm_rank(), ctx.comm_size()

    def allgather(self, buf):
        return self.ctx.comm_allgather(buf)


class TorchGroup:
    """A torch.distributed process group (gloo on CPU, nccl on GPUs) for callers that live in one already."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def allgather(self, buf):
        import torch
        import torch.distributed as dist
        device = self.device
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
        mine = torch.from_numpy(np.ascontiguousarray(buf)).to(device)
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        return np.stack([p.cpu().numpy() for p in parts])


def resolve_group(process, group=None):
    """LibGroup when the process' device context carries a communicator, a TorchGroup when torch.distributed is
    initialised, None for a single rank."""
    if group is not None:
        return group if hasattr(group, "allgather") else TorchGroup(group)
    ctx = process.ctx
    if hasattr(ctx, "comm_size") and ctx.comm_size() > 1:
        return LibGroup(ctx)
    import sys
    dist = sys.modules.get("torch.distributed")
    if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return TorchGroup()
    return None


def _gather_rows(grp, total, local, width):
    """all-gather of per-rank row blocks of unequal height: padded to the largest shard, reassembled in rank order."""
    rows = max(shard_bounds(total, r, grp.world)[1] - shard_bounds(total, r, grp.world)[0] for r in range(grp.world))
    buf = np.zeros((rows, width))
    buf[:local.shape[0]] = local
    parts = grp.allgather(buf)
    out = np.empty((total, width))
    for r in range(grp.world):
        a, b = shard_bounds(total, r, grp.world)
        out[a:b] = parts[r][: b - a]
    return out


def logp_dlogp_batch_sharded(process, Theta, group=None):
    """Evaluate `process.logp_dlogp_batch` on this rank's rows of Theta and all-gather (logp, dlogp).
    Returns the full (B,) and (B, P) arrays on every rank."""
    Theta = np.atleast_2d(np.asarray(Theta, dtype=np.float64))
    B, P = Theta.shape
    grp = resolve_group(process, group)
    if grp is None or grp.world == 1:
        lp, g, _ = process.logp_dlogp_batch(Theta)
        return lp, g
    lo, hi = shard_bounds(B, grp.rank, grp.world)
    local = np.zeros((hi - lo, P + 1))
    if hi > lo:
        lp, g, _ = process.logp_dlogp_batch(Theta[lo:hi])
        local[:, 0] = lp
        local[:, 1:] = g
    out = _gather_rows(grp, B, local, P + 1)
    return out[:, 0], out[:, 1:]


def logp_batch_sharded(process, Theta, group=None):
    """logp only (ensemble samplers): rows of Theta split over the ranks, one all-gather of B doubles."""
    Theta = np.atleast_2d(np.asarray(Theta, dtype=np.float64))
    B = Theta.shape[0]
    grp = resolve_group(process, group)
    if grp is None or grp.world == 1:
        return process.logp_batch(Theta)
    lo, hi = shard_bounds(B, grp.rank, grp.world)
    local = np.zeros((hi - lo, 1))
    if hi > lo:
        local[:, 0] = process.logp_batch(Theta[lo:hi])
    return _gather_rows(grp, B, local, 1)[:, 0]


def predict_sharded(process, params, space, group=None, array=True, noise=False):
    """Posterior mean / variance at many test points with the rows of `space` split over the ranks (SURVEY §8e,
    "posterior at large M": independent test tiles, the factor of K is replicated - every rank factors the same
    K - so no collective sits on the data path).  Returns the full (mean, variance) on every rank."""
    space = np.asarray(space, dtype=np.float64)
    if space.ndim < 2:
        space = space.reshape(len(space), 1)
    M = space.shape[0]
    grp = resolve_group(process, group)
    if grp is None or grp.world == 1:
        v = process.predict(params, space=space, array=array, var=True, std=False, noise=noise)
        return v["mean"], v["variance"]
    lo, hi = shard_bounds(M, grp.rank, grp.world)
    local = np.zeros((hi - lo, 2))
    if hi > lo:
        v = process.predict(params, space=space[lo:hi], array=array, var=True, std=False, noise=noise)
        local[:, 0] = v["mean"]
        local[:, 1] = v["variance"]
    out = _gather_rows(grp, M, local, 2)
    return out[:, 0], out[:, 1]


def chains_sharded(process, start, samples, chains_per_rank=1, group=None, seed=0, **hmc_kwargs):
    """MCMC chains, `chains_per_rank` per GPU (BASELINE config 3: "8 chains, one per GPU"): every rank advances its own
    chains with `process.sample_hmc` (rank-dependent seed), no communication while sampling; the chains are gathered at
    the end.  Returns (chain [samples, world * chains_per_rank, P], logp [samples, world * chains_per_rank])."""
    grp = resolve_group(process, group)
    rank, world = (0, 1) if grp is None else (grp.rank, grp.world)
    ch, lp, _ = process.sample_hmc(start=start, samples=samples, chains=chains_per_rank, seed=seed + 1000 * rank, **hmc_kwargs)
    if world == 1:
        return ch, lp
    P = ch.shape[2]
    flat = np.concatenate([ch.reshape(samples, -1), lp], axis=1)             # [samples, c*P + c]
    parts = grp.allgather(flat)
    chain = np.concatenate([p[:, :chains_per_rank * P].reshape(samples, chains_per_rank, P) for p in parts], axis=1)
    logp = np.concatenate([p[:, chains_per_rank * P:] for p in parts], axis=1)
    return chain, logp
