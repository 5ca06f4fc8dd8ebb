"""theta-batch sharding across GPUs.

The only axis of the hot path that shards naturally at N=4096 is "independent theta evaluations"
(emcee walkers, fixed-chain rows, multi-start points): the reference fans them out with
`multiprocessing.Pool.map` over chain groups (g3py/processes/stochastic.py:775-783).  Here every rank
(one process per GPU) evaluates a contiguous slice of the rows on its own device; X, y are replicated
(KBs..MBs).  No collective sits on the data path; the results (8*B*(P+1) bytes) are gathered with one
`all_gather` over `torch.distributed` (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(B, rank, world):
    """Contiguous, balanced split of B rows: first (B % world) ranks get one extra row."""
    base, extra = divmod(int(B), int(world))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def logp_dlogp_batch_sharded(process, Theta, group=None, device=None):
    """Evaluate `process.logp_dlogp_batch` on this rank's rows of Theta and all-gather (logp, dlogp).

    Returns the full (B,) and (B, P) arrays on every rank.  `device`: torch device used for the
    collective buffers (cuda for NCCL, cpu for gloo)."""
    import torch
    import torch.distributed as dist
    Theta = np.atleast_2d(np.asarray(Theta, dtype=np.float64))
    B, P = Theta.shape
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        lp, g, _ = process.logp_dlogp_batch(Theta)
        return lp, g
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(B, rank, world)
    rows = max(shard_bounds(B, r, world)[1] - shard_bounds(B, r, world)[0] for r in range(world))
    buf = np.zeros((rows, P + 1))
    if hi > lo:
        lp, g, _ = process.logp_dlogp_batch(Theta[lo:hi])
        buf[:hi - lo, 0] = lp
        buf[:hi - lo, 1:] = g
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(buf).to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    out = np.empty((B, P + 1))
    for r in range(world):
        a, b = shard_bounds(B, r, world)
        out[a:b] = parts[r][: b - a].cpu().numpy()
    return out[:, 0], out[:, 1:]


def predict_sharded(process, params, space, group=None, device=None, array=True, noise=False):
    """Posterior mean / variance at many test points with the rows of `space` split over the ranks (SURVEY §8e,
    "posterior at large M": independent test tiles, the factor of K is replicated - every rank factors the same
    K - so no collective sits on the data path).  Returns the full (mean, variance) on every rank."""
    import torch
    import torch.distributed as dist
    space = np.asarray(space, dtype=np.float64)
    if space.ndim < 2:
        space = space.reshape(len(space), 1)
    M = space.shape[0]
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        v = process.predict(params, space=space, array=array, var=True, std=False, noise=noise)
        return v["mean"], v["variance"]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(M, rank, world)
    rows = max(shard_bounds(M, r, world)[1] - shard_bounds(M, r, world)[0] for r in range(world))
    buf = np.zeros((rows, 2))
    if hi > lo:
        v = process.predict(params, space=space[lo:hi], array=array, var=True, std=False, noise=noise)
        buf[:hi - lo, 0] = v["mean"]
        buf[:hi - lo, 1] = v["variance"]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(buf).to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    out = np.empty((M, 2))
    for r in range(world):
        a, b = shard_bounds(M, r, world)
        out[a:b] = parts[r][: b - a].cpu().numpy()
    return out[:, 0], out[:, 1]
