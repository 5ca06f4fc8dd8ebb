"""Process-group plumbing for the multi-GPU paths, PyTorch-free.

One process per GPU (the launcher - `python -m torch.distributed.run`, mpirun, a shell loop - only has to export
RANK / WORLD_SIZE / LOCAL_RANK and, for the rendezvous, MASTER_PORT).  The NCCL communicator lives inside libg3b.so
(`g3_comm_init`, include/g3b.h); all this module does is hand rank 0's 128-byte NCCL id to the other ranks of the
node through a file in a directory every rank can see (single-node contract: NVLink / NVSwitch domain).

The reference has nothing comparable: its only parallelism is `multiprocessing.Pool.map` over chain groups
(g3py/processes/stochastic.py:775-783).
"""
import os
import time

from . import _cabi as cabi

_SEQ = [0]


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def _rdv_path(seq):
    explicit = os.environ.get("G3_RDV_FILE")
    if explicit:
        return "%s.%d" % (explicit, seq)
    # all ranks of one launch share the launcher as parent; MASTER_PORT separates concurrent launches
    tag = "%s_%d_%s" % (os.environ.get("MASTER_PORT", "0"), os.getppid(), os.environ.get("TORCHELASTIC_RUN_ID", "none"))
    return os.path.join(os.environ.get("G3_RDV_DIR", "/tmp"), "g3b_rdv_%s.%d" % (tag, seq))


def exchange_id(rank, world, timeout=300.0):
    """Rank 0 creates the NCCL id and publishes it (atomic rename); the others wait for the file."""
    seq = _SEQ[0]
    _SEQ[0] += 1
    path = _rdv_path(seq)
    if rank == 0:
        uid = cabi.comm_unique_id()
        tmp = "%s.tmp%d" % (path, os.getpid())
        with open(tmp, "wb") as f:
            f.write(uid)
        os.replace(tmp, path)
        return uid, path
    t0 = time.time()
    while True:
        try:
            with open(path, "rb") as f:
                uid = f.read()
            if len(uid) == 128:
                return uid, path
        except FileNotFoundError:
            pass
        if time.time() - t0 > timeout:
            raise TimeoutError("rank %d: no NCCL id at %s after %.0f s" % (rank, path, timeout))
        time.sleep(0.01)


def init(ctx, rank=None, world=None):
    """Create the communicator of `ctx` (a _cabi.Context on this rank's GPU).  Collective: every rank calls it."""
    r, w, _ = env_rank()
    rank = r if rank is None else rank
    world = w if world is None else world
    if world == 1:
        ctx.comm_init(1, 0, None)
        return ctx
    uid, path = exchange_id(rank, world)
    ctx.comm_init(world, rank, uid)              # blocks until all ranks have joined
    ctx.comm_barrier()
    if rank == 0:
        try:
            os.remove(path)
        except OSError:
            pass
    return ctx


def grid_for(world):
    """Default process grid Pr x Pc for `world` ranks: the most square one with Pr <= Pc (8 -> 2 x 4)."""
    pr = 1
    for k in range(1, int(world ** 0.5) + 1):
        if world % k == 0:
            pr = k
    return pr, world // pr
