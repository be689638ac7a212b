"""Sharding of independent robot instances over ranks: contiguous slices, no
data-path collective; the only exchange is one small stats reduction after the timed region."""
from __future__ import annotations

import numpy as np

STAT_SUM = ("instances", "iters_sum", "not_optimal", "bad_input")
STAT_MAX = ("iters_max", "mu_max", "rd_max")


def shard_slice(n_total: int, rank: int, world: int) -> slice:
    """Contiguous slice of ``n_total`` instances owned by ``rank`` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_total), int(world))
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def local_stats(status, iters, resid):
    """(sum-vector, max-vector) of one shard's solver outcome, float64, fixed order (STAT_SUM / STAT_MAX)."""
    status, iters, resid = np.asarray(status), np.asarray(iters), np.asarray(resid, dtype=np.float64).reshape(-1, 2)
    n = status.shape[0]
    s = np.array([n, iters.sum(), (status != 0).sum(), (status == 3).sum()], dtype=np.float64)
    m = np.array([iters.max() if n else 0, resid[:, 0].max() if n else 0.0, resid[:, 1].max() if n else 0.0],
                 dtype=np.float64)
    return s, m


def reduce_stats(ssum, smax, device=None):
    """All-reduce the two stats vectors over the default process group (NCCL on GPUs, gloo in tests).
    Returns a dict; with no initialised process group it just names the local values."""
    import torch
    import torch.distributed as dist
    ts = torch.as_tensor(ssum, dtype=torch.float64, device=device).clone()
    tm = torch.as_tensor(smax, dtype=torch.float64, device=device).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    out = dict(zip(STAT_SUM, ts.cpu().tolist()))
    out.update(zip(STAT_MAX, tm.cpu().tolist()))
    out["mean_iters"] = out["iters_sum"] / max(1.0, out["instances"])
    return out
