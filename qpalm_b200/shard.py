"""Sharding of independent QPs over the GPUs of one box (BASELINE config 4).

The batched sweep partitions by instance: rank r of W owns the contiguous block ``shard_range(nb, r, W)``; the shared
(Q, A, settings) are replicated, there is NO data-path collective during the solve, and one gather of (x, y, info) at
the end returns the whole sweep on rank 0.  ``torch.distributed`` is plumbing only (NCCL on the GPU box, gloo in the
CPU tests); the solve itself is ``qpalm_b200.batch.Batch`` (CUDA) -- tests inject a different per-shard solver.
"""
from __future__ import annotations

import numpy as np


def shard_range(nb: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of instance indices owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(nb, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _default_solver(b, lo, hi):
    from . import batch as qb
    from .problems import BatchQP
    sub = BatchQP(b.Q, b.A, b.q[lo:hi], b.bmin[lo:hi], b.bmax[lo:hi], b.settings)
    return qb.solve_batch(sub)


def solve_batch_sharded(b, solver=None, device=None):
    """Solve every instance of the BatchQP `b` with the ranks of the default process group.

    Returns (x [nb, n], y [nb, m], infos [nb]) on rank 0 and (None, None, None) elsewhere.  `solver(b, lo, hi)` must
    return (x, y, infos) for instances lo..hi-1 (default: the CUDA batch entry point)."""
    import torch
    import torch.distributed as dist
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    nb, n, m = b.q.shape[0], b.q.shape[1], b.bmin.shape[1]
    lo, hi = shard_range(nb, rank, world)
    xs, ys, infos = (solver or _default_solver)(b, lo, hi) if hi > lo else (np.zeros((0, n)), np.zeros((0, m)), [])
    if world == 1:
        return xs, ys, infos
    # pack [x | y | status, iter, iter_out, pri, dua, obj] per instance and gather equal-sized padded blocks
    cap = -(-nb // world)
    blk = np.zeros((cap, n + m + 6))
    for k in range(hi - lo):
        i = infos[k]
        blk[k, :n], blk[k, n:n + m] = xs[k], ys[k]
        blk[k, n + m:] = [i["status_val"], i["iter"], i["iter_out"], i["pri_res_norm"], i["dua_res_norm"], i["objective"]]
    t = torch.from_numpy(blk)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, out, dst=0)
    if rank != 0:
        return None, None, None
    X, Y, I = np.zeros((nb, n)), np.zeros((nb, m)), []
    for r in range(world):
        l, h = shard_range(nb, r, world)
        a = out[r].cpu().numpy()[: h - l]
        X[l:h], Y[l:h] = a[:, :n], a[:, n:n + m]
        for row in a:
            s = row[n + m:]
            I.append(dict(status_val=int(s[0]), iter=int(s[1]), iter_out=int(s[2]), pri_res_norm=float(s[3]),
                          dua_res_norm=float(s[4]), objective=float(s[5])))
    return X, Y, I
