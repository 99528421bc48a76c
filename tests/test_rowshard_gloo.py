"""CPU suite, world_size 2 over gloo: the exchange plan of the row-sharded dense QP (csrc/shard.cu).

Each rank keeps a contiguous block of constraint rows of A (rowshard.row_block = the partition engine_create applies) and
everything else replicated.  The three exchanges of an iteration are emulated with numpy on the rank's rows and
torch.distributed collectives, and must reproduce the unsharded quantities:
  A d        -> allgather of equal padded blocks,
  A' yh      -> allreduce(sum) of the n-vector partials,
  A_J' S A_J -> allreduce(sum) of the n x n SYRK partials.
The CUDA path itself is covered by tests/test_gpu_rowshard.py on two GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from qpalm_b200.rowshard import row_block  # noqa: E402


def test_row_block_partition():
    for m, world in ((16000, 8), (10, 4), (3, 8), (949, 2), (1, 2)):
        blocks = [row_block(m, r, world) for r in range(world)]
        cap = blocks[0][2]
        assert cap * world >= m and all(b[2] == cap for b in blocks)
        covered = []
        for lo, cnt, _ in blocks:
            covered += list(range(lo, lo + cnt))
        assert covered == list(range(m))


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)                       # same data on every rank (replicated inputs)
    n, m = 40, 101
    A = rng.standard_normal((m, n)); d = rng.standard_normal(n); yh = rng.standard_normal(m)
    sigma = 10.0 ** rng.uniform(-1, 1, m); active = rng.random(m) < 0.4
    lo, cnt, cap = row_block(m, rank, world)
    Ag = A[lo:lo + cnt]
    # A d: local rows into this rank's padded block, allgather in place
    blk = torch.zeros(cap, dtype=torch.float64); blk[:cnt] = torch.from_numpy(Ag @ d)
    gathered = [torch.zeros(cap, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, blk)
    Ad = torch.cat(gathered)[:m].numpy()
    # A' yh: local partial, allreduce
    part = torch.from_numpy(Ag.T @ yh[lo:lo + cnt]); dist.all_reduce(part)
    # SYRK partial over the active local rows, allreduce
    J = active[lo:lo + cnt]
    H = torch.from_numpy((Ag[J].T * sigma[lo:lo + cnt][J]) @ Ag[J]); dist.all_reduce(H)
    if rank == 0:
        np.savez(out_path, Ad=Ad, Aty=part.numpy(), H=H.numpy(), Ad_ref=A @ d, Aty_ref=A.T @ yh,
                 H_ref=(A[active].T * sigma[active]) @ A[active])
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_plan_two_ranks(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = np.load(out)
    assert np.max(np.abs(r["Ad"] - r["Ad_ref"])) < 1e-13 * np.max(np.abs(r["Ad_ref"]))   # numpy's BLAS blocks by row count
    assert np.max(np.abs(r["Aty"] - r["Aty_ref"])) < 1e-12 * np.max(np.abs(r["Aty_ref"]))
    assert np.max(np.abs(r["H"] - r["H_ref"])) < 1e-12 * np.max(np.abs(r["H_ref"]))
