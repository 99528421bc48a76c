"""Row-sharded dense QP (BASELINE config 3) on 2 GPUs: NCCL allgather of A d, allreduce of A' yh and of
the Schur-complement SYRK partials.  Every rank must return the reference solution (status, x / y 1e-8, iterations 5 %).
Skipped on a single-GPU box (the driver's multi-GPU tier and `gpurun --gpus 2` run it)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, n, m, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)     # id exchange only; the data path is the library's NCCL
    from qpalm_b200 import problems, rowshard
    from qpalm_b200.interface import solve_qp
    rowshard.init(rank, world)
    for k, (nn, mm, seed) in enumerate([(n, m, 2), (n // 2 + 3, m + 1, 5)]):     # second case: ragged row blocks
        p = problems.dense_qp(nn, mm, seed=seed)
        r = solve_qp("b200", p.Q, p.A, p.q, p.bmin, p.bmax, **p.settings)
        np.savez(os.path.join(out_dir, f"r{rank}_{k}.npz"), x=r.x, y=r.y, it=r.iter, out=r.iter_out, st=r.status_val)
    dist.barrier()
    rowshard.finalize()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(600)
def test_row_sharded_dense_qp_two_gpus(tmp_path):
    import torch.multiprocessing as mp
    from qpalm_b200 import problems
    from qpalm_b200.interface import solve_qp
    n, m = 300, 640
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n, m, str(tmp_path)), nprocs=2, join=True)
    for k, (nn, mm, seed) in enumerate([(n, m, 2), (n // 2 + 3, m + 1, 5)]):
        p = problems.dense_qp(nn, mm, seed=seed)
        o = solve_qp("oracle", p.Q, p.A, p.q, p.bmin, p.bmax, **p.settings)
        got = [np.load(os.path.join(str(tmp_path), f"r{r}_{k}.npz")) for r in range(2)]
        assert np.array_equal(got[0]["x"], got[1]["x"]) and np.array_equal(got[0]["y"], got[1]["y"])   # replicated control flow
        for g in got:
            assert int(g["st"]) == o.status_val == 1
            assert np.max(np.abs(g["x"] - o.x)) / max(1.0, np.max(np.abs(o.x))) < 1e-8
            assert np.max(np.abs(g["y"] - o.y)) / max(1.0, np.max(np.abs(o.y))) < 1e-8
            assert abs(int(g["it"]) - o.iter) <= max(1, int(np.ceil(0.05 * o.iter)))
