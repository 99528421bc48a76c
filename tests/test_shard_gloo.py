"""CPU suite: the N > 1 host logic (instance sharding + final gather) with world_size 2 over gloo.
The per-shard solver is injected (the oracle, test infrastructure) because there is no GPU here; on the GPU box the same
code path runs with the CUDA batch entry point."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from qpalm_b200 import problems  # noqa: E402
from qpalm_b200.shard import shard_range, solve_batch_sharded  # noqa: E402


def _oracle_solver(b, lo, hi):
    from oracle.refbind import solve_qp   # spawned workers do not run conftest: register the checker here
    xs, ys, infos = [], [], []
    for k in range(lo, hi):
        q = b.instance(k)
        r = solve_qp("oracle", q.Q, q.A, q.q, q.bmin, q.bmax, **q.settings)
        xs.append(r.x), ys.append(r.y)
        infos.append(dict(status_val=r.status_val, iter=r.iter, iter_out=r.iter_out, pri_res_norm=r.pri_res_norm,
                          dua_res_norm=r.dua_res_norm, objective=r.objective))
    return np.array(xs), np.array(ys), infos


def _worker(rank, world, port, nb, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = problems.mpc_batch(nb, n=24, m0=30, seed=4)
    X, Y, I = solve_batch_sharded(b, solver=_oracle_solver)
    if rank == 0:
        np.savez(out_path, X=X, Y=Y, it=np.array([i["iter"] for i in I]), st=np.array([i["status_val"] for i in I]))
    else:
        assert X is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for nb in (0, 1, 7, 8, 4096, 4099):
        for w in (1, 2, 3, 8):
            r = [shard_range(nb, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == nb
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [h - l for l, h in r]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_sweep_equals_single_rank(tmp_path):
    nb = 7          # ragged: 4 + 3
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, port, nb, out), nprocs=2, join=True)
    got = np.load(out)
    b = problems.mpc_batch(nb, n=24, m0=30, seed=4)
    X, Y, I = _oracle_solver(b, 0, nb)
    assert np.array_equal(got["X"], X) and np.array_equal(got["Y"], Y)
    assert list(got["it"]) == [i["iter"] for i in I] and list(got["st"]) == [i["status_val"] for i in I]
