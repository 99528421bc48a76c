#!/usr/bin/env python
"""Regenerates tests/golden/ref_outputs.json by running the UNMODIFIED reference (oracle/_ref/libqpalm_ref.so, built from
/root/reference by oracle/Makefile) on the seeded problems of qpalm_b200.problems.  Run in the build container only
(`python tests/golden/make_golden.py`); the GPU box and the test-suite read the committed JSON.

Recorded per case: status_val, iter, iter_out, objective, pri/dua residual norms, solution x / y, the Ruiz scaling
vectors D / E and cost scale c after qpalm_setup (scaling.c:34-113) and the final gamma.
"""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qpalm_b200 import problems  # noqa: E402
from oracle.refbind import Qpalm  # noqa: E402

libc = ctypes.CDLL("libc.so.6")


def cases():
    """name -> QP.  Keep in sync with tests/test_oracle.py::golden_cases (it imports this function)."""
    c = {
        "basic_qp": problems.basic_qp(),
        "basic_qp_unscaled": problems.basic_qp(scaling=0),
        "basic_qp_noprox": problems.basic_qp(proximal=0, scaling=2),
        "basic_qp_dual_term": problems.basic_qp(enable_dual_termination=1),
        "medium_qp": problems.medium_qp(),
        "ls_qp": problems.ls_qp(),
        "degen_hess": problems.degen_hess_qp(),
        "prim_inf": problems.prim_inf_qp(),
        "dua_inf": problems.dua_inf_qp(),
        "nonconvex_qp": problems.nonconvex_qp(),
        "update_qp": problems.update_qp(),
        "random_60_120_s0": problems.random_qp(60, 120, 0.3, 0.1, seed=0),
        "random_60_120_s0_rankupd": problems.random_qp(60, 120, 0.3, 0.1, seed=0, max_rank_update_fraction=1.0),
        "random_300_600_s5": problems.random_qp(300, 600, 0.1, 0.02, seed=5),
        "dense_250_400_s2": problems.random_qp(250, 400, 1.0, 1.0, seed=2),
        "c1_random_1000_2000_s1": problems.random_qp(1000, 2000, 0.05, 0.007, seed=1),
        "nonconvex_random_100_200_s3": problems.nonconvex_random_qp(100, 200, seed=3),
        "mpc_small_k0": problems.mpc_batch(3, n=48, m0=80, seed=3).instance(0),
        "mpc_chain80w_k0": problems.mpc_batch(2, seed=1).instance(0),
    }
    return c


def run(impl, p):
    libc.srand(1)          # LOBPCG start vector comes from rand() (nonconvex.c:41-44)
    s = Qpalm(impl)
    for k, v in p.settings.items():
        setattr(s.settings, k, v)
    s.set_data(p.Q.copy(), p.A.copy(), p.q.copy(), p.bmin.copy(), p.bmax.copy())
    assert s._allocate_work()
    as_np = lambda ptr, k: np.ctypeslib.as_array(ptr, shape=(k,)).copy() if k and ptr else np.zeros(0)
    D, E, cs = np.zeros(0), np.zeros(0), 1.0
    if s.settings.scaling and s.work.scaling:       # the reference allocates work->scaling only when scaling > 0
        sc = s.work.scaling.contents
        D, E, cs = as_np(sc.D, p.n), as_np(sc.E, p.m), float(sc.c)
    s._solve()
    r = s.result()
    s.cleanup()
    return dict(status_val=r.status_val, iter=r.iter, iter_out=r.iter_out, objective=r.objective, dual_objective=r.dual_objective,
                pri_res_norm=r.pri_res_norm, dua_res_norm=r.dua_res_norm, gamma=r.gamma, x=r.x.tolist(), y=r.y.tolist(),
                D=D.tolist(), E=E.tolist(), c=cs)


if __name__ == "__main__":
    out = {name: run("reference", p) for name, p in cases().items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_outputs.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, {k: (v["status_val"], v["iter"], v["iter_out"]) for k, v in out.items()})
