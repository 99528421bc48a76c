"""Small workloads for compute-sanitizer (tools/gpu_sanitize.sh): which = batch | dense | sparse | front"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpalm_b200 import problems
from qpalm_b200.interface import Qpalm
which = sys.argv[1]
if which == "batch":
    from qpalm_b200 import batch as qb
    b = problems.mpc_batch(6, n=48, m0=80, seed=3)
    x, y, infos = qb.solve_batch(b)
    print("batch:", [(i["status_val"], i["iter"]) for i in infos])
elif which == "front":   # one front above the shared-memory size: the cluster-per-front kernel (mfc::k_mf_front), 16 CTAs
    import numpy as np
    from qpalm_b200.sparse import sparse_newton
    p = problems.random_qp(300, 500, 0.2, 0.1, seed=4)
    rng = np.random.default_rng(1)
    sigma = 0.5 + 20 * rng.random(p.m); act = (rng.random(p.m) < 0.5).astype(np.int64); rhs = rng.standard_normal(p.n)
    d, L, perm, _ = sparse_newton(p.Q, p.A, sigma, act, 1e-3, rhs, want_factor=True)
    print("front: |d|", float(np.max(np.abs(d))), "L finite", bool(np.all(np.isfinite(L))))
else:
    p = problems.random_qp(300, 600, 1.0, 1.0, seed=3) if which == "dense" else problems.grid_qp(24, seed=2)
    s = Qpalm("b200")
    for k, v in p.settings.items():
        setattr(s.settings, k, v)
    s.set_data(p.Q, p.A, p.q, p.bmin, p.bmax)
    assert s._allocate_work()
    s._solve()
    r, st = s.result(), s.stats()
    print(which, r.status, r.iter, r.iter_out, "refactor", st.refactorizations, "updown sweeps", st.updown_calls, "launches", st.kernel_launches)
    s.cleanup()
