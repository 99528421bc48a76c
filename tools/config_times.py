"""Time-to-solution of the BASELINE.json single-problem configs through the drop-in API, next to the reference on the
host: C1 (random sparse n=1000, m=2000, density 0.05), C2 stand-in (grid QP, SYNTHETIC: the Maros-Meszaros files are not
in the reference checkout) written to a QPS file and read back through the QPS front end, C5 (nonconvex random n=5000)."""
import json
import os
import sys
import tempfile
import time
import ctypes

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import oracle.refbind  # noqa: F401  (registers the "reference" / "oracle" checker libraries)
from qpalm_b200 import problems, qps
from qpalm_b200.interface import Qpalm, solve_qp
from qpalm_b200.problems import CSC

libc = ctypes.CDLL("libc.so.6")
which = sys.argv[1:] or ["c1", "c2", "c5"]


def timed(impl, p, reps):
    best, res, setup = None, None, None
    for _ in range(reps):
        libc.srand(1)
        s = Qpalm(impl)
        for k, v in p.settings.items():
            setattr(s.settings, k, v)
        s.set_data(p.Q.copy(), p.A.copy(), p.q.copy(), p.bmin.copy(), p.bmax.copy(), p.c)
        t0 = time.perf_counter(); assert s._allocate_work(); t1 = time.perf_counter()
        s._solve(); t2 = time.perf_counter()
        res = s.result(); s.cleanup()
        if best is None or t2 - t1 < best:
            best, setup = t2 - t1, t1 - t0
    return res, setup, best


def report(name, p, ref_reps=1):
    g, gs, gt = timed("b200", p, 3)
    r, rs, rt = timed("reference", p, ref_reps)
    rel = lambda a, b: float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b)))) if a.size else 0.0
    print(json.dumps(dict(config=name, n=p.n, m=p.m, status=g.status, ref_status=r.status, iter=[g.iter, g.iter_out], ref_iter=[r.iter, r.iter_out],
                          solve_s=gt, setup_s=gs, ref_solve_s=rt, ref_setup_s=rs, speedup_solve=rt / gt, speedup_setup_plus_solve=(rt + rs) / (gt + gs),
                          rel_dx=rel(g.x, r.x), rel_dy=rel(g.y, r.y), host_cores=os.cpu_count())), flush=True)


if "c1" in which:
    report("C1 random sparse convex QP n=1000 m=2000 density 0.05", problems.random_qp(1000, 2000, 0.05, 0.007, seed=0), ref_reps=2)
if "c2" in which:
    gq = problems.grid_qp(int(os.environ.get("C2_GRID", "150")), seed=0)
    n, m0 = gq.n, gq.m - gq.n
    Afull = sp.csc_matrix((gq.A.x, gq.A.i, gq.A.p), shape=(gq.m, n))
    path = os.path.join(tempfile.mkdtemp(), "grid.qps")
    qps.write_qps(path, "GRID", Afull[:m0], gq.bmin[:m0], gq.bmax[:m0], gq.q, sp.csc_matrix((gq.Q.x, gq.Q.i, gq.Q.p), shape=(n, n)),
                  var_lo=gq.bmin[m0:], var_up=gq.bmax[m0:])
    t0 = time.perf_counter(); pr = qps.read_qps(path); t_read = time.perf_counter() - t0
    p = problems.QP("grid_from_qps", CSC(n, n, pr.Q_p, pr.Q_i, pr.Q_x, -1), CSC(pr.m, n, pr.A_p, pr.A_i, pr.A_x, 0), pr.q, pr.bmin, pr.bmax, pr.c,
                    dict(eps_abs=1e-6, eps_rel=1e-6, verbose=0))
    print(f"# C2 stand-in: QPS file of {os.path.getsize(path) / 1e6:.1f} MB read in {t_read:.3f} s (qpalm_b200_qps_read)", flush=True)
    report(f"C2 stand-in (SYNTHETIC grid QP via QPS) n={n} m={pr.m}", p)
if "c5" in which:
    n5 = int(os.environ.get("C5_N", "5000"))
    seed5 = int(os.environ.get("C5_SEED", "1"))   # seed 0 and 2 make the REFERENCE return NaN at n=5000 (see profiles/r01d_config_times_*)
    report(f"C5 nonconvex random QP n={n5} m={2 * n5} seed={seed5}", problems.nonconvex_random_qp(n5, 2 * n5, seed=seed5))
