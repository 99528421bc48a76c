#!/usr/bin/env python
"""Reference records at the BASELINE.json sizes (C3 dense n=8000, m=16000; C5 nonconvex n=5000, m=10000).

Runs the UNMODIFIED reference (oracle/_ref/libqpalm_ref.so, built from /root/reference by oracle/Makefile) ONCE on the seeded
problem of qpalm_b200.problems and commits  tests/golden/<case>.json  (status, iterations, objective, residuals, setup / solve
seconds, host core count, BLAS threads, sha256 of the input bytes)  and  tests/golden/<case>.npz  (x, y).  Build container
only; the GPU box reads the committed files (tests/test_gpu_big.py, bench.py).

    OPENBLAS_NUM_THREADS=8 python tests/golden/make_golden_big.py c3 [threads-label]
    python tests/golden/make_golden_big.py c5
"""
import ctypes
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qpalm_b200 import problems  # noqa: E402

CASES = {
    "c3": ("c3_dense_n8000_m16000_s0", lambda: problems.dense_qp(8000, 16000, seed=0)),
    "c3_small": ("c3_dense_n2000_m4000_s0", lambda: problems.dense_qp(2000, 4000, seed=0)),
    "c5": ("c5_nonconvex_n5000_m10000_s1", lambda: problems.nonconvex_random_qp(5000, 10000, seed=1)),
}


def input_sha(p):
    h = hashlib.sha256()
    for a in (p.Q.p, p.Q.i, p.Q.x, p.A.p, p.A.i, p.A.x, p.q, p.bmin, p.bmax):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    key = sys.argv[1]
    name, make = CASES[key]
    suffix = ("_" + sys.argv[2]) if len(sys.argv) > 2 else ""
    from oracle.refbind import Qpalm   # the reference binding is test infrastructure, not product code
    t0 = time.time()
    p = make()
    sha = input_sha(p)
    print(f"{name}: generated in {time.time() - t0:.1f} s, sha256 {sha[:16]}", flush=True)
    ctypes.CDLL("libc.so.6").srand(1)
    s = Qpalm("reference")
    for k, v in p.settings.items():
        setattr(s.settings, k, v)
    s.set_data(p.Q, p.A, p.q, p.bmin, p.bmax)
    t0 = time.time()
    assert s._allocate_work()
    t_setup = time.time() - t0
    print(f"setup {t_setup:.1f} s", flush=True)
    t0 = time.time()
    s._solve()
    t_solve = time.time() - t0
    r = s.result()
    rec = dict(case=name, n=p.n, m=p.m, settings=p.settings, input_sha256=sha,
               status_val=r.status_val, status=r.status, iter=r.iter, iter_out=r.iter_out, objective=r.objective,
               pri_res_norm=r.pri_res_norm, dua_res_norm=r.dua_res_norm, gamma=r.gamma,
               setup_seconds_wall=t_setup, solve_seconds_wall=t_solve,
               info_setup_time=r.setup_time, info_solve_time=r.solve_time,
               host_cores=os.cpu_count(), blas_threads=int(os.environ.get("OPENBLAS_NUM_THREADS", os.cpu_count())),
               host="build container (Xeon, see profiles/README.md)", reference="oracle/_ref/libqpalm_ref.so (CHOLMOD build, OpenBLAS 0.3.15)",
               x_sha256=hashlib.sha256(r.x.tobytes()).hexdigest(), y_sha256=hashlib.sha256(r.y.tobytes()).hexdigest())
    s.cleanup()
    d = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(d, name + suffix + ".json"), "w") as f:
        json.dump(rec, f, indent=1)
    if not suffix:
        np.savez_compressed(os.path.join(d, name + ".npz"), x=r.x, y=r.y)
    print(json.dumps({k: v for k, v in rec.items() if k not in ("settings",)}))


if __name__ == "__main__":
    main()
