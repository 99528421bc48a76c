"""Warm qpalm_setup / qpalm_cleanup wall times of the dense config (host buffers -> device, Ruiz scaling)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpalm_b200 import problems
from qpalm_b200.interface import Qpalm
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
p = problems.dense_qp(n, 2 * n, seed=0)
for rep in range(4):
    s = Qpalm("b200")
    for k, v in p.settings.items():
        setattr(s.settings, k, v)
    s.set_data(p.Q, p.A, p.q, p.bmin, p.bmax)
    t0 = time.perf_counter(); s._allocate_work(); t1 = time.perf_counter()
    if rep >= 2: s._solve()
    t2 = time.perf_counter(); s.cleanup(); t3 = time.perf_counter()
    print(f"rep {rep}: setup {t1 - t0:.3f} s, solve {t2 - t1:.3f} s, cleanup {t3 - t2:.3f} s", flush=True)
