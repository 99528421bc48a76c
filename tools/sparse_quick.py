import os, sys, time
sys.path.insert(0, os.getcwd())
from qpalm_b200 import problems
from qpalm_b200.interface import Qpalm
p = problems.grid_qp(300, seed=0)
for rep in range(2):
    s = Qpalm("b200")
    for k, v in p.settings.items(): setattr(s.settings, k, v)
    s.set_data(p.Q, p.A, p.q, p.bmin, p.bmax)
    t0 = time.perf_counter(); s._allocate_work(); t1 = time.perf_counter()
    if rep == 0: s._solve()
    t2 = time.perf_counter()
    r = s.result(); s.cleanup()
    print(f"grid g=300 rep {rep}: setup {t1-t0:.3f} s solve {t2-t1:.3f} s status {r.status} iter {r.iter}/{r.iter_out}", flush=True)
