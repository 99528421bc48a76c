"""C3 (dense n=8000, m=16000) through the drop-in API with the per-iteration trace (QPALM_B200_TRACE): active-set changes per
iteration, the refactor-vs-update decision, device times.  usage: python tools/c3_trace.py [n m [seed]]"""
import os
import sys
import time

os.environ.setdefault("QPALM_B200_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpalm_b200 import problems
from qpalm_b200.interface import Qpalm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * n
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
p = problems.dense_qp(n, m, seed=seed)
for rep in range(2):
    s = Qpalm("b200")
    for k, v in p.settings.items():
        setattr(s.settings, k, v)
    s.set_data(p.Q, p.A, p.q, p.bmin, p.bmax)
    t0 = time.perf_counter()
    assert s._allocate_work()
    t1 = time.perf_counter()
    s._solve()
    t2 = time.perf_counter()
    r, st = s.result(), s.stats()
    print(f"rep {rep}: {r.status} iter {r.iter}/{r.iter_out} setup {t1 - t0:.3f} s solve {t2 - t1:.3f} s device {st.device_ms_total:.1f} ms | "
          f"refactorizations {st.refactorizations} ({st.device_ms_factor:.1f} ms) updown sweeps {st.updown_calls} ranks {st.updown_rank_sum} "
          f"({st.device_ms_updown:.1f} ms) sigma updates {st.sigma_update_calls} obj {r.objective:.10e}", flush=True)
    s.cleanup()
    os.environ.pop("QPALM_B200_TRACE", None)
