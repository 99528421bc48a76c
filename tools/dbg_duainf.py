import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.refbind  # noqa: F401  (registers the "reference" / "oracle" checker libraries)
from qpalm_b200 import problems
from qpalm_b200.interface import solve_qp
p = problems.dua_inf_qp()
st = dict(p.settings); st["verbose"] = 1
for impl in ("b200", "reference"):
    print("=====", impl, flush=True)
    r = solve_qp(impl, p.Q, p.A, p.q, p.bmin, p.bmax, **st)
    sys.stdout.flush()
    print(impl, r.status, r.iter, r.iter_out, flush=True)
