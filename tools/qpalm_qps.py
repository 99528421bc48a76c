"""Command-line twin of the reference's `qpalm_qps problem.qps [settings.txt]` (interfaces/qps/src/qpalm_qps.c:692-831),
running on the GPU through qpalm_b200_qps_solve."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpalm_b200 import qps  # noqa: E402

if len(sys.argv) not in (2, 3):
    sys.exit("Wrong number of arguments. Correct usage is qpalm_qps problem.qps or qpalm_qps problem.qps settings.txt.")
info, x, y = qps.solve_qps(sys.argv[1], sys.argv[2] if len(sys.argv) == 3 else None)
print(f"Iter: {info['iter']}")
print(f"Status: {info['status']}")
print(f"Objective: {info['objective']:.10e}")
print(f"Runtime: {info['setup_time'] + info['solve_time']:f} seconds")
