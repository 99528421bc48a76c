"""Sequential warm-started MPC loop: seconds per step on the GPU next to the reference on the host (f4 measurement)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle.refbind  # noqa: F401  (registers the "reference" / "oracle" checker libraries)
from qpalm_b200 import mpc
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
b, seq = mpc.mpc_sequence(steps, seed=1)
out = {}
for impl in ("b200", "reference"):
    try:
        res, t = mpc.run_sequence(impl, b, seq)
    except Exception as e:   # reference build absent
        out[impl] = str(e); continue
    out[impl] = {"ms_per_step_median": 1e3 * float(np.median(t)), "ms_per_step_mean": 1e3 * float(np.mean(t)),
                 "mean_iter": float(np.mean([r.iter for r in res[1:]])), "all_solved": all(r.status_val == 1 for r in res)}
print(json.dumps({"workload": f"{steps} warm-started steps, n=240 m=949 (chain80w size), update_q + update_bounds + warm_start per step", **out}))
