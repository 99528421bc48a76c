#!/usr/bin/env python
"""Device time of the resident batch solve (no phase clocks): python tools/batch_time.py [nb] [reps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qpalm_b200 import batch as qb, problems

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
b = problems.mpc_batch(nb, seed=0)
h = qb.Batch(b.Q, b.A, b.settings, nb)
h.upload(b.q, b.bmin, b.bmax)
ms = [h.solve_resident(nb) for _ in range(reps)]
x, y, infos = h.download(nb)
iters = np.array([i["iter"] for i in infos])
st = h.stats(nb)
print(f"nb={nb} ms min {min(ms[1:]):.2f} mean {np.mean(ms[1:]):.2f} -> {nb / min(ms[1:]) * 1e3:.0f} solves/s; iters mean {iters.mean():.2f} max {iters.max()}; solved {sum(i['status_val'] == 1 for i in infos)}; "
      f"refac {st['refactorizations'] / nb:.1f} sweeps {st['updown_sweeps'] / nb:.1f}")
