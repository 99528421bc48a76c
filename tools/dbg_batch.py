import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from collections import Counter
from qpalm_b200 import batch as qb, problems
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 148
b = problems.mpc_batch(nb, seed=0)
h = qb.Batch(b.Q, b.A, b.settings, nb)
h.upload(b.q, b.bmin, b.bmax)
for rep in range(3):
    ms = h.solve_resident(nb)
    x, y, infos = h.download(nb)
    it = np.array([i["iter"] for i in infos])
    print(rep, f"{ms:.2f} ms", "iters mean", it.mean(), "max", it.max(), Counter(i["status_val"] for i in infos), "x0 norm", np.abs(x[0]).max())
