"""Regenerates tests/golden/qps/gen_*.qps and tests/golden/qps_ref_outputs.json.

Run in the build container (needs /root/reference compiled into oracle/_ref by `make -C oracle ref`): every .qps file
under tests/golden/qps/ is parsed by the UNMODIFIED reference reader (interfaces/qps/src/qpalm_qps.c, driven through
oracle/qps_ref_shim.c) and the resulting QPALMData is stored as JSON.  tests/test_qps.py pins qpalm_b200_qps_read to it.
"""
import glob
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from qpalm_b200 import qps  # noqa: E402
from oracle import refbind  # noqa: E402

INF = 1e20


def gen(name, n, m0, seed, **kw):
    rng = np.random.default_rng(seed)
    A = sp.random(m0, n, density=0.4, random_state=np.random.RandomState(seed), format="csc")
    A.data = np.round(rng.standard_normal(A.nnz), 3)
    lo = np.round(-rng.random(m0), 3); up = np.round(rng.random(m0), 3)
    kind = rng.integers(0, 4, m0)
    lo[kind == 0] = -INF; up[kind == 1] = INF; up[kind == 2] = lo[kind == 2]
    M = sp.random(n, n, density=0.3, random_state=np.random.RandomState(seed + 1), format="csc")
    Q = sp.tril(M @ M.T + sp.eye(n), format="csc"); Q.data = np.round(Q.data, 4)
    q = np.round(rng.standard_normal(n), 3); q[rng.random(n) < 0.3] = 0.0
    vlo = np.zeros(n); vup = np.full(n, INF)
    vk = rng.integers(0, 5, n)
    vlo[vk == 1] = -INF; vup[vk == 1] = INF                                  # FR
    vlo[vk == 2] = np.round(-rng.random((vk == 2).sum()), 2)                 # LO
    vup[vk == 3] = np.round(1 + rng.random((vk == 3).sum()), 2)              # UP
    fx = vk == 4; vlo[fx] = vup[fx] = np.round(rng.random(fx.sum()), 2)      # FX
    qps.write_qps(os.path.join(HERE, "qps", f"{name}.qps"), name.upper(), A, lo, up, q, Q, c=float(np.round(rng.standard_normal(), 2)),
                  var_lo=vlo, var_up=vup, **kw)


def main():
    gen("gen_a", 8, 6, 1)
    gen("gen_b", 12, 9, 2, rhs_name=None, bnd_name=None, two_per_line=False)
    gen("gen_c", 10, 14, 3, two_per_line=True, rhs_name="B")
    out = {}
    for path in sorted(glob.glob(os.path.join(HERE, "qps", "*.qps"))):
        if os.path.basename(path).startswith("mine_"):
            continue                      # inputs the reference reader cannot handle (see tests/test_qps.py)
        p = refbind.read_qps_reference(path)
        out[os.path.basename(path)] = {k: (getattr(p, k).tolist() if hasattr(getattr(p, k), "tolist") else getattr(p, k))
                                       for k in ("n", "m", "A_p", "A_i", "A_x", "Q_p", "Q_i", "Q_x", "q", "c", "bmin", "bmax")}
    out["qps_settings.txt"] = refbind.read_settings_reference(os.path.join(HERE, "qps_settings.txt"))
    with open(os.path.join(HERE, "qps_ref_outputs.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out), "entries")


if __name__ == "__main__":
    main()
