import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from qpalm_b200 import abi  # noqa: E402
from oracle import refbind  # noqa: E402  (registers the "oracle" / "reference" checker libraries)
from qpalm_b200.interface import load_library  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAS_GPU = _has_gpu()
HAS_REF = refbind.have_reference()


class Ops:
    """Typed access to the operator-level ABI of one library (`qpalm_b200_X` or `oracle_X`)."""

    def __init__(self, impl):
        self.impl = impl
        self.lib = load_library(impl)
        self.pre = "oracle_" if impl == "oracle" else "qpalm_b200_"
        SP = C.POINTER(abi.SolverSparse)
        fp, ip = abi.c_float_p, abi.c_int_p
        sig = {
            "mat_vec": [SP, fp, fp], "mat_tpose_vec": [SP, fp, fp],
            "mat_inf_norm_cols": [SP, fp], "mat_inf_norm_rows": [SP, fp],
            "scale_data": [SP, SP, fp, fp, fp, abi.c_int, fp, fp, fp],
            "residuals_active_set": [SP, fp, fp, fp, fp, fp, fp, fp, fp, abi.c_int, abi.c_float, ip,
                                     fp, fp, fp, fp, fp, fp, fp, ip, ip, ip, ip, ip, ip],
            "linesearch": [abi.c_int, abi.c_float, abi.c_float, fp, fp, fp, fp, fp, fp, fp, fp, fp, ip, ip],
            "newton_solve": [SP, SP, fp, ip, abi.c_float, fp, fp, fp],
            "updown": [abi.c_int, abi.c_int, fp, fp, abi.c_int],
            "lobpcg": [SP, fp, fp, ip],
        }
        for name, args in sig.items():
            f = getattr(self.lib, self.pre + name)
            f.argtypes = args
            f.restype = C.c_int
            setattr(self, "_" + name, f)

    # ---- numpy-friendly wrappers ----
    def mat_vec(self, A, x):
        y = np.zeros(A.nrow)
        assert self._mat_vec(A.ref(), abi.fptr(np.ascontiguousarray(x, dtype=float)), abi.fptr(y)) == 0
        return y

    def mat_tpose_vec(self, A, x):
        y = np.zeros(A.ncol)
        assert self._mat_tpose_vec(A.ref(), abi.fptr(np.ascontiguousarray(x, dtype=float)), abi.fptr(y)) == 0
        return y

    def norm_cols(self, A):
        E = np.zeros(A.ncol)
        assert self._mat_inf_norm_cols(A.ref(), abi.fptr(E)) == 0
        return E

    def norm_rows(self, A):
        E = np.zeros(A.nrow)
        assert self._mat_inf_norm_rows(A.ref(), abi.fptr(E)) == 0
        return E

    def scale_data(self, A, Q, q, bmin, bmax, iters):
        A, Q = A.copy(), Q.copy()
        q, bmin, bmax = q.copy(), bmin.copy(), bmax.copy()
        D, E, c = np.zeros(Q.ncol), np.zeros(A.nrow), C.c_double(0)
        assert self._scale_data(A.ref(), Q.ref(), abi.fptr(q), abi.fptr(bmin), abi.fptr(bmax), iters,
                                abi.fptr(D), abi.fptr(E), C.cast(C.byref(c), abi.c_float_p)) == 0
        return dict(Ax=A.x, Qx=Q.x, q=q, bmin=bmin, bmax=bmax, D=D, E=E, c=c.value)

    def residuals(self, A, Ax, y, sigma, bmin, bmax, Qx, q, x0, proximal, gamma, active_old):
        m, n = A.nrow, A.ncol
        o = {k: np.zeros(m) for k in ("Axys", "z", "pri_res", "yh")}
        o.update({k: np.zeros(n) for k in ("Atyh", "df", "dphi")})
        act, ent, lea = (np.zeros(m + 1, dtype=np.int64) for _ in range(3))
        na, ne, nl = (np.zeros(1, dtype=np.int64) for _ in range(3))
        f = lambda a: abi.fptr(np.ascontiguousarray(a, dtype=float))
        keep = [np.ascontiguousarray(a, dtype=float) for a in (Ax, y, sigma, bmin, bmax, Qx, q, x0)]
        ao = np.ascontiguousarray(active_old, dtype=np.int64)
        rc = self._residuals_active_set(A.ref(), *[abi.fptr(k) for k in keep], int(proximal), float(gamma), abi.iptr(ao),
                                        f(o["Axys"]) if False else abi.fptr(o["Axys"]), abi.fptr(o["z"]), abi.fptr(o["pri_res"]),
                                        abi.fptr(o["yh"]), abi.fptr(o["Atyh"]), abi.fptr(o["df"]), abi.fptr(o["dphi"]),
                                        abi.iptr(act), abi.iptr(na), abi.iptr(ent), abi.iptr(ne), abi.iptr(lea), abi.iptr(nl))
        assert rc == 0
        o.update(active=act[:m], nb_active=int(na[0]), enter=ent[:int(ne[0])], leave=lea[:int(nl[0])])
        return o

    def linesearch(self, eta, beta, Ad, Ax, y, sigma, bmin, bmax):
        m = len(Ad)
        arrs = [np.ascontiguousarray(a, dtype=float) for a in (Ad, Ax, y, sigma, np.sqrt(sigma), bmin, bmax)]
        tau = np.zeros(1)
        ss, si, nL = np.zeros(2 * m + 1), np.zeros(2 * m + 1, dtype=np.int64), np.zeros(1, dtype=np.int64)
        rc = self._linesearch(m, float(eta), float(beta), *[abi.fptr(a) for a in arrs], abi.fptr(tau), abi.fptr(ss),
                              abi.iptr(si), abi.iptr(nL))
        assert rc == 0
        k = int(nL[0])
        return float(tau[0]), ss[:k].copy(), si[:k].copy()

    def newton_solve(self, Q, A, sigma, active, beta, rhs, want_L=True):
        n = Q.ncol
        d = np.zeros(n)
        L = np.zeros((n, n), order="F") if want_L else None
        sig = np.ascontiguousarray(sigma if sigma is not None else np.ones(max(A.nrow if A else 0, 1)), dtype=float)
        act = None if active is None else np.ascontiguousarray(active, dtype=np.int64)
        rhs = np.ascontiguousarray(rhs, dtype=float)
        rc = self._newton_solve(Q.ref(), A.ref() if A is not None else None, abi.fptr(sig),
                                None if act is None else abi.iptr(act), float(beta), abi.fptr(rhs), abi.fptr(d),
                                None if L is None else L.ctypes.data_as(abi.c_float_p))
        assert rc == 0, rc
        return d, L

    def updown(self, L, W, update):
        n, k = W.shape
        Lc = np.asfortranarray(L.copy())
        Wc = np.asfortranarray(W.copy())
        rc = self._updown(n, k, Lc.ctypes.data_as(abi.c_float_p), Wc.ctypes.data_as(abi.c_float_p), int(update))
        assert rc == 0, rc
        return Lc

    def lobpcg(self, Q, x0):
        lam, its = np.zeros(1), np.zeros(1, dtype=np.int64)
        assert self._lobpcg(Q.ref(), abi.fptr(np.ascontiguousarray(x0, dtype=float)), abi.fptr(lam), abi.iptr(its)) == 0
        return float(lam[0]), int(its[0])


@pytest.fixture(scope="session")
def oracle_ops():
    return Ops("oracle")


@pytest.fixture(scope="session")
def gpu_ops():
    if not HAS_GPU:
        pytest.skip("no CUDA device")
    return Ops("b200")
