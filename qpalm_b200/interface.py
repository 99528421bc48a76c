"""Host-side mirror of the reference's Python interface (interfaces/python/qpalm.py:190-330).

``Qpalm`` keeps the reference class's method names (``set_data``, ``_solve``, ``_warm_start``,
``_update_bounds``, ``_update_q``, ``_update_settings``) and argument meaning, but talks to the
drop-in C ABI of this repository directly (no ``python_allocate_*`` helper library is needed because
the structs are built with ctypes).  ``impl="b200"`` binds qpalm_b200/libqpalm_b200.so -- the product.
Loading fails loudly if the CUDA library has not been built; there is no CPU fallback.

The class is generic over any library that exports the reference API, so the test infrastructure
(oracle/refbind.py, which this package never imports) can register the compiled reference and the CPU
oracle as extra ``impl`` names through ``register_checker`` and drive them with the same code the tests
use on the product.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import abi
from .abi import (CSC, QPALMData, QPALMInfo, QPALMSettings, QPALMWorkspace, c_float_p, fptr)

_LIBS: dict[str, C.CDLL] = {}


_CHECKERS: dict[str, tuple] = {}


def register_checker(impl: str, path: str, prefix: str, loader) -> None:
    """Called by oracle/refbind.py (test infrastructure) only: makes ``Qpalm(impl)`` bind another library that exports the
    reference API (symbol prefix ``prefix``).  The product never registers anything."""
    if impl == "b200":
        raise ValueError("the product library cannot be re-bound")
    _CHECKERS[impl] = (path, prefix, loader)


def load_library(impl: str = "b200") -> C.CDLL:
    """Load (once) and type the product library (or a checker library registered by the test infrastructure)."""
    if impl in _LIBS:
        return _LIBS[impl]
    if impl == "b200":
        path, pre, loader = abi.PRODUCT_LIB, "", None
    elif impl in _CHECKERS:
        path, pre, loader = _CHECKERS[impl]
    else:
        raise RuntimeError(f"unknown implementation {impl!r}: the product binds \"b200\" only (checker libraries are "
                           "registered by the test infrastructure, oracle/refbind.py)")
    if not os.path.exists(path):
        raise RuntimeError(
            f"{impl}: shared library {path} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'`"
            + (" (the product has no CPU fallback)" if impl == "b200" else ""))
    lib = loader() if loader else C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    W = C.POINTER(QPALMWorkspace)

    def sym(name):
        return getattr(lib, pre + name)

    sym("qpalm_set_default_settings").argtypes = [C.POINTER(QPALMSettings)]
    sym("qpalm_set_default_settings").restype = None
    sym("qpalm_setup").argtypes = [C.POINTER(QPALMData), C.POINTER(QPALMSettings)]
    sym("qpalm_setup").restype = W
    sym("qpalm_warm_start").argtypes = [W, c_float_p, c_float_p]
    sym("qpalm_warm_start").restype = None
    sym("qpalm_solve").argtypes = [W]
    sym("qpalm_solve").restype = None
    sym("qpalm_update_settings").argtypes = [W, C.POINTER(QPALMSettings)]
    sym("qpalm_update_settings").restype = None
    sym("qpalm_update_bounds").argtypes = [W, c_float_p, c_float_p]
    sym("qpalm_update_bounds").restype = None
    sym("qpalm_update_q").argtypes = [W, c_float_p]
    sym("qpalm_update_q").restype = None
    sym("qpalm_cleanup").argtypes = [W]
    sym("qpalm_cleanup").restype = None
    lib._prefix = pre
    lib._impl = impl
    _LIBS[impl] = lib
    return lib


@dataclass
class Result:
    x: np.ndarray
    y: np.ndarray
    status_val: int
    status: str
    iter: int
    iter_out: int
    pri_res_norm: float
    dua_res_norm: float
    objective: float
    dual_objective: float
    setup_time: float
    solve_time: float
    gamma: float
    delta_x: np.ndarray
    delta_y: np.ndarray


class Qpalm:
    """Reference-compatible wrapper (see module docstring)."""

    def __init__(self, impl: str = "b200"):
        self.impl = impl
        self.lib = load_library(impl)
        self._pre = self.lib._prefix
        self._work = None
        self._data = None
        self._keep = {}
        self._settings = QPALMSettings()
        self._f("qpalm_set_default_settings")(C.byref(self._settings))

    def _f(self, name):
        return getattr(self.lib, self._pre + name)

    def __del__(self):
        try:
            self.cleanup()
        except Exception:
            pass

    def cleanup(self):
        if self._work:
            self._f("qpalm_cleanup")(self._work)
        self._work = None

    # ---- reference-named methods -------------------------------------------------------------
    def set_data(self, Q, A, q, bmin, bmax, c=0.0):
        """Q: symmetric n x n (scipy sparse, dense ndarray or abi.CSC with stype -1);
        A: m x n (scipy sparse, dense ndarray or abi.CSC).  As in the reference only the lower
        triangle of Q is read."""
        Qc = Q if isinstance(Q, CSC) else (CSC.from_dense(Q, -1) if isinstance(Q, np.ndarray) else CSC.from_scipy(Q, -1))
        Ac = A if isinstance(A, CSC) else (CSC.from_dense(A, 0) if isinstance(A, np.ndarray) else CSC.from_scipy(A, 0))
        n, m = Qc.ncol, Ac.nrow
        q = np.ascontiguousarray(q, dtype=np.float64)
        bmin = np.ascontiguousarray(bmin, dtype=np.float64)
        bmax = np.ascontiguousarray(bmax, dtype=np.float64)
        assert q.size == n and bmin.size == m and bmax.size == m and (m == 0 or Ac.ncol == n)
        if m == 0:  # keep valid pointers for zero-length arrays
            bmin = np.zeros(1)[:0].copy()
            bmax = np.zeros(1)[:0].copy()
        self._keep = dict(Q=Qc, A=Ac, q=q, bmin=bmin, bmax=bmax)
        d = QPALMData()
        d.n, d.m = n, m
        d.Q, d.A = Qc.ptr(), Ac.ptr()
        d.q = fptr(q)
        d.c = float(c)
        d.bmin = bmin.ctypes.data_as(c_float_p)
        d.bmax = bmax.ctypes.data_as(c_float_p)
        self._data = d
        self.n, self.m = n, m

    @property
    def settings(self) -> QPALMSettings:
        return self._settings

    def _allocate_work(self):
        self._work = self._f("qpalm_setup")(C.byref(self._data), C.byref(self._settings))
        if not self._work:
            self._work = None
        return self._work is not None

    def _solve(self):
        if self._data is None:
            raise RuntimeError("No data given")
        if self._work is None and not self._allocate_work():
            raise RuntimeError("qpalm_setup returned NULL")
        self._f("qpalm_solve")(self._work)

    def _warm_start(self, x=None, y=None):
        xs = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
        ys = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        self._f("qpalm_warm_start")(self._work,
                                    None if xs is None else fptr(xs),
                                    None if ys is None else fptr(ys))

    def _update_bounds(self, bmin=None, bmax=None):
        if bmin is not None:
            self._keep["bmin"][:] = bmin
        if bmax is not None:
            self._keep["bmax"][:] = bmax
        self._f("qpalm_update_bounds")(self._work,
                                       None if bmin is None else fptr(self._keep["bmin"]),
                                       None if bmax is None else fptr(self._keep["bmax"]))

    def _update_q(self, q=None):
        if q is not None:
            self._keep["q"][:] = q
        self._f("qpalm_update_q")(self._work, fptr(self._keep["q"]))

    def _update_settings(self):
        self._f("qpalm_update_settings")(self._work, C.byref(self._settings))

    # ---- result access -----------------------------------------------------------------------
    @property
    def work(self) -> QPALMWorkspace:
        return self._work.contents

    @property
    def info(self) -> QPALMInfo:
        return self._work.contents.info.contents

    def stats(self) -> abi.QPALMB200Stats:
        """Counters of the CUDA engine (qpalm_b200_get_stats, an additive entry point: "b200" only)."""
        f = self.lib.qpalm_b200_get_stats
        f.argtypes = [C.POINTER(QPALMWorkspace), C.POINTER(abi.QPALMB200Stats)]
        f.restype = C.c_int
        st = abi.QPALMB200Stats()
        if f(self._work, C.byref(st)) != 0:
            raise RuntimeError("qpalm_b200_get_stats failed")
        return st

    def vec(self, name: str, length: int) -> np.ndarray:
        p = getattr(self._work.contents, name)
        return np.ctypeslib.as_array(p, shape=(length,)).copy() if length else np.zeros(0)

    def result(self) -> Result:
        w = self._work.contents
        info = w.info.contents
        n, m = self.n, self.m
        as_np = lambda p, k: (np.ctypeslib.as_array(p, shape=(k,)).copy() if k else np.zeros(0))
        return Result(x=as_np(w.solution.contents.x, n), y=as_np(w.solution.contents.y, m),
                      status_val=int(info.status_val), status=info.status.decode(),
                      iter=int(info.iter), iter_out=int(info.iter_out),
                      pri_res_norm=float(info.pri_res_norm), dua_res_norm=float(info.dua_res_norm),
                      objective=float(info.objective), dual_objective=float(info.dual_objective),
                      setup_time=float(info.setup_time), solve_time=float(info.solve_time),
                      gamma=float(w.gamma), delta_x=as_np(w.delta_x, n), delta_y=as_np(w.delta_y, m))


def solve_qp(impl, Q, A, q, bmin, bmax, c=0.0, warm_x=None, warm_y=None, **settings) -> Result:
    """One-shot convenience used all over the tests and by bench.py."""
    s = Qpalm(impl)
    for k, v in settings.items():
        setattr(s.settings, k, v)
    s.set_data(Q, A, q, bmin, bmax, c)
    if not s._allocate_work():
        raise RuntimeError("qpalm_setup returned NULL")
    if warm_x is not None or warm_y is not None:
        s._warm_start(warm_x, warm_y)
    s._solve()
    r = s.result()
    s.cleanup()
    return r
