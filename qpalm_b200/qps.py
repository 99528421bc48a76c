"""QPS front end: host-side mirror of the reference's ``interfaces/qps`` (qpalm_qps.c).

``read_qps``       -> qpalm_b200_qps_read       (get_sizes_and_check_format + read_data, qpalm_qps.c:69-575)
``read_settings``  -> qpalm_b200_read_settings  (read_settings, qpalm_qps.c:610-689)
``solve_qps``      -> qpalm_b200_qps_solve      (main, qpalm_qps.c:692-831): read, upload, solve on the GPU
``write_qps`` writes a problem in the free QPS format the Maros-Meszaros set uses -- the reference has no writer; the
tests and ``bench.py`` use it to make synthetic stand-ins for the absent .qps files (SURVEY.md 8(c) "data not available").
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import abi
from .abi import QPALMData, QPALMInfo, QPALMSettings
from .interface import load_library


@dataclass
class QPSProblem:
    name: str
    n: int
    m: int
    A_p: np.ndarray
    A_i: np.ndarray
    A_x: np.ndarray
    Q_p: np.ndarray
    Q_i: np.ndarray
    Q_x: np.ndarray
    q: np.ndarray
    c: float
    bmin: np.ndarray
    bmax: np.ndarray


def _arr(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    ct = C.c_int64 if dtype == np.int64 else C.c_double
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()


def _unpack(d: QPALMData, name: str) -> QPSProblem:
    n, m = int(d.n), int(d.m)
    A, Q = d.A.contents, d.Q.contents
    Ap = _arr(A.p, n + 1, np.int64)
    Qp = _arr(Q.p, n + 1, np.int64)
    assert A.stype == 0 and Q.stype == -1 and A.nrow == m and A.ncol == n and Q.nrow == n and Q.ncol == n
    return QPSProblem(name, n, m, Ap, _arr(A.i, int(Ap[-1]), np.int64), _arr(A.x, int(Ap[-1]), np.float64),
                      Qp, _arr(Q.i, int(Qp[-1]), np.int64), _arr(Q.x, int(Qp[-1]), np.float64),
                      _arr(d.q, n, np.float64), float(d.c), _arr(d.bmin, m, np.float64), _arr(d.bmax, m, np.float64))


def _lib():
    lib = load_library("b200")
    if not getattr(lib, "_qps_typed", False):
        lib.qpalm_b200_qps_read.argtypes = [C.c_char_p, C.POINTER(C.POINTER(QPALMData)), C.c_char_p, C.c_size_t]
        lib.qpalm_b200_qps_read.restype = C.c_int
        lib.qpalm_b200_qps_free.argtypes = [C.POINTER(QPALMData)]
        lib.qpalm_b200_qps_free.restype = None
        lib.qpalm_b200_read_settings.argtypes = [C.c_char_p, C.POINTER(QPALMSettings)]
        lib.qpalm_b200_read_settings.restype = C.c_int
        lib.qpalm_b200_qps_solve.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(QPALMInfo), abi.c_float_p, abi.c_float_p,
                                             C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        lib.qpalm_b200_qps_solve.restype = C.c_int
        lib._qps_typed = True
    return lib


def read_qps(path: str) -> QPSProblem:
    lib = _lib()
    d = C.POINTER(QPALMData)()
    name = C.create_string_buffer(256)
    rc = lib.qpalm_b200_qps_read(os.fsencode(path), C.byref(d), name, len(name))
    if rc != 0:
        raise RuntimeError(f"qpalm_b200_qps_read({path}) failed with code {rc}")
    try:
        return _unpack(d.contents, name.value.decode())
    finally:
        lib.qpalm_b200_qps_free(d)


def read_settings(path: str) -> dict:
    lib = _lib()
    s = QPALMSettings()
    rc = lib.qpalm_b200_read_settings(os.fsencode(path), C.byref(s))
    out = {f: getattr(s, f) for f, _ in QPALMSettings._fields_}
    out["_rc"] = rc
    return out


def solve_qps(path: str, settings_path: str | None = None):
    """Returns (info dict, x, y) of the GPU solve of a QPS file."""
    lib = _lib()
    n, m = C.c_size_t(0), C.c_size_t(0)
    d = C.POINTER(QPALMData)()
    rc = lib.qpalm_b200_qps_read(os.fsencode(path), C.byref(d), None, 0)
    if rc != 0:
        raise RuntimeError(f"qpalm_b200_qps_read({path}) failed with code {rc}")
    nn, mm = int(d.contents.n), int(d.contents.m)
    lib.qpalm_b200_qps_free(d)
    x, y = np.zeros(nn), np.zeros(mm)
    n.value, m.value = nn, mm
    info = QPALMInfo()
    rc = lib.qpalm_b200_qps_solve(os.fsencode(path), os.fsencode(settings_path) if settings_path else None, C.byref(info),
                                  x.ctypes.data_as(abi.c_float_p), y.ctypes.data_as(abi.c_float_p), C.byref(n), C.byref(m))
    if rc != 0:
        raise RuntimeError(f"qpalm_b200_qps_solve({path}) failed with code {rc}")
    return ({"status": info.status.decode(), "status_val": int(info.status_val), "iter": int(info.iter),
             "iter_out": int(info.iter_out), "objective": float(info.objective), "solve_time": float(info.solve_time),
             "setup_time": float(info.setup_time)}, x, y)


# ---------------------------------------------------------------------------------------------
# writer (free QPS format)
# ---------------------------------------------------------------------------------------------
def _num(v: float) -> str:
    return repr(float(v))


def write_qps(path: str, name: str, A, row_lo, row_up, q, Q_lower=None, c: float = 0.0, var_lo=None, var_up=None,
              use_ranges: bool = True, rhs_name: str | None = "RHS", bnd_name: str | None = "BND",
              two_per_line: bool = True) -> None:
    """Write  min 1/2 x'Qx + q'x + c  s.t. row_lo <= A x <= row_up, var_lo <= x <= var_up  as a QPS file.

    A: scipy sparse (m0 x n) of the proper constraints (NOT the bound rows: the reader appends those).
    Row types: E if lo == up, L if lo = -inf, G if up = +inf, otherwise L with a RANGES entry (or G when use_ranges is False
    is not expressible: two-sided rows need RANGES).  Variable bounds: default [0, +inf) emits nothing; (-inf, +inf) -> FR;
    lo == up -> FX; otherwise LO / UP lines as needed (the reference ignores MI, so a (-inf, u] variable cannot be written).
    rhs_name / bnd_name = None writes the RHS / BOUNDS sections without set names (both forms occur in the test set).
    """
    import scipy.sparse as sp
    A = sp.csc_matrix(A)
    m0, n = A.shape
    INF = abi.QPALM_INFTY
    row_lo = np.asarray(row_lo, dtype=float); row_up = np.asarray(row_up, dtype=float)
    var_lo = np.zeros(n) if var_lo is None else np.asarray(var_lo, dtype=float)
    var_up = np.full(n, INF) if var_up is None else np.asarray(var_up, dtype=float)
    rn = [f"R{r + 1}" for r in range(m0)]
    cn = [f"C{j + 1}" for j in range(n)]
    types, rhs, ranges = [], {}, {}
    for r in range(m0):
        lo, up = row_lo[r], row_up[r]
        if lo == up:
            types.append("E"); rhs[r] = lo
        elif lo <= -INF:
            types.append("L"); rhs[r] = up
        elif up >= INF:
            types.append("G"); rhs[r] = lo
        else:
            if not use_ranges:
                raise ValueError("two-sided row needs RANGES")
            types.append("L"); rhs[r] = up; ranges[r] = up - lo
    with open(path, "w") as f:
        f.write(f"NAME          {name}\nROWS\n N  OBJ\n")
        for r in range(m0):
            f.write(f" {types[r]}  {rn[r]}\n")
        f.write("COLUMNS\n")
        for j in range(n):
            ent = []
            if q[j] != 0.0:
                ent.append(("OBJ", q[j]))
            for k in range(A.indptr[j], A.indptr[j + 1]):
                ent.append((rn[A.indices[k]], A.data[k]))
            if not ent:
                ent.append(("OBJ", 0.0))
            step = 2 if two_per_line else 1
            for k in range(0, len(ent), step):
                f.write(f"    {cn[j]}  " + "  ".join(f"{nm}  {_num(v)}" for nm, v in ent[k:k + step]) + "\n")
        pre = f"    {rhs_name}  " if rhs_name else "    "
        f.write("RHS\n")
        if c != 0.0:
            f.write(f"{pre}OBJ  {_num(-c)}\n")
        for r in range(m0):
            if rhs[r] != 0.0:
                f.write(f"{pre}{rn[r]}  {_num(rhs[r])}\n")
        if ranges:
            f.write("RANGES\n")
            for r, v in ranges.items():
                f.write(f"    RNG  {rn[r]}  {_num(v)}\n")
        bl = []
        bpre = f"{bnd_name}  " if bnd_name else ""
        for j in range(n):
            lo, up = var_lo[j], var_up[j]
            if lo <= -INF and up >= INF:
                bl.append(f" FR {bpre}{cn[j]}\n")
            elif lo == up:
                bl.append(f" FX {bpre}{cn[j]}  {_num(lo)}\n")
            else:
                if lo <= -INF:
                    raise ValueError("(-inf, u] variables are not representable (the reference ignores MI)")
                if lo != 0.0:
                    bl.append(f" LO {bpre}{cn[j]}  {_num(lo)}\n")
                if up < INF:
                    bl.append(f" UP {bpre}{cn[j]}  {_num(up)}\n")
        if bl:
            f.write("BOUNDS\n" + "".join(bl))
        if Q_lower is not None:
            Ql = sp.csc_matrix(Q_lower)
            if Ql.nnz:
                f.write("QUADOBJ\n")
                for j in range(n):
                    for k in range(Ql.indptr[j], Ql.indptr[j + 1]):
                        f.write(f"    {cn[j]}  {cn[Ql.indices[k]]}  {_num(Ql.data[k])}\n")
        f.write("ENDATA\n")
