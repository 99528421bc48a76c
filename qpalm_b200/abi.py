"""ctypes mirror of include/qpalm_b200.h (Part 1) -- the drop-in ABI.

The struct layouts are the reference's, so the same classes describe ``qpalm_b200/libqpalm_b200.so`` (the product,
CUDA, sm_100a) and any other library exporting the reference API (the test infrastructure binds two, oracle/refbind.py).

Reference for the layouts: /root/reference/include/types.h:37-314 and the field-by-field ctypes
mirror in /root/reference/interfaces/python/qpalm.py:15-190.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

c_int = C.c_int64
c_float = C.c_double
c_int_p = C.POINTER(c_int)
c_float_p = C.POINTER(c_float)

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_LIB = os.environ.get("QPALM_B200_LIB") or os.path.join(REPO_ROOT, "qpalm_b200", "libqpalm_b200.so")   # override: A/B builds

QPALM_SOLVED = 1
QPALM_DUAL_TERMINATED = 2
QPALM_MAX_ITER_REACHED = -2
QPALM_PRIMAL_INFEASIBLE = -3
QPALM_DUAL_INFEASIBLE = -4
QPALM_TIME_LIMIT_REACHED = -5
QPALM_UNSOLVED = -10
QPALM_ERROR = 0
QPALM_INFTY = 1e20


class SolverSparse(C.Structure):
    """cholmod_sparse layout (cholmod_core.h:1214-1263)."""
    _fields_ = [("nrow", C.c_size_t), ("ncol", C.c_size_t), ("nzmax", C.c_size_t),
                ("p", C.c_void_p), ("i", C.c_void_p), ("nz", C.c_void_p),
                ("x", C.c_void_p), ("z", C.c_void_p),
                ("stype", C.c_int), ("itype", C.c_int), ("xtype", C.c_int),
                ("dtype", C.c_int), ("sorted", C.c_int), ("packed", C.c_int)]


class QPALMSettings(C.Structure):
    _fields_ = [("max_iter", c_int), ("inner_max_iter", c_int),
                ("eps_abs", c_float), ("eps_rel", c_float), ("eps_abs_in", c_float), ("eps_rel_in", c_float),
                ("rho", c_float), ("eps_prim_inf", c_float), ("eps_dual_inf", c_float),
                ("theta", c_float), ("delta", c_float), ("sigma_max", c_float), ("sigma_init", c_float),
                ("proximal", c_int), ("gamma_init", c_float), ("gamma_upd", c_float), ("gamma_max", c_float),
                ("scaling", c_int), ("nonconvex", c_int), ("verbose", c_int), ("print_iter", c_int),
                ("warm_start", c_int), ("reset_newton_iter", c_int), ("enable_dual_termination", c_int),
                ("dual_objective_limit", c_float), ("time_limit", c_float),
                ("ordering", c_int), ("factorization_method", c_int), ("max_rank_update", c_int),
                ("max_rank_update_fraction", c_float)]


class QPALMData(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t),
                ("Q", C.POINTER(SolverSparse)), ("A", C.POINTER(SolverSparse)),
                ("q", c_float_p), ("c", c_float), ("bmin", c_float_p), ("bmax", c_float_p)]


class QPALMInfo(C.Structure):
    _fields_ = [("iter", c_int), ("iter_out", c_int), ("status", C.c_char * 32), ("status_val", c_int),
                ("pri_res_norm", c_float), ("dua_res_norm", c_float), ("dua2_res_norm", c_float),
                ("objective", c_float), ("dual_objective", c_float),
                ("setup_time", c_float), ("solve_time", c_float), ("run_time", c_float)]


class QPALMSolution(C.Structure):
    _fields_ = [("x", c_float_p), ("y", c_float_p)]


class QPALMScaling(C.Structure):
    _fields_ = [("D", c_float_p), ("Dinv", c_float_p), ("E", c_float_p), ("Einv", c_float_p),
                ("c", c_float), ("cinv", c_float)]


class ArrayElement(C.Structure):
    _fields_ = [("x", c_float), ("i", C.c_size_t)]


class QPALMSolver(C.Structure):
    _fields_ = [("factorization_method", c_int),
                ("kkt", C.c_void_p), ("kkt_full", C.c_void_p), ("At", C.c_void_p),
                ("first_row_A", c_int_p), ("first_elem_A", c_float_p),
                ("LD", C.c_void_p), ("sym", C.c_void_p), ("LD_Q", C.c_void_p), ("sym_Q", C.c_void_p),
                ("E_temp", C.c_void_p), ("D_temp", C.c_void_p), ("neg_dphi", C.c_void_p),
                ("rhs_kkt", C.c_void_p), ("sol_kkt", C.c_void_p), ("d", C.c_void_p),
                ("Ad", C.c_void_p), ("Qd", C.c_void_p), ("yh", C.c_void_p), ("Atyh", C.c_void_p),
                ("first_factorization", c_int), ("reset_newton", c_int),
                ("active_constraints", c_int_p), ("active_constraints_old", c_int_p),
                ("nb_active_constraints", c_int),
                ("enter", c_int_p), ("nb_enter", c_int), ("leave", c_int_p), ("nb_leave", c_int),
                ("At_scale", C.c_void_p), ("At_sqrt_sigma", C.c_void_p)]


class QPALMWorkspace(C.Structure):
    _fields_ = [("data", C.POINTER(QPALMData)),
                ("x", c_float_p), ("y", c_float_p), ("Ax", c_float_p), ("Qx", c_float_p),
                ("Aty", c_float_p), ("x_prev", c_float_p), ("initialized", c_int),
                ("temp_m", c_float_p), ("temp_n", c_float_p), ("sigma", c_float_p), ("sigma_inv", c_float_p),
                ("sqrt_sigma_max", c_float), ("nb_sigma_changed", c_int), ("gamma", c_float),
                ("gamma_maxed", c_int),
                ("Axys", c_float_p), ("z", c_float_p), ("pri_res", c_float_p), ("pri_res_in", c_float_p),
                ("yh", c_float_p), ("Atyh", c_float_p), ("df", c_float_p), ("x0", c_float_p),
                ("xx0", c_float_p), ("dphi", c_float_p), ("neg_dphi", c_float_p), ("dphi_prev", c_float_p),
                ("d", c_float_p),
                ("tau", c_float), ("Qd", c_float_p), ("Ad", c_float_p), ("sqrt_sigma", c_float_p),
                ("sqrt_delta", c_float), ("eta", c_float), ("beta", c_float),
                ("delta", c_float_p), ("alpha", c_float_p), ("temp_2m", c_float_p), ("delta2", c_float_p),
                ("delta_alpha", c_float_p), ("s", C.POINTER(ArrayElement)),
                ("index_L", c_int_p), ("index_P", c_int_p), ("index_J", c_int_p),
                ("eps_pri", c_float), ("eps_dua", c_float), ("eps_dua_in", c_float),
                ("eps_abs_in", c_float), ("eps_rel_in", c_float),
                ("delta_y", c_float_p), ("Atdelta_y", c_float_p),
                ("delta_x", c_float_p), ("Qdelta_x", c_float_p), ("Adelta_x", c_float_p),
                ("D_temp", c_float_p), ("E_temp", c_float_p),
                ("solver", C.POINTER(QPALMSolver)), ("settings", C.POINTER(QPALMSettings)),
                ("scaling", C.POINTER(QPALMScaling)), ("solution", C.POINTER(QPALMSolution)),
                ("info", C.POINTER(QPALMInfo)), ("timer", C.c_void_p)]


class QPALMB200Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("inner_iterations", C.c_int64),
                ("outer_iterations", C.c_int64), ("refactorizations", C.c_int64),
                ("refactor_active_sum", C.c_int64), ("updown_calls", C.c_int64),
                ("updown_rank_sum", C.c_int64), ("spmv_calls", C.c_int64),
                ("algorithmic_bytes", C.c_double), ("dense_flops", C.c_double),
                ("device_ms_factor", C.c_double), ("device_ms_updown", C.c_double),
                ("device_ms_total", C.c_double),
                ("sparse_factor_nnz", C.c_int64), ("sparse_supernodes", C.c_int64), ("sparse_levels", C.c_int64),
                ("sigma_update_calls", C.c_int64), ("sigma_update_rank_sum", C.c_int64),
                ("kkt_factorizations", C.c_int64), ("kkt_refinement_steps", C.c_int64)]


# --------------------------------------------------------------------------------------------
# helpers: numpy <-> ABI
# --------------------------------------------------------------------------------------------
def fptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_float_p)


def iptr(a: np.ndarray):
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int_p)


class CSC:
    """Owns the numpy arrays behind a SolverSparse (CHOLMOD compressed-column, int64 indices)."""

    def __init__(self, nrow, ncol, p, i, x, stype=0):
        self.p = np.ascontiguousarray(p, dtype=np.int64)
        self.i = np.ascontiguousarray(i, dtype=np.int64)
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        self.nrow, self.ncol, self.stype = int(nrow), int(ncol), int(stype)
        nz = int(self.p[-1])
        # nzmax must be >= 1 for CHOLMOD's copy_sparse
        if self.i.size == 0:
            self.i = np.zeros(1, dtype=np.int64)
            self.x = np.zeros(1, dtype=np.float64)
        self.struct = SolverSparse(self.nrow, self.ncol, max(nz, 1),
                                   self.p.ctypes.data, self.i.ctypes.data, None,
                                   self.x.ctypes.data, None,
                                   self.stype, 2, 1, 0, 1, 1)

    @classmethod
    def from_scipy(cls, M, stype=0):
        import scipy.sparse as sp
        M = sp.csc_matrix(M)
        if stype == -1:
            M = sp.tril(M, format="csc")
        M.sort_indices()
        return cls(M.shape[0], M.shape[1], M.indptr, M.indices, M.data, stype)

    @classmethod
    def from_dense(cls, M, stype=0, keep_zeros=True):
        """Dense matrix stored in CSC with every entry explicit (the reference's `dense' case)."""
        M = np.asarray(M, dtype=np.float64)
        nrow, ncol = M.shape
        if stype == -1:
            p = np.zeros(ncol + 1, dtype=np.int64)
            rows, vals = [], []
            for j in range(ncol):
                rows.append(np.arange(j, nrow, dtype=np.int64))
                vals.append(M[j:, j])
                p[j + 1] = p[j] + nrow - j
            return cls(nrow, ncol, p, np.concatenate(rows), np.concatenate(vals), stype)
        p = np.arange(0, (ncol + 1) * nrow, nrow, dtype=np.int64)
        i = np.tile(np.arange(nrow, dtype=np.int64), ncol)
        return cls(nrow, ncol, p, i, np.asfortranarray(M).ravel(order="F"), stype)

    def to_scipy(self):
        import scipy.sparse as sp
        nz = int(self.p[-1])
        M = sp.csc_matrix((self.x[:nz], self.i[:nz], self.p), shape=(self.nrow, self.ncol))
        if self.stype == -1:
            L = sp.tril(M, format="csc")
            M = L + sp.tril(L, -1).T
        return M

    def copy(self):
        return CSC(self.nrow, self.ncol, self.p.copy(), self.i.copy(), self.x.copy(), self.stype)

    def ref(self):
        return C.byref(self.struct)

    def ptr(self):
        return C.pointer(self.struct)


def default_settings_py() -> QPALMSettings:
    """Defaults of /root/reference/include/constants.h:65-116 (used to cross-check the libraries)."""
    s = QPALMSettings()
    s.max_iter, s.inner_max_iter = 10000, 100
    s.eps_abs = s.eps_rel = 1e-4
    s.eps_abs_in = s.eps_rel_in = 1.0
    s.rho = 0.1
    s.eps_prim_inf = s.eps_dual_inf = 1e-5
    s.theta, s.delta, s.sigma_max, s.sigma_init = 0.25, 100.0, 1e9, 20.0
    s.proximal, s.gamma_init, s.gamma_upd, s.gamma_max = 1, 1e7, 10.0, 1e7
    s.scaling, s.nonconvex, s.verbose, s.print_iter, s.warm_start = 10, 0, 1, 1, 0
    s.reset_newton_iter, s.enable_dual_termination = 10000, 0
    s.dual_objective_limit = s.time_limit = 1e20
    s.ordering, s.factorization_method, s.max_rank_update = 0, 2, 160
    s.max_rank_update_fraction = 0.1
    return s
