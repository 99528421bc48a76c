"""TEST INFRASTRUCTURE -- ctypes bindings of the two checker libraries.

* ``"reference"``  oracle/_ref/libqpalm_ref.so: the UNMODIFIED reference (CHOLMOD build) compiled by oracle/Makefile,
* ``"oracle"``     oracle/liboracle.so: the plain-C restatement (``oracle_`` symbol prefix),
* the unmodified reference QPS reader (oracle/_ref/libqpalm_qps_ref.so).

Importing this module registers the two libraries with ``qpalm_b200.interface`` so that ``Qpalm("reference")`` /
``solve_qp("oracle", ...)`` work.  Only tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference leg and
the golden-vector scripts import it; nothing under qpalm_b200/ does (tests/test_abi.py checks that), and the product's
``Qpalm("b200")`` never reaches these libraries.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import sysconfig

import numpy as np

from qpalm_b200 import abi, interface
from qpalm_b200.abi import QPALMData, QPALMSettings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libqpalm_ref.so")
REF_QPS_LIB = os.path.join(HERE, "_ref", "libqpalm_qps_ref.so")
ORACLE_LIB = os.path.join(HERE, "liboracle.so")


_BLAS = []


def preload_blas():
    """The compiled reference links the OpenBLAS bundled in the opencv wheel (oracle/Makefile); that library needs its
    sibling libgfortran/libquadmath, which are not on the loader path."""
    d = os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs")
    for pat in ("libquadmath*", "libgfortran*", "libopenblasp*"):
        for f in sorted(glob.glob(os.path.join(d, pat))):
            try:
                h = C.CDLL(f, mode=os.RTLD_GLOBAL | os.RTLD_NOW)
                if "openblas" in os.path.basename(f):
                    _BLAS.append(h)
            except OSError:
                pass


def set_blas_threads(n: int) -> None:
    """Thread count of the OpenBLAS the compiled reference uses (its supernodal Cholesky calls BLAS-3); QPALM itself is
    single-threaded.  bench.py: 1 per worker process for the batch sample, all host cores for the dense sample."""
    for h in _BLAS:
        try:
            h.openblas_set_num_threads(int(n))
        except AttributeError:
            pass


def have_reference() -> bool:
    return os.path.exists(REF_LIB)


def _load_reference():
    preload_blas()
    return C.CDLL(REF_LIB, mode=os.RTLD_LOCAL | os.RTLD_NOW)


def _load_oracle():
    return C.CDLL(ORACLE_LIB, mode=os.RTLD_LOCAL | os.RTLD_NOW)


interface.register_checker("reference", REF_LIB, "", _load_reference)
interface.register_checker("oracle", ORACLE_LIB, "oracle_", _load_oracle)

Qpalm = interface.Qpalm
solve_qp = interface.solve_qp

# ---------------------------------------------------------------------------------------------
# the unmodified reference QPS reader (interfaces/qps/src/qpalm_qps.c through oracle/qps_ref_shim.c)
# ---------------------------------------------------------------------------------------------
_REF = None


def reference_reader_available() -> bool:
    return os.path.exists(REF_QPS_LIB) and os.path.exists(REF_LIB)


def _ref():
    global _REF
    if _REF is None:
        preload_blas()
        C.CDLL(REF_LIB, mode=os.RTLD_GLOBAL | os.RTLD_NOW)
        lib = C.CDLL(REF_QPS_LIB, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        lib.qps_ref_read.argtypes = [C.c_char_p]
        lib.qps_ref_read.restype = C.POINTER(QPALMData)
        lib.qps_ref_read_settings.argtypes = [C.c_char_p, C.POINTER(QPALMSettings)]
        lib.qps_ref_read_settings.restype = None
        _REF = lib
    return _REF


def read_qps_reference(path: str):
    from qpalm_b200 import qps
    d = _ref().qps_ref_read(os.fsencode(path))
    if not d:
        raise RuntimeError(f"reference reader failed on {path}")
    return qps._unpack(d.contents, "")          # the few reference-owned buffers are left to the process


def read_settings_reference(path: str) -> dict:
    s = QPALMSettings()
    _ref().qps_ref_read_settings(os.fsencode(path), C.byref(s))
    return {f: getattr(s, f) for f, _ in QPALMSettings._fields_}
