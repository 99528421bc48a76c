"""Time-to-solution of the config-2 stand-in (problems.grid_qp) through the supernodal sparse path vs the reference build."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle.refbind  # noqa: F401  (registers the "reference" / "oracle" checker libraries)
from qpalm_b200 import abi, problems
from qpalm_b200.interface import Qpalm, load_library, solve_qp

lib = load_library("b200")
lib.qpalm_b200_get_stats.argtypes = [C.POINTER(abi.QPALMWorkspace), C.POINTER(abi.QPALMB200Stats)]
lib.qpalm_b200_prof_enable.argtypes = [C.c_char_p]
lib.qpalm_b200_prof_report.argtypes = [C.c_char_p, C.c_size_t]
PROF = os.environ.get("SPARSE_TIME_PROF", "0") == "1"
for g in [int(a) for a in sys.argv[1:]] or [100]:
    p = problems.grid_qp(g, seed=0)
    s = Qpalm("b200")
    for k, v in p.settings.items():
        setattr(s.settings, k, v)
    s.set_data(p.Q, p.A, p.q, p.bmin, p.bmax)
    t0 = time.perf_counter(); ok = s._allocate_work(); t_setup = time.perf_counter() - t0
    assert ok
    s._solve()
    st0 = abi.QPALMB200Stats(); lib.qpalm_b200_get_stats(s._work, C.byref(st0))
    t0 = time.perf_counter(); s._solve(); t_solve = time.perf_counter() - t0
    st = abi.QPALMB200Stats(); lib.qpalm_b200_get_stats(s._work, C.byref(st))
    r = s.result()
    if PROF:   # third solve with a CUDA-event pair around every launch: per-kernel totals (serialises nothing, adds event overhead)
        lib.qpalm_b200_prof_enable(b"*")
        s._solve()
        buf = C.create_string_buffer(1 << 16)
        lib.qpalm_b200_prof_report(buf, len(buf))
        lib.qpalm_b200_prof_enable(b"")
        prof = json.loads(buf.value.decode() or "{}")
        tot = sum(v["ms"] for v in prof.values())
        print(f"# g={g}: per-kernel CUDA-event time of one solve ({tot:.1f} ms in kernels)")
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:18]:
            print(f"#   {k:28s} launches {v['launches']:7d}  ms {v['ms']:9.2f}  mean_us {1e3 * v['ms'] / v['launches']:8.2f}  {100 * v['ms'] / tot:5.1f}%")
    s.cleanup()
    t0 = time.perf_counter()
    ref = solve_qp("reference", p.Q.copy(), p.A.copy(), p.q.copy(), p.bmin.copy(), p.bmax.copy(), **p.settings)
    t_ref = time.perf_counter() - t0
    ex = np.max(np.abs(r.x - ref.x)) / max(1.0, np.max(np.abs(ref.x)))
    ey = np.max(np.abs(r.y - ref.y)) / max(1.0, np.max(np.abs(ref.y)))
    print(json.dumps(dict(g=g, n=p.n, m=p.m, status=r.status, iter=r.iter, iter_out=r.iter_out, ref_status=ref.status, ref_iter=ref.iter,
                          ref_iter_out=ref.iter_out, setup_s=t_setup, solve_s=t_solve, device_ms=st.device_ms_total - st0.device_ms_total,
                          ref_setup_s=ref.setup_time, ref_solve_s=ref.solve_time, ref_wall_s=t_ref, rel_dx=ex, rel_dy=ey,
                          nnzL=st.sparse_factor_nnz, supernodes=st.sparse_supernodes, levels=st.sparse_levels,
                          refactorizations=st.refactorizations - st0.refactorizations, updown=st.updown_calls - st0.updown_calls,
                          ms_factor=st.device_ms_factor - st0.device_ms_factor, ms_updown=st.device_ms_updown - st0.device_ms_updown,
                          launches=st.kernel_launches - st0.kernel_launches)))
