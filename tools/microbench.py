import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpalm_b200.interface import load_library
lib = load_library("b200")
out = (C.c_longlong * 5)()
for ctas, thr in ((1, 32), (1, 256), (1, 1024), (148, 256), (148 * 3, 256), (148, 1024)):
    lib.qpalm_b200_microbench(ctas, thr, out)
    print(f"ctas={ctas} threads={thr}: dfma_lat={out[0]} sqrt={out[1]} rcp={out[2]} shfl+add={out[3]} 2048 indep DFMA/thread: {out[4]} clks -> {2048*thr/ max(out[4],1):.1f} DFMA/clk/SM")
o2 = (C.c_longlong * 2)()
lib.qpalm_b200_microbench_diag(o2)
print(f"16x16 in-warp Cholesky: rolled smem {o2[0]} clks, unrolled registers {o2[1]} clks")
if hasattr(lib, "qpalm_b200_microbench_diag_phases"):
    o = (C.c_longlong * 32)()
    lib.qpalm_b200_microbench_diag_phases(o)
    t = [o[i] - o[0] for i in range(32) if o[i]]
    names = ["start", "load"]
    for kb in range(4):
        names += [f"kb{kb}:factor32", f"kb{kb}:inv32", f"kb{kb}:S1 barrier(R)"]
        if kb < 3: names += [f"kb{kb}:X+panel_solve"]
        names += [f"kb{kb}:trailing/end"]
    names += ["stores"]
    prev = 0
    for nm, v in zip(names, t):
        print(f"  diag phase {nm:24s} +{v - prev:7d} clks  (t={v})")
        prev = v
