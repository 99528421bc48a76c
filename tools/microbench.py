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
