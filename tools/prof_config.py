"""Per-kernel CUDA-event breakdown of one solve of a BASELINE single-problem config (c1 | c3 | c5 | grid<g>): launches and device time per
kernel name, next to the un-profiled wall time of the same solve (the difference is launch latency + host control flow)."""
import ctypes as C
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpalm_b200 import problems
from qpalm_b200.interface import Qpalm, load_library

which = sys.argv[1] if len(sys.argv) > 1 else "c1"
if which == "c1":
    p = problems.random_qp(1000, 2000, 0.05, 0.007, seed=0)
elif which == "c3":
    p = problems.dense_qp(8000, 16000, seed=0)
elif which == "c5":
    p = problems.nonconvex_random_qp(5000, 10000, seed=1)
elif which.startswith("grid"):
    p = problems.grid_qp(int(which[4:] or 150), seed=0)
else:
    raise SystemExit("c1 | c3 | c5 | grid<g>")
lib = load_library("b200")


def solve(profile):
    s = Qpalm("b200")
    for k, v in p.settings.items():
        setattr(s.settings, k, v)
    s.set_data(p.Q.copy(), p.A.copy(), p.q.copy(), p.bmin.copy(), p.bmax.copy(), p.c)
    assert s._allocate_work()
    if profile:
        lib.qpalm_b200_prof_enable(b"*")
    t0 = time.perf_counter(); s._solve(); dt = time.perf_counter() - t0
    r = s.result()
    rep = None
    if profile:
        buf = C.create_string_buffer(1 << 18)
        lib.qpalm_b200_prof_report(buf, len(buf))
        rep = json.loads(buf.value.decode())
        lib.qpalm_b200_prof_enable(b"")
    s.cleanup()
    return r, dt, rep


solve(False)
r, dt, _ = min((solve(False) for _ in range(3)), key=lambda t: t[1])
print(f"{which}: n={p.n} m={p.m} status={r.status} iter={r.iter}/{r.iter_out} solve wall {1e3 * dt:.2f} ms = {1e3 * dt / max(1, r.iter):.3f} ms per iteration")
r, dtp, rep = solve(True)
tot_ms = sum(v["ms"] for v in rep.values()); tot_l = sum(v["launches"] for v in rep.values())
print(f"profiled: {tot_l} launches ({tot_l / max(1, r.iter):.1f} per iteration), device time in kernels {tot_ms:.2f} ms")
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])[:40]:
    print(f"  {k:34s} launches {v['launches']:6d}  ms {v['ms']:9.3f}  mean_us {1e3 * v['ms'] / v['launches']:8.2f}")

if os.environ.get("QPALM_B200_MF_CLOCKS"):
    lib.qpalm_b200_mf_clocks.argtypes = [C.POINTER(C.c_ulonglong)]
    buf = (C.c_ulonglong * 16)()
    lib.qpalm_b200_mf_clocks(buf)      # clear what the solves above accumulated
    solve(False)
    lib.qpalm_b200_mf_clocks(buf)
    names = ["extend", "wait B0", "block 0 + sync", "Ls load", "trsm", "wait B1", "chain (CTA 0)", "tiles (CTA 1)", "wait B2 (CTA 0)",
             "wait B2 (CTA 1)"]
    steps, launches = max(1, buf[10]), max(1, buf[11])
    print(f"k_mf_front phase clocks of the first front of each level: {launches} launches, {steps} block steps")
    for i, nm in enumerate(names):
        per = "launch" if i < 3 else "step"
        print(f"  {nm:18s} {buf[i] / (launches if i < 3 else steps):10.0f} clocks per {per}   total {buf[i] / 1.965e6:9.2f} ms")
