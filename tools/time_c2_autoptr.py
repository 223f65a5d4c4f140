"""C2 (SrVO3 Green's-function trace, AutoPTR a = eta = 1e-2 on the cubic IBZ): per-solve wall time over repeated solves, and the
time of its parts (rule construction per grid size, sums).  Usage: python tools/time_c2_autoptr.py"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import autobz_b200 as ab
from autobz_b200 import _lib as L

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
d = np.load(os.path.join(ROOT, "tests", "golden", "svo_hr.npz"))
H, lo, A = np.asfortranarray(d["H_R"]), tuple(int(x) for x in d["lo"]), d["A"]
fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
ibz = ab.load_bz(ab.CubicSymIBZ(), A)
ctx = ab.default_context(0)
f2 = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=1e-2)
alg = ab.EvalCounter(ab.AutoPTR(a=1e-2, nmin=50, nmax=1000))
ts = []
for i in range(12):
    t = time.perf_counter()
    sol = ab.solve(ab.IntegralProblem(f2, ibz, {"omega": 12.5}), alg, abstol=1e-3)
    ts.append(1e3 * (time.perf_counter() - t))
print("solve ms:", " ".join(f"{x:.1f}" for x in ts), "numevals", sol.numevals)
S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
for npt in (600, 831):
    tt = []
    for i in range(6):
        t = time.perf_counter()
        R = L.DeviceRule(ctx, S, npt, syms=ibz.syms)
        t1 = time.perf_counter()
        v = R.resolvent_sum(np.array([12.5 + 0.01j]))
        t2 = time.perf_counter()
        R.close()
        t3 = time.perf_counter()
        tt.append((1e3 * (t1 - t), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
    print(f"npt {npt}: (rule, sum, close) ms:", " ".join(f"({a:.2f},{b:.2f},{c:.2f})" for a, b, c in tt))
