"""C3 (SrVO3 DOS via IAI, eta = 1e-4, abstol = 1e-3) with and without device-side middle integrals and look-ahead on the outermost
integral: time, numevals, device rounds.
Usage: python tools/time_c3.py"""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import autobz_b200 as ab

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
d = np.load(os.path.join(ROOT, "tests", "golden", "svo_hr.npz"))
H, lo, A = np.asfortranarray(d["H_R"]), tuple(int(x) for x in d["lo"]), d["A"]
fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
ibz = ab.load_bz(ab.CubicSymIBZ(), A)
f = ab.FourierIntegrand(ab.dos_integrand, fs, 1e-4)
ctx = ab.default_context(0)
for omega in (12.0, 12.975161):
    for mids, spec in ((False, False), (False, True), (True, False), (True, True)):
        be = ab.DeviceBackend(ctx=ctx, iai_engine="native", iai_device_leaves=True, iai_device_middles=mids, iai_speculate=spec)
        best, sol, rounds = 1e30, None, 0
        for _ in range(3):
            cache = ab.init(ab.IntegralProblem(f, ibz, omega), ab.EvalCounter(ab.IAI()), abstol=1e-3, backend=be)
            t0 = time.perf_counter(); sol = ab.solve_(cache); best = min(best, time.perf_counter() - t0)
            rounds = cache.cacheval["iai_rounds"]
        print(json.dumps({"omega": omega, "device_middles": mids, "lookahead": spec, "s": best, "numevals": sol.numevals, "rounds": rounds,
                          "evals_per_s": sol.numevals / best, "u": sol.u}), flush=True)
