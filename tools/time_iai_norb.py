"""3-d IAI of the DOS integrand on the cubic IBZ for synthetic cubic Wannier models with 3 ... 6 orbitals: seconds per solve with
device-side middle integrals + look-ahead, device-side innermost integrals only, and host-driven panels (abz_iai_solve flags).
usage: python tools/time_iai_norb.py [eta] [abstol]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import autobz_b200 as ab

eta = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
atol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-5
ctx = ab.default_context(0)
mult = (2 * np.pi) ** 3 * 48
print(f"# 3-d IAI, DOS integrand on the cubic IBZ, synthetic cubic Wannier model (R in [-2,2]^3), eta = {eta}, abstol = {atol} (per unit IBZ volume): best of 2, seconds per solve")
for n in (3, 4, 5, 6):
    H, lo = ab.synthetic.wannier_hamiltonian(n, 2, cubic=True)
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=n)
    ibz = ab.load_bz(ab.CubicSymIBZ(), np.eye(3))
    f = ab.FourierIntegrand(ab.dos_integrand, fs, eta)
    row = []
    for leaves, middles, spec in ((True, True, True), (True, False, False), (False, False, False)):
        be = ab.DeviceBackend(ctx=ctx, iai_engine="native", iai_device_leaves=leaves, iai_device_middles=middles, iai_speculate=spec)
        best = 1e30
        for _ in range(2):
            t = time.perf_counter()
            sol = ab.solve(ab.IntegralProblem(f, ibz, 0.3), ab.EvalCounter(ab.IAI()), abstol=atol * mult, backend=be)
            best = min(best, time.perf_counter() - t)
        row.append((best, sol.numevals))
    print(f"norb {n}: numevals {row[0][1]:>10d}  middles+look-ahead {row[0][0]:.4f} s ({row[0][1] / row[0][0] / 1e6:.0f} M evals/s)  leaves only {row[1][0]:.4f} s  "
          f"host-driven panels {row[2][0]:.4f} s  (numevals equal: {row[0][1] == row[1][1] == row[2][1]})", flush=True)
