#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_iai_norb6.py -m gpu -x -q > gpurun_out/r2_31_tests_norb6.log 2>&1; echo "norb6 tests rc=$?"; tail -n 15 gpurun_out/r2_31_tests_norb6.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_iai_middles.py tests/test_gpu_general_limits.py tests/test_gpu_matrix_iai.py tests/test_gpu_gk_orders.py -m gpu -x -q > gpurun_out/r2_31_tests_iai.log 2>&1; echo "iai tests rc=$?"; tail -n 3 gpurun_out/r2_31_tests_iai.log
python - <<'PY' > gpurun_out/r2_31_norb6_timing.log 2>&1
import time, numpy as np
import autobz_b200 as ab
ctx = ab.default_context(0)
print("# 3-d IAI, DOS integrand on the cubic IBZ, synthetic cubic Wannier model (R in [-2,2]^3), eta = 0.05, abstol 1e-3 (x (2 pi)^3 48): seconds per solve")
for n in (3, 4, 5, 6):
    H, lo = ab.synthetic.wannier_hamiltonian(n, 2, cubic=True)
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=n)
    ibz = ab.load_bz(ab.CubicSymIBZ(), np.eye(3))
    f = ab.FourierIntegrand(ab.dos_integrand, fs, 0.05)
    row = []
    for leaves, middles, spec in ((True, True, True), (True, False, False), (False, False, False)):
        be = ab.DeviceBackend(ctx=ctx, iai_engine="native", iai_device_leaves=leaves, iai_device_middles=middles, iai_speculate=spec)
        best = 1e30
        for _ in range(2):
            t = time.perf_counter()
            sol = ab.solve(ab.IntegralProblem(f, ibz, 0.3), ab.EvalCounter(ab.IAI()), abstol=1e-3 * (2 * np.pi) ** 3 * 48, backend=be)
            best = min(best, time.perf_counter() - t)
        row.append((best, sol.numevals))
    print(f"norb {n}: numevals {row[0][1]}  middles+look-ahead {row[0][0]:.4f} s  leaves only {row[1][0]:.4f} s  host-driven panels {row[2][0]:.4f} s  (numevals equal: {row[0][1] == row[1][1] == row[2][1]})")
PY
cat gpurun_out/r2_31_norb6_timing.log
