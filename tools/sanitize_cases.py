"""Tiny invocations of every kernel family, sized for `compute-sanitizer --tool memcheck` (one tool per gpurun call):
   compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_cases.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import autobz_b200 as ab
from autobz_b200 import _lib as L

ctx = ab.default_context(0)
syms = np.array(ab.cube_automorphisms(3), dtype=np.int32)
z = np.array([0.3 + 0.2j, -0.4 + 0.1j, 0.9 + 0.3j])
if "multitile" in sys.argv:
    # stage 1 of the contraction with several phase tiles per row: the double-buffered TMA bulk copies and their mbarrier
    # parity flips (N = 70: 3 tiles, the last one partial), full grid and a symmetric rule with rows longer than one tile
    for n, rmax in ((4, 8), (32, 1)):
        H, lo = ab.synthetic.wannier_hamiltonian(n, rmax, cubic=True)
        S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
        for R in (L.DeviceRule(ctx, S, 70, k3_lo=3, k3_hi=4), L.DeviceRule(ctx, S, 70, syms=syms, k3_lo=0, k3_stride=35)):
            R.resolvent_sum(z)
            R.copy_out()
            R.close()
        S.close()
    print("sanitize multitile cases done, launches:", ctx.launch_count)
    sys.exit(0)
for n, N in ((1, 5), (2, 5), (3, 6), (5, 4), (17, 3), (32, 3), (33, 3), (64, 2)):
    H, lo = ab.synthetic.wannier_hamiltonian(n, 1, cubic=True)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    for R in (L.DeviceRule(ctx, S, N), L.DeviceRule(ctx, S, N, syms=syms), L.DeviceRule(ctx, S, N, syms=syms, k3_lo=1, k3_stride=2)):
        R.resolvent_sum(z)
        sig = 0.05 * np.ones((n, n, 3), dtype=complex)
        R.resolvent_sum(z, sigma=sig)
        R.resolvent_sum(None, fkind=L.F_TRACE_H)
        R.resolvent_matrix_sum(z[:2])
        R.eig_sum(L.EIG_FERMI_ENERGY, (0.1, 0.3))
        R.eigvals()
        R.ggr_data(3)
        R.ggr_sum([0.0, 0.2])
        for algo in (1, 3):
            ctx.set_option(L.OPT_RESOLVENT_ALGO, algo)
            R.resolvent_sum(z)
            ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)
        for ea in (1, 2):
            ctx.set_option(L.OPT_EIG_ALGO, ea)
            R.eigvals()
            ctx.set_option(L.OPT_EIG_ALGO, 0)
        R.copy_out()
        R.materialize()
        R.resolvent_sum(z)
        R.close()
    S.points_resolvent(np.random.default_rng(0).random((5, 3)), z)
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=n)
    ibz = ab.load_bz(ab.CubicSymIBZ(), np.eye(3))
    if n in (1, 3, 5):
        for leaves in (True, False):
            be = ab.DeviceBackend(ctx=ctx, iai_engine="native", iai_device_leaves=leaves)
            ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.dos_integrand, fs, 0.3), ibz, 0.2), ab.EvalCounter(ab.IAI()), abstol=1.0, backend=be)
        be = ab.DeviceBackend(ctx=ctx, iai_engine="python")
        ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.dos_integrand, fs, 0.3), ibz, 0.2), ab.IAI(), abstol=5.0, backend=be)
    S.close()
print("sanitize cases done, launches:", ctx.launch_count)
