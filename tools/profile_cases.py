"""Short single-GPU cases that launch every kernel family of the library a few times, for `ncu --set full -k regex:...`
(one capture per change; see profiles/).  Usage: python tools/profile_cases.py [contract] [small] [eig] [iai]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import autobz_b200 as ab
from autobz_b200 import _lib as L

ctx = ab.default_context(0)
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
which = sys.argv[1:] or ["contract", "small", "eig", "iai", "sweep", "ggr", "matrix", "team", "mid"]
d = np.load(os.path.join(ROOT, "tests", "golden", "svo_hr.npz"))
Hs, los, A = np.asfortranarray(d["H_R"]), tuple(int(x) for x in d["lo"]), d["A"]

if "contract" in which:
    # C4 shape: norb 32, M 17, one k3 plane of the 256^3 grid: stage 3 / 2 / 1 contractions, then 8 resolvents per node
    H, lo = ab.synthetic.wannier_hamiltonian(32, 8)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    R = L.DeviceRule(ctx, S, 256, k3_lo=0, k3_hi=1)
    R.materialize()
    ext = ab.synthetic.band_extent(H)
    z = np.linspace(-0.2 * ext, 0.2 * ext, 8) + 1j * 0.01 * ext
    print("contract", R.resolvent_sum(z)[:2], ctx.last_timings())
    R.close(); S.close()

if "fused" in which:
    # C4 shape streamed (not materialised): stages 3 / 2, then K3-fused (stage 1 folded into the DMMA resolvent kernel), 128 frequencies
    # as in the bench (the kernel synchronises its 8 warps once per node: nw / 8 matrices per warp between barriers)
    H, lo = ab.synthetic.wannier_hamiltonian(32, 8)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    R = L.DeviceRule(ctx, S, 256, k3_lo=0, k3_hi=1)
    ext = ab.synthetic.band_extent(H)
    z = np.linspace(-0.2 * ext, 0.2 * ext, 128) + 1j * 0.01 * ext
    print("fused", R.resolvent_sum(z)[:2], ctx.last_timings())
    R.close(); S.close()

if "small" in which:
    fs = ab.FourierSeries(Hs, period=1.0, lo=los, norb=3)
    f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=1e-2)
    ws = [{"omega": w} for w in np.linspace(11.0, 14.0, 16)]
    for npt, dom in ((200, ab.load_bz(ab.FBZ(), A)), (400, ab.load_bz(ab.CubicSymIBZ(), A))):
        solver = ab.IntegralSolver(f, dom, ab.PTR(npt=npt))
        print("small", npt, ab.batchsolve(solver, ws)[:1], ctx.last_timings())

if "eig" in which:
    n = 64
    H, lo = ab.synthetic.wannier_hamiltonian(n, 4, cubic=True)
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=n)
    ibz = ab.load_bz(ab.CubicSymIBZ(), 2 * np.pi * np.eye(3))
    f = ab.FourierIntegrand(ab.EigenIntegrand("fermi_energy"), fs, 0.0, 0.5)
    sol = ab.solve(ab.IntegralProblem(f, ibz), ab.PTR(npt=96))
    print("eig", sol.u, ctx.last_timings())

if "iai" in which:
    fs = ab.FourierSeries(Hs, period=1.0, lo=los, norb=3)
    ibz = ab.load_bz(ab.CubicSymIBZ(), A)
    f = ab.FourierIntegrand(ab.dos_integrand, fs, 0.05)
    for leaves in (True, False):
        be = ab.DeviceBackend(ctx=ctx, iai_engine="native", iai_device_leaves=leaves)
        sol = ab.solve(ab.IntegralProblem(f, ibz, 12.5), ab.EvalCounter(ab.IAI()), abstol=3e-1, backend=be)
        print("iai", leaves, sol.u, sol.numevals)

if "sweep" in which:
    # C4 shape, one k3 plane, 128 frequencies through the opt-in sweep path (tridiagonalise once per k, p'/p per frequency)
    H, lo = ab.synthetic.wannier_hamiltonian(32, 8)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    R = L.DeviceRule(ctx, S, 256, k3_lo=0, k3_hi=1)
    R.materialize()
    ext = ab.synthetic.band_extent(H)
    z = np.linspace(-0.2 * ext, 0.2 * ext, 128) + 1j * 0.01 * ext
    ctx.set_option(L.OPT_RESOLVENT_ALGO, 3)
    print("sweep", R.resolvent_sum(z)[:2], ctx.last_timings())
    ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)
    R.close(); S.close()

if "ggr" in which:
    # GGR data pass + sum: graphene-like 2-band model is too small to time; use a 16-orbital cubic model on the IBZ, npt = 32
    H, lo = ab.synthetic.wannier_hamiltonian(16, 2, cubic=True)
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=16)
    cache = ab.init(ab.DOSProblem(fs, np.linspace(-1.0, 1.0, 64), ab.load_bz(ab.CubicSymIBZ(), np.eye(3))), ab.GGR(npt=32))
    print("ggr", ab.solve_(cache).u[:3], ctx.last_timings())

if "matrix" in which:
    H, lo = ab.synthetic.wannier_hamiltonian(32, 2)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    R = L.DeviceRule(ctx, S, 24)
    z = np.linspace(-1.0, 1.0, 8) + 0.05j
    print("matrix", R.resolvent_matrix_sum(z)[0, 0, :2], ctx.last_timings())
    R.close(); S.close()

if "team" in which:
    # norb 64 resolvent traces: the four-warp DMMA block-LU team kernel (32 < norb <= 64), 32^3 grid x 8 frequencies
    H, lo = ab.synthetic.wannier_hamiltonian(64, 2)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    R = L.DeviceRule(ctx, S, 32)
    R.materialize()
    ext = ab.synthetic.band_extent(H)
    z = np.linspace(-0.2 * ext, 0.2 * ext, 8) + 1j * 0.01 * ext
    print("team", R.resolvent_sum(z)[:2], ctx.last_timings())
    R.close(); S.close()

if "mid" in which:
    # C3 (SrVO3 DOS by IAI on the cubic IBZ, eta = 1e-3): device-side middle integrals (iai_mid_kernel, one 2-CTA cluster per task)
    fs = ab.FourierSeries(Hs, period=1.0, lo=los, norb=3)
    ibz = ab.load_bz(ab.CubicSymIBZ(), A)
    f = ab.FourierIntegrand(ab.dos_integrand, fs, 1e-3)
    sol = ab.solve(ab.IntegralProblem(f, ibz, 12.0), ab.EvalCounter(ab.IAI()), abstol=1e-3)
    print("mid", sol.u, sol.numevals)
