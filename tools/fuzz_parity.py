"""Randomised differential check of the C ABI against the CPU oracle (test infrastructure, like tests/): random orbital counts,
coefficient ranges, grid sizes, frequency lists, self-energies and resolvent algorithms; rule sums (full grid, k3 slabs),
per-point values, matrix-valued sums and eigenvalue sums.  Usage: python tools/fuzz_parity.py [cases=200] [seed=0]"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import autobz_b200 as ab
from autobz_b200 import _lib as L
import orc

ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = ab.default_context(0)
rng = np.random.default_rng(seed)


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


worst, fails, t0 = {}, 0, time.time()
for case in range(ncase):
    n = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 12, 13, 16, 17, 24, 31, 32, 33, 40, 48, 63, 64]))
    rmax = int(rng.integers(0, 3))
    N = int(rng.integers(1, 7 if n <= 32 else 5))
    nw = int(rng.choice([1, 2, 3, 9, 17]))
    # a slice of the cases on grids with several 32-node phase tiles per (k2,k3) row (the multi-tile TMA pipeline of stage 1)
    if n <= 8 and rng.integers(0, 5) == 0:
        N = int(rng.integers(33, 101))
    elif n <= 32 and rng.integers(0, 8) == 0:
        N, nw = int(rng.integers(33, 49)), min(nw, 3)
    eta = float(rng.choice([0.3, 0.03, 2e-3]))
    H, lo = ab.synthetic.wannier_hamiltonian(n, rmax)
    ext = ab.synthetic.band_extent(H)
    z = rng.uniform(-1.2 * ext, 1.2 * ext, nw) + 1j * eta * ext
    use_sig = bool(rng.integers(0, 2))
    sig = None
    if use_sig:
        sig = 0.05 * ext * (rng.standard_normal((n, n, nw)) + 1j * rng.standard_normal((n, n, nw)))
        sig = sig - 0.05j * ext * np.eye(n)[:, :, None]
    algos = [0, 1] + ([4] if n <= 64 else []) + ([2] if 4 <= n <= 32 else []) + ([3] if not use_sig else [])
    algo = int(rng.choice(algos))
    tag = f"case {case}: n={n} rmax={rmax} N={N} nw={nw} eta={eta} sigma={use_sig} algo={algo}"
    try:
        S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
        So = orc.Series(H, lo)
        ctx.set_option(L.OPT_RESOLVENT_ALGO, algo)
        R = L.DeviceRule(ctx, S, N)
        ref = orc.ptr_sum(So, N, z, sigma=sig)
        errs = {"sum": rel(R.resolvent_sum(z, sigma=sig, scale=1 / N ** 3), ref)}
        if N >= 2:
            a = int(rng.integers(1, N))
            parts = [L.DeviceRule(ctx, S, N, k3_lo=lo_, k3_hi=hi_).resolvent_sum(z, sigma=sig, scale=1 / N ** 3) for lo_, hi_ in ((0, a), (a, N))]
            errs["slabs"] = rel(sum(parts), ref)
        kp = rng.random((int(rng.integers(1, 12)), 3))
        Hk = orc.eval_points(So, kp)
        errs["points"] = rel(S.points_resolvent(kp, z, sigma=sig), orc.resolvent_trace_batch(Hk, z, sig))
        if rng.integers(0, 3) == 0:
            G = R.resolvent_matrix_sum(z, sigma=sig, scale=1 / N ** 3)
            Hg = orc.grid_eval_full(So, N).reshape(n, n, -1, order="F")
            want = np.zeros((nw, n, n), complex)
            for w in range(nw):
                A = z[w] * np.eye(n)[None] - np.moveaxis(Hg, 2, 0) - (0 if sig is None else sig[:, :, w][None])
                want[w] = np.linalg.inv(A).sum(0) / N ** 3
            errs["matrix"] = rel(G, want)
        if rng.integers(0, 3) == 0:
            kind = int(rng.integers(0, 4))
            prm = (0.1 * ext, 0.3 * ext)
            errs["eig"] = abs(R.eig_sum(kind, prm, 1 / N ** 3) - orc.ptr_eig_sum(So, N, kind, prm, scale=1 / N ** 3)[0]) / max(1.0, n * ext)
        tol = 2e-10 if eta <= 2e-3 else 1e-10
        bad = {k: v for k, v in errs.items() if not (v < tol)}
        for k, v in errs.items():
            worst[k] = max(worst.get(k, 0.0), v if np.isfinite(v) else np.inf)
        if bad:
            fails += 1
            print("FAIL", tag, bad, flush=True)
        R.close(); S.close()
    except Exception as e:                                  # noqa: BLE001
        fails += 1
        print("ERROR", tag, type(e).__name__, str(e)[:200], flush=True)
    finally:
        ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)
print(f"{ncase} cases, {fails} failures, worst relative errors {worst}, {time.time() - t0:.1f} s")

# ---- symmetry-reduced rules through the public API: the IBZ (device-side symptr_rule) must reproduce the full-BZ integral
fails2, worst2 = 0, 0.0
for case in range(max(10, ncase // 8)):
    n = int(rng.choice([1, 2, 3, 5, 8, 16, 32, 33, 64]))
    npt = int(rng.integers(2, 14 if n <= 8 else 8))
    H, lo = ab.synthetic.wannier_hamiltonian(n, int(rng.integers(1, 3)), cubic=True)
    ext = ab.synthetic.band_extent(H)
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=n)
    A = np.eye(3) * float(rng.uniform(0.5, 2.0))
    tag = f"sym case {case}: n={n} npt={npt}"
    try:
        if rng.integers(0, 2):
            f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=0.05 * ext)
            p = {"omega": float(rng.uniform(-ext, ext))}
        else:
            f = ab.FourierIntegrand(ab.EigenIntegrand("fermi_energy"), fs, 0.1 * ext, 0.2 * ext)
            p = None
        vals = [ab.solve(ab.IntegralProblem(f, ab.load_bz(kind, A), p), ab.PTR(npt=npt)).u for kind in (ab.FBZ(), ab.InversionSymIBZ(), ab.CubicSymIBZ())]
        e = max(rel(vals[1], vals[0]), rel(vals[2], vals[0]))
        worst2 = max(worst2, e)
        if not e < 1e-10:
            fails2 += 1
            print("FAIL", tag, e, vals, flush=True)
    except Exception as ex:                                 # noqa: BLE001
        fails2 += 1
        print("ERROR", tag, type(ex).__name__, str(ex)[:200], flush=True)
print(f"symmetric rules: {max(10, ncase // 8)} cases, {fails2} failures, worst IBZ-vs-FBZ relative difference {worst2}")
sys.exit(1 if fails or fails2 else 0)
