"""Times the DMMA team resolvent (one matrix per team of warps, abz_resolvent_mma_team.cuh) against the pivoted Gauss-Jordan teams
(algo 1) for 32 < norb <= 64, and - with ABZ_MMA_TEAM=1 in the environment - the two-warp teams against the one-warp DMMA kernel at
norb = 32.  Usage: python tools/time_team_resolvent.py [norb ...]"""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import autobz_b200 as ab
from autobz_b200 import _lib as L

ctx = ab.default_context(0)


def best_ms(fn, reps=4):
    fn()
    ts = []
    for _ in range(reps):
        fn(); ts.append(ctx.last_timings()[1])
    return min(ts)


norbs = [int(a) for a in sys.argv[1:]] or [64, 56, 48, 40, 32]
for n in norbs:
    N, nw = (24, 64) if n > 32 else (32, 128)
    H, lo = ab.synthetic.wannier_hamiltonian(n, 2)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    R = L.DeviceRule(ctx, S, N)
    R.materialize()
    ext = ab.synthetic.band_extent(H)
    z = np.linspace(-0.25 * ext, 0.25 * ext, nw) + 1j * 0.005 * ext
    nmat = N ** 3 * nw
    res = {}
    for algo in (0, 1):
        ctx.set_option(L.OPT_RESOLVENT_ALGO, algo)
        v = R.resolvent_sum(z)
        ms = best_ms(lambda: R.resolvent_sum(z))
        res[algo] = (v, ms)
    ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)
    print(json.dumps({"norb": n, "matrices": nmat, "team_env": os.environ.get("ABZ_MMA_TEAM", ""),
                      "ms_default": res[0][1], "ms_pivoted_gj": res[1][1],
                      "tflops_credited_default": 8 * n ** 3 * nmat / res[0][1] / 1e9, "tflops_credited_pivoted_gj": 8 * n ** 3 * nmat / res[1][1] / 1e9,
                      "rel_diff": float(np.max(np.abs(res[0][0] - res[1][0]) / np.abs(res[1][0])))}), flush=True)
    R.close(); S.close()
