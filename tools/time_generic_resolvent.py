"""Times the pivoted (generic) resolvent paths: register-resident Gauss-Jordan (algo 1) against the earlier shared-memory
formulation (algo 4), trace sums for norb in {8, 24, 32, 48, 64} and the matrix-valued Green's function sums.
Usage: python tools/time_generic_resolvent.py"""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import autobz_b200 as ab
from autobz_b200 import _lib as L

ctx = ab.default_context(0)


def best_ms(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return 1e3 * min(ts)


CASES = ((8, 48, 8), (24, 40, 8), (32, 40, 8), (48, 28, 8), (64, 28, 8))
if len(sys.argv) > 1:      # python tools/time_generic_resolvent.py 12 16 20: other orbital counts on a 40^3 grid
    CASES = tuple((int(a), 40, 8) for a in sys.argv[1:])
for n, N, nw in CASES:
    H, lo = ab.synthetic.wannier_hamiltonian(n, 2)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    R = L.DeviceRule(ctx, S, N)
    R.materialize()
    ext = ab.synthetic.band_extent(H)
    z = np.linspace(-0.2 * ext, 0.2 * ext, nw) + 1j * 0.01 * ext
    out = {}
    for algo in (1, 4):
        ctx.set_option(L.OPT_RESOLVENT_ALGO, algo)
        v = R.resolvent_sum(z)
        ms = best_ms(lambda: R.resolvent_sum(z))
        out[algo] = (v, ms)
    ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)
    nmat = N ** 3 * nw
    rel = float(np.max(np.abs(out[1][0] - out[4][0]) / np.abs(out[4][0])))
    print(json.dumps({"case": "trace", "norb": n, "matrices": nmat, "ms_reg": out[1][1], "ms_smem": out[4][1],
                      "Mmat_per_s_reg": nmat / out[1][1] / 1e3, "tflops_credited_reg": 8 * n ** 3 * nmat / out[1][1] / 1e9,
                      "speedup": out[4][1] / out[1][1], "rel_diff": rel}))
    out = {}
    for algo in (1, 4):
        ctx.set_option(L.OPT_RESOLVENT_ALGO, algo)
        v = R.resolvent_matrix_sum(z)
        ms = best_ms(lambda: R.resolvent_matrix_sum(z), reps=3)
        out[algo] = (v, ms)
    ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)
    rel = float(np.max(np.abs(out[1][0] - out[4][0])) / np.max(np.abs(out[4][0])))
    print(json.dumps({"case": "matrix", "norb": n, "matrices": nmat, "ms_reg": out[1][1], "ms_smem": out[4][1],
                      "Mmat_per_s_reg": nmat / out[1][1] / 1e3, "speedup": out[4][1] / out[1][1], "rel_diff": rel}))
    if n <= 32:      # the unpivoted DMMA fast path on the same matrices, for scale
        ms = best_ms(lambda: R.resolvent_sum(z))
        print(json.dumps({"case": "trace, DMMA fast path (default)", "norb": n, "matrices": nmat, "ms": ms, "Mmat_per_s": nmat / ms / 1e3}))
    R.close(); S.close()
