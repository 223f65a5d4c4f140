"""Stage B of the eigenvalue path: thread-per-matrix QL (ABZ_OPT_EIG_ALGO 3) against warp-per-matrix Sturm bisection (4) for several batch
sizes and orbital counts; prints the matrix-function time of abz_rule_eig_sum (tridiagonalisation + stage B, CUDA events) per case."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import autobz_b200 as ab
from autobz_b200 import _lib as L

ctx = ab.default_context(0)
ibz = ab.load_bz(ab.CubicSymIBZ(), np.eye(3))
for n in (64, 48, 32, 16):
    H, lo = ab.synthetic.wannier_hamiltonian(n, 2, cubic=True)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    for npt in (24, 48, 72, 96, 120, 144):
        R = L.DeviceRule(ctx, S, npt, syms=ibz.syms)
        row = {}
        for algo in (3, 4):
            ctx.set_option(L.OPT_EIG_ALGO, algo)
            best = 1e9
            for rep in range(4):
                v = R.eig_sum(1, (0.0, 0.5))
                best = min(best, ctx.last_timings()[1])
            row[algo] = (best, v)
        ctx.set_option(L.OPT_EIG_ALGO, 0)
        print(f"n={n} npt={npt} nodes={R.nnodes}: QL {row[3][0]:.3f} ms, bisection {row[4][0]:.3f} ms, rel diff {abs(row[3][1] - row[4][1]) / abs(row[3][1]):.1e}", flush=True)
        R.close()
    S.close()
