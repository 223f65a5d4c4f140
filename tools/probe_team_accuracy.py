"""accuracy of the unpivoted DMMA team resolvent (32 < norb <= 64) against the pivoted teams as eta shrinks, per (k, omega) value;
whether the pivot monitor sent the call to the pivoted path is visible in the launch count.  Usage: python tools/probe_team_accuracy.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import autobz_b200 as ab
from autobz_b200 import _lib as L

ctx = ab.default_context(0)
rng = np.random.default_rng(1)
print("ABZ_MMA_TEAM_PIVOT_THR =", os.environ.get("ABZ_MMA_TEAM_PIVOT_THR", "(default)"))
for n in (40, 64):
    for rmax in (1, 2):
        H, lo = ab.synthetic.wannier_hamiltonian(n, rmax)
        S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
        ext = ab.synthetic.band_extent(H)
        kp = rng.random((256, 3))
        for eta_rel in (3e-2, 1e-2, 3e-3, 1e-3, 3e-4, 1e-4):
            z = rng.uniform(-0.6 * ext, 0.6 * ext, 16) + 1j * eta_rel * ext
            ctx.set_option(L.OPT_RESOLVENT_ALGO, 1)
            ref = S.points_resolvent(kp, z)
            ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)
            l0 = ctx.launch_count
            got = S.points_resolvent(kp, z)
            dl = ctx.launch_count - l0
            err = np.abs(got - ref) / np.abs(ref)
            print(f"n={n} rmax={rmax} eta/ext={eta_rel:7.0e}  max rel err {err.max():.2e}  median {np.median(err):.2e}  launches {dl} {'(pivoted rerun)' if dl > 2 else ''}", flush=True)
        S.close()
