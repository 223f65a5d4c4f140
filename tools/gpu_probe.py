"""ad-hoc GPU probes (development aid): accuracy of the unpivoted DMMA team kernel on single matrices"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import autobz_b200 as ab
from autobz_b200 import _lib as L

ctx = ab.default_context(0)
for n in (32, 40, 64):
    for cubic in (True, False):
        H, lo = ab.synthetic.wannier_hamiltonian(n, 1, cubic=cubic)
        S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
        z = np.array([0.2 + 0.3j, 0.5 + 0.1j, 0.1 + 0.01j])
        R = L.DeviceRule(ctx, S, 1)
        H0 = H.sum(axis=(2, 3, 4))
        ref = np.array([np.trace(np.linalg.inv(zz * np.eye(n) - H0)) for zz in z])
        out = {}
        for algo in (0, 1):
            ctx.set_option(L.OPT_RESOLVENT_ALGO, algo)
            got = R.resolvent_sum(z)
            out[algo] = np.abs(got - ref) / np.abs(ref)
        ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)
        print(n, cubic, "rel err algo0", out[0], "algo1", out[1], flush=True)
        R.close(); S.close()
