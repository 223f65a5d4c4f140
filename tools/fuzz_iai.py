"""Randomised check of the device IAI engine against the oracle's recursion (identical numevals, values to 1e-10): random
Hermitian series (1-3 orbitals, 1-3 dimensions, 3-5 coefficients per dimension), cubic / tetrahedral limits, real / complex
values, with and without device-side innermost integrals.  Usage: python tools/fuzz_iai.py [cases=100] [seed=0]"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import autobz_b200 as ab
from autobz_b200 import _lib as L
import orc

ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
ctx = ab.default_context(0)
fails, worst, tot, t0 = 0, 0.0, 0, time.time()
for case in range(ncase):
    ndim, n = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    lkind, vkind = int(rng.integers(0, 2)), int(rng.integers(0, 2))
    m = int(rng.choice([3, 5]))
    M = (m,) * ndim
    c = rng.standard_normal((n, n) + M) + 1j * rng.standard_normal((n, n) + M)
    rev = (slice(None), slice(None)) + (slice(None, None, -1),) * ndim
    c = 0.5 * (c + np.conj(np.swapaxes(c[rev], 0, 1))) / m
    lo = (-(m // 2),) * ndim
    fs = ab.FourierSeries(c, period=1.0, lo=lo, norb=n)
    So = orc.Series(c.reshape((n, n) + M + (1,) * (3 - ndim)), lo + (0,) * (3 - ndim))
    z = complex(rng.uniform(-1, 1), float(rng.choice([0.3, 0.1, 0.03])))
    tol = float(rng.choice([3e-1, 3e-2, 3e-3, 3e-4])) * (10.0 if ndim == 3 else 1.0)
    la = [0.5] * ndim if lkind else [0.0] * ndim
    lb = None if lkind else [1.0] * ndim
    tag = f"case {case}: ndim={ndim} n={n} M={m} lkind={lkind} vkind={vkind} z={z} atol={tol}"
    try:
        Io, Eo, neo = orc.iai(So, ndim, lkind, la, lb, vkind=vkind, z=z, atol=tol)
        nest = L.DeviceNest(ctx, fs.device(ctx), ndim, 256 if ndim == 3 else 0, 4096 if ndim >= 2 else 0)
        first = None
        # (device leaves, device middles, look-ahead): every engine configuration must take the oracle's decisions
        for leaves, mids, spec in ((True, True, True), (True, True, False), (True, False, True), (True, False, False), (False, False, False)):
            I, E, ne, rounds, launches = nest.iai_solve(lkind, la, lb, L.F_RESOLVENT_TRACE, vkind, z, None, None, tol, 0.0, 2 ** 62,
                                                        device_leaves=leaves, device_middles=mids, speculate=spec)
            e = abs(I - Io) / max(abs(Io), 1e-12)
            worst = max(worst, e)
            first = (I, E) if first is None else first
            if ne != neo or not e <= 1e-10 or (leaves and (I, E) != first):     # device-task configurations agree bit for bit
                fails += 1
                print("FAIL", tag, "leaves/middles/lookahead", leaves, mids, spec, "numevals", ne, neo, "rel", e, flush=True)
        tot += neo
    except Exception as ex:                                 # noqa: BLE001
        fails += 1
        print("ERROR", tag, type(ex).__name__, str(ex)[:200], flush=True)
print(f"{ncase} IAI cases ({tot} oracle evaluations), {fails} failures, worst relative difference {worst}, {time.time() - t0:.1f} s")
sys.exit(1 if fails else 0)
