import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np
import autobz_b200 as ab
from autobz_b200 import _lib as L
ctx = ab.default_context(0)
d = np.load('/root/repo/tests/golden/svo_hr.npz')
Hs, los = np.asfortranarray(d["H_R"]), tuple(int(x) for x in d["lo"])
S = L.DeviceSeries(ctx, Hs, los, (1.0,) * 3)
syms = np.array(ab.cube_automorphisms(3), dtype=np.int32)
for npt in (400, 831, 831, 831):
    l0 = ctx.launch_count
    t = time.perf_counter(); R = L.DeviceRule(ctx, S, npt, syms=syms); dt = time.perf_counter() - t
    print(npt, "create_symptr ms", round(1e3 * dt, 2), "nodes", len(R), "launches", ctx.launch_count - l0, flush=True)
    R.close()
