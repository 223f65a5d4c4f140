"""Golden values for BASELINE config 3 at its own eta: SrVO3 DOS via IAI on the cubic IBZ, eta = 1e-4, abstol = 1e-3
(physical units, i.e. atol = 1e-3 / (j * 48) for the nested integral over TetrahedralLimits(1/2)), omega = 12.0 and 12.975161.
Computed with the CPU oracle's sequential recursion (oracle/autobz_oracle.c: orc_iai), about 40-60 s of one core each, which is
too slow for the test suite - hence committed here.  The -m gpu test asserts identical `numevals` and I within 1e-10 relative.
Needs only tests/golden/svo_hr.npz (made by make_golden.py from the reference's svo_hr.dat)."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc

d = np.load(os.path.join(HERE, "svo_hr.npz"))
H, lo, A = np.asfortranarray(d["H_R"]), tuple(int(x) for x in d["lo"]), d["A"]
S = orc.Series(H, lo)
B = 2 * np.pi * np.linalg.inv(A).T
j = abs(np.linalg.det(B))
out = {"eta": 1e-4, "abstol_physical": 1e-3, "j": j, "nsyms": 48, "limits": "TetrahedralLimits(0.5, 0.5, 0.5)", "cases": []}
for omega in (12.0, 12.975161):
    t0 = time.time()
    Iv, E, ne = orc.iai(S, 3, 1, [0.5] * 3, vkind=1, z=complex(omega, 1e-4), atol=1e-3 / (j * 48))
    out["cases"].append({"omega": omega, "I": Iv.real, "E": E, "numevals": ne, "oracle_seconds": round(time.time() - t0, 1)})
    print(out["cases"][-1], flush=True)
json.dump(out, open(os.path.join(HERE, "c3_eta1e-4.json"), "w"), indent=1)
