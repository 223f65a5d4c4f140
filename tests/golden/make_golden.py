"""Generates the golden fixtures under tests/golden/ (run in the build container, where /root/reference exists;
nothing under tests/ reads /root/reference at run time).

  svo_hr.npz      SrVO3 t2g H_R tensor [3,3,11,11,11] (entries / degeneracy), lo, lattice A — parsed from the
                  reference's bundled aps_example/svo_hr.dat + svo.wout with autobz_b200.wannier
  golden.json     known-answer values quoted from the reference's docs/tests (with file:line) and oracle values
                  for the SrVO3 / synthetic configs (regression anchors; the oracle itself is pinned on the former)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import autobz_b200 as ab
import orc

REF = "/root/reference/aps_example"
H, lo = ab.read_w90_hrdat(os.path.join(REF, "svo_hr.dat"))
A = ab.read_wout_lattice(os.path.join(REF, "svo.wout"))
np.savez_compressed(os.path.join(HERE, "svo_hr.npz"), H_R=H, lo=np.array(lo), A=A)

S = orc.Series(H, lo)
g = {
    "reference_known_answers": {
        "docs/src/examples.md:60 (QuadGKJL abstol=1e-3, 1-D gloc, eta=0.1, omega=0)": [-2.7755575615628914e-17, -0.9950375451895513],
        "docs/src/examples.md:105 (IAI abstol=1e-3, 2-D gloc on FBZ(2), eta=0.1, omega=0)": [1.5265566588595902e-16, -1.3941704019631334],
        "test/fourier.jl:40-56 integral of 1.3*H(k)+1 over the BZ of A=I(d)": "(2*pi)**d",
        "test/brillouin.jl:96 EvalCounter(QuadGKJL(order=7)) on a constant": 15,
    },
    "oracle_values": {},
}
zs = [11.0 + 0.01j, 12.0 + 0.01j, 12.975161 + 0.01j, 13.5 + 0.01j]
for N in (24, 50):
    v = orc.ptr_sum(S, N, zs)
    g["oracle_values"][f"svo_fbz_ptr_N{N}_eta1e-2"] = [[float(x.real), float(x.imag)] for x in v]
c1, lo1 = ab.synthetic.integer_lattice(3)
v = orc.ptr_sum(orc.Series(c1[..., :, :, :].astype(complex), lo1), 64, [0.1j, 0.5 + 0.1j])
g["oracle_values"]["c1_ptr_N64_eta0.1"] = [[float(x.real), float(x.imag)] for x in v]
Iv, E, ne = orc.iai(S, 3, 1, [0.5, 0.5, 0.5], vkind=1, z=12.5 + 0.05j, atol=1e-2)
g["oracle_values"]["svo_iai_tetra_dos_w12.5_eta0.05_atol1e-2"] = {"I": Iv.real, "E": E, "numevals": ne}
json.dump(g, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
print(json.dumps(g["oracle_values"], indent=1))
