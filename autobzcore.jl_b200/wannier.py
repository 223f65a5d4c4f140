"""Minimal Wannier90 readers (no WannierIO): `seedname_hr.dat` -> H_R tensor, `seedname.wout` -> lattice.
Replaces WannierIO.read_w90_hrdat + the H_R assembly of aps_example/aps_example.jl:5-21 and the lattice
part of ext/WannierIOExt.jl:12-23.  Host-side I/O, outside the timed path."""
import re

import numpy as np


def read_w90_hrdat(path):
    """Returns (H_R, lo): H_R complex [n, n, M1, M2, M3] (entries divided by the degeneracy of R, as
    aps_example.jl:19-21 does), lo = lowest R index per dimension.
    File format: comment line, num_wann, nrpts, degeneracies (15 per line), then `R1 R2 R3 m n Re Im`."""
    with open(path) as fh:
        fh.readline()
        nw = int(fh.readline().split()[0])
        nr = int(fh.readline().split()[0])
        deg = []
        while len(deg) < nr:
            deg += [int(x) for x in fh.readline().split()]
        data = np.loadtxt(fh, dtype=np.float64)
    if data.shape[0] != nr * nw * nw:
        raise ValueError("unexpected number of matrix-element lines in hr.dat")
    R = data[:, :3].astype(np.int64)
    lo = R.min(axis=0)
    hi = R.max(axis=0)
    M = hi - lo + 1
    H = np.zeros((nw, nw) + tuple(M), dtype=np.complex128, order="F")
    blk = np.arange(data.shape[0]) // (nw * nw)
    d = np.asarray(deg, dtype=np.float64)[blk]
    m = data[:, 3].astype(np.int64) - 1
    n = data[:, 4].astype(np.int64) - 1
    H[m, n, R[:, 0] - lo[0], R[:, 1] - lo[1], R[:, 2] - lo[2]] = (data[:, 5] + 1j * data[:, 6]) / d
    return H, tuple(int(x) for x in lo)


def read_wout_lattice(path):
    """Real-space lattice vectors (Angstrom) from a .wout file as the COLUMNS of A."""
    vec = {}
    with open(path) as fh:
        for line in fh:
            m = re.match(r"\s*a_(\d)\s+([-\d.Ee+]+)\s+([-\d.Ee+]+)\s+([-\d.Ee+]+)", line)
            if m:
                vec[int(m.group(1))] = [float(m.group(i)) for i in (2, 3, 4)]
                if len(vec) == 3:
                    break
    if len(vec) != 3:
        raise ValueError("lattice vectors not found in .wout")
    return np.array([vec[1], vec[2], vec[3]]).T
