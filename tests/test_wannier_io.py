"""Wannier90 readers (replaces WannierIO.read_w90_hrdat + the H_R assembly of
aps_example/aps_example.jl:5-21 and the lattice part of ext/WannierIOExt.jl:12-23): a synthetic seedname_hr.dat / .wout pair
written in the on-disk format, and - in the build container, where the reference tree exists - the bundled SrVO3 files
against the committed fixture tests/golden/svo_hr.npz (which is what every other test and the bench use)."""
import os

import numpy as np
import pytest

import autobz_b200 as ab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_hr(path, H, lo, deg):
    n = H.shape[0]
    M = H.shape[2:]
    nr = int(np.prod(M))
    with open(path, "w") as fh:
        fh.write(" written by tests/test_wannier_io.py\n")
        fh.write(f"{n:12d}\n{nr:12d}\n")
        for a in range(0, nr, 15):                         # degeneracies, 15 per line
            fh.write("".join(f"{d:5d}" for d in deg[a:a + 15]) + "\n")
        r = 0
        for i1 in range(M[0]):                             # R3 fastest among R, m fastest within a block
            for i2 in range(M[1]):
                for i3 in range(M[2]):
                    for nn in range(n):
                        for m in range(n):
                            v = H[m, nn, i1, i2, i3] * deg[r]
                            fh.write(f"{i1 + lo[0]:5d}{i2 + lo[1]:5d}{i3 + lo[2]:5d}{m + 1:5d}{nn + 1:5d}{v.real:12.6f}{v.imag:12.6f}\n")
                    r += 1


def test_hrdat_and_wout_roundtrip(tmp_path):
    rng = np.random.default_rng(5)
    n, M, lo = 2, (3, 5, 3), (-1, -2, -1)
    H = np.round(rng.standard_normal((n, n) + M) + 1j * rng.standard_normal((n, n) + M), 3)
    deg = rng.integers(1, 4, size=int(np.prod(M)))
    H = np.round(H * 1.0, 3)
    Hd = H.copy()
    _write_hr(tmp_path / "t_hr.dat", Hd, lo, deg)
    got, glo = ab.read_w90_hrdat(tmp_path / "t_hr.dat")
    assert glo == lo and got.shape == H.shape
    # the file stores H_R * degeneracy rounded to 6 decimals; the reader divides by the degeneracy (aps_example.jl:19-21)
    assert np.max(np.abs(got - H)) < 1e-6
    (tmp_path / "t.wout").write_text("""
                              Lattice Vectors (Ang)
                    a_1     3.858560   0.000000   0.000000
                    a_2     0.100000   3.858560   0.000000
                    a_3     0.000000   0.200000   3.858560

                   Unit Cell Volume:      57.44810  (Ang^3)
""")
    A = ab.read_wout_lattice(tmp_path / "t.wout")
    assert np.allclose(A[:, 0], [3.85856, 0, 0]) and np.allclose(A[:, 1], [0.1, 3.85856, 0]) and np.allclose(A[:, 2], [0, 0.2, 3.85856])
    bz = ab.load_bz(ab.FBZ(), A)
    assert np.allclose(bz.B.T @ A, 2 * np.pi * np.eye(3))          # B = 2 pi A^-T (src/brillouin.jl:9)
    with pytest.raises(ValueError):
        (tmp_path / "bad.wout").write_text("no lattice here\n")
        ab.read_wout_lattice(tmp_path / "bad.wout")


def test_bundled_svo_files_match_the_committed_fixture():
    ref = "/root/reference/aps_example"
    if not os.path.exists(os.path.join(ref, "svo_hr.dat")):
        pytest.skip("reference tree not present (GPU box): the committed fixture is the input there")
    d = np.load(os.path.join(ROOT, "tests", "golden", "svo_hr.npz"))
    H, lo = ab.read_w90_hrdat(os.path.join(ref, "svo_hr.dat"))
    A = ab.read_wout_lattice(os.path.join(ref, "svo.wout"))
    assert H.shape == (3, 3, 11, 11, 11) and lo == (-5, -5, -5)
    assert np.array_equal(H, d["H_R"]) and np.array_equal(np.array(lo), d["lo"]) and np.array_equal(A, d["A"])
    assert np.max(np.abs(H.imag)) == 0.0                                     # svo_hr.dat is real
    assert abs(abs(np.linalg.det(2 * np.pi * np.linalg.inv(A).T)) - 4.31781301953062) < 1e-10   # j = |det B|
