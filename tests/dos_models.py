"""Tight-binding models and exact densities of states of the reference's DOS tests (test/dos.jl:8-86), restated."""
import numpy as np
from scipy.integrate import quad
from scipy.special import ellipk

import autobz_b200 as ab


def tb_integer(n, t=1.0):
    """test/dos.jl:34-41: H(k) = 2 t sum_i cos(2 pi k_i), 1 x 1 matrices on the offsets -1:1"""
    c = np.zeros((1, 1) + (3,) * n)
    for i in range(n):
        for j in (0, 2):
            idx = [0, 0] + [1] * n
            idx[2 + i] = j
            c[tuple(idx)] = t
    return ab.FourierSeries(c, period=1.0, lo=(-1,) * n, norb=1)


def tb_graphene(t=1.0):
    """test/dos.jl:8-15: 2 x 2 matrices on the offsets -2:2 x -2:2"""
    c = np.zeros((2, 2, 5, 5))
    o = 2
    for (r1, r2) in ((1, 1), (1, -2), (-2, 1)):
        c[0, 1, r1 + o, r2 + o] = t
    for (r1, r2) in ((-1, -1), (-1, 2), (2, -1)):
        c[1, 0, r1 + o, r2 + o] = t
    return ab.FourierSeries(c, period=1.0, lo=(-2, -2), norb=2)


def dos_graphene_exact(E, t=1.0):
    E = abs(E)
    x = abs(E / t)
    if x == 0:
        return 0.0
    if x <= 1:
        f = (1 + x) ** 2 - (x * x - 1) ** 2 / 4
        return 2 * E / ((np.pi * t) ** 2 * np.sqrt(f)) * ellipk(4 * x / f)
    if 1 < x < 3:
        f = (1 + x) ** 2 - (x * x - 1) ** 2 / 4
        return 2 * E / ((np.pi * t) ** 2 * np.sqrt(4 * x)) * ellipk(f / (4 * x))
    return 0.0


def dos_integer_1d_exact(E, t=1.0):
    x = abs(E / (2 * t))
    return 1 / np.sqrt(1 - x * x) / (np.pi * 2 * t) if x <= 1 else 0.0


def dos_integer_2d_exact(E, t=1.0):
    x = abs(E / (4 * t))
    return 1 / (np.pi ** 2 * 2 * t) * ellipk(1 - x * x) if x <= 1 else 0.0


def dos_integer_3d_exact(E, t=1.0):
    x = abs(E / (6 * t))
    f = lambda u: ellipk(1 - ((3 * x - np.cos(u)) / 2) ** 2)
    if 3 * x < 1:
        up = np.arccos(3 * x)
        return (quad(f, 0, up, limit=200)[0] + quad(f, up, np.pi, limit=200)[0]) / (np.pi ** 3 * 2 * t)
    if x < 1:
        return quad(f, 0, np.arccos(3 * x - 2), limit=200)[0] / (np.pi ** 3 * 2 * t)
    return 0.0


# (model, exact DOS, bandwidth, BZ kind) exactly as the loop of test/dos.jl:88-99
CASES = [
    ("graphene", tb_graphene, dos_graphene_exact, 4, "FBZ"),
    ("integer_1d", lambda: tb_integer(1), dos_integer_1d_exact, 2, "FBZ"),
    ("integer_2d", lambda: tb_integer(2), dos_integer_2d_exact, 4, "FBZ"),
    ("integer_3d", lambda: tb_integer(3), dos_integer_3d_exact, 6, "FBZ"),
    ("integer_1d", lambda: tb_integer(1), dos_integer_1d_exact, 2, "InversionSymIBZ"),
    ("integer_2d", lambda: tb_integer(2), dos_integer_2d_exact, 4, "InversionSymIBZ"),
    ("integer_3d", lambda: tb_integer(3), dos_integer_3d_exact, 6, "InversionSymIBZ"),
    ("integer_1d", lambda: tb_integer(1), dos_integer_1d_exact, 2, "CubicSymIBZ"),
    ("integer_2d", lambda: tb_integer(2), dos_integer_2d_exact, 4, "CubicSymIBZ"),
    ("integer_3d", lambda: tb_integer(3), dos_integer_3d_exact, 6, "CubicSymIBZ"),
]


def energies(B):
    return np.array([-B - 1, -0.8 * B, -0.6 * B, -0.2 * B, 0.1 * B, 0.3 * B, 0.5 * B, 0.7 * B, 0.9 * B, B + 2.0])


def bz_of(kind, ndim):
    return ab.load_bz({"FBZ": ab.FBZ, "InversionSymIBZ": ab.InversionSymIBZ, "CubicSymIBZ": ab.CubicSymIBZ}[kind](), np.eye(ndim))
