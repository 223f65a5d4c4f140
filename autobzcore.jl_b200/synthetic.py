"""Synthetic Wannier Hamiltonians and lattice models (inputs of the BASELINE.json configs).

Counter-based generator so that any host language reproduces the same H_R bit for
bit: u(R,a,b,c) = splitmix64(seed xor key(R,a,b,c)) / 2^64.
"""
import numpy as np

DEFAULT_SEED = 20240607
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _uniform(seed, R1, R2, R3, a, b, c):
    key = (((((R1 + 64) * 128 + (R2 + 64)) * 128 + (R3 + 64)) * 256 + a) * 256 + b) * 2 + c
    z = _splitmix64(np.uint64(seed) ^ key.astype(np.uint64))
    return z.astype(np.float64) / 18446744073709551616.0


def wannier_hamiltonian(norb, rmax, seed=DEFAULT_SEED, cubic=False, decay=0.5):
    """H_R for R in [-rmax, rmax]^3 as a complex array [norb, norb, M, M, M] (M = 2 rmax + 1), lo = -rmax.

    H_R = 1/2 exp(-decay |R|_2) (X_R + X_{-R}^dagger), X_R[a,b] = (2u0-1) + i(2u1-1); H_0 += diag(linspace(-1,1)).
    Hermiticity H_{-R} = H_R^dagger is exact.  cubic=True uses one Hermitian block per orbit of R under the 48
    signed permutations (trivial orbital action) so that H(Sk) = H(k) and CubicSymIBZ is valid.
    """
    M = 2 * rmax + 1
    r = np.arange(-rmax, rmax + 1)
    R1, R2, R3 = np.meshgrid(r, r, r, indexing="ij")
    a = np.arange(norb)
    A, B = np.meshgrid(a, a, indexing="ij")

    def X(R1_, R2_, R3_):
        sh = (norb, norb) + R1_.shape
        Rb = [np.broadcast_to(x[None, None], sh) for x in (R1_, R2_, R3_)]
        Ab = np.broadcast_to(A[:, :, None, None, None], sh)
        Bb = np.broadcast_to(B[:, :, None, None, None], sh)
        u0 = _uniform(seed, Rb[0], Rb[1], Rb[2], Ab, Bb, np.zeros(sh, dtype=np.int64))
        u1 = _uniform(seed, Rb[0], Rb[1], Rb[2], Ab, Bb, np.ones(sh, dtype=np.int64))
        return (2 * u0 - 1) + 1j * (2 * u1 - 1)

    if cubic:
        S = np.sort(np.abs(np.stack([R1, R2, R3])), axis=0)[::-1]
        Xc = X(S[0], S[1], S[2])
        H = 0.5 * (Xc + np.conj(np.swapaxes(Xc, 0, 1)))
    else:
        Xp = X(R1, R2, R3)
        Xm = X(-R1, -R2, -R3)
        H = 0.5 * (Xp + np.conj(np.swapaxes(Xm, 0, 1)))
    nrm = np.sqrt((R1 ** 2 + R2 ** 2 + R3 ** 2).astype(np.float64))
    H = H * np.exp(-decay * nrm)[None, None]
    H[a, a, rmax, rmax, rmax] += np.linspace(-1.0, 1.0, norb) if norb > 1 else 0.0
    return np.asfortranarray(H), (-rmax,) * 3


def integer_lattice(ndim):
    """test/utils.jl:3-9: coefficients 1/(2 ndim) at +-e_i => H(k) = (1/ndim) sum_i cos(2 pi k_i).
    Returns (coeffs [1,1,3,...], lo)."""
    c = np.zeros((1, 1) + (3,) * ndim)
    for i in range(ndim):
        for j in (-1, 1):
            idx = [1] * ndim
            idx[i] += j
            c[(0, 0) + tuple(idx)] = 1.0 / (2 * ndim)
    return c, (-1,) * ndim


def band_extent(coeffs):
    """Crude bound on the spectrum: |H(k)| <= sum_R ||H_R||_2 (used to place synthetic frequency sweeps)."""
    c = np.asarray(coeffs)
    n = c.shape[0]
    flat = c.reshape(n, n, -1)
    return float(sum(np.linalg.norm(flat[:, :, i], 2) for i in range(flat.shape[2])))
