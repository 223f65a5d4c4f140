"""The block algebra the resolvent kernels implement, restated in numpy and checked against `numpy.linalg.inv` (CPU only).

These are not tests of the CUDA code (tests/test_gpu_*.py compare that with the oracle on a B200); they pin the FORMULAS the
kernels are built on, so that a reader can check a kernel against a dozen lines of numpy:

* K3-fast / K3-fused / K3-team (csrc/abz_resolvent_mma.cuh): block LU without inter-block pivoting with explicit diagonal-block
  inverses D_s, X = D U, V = (U~)^-1 and M = (L~)^-1 by block substitutions,  tr A^-1 = sum_s tr D_s + sum_{i<j} tr(V_ij M_ji);
  the W / N variant (VAR 2):  tr A^-1 = sum_s tr D_s + sum_{i<j} tr(W_ij D_j N_ji).
* `inv8` / `inv8_ff`: in-place Gauss-Jordan without row exchanges; the division-free form with one scale per row, an exact
  power-of-two rescaling after four steps and one reciprocal per row at the end.
* `small_resolvent_trace<n>` for 4 <= n <= 6 (csrc/abz_kernels.cuh): in-place Gauss-Jordan with partial pivoting, the row
  permutation undone on the columns at the end.

The matrices are of the kind the path produces: A = z - H - Sigma with Hermitian H and Im z > 0 (positive-definite
anti-Hermitian part, hence nonsingular leading blocks)."""
import numpy as np
import pytest


def resolvent_matrix(n, rng, eta=0.05):
    H = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    H = H + H.conj().T
    return (0.3 + 1j * eta) * np.eye(n) - H


def gj_inplace_nopivot(a):
    """inv8: for every pivot p, row g gets row_g - (a_gp / a_pp) row_p (row p: scaled), the pivot column is overwritten."""
    a = a.copy()
    n = a.shape[0]
    for p in range(n):
        piv = a[p, p]
        m = a[:, p] / piv
        m[p] = 1.0 - 1.0 / piv
        row = a[p, :].copy()
        a -= np.outer(m, row)
        col = -m
        col[p] = 1.0 / piv
        a[:, p] = col
    return a


def gj_division_free(a):
    """inv8_ff: row_g <- a_pp row_g - a_gp row_p (g != p); z_g = the row's scale until its own pivot step, its diagonal entry
    afterwards; column p becomes the new column of R (R_pp = scale of row p); rows rescaled by 2^-exponent(z) after four steps."""
    a = a.copy()
    n = a.shape[0]
    z = np.ones(n, dtype=complex)
    for p in range(n):
        pp = a[p, p]
        f = a[:, p].copy()
        a[:, p] = 0.0
        a[p, p] = z[p]
        prow = a[p, :].copy()
        f[p] = 0.0                                  # the pivot row is only scaled
        w = z.copy()
        w[p] = pp
        a = pp * a - np.outer(f, prow)
        z = pp * w
        if p == 3:
            r = 2.0 ** -np.floor(np.log2(np.maximum(np.abs(z.real), np.abs(z.imag))))
            a *= r[:, None]
            z *= r
    return a / z[:, None]


def gj_partial_pivoting(a):
    """small_resolvent_trace<n>, n = 4..6: physical row exchanges, in-place elimination, columns exchanged back in reverse order."""
    a = a.copy()
    n = a.shape[0]
    perm = []
    for p in range(n):
        r = p + int(np.argmax(np.abs(a[p:, p]) ** 2))
        perm.append(r)
        a[[p, r], :] = a[[r, p], :]
        inv = 1.0 / a[p, p]
        a[p, p] = 1.0
        a[p, :] *= inv
        for i in range(n):
            if i != p:
                f = a[i, p]
                a[i, p] = 0.0
                a[i, :] -= f * a[p, :]
    for p in range(n - 1, -1, -1):
        r = perm[p]
        a[:, [p, r]] = a[:, [r, p]]
    return a


def block_lu_trace(A, b=8, variant=1):
    """tr A^-1 the way K3-fast computes it (variant 1) and in the W / N form of VAR 2 (variant 2)."""
    n = A.shape[0]
    nb = n // b
    B = {(i, j): A[i * b:(i + 1) * b, j * b:(j + 1) * b].copy() for i in range(nb) for j in range(nb)}
    D = {}
    for s in range(nb):
        D[s] = gj_inplace_nopivot(B[s, s])                        # D_s = S_s^-1
        for i in range(s + 1, nb):
            B[i, s] = B[i, s] @ D[s]                              # L_is
        for j in range(s + 1, nb):
            U = B[s, j]
            for i in range(s + 1, nb):
                B[i, j] = B[i, j] - B[i, s] @ U                   # trailing update
            B[s, j] = D[s] @ U                                    # X_sj
    tr = sum(np.trace(D[s]) for s in range(nb))
    if variant == 1:
        V = {}
        for jv in range(nb - 1, 0, -1):                           # V = (U~)^-1, U~ = diag(S)(I + X~): right-looking, columns right to left
            for iv in range(jv):
                V[iv, jv] = -B[iv, jv] @ D[jv]
            for t in range(jv - 1, 0, -1):
                for iv in range(t):
                    V[iv, jv] = V[iv, jv] - B[iv, t] @ V[t, jv]
        M = {(i, j): -B[i, j] for j in range(nb) for i in range(j + 1, nb)}
        for jm in range(nb - 1):                                  # M = (L~)^-1, columns left to right
            for t in range(jm + 1, nb):
                tr += np.trace(V[jm, t] @ M[t, jm])
                for im in range(t + 1, nb):
                    M[im, jm] = M[im, jm] - B[im, t] @ M[t, jm]
        return tr
    W = {(i, j): B[i, j].copy() for i in range(nb) for j in range(i + 1, nb)}      # W = -(I + X~)^-1 + I, rows top down
    for i in range(nb):
        for j in range(i + 1, nb):
            for t in range(i + 1, j):
                W[i, j] = W[i, j] - W[i, t] @ B[t, j]
    N = {(i, j): B[i, j].copy() for j in range(nb) for i in range(j + 1, nb)}      # N = -(L~)^-1 + I, columns left to right
    for j in range(nb):
        for i in range(j + 1, nb):
            for t in range(j + 1, i):
                N[i, j] = N[i, j] - B[i, t] @ N[t, j]
    for i in range(nb):
        for j in range(i + 1, nb):
            tr += np.trace(W[i, j] @ D[j] @ N[j, i])
    return tr


@pytest.mark.parametrize("seed", range(5))
def test_diagonal_block_inversions(seed):
    rng = np.random.default_rng(seed)
    A = resolvent_matrix(8, rng)
    ref = np.linalg.inv(A)
    for f in (gj_inplace_nopivot, gj_division_free, gj_partial_pivoting):
        assert np.max(np.abs(f(A) - ref)) < 1e-12 * np.max(np.abs(ref)), f.__name__


@pytest.mark.parametrize("n", [4, 5, 6])
def test_pivoted_elimination_exchanges_rows(n):
    rng = np.random.default_rng(n)
    for trial in range(50):
        A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        A[0, 0] = 0.0                                             # forces an exchange in the first step
        if trial % 2:
            A[1, 1] = 0.0
        ref = np.linalg.inv(A)
        assert np.max(np.abs(gj_partial_pivoting(A) - ref)) < 1e-10 * np.max(np.abs(ref))


@pytest.mark.parametrize("n", [16, 24, 32, 64])
@pytest.mark.parametrize("variant", [1, 2])
def test_trace_of_the_inverse_from_the_block_factors(n, variant):
    rng = np.random.default_rng(100 + n)
    for _ in range(3):
        A = resolvent_matrix(n, rng)
        ref = np.trace(np.linalg.inv(A))
        got = block_lu_trace(A, 8, variant)
        # unpivoted between blocks: the error grows with the conditioning of the leading blocks (eta / |A| ~ 3e-3 here) - 1e-10 for
        # n > 32 as in the GPU tests, where the growth monitor hands harder cases to the pivoted kernel
        assert abs(got - ref) < (1e-11 if n <= 32 else 1e-10) * abs(ref), (n, variant, got, ref)
