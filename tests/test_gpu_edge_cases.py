"""Edge cases of the C ABI on the device: empty shards, one-point grids, constant series, more frequencies than one CTA pass
holds, frequency lists that do not divide the chunk sizes — each against the oracle or an exact value."""
import numpy as np
import pytest

import autobz_b200 as ab
from autobz_b200 import _lib as L

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


@pytest.mark.parametrize("n", [1, 3, 8, 32, 64])
def test_empty_shards_and_one_point_grids(ctx, orc, n):
    H, lo = ab.synthetic.wannier_hamiltonian(n, 1, cubic=True)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    z = np.array([0.2 + 0.3j, 0.5 + 0.1j])
    syms = np.array(ab.cube_automorphisms(3), dtype=np.int32)
    # a rank whose share of the k3 planes is empty (more ranks than planes)
    for R in (L.DeviceRule(ctx, S, 4, k3_lo=4, k3_hi=4), L.DeviceRule(ctx, S, 4, syms=syms, k3_lo=5, k3_stride=7)):
        assert len(R) == 0
        assert np.all(R.resolvent_sum(z) == 0) and R.eig_sum(L.EIG_SUM) == 0.0 and np.all(R.resolvent_matrix_sum(z) == 0)
        assert R.eigvals().shape == (0, n)
        R.ggr_data(3, copy=False)
        assert np.all(R.ggr_sum([0.0, 1.0]) == 0)
        Hk, k, w = R.copy_out()
        assert Hk.shape == (n, n, 0)
        R.materialize()
    # one-point grid: H(k = 0) = sum_R H_R
    R = L.DeviceRule(ctx, S, 1)
    H0 = H.sum(axis=(2, 3, 4))
    # one matrix, no averaging over k: the unpivoted block eliminations (DMMA kernels) keep ~2 digits less than pivoted LU
    assert rel(R.resolvent_sum(z), [np.trace(np.linalg.inv(zz * np.eye(n) - H0)) for zz in z]) < (1e-12 if n <= 32 else 2e-11)
    assert rel(R.eigvals()[0], np.linalg.eigvalsh((H0 + H0.conj().T) / 2)) < 1e-13
    Rs = L.DeviceRule(ctx, S, 1, syms=syms)
    assert len(Rs) == 1 and Rs.copy_out()[2][0] == 1.0


@pytest.mark.parametrize("n,nw", [(1, 1030), (3, 2500), (5, 300), (32, 257), (64, 130)])
def test_many_frequencies(ctx, orc, n, nw):
    """frequency lists longer than one CTA pass (small-norb kernel: 1024 per pass; DMMA / Gauss-Jordan kernels: shared
    accumulators sized by nw) and not multiples of any chunk size; checked at both ends and in the middle"""
    H, lo = ab.synthetic.wannier_hamiltonian(n, 1)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    So = orc.Series(H, lo)
    N = 5 if n <= 5 else 3
    z = np.linspace(-1.0, 1.0, nw) + 0.2j
    got = L.DeviceRule(ctx, S, N).resolvent_sum(z, scale=1 / N ** 3)
    pick = np.array([0, 1, nw // 2, nw - 2, nw - 1])
    assert got.shape == (nw,) and rel(got[pick], orc.ptr_sum(So, N, z[pick])) < 1e-11
    ctx.set_option(L.OPT_RESOLVENT_ALGO, 3)
    try:
        sw = L.DeviceRule(ctx, S, N).resolvent_sum(z, scale=1 / N ** 3)
    finally:
        ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)
    assert rel(sw, got) < 1e-11


def test_constant_series_and_real_coefficients(ctx, orc):
    """M = 1 in every dimension (a k-independent H) and real-valued coefficient input (is_complex = 0, as svo_hr.dat)"""
    rng = np.random.default_rng(3)
    A = rng.standard_normal((4, 4))
    A = (A + A.T) / 2
    S = L.DeviceSeries(ctx, A.reshape(4, 4, 1, 1, 1), (0, 0, 0), (1.0,) * 3)
    R = L.DeviceRule(ctx, S, 3)
    z = np.array([0.1 + 0.5j])
    assert rel(R.resolvent_sum(z, scale=1 / 27), [np.trace(np.linalg.inv(z[0] * np.eye(4) - A))]) < 1e-12
    assert rel(R.eigvals(), np.tile(np.linalg.eigvalsh(A), (27, 1))) < 1e-13
