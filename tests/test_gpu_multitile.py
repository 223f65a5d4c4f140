"""Oracle-tight parity of the FULL-GRID interface at grid sizes where stage 1 of the contraction runs several
phase tiles per (k2,k3) row: N > 32 makes `contract_stage_kernel` prefetch phase tiles by TMA bulk copies into its two
buffers and flip the mbarrier parity (csrc/abz_kernels.cuh), which the N <= 16 cases of test_gpu_parity.py never reach.
This is the code path of the headline workload (BASELINE config 4: norb 32, M = 17, N = 256, full grid,
`DeviceRule(ctx, S, 256, k3_lo, k3_hi)`), compared here with the CPU oracle (src/fourier.jl:132-164 restated in
oracle/autobz_oracle.c) - not with another CUDA path.

Tolerances: H(k) <= 1e-13 relative (phase rounding only), rule sums <= 1e-11 relative."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import autobz_b200 as ab
from autobz_b200 import _lib as L

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def oracle_plane_sum(orc, So, N, z, k3_lo, k3_hi, sigma=None):
    """orc.ptr_sum over planes [k3_lo, k3_hi) with the k2 rows dealt to host threads (the oracle itself threads over k3
    only, like the reference, src/fourier.jl:156; ctypes releases the GIL)."""
    nt = max(1, min(os.cpu_count() or 1, N))
    bounds = [(i * N // nt, (i + 1) * N // nt) for i in range(nt)]
    with ThreadPoolExecutor(nt) as ex:
        parts = list(ex.map(lambda ab_: orc.ptr_sum(So, N, z, sigma=sigma, k3_lo=k3_lo, k3_hi=k3_hi, k2_lo=ab_[0], k2_hi=ab_[1], nthreads=1),
                            [b for b in bounds if b[1] > b[0]]))
    return sum(parts)


def test_c4_plane_full_grid_interface_vs_oracle(ctx, orc):
    """The headline shape itself: norb 32, R in [-8,8]^3 (M = 17), N = 256, one whole k3 plane (65 536 nodes, 8 phase tiles
    per row) through the full-grid interface, 4 frequencies, against the oracle's LU (<= 1e-11)."""
    n, rmax, N = 32, 8, 256
    H, lo = ab.synthetic.wannier_hamiltonian(n, rmax)
    ext = ab.synthetic.band_extent(H)
    z = np.array([-0.2 * ext, -0.05 * ext, 0.07 * ext, 0.22 * ext]) + 1j * 5e-3 * ext
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    R = L.DeviceRule(ctx, S, N, k3_lo=5, k3_hi=6)
    assert len(R) == N * N
    got = R.resolvent_sum(z, scale=1.0 / N ** 3)
    ref = oracle_plane_sum(orc, orc.Series(H, lo), N, z, 5, 6)
    assert rel(got, ref) < 1e-11, (got, ref)
    R.close(); S.close()


def test_c4_plane_h_of_k_m17_vs_oracle(ctx, orc):
    """M = 17, N = 256, one k3 plane at norb 4: every H(k) of the plane (16 MB) against the oracle's nested contraction
    (<= 1e-13), plus the sum with a matrix self-energy."""
    n, rmax, N = 4, 8, 256
    H, lo = ab.synthetic.wannier_hamiltonian(n, rmax)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    So = orc.Series(H, lo)
    R = L.DeviceRule(ctx, S, N, k3_lo=5, k3_hi=6)
    Hk, k, w = R.copy_out()
    ref = orc.grid_eval_full(So, N, k3_lo=5, k3_hi=6)
    assert rel(Hk.reshape(n, n, N, N, 1, order="F"), ref) < 1e-13
    assert np.all(w == 1.0) and abs(k[N + 3, 0] - 3 / N) < 1e-16 and abs(k[N + 3, 1] - 1 / N) < 1e-16 and abs(k[0, 2] - 5 / N) < 1e-16
    rng = np.random.default_rng(3)
    z = np.array([0.3 + 0.05j, -0.7 + 0.2j])
    sig = 0.1 * (rng.standard_normal((n, n, 2)) + 1j * rng.standard_normal((n, n, 2))) - 0.1j * np.eye(n)[:, :, None]
    assert rel(R.resolvent_sum(z, sigma=sig, scale=1.0 / N ** 3), oracle_plane_sum(orc, So, N, z, 5, 6, sigma=sig)) < 1e-11
    R.close(); S.close()


@pytest.mark.parametrize("N", [33, 65, 97, 129])
@pytest.mark.parametrize("n,rmax", [(4, 8), (32, 2)])
def test_multi_tile_grids_vs_oracle(ctx, orc, n, rmax, N):
    """ntiles = 2 ... 5 phase tiles per row (the last one partial), so that both phase buffers are refilled and the mbarrier
    parity wraps: H(k) on two planes (<= 1e-13) and the plane sums (<= 1e-11) through the full-grid interface."""
    H, lo = ab.synthetic.wannier_hamiltonian(n, rmax)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    So = orc.Series(H, lo)
    k3 = (N // 3, N // 3 + 2)
    R = L.DeviceRule(ctx, S, N, k3_lo=k3[0], k3_hi=k3[1])
    Hk, _, _ = R.copy_out()
    ref = orc.grid_eval_full(So, N, k3_lo=k3[0], k3_hi=k3[1])
    assert rel(Hk.reshape(n, n, N, N, 2, order="F"), ref) < 1e-13
    ext = ab.synthetic.band_extent(H)
    z = np.array([0.1 * ext + 0.01j * ext, -0.3 * ext + 0.05j * ext, 0.25 * ext + 0.002j * ext])
    assert rel(R.resolvent_sum(z, scale=1.0 / N ** 3), oracle_plane_sum(orc, So, N, z, k3[0], k3[1])) < 1e-11
    # the same planes from a materialised rule (cached H(k) reused across parameters, src/interfaces.jl:234-243)
    R.materialize()
    assert rel(R.resolvent_sum(z, scale=1.0 / N ** 3), oracle_plane_sum(orc, So, N, z, k3[0], k3[1])) < 1e-11
    R.close(); S.close()


@pytest.mark.parametrize("n,npt", [(3, 72), (5, 80)])
def test_symmetric_rule_rows_longer_than_one_tile(ctx, orc, n, npt):
    """Symmetry-reduced rule whose (k2,k3) rows hold more than 32 nodes: the gathered (CSR node list) phase tiles of stage 1
    also run several tiles per row.  Sum against the oracle's symmetric sum on the oracle's own symptr weights."""
    H, lo = ab.synthetic.wannier_hamiltonian(n, 2, cubic=True)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    So = orc.Series(H, lo)
    syms = np.array(ab.cube_automorphisms(3), dtype=np.int32)
    w, nirr = orc.symptr_rule(npt, syms)
    assert int((w != 0).sum(axis=0).max()) > 32          # rows longer than one 32-node tile
    ext = ab.synthetic.band_extent(H)
    z = np.array([0.1 * ext + 0.02j * ext, -0.4 * ext + 0.05j * ext])
    ref, cnt = orc.symptr_sum(So, npt, w, z, scale=1.0 / npt ** 3)
    for R in (L.DeviceRule(ctx, S, npt, wsym=w), L.DeviceRule(ctx, S, npt, syms=syms)):
        assert len(R) == nirr == cnt
        assert rel(R.resolvent_sum(z, scale=1.0 / npt ** 3), ref) < 1e-11
        if n > 3:
            Hk, kk, ww = R.copy_out()
            assert rel(Hk, orc.eval_points(So, kk)) < 1e-13
        R.close()
    S.close()
