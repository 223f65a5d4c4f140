"""K3-fused (`resolvent_mma_fused_kernel`, csrc/abz_resolvent_mma.cuh): the one-warp DMMA resolvent kernel with the innermost
contraction stage folded in (H(k) = sum_m C1[row][m] e^{2 pi i k1 R_m} formed per node in shared memory, never written to HBM).
It serves streamed (not materialised) rules with 4 <= norb <= 32 and at least 8 frequencies - the headline workload's shape.
Every case is compared with the CPU oracle (src/fourier.jl:127-164 + the docs' `tr(inv(...))` integrand restated in
oracle/autobz_oracle.c); a materialised rule (separate stage-1 kernel, then the same kernel in its direct mode: H(k) copied
into the shared-memory buffer) must agree with the streamed one.

Tolerances: rule sums <= 1e-11 relative."""
import numpy as np
import pytest

import autobz_b200 as ab
from autobz_b200 import _lib as L

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def freqs(H, nw, eta=0.02):
    ext = ab.synthetic.band_extent(H)
    return np.linspace(-0.45, 0.45, nw) * ext + 1j * eta * ext


@pytest.mark.parametrize("n,rmax,N,nw", [(4, 2, 9, 8), (7, 2, 10, 11), (8, 1, 12, 16), (12, 2, 8, 9), (17, 1, 9, 8), (24, 1, 7, 13),
                                         (25, 1, 6, 8), (31, 1, 6, 10), (32, 2, 8, 24)])
def test_fused_full_grid_vs_oracle(ctx, orc, n, rmax, N, nw):
    """Full grids, every block count NB = 1..4 with ragged norb (padding rows), frequency counts that are not multiples of the
    8 warps, scalar and matrix self-energies; the same rule materialised (separate stage 1) gives the same sums."""
    H, lo = ab.synthetic.wannier_hamiltonian(n, rmax)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    So = orc.Series(H, lo)
    z = freqs(H, nw)
    R = L.DeviceRule(ctx, S, N)
    l0 = ctx.launch_count
    got = R.resolvent_sum(z, scale=1.0 / N ** 3)
    nl = ctx.launch_count - l0
    ref = orc.ptr_sum(So, N, z)
    assert rel(got, ref) < 1e-11, (n, N, nw, rel(got, ref))
    rng = np.random.default_rng(n + N)
    sig = 0.05 * (rng.standard_normal((n, n, nw)) + 1j * rng.standard_normal((n, n, nw))) - 0.1j * np.eye(n)[:, :, None]
    got_s = R.resolvent_sum(z, sigma=sig, scale=1.0 / N ** 3)
    assert rel(got_s, orc.ptr_sum(So, N, z, sigma=sig)) < 1e-11
    R.materialize()
    l1 = ctx.launch_count
    mat = R.resolvent_sum(z, scale=1.0 / N ** 3)
    assert rel(mat, got) < 1e-12
    assert rel(R.resolvent_sum(z, sigma=sig, scale=1.0 / N ** 3), got_s) < 1e-12
    assert nl >= 1 and ctx.launch_count > l1
    R.close(); S.close()


def test_fused_symmetric_rule_and_slabs_vs_oracle(ctx, orc):
    """Symmetry-reduced (CSR) rules: rows of different lengths, empty rows, weights, the k1 gather list; k3 shards add up."""
    syms = ab.cube_automorphisms(3)
    for n, rmax, N, nw in [(6, 1, 10, 8), (16, 1, 9, 12), (32, 1, 8, 8)]:
        H, lo = ab.synthetic.wannier_hamiltonian(n, rmax, cubic=True)
        S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
        So = orc.Series(H, lo)
        w_d, nirr = ctx.symptr_rule(N, np.array(syms, dtype=np.int32))
        z = freqs(H, nw)
        R = L.DeviceRule(ctx, S, N, wsym=w_d)
        ref, cnt = orc.symptr_sum(So, N, w_d, z, scale=1 / N ** 3)
        assert len(R) == cnt == nirr
        got = R.resolvent_sum(z, scale=1 / N ** 3)
        assert rel(got, ref) < 1e-11, (n, N, rel(got, ref))
        shards = [L.DeviceRule(ctx, S, N, wsym=w_d, k3_lo=r, k3_stride=3).resolvent_sum(z, scale=1 / N ** 3) for r in range(3)]
        assert rel(sum(shards), ref) < 1e-11
        full = [L.DeviceRule(ctx, S, N, k3_lo=a, k3_hi=b).resolvent_sum(z, scale=1 / N ** 3) for a, b in ((0, 3), (3, 4), (4, N))]
        assert rel(sum(full), orc.ptr_sum(So, N, z)) < 1e-11
        S.close()


def test_fused_small_budget_splits_rows(ctx, orc):
    """A workspace budget that cuts planes into row ranges: chunks that start in the middle of a plane."""
    n, rmax, N, nw = 32, 2, 16, 8          # 80 KB of C1 per row: 12 rows per chunk, planes of 16 rows are split
    H, lo = ab.synthetic.wannier_hamiltonian(n, rmax)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    z = freqs(H, nw)
    ref = orc.ptr_sum(orc.Series(H, lo), N, z)
    ctx.set_option(L.OPT_MEM_BUDGET_MB, 1)
    try:
        got = L.DeviceRule(ctx, S, N).resolvent_sum(z, scale=1.0 / N ** 3)
    finally:
        ctx.set_option(L.OPT_MEM_BUDGET_MB, 4096)
    assert rel(got, ref) < 1e-11
    S.close()


def test_fused_path_keeps_the_pivot_monitor(ctx, orc):
    """A matrix whose leading block is nearly singular raises the growth flag inside the fused kernel too; the call is repeated
    with the pivoted kernel and stays at oracle accuracy."""
    n, N, nw = 32, 4, 8
    rng = np.random.default_rng(11)
    H = np.zeros((n, n, 1, 1, 1), dtype=np.complex128)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = A + A.conj().T
    A[0, 0] = 0.0
    H[:, :, 0, 0, 0] = A
    lo = (0, 0, 0)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    z = np.linspace(-1.0, 1.0, nw) * 1e-9 + 1e-12j          # z ~ 0: the (0,0) pivot of z - H is ~ 1e-9
    got = L.DeviceRule(ctx, S, N).resolvent_sum(z, scale=1.0 / N ** 3)
    ref = orc.ptr_sum(orc.Series(H, lo), N, z)
    assert rel(got, ref) < 1e-10
    S.close()


def test_fused_many_frequencies_streamed_and_materialised(ctx, orc):
    """128 frequencies (the bench's count): the kernel's dynamic shared memory (8 x nw accumulators + two H(k) buffers = 50 KB) is above
    the 48 KB that need the per-device opt-in, in the streamed and in the direct (materialised) instantiation."""
    n, rmax, N, nw = 32, 1, 5, 128
    H, lo = ab.synthetic.wannier_hamiltonian(n, rmax)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    z = freqs(H, nw)
    ref = orc.ptr_sum(orc.Series(H, lo), N, z)
    R = L.DeviceRule(ctx, S, N)
    got = R.resolvent_sum(z, scale=1.0 / N ** 3)
    assert rel(got, ref) < 1e-11
    R.materialize()
    assert rel(R.resolvent_sum(z, scale=1.0 / N ** 3), ref) < 1e-11
    R.close(); S.close()
