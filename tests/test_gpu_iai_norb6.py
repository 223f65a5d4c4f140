"""Device-side IAI integrals for 4 <= norb <= 6 (round 2): the per-thread innermost evaluation (`nest_point_small`) inverts
z - H(k) - Sigma by an in-register Gauss-Jordan elimination with partial pivoting (csrc/abz_kernels.cuh,
`small_resolvent_trace<NORB>`), so the leaf kernel (one warp per innermost adaptive integral), the middle-integral clusters and
the look-ahead of the outermost integral serve models with up to six orbitals - the reference's nest is value-type generic
(src/fourier.jl:432-510) and `inv` of an SMatrix larger than 3 x 3 is a pivoted LU as well.

Every case is compared with the CPU oracle's sequential recursion (oracle/autobz_oracle.c): identical `numevals`, integrals
<= 1e-10 relative; the engine configurations (device leaves / middles / look-ahead on and off) agree with each other."""
import numpy as np
import pytest

import autobz_b200 as ab

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [4, 5, 6])
def test_three_dimensional_iai_vs_oracle(ctx, orc, n):
    """3-d DOS integrand on the cubic IBZ (TetrahedralLimits) and the complex trace on the full BZ."""
    H, lo = ab.synthetic.wannier_hamiltonian(n, 2, cubic=True)
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=n)
    S = orc.Series(H, lo)
    ibz = ab.load_bz(ab.CubicSymIBZ(), np.eye(3))
    fbz = ab.load_bz(ab.FBZ(), np.eye(3))
    mult = (2 * np.pi) ** 3
    eta, omega, atol = 0.25, 0.3, 2e-3
    Io, Eo, neo = orc.iai(S, 3, 1, [0.5] * 3, vkind=1, z=complex(omega, eta), atol=atol)
    f = ab.FourierIntegrand(ab.dos_integrand, fs, eta)
    res = []
    for leaves, middles, spec in ((True, True, True), (True, True, False), (True, False, False), (False, False, False)):
        be = ab.DeviceBackend(ctx=ctx, iai_engine="native", iai_device_leaves=leaves, iai_device_middles=middles, iai_speculate=spec)
        sol = ab.solve(ab.IntegralProblem(f, ibz, omega), ab.EvalCounter(ab.IAI()), abstol=atol * mult * 48, backend=be)
        assert sol.numevals == neo, (n, leaves, middles, spec, sol.numevals, neo)
        assert abs(sol.u - mult * 48 * Io.real) <= 1e-10 * abs(sol.u)
        res.append(sol.u)
    assert res[0] == res[1] == res[2]                      # the device-task configurations take bit-identical decisions
    Io, Eo, neo = orc.iai(S, 3, 0, [0.0] * 3, [1.0] * 3, vkind=0, z=complex(0.1, 0.4), atol=0.0, rtol=1e-3)
    g = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=0.4)
    sol = ab.solve(ab.IntegralProblem(g, fbz, {"omega": 0.1}), ab.EvalCounter(ab.IAI()), reltol=1e-3)
    assert sol.numevals == neo
    assert abs(sol.u - mult * Io) <= 1e-10 * abs(sol.u)


@pytest.mark.parametrize("n", [4, 6])
def test_two_dimensional_iai_with_self_energy_and_pivoting(ctx, orc, n):
    """2-d solve (device leaves only) with a matrix self-energy; and a Hamiltonian whose leading diagonal entries vanish, so that
    the elimination has to exchange rows (z ~ 0)."""
    H, lo = ab.synthetic.wannier_hamiltonian(n, 2)
    H2, lo2 = np.asfortranarray(H[:, :, :, :, 2]), lo[:2]
    m0 = tuple(-l for l in lo2)
    H2[0, 0, m0[0], m0[1]] = 0.0          # on-site (0,0) entry
    H2[1, 1, m0[0], m0[1]] = 0.0
    fs = ab.FourierSeries(H2, period=1.0, lo=lo2, norb=n)
    So = orc.Series(H2[..., None], tuple(lo2) + (0,))
    bz2 = ab.load_bz(ab.FBZ(2), 2 * np.pi * np.eye(2))
    rng = np.random.default_rng(n)
    sig = 0.05 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) - 0.2j * np.eye(n)
    for z, sigma in ((complex(0.0, 0.15), None), (complex(0.2, 0.05), sig)):
        Io, Eo, neo = orc.iai(So, 2, 0, [0.0] * 2, [1.0] * 2, vkind=0, z=z, sigma=sigma, atol=1e-4)
        f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=z.imag, sigma=sigma) if sigma is not None else \
            ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=z.imag)
        for leaves in (True, False):
            be = ab.DeviceBackend(ctx=ctx, iai_engine="native", iai_device_leaves=leaves)
            sol = ab.solve(ab.IntegralProblem(f, bz2, {"omega": z.real}), ab.EvalCounter(ab.IAI()), abstol=1e-4, backend=be)
            assert sol.numevals == neo, (n, z, leaves, sol.numevals, neo)
            assert abs(sol.u - Io) <= 1e-10 * abs(Io)
