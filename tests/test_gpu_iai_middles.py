"""Device-side MIDDLE integrals of IAI (iai_mid_kernel: one CTA runs a whole level-1 adaptive integral, its warps the innermost
ones): identical numevals and integrals to the oracle's sequential recursion and to the engine with host-driven middle levels,
with one device round per refinement of the OUTERMOST integral only (src/fourier.jl:432-510)."""
import numpy as np
import pytest

import autobz_b200 as ab

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lims", ["tetra", "cubic"])
@pytest.mark.parametrize("n", [1, 2, 3])
def test_device_middles_match_oracle_and_host_driven_levels(ctx, orc, svo, n, lims):
    if n == 3:
        H, lo, A = svo
        z = complex(12.5, 0.05)
    else:
        H, lo = ab.synthetic.wannier_hamiltonian(n, 2, cubic=True)
        A = np.eye(3)
        z = complex(0.3, 0.05)
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=n)
    S = orc.Series(H, lo)
    f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=z.imag)
    bz = ab.load_bz(ab.CubicSymIBZ() if lims == "tetra" else ab.InversionSymIBZ(), A)
    mult = abs(np.linalg.det(bz.B)) * bz.nsyms
    atol = 2e-3
    if lims == "tetra":
        Io, Eo, neo = orc.iai(S, 3, 1, [0.5] * 3, vkind=0, z=z, atol=atol)
    else:
        Io, Eo, neo = orc.iai(S, 3, 0, [0.0] * 3, [0.5] * 3, vkind=0, z=z, atol=atol)
    res = {}
    for mids in (True, False):
        be = ab.DeviceBackend(ctx=ctx, iai_engine="native", iai_device_leaves=True, iai_device_middles=mids)
        cache = ab.init(ab.IntegralProblem(f, bz, {"omega": z.real}), ab.EvalCounter(ab.IAI()), abstol=atol * mult, backend=be)
        sol = ab.solve_(cache)
        res[mids] = (sol, cache.cacheval["iai_rounds"])
        assert sol.numevals == neo
        assert abs(sol.u - mult * Io) <= 1e-10 * abs(sol.u)
    assert res[True][0].u == res[False][0].u and res[True][0].resid == res[False][0].resid     # bit-identical
    assert res[True][1] < res[False][1]                                                        # rounds: outermost refinements only


def test_device_middles_maxevals_rtol_and_errors(ctx, orc, svo):
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    S = orc.Series(H, lo)
    ibz = ab.load_bz(ab.CubicSymIBZ(), A)
    mult = abs(np.linalg.det(ibz.B)) * 48
    f = ab.FourierIntegrand(ab.dos_integrand, fs, 0.02)
    # relative tolerance only
    Io, Eo, neo = orc.iai(S, 3, 1, [0.5] * 3, vkind=1, z=complex(12.2, 0.02), atol=0.0, rtol=1e-2)
    sol = ab.solve(ab.IntegralProblem(f, ibz, 12.2), ab.EvalCounter(ab.IAI()), reltol=1e-2)
    assert sol.numevals == neo and abs(sol.u - mult * Io.real) <= 1e-10 * abs(sol.u)
    # maxiters cuts every 1-D integral after its first refinement
    Io, Eo, neo = orc.iai(S, 3, 1, [0.5] * 3, vkind=1, z=complex(12.2, 0.02), atol=1e-12, maxevals=45)
    sol = ab.solve(ab.IntegralProblem(f, ibz, 12.2), ab.EvalCounter(ab.IAI()), abstol=1e-12 * mult, maxiters=45)
    assert sol.numevals == neo and abs(sol.u - mult * Io.real) <= 1e-10 * abs(sol.u)
    # a pole on the integration path: NaN/Inf surfaces as an error (QuadGK's DomainError), no hang
    g = ab.FourierIntegrand(ab.gloc_trace_integrand, ab.FourierSeries(np.zeros((1, 1, 1, 1, 1)), period=1.0, lo=(0, 0, 0), norb=1), eta=0.0)
    with pytest.raises(FloatingPointError):
        ab.solve(ab.IntegralProblem(g, ab.load_bz(ab.CubicSymIBZ(), np.eye(3)), {"omega": 0.0}), ab.IAI(), abstol=1e-3)


@pytest.mark.parametrize("lims", ["tetra", "cubic"])
def test_lookahead_on_the_outermost_integral(ctx, orc, svo, lims):
    """ABZ_IAI_SPECULATE: two bisections of the outermost integral per device round.  Bit-identical value and error estimate, identical
    numevals (oracle's sequential recursion), fewer rounds; 2-d solves (innermost integrals as tasks) take part too."""
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    S = orc.Series(H, lo)
    z = complex(12.5, 0.005)
    f = ab.FourierIntegrand(ab.dos_integrand, fs, z.imag)
    bz = ab.load_bz(ab.CubicSymIBZ() if lims == "tetra" else ab.InversionSymIBZ(), A)
    mult = abs(np.linalg.det(bz.B)) * bz.nsyms
    atol = 2e-5 if lims == "tetra" else 1e-4
    if lims == "tetra":
        Io, Eo, neo = orc.iai(S, 3, 1, [0.5] * 3, vkind=1, z=z, atol=atol)
    else:
        Io, Eo, neo = orc.iai(S, 3, 0, [0.0] * 3, [0.5] * 3, vkind=1, z=z, atol=atol)
    res = {}
    for spec in (False, True):
        be = ab.DeviceBackend(ctx=ctx, iai_speculate=spec)
        cache = ab.init(ab.IntegralProblem(f, bz, z.real), ab.EvalCounter(ab.IAI()), abstol=atol * mult, backend=be)
        sol = ab.solve_(cache)
        res[spec] = (sol, cache.cacheval["iai_rounds"])
        assert sol.numevals == neo and abs(sol.u - mult * Io.real) <= 1e-10 * abs(sol.u)
    assert res[True][0].u == res[False][0].u and res[True][0].resid == res[False][0].resid
    assert res[True][1] < res[False][1]
    # 2-d
    c2 = np.zeros((1, 1, 3, 3)); c2[0, 0, 0, 1] = c2[0, 0, 2, 1] = c2[0, 0, 1, 0] = c2[0, 0, 1, 2] = 0.5
    h2 = ab.FourierSeries(c2, period=1.0, lo=(-1, -1), norb=1)
    bz2 = ab.load_bz(ab.InversionSymIBZ(2), np.eye(2))
    g = ab.FourierIntegrand(ab.gloc_trace_integrand, h2, eta=0.005)
    out = []
    for spec in (False, True):
        cache = ab.init(ab.IntegralProblem(g, bz2, {"omega": 0.3}), ab.EvalCounter(ab.IAI()), abstol=1e-6, backend=ab.DeviceBackend(ctx=ctx, iai_speculate=spec))
        out.append((ab.solve_(cache), cache.cacheval["iai_rounds"]))
    assert out[0][0].u == out[1][0].u and out[0][0].numevals == out[1][0].numevals and out[1][1] < out[0][1]
