"""IAI over general iterated limits on the device (abz_iai_solve_general): the OUTPUT of the reference's IBZ loader - a convex
polyhedron's `segments` / `fixandeliminate` (ext/SymmetryReduceBZExt.jl:33-58) - and several initial segments per 1-D integral
(PuncturedInterval(segs), src/fourier.jl:493-500), against the oracle's recursion over the same limits object (identical
numevals, integrals <= 1e-10 relative) and against the built-in limit kinds."""
import numpy as np
import pytest

import autobz_b200 as ab

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("leaves", [True, False])
def test_segmented_and_callback_limits_vs_oracle(ctx, orc, svo, leaves):
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    S = orc.Series(H, lo)
    be = ab.DeviceBackend(ctx=ctx, iai_engine="native", iai_device_leaves=leaves)
    f = ab.FourierIntegrand(ab.dos_integrand, fs, 0.05)
    seg = ab.SegmentedLimits((0.0, 0.2, 0.5), (0.0, 0.25, 0.3, 0.5), (0.0, 0.1, 0.5))
    Io, Eo, neo = orc.iai_general(S, 3, seg, vkind=1, z=complex(12.5, 0.05), atol=2e-3)
    sol = ab.solve(ab.IntegralProblem(f, seg, 12.5), ab.EvalCounter(ab.NestedQuad(ab.AuxQuadGKJL())), abstol=2e-3, backend=be)
    assert sol.numevals == neo and abs(sol.u - Io.real) <= 1e-10 * abs(Io.real)
    # the Python-driven engine takes the same path
    bp = ab.DeviceBackend(ctx=ctx, iai_engine="python")
    sp = ab.solve(ab.IntegralProblem(f, seg, 12.5), ab.EvalCounter(ab.NestedQuad(ab.AuxQuadGKJL())), abstol=2e-3, backend=bp)
    assert sp.numevals == neo and abs(sp.u - Io.real) <= 1e-10 * abs(Io.real)
    # TetrahedralLimits with s != 1 has no built-in kind: it is served through the callback
    tet = ab.TetrahedralLimits([0.5] * 3, s=0.8)
    It, Et, net = orc.iai_general(S, 3, tet, vkind=1, z=complex(12.5, 0.05), atol=1e-3)
    st = ab.solve(ab.IntegralProblem(f, tet, 12.5), ab.EvalCounter(ab.NestedQuad(ab.AuxQuadGKJL())), abstol=1e-3, backend=be)
    assert st.numevals == net and abs(st.u - It.real) <= 1e-10 * abs(It.real)


def test_polyhedral_ibz_limits_on_device(ctx, orc, svo):
    """The cubic IBZ as a hand-built polyhedron (the tetrahedron 0 <= x <= y <= z <= 1/2 of load_bz(CubicSymIBZ),
    src/brillouin.jl:301-307) in a SymmetricBZ with the 48 cube symmetries: same integral as CubicSymIBZ's TetrahedralLimits within
    the requested tolerance, identical numevals to the oracle on the same polyhedron; and volumes of hand-built polyhedra."""
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    S = orc.Series(H, lo)
    ibz = ab.load_bz(ab.CubicSymIBZ(), A)
    verts = np.array([(0, 0, 0), (0, 0, 0.5), (0, 0.5, 0.5), (0.5, 0.5, 0.5)], dtype=float)
    poly = ab.PolyhedronLimits(verts)
    pbz = ab.SymmetricBZ(ibz.A, ibz.B, poly, ibz.syms)
    f = ab.FourierIntegrand(ab.dos_integrand, fs, 0.05)
    mult = abs(np.linalg.det(ibz.B)) * 48
    a = ab.solve(ab.IntegralProblem(f, ibz, 12.5), ab.EvalCounter(ab.IAI()), abstol=1e-4 * mult)
    b = ab.solve(ab.IntegralProblem(f, pbz, 12.5), ab.EvalCounter(ab.IAI()), abstol=1e-4 * mult)
    assert abs(a.u - b.u) <= 3e-4 * mult
    Io, Eo, neo = orc.iai_general(S, 3, poly, vkind=1, z=complex(12.5, 0.05), atol=1e-4)
    assert b.numevals == neo and abs(b.u - mult * Io.real) <= 1e-10 * abs(b.u)
    # nested_quad(1, lims) = volume (test/test_ibz.jl:121-149), through the affine integrand 0 * tr H + 1
    one = ab.FourierIntegrand(ab.AffineTraceIntegrand(), ab.FourierSeries(np.zeros((1, 1, 1, 1, 1)), period=1.0, lo=(0, 0, 0), norb=1), 0.0, 1.0)
    for vs, vol in (([(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)], 4 / 3),
                    ([(0, 0, 0), (1, 0, 0), (0.3, 1, 0), (0.2, 0.1, 0.7), (1.2, 0.1, 0.7), (0.5, 1.1, 0.7)], 0.35)):
        sol = ab.solve(ab.IntegralProblem(one, ab.PolyhedronLimits(np.array(vs, dtype=float))), ab.NestedQuad(ab.AuxQuadGKJL()), abstol=1e-10)
        assert abs(sol.u - vol) < 1e-8


def test_limits_callback_errors_surface(ctx, svo):
    """a limits object that misbehaves (descending breakpoints / an exception) ends the solve with an error, not a hang"""
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    f = ab.FourierIntegrand(ab.dos_integrand, fs, 0.05)

    class Bad(ab.SegmentedLimits):
        def fix(self, x):
            raise RuntimeError("no inner limits")

    with pytest.raises(RuntimeError):
        ab.solve(ab.IntegralProblem(f, Bad((0.0, 0.5), (0.0, 0.5), (0.0, 0.5)), 12.5), ab.NestedQuad(ab.AuxQuadGKJL()), abstol=1e-2)

    class Desc(ab.SegmentedLimits):
        def segments(self):
            return (0.5, 0.0)

    with pytest.raises(ValueError):
        ab.solve(ab.IntegralProblem(f, Desc((0.0, 0.5), (0.0, 0.5), (0.0, 0.5)), 12.5), ab.NestedQuad(ab.AuxQuadGKJL()), abstol=1e-2)
