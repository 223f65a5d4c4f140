"""IAI with a Gauss-Kronrod order per variable (IAI(algs...), src/brillouin.jl:368-377; AuxQuadGKJL(order), src/algorithms.jl:202-208;
algs[dim] belongs to variable dim, src/algorithms.jl:462-463).  Orders other than 7 run the engine on the host with the device
evaluating the panel nodes (abz_nest_contract3 / contract2 / eval): identical decisions to the same engine over the CPU oracle, whose
control flow tests/test_host_logic.py checks against the sequential recursion."""
import numpy as np
import pytest

import autobz_b200 as ab
from oracle_backend import OracleBackend

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("orders", [(10, 10, 10), (5, 7, 3), (15, 4, 9)])
def test_gk_orders_device_vs_oracle(ctx, svo, orders):
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    ibz = ab.load_bz(ab.CubicSymIBZ(), A)
    mult = abs(np.linalg.det(ibz.B)) * 48
    f = ab.FourierIntegrand(ab.dos_integrand, fs, 0.03)
    alg = ab.EvalCounter(ab.IAI(*[ab.AuxQuadGKJL(order=o) for o in orders]))
    a = ab.solve(ab.IntegralProblem(f, ibz, 12.5), alg, abstol=1e-3 * mult, backend=ab.DeviceBackend(ctx=ctx))
    b = ab.solve(ab.IntegralProblem(f, ibz, 12.5), alg, abstol=1e-3 * mult, backend=OracleBackend())
    assert a.numevals == b.numevals and a.numevals > np.prod([2 * o + 1 for o in orders])
    assert abs(a.u - b.u) <= 1e-10 * abs(b.u)
    # and the default rule gives the same integral within the two tolerances
    c = ab.solve(ab.IntegralProblem(f, ibz, 12.5), ab.IAI(), abstol=1e-3 * mult, backend=ab.DeviceBackend(ctx=ctx))
    assert abs(a.u - c.u) <= 2e-3 * mult


def test_gk_order_matrix_valued_and_2d(ctx):
    """2-d, matrix-valued, orders (12, 6): device against the CPU oracle under the same engine.  omega != 0 on the inversion-reduced BZ:
    no two panels are mirror images of each other, so no accept/refine decision hangs on a rounding-level tie"""
    n = 2
    H, lo = ab.synthetic.wannier_hamiltonian(n, 2, cubic=True)
    H = np.asfortranarray(H[:, :, :, :, 2])
    fs = ab.FourierSeries(H, period=1.0, lo=lo[:2], norb=n)
    bz2 = ab.load_bz(ab.InversionSymIBZ(2), np.eye(2))
    alg = ab.EvalCounter(ab.IAI(ab.AuxQuadGKJL(order=12), ab.AuxQuadGKJL(order=6)))
    sym = ab.GlocIntegrand(symmetrize=lambda bz, x: bz.nsyms * x)
    dev, cpu = ab.DeviceBackend(ctx=ctx), OracleBackend()
    a = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(sym, fs, eta=0.3), bz2, {"omega": 0.3}), alg, abstol=1e-3, backend=dev)
    b = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(sym, fs, eta=0.3), bz2, {"omega": 0.3}), alg, abstol=1e-3, backend=cpu)
    assert a.u.shape == (n, n) and a.numevals == b.numevals > 25 * 13
    assert np.max(np.abs(a.u - b.u)) <= 1e-11 * np.max(np.abs(b.u))
    t = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=0.3), bz2, {"omega": 0.3}), ab.IAI(), abstol=1e-3, backend=dev)
    assert abs(np.trace(a.u) - t.u) < 3e-3
