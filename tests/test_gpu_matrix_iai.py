"""IAI on the MATRIX-valued Green's function (docs/src/examples.md:20,90 `gloc_integrand` = inv(complex(omega, eta) I - h_k.s);
the reference's nest is generic in the value type, src/fourier.jl:432-510; error norm = LinearAlgebra.norm = Frobenius).
Device path: abz_nest_eval_matrix gives (z - H - Sigma)^-1 at the nodes of every innermost panel; the host engine keeps the
15 matrices of each panel.  Checked against the same engine over the CPU oracle (identical decisions => identical numevals)."""
import numpy as np
import pytest

import autobz_b200 as ab
from oracle_backend import OracleBackend

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 3, 24])
def test_nest_eval_matrix_vs_lapack(ctx, orc, n):
    """the raw entry point: matrices at arbitrary nodes of a 1-D / 3-D nest against numpy's inverse of the oracle's H(k)"""
    from autobz_b200 import _lib as L
    H, lo = ab.synthetic.wannier_hamiltonian(n, 2)
    rng = np.random.default_rng(n)
    ser = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    nest = L.DeviceNest(ctx, ser, 3, 4, 8)
    x3, x2, x1 = rng.uniform(-1, 1, 2), rng.uniform(-1, 1, 5), rng.uniform(-1, 1, 37)
    nest.contract3(x3, np.arange(2))
    par = rng.integers(0, 2, 5)
    nest.contract2(x2, par, np.arange(5))
    s1 = rng.integers(0, 5, 37)
    z = complex(0.2, 0.07)
    sig = 0.05 * (rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n)))
    for sigma in (None, sig):
        G = nest.eval_matrix(x1, s1, z, sigma)
        assert G.shape == (37, n, n)
        k = np.stack([x1, x2[s1], x3[par[s1]]], axis=1)
        Hk = np.moveaxis(orc.eval_points(orc.Series(H, lo), k), 2, 0)
        ref = np.linalg.inv(z * np.eye(n)[None] - Hk - (0 if sigma is None else sigma[None]))
        assert np.max(np.abs(G - ref)) <= 1e-11 * np.max(np.abs(ref))


def test_matrix_iai_device_vs_oracle_engine(ctx, orc):
    n = 3
    H, lo = ab.synthetic.wannier_hamiltonian(n, 1, cubic=True)
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=n)
    fbz, ibz = ab.load_bz(ab.FBZ(), np.eye(3)), ab.load_bz(ab.CubicSymIBZ(), np.eye(3))
    p = {"omega": 0.3}
    sym = ab.GlocIntegrand(symmetrize=lambda bz, x: bz.nsyms * x)
    dev, cpu = ab.DeviceBackend(ctx=ctx), OracleBackend()
    for f, bz in ((ab.gloc_integrand, fbz), (sym, ibz)):
        a = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(f, fs, eta=4.0), bz, p), ab.EvalCounter(ab.IAI()), abstol=0.1, backend=dev)
        b = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(f, fs, eta=4.0), bz, p), ab.EvalCounter(ab.IAI()), abstol=0.1, backend=cpu)
        assert a.u.shape == (n, n) and a.numevals == b.numevals
        assert np.max(np.abs(a.u - b.u)) <= 1e-11 * np.max(np.abs(b.u)) and abs(a.resid - b.resid) <= 1e-8 * b.resid + 1e-13
    fine = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_integrand, fs, eta=4.0), fbz, p), ab.PTR(npt=32), backend=dev).u
    assert np.max(np.abs(a.u - fine)) < 0.1
    # the trace of the matrix integral agrees with the scalar (native-engine) integral within the two tolerances
    t = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=4.0), fbz, p), ab.IAI(), abstol=0.1, backend=dev).u
    assert abs(np.trace(fine) - t) < 0.1


def test_matrix_iai_docs_example_2d(ctx):
    """docs/src/examples.md:90-105: IAI(abstol 1e-3) of gloc_integrand for the 2-D scalar series; a 1 x 1 matrix takes exactly the
    scalar integrand's decisions"""
    c2 = np.zeros((1, 1, 3, 3)); c2[0, 0, 0, 1] = c2[0, 0, 2, 1] = c2[0, 0, 1, 0] = c2[0, 0, 1, 2] = 0.5
    h2 = ab.FourierSeries(c2, period=1.0, lo=(-1, -1), norb=1)
    bz2 = ab.load_bz(ab.FBZ(2), np.eye(2) * 2 * np.pi)
    be = ab.DeviceBackend(ctx=ctx)
    sm = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_integrand, h2, eta=0.1), bz2, {"omega": 0.0}), ab.EvalCounter(ab.IAI()), abstol=1e-3, backend=be)
    ss = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_trace_integrand, h2, eta=0.1), bz2, {"omega": 0.0}), ab.EvalCounter(ab.IAI()), abstol=1e-3, backend=be)
    assert sm.u.shape == (1, 1) and sm.numevals == ss.numevals and abs(sm.u[0, 0] - ss.u) <= 1e-11 * abs(ss.u)
