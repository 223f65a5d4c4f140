"""GGR density of states on the device (abz_rule_ggr_data / abz_rule_ggr_sum) against the oracle and against the
reference's own known answers (test/dos.jl:88-111, atol 1e-2 at npt = 200)."""
import numpy as np
import pytest

import autobz_b200 as ab
from autobz_b200 import _lib as L
from dos_models import CASES, bz_of, energies

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,model,exact,B,bzkind", CASES, ids=[f"{c[0]}-{c[4]}" for c in CASES])
def test_reference_dos_known_answers_device(name, model, exact, B, bzkind):
    h = model()
    bz = bz_of(bzkind, h.ndim)
    cache = ab.init(ab.DOSProblem(h, 0.0, bz), ab.GGR(npt=200))
    for e in energies(B):
        cache.domain = float(e)
        sol = ab.solve_(cache)
        assert abs(sol.u - exact(e)) < 1e-2, (name, bzkind, e, sol.u, exact(e))
    cache.domain = energies(B)
    us = ab.solve_(cache).u
    assert all(abs(u - exact(e)) < 1e-2 for u, e in zip(us, energies(B)))
    assert np.array_equal(us, ab.solve_(cache).u)            # bit-reproducible


@pytest.mark.parametrize("n,sym", [(1, False), (3, False), (3, True), (8, True), (33, False), (64, True)])
def test_ggr_data_and_sum_vs_oracle(ctx, orc, n, sym):
    """energies ascending and band velocities per node vs the oracle (<= 1e-10 of the band width), sums <= 1e-10"""
    N = 6 if n <= 8 else 4
    H, lo = ab.synthetic.wannier_hamiltonian(n, 1, cubic=sym)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    So = orc.Series(H, lo)
    syms = np.array(ab.cube_automorphisms(3), dtype=np.int32)
    ws = orc.symptr_rule(N, syms)[0] if sym else None
    R = L.DeviceRule(ctx, S, N, syms=syms) if sym else L.DeviceRule(ctx, S, N)
    wo, eo, vo = orc.ggr_data(So, 3, N, ws)
    e, v = R.ggr_data(3)
    scale = max(1.0, float(np.max(np.abs(eo))))
    assert e.shape == eo.shape and v.shape == vo.shape
    assert np.max(np.abs(e - eo)) < 1e-11 * scale
    # velocities of (nearly) degenerate bands depend on the eigenvector basis: compare where the gap is healthy
    gap = np.minimum(np.diff(eo, axis=1, prepend=-np.inf), np.diff(eo, axis=1, append=np.inf))
    ok = gap > 1e-6 * scale
    assert ok.mean() > 0.9
    vs = max(1.0, float(np.max(np.abs(vo))))
    assert np.max(np.abs(v - vo)[np.broadcast_to(ok[:, None, :], v.shape)]) < 1e-8 * vs
    E = np.linspace(eo.min() - 0.1, eo.max() + 0.1, 23)
    ref = orc.ggr_sum(3, N, E, wo, eo, vo)
    got = R.ggr_sum(E)
    # the formula is piecewise: tiny velocity differences move nodes across the window edges only at measure-zero E
    assert np.max(np.abs(got - ref)) < 1e-6 * max(1.0, np.max(np.abs(ref)))
    # k3 shards add up
    if sym:
        parts = [L.DeviceRule(ctx, S, N, syms=syms, k3_lo=r, k3_stride=2) for r in range(2)]
    else:
        parts = [L.DeviceRule(ctx, S, N, k3_lo=0, k3_hi=N // 2), L.DeviceRule(ctx, S, N, k3_lo=N // 2, k3_hi=N)]
    for p in parts:
        p.ggr_data(3, copy=False)
    assert np.max(np.abs(sum(p.ggr_sum(E) for p in parts) - got)) < 1e-12 * max(1.0, np.max(np.abs(got)))


def test_ggr_errors(ctx):
    H, lo = ab.synthetic.wannier_hamiltonian(2, 1)
    R = L.DeviceRule(ctx, L.DeviceSeries(ctx, H, lo, (1.0,) * 3), 4)
    with pytest.raises(ValueError):
        R.ggr_sum([0.0])                  # data pass missing
    with pytest.raises(ValueError):
        R.ggr_data(2)                     # a 3-d series on a 2-d domain: "variables in Fourier series don't match domain"
