"""GGR density of states (src/dos_ggr.jl): the reference's own known-answer test, test/dos.jl:88-111
(ten model x BZ combinations against exact DOS formulas, atol 1e-2 at npt = 200), pins the oracle's restatement of the
data pass (energies + band velocities from the JacobianSeries) and of ggr_formula, and the host mirror of
DOSProblem / GGR / init / solve! including the cache semantics (cache.domain = e; solve!(cache))."""
import numpy as np
import pytest

import autobz_b200 as ab
from dos_models import CASES, bz_of, energies
from oracle_backend import OracleBackend


@pytest.mark.parametrize("name,model,exact,B,bzkind", CASES, ids=[f"{c[0]}-{c[4]}" for c in CASES])
def test_reference_dos_known_answers_oracle(orc, name, model, exact, B, bzkind):
    h = model()
    bz = bz_of(bzkind, h.ndim)
    prob = ab.DOSProblem(h, 0.0, bz)
    cache = ab.init(prob, ab.GGR(npt=200), backend=OracleBackend())
    for e in energies(B):
        cache.domain = float(e)
        sol = ab.solve_(cache)
        assert sol.retcode and sol.numevals == -1 and sol.err is None
        assert abs(sol.u - exact(e)) < 1e-2, (name, bzkind, e, sol.u, exact(e))
    # one data pass serves a whole energy list
    cache.domain = energies(B)
    us = ab.solve_(cache).u
    assert us.shape == (10,) and all(abs(u - exact(e)) < 1e-2 for u, e in zip(us, energies(B)))


def test_dos_cache_and_argument_errors(orc):
    from dos_models import tb_integer
    h = tb_integer(2)
    bz = bz_of("FBZ", 2)
    be = OracleBackend()
    cache = ab.init(ab.DOSProblem(h, 0.4, bz), ab.GGR(npt=40), backend=be)
    u1 = ab.solve_(cache).u
    assert not cache.isfresh
    cache.H = tb_integer(2, t=0.5)         # replacing H marks the cache fresh: the data pass is redone (src/dos_interfaces.jl:59-64)
    assert cache.isfresh
    u2 = ab.solve_(cache).u
    assert not cache.isfresh and abs(u2 - u1) > 1e-3
    with pytest.raises(TypeError):
        ab.init(ab.DOSProblem(lambda k: k, 0.0, bz), ab.GGR(), backend=be)                     # src/dos_ggr.jl:2
    with pytest.raises(TypeError):
        ab.init(ab.DOSProblem(h, 0.0, ab.CubicLimits([0.0] * 2, [1.0] * 2)), ab.GGR(), backend=be)   # src/dos_ggr.jl:3
    with pytest.raises(ValueError):
        ab.init(ab.DOSProblem(h, 0.0, bz), ab.GGR(), backend=be, tol=1)                        # checkkwargs
    assert ab.GGR().npt == 50


def test_ggr_velocities_are_band_derivatives(orc):
    """v_n,d = real(diag(U' dH/dk_d U)) * period_d is the band derivative (Hellmann-Feynman): compare with a centred
    finite difference of the oracle's eigenvalues on a 3-orbital synthetic Hamiltonian with non-degenerate bands."""
    H, lo = ab.synthetic.wannier_hamiltonian(3, 1)
    S = orc.Series(H, lo)
    N = 6
    w, e, v = orc.ggr_data(S, 3, N)
    k = np.array([[i1, i2, i3] for i3 in range(N) for i2 in range(N) for i1 in range(N)], dtype=float) / N
    hstep = 1e-5
    for d in range(3):
        dk = np.zeros(3); dk[d] = hstep
        ep = orc.eigvals_batch(orc.eval_points(S, k + dk))
        em = orc.eigvals_batch(orc.eval_points(S, k - dk))
        fd = (ep - em) / (2 * hstep)
        assert np.max(np.abs(fd - v[:, d, :])) < 1e-6 * max(1.0, np.max(np.abs(fd)))
    assert np.allclose(e, orc.eigvals_batch(orc.eval_points(S, k)), atol=1e-12)
