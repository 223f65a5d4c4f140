"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs,
against the committed golden fixtures, and — at BASELINE.json's full sizes — through size-independent
properties (slab additivity, IBZ == FBZ, run-to-run bit reproducibility).

Tolerances (north_star): integrals <= 1e-10 relative; fixed-N PTR sums <= 1e-12 relative (only the
summation order and phase rounding differ); IAI: identical numevals to the oracle's restated control flow."""
import json
import os

import numpy as np
import pytest

import autobz_b200 as ab
from autobz_b200 import _lib as L

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def lattice_series(d):
    c, lo = ab.synthetic.integer_lattice(d)
    return ab.FourierSeries(c[0, 0], period=1.0, lo=lo)


# ---------------------------------------------------------------------------------------------------
# low-level ABI vs oracle
@pytest.mark.parametrize("n,rmax,N", [(1, 1, 16), (2, 1, 9), (3, 2, 12), (5, 2, 10), (8, 1, 8), (17, 1, 5), (32, 1, 6), (64, 1, 4)])
def test_rule_sums_and_values_vs_oracle(ctx, orc, n, rmax, N):
    H, lo = ab.synthetic.wannier_hamiltonian(n, rmax)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    So = orc.Series(H, lo)
    z = np.array([0.3 + 0.05j, -0.7 + 0.2j, 1.5 + 0.01j])
    rng = np.random.default_rng(n)
    sig = 0.1 * (rng.standard_normal((n, n, 3)) + 1j * rng.standard_normal((n, n, 3)))
    R = L.DeviceRule(ctx, S, N)
    ref = orc.ptr_sum(So, N, z)
    # norb > 32: the default path is the block LU without pivoting between blocks (DMMA teams); its error grows like
    # norb * eps * |A| / eta (profiles/r02_team_resolvent_accuracy.log) - within the 1e-10 of north_star, not 1e-11, at eta = 0.01
    tol = 1e-11 if n <= 32 else 1e-10
    assert rel(R.resolvent_sum(z, scale=1 / N ** 3), ref) < tol
    assert rel(R.resolvent_sum(z, sigma=sig, scale=1 / N ** 3), orc.ptr_sum(So, N, z, sigma=sig)) < tol
    if n > 32:       # the pivoted teams keep 1e-11
        ctx.set_option(L.OPT_RESOLVENT_ALGO, 1)
        try:
            assert rel(R.resolvent_sum(z, scale=1 / N ** 3), ref) < 1e-11
        finally:
            ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)
    Hk, k, w = R.copy_out()
    assert rel(Hk.reshape(n, n, N, N, N, order="F"), orc.grid_eval_full(So, N)) < 1e-13
    assert np.all(w == 1.0) and k.shape == (N ** 3, 3) and abs(k[1, 0] - 1 / N) < 1e-16
    R.materialize()
    assert rel(R.resolvent_sum(z, scale=1 / N ** 3), ref) < tol
    # k3 slabs are additive (the multi-GPU shard unit)
    parts = [L.DeviceRule(ctx, S, N, k3_lo=a, k3_hi=b).resolvent_sum(z, scale=1 / N ** 3) for a, b in ((0, 1), (1, 3), (3, N))]
    assert rel(sum(parts), ref) < tol
    # eigenvalues vs LAPACK and eigenvalue sums vs oracle
    ev = R.eigvals()
    assert rel(ev, np.linalg.eigvalsh(np.moveaxis(Hk, 2, 0))) < 1e-11
    for kind, prm in ((L.EIG_SUM, (0.0, 1.0)), (L.EIG_FERMI_ENERGY, (0.1, 0.3)), (L.EIG_FERMI_COUNT, (0.1, 0.3)), (L.EIG_GAUSS_DOS, (0.2, 0.4))):
        assert abs(R.eig_sum(kind, prm, 1 / N ** 3) - orc.ptr_eig_sum(So, N, kind, prm, scale=1 / N ** 3)[0]) < 1e-11 * n
    # scattered points
    kp = rng.random((9, 3))
    assert rel(S.eval_points(kp), orc.eval_points(So, kp)) < 1e-13
    assert rel(S.points_resolvent(kp, z), orc.resolvent_trace_batch(orc.eval_points(So, kp), z)) < tol


@pytest.mark.parametrize("algo", [1, 2])
def test_resolvent_algorithms_agree(ctx, orc, algo):
    """generic pivoted Gauss-Jordan (1) and the DMMA register path (2) against the oracle's pivoted LU, incl. a
    matrix-valued self-energy and a small eta (cond ~ 1e4)"""
    n, N = 32, 6
    H, lo = ab.synthetic.wannier_hamiltonian(n, 2)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    So = orc.Series(H, lo)
    rng = np.random.default_rng(7)
    z = np.concatenate([rng.uniform(-3, 3, 6) + 0.02j, rng.uniform(-3, 3, 3) + 1e-4j])
    sig = 0.05 * (rng.standard_normal((n, n, 9)) + 1j * rng.standard_normal((n, n, 9)))
    sig = sig - 0.1j * np.eye(n)[:, :, None]
    ctx.set_option(L.OPT_RESOLVENT_ALGO, algo)
    try:
        R = L.DeviceRule(ctx, S, N)
        assert rel(R.resolvent_sum(z, scale=1 / N ** 3), orc.ptr_sum(So, N, z)) < 1e-10
        assert rel(R.resolvent_sum(z, sigma=sig, scale=1 / N ** 3), orc.ptr_sum(So, N, z, sigma=sig)) < 1e-10
        kp = rng.random((5, 3))
        assert rel(S.points_resolvent(kp, z, sigma=sig), orc.resolvent_trace_batch(orc.eval_points(So, kp), z, sig)) < 1e-10
    finally:
        ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)


@pytest.mark.parametrize("n", [4, 9, 31, 32, 33, 47, 64])
@pytest.mark.parametrize("algo", [1, 4])
def test_pivoted_gauss_jordan_needs_its_pivoting(ctx, orc, n, algo):
    """the pivoted kernels (1: register-resident teams, 4: shared-memory) on matrices whose leading minors vanish:
    H has a zero diagonal and z = 0 + i 1e-9, so elimination without row exchanges would divide by ~1e-9;
    traces (weighted sum and per point) and the matrix-valued sum against the oracle's pivoted LU / numpy"""
    rng = np.random.default_rng(100 + n)
    c = np.zeros((n, n, 3, 1, 1), complex)
    for m in range(3):
        a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        np.fill_diagonal(a, 0.0)
        c[:, :, m, 0, 0] = a
    c[:, :, 2, 0, 0] = c[:, :, 0, 0, 0].conj().T
    c[:, :, 1, 0, 0] = c[:, :, 1, 0, 0] + c[:, :, 1, 0, 0].conj().T
    lo = (-1, 0, 0)
    S = L.DeviceSeries(ctx, c, lo, (1.0,) * 3)
    So = orc.Series(c, lo)
    z = np.array([1e-9j, 0.3 + 1e-9j, -0.7 + 0.05j])
    N = 4
    ctx.set_option(L.OPT_RESOLVENT_ALGO, algo)
    try:
        R = L.DeviceRule(ctx, S, N)
        assert rel(R.resolvent_sum(z, scale=1 / N ** 3), orc.ptr_sum(So, N, z)) < 1e-10
        kp = rng.random((7, 3))
        Hk = orc.eval_points(So, kp)
        assert rel(S.points_resolvent(kp, z), orc.resolvent_trace_batch(Hk, z)) < 1e-10
        G = R.resolvent_matrix_sum(z, scale=1 / N ** 3)
        want = np.zeros((len(z), n, n), complex)
        for i in range(N):
            Hi = orc.eval_points(So, np.array([[i / N, 0.0, 0.0]]))[:, :, 0]
            for w, zz in enumerate(z):
                want[w] += np.linalg.inv(zz * np.eye(n) - Hi) / N
        assert np.max(np.abs(G - want)) < 1e-10 * np.max(np.abs(want))
    finally:
        ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)


def test_singular_matrix_is_an_error(ctx):
    c = np.zeros((2, 2, 1, 1, 1))
    S = L.DeviceSeries(ctx, c, (0, 0, 0), (1.0,) * 3)
    R = L.DeviceRule(ctx, S, 2)
    with pytest.raises(ab.SingularIntegrandError):
        R.resolvent_sum([0.0 + 0.0j])          # z I - 0 = 0: singular -> DomainError-like
    c4 = np.zeros((5, 5, 1, 1, 1))
    R4 = L.DeviceRule(ctx, L.DeviceSeries(ctx, c4, (0, 0, 0), (1.0,) * 3), 2)
    with pytest.raises(ab.SingularIntegrandError):
        R4.resolvent_sum([0.0 + 0.0j])
    assert abs(R4.resolvent_sum([2.0 + 0.0j], scale=1 / 8)[0] - 2.5) < 1e-14     # ctx usable afterwards


def test_invalid_arguments_raise(ctx):
    H, lo = ab.synthetic.wannier_hamiltonian(2, 1)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    with pytest.raises(ValueError):
        L.DeviceRule(ctx, S, 4, k3_lo=3, k3_hi=9)
    with pytest.raises(ValueError):
        L.DeviceRule(ctx, S, 4, nodes=[[1, 0, 0], [0, 0, 0]])           # unsorted
    R = L.DeviceRule(ctx, S, 4, nodes=np.zeros((0, 3), dtype=np.int32))  # empty rule is fine (a rank with no planes)
    assert len(R) == 0 and R.resolvent_sum([1.0 + 1j])[0] == 0
    with pytest.raises(ValueError):
        L.DeviceSeries(ctx, np.zeros((65, 65, 1, 1, 1)), (0, 0, 0), (1.0,) * 3) if False else L.DeviceRule(ctx, S, 0)


def test_symmetric_rules_vs_oracle(ctx, orc):
    syms = ab.cube_automorphisms(3)
    for n, rmax, N in [(3, 2, 12), (3, 2, 15), (6, 1, 10), (32, 1, 8)]:
        H, lo = ab.synthetic.wannier_hamiltonian(n, rmax, cubic=True)
        S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
        So = orc.Series(H, lo)
        w_o, nirr_o = orc.symptr_rule(N, syms)
        w_d, nirr_d = ctx.symptr_rule(N, np.array(syms, dtype=np.int32))
        assert np.array_equal(w_o, w_d) and nirr_o == nirr_d and w_d.sum() == N ** 3
        z = np.array([0.3 + 0.05j, -0.7 + 0.2j])
        R = L.DeviceRule(ctx, S, N, wsym=w_d)
        ref, cnt = orc.symptr_sum(So, N, w_o, z, scale=1 / N ** 3)
        assert len(R) == cnt == nirr_o
        assert rel(R.resolvent_sum(z, scale=1 / N ** 3), ref) < 1e-11
        assert rel(ref, orc.ptr_sum(So, N, z)) < 1e-12                       # IBZ-weighted == FBZ
        R.materialize()
        assert rel(R.resolvent_sum(z, scale=1 / N ** 3), ref) < 1e-11
        shards = [L.DeviceRule(ctx, S, N, wsym=w_d, k3_lo=r, k3_stride=3).resolvent_sum(z, scale=1 / N ** 3) for r in range(3)]
        assert rel(sum(shards), ref) < 1e-11
        assert abs(R.eig_sum(L.EIG_SUM, scale=1 / N ** 3) - orc.ptr_eig_sum(So, N, 0, (0, 1), wsym=w_o, scale=1 / N ** 3)[0]) < 1e-11 * n


def test_nest_arena_vs_oracle(ctx, orc):
    for n in (1, 3, 6):
        H, lo = ab.synthetic.wannier_hamiltonian(n, 2)
        S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
        So = orc.Series(H, lo)
        nest = L.DeviceNest(ctx, S, 3, 4, 8)
        rng = np.random.default_rng(5)
        x3 = rng.random(3); nest.contract3(x3, [0, 1, 3])
        x2 = rng.random(5); par = np.array([0, 1, 3, 3, 0]); sl1 = [7, 0, 2, 3, 5]; nest.contract2(x2, par, sl1)
        x1 = rng.random(11); s1 = rng.choice(sl1, 11)
        m = {s: i for i, s in enumerate(sl1)}
        m3 = {0: 0, 1: 1, 3: 2}
        kp = np.array([[x1[i], x2[m[s1[i]]], x3[m3[par[m[s1[i]]]]]] for i in range(11)])
        y = nest.eval(x1, s1, 0.2 + 0.1j)
        assert rel(y, orc.resolvent_trace_batch(orc.eval_points(So, kp), [0.2 + 0.1j])[:, 0]) < 1e-11
        with pytest.raises(ValueError):
            nest.contract3([0.1], [99])


# ---------------------------------------------------------------------------------------------------
# the reference's own tests through the public API on the device
@pytest.mark.parametrize("d", [1, 2, 3])
@pytest.mark.parametrize("bzkind", ["fbz", "inv"])
def test_fourier_jl_algorithms_on_device(d, bzkind):
    """test/fourier.jl:40-56"""
    vol = (2 * np.pi) ** d
    s = lattice_series(d)
    bz = ab.load_bz(ab.FBZ() if bzkind == "fbz" else ab.InversionSymIBZ(), np.eye(d))
    prob = ab.IntegralProblem(ab.FourierIntegrand(ab.AffineTraceIntegrand(), s, 1.3, b=1.0), bz)
    for alg in (ab.IAI(), ab.PTR(), ab.AutoPTR()):
        for counter in (False, True):
            solver = ab.IntegralSolver(prob, ab.EvalCounter(alg) if counter else alg, reltol=0, abstol=1e-6)
            assert abs(solver() - vol) < 1e-6


def test_docs_goldens_on_device():
    # docs/src/examples.md:44-60 (1-D) and :79-106 (2-D IAI on FBZ(2)), abstol = 1e-3
    h1 = ab.FourierSeries([0.5, 0.0, 0.5], period=1, offset=-2)
    bz1 = ab.SymmetricBZ(np.eye(1) * 2 * np.pi, np.eye(1), ab.CubicLimits([0.0], [1.0]), None)
    g1 = ab.IntegralSolver(ab.FourierIntegrand(ab.gloc_trace_integrand, h1, eta=0.1), bz1, ab.IAI(), abstol=1e-3)(omega=0.0)
    ref1 = GOLD["reference_known_answers"]["docs/src/examples.md:60 (QuadGKJL abstol=1e-3, 1-D gloc, eta=0.1, omega=0)"]
    assert abs(g1.imag - ref1[1]) < 1e-14 and abs(g1.real) < 1e-14
    C2 = np.array([[0.0, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.0]])
    h2 = ab.FourierSeries(C2, period=1, offset=-2)
    bz2 = ab.load_bz(ab.FBZ(2), 2 * np.pi * np.eye(2))
    g2 = ab.IntegralSolver(ab.FourierIntegrand(ab.gloc_trace_integrand, h2, eta=0.1), bz2, ab.IAI(), abstol=1e-3)(omega=0.0)
    ref2 = GOLD["reference_known_answers"]["docs/src/examples.md:105 (IAI abstol=1e-3, 2-D gloc on FBZ(2), eta=0.1, omega=0)"]
    assert abs(g2.imag - ref2[1]) < 1e-13 and abs(g2.real) < 1e-13


def test_c1_config_values():
    """BASELINE config 1: cubic 1-orbital cos band, PTR 64^3, eta = 0.1"""
    s = lattice_series(3)
    bz = ab.load_bz(ab.FBZ(), 2 * np.pi * np.eye(3))      # B = I => j = 1
    solver = ab.IntegralSolver(ab.FourierIntegrand(ab.gloc_trace_integrand, s, eta=0.1), bz, ab.PTR(npt=64))
    g = ab.batchsolve(solver, [{"omega": 0.0}, {"omega": 0.5}])
    assert abs(g[0] - (-2.361629003144814j)) < 1e-12
    assert abs(g[1] - (1.448114087711081 - 1.4016191114904277j)) < 1e-12
    ref = GOLD["oracle_values"]["c1_ptr_N64_eta0.1"]
    assert abs(g[1] - complex(*ref[1])) < 1e-12 * abs(g[1])


def test_c2_svo_ptr_autoptr(orc, svo):
    """BASELINE config 2: SrVO3 Green's-function trace, PTR goldens + AutoPTR on CubicSymIBZ vs FBZ"""
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    bz, ibz = ab.load_bz(ab.FBZ(), A), ab.load_bz(ab.CubicSymIBZ(), A)
    j = abs(np.linalg.det(bz.B))
    assert abs(j - 4.31781301953062) < 1e-12
    f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=0.01)
    ws = [11.0, 12.0, 12.975161, 13.5]
    for N in (24, 50):
        ref = np.array([complex(*v) for v in GOLD["oracle_values"][f"svo_fbz_ptr_N{N}_eta1e-2"]]) * j
        for dom in (bz, ibz):
            got = ab.batchsolve(ab.IntegralSolver(f, dom, ab.PTR(npt=N)), [{"omega": w} for w in ws])
            assert rel(got, ref) < 1e-12
    # DOS integrand of aps_example.jl:30-34 (positional eta, omega)
    dos = ab.IntegralSolver(ab.FourierIntegrand(ab.dos_integrand, fs, 0.01), ibz, ab.PTR(npt=50))
    ref = -np.imag(complex(*GOLD["oracle_values"]["svo_fbz_ptr_N50_eta1e-2"][1]) * j) / np.pi
    assert abs(dos(12.0) - ref) < 1e-12 * abs(ref)
    # AutoPTR: same decisions and value as the oracle-driven control flow
    from oracle_backend import OracleBackend
    alg = ab.EvalCounter(ab.AutoPTR(a=0.2, nmin=20, nmax=400))          # grids 30, 42, 54, ...
    f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=0.1)
    for w in (11.0, 12.5):
        sd = ab.solve(ab.IntegralProblem(f, ibz, {"omega": w}), alg, abstol=1e-2)
        so = ab.solve(ab.IntegralProblem(f, ibz, {"omega": w}), alg, abstol=1e-2, backend=OracleBackend())
        assert sd.numevals == so.numevals
        assert abs(sd.u - so.u) < 1e-10 * abs(so.u)


def test_c3_svo_iai_counts_and_values(orc, svo):
    """BASELINE config 3: SrVO3 DOS via IAI (nested GK panels): identical adaptive evaluation counts to the
    oracle's sequential recursion and integrals within 1e-10 relative."""
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    S = orc.Series(H, lo)
    ibz = ab.load_bz(ab.CubicSymIBZ(), A)
    mult = abs(np.linalg.det(ibz.B)) * 48
    for eta, omega, atol in ((0.05, 12.5, 1e-2), (0.01, 12.0, 1e-2), (1e-3, 12.975161, 2e-2)):
        Io, Eo, neo = orc.iai(S, 3, 1, [0.5] * 3, vkind=1, z=complex(omega, eta), atol=atol)
        f = ab.FourierIntegrand(ab.dos_integrand, fs, eta)
        sol = ab.solve(ab.IntegralProblem(f, ibz, omega), ab.EvalCounter(ab.IAI()), abstol=atol * mult)
        assert sol.numevals == neo
        assert abs(sol.u - mult * Io.real) <= 1e-10 * abs(sol.u)
    g = GOLD["oracle_values"]["svo_iai_tetra_dos_w12.5_eta0.05_atol1e-2"]
    sol = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.dos_integrand, fs, 0.05), ibz, 12.5), ab.EvalCounter(ab.IAI()), abstol=1e-2 * mult)
    assert sol.numevals == g["numevals"] and abs(sol.u / mult - g["I"]) < 1e-10 * abs(g["I"])


def test_c5_band_energy_ibz_autoptr(orc):
    """BASELINE config 5 (reduced grid for the oracle): norb=64 band-energy integrand on CubicSymIBZ, AutoPTR"""
    n = 64
    H, lo = ab.synthetic.wannier_hamiltonian(n, 1, cubic=True)
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=n)
    So = orc.Series(H, lo)
    ibz = ab.load_bz(ab.CubicSymIBZ(), 2 * np.pi * np.eye(3))
    f = ab.FourierIntegrand(ab.EigenIntegrand("fermi_energy"), fs, 0.0, 0.5)
    sol = ab.solve(ab.IntegralProblem(f, ibz), ab.PTR(npt=8))
    w, nirr = orc.symptr_rule(8, ab.cube_automorphisms(3))
    ref, cnt = orc.ptr_eig_sum(So, 8, 1, (0.0, 0.5), wsym=w, scale=1 / 8 ** 3)
    assert abs(sol.u - ref) < 1e-10 * abs(ref)
    sol = ab.solve(ab.IntegralProblem(f, ibz), ab.EvalCounter(ab.AutoPTR(nmin=6, a=1.0)), reltol=1e-3)
    assert sol.numevals > 0 and np.isfinite(sol.u)


def test_generic_python_integrand_copies_h_back():
    s = lattice_series(2)
    bz = ab.load_bz(ab.FBZ(), np.eye(2))
    f = ab.FourierIntegrand(lambda x, a, b=0.0: a * x.s + b, s, 1.3, b=1.0)
    assert abs(ab.solve(ab.IntegralProblem(f, bz), ab.PTR(npt=8)).u - (2 * np.pi) ** 2) < 1e-10


# ---------------------------------------------------------------------------------------------------
# full-size properties (no oracle at this size)
def test_c4_full_size_properties(ctx):
    """BASELINE config 4 shape: norb=32, R in [-8,8]^3, 256^3 grid (one k3 plane here), many frequencies.
    Properties: bit-reproducible run to run; plane sums additive; generic pivoted path == fast path;
    Im tr G <= 0 for every frequency (causality)."""
    n = 32
    H, lo = ab.synthetic.wannier_hamiltonian(n, 8)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    ext = ab.synthetic.band_extent(H)
    z = np.linspace(-0.2 * ext, 0.2 * ext, 16) + 1j * 0.01 * ext
    N = 256
    # a 1/16 plane through the node-list interface keeps the test short: rows k2 < 16 of plane k3 = 5
    i1, i2 = np.meshgrid(np.arange(N), np.arange(16), indexing="xy")
    nodes = np.stack([i1.ravel(), i2.ravel(), np.full(i1.size, 5)], axis=1).astype(np.int32)
    R = L.DeviceRule(ctx, S, N, nodes=nodes)
    a = R.resolvent_sum(z)
    b = R.resolvent_sum(z)
    assert np.array_equal(a, b)                                   # deterministic reduction
    assert np.all(a.imag < 0)
    half = [L.DeviceRule(ctx, S, N, nodes=nodes[: nodes.shape[0] // 2]).resolvent_sum(z),
            L.DeviceRule(ctx, S, N, nodes=nodes[nodes.shape[0] // 2:]).resolvent_sum(z)]
    assert rel(half[0] + half[1], a) < 1e-12
    ctx.set_option(L.OPT_RESOLVENT_ALGO, 1)
    try:
        c = R.resolvent_sum(z[:4])
    finally:
        ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)
    assert rel(c, a[:4]) < 1e-11


# ---------------------------------------------------------------------------------------------------
# abz_iai_solve: the library's C++ host engine, with host-driven panels and with device-side innermost integrals
@pytest.mark.parametrize("engine,leaves", [("python", False), ("native", False), ("native", True)])
def test_iai_engines_agree_with_oracle(ctx, orc, svo, engine, leaves):
    """All three ways of driving IAI (Python round loop over abz_nest_*, abz_iai_solve with host-driven innermost
    panels, abz_iai_solve with one warp per innermost adaptive integral) make the oracle's decisions: identical
    numevals, integrals <= 1e-10 relative."""
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    S = orc.Series(H, lo)
    be = ab.DeviceBackend(ctx=ctx, iai_engine=engine, iai_device_leaves=leaves)
    ibz = ab.load_bz(ab.CubicSymIBZ(), A)
    fbz = ab.load_bz(ab.FBZ(), A)
    j = abs(np.linalg.det(ibz.B))
    for eta, omega, atol in ((0.05, 12.5, 1e-2), (0.01, 12.0, 1e-2), (2e-3, 12.975161, 2e-2)):
        Io, Eo, neo = orc.iai(S, 3, 1, [0.5] * 3, vkind=1, z=complex(omega, eta), atol=atol)
        f = ab.FourierIntegrand(ab.dos_integrand, fs, eta)
        sol = ab.solve(ab.IntegralProblem(f, ibz, omega), ab.EvalCounter(ab.IAI()), abstol=atol * j * 48, backend=be)
        assert sol.numevals == neo
        assert abs(sol.u - j * 48 * Io.real) <= 1e-10 * abs(sol.u)
        assert abs(sol.resid - j * 48 * Eo) <= 1e-6 * sol.resid
    # complex-valued integrand on the full BZ (CubicLimits), relative tolerance only
    Io, Eo, neo = orc.iai(S, 3, 0, [0.0] * 3, [1.0] * 3, vkind=0, z=complex(12.3, 0.1), atol=0.0, rtol=1e-4)
    f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=0.1)
    sol = ab.solve(ab.IntegralProblem(f, fbz, {"omega": 12.3}), ab.EvalCounter(ab.IAI()), reltol=1e-4, backend=be)
    assert sol.numevals == neo
    assert abs(sol.u - j * Io) <= 1e-10 * abs(sol.u)


@pytest.mark.parametrize("leaves", [False, True])
def test_iai_native_low_dim_large_norb_and_errors(ctx, orc, leaves):
    be = ab.DeviceBackend(ctx=ctx, iai_engine="native", iai_device_leaves=leaves)
    # docs/src/examples.md:44-60 (1-D) and :79-106 (2-D IAI on FBZ(2)), abstol = 1e-3: reference known answers
    h1 = ab.FourierSeries([0.5, 0.0, 0.5], period=1, offset=-2)
    bz1 = ab.SymmetricBZ(np.eye(1) * 2 * np.pi, np.eye(1), ab.CubicLimits([0.0], [1.0]), None)
    g1 = ab.IntegralSolver(ab.FourierIntegrand(ab.gloc_trace_integrand, h1, eta=0.1), bz1, ab.IAI(), abstol=1e-3, backend=be)(omega=0.0)
    ref1 = GOLD["reference_known_answers"]["docs/src/examples.md:60 (QuadGKJL abstol=1e-3, 1-D gloc, eta=0.1, omega=0)"]
    assert abs(g1.imag - ref1[1]) < 1e-14 and abs(g1.real) < 1e-14
    C2 = np.array([[0.0, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.0]])
    h2 = ab.FourierSeries(C2, period=1, offset=-2)
    bz2 = ab.load_bz(ab.FBZ(2), 2 * np.pi * np.eye(2))
    g2 = ab.IntegralSolver(ab.FourierIntegrand(ab.gloc_trace_integrand, h2, eta=0.1), bz2, ab.IAI(), abstol=1e-3, backend=be)(omega=0.0)
    ref2 = GOLD["reference_known_answers"]["docs/src/examples.md:105 (IAI abstol=1e-3, 2-D gloc on FBZ(2), eta=0.1, omega=0)"]
    assert abs(g2.imag - ref2[1]) < 1e-13 and abs(g2.real) < 1e-13
    # 5-orbital resolvent through the generic (nodes -> H -> resolvent kernel -> combine) path, 2-D, vs oracle
    n = 5
    H, lo = ab.synthetic.wannier_hamiltonian(n, 2)
    H2, lo2 = np.asfortranarray(H[:, :, :, :, 2]), lo[:2]
    fs = ab.FourierSeries(H2, period=1.0, lo=lo2, norb=n)
    So = orc.Series(H2[..., None], tuple(lo2) + (0,))
    z = complex(0.2, 0.15)
    Io, Eo, neo = orc.iai(So, 2, 0, [0.0] * 2, [1.0] * 2, vkind=0, z=z, atol=1e-4)
    bz2 = ab.load_bz(ab.FBZ(2), 2 * np.pi * np.eye(2))
    f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=0.15)
    sol = ab.solve(ab.IntegralProblem(f, bz2, {"omega": 0.2}), ab.EvalCounter(ab.IAI()), abstol=1e-4, backend=be)
    assert sol.numevals == neo and abs(sol.u - Io) <= 1e-10 * abs(Io)
    # NaN/Inf in the integrand surfaces as an error (QuadGK DomainError): z on a pole of a diagonal 1-orbital band
    hc = ab.FourierSeries([0.0, 1.0, 0.0], period=1, offset=-2)
    with pytest.raises(FloatingPointError):
        ab.IntegralSolver(ab.FourierIntegrand(ab.gloc_trace_integrand, hc, eta=0.0), bz1, ab.IAI(), abstol=1e-3, backend=be)(omega=1.0)


@pytest.mark.parametrize("npt", [7, 12, 33])
def test_rule_create_symptr_on_device_matches_host_path(ctx, orc, npt):
    """abz_rule_create_symptr (weights + CSR compaction on the device) == abz_symptr_rule -> abz_rule_create_sym
    == the oracle's symptr_rule: same nodes, order, weights, sums; sharded by k3 planes the parts add up."""
    n = 3
    H, lo = ab.synthetic.wannier_hamiltonian(n, 2, cubic=True)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    syms = np.array(ab.cube_automorphisms(3), dtype=np.int32)
    w_host, nirr = ctx.symptr_rule(npt, syms)
    w_orc, nirr_o = orc.symptr_rule(npt, syms)
    assert nirr == nirr_o and np.array_equal(w_host, w_orc)
    Ra = L.DeviceRule(ctx, S, npt, wsym=w_host)
    Rb = L.DeviceRule(ctx, S, npt, syms=syms)
    assert len(Ra) == len(Rb) == nirr == Rb.nirr_total
    Ha, ka, wa = Ra.copy_out()
    Hb, kb, wb = Rb.copy_out()
    assert np.array_equal(ka, kb) and np.array_equal(wa, wb) and np.array_equal(Ha, Hb)
    assert wb.sum() == npt ** 3
    z = np.array([0.3 + 0.05j, -0.7 + 0.2j])
    assert np.array_equal(Ra.resolvent_sum(z), Rb.resolvent_sum(z))
    parts = [L.DeviceRule(ctx, S, npt, syms=syms, k3_lo=r, k3_stride=3) for r in range(3)]
    assert all(p.nirr_total == nirr for p in parts) and sum(len(p) for p in parts) == nirr
    assert rel(sum(p.resolvent_sum(z) for p in parts), Ra.resolvent_sum(z)) < 1e-13
    # inversion-only group and the trivial group
    inv = np.array([np.eye(3), -np.eye(3)], dtype=np.int32)
    Rc = L.DeviceRule(ctx, S, npt, syms=inv)
    wi, ni = orc.symptr_rule(npt, inv)
    assert len(Rc) == ni
    assert rel(Rc.resolvent_sum(z), L.DeviceRule(ctx, S, npt).resolvent_sum(z)) < 1e-12


@pytest.mark.parametrize("n", [2, 7, 31, 32, 33, 40, 63, 64])
def test_eig_algorithms_agree_with_lapack(ctx, n):
    """Householder tridiagonalisation + implicit QL (default) and cyclic Jacobi (option) against LAPACK zheev on
    H(k) of a synthetic Wannier Hamiltonian, incl. a degenerate spectrum (H = 0 -> all eigenvalues 0; diagonal H)."""
    H, lo = ab.synthetic.wannier_hamiltonian(n, 1)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    R = L.DeviceRule(ctx, S, 5)
    Hk, _, _ = R.copy_out()
    ref = np.linalg.eigvalsh(np.moveaxis(Hk, 2, 0))
    ev_fast = R.eigvals()
    ctx.set_option(L.OPT_EIG_ALGO, 1)
    try:
        ev_jac = R.eigvals()
        sj = R.eig_sum(L.EIG_FERMI_ENERGY, (0.1, 0.3))
    finally:
        ctx.set_option(L.OPT_EIG_ALGO, 0)
    assert rel(ev_fast, ref) < 1e-13 and rel(ev_jac, ref) < 1e-12
    assert np.all(np.diff(ev_fast, axis=1) >= 0)
    assert abs(R.eig_sum(L.EIG_FERMI_ENERGY, (0.1, 0.3)) - sj) < 1e-11 * abs(sj)
    Hd = np.zeros_like(H)
    Hd[:, :, -lo[0], -lo[1], -lo[2]] = np.diag(np.repeat(np.arange((n + 1) // 2), 2)[:n])
    Rd = L.DeviceRule(ctx, L.DeviceSeries(ctx, Hd, lo, (1.0,) * 3), 3)
    want = np.tile(np.repeat(np.arange((n + 1) // 2), 2)[:n].astype(float), (27, 1))
    assert np.max(np.abs(Rd.eigvals() - want)) <= 1e-13 * max(1.0, want.max())      # bisection (small batches): to rounding
    ctx.set_option(L.OPT_EIG_ALGO, 3)
    try:
        assert np.array_equal(Rd.eigvals(), want)                                       # QL leaves a diagonal matrix untouched
    finally:
        ctx.set_option(L.OPT_EIG_ALGO, 0)


def test_generic_and_batch_integrands_on_device(ctx, orc, svo):
    """S3/S4 seams for user integrands evaluated on the host: H(k) at the rule nodes (abz_rule_copy_out) or at IAI panel
    nodes (abz_nest_eval_h) comes from the device; a plain callable and its BatchIntegrand form reproduce the native
    DOS integrand: same value to 1e-10 and the same adaptive evaluation count."""
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    ibz = ab.load_bz(ab.CubicSymIBZ(), A)
    eta, omega = 0.1, 12.5

    def dos(x, eta, omega):
        return -np.imag(np.trace(np.linalg.inv(complex(omega, eta) * np.eye(3) - x.s))) / np.pi

    def dos_batch(y, x, eta, omega):
        assert x.s.shape == (len(y), 3, 3) and x.x.shape == (len(y), 3)
        y[:] = -np.imag(np.trace(np.linalg.inv(complex(omega, eta) * np.eye(3) - x.s), axis1=1, axis2=2)) / np.pi

    for alg, kw in ((ab.IAI(), dict(abstol=5e-1)), (ab.PTR(npt=12), {})):
        ref = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.dos_integrand, fs, eta), ibz, omega), ab.EvalCounter(alg), **kw)
        for integrand in (ab.FourierIntegrand(dos, fs, eta), ab.FourierIntegrand(ab.BatchIntegrand(dos_batch, dtype=np.float64, max_batch=4096), fs, eta)):
            sol = ab.solve(ab.IntegralProblem(integrand, ibz, omega), ab.EvalCounter(alg), **kw)
            assert sol.numevals == ref.numevals
            assert abs(sol.u - ref.u) <= 1e-10 * abs(ref.u)
            assert isinstance(sol.u, float)


@pytest.mark.parametrize("n,N", [(1, 8), (3, 7), (8, 6), (32, 4), (64, 3)])
def test_matrix_valued_green_function_sum(ctx, orc, n, N):
    """abz_rule_resolvent_matrix_sum (docs/src/examples.md:20,90, gloc_integrand = inv(...)): full matrices against LAPACK
    inverses of the oracle's H(k), with and without a self-energy, streamed and materialised, IBZ shards adding up;
    the trace of the matrix sum equals the trace kernel's result."""
    H, lo = ab.synthetic.wannier_hamiltonian(n, 1, cubic=True)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    So = orc.Series(H, lo)
    z = np.array([0.3 + 0.25j, -0.6 + 0.4j])
    rng = np.random.default_rng(n)
    sig = 0.1 * (rng.standard_normal((n, n, 2)) + 1j * rng.standard_normal((n, n, 2)))
    R = L.DeviceRule(ctx, S, N)
    Hk = np.moveaxis(orc.grid_eval_full(So, N).reshape(n, n, -1, order="F"), 2, 0)
    for sg in (None, sig):
        ref = np.array([np.linalg.inv(zz * np.eye(n) - Hk - (0 if sg is None else sg[:, :, w])).sum(axis=0) for w, zz in enumerate(z)])
        got = R.resolvent_matrix_sum(z, sigma=sg)
        assert got.shape == (2, n, n) and rel(got, ref) < 1e-11
        assert rel(np.trace(got, axis1=1, axis2=2), R.resolvent_sum(z, sigma=sg)) < 1e-11
    R.materialize()
    assert rel(R.resolvent_matrix_sum(z), np.array([np.linalg.inv(zz * np.eye(n) - Hk).sum(axis=0) for zz in z])) < 1e-11
    syms = np.array(ab.cube_automorphisms(3), dtype=np.int32)
    parts = [L.DeviceRule(ctx, S, N, syms=syms, k3_lo=r, k3_stride=2) for r in range(2)]
    assert rel(sum(p.resolvent_matrix_sum(z) for p in parts), np.array([np.linalg.inv(zz * np.eye(n) - Hk).sum(axis=0) for zz in z])) < 1e-11
    # through the public API: SymRep hook on the IBZ == full BZ
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=n)
    p = {"omega": 0.3}
    a = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_integrand, fs, eta=0.25), ab.load_bz(ab.FBZ(), np.eye(3)), p), ab.PTR(npt=N)).u
    b = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.GlocIntegrand(symmetrize=lambda bz, x: bz.nsyms * x), fs, eta=0.25),
                                    ab.load_bz(ab.CubicSymIBZ(), np.eye(3)), p), ab.PTR(npt=N)).u
    assert rel(b, a) < 1e-11


@pytest.mark.parametrize("n,rmax,N", [(4, 1, 6), (7, 2, 5), (32, 1, 5), (33, 1, 4), (64, 1, 3)])
def test_frequency_sweep_from_one_tridiagonalisation(ctx, orc, n, rmax, N):
    """ABZ_OPT_RESOLVENT_ALGO = 3: tr (z - H(k))^-1 for all frequencies from one Householder reduction per k
    (p'(z)/p(z) of the tridiagonal form) against the oracle's pivoted-LU resolvent, sums and per-node values; a matrix
    self-energy silently takes the LU path."""
    H, lo = ab.synthetic.wannier_hamiltonian(n, rmax)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    So = orc.Series(H, lo)
    ext = ab.synthetic.band_extent(H)
    z = np.linspace(-0.3 * ext, 0.3 * ext, 37) + 1j * 0.01 * ext
    R = L.DeviceRule(ctx, S, N)
    ref = orc.ptr_sum(So, N, z)
    rng = np.random.default_rng(n)
    sig = 0.1 * (rng.standard_normal((n, n, z.size)) + 1j * rng.standard_normal((n, n, z.size)))
    kp = rng.random((11, 3))
    ctx.set_option(L.OPT_RESOLVENT_ALGO, 3)
    try:
        assert rel(R.resolvent_sum(z, scale=1 / N ** 3), ref) < 1e-11
        R.materialize()
        a = R.resolvent_sum(z, scale=1 / N ** 3)
        assert rel(a, ref) < 1e-11 and np.array_equal(a, R.resolvent_sum(z, scale=1 / N ** 3))
        assert rel(R.resolvent_sum(z, sigma=sig, scale=1 / N ** 3), orc.ptr_sum(So, N, z, sigma=sig)) < 1e-11
        y = S.points_resolvent(kp, z)
    finally:
        ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)
    assert rel(y, orc.resolvent_trace_batch(orc.eval_points(So, kp), z)) < 1e-11


@pytest.mark.parametrize("n", [2, 5, 8, 9, 16, 23, 24, 31, 32])
def test_register_resident_tridiagonalisation(ctx, n):
    """norb <= 32: the warp-per-matrix register-resident Householder kernel (default) and the shared-memory kernel
    (ABZ_OPT_EIG_ALGO = 2) give the same spectrum as LAPACK, including zero-padded sizes that are not multiples of 8"""
    H, lo = ab.synthetic.wannier_hamiltonian(n, 1)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    R = L.DeviceRule(ctx, S, 6)
    Hk, _, _ = R.copy_out()
    ref = np.linalg.eigvalsh(np.moveaxis(Hk, 2, 0))
    ev = R.eigvals()
    ctx.set_option(L.OPT_EIG_ALGO, 2)
    try:
        ev2 = R.eigvals()
    finally:
        ctx.set_option(L.OPT_EIG_ALGO, 0)
    assert rel(ev, ref) < 1e-13 and rel(ev2, ref) < 1e-13


@pytest.mark.parametrize("n", [6, 40])
def test_frequency_sweep_rejects_non_hermitian_h(ctx, orc, n):
    """the sweep path is only valid for Hermitian H(k): a series without H_{-R} = H_R^dagger is refused (ValueError), while the
    LU path (Julia's inv) handles it"""
    rng = np.random.default_rng(n)
    H = rng.standard_normal((n, n, 3, 3, 3)) + 1j * rng.standard_normal((n, n, 3, 3, 3))
    S = L.DeviceSeries(ctx, H, (-1, -1, -1), (1.0,) * 3)
    z = np.array([0.3 + 2.5j, -0.4 + 3.0j])
    R = L.DeviceRule(ctx, S, 4)
    ref = orc.ptr_sum(orc.Series(H, (-1, -1, -1)), 4, z)
    assert rel(R.resolvent_sum(z, scale=1 / 64), ref) < 1e-10
    ctx.set_option(L.OPT_RESOLVENT_ALGO, 3)
    try:
        with pytest.raises(ValueError, match="not Hermitian"):
            R.resolvent_sum(z)
    finally:
        ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)
    assert rel(R.resolvent_sum(z, scale=1 / 64), ref) < 1e-10       # the context stays usable


@pytest.mark.parametrize("seed", range(12))
def test_iai_device_vs_oracle_on_random_problems(ctx, orc, seed):
    """random Hermitian series (1-3 orbitals, 1-3 dimensions, cubic / tetrahedral limits, real DOS / complex trace values):
    the device engine with one warp per innermost integral takes the oracle recursion's decisions (same numevals) and agrees
    to 1e-10"""
    rng = np.random.default_rng(1000 + seed)
    ndim, n = 1 + seed % 3, 1 + (seed // 3) % 3
    lkind, vkind = seed % 2, (seed // 2) % 2
    M = (3,) * ndim
    c = rng.standard_normal((n, n) + M) + 1j * rng.standard_normal((n, n) + M)
    rev = (slice(None), slice(None)) + (slice(None, None, -1),) * ndim
    c = 0.5 * (c + np.conj(np.swapaxes(c[rev], 0, 1)))
    fs = ab.FourierSeries(c, period=1.0, lo=(-1,) * ndim, norb=n)
    So = orc.Series(c.reshape((n, n) + M + (1,) * (3 - ndim)), (-1,) * ndim + (0,) * (3 - ndim))
    z = complex(rng.uniform(-1, 1), [0.3, 0.1, 0.05][seed % 3])
    tol = [3e-1, 3e-2, 3e-3][(seed // 4) % 3]
    la = [0.5] * ndim if lkind else [0.0] * ndim
    lb = None if lkind else [1.0] * ndim
    Io, Eo, neo = orc.iai(So, ndim, lkind, la, lb, vkind=vkind, z=z, atol=tol)
    nest = L.DeviceNest(ctx, fs.device(ctx), ndim, 64 if ndim == 3 else 0, 2048 if ndim >= 2 else 0)
    for leaves in (True, False):
        I, E, ne, rounds, launches = nest.iai_solve(lkind, la, lb, L.F_RESOLVENT_TRACE, vkind, z, None, None, tol, 0.0, 2 ** 62, device_leaves=leaves)
        assert ne == neo
        assert abs(I - Io) <= 1e-10 * max(abs(Io), 1e-12)


@pytest.mark.parametrize("leaves", [True, False])
def test_iai_rounds_in_flight_do_not_change_results(ctx, svo, leaves):
    """ABZ_OPT_IAI_LANES: 1, 2, 4 and 7 rounds in flight (one stream each) give bit-identical integrals, error estimates and
    numevals (every 1-D integral sees its own evaluations in its own order); more lanes mean more, smaller rounds"""
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    nest = L.DeviceNest(ctx, fs.device(ctx), 3, 64, 4096)
    res = []
    try:
        for lanes in (1, 2, 4, 7):
            ctx.set_option(L.OPT_IAI_LANES, lanes)
            I, E, ne, rounds, launches = nest.iai_solve(1, [0.5] * 3, None, L.F_RESOLVENT_TRACE, 1, complex(12.5, 0.02), None, None, 1e-3, 0.0,
                                                        2 ** 62, device_leaves=leaves)
            res.append((I, E, ne, rounds))
    finally:
        ctx.set_option(L.OPT_IAI_LANES, 4)
    assert all(r[:3] == res[0][:3] for r in res)
    assert res[2][3] > res[0][3]


def test_iai_leaf_heap_overflow_falls_back_to_host_panels(ctx, orc, svo):
    """an innermost integral that outgrows the device segment heap (shrunk here to 2 segments in total) is redone with
    host-driven panels: same numevals and value as the oracle, no error"""
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    S = orc.Series(H, lo)
    Io, Eo, neo = orc.iai(S, 3, 1, [0.5] * 3, vkind=1, z=complex(12.975161, 2e-3), atol=1e-3)
    nest = L.DeviceNest(ctx, fs.device(ctx), 3, 64, 2048)
    ctx.set_option(L.OPT_IAI_LEAF_SPILL, -2)
    try:
        I, E, ne, rounds, launches = nest.iai_solve(1, [0.5] * 3, None, L.F_RESOLVENT_TRACE, 1, complex(12.975161, 2e-3), None, None, 1e-3, 0.0,
                                                    2 ** 62, device_leaves=True)
    finally:
        ctx.set_option(L.OPT_IAI_LEAF_SPILL, 1024)
    assert ne == neo and abs(I.real - Io.real) <= 1e-10 * abs(Io.real)
    # the fallback really happened: its round count = the aborted device-leaf rounds + all rounds of the host-panel solve
    _, _, _, rounds_host, _ = nest.iai_solve(1, [0.5] * 3, None, L.F_RESOLVENT_TRACE, 1, complex(12.975161, 2e-3), None, None, 1e-3, 0.0,
                                             2 ** 62, device_leaves=False)
    assert rounds > rounds_host


def test_symptr_two_phase_path_on_a_large_grid(ctx, orc):
    """npt^3 >= 2^22 takes the two-phase (filter + compaction, then filter + orbit-stabiliser weights) version of symptr_rule: identical weight
    array to the oracle's for the cubic group and for a group given in an unusual order (identity last, inversion first)"""
    npt = 165
    syms = np.array(ab.cube_automorphisms(3), dtype=np.int32)
    w_o, n_o = orc.symptr_rule(npt, syms)
    w_d, n_d = ctx.symptr_rule(npt, syms)
    assert n_d == n_o and np.array_equal(w_d, w_o) and int(w_d.sum()) == npt ** 3
    perm = np.concatenate([syms[::-1][:1], syms[1:-1][::-1], syms[:1]])
    w_p, n_p = ctx.symptr_rule(npt, perm)
    assert n_p == n_o and np.array_equal(w_p, w_o)
    H, lo = ab.synthetic.wannier_hamiltonian(2, 1, cubic=True)
    R = L.DeviceRule(ctx, L.DeviceSeries(ctx, H, lo, (1.0,) * 3), npt, syms=syms)
    assert len(R) == n_o and R.copy_out()[2].sum() == npt ** 3
    # a list that is NOT a group is refused (AutoSymPTR's scan is order-dependent then; load_bz only produces groups)
    for bad in (syms[1:21], syms[:20], np.concatenate([syms, syms[:1]])):
        with pytest.raises(ValueError, match="group"):
            ctx.symptr_rule(12, bad)
        with pytest.raises(ValueError, match="group"):
            L.DeviceRule(ctx, L.DeviceSeries(ctx, H, lo, (1.0,) * 3), 12, syms=bad)
