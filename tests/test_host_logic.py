"""Host control flow (BZ conventions, tolerance rescaling, AutoPTR loop, IAI engine, sharding) exercised on CPU
through the oracle-backed test double, against the reference's own tests (test/fourier.jl, test/brillouin.jl).
CPU only; the same tests run against the device backend in test_gpu_parity.py."""
import json
import math
import os
import subprocess
import sys

import numpy as np
import pytest

import autobz_b200 as ab
from oracle_backend import OracleBackend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lattice_series(d):
    c, lo = ab.synthetic.integer_lattice(d)
    return ab.FourierSeries(c[0, 0], period=1.0, lo=lo)


def test_bz_construction():
    # test/brillouin.jl:7-31
    for d in (1, 2, 3):
        A = np.eye(d)
        assert ab.load_bz(ab.FBZ(), A).nsyms == 1
        assert ab.load_bz(ab.InversionSymIBZ(), A).nsyms == 2 ** d
        assert ab.load_bz(ab.CubicSymIBZ(), A).nsyms == math.factorial(d) * 2 ** d
        assert isinstance(ab.load_bz(ab.FBZ(), A).lims, ab.CubicLimits)
        assert isinstance(ab.load_bz(ab.CubicSymIBZ(), A).lims, ab.TetrahedralLimits)
    bz = ab.load_bz(ab.FBZ(2), 2 * np.pi * np.eye(2))
    assert np.allclose(bz.B, np.eye(2))
    with pytest.raises(ValueError):
        ab.load_bz(ab.FBZ(), np.eye(2), np.eye(2))          # non-reciprocal bases
    syms = ab.cube_automorphisms(3)
    assert len({S.tobytes() for S in syms}) == 48


def test_tetrahedral_limits_volume():
    lims = ab.TetrahedralLimits([0.5, 0.5, 0.5])
    assert lims.segments() == (0.0, 0.5)
    l2 = lims.fix(0.25)
    assert l2.segments() == (0.0, 0.25)
    assert l2.fix(0.1).segments() == (0.0, 0.1)


def test_schedule_defaults():
    from autobz_b200.algorithms import monkhorst_pack_schedule
    assert monkhorst_pack_schedule(1.0, 50, 1000, 6.0, math.log(10)) == (50, 3)
    assert monkhorst_pack_schedule(0.01, 50, 1000, 6.0, math.log(10)) == (600, 231)


@pytest.mark.parametrize("d", [1, 2, 3])
@pytest.mark.parametrize("bzkind", ["fbz", "inv"])
def test_fourier_jl_algorithms(d, bzkind):
    """test/fourier.jl:40-56: for alg in (IAI, PTR, AutoPTR) x (plain, EvalCounter):
    integral of 1.3*H(k) + 1 over the BZ of A = I(d) equals (2 pi)^d to atol 1e-6 (abstol=1e-6, reltol=0)."""
    vol = (2 * np.pi) ** d
    s = lattice_series(d)
    bz = ab.load_bz(ab.FBZ() if bzkind == "fbz" else ab.InversionSymIBZ(), np.eye(d))
    integrand = ab.FourierIntegrand(ab.AffineTraceIntegrand(), s, 1.3, b=1.0)
    prob = ab.IntegralProblem(integrand, bz)
    be = OracleBackend()
    for alg in (ab.IAI(), ab.PTR(), ab.AutoPTR()):
        for counter in (False, True):
            new_alg = ab.EvalCounter(alg) if counter else alg
            solver = ab.IntegralSolver(prob, new_alg, reltol=0, abstol=1e-6, backend=be)
            assert abs(solver() - vol) < 1e-6
            sol = solver.solve_p(None)
            assert (sol.numevals > 0) == counter


def test_evalcounter_counts():
    s = lattice_series(3)
    bz = ab.load_bz(ab.FBZ(), np.eye(3))
    ibz = ab.load_bz(ab.CubicSymIBZ(), np.eye(3))
    f = ab.FourierIntegrand(ab.AffineTraceIntegrand(), s, 0.0, b=1.0)     # constant integrand
    be = OracleBackend()
    # constant integrand: every GK panel converges at once -> 15^3 evaluations (test/brillouin.jl:96 per level)
    sol = ab.solve(ab.IntegralProblem(f, bz), ab.EvalCounter(ab.IAI()), abstol=1e-8, backend=be)
    assert sol.numevals == 15 ** 3 and abs(sol.u - (2 * np.pi) ** 3) < 1e-9
    sol = ab.solve(ab.IntegralProblem(f, ibz), ab.EvalCounter(ab.IAI()), abstol=1e-8, backend=be)
    assert sol.numevals == 15 ** 3 and abs(sol.u - (2 * np.pi) ** 3) < 1e-9
    # PTR: numevals = number of (irreducible) nodes
    sol = ab.solve(ab.IntegralProblem(f, bz), ab.EvalCounter(ab.PTR(npt=10)), backend=be)
    assert sol.numevals == 1000
    sol = ab.solve(ab.IntegralProblem(f, ibz), ab.EvalCounter(ab.PTR(npt=10)), backend=be)
    assert sol.numevals == 56
    # AutoPTR default schedule 50, 53 (converges immediately for a constant)
    sol = ab.solve(ab.IntegralProblem(f, ibz), ab.EvalCounter(ab.AutoPTR()), abstol=1e-8, backend=be)
    w50 = 50 * 51 * 52 // 6 if False else None
    assert sol.numevals > 0 and abs(sol.u - (2 * np.pi) ** 3) < 1e-9


@pytest.mark.parametrize("lims", ["cubic", "tetra"])
def test_iai_engine_matches_recursive_oracle(orc, svo, lims):
    """The level-synchronous engine must reproduce the sequential recursion decision for decision:
    identical numevals and integrals equal to rounding (same oracle arithmetic on both sides)."""
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    S = orc.Series(H, lo)
    be = OracleBackend()
    eta, omega = 0.05, 12.5
    if lims == "cubic":
        bz = ab.load_bz(ab.FBZ(), A)
        Io, Eo, neo = orc.iai(S, 3, 0, [0.0] * 3, [1.0] * 3, vkind=1, z=complex(omega, eta), atol=3e-2)
        mult = abs(np.linalg.det(bz.B))
        abstol = 3e-2 * mult
    else:
        bz = ab.load_bz(ab.CubicSymIBZ(), A)
        Io, Eo, neo = orc.iai(S, 3, 1, [0.5] * 3, vkind=1, z=complex(omega, eta), atol=1e-3)
        mult = abs(np.linalg.det(bz.B)) * 48
        abstol = 1e-3 * mult
    f = ab.FourierIntegrand(ab.dos_integrand, fs, eta)
    sol = ab.solve(ab.IntegralProblem(f, bz, omega), ab.EvalCounter(ab.IAI()), abstol=abstol, backend=be)
    assert sol.numevals == neo
    assert abs(sol.u - mult * Io.real) <= 1e-12 * abs(sol.u)


def test_ptr_and_autoptr_svo_vs_oracle(orc, svo):
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    S = orc.Series(H, lo)
    be = OracleBackend()
    bz = ab.load_bz(ab.FBZ(), A)
    ibz = ab.load_bz(ab.CubicSymIBZ(), A)
    j = abs(np.linalg.det(bz.B))
    f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=0.01)
    ref = orc.ptr_sum(S, 24, [11.0 + 0.01j])[0] * j
    for dom in (bz, ibz):
        u = ab.solve(ab.IntegralProblem(f, dom, {"omega": 11.0}), ab.PTR(npt=24), backend=be).u
        assert abs(u - ref) < 1e-11 * abs(ref)
    # batchsolve == serial loop (test/brillouin.jl:98-111)
    solver = ab.IntegralSolver(f, ibz, ab.PTR(npt=12), backend=be)
    ws = [11.0, 12.0, 13.0]
    assert np.allclose(ab.batchsolve(solver, [{"omega": w} for w in ws]), [solver(omega=w) for w in ws], rtol=1e-13)
    # AutoPTR outside the band converges on the first pair of grids
    sol = ab.solve(ab.IntegralProblem(f, ibz, {"omega": 11.0}), ab.EvalCounter(ab.AutoPTR(nmin=12, a=1.0)), abstol=1e-6, backend=be)
    assert abs(sol.u - ref) < 1e-6


def test_generic_python_integrand_host_path():
    s = lattice_series(2)
    bz = ab.load_bz(ab.FBZ(), np.eye(2))
    f = ab.FourierIntegrand(lambda x, a, b=0.0: a * x.s + b, s, 1.3, b=1.0)
    u = ab.solve(ab.IntegralProblem(f, bz), ab.PTR(npt=8), backend=OracleBackend()).u
    assert abs(u - (2 * np.pi) ** 2) < 1e-10


_GLOO_SCRIPT = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle")); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch.distributed as dist
import autobz_b200 as ab
from oracle_backend import OracleBackend
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
shard = ab.Shard(rank, 2, ab.torch_allreduce())
H, lo = ab.synthetic.wannier_hamiltonian(3, 1, cubic=True)
fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=2.0)
out = []
for dom in (ab.load_bz(ab.FBZ(), np.eye(3)), ab.load_bz(ab.CubicSymIBZ(), np.eye(3))):
    for alg in (ab.PTR(npt=9), ab.EvalCounter(ab.AutoPTR(nmin=6, nmax=40))):
        s2 = ab.solve(ab.IntegralProblem(f, dom, {{"omega": 0.3}}), alg, abstol=1e-5, backend=OracleBackend(), shard=shard)
        s1 = ab.solve(ab.IntegralProblem(f, dom, {{"omega": 0.3}}), alg, abstol=1e-5, backend=OracleBackend())
        out.append((abs(s2.u - s1.u) / abs(s1.u), s2.numevals == s1.numevals))
# IAI through the Python-driven engine (what generic / batch host integrands use): outermost panel nodes dealt to the ranks,
# one allreduce per outer step; bit-identical to one rank, total numevals over the ranks
def generic(x, eta, omega):
    return np.trace(np.linalg.inv(complex(omega, eta) * np.eye(3) - x.s))
ibz = ab.load_bz(ab.CubicSymIBZ(), np.eye(3))
for fi, p in ((f, {{"omega": 0.3}}), (ab.FourierIntegrand(generic, fs, 2.0), 0.3)):
    s2 = ab.solve(ab.IntegralProblem(fi, ibz, p), ab.EvalCounter(ab.IAI()), abstol=1e-4, backend=OracleBackend(), shard=shard)
    s1 = ab.solve(ab.IntegralProblem(fi, ibz, p), ab.EvalCounter(ab.IAI()), abstol=1e-4, backend=OracleBackend())
    out.append((abs(s2.u - s1.u) / abs(s1.u), s2.numevals == s1.numevals and s2.u == s1.u))
ok = all(e < 1e-13 and c for e, c in out)
print("RANK", rank, "OK" if ok else "FAIL", out)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_two_rank_gloo_sharding(tmp_path):
    """world_size 2 over gloo: k3 planes sharded (contiguous for FBZ, round-robin for IBZ), one allreduce per rule,
    identical convergence decisions and results equal to the single-rank solve to rounding."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "gloo2.py"
    script.write_text(_GLOO_SCRIPT.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


@pytest.mark.parametrize("d", [1, 2, 3])
def test_generic_and_batch_integrands_match_native(d):
    """test/fourier.jl:24-56: a plain user integrand f(x::FourierValue, a; b) = a*x.s + b and its BatchIntegrand form
    f!(y, x, a; b) (src/batch.jl:10-38, the S3 seam) give the same integrals - and, for IAI, the same adaptive
    evaluation counts - as the device-native affine integrand, for IAI / PTR / AutoPTR."""
    vol = (2 * np.pi) ** d
    s = lattice_series(d)
    be = OracleBackend()

    def f(x, a, b=0.0):
        return a * x.s + b

    calls = []

    def fb(y, x, a, b=0.0):
        calls.append(len(y))
        assert x.x.shape == (len(y), d) and x.s.shape == (len(y),)
        y[:] = a * x.s + b

    for bz in (ab.load_bz(ab.FBZ(), np.eye(d)), ab.load_bz(ab.InversionSymIBZ(), np.eye(d))):
        for alg in (ab.IAI(), ab.PTR(npt=12), ab.AutoPTR()):
            ref = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.AffineTraceIntegrand(), s, 1.3, b=1.0), bz), ab.EvalCounter(alg),
                           reltol=0, abstol=1e-6, backend=be)
            for integrand in (ab.FourierIntegrand(f, s, 1.3, b=1.0), ab.FourierIntegrand(ab.BatchIntegrand(fb, max_batch=100), s, 1.3, b=1.0)):
                sol = ab.solve(ab.IntegralProblem(integrand, bz), ab.EvalCounter(alg), reltol=0, abstol=1e-6, backend=be)
                assert abs(sol.u - vol) < 1e-6 and abs(sol.u - ref.u) < 1e-9
                assert sol.numevals == ref.numevals
    assert calls and max(calls) <= 100
    assert ab.BatchIntegralFunction is ab.BatchIntegrand and ab.FourierIntegralFunction is ab.FourierIntegrand
    with pytest.raises(ValueError):
        ab.BatchIntegrand(fb, max_batch=0)


def test_generic_integrand_sees_full_k_points_in_iai():
    """the innermost closure passes FourierValue(limit_iterate(lims, state, x), H(x)) (src/fourier.jl:454): the integrand
    sees the full k = (k1, k2, k3).  Integrate k1 + 2 k2 + 4 k3 over the unit cube = 3.5 (per unit |det B|)."""
    s = lattice_series(3)
    bz = ab.load_bz(ab.FBZ(), 2 * np.pi * np.eye(3))           # B = I  =>  j = 1
    f = ab.FourierIntegrand(lambda x: x.x[0] + 2 * x.x[1] + 4 * x.x[2] + 0 * x.s.real, s)
    sol = ab.solve(ab.IntegralProblem(f, bz), ab.EvalCounter(ab.IAI()), abstol=1e-10, backend=OracleBackend())
    assert abs(sol.u - 3.5) < 1e-12 and sol.numevals == 15 ** 3 and isinstance(sol.u, float)


def test_matrix_valued_gloc_integrand_and_symrep(orc):
    """docs/src/examples.md:20,90: gloc_integrand returns inv(complex(omega, eta) I - h_k.s), a matrix.  On a symmetric BZ the
    reference symmetrises with the integrand's SymRep (src/brillouin.jl:86-107); with UnknownRep it warns and repeats the
    calculation on the full BZ (src/brillouin.jl:348-353)."""
    n = 3
    H, lo = ab.synthetic.wannier_hamiltonian(n, 1, cubic=True)      # orbital action trivial: G(Sk) = G(k)
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=n)
    be = OracleBackend()
    fbz, ibz = ab.load_bz(ab.FBZ(), np.eye(3)), ab.load_bz(ab.CubicSymIBZ(), np.eye(3))
    p = {"omega": 0.3}
    ref = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_integrand, fs, eta=0.4), fbz, p), ab.PTR(npt=8), backend=be).u
    assert ref.shape == (n, n)
    tr = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=0.4), fbz, p), ab.PTR(npt=8), backend=be).u
    assert abs(np.trace(ref) - tr) < 1e-12 * abs(tr)
    # a SymRep for which the representation is trivial: symmetrize(bz, x) = nsyms x
    sym = ab.GlocIntegrand(symmetrize=lambda bz, x: bz.nsyms * x)
    got = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(sym, fs, eta=0.4), ibz, p), ab.PTR(npt=8), backend=be).u
    assert np.max(np.abs(got - ref)) < 1e-12 * np.max(np.abs(ref))
    # UnknownRep: warning + full-BZ recomputation
    with pytest.warns(UserWarning, match="symmetry representation is unknown"):
        got = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_integrand, fs, eta=0.4), ibz, p), ab.PTR(npt=8), backend=be).u
    assert np.max(np.abs(got - ref)) < 1e-13 * np.max(np.abs(ref))
    # AutoPTR converges in the Frobenius norm; batchsolve returns one matrix per frequency
    sol = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(sym, fs, eta=0.4), ibz, p), ab.EvalCounter(ab.AutoPTR(nmin=4, a=0.4)), abstol=1e-6, backend=be)
    assert sol.u.shape == (n, n) and sol.resid <= 1e-6 and sol.numevals > 0
    solver = ab.IntegralSolver(ab.FourierIntegrand(ab.gloc_integrand, fs, eta=0.4), fbz, ab.PTR(npt=6), backend=be)
    G = ab.batchsolve(solver, [{"omega": w} for w in (0.0, 0.3)])
    assert G.shape == (2, n, n)
    # IAI on the matrix-valued integrand (the nest is generic in the value type, src/fourier.jl:452-456; norm = Frobenius):
    # same integral as PTR within the tolerance, SymRep applied to the IBZ value, trace consistent with the scalar integrand
    # (eta = 4 and abstol = 0.1 on a value of size 50 keep the Python engine + CPU oracle to ~1e5 evaluations)
    gi = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_integrand, fs, eta=4.0), fbz, p), ab.EvalCounter(ab.IAI()), abstol=0.1, backend=be)
    assert gi.u.shape == (n, n) and gi.numevals > 0 and gi.resid <= 0.1
    fine = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_integrand, fs, eta=4.0), fbz, p), ab.PTR(npt=32), backend=be).u
    assert np.max(np.abs(gi.u - fine)) < 0.1
    gt = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=4.0), fbz, p), ab.IAI(), abstol=0.1, backend=be).u
    assert abs(np.trace(gi.u) - gt) < 0.2
    gs = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(sym, fs, eta=4.0), ibz, p), ab.IAI(), abstol=0.1, backend=be).u
    assert np.max(np.abs(gs - fine)) < 0.1
    # a 1 x 1 "matrix" takes exactly the decisions of the scalar integrand (docs/src/examples.md:90-105: IAI on gloc_integrand of a
    # scalar series, 2-d): same numevals, same value, and the value the reference's docs print
    c2 = np.zeros((1, 1, 3, 3)); c2[0, 0, 0, 1] = c2[0, 0, 2, 1] = c2[0, 0, 1, 0] = c2[0, 0, 1, 2] = 0.5
    h2 = ab.FourierSeries(c2, period=1.0, lo=(-1, -1), norb=1)
    bz2 = ab.load_bz(ab.FBZ(2), np.eye(2) * 2 * np.pi)
    sm = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_integrand, h2, eta=0.1), bz2, {"omega": 0.0}), ab.EvalCounter(ab.IAI()), abstol=1e-3, backend=be)
    ss = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_trace_integrand, h2, eta=0.1), bz2, {"omega": 0.0}), ab.EvalCounter(ab.IAI()), abstol=1e-3, backend=be)
    assert sm.numevals == ss.numevals and abs(sm.u[0, 0] - ss.u) <= 1e-13 * abs(ss.u)
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))["reference_known_answers"]
    assert abs(ss.u - complex(*gold["docs/src/examples.md:105 (IAI abstol=1e-3, 2-D gloc on FBZ(2), eta=0.1, omega=0)"])) < 1e-3


def test_batchsolve_log_archive(tmp_path, orc, svo):
    """ext/HDF5Ext.jl:116-158: batchsolve into an archive with datasets I, E, t, retcode, numevals and the parameters"""
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    bz = ab.load_bz(ab.CubicSymIBZ(), A)
    solver = ab.IntegralSolver(ab.FourierIntegrand(ab.dos_integrand, fs, 0.1), bz, ab.EvalCounter(ab.PTR(npt=8)), backend=OracleBackend())
    ws = [12.0, 12.5, 13.0]
    out = ab.batchsolve_log(tmp_path / "sweep.npz", solver, ws)
    arc = np.load(tmp_path / "sweep.npz")
    assert set(arc.files) >= {"I", "E", "t", "retcode", "numevals", "args/1"}
    assert np.array_equal(arc["I"], out) and np.array_equal(arc["args/1"], ws)
    assert np.all(arc["retcode"] == 1) and np.all(arc["numevals"] == len(solver.cache.cacheval["rule"])) and np.all(np.isnan(arc["E"]))
    solver2 = ab.IntegralSolver(ab.FourierIntegrand(ab.gloc_trace_integrand, fs), bz, ab.PTR(npt=6), backend=OracleBackend())
    ab.batchsolve_log(tmp_path / "kw.npz", solver2, [{"eta": 0.1, "omega": w} for w in ws])
    arc = np.load(tmp_path / "kw.npz")
    assert np.array_equal(arc["kwargs/omega"], ws) and arc["I"].dtype == np.complex128


@pytest.mark.parametrize("d", [1, 2, 3])
def test_nested_batch_integrand_equals_plain(d):
    """test/fourier.jl:24-37: prob1 = FourierIntegrand(p, s), prob2 = FourierIntegrand(p, ws, nest) with
    nest = NestedBatchIntegrand(ntuple(n -> deepcopy(p), nouter), ...): solve(prob1, alg).u ~ solve(prob2, alg).u for
    NestedQuad(AuxQuadGKJL()) on CubicLimits and MonkhorstPack on the basis."""
    s = lattice_series(d)
    be = OracleBackend()

    def p(x, a, b=0.0):
        return a * x.s + b

    nest = ab.NestedBatchIntegrand(tuple(p for _ in range(3)), dtype=np.complex128, max_batch=50)
    f1 = ab.FourierIntegrand(p, s, 1.3, b=4.2)
    f2 = ab.FourierIntegrand(p, s, 1.3, nest=nest, b=4.2)
    for alg, dom in ((ab.NestedQuad(ab.AuxQuadGKJL()), ab.CubicLimits([0.0] * d, [1.0] * d)), (ab.MonkhorstPack(npt=9), ab.Basis(np.eye(d)))):
        u1 = ab.solve(ab.IntegralProblem(f1, dom), alg, abstol=1e-8, backend=be).u
        u2 = ab.solve(ab.IntegralProblem(f2, dom), alg, abstol=1e-8, backend=be).u
        assert abs(u1 - u2) < 1e-12 * abs(u1) and abs(u1 - 4.2) < 1e-7
    with pytest.raises(TypeError):
        ab.FourierIntegrand(p, s, nest=object())


def test_mixed_parameters_paramzip_paramproduct():
    """test/brillouin.jl:46-61 (MixedParameters merge rules) and :98-111 (batchsolve over paramzip / paramproduct arrays
    equals the loops over solver(a, b=b)), with the affine Fourier integrand of test/fourier.jl:41 as f(x, a; b)"""
    args, kwargs = (1, 2), {"a": "a", "b": "b"}
    p, q = ab.MixedParameters(*args), ab.MixedParameters(**kwargs)
    for pq in (ab.merge(p, q), ab.merge(p, kwargs), ab.merge(q, args)):
        assert pq[0] == args[0] and pq[1] == args[1] and pq.a == "a" and pq.b == "b"
    assert ab.merge(p, 3)[2] == 3 and ab.merge(q, 3)[0] == 3            # generic values are appended
    assert ab.merge(p, {"a": "c"}).a == "c" and ab.merge(q, {"a": "c"}).a == "c"   # keywords overwritten
    z = ab.paramzip([1, 2, 3], b=[4, 5, 6])
    assert [(m[0], m.b) for m in z] == [(1, 4), (2, 5), (3, 6)]
    pp = ab.paramproduct([1, 2, 3], b=[4, 5])
    assert pp.shape == (3, 2) and pp[2, 1] == ab.MixedParameters(3, b=5)

    s = lattice_series(2)
    bz = ab.load_bz(ab.FBZ(), np.eye(2))
    f = ab.FourierIntegrand(lambda x, a, b=0.0: a * x.s + b, s)
    solver = ab.IntegralSolver(f, bz, ab.PTR(npt=6), backend=OracleBackend())
    rng = np.random.default_rng(3)
    as_, bs = rng.random(3), rng.random(2)
    want_zip = [solver(a, b=b) for a, b in zip(as_, bs)]
    assert np.allclose(ab.batchsolve(solver, ab.paramzip(as_, b=bs)), want_zip, rtol=1e-14)
    want_prod = np.array([[solver(a, b=b) for b in bs] for a in as_])
    got = ab.batchsolve(solver, ab.paramproduct(as_, b=bs))
    assert got.shape == (3, 2) and np.allclose(got, want_prod, rtol=1e-14)
    # parameters preloaded in the integrand merge with the solver's call (ParameterIntegrand, test/brillouin.jl:83-96)
    f2 = ab.FourierIntegrand(lambda x, a, b=0.0: a * x.s + b, s, b=bs[0])
    solver2 = ab.IntegralSolver(f2, bz, ab.PTR(npt=6), backend=OracleBackend())
    assert np.allclose(solver2(as_[0]), solver(as_[0], b=bs[0]), rtol=1e-14)
    assert np.allclose(ab.batchsolve(solver2, [ab.MixedParameters(a) for a in as_]), [solver(a, b=bs[0]) for a in as_], rtol=1e-14)


def test_absolute_estimate_ptr_iai():
    """AbsoluteEstimate / PTR_IAI / AutoPTR_IAI (src/algorithms.jl:614-653, src/brillouin.jl:466-490): the estimate from the first
    algorithm turns reltol into the abstol of the second; the result is the IAI solve with that abstol and reltol = 0, and
    EvalCounter counts the evaluations of both solves"""
    be = OracleBackend()
    H, lo = ab.synthetic.wannier_hamiltonian(2, 1, cubic=True)
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=2)
    ext = ab.synthetic.band_extent(H)
    bz = ab.load_bz(ab.CubicSymIBZ(), np.eye(3))
    f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=0.3 * ext)
    p = {"omega": 0.1 * ext}
    prob = ab.IntegralProblem(f, bz, p)
    rtol = 1e-3
    est = ab.solve(prob, ab.EvalCounter(ab.PTR(npt=6)), backend=be)
    want = ab.solve(prob, ab.EvalCounter(ab.IAI()), abstol=rtol * abs(est.u), reltol=0.0, backend=be)
    got = ab.solve(prob, ab.EvalCounter(ab.PTR_IAI(ptr=ab.PTR(npt=6))), reltol=rtol, backend=be)
    assert got.u == want.u and got.numevals == est.numevals + want.numevals
    # an explicit abstol larger than reltol * |estimate| wins (abstol = max(abstol, reltol * norm(I)))
    big = 10 * rtol * abs(est.u)
    got2 = ab.solve(prob, ab.PTR_IAI(ptr=ab.PTR(npt=6)), reltol=rtol, abstol=big, backend=be)
    want2 = ab.solve(prob, ab.IAI(), abstol=big, reltol=0.0, backend=be)
    assert got2.u == want2.u
    # AutoPTR_IAI: the estimate is an AutoPTR solve at its own (loose) reltol
    est3 = ab.solve(prob, ab.AutoPTR(nmin=6, a=0.3), reltol=1.0, backend=be)
    got3 = ab.solve(prob, ab.AutoPTR_IAI(ptr=ab.AutoPTR(nmin=6, a=0.3)), reltol=rtol, backend=be)
    want3 = ab.solve(prob, ab.IAI(), abstol=rtol * abs(est3.u), reltol=0.0, backend=be)
    assert got3.u == want3.u
    # through IntegralSolver / batchsolve as well
    solver = ab.IntegralSolver(f, bz, ab.PTR_IAI(ptr=ab.PTR(npt=6)), reltol=rtol, backend=be)
    assert solver(omega=0.1 * ext) == want.u
    with pytest.raises(ValueError):
        ab.AbsoluteEstimate(ab.PTR(), ab.IAI(), tolerance=1.0)
    with pytest.raises(TypeError):
        ab.solve(prob, ab.TAI(), backend=be)


def test_symrep_names():
    """src/brillouin.jl:44-114: TrivialRep multiplies by nsyms, UnknownRep returns the IBZ value, the full BZ is the identity;
    load_bz(IBZ()) needs an extension that is out of scope"""
    ibz, fbz = ab.load_bz(ab.CubicSymIBZ(), np.eye(3)), ab.load_bz(ab.FBZ(), np.eye(3))
    assert ab.symmetrize(None, ibz, 2.0) == 96.0 and ab.symmetrize(None, fbz, 2.0) == 2.0
    x = np.ones((2, 2))
    assert ab.symmetrize(object(), ibz, x) is x and isinstance(ab.SymRep(object()), ab.UnknownRep)
    assert np.all(ab.symmetrize(ab.TrivialRep(), ibz, x) == 48 * x)
    g = ab.FourierIntegrand(ab.GlocIntegrand(symmetrize=lambda bz, v: 2 * v), lattice_series(3), eta=0.1)
    assert isinstance(ab.SymRep(g), ab.FunctionRep) and np.all(ab.symmetrize(g, ibz, x) == 2 * x)
    with pytest.raises(NotImplementedError):
        ab.load_bz(ab.IBZ(), np.eye(3))


def test_ibz_with_supplied_polyhedron(orc):
    """src/brillouin.jl:205-247 + ext/SymmetryReduceBZExt.jl: load_bz(IBZ) = SymmetricBZ(A, B, polyhedral limits, point group).  The
    polyhedron and the group (what SymmetryReduceBZ computes) are supplied by the caller; for the cubic group they are the
    tetrahedron and the 48 automorphisms of load_bz(CubicSymIBZ), and the integral over the two must agree
    (test/test_ibz.jl:151-180 compares IBZ against FBZ integrals the same way)."""
    from autobz_b200.bz import cube_automorphisms
    verts = [(0, 0, 0), (0, 0, 0.5), (0, 0.5, 0.5), (0.5, 0.5, 0.5)]
    pbz = ab.load_bz(ab.IBZ(3, polyhedron=verts, syms=cube_automorphisms(3)), np.eye(3))
    cbz, fbz = ab.load_bz(ab.CubicSymIBZ(), np.eye(3)), ab.load_bz(ab.FBZ(), np.eye(3))
    assert pbz.nsyms == 48 and isinstance(pbz.lims, ab.PolyhedronLimits)
    H, lo = ab.synthetic.wannier_hamiltonian(2, 1, cubic=True)
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=2)
    be = OracleBackend()
    f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=4.0)
    a = ab.solve(ab.IntegralProblem(f, pbz, {"omega": 0.3}), ab.EvalCounter(ab.IAI()), abstol=0.05, backend=be)
    b = ab.solve(ab.IntegralProblem(f, cbz, {"omega": 0.3}), ab.EvalCounter(ab.IAI()), abstol=0.05, backend=be)
    c = ab.solve(ab.IntegralProblem(f, fbz, {"omega": 0.3}), ab.PTR(npt=32), backend=be)
    assert abs(a.u - c.u) < 0.05 and abs(b.u - c.u) < 0.05 and a.numevals > 0
    with pytest.raises(ValueError):
        ab.load_bz(ab.IBZ(polyhedron=[(0, 0, 0), (1, 0, 0)], syms=[np.eye(3)]), np.eye(3))


def test_kronrod_rules_match_quadpack_tables():
    """AuxQuadGKJL(order=n) (src/algorithms.jl:202-208): the rule generator against QUADPACK's published qk15 / qk21 / qk31 constants,
    polynomial exactness (degree 3n+1, or 3n+2 for odd n) and the embedded Gauss rule, n = 2 ... 20"""
    from autobz_b200 import iai
    x, w, gw = iai.kronrod(7)
    assert np.abs(x - iai.GK_X).max() < 5e-16 and np.abs(w - iai.GK_W).max() < 5e-16 and np.abs(gw - iai.GK_GW).max() < 5e-16
    x, w, gw = iai.kronrod(10)
    assert abs(x[0] + 0.995657163025808080735527280689003) < 5e-16 and abs(w[-1] - 0.149445554002916905664936468389821) < 5e-16
    x, w, gw = iai.kronrod(15)
    assert abs(x[0] + 0.998002298693397060285172840152271) < 5e-16 and abs(w[-1] - 0.101330007014791549017374792767493) < 5e-16
    for n in range(2, 21):
        r = iai.GKRule(n)
        assert r.K == 2 * n + 1 and r.off.size == r.K and np.all(r.w > 0)
        xs = np.concatenate([r.x, -r.x[:-1][::-1]])
        ws = np.concatenate([r.w, r.w[:-1][::-1]])
        deg = 3 * n + 1 + (n % 2)
        for k in range(0, deg + 1, 2):
            assert abs(np.sum(ws * xs ** k) - 2 / (k + 1)) < 2e-15
        xg, wg = np.polynomial.legendre.leggauss(n)
        assert np.abs(r.x[1::2] - xg[: n // 2 + (n % 2)][: len(r.x[1::2])]).max() < 5e-16
        # evalrule of a polynomial of degree 2n - 1: Gauss and Kronrod agree, E ~ 0, I exact
        f = lambda t: t ** (2 * n - 1) + 3 * t ** 2
        I, E = r.combine(0.5, 2.0, f(r.nodes(0.5, 2.0)))
        exact = (2.0 ** (2 * n) - 0.5 ** (2 * n)) / (2 * n) + 2.0 ** 3 - 0.5 ** 3
        assert abs(I - exact) < 1e-13 * abs(exact) and E < 1e-12 * abs(exact)
    assert iai.GKRule(7) is iai.GKRule(7)
    with pytest.raises(ValueError):
        ab.AuxQuadGKJL(order=1)


def _nested_quadgk_sequential(point_value, lims, orders, atol, rtol, maxevals):
    """The reference's recursion written down directly (QuadGK do_quadgk + adapt per level, abstol / len for the inner levels,
    src/fourier.jl:474-481): an independent control flow against which the level-synchronous engine is checked."""
    from autobz_b200 import iai
    count = [0]

    def quadgk(f, segs, rule, atol_):
        heap = []
        for a, b in zip(segs[:-1], segs[1:]):
            I, E = rule.combine(a, b, np.array([f(x) for x in rule.nodes(a, b)]))
            heap.append((float(E), a, b, I[()]))
        I, E = heap[0][3], heap[0][0]
        for s in heap[1:]:
            I, E = I + s[3], E + s[0]
        ne = rule.K * len(heap)
        if ne >= maxevals or E <= atol_ or E <= rtol * abs(I):
            return I
        for i in range(len(heap) // 2, 0, -1):
            iai._percolate_down(heap, i, heap[i - 1], len(heap))
        while E > atol_ and E > rtol * abs(I) and ne < maxevals:
            s = iai.heappop(heap)
            mid = (s[1] + s[2]) / 2
            new = []
            for a, b in ((s[1], mid), (mid, s[2])):
                Ii, Ei = rule.combine(a, b, np.array([f(x) for x in rule.nodes(a, b)]))
                new.append((float(Ei), a, b, Ii[()]))
            I = (I - s[3]) + new[0][3] + new[1][3]
            E = (E - s[0]) + new[0][0] + new[1][0]
            ne += 2 * rule.K
            iai.heappush(heap, new[0])
            iai.heappush(heap, new[1])
        I = heap[0][3]
        for s in heap[1:]:
            I = I + s[3]
        return I

    def level(l, lims_, outer, atol_):
        rule = iai.GKRule(orders[l])
        segs = tuple(lims_.segments())
        if l == 0:
            def f(x):
                count[0] += 1
                return point_value((float(x),) + outer)
        else:
            def f(x):
                cl = lims_.fix(float(x))
                cs = tuple(cl.segments())
                return level(l - 1, cl, (float(x),) + outer, atol_ / (cs[-1] - cs[0]))
        return quadgk(f, segs, rule, atol_)

    return level(len(orders) - 1, lims, (), atol), count[0]


@pytest.mark.parametrize("ndim,orders", [(1, (10,)), (2, (4, 9)), (3, (5, 7, 3)), (3, (7, 7, 7))])
def test_per_level_gauss_kronrod_orders(orc, ndim, orders):
    """IAI(algs...) with a different AuxQuadGKJL order per variable (src/brillouin.jl:368-377, src/algorithms.jl:462-463: algs[dim]
    belongs to variable dim, the innermost is algs[1]): same evaluation count and value as the sequential recursion"""
    n = 2
    H, lo = ab.synthetic.wannier_hamiltonian(n, 1, cubic=True)
    H = np.asfortranarray(H[(slice(None),) * (2 + ndim) + (1,) * (3 - ndim)])
    fs = ab.FourierSeries(H, period=1.0, lo=lo[:ndim], norb=n)
    S = orc.Series(np.asfortranarray(H.reshape(H.shape + (1,) * (3 - ndim))), tuple(lo[:ndim]) + (0,) * (3 - ndim))
    z = complex(0.3, 3.0)
    bz = ab.load_bz(ab.InversionSymIBZ(ndim), np.eye(ndim))
    mult = abs(np.linalg.det(bz.B)) * bz.nsyms
    atol = 2e-3

    def point_value(k):
        Hk = orc.eval_points(S, np.array([tuple(k) + (0.0,) * (3 - ndim)]))[:, :, 0]
        return np.trace(np.linalg.inv(z * np.eye(n) - Hk))

    Iref, nref = _nested_quadgk_sequential(point_value, bz.lims, orders, atol, 0.0, 10 ** 7)
    alg = ab.IAI(*[ab.AuxQuadGKJL(order=o) for o in orders])
    sol = ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=z.imag), bz, {"omega": z.real}), ab.EvalCounter(alg),
                   abstol=atol * mult, backend=OracleBackend())
    assert sol.numevals == nref
    assert abs(sol.u - mult * Iref) <= 1e-12 * abs(sol.u)
    with pytest.raises(ValueError):
        ab.solve(ab.IntegralProblem(ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=1.0), bz, {"omega": 0.0}),
                 ab.IAI(ab.AuxQuadGKJL(), ab.AuxQuadGKJL(), ab.AuxQuadGKJL(), ab.AuxQuadGKJL()), abstol=1.0, backend=OracleBackend())


def test_serpentine_dealing_balances_the_irreducible_wedge():
    """`k3_stride = -nranks` (backend.share_planes mirrors csrc/abz_common.cuh:share_plane): every plane goes to exactly one rank, and
    on the cubic IBZ the largest share is within 5 % of the average where the reference's round-robin dealing (src/fourier.jl:246-255)
    is 22 % above it (npt = 96, 8 ranks: 2 709 against 3 171 of 20 825 nodes)."""
    import orc
    from autobz_b200.backend import share_planes
    orc.build()
    N, W = 96, 8
    w, nirr = orc.symptr_rule(N, ab.cube_automorphisms(3))
    per_plane = (np.asarray(w) != 0).sum(axis=(0, 1))
    for stride in (W, -W):
        planes = [share_planes(N, r, stride) for r in range(W)]
        assert sorted(p for pl in planes for p in pl) == list(range(N))
    rr = max(int(per_plane[share_planes(N, r, W)].sum()) for r in range(W))
    sp = max(int(per_plane[share_planes(N, r, -W)].sum()) for r in range(W))
    assert (rr, sp, nirr) == (3171, 2709, 20825)
    assert share_planes(10, 1, -2) == [1, 2, 5, 6, 9] and share_planes(10, 0, -2) == [0, 3, 4, 7, 8]
