"""Round-2 GPU tests: BASELINE config 3 at its own eta against committed oracle goldens, AutoPTR against an independent
restatement of its refinement loop, parameter sweeps over one diagonalisation, plane-wise (sharded) construction of
symmetry-reduced rules, object lifetimes across the C ABI, several contexts in one process."""
import json
import os

import numpy as np
import pytest

import autobz_b200 as ab
from autobz_b200 import _lib as L

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def test_c3_at_baseline_eta_matches_oracle_goldens(svo):
    """BASELINE config 3 as written: SrVO3 DOS via IAI on the cubic IBZ at eta = 1e-4, abstol = 1e-3.  The oracle's sequential
    recursion needs ~70 s of one core per frequency, so its `numevals` and integrals are committed goldens
    (tests/golden/c3_eta1e-4.json, made by tests/golden/make_golden_c3.py): identical adaptive evaluation counts
    (121 241 805 and 156 928 605) and integrals within 1e-10 relative, for the device-leaf engine and the host-driven panels."""
    G = json.load(open(os.path.join(HERE, "golden", "c3_eta1e-4.json")))
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    ibz = ab.load_bz(ab.CubicSymIBZ(), A)
    mult = abs(np.linalg.det(ibz.B)) * 48
    assert abs(abs(np.linalg.det(ibz.B)) - G["j"]) < 1e-12 * G["j"]
    f = ab.FourierIntegrand(ab.dos_integrand, fs, G["eta"])
    for case in G["cases"]:
        sol = ab.solve(ab.IntegralProblem(f, ibz, case["omega"]), ab.EvalCounter(ab.IAI()), abstol=G["abstol_physical"])
        assert sol.numevals == case["numevals"], (case["omega"], sol.numevals, case["numevals"])
        assert abs(sol.u / mult - case["I"]) <= 1e-10 * abs(case["I"])
        assert abs(sol.resid / mult - case["E"]) <= 1e-6 * abs(case["E"])
    # the engine with host-driven innermost panels takes the same decisions (one frequency: ~1 s)
    be = ab.DeviceBackend(iai_engine="native", iai_device_leaves=False)
    case = G["cases"][0]
    sol = ab.solve(ab.IntegralProblem(f, ibz, case["omega"]), ab.EvalCounter(ab.IAI()), abstol=G["abstol_physical"], backend=be)
    assert sol.numevals == case["numevals"] and abs(sol.u / mult - case["I"]) <= 1e-10 * abs(case["I"])


def _independent_autoptr(orc, So, syms, n0, dn, z, atol, rtol):
    """AutoSymPTR.autosymptr restated HERE, independently of autobz_b200.interfaces._autosymptr, on the oracle's own symmetric
    rules: I_i on grids n0, n0 + dn, ...; stop when |I_i - I_{i-1}| <= max(rtol |I_i|, atol); numevals = sum of rule lengths.
    (The dependency itself is not in the reference tree - SURVEY.md App. A.2 - so this pins the count parity against a second
    restatement, not against AutoSymPTR.jl: "parity unpinned" for the exact rounding of n0/a, dn/a remains.)"""
    def rule(npt):
        w, _ = orc.symptr_rule(npt, syms)
        v, cnt = orc.symptr_sum(So, npt, w, [z], scale=1.0 / npt ** 3)
        return v[0], cnt
    npt = n0
    I1, ne = rule(npt)
    npt += dn
    I2, c = rule(npt)
    ne += c
    err = abs(I1 - I2)
    while not err <= max(rtol * abs(I2), atol):
        npt += dn
        I1 = I2
        I2, c = rule(npt)
        ne += c
        err = abs(I1 - I2)
    return I2, err, ne, npt


def test_autoptr_counts_vs_independent_restatement(orc, svo):
    H, lo, A = svo
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=3)
    So = orc.Series(H, lo)
    ibz = ab.load_bz(ab.CubicSymIBZ(), A)
    j = abs(np.linalg.det(ibz.B))
    syms = ab.cube_automorphisms(3)
    from autobz_b200.algorithms import monkhorst_pack_schedule
    assert monkhorst_pack_schedule(1.0, 50, 1000, 6.0, np.log(10)) == (50, 3)          # the reference's defaults: 50, 53, 56, ...
    f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=0.1)
    for a, nmin, omega, abstol in ((0.2, 20, 11.0, 1e-2), (0.2, 20, 12.5, 1e-3), (0.5, 12, 12.9, 1e-2)):
        n0, dn = monkhorst_pack_schedule(a, nmin, 400, 6.0, np.log(10))
        Io, Eo, neo, last = _independent_autoptr(orc, So, syms, n0, dn, complex(omega, 0.1), abstol / j, 0.0)
        cache = ab.init(ab.IntegralProblem(f, ibz, {"omega": omega}), ab.EvalCounter(ab.AutoPTR(a=a, nmin=nmin, nmax=400)), abstol=abstol)
        sol = ab.solve_(cache)
        assert sol.numevals == neo and cache.cacheval["last_npt"] == last
        assert abs(sol.u - j * Io) <= 1e-10 * abs(sol.u)
        assert abs(sol.resid - j * Eo) <= 1e-6 * abs(sol.resid) + 1e-14


@pytest.mark.parametrize("n", [5, 32, 64])
def test_eig_parameter_sweep_over_one_diagonalisation(ctx, orc, n):
    """abz_rule_eig_sum_batch: all (mu, T) of a sweep from ONE diagonalisation per node equal the per-parameter sums and the oracle
    (src/interfaces.jl:199-243: the reference's sweep shares the cached grid); a materialised rule keeps the eigenvalues cached."""
    N = 6 if n < 64 else 4
    H, lo = ab.synthetic.wannier_hamiltonian(n, 1, cubic=True)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    So = orc.Series(H, lo)
    syms = np.array(ab.cube_automorphisms(3), dtype=np.int32)
    prm = np.array([(mu, T) for mu in (-0.3, 0.0, 0.2) for T in (0.05, 0.4)])
    for R, w in ((L.DeviceRule(ctx, S, N), None), (L.DeviceRule(ctx, S, N, syms=syms), orc.symptr_rule(N, syms)[0])):
        for kind in (L.EIG_FERMI_ENERGY, L.EIG_FERMI_COUNT, L.EIG_GAUSS_DOS):
            got = R.eig_sum_batch(kind, prm, scale=1.0 / N ** 3)
            one = np.array([R.eig_sum(kind, p, 1.0 / N ** 3) for p in prm])
            ref = np.array([orc.ptr_eig_sum(So, N, kind, p, wsym=w, scale=1.0 / N ** 3)[0] for p in prm])
            assert np.max(np.abs(got - one)) <= 1e-13 * n and np.max(np.abs(got - ref)) <= 1e-11 * n
        R.materialize()
        l0 = ctx.launch_count
        first = R.eig_sum_batch(L.EIG_FERMI_ENERGY, prm, scale=1.0 / N ** 3)
        l1 = ctx.launch_count
        again = R.eig_sum_batch(L.EIG_FERMI_ENERGY, prm[::-1], scale=1.0 / N ** 3)
        l2 = ctx.launch_count
        assert np.array_equal(first, again[::-1]) and (l2 - l1) < (l1 - l0)       # second sweep: sums over cached eigenvalues only
        R.close()
    # through the public API: batchsolve over (mu, T) = one device pass
    fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=n)
    ibz = ab.load_bz(ab.CubicSymIBZ(), 2 * np.pi * np.eye(3))
    solver = ab.IntegralSolver(ab.FourierIntegrand(ab.EigenIntegrand("fermi_energy"), fs), ibz, ab.PTR(npt=N))
    l0 = ctx.launch_count
    many = ab.batchsolve(solver, [tuple(p) for p in prm])
    dl_many = ctx.launch_count - l0
    l0 = ctx.launch_count
    one = solver(*prm[0])
    dl_one = ctx.launch_count - l0
    assert abs(many[0] - one) <= 1e-12 * abs(one) and dl_many <= dl_one + 2
    S.close()


@pytest.mark.parametrize("npt", [12, 45, 200])
def test_symmetric_rule_built_plane_wise_equals_whole(ctx, orc, npt):
    """abz_rule_create_symptr with nirr_total == NULL computes the orbit weights of the selected k3 planes only (one rank's share
    of the construction): same nodes, order and weights as the slice of the whole-grid construction, for every rank of 3
    (npt = 200: the two-phase compaction path on 2.7 M points per rank)."""
    H, lo = ab.synthetic.wannier_hamiltonian(2, 1, cubic=True)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    syms = np.array(ab.cube_automorphisms(3), dtype=np.int32)
    w, nirr = orc.symptr_rule(npt, syms) if npt <= 64 else ctx.symptr_rule(npt, syms)
    tot = 0
    z = np.array([0.3 + 0.1j])
    whole = L.DeviceRule(ctx, S, npt, syms=syms)
    assert whole.nirr_total == nirr == len(whole)
    acc = 0
    for rank in range(3):
        Ra = L.DeviceRule(ctx, S, npt, syms=syms, k3_lo=rank, k3_stride=3, count_all=False)
        Rb = L.DeviceRule(ctx, S, npt, syms=syms, k3_lo=rank, k3_stride=3, count_all=True)
        assert Ra.nirr_total is None and Rb.nirr_total == nirr and len(Ra) == len(Rb)
        _, ka, wa = Ra.copy_out(want_h=False)
        _, kb, wb = Rb.copy_out(want_h=False)
        assert np.array_equal(ka, kb) and np.array_equal(wa, wb)
        tot += len(Ra)
        acc = acc + Ra.resolvent_sum(z)
        Ra.close(); Rb.close()
    assert tot == nirr
    assert rel(acc, whole.resolvent_sum(z)) < 1e-13
    whole.close(); S.close()


def test_series_handle_can_be_retired_before_its_rules(ctx, orc):
    """abz_series_destroy while rules / nests still use the series (a Python GC order, FourierSeries.drop_device()): the
    dependants keep the coefficients alive, the handle itself is gone."""
    H, lo = ab.synthetic.wannier_hamiltonian(4, 1)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    R = L.DeviceRule(ctx, S, 6)
    nest = L.DeviceNest(ctx, S, 3, 4, 16)
    z = np.array([0.2 + 0.1j])
    want = R.resolvent_sum(z)
    S.close()
    # churn the pool so that a recycled coefficient block would show
    junk = [L.DeviceSeries(ctx, H * (i + 2), lo, (1.0,) * 3) for i in range(3)]
    assert np.array_equal(R.resolvent_sum(z), want)
    nest.contract3([0.1], [0]); nest.contract2([0.2], [0], [0])
    y = nest.eval([0.3], [0], z)
    So = orc.Series(H, lo)
    assert rel(y, orc.resolvent_trace_batch(orc.eval_points(So, [[0.3, 0.2, 0.1]]), z)[0]) < 1e-12
    with pytest.raises(ValueError):
        ctx.check(ctx.lib.abz_series_destroy(ctx.h, 10 ** 9))
    for j in junk:
        j.close()
    R.close(); nest.close()


def test_several_contexts_in_one_process(orc):
    """One context per device (and several per device) in ONE process: the > 48 KB dynamic shared-memory opt-in is made per
    device at context creation, so kernels that need it run in every context (n = 32: DMMA resolvent with 128 frequencies,
    contraction stages; n = 64: eigenvalues)."""
    import torch
    ndev = torch.cuda.device_count()
    H, lo = ab.synthetic.wannier_hamiltonian(32, 1)
    So = orc.Series(H, lo)
    z = np.linspace(-1, 1, 128) + 0.05j
    ref = orc.ptr_sum(So, 4, z)
    ctxs = [L.Context(d) for d in ([0, 0] + ([1] if ndev > 1 else []))]
    for c in ctxs:
        S = L.DeviceSeries(c, H, lo, (1.0,) * 3)
        R = L.DeviceRule(c, S, 4)
        assert rel(R.resolvent_sum(z, scale=1 / 64), ref) < 1e-11
        assert np.isfinite(R.eig_sum(L.EIG_SUM, (0.0, 1.0)))
        R.close(); S.close()
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("n", [33, 40, 47, 48, 56, 63, 64])
def test_dmma_team_resolvent_vs_oracle(ctx, orc, n):
    """32 < norb <= 64 through the default algorithm = the DMMA block LU shared by a team of 4 warps
    (csrc/abz_resolvent_mma_team.cuh): rule sums (with and without a matrix self-energy, 130 frequencies so that a team walks
    several matrices), per-point values, against the oracle's pivoted LU and the pivoted Gauss-Jordan teams (<= 1e-10)."""
    N = 4
    H, lo = ab.synthetic.wannier_hamiltonian(n, 1)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    So = orc.Series(H, lo)
    ext = ab.synthetic.band_extent(H)
    rng = np.random.default_rng(n)
    z = np.concatenate([rng.uniform(-ext, ext, 127) + 0.02j * ext, rng.uniform(-0.5 * ext, 0.5 * ext, 3) + 2e-3j * ext])
    sig = 0.05 * ext * (rng.standard_normal((n, n, z.size)) + 1j * rng.standard_normal((n, n, z.size))) - 0.1j * ext * np.eye(n)[:, :, None]
    R = L.DeviceRule(ctx, S, N)
    l0 = ctx.launch_count
    got = R.resolvent_sum(z, scale=1 / N ** 3)
    assert ctx.launch_count - l0 <= 6           # stages 3, 2, 1 + ONE resolvent launch + reduction: no pivoted rerun
    assert rel(got, orc.ptr_sum(So, N, z)) < 1e-10
    assert rel(R.resolvent_sum(z[:9], sigma=sig[:, :, :9], scale=1 / N ** 3), orc.ptr_sum(So, N, z[:9], sigma=sig[:, :, :9])) < 1e-10
    ctx.set_option(L.OPT_RESOLVENT_ALGO, 1)
    try:
        assert rel(got, R.resolvent_sum(z, scale=1 / N ** 3)) < 1e-10
    finally:
        ctx.set_option(L.OPT_RESOLVENT_ALGO, 0)
    kp = rng.random((11, 3))
    assert rel(S.points_resolvent(kp, z[:5]), orc.resolvent_trace_batch(orc.eval_points(So, kp), z[:5])) < 1e-10
    R.close(); S.close()


@pytest.mark.parametrize("n", [33, 64])
def test_dmma_team_resolvent_falls_back_to_pivoting(ctx, orc, n):
    """matrices that need row exchanges (zero diagonal, z ~ 0) through the DEFAULT algorithm: the team kernel's pivot monitor
    raises its flag and the call is rerun with the pivoted teams - the answer is the oracle's"""
    rng = np.random.default_rng(200 + n)
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
    R = L.DeviceRule(ctx, S, 4)
    assert rel(R.resolvent_sum(z, scale=1 / 64), orc.ptr_sum(So, 4, z)) < 1e-10
    R.close(); S.close()


@pytest.mark.parametrize("n", [1, 3, 16, 17, 31, 32, 33, 40, 63, 64])
def test_eigenvalues_stage_b_ql_and_bisection_vs_lapack(ctx, n):
    """Stage B of the eigenvalue path (eigen(Hermitian(H(k))), src/dos_ggr.jl:19,34): the thread-per-matrix QL kernel (ABZ_OPT_EIG_ALGO 3)
    and the warp-per-matrix Sturm bisection kernel (4) against LAPACK, sorted output, and the band sums of both against numpy;
    includes sizes that leave lanes without an eigenvalue and degenerate spectra (cubic symmetry at the zone centre)"""
    from autobz_b200 import _lib as L
    H, lo = ab.synthetic.wannier_hamiltonian(n, 1, cubic=True)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    R = L.DeviceRule(ctx, S, 5)
    Hk, _, _ = R.copy_out()
    ref = np.linalg.eigvalsh(np.moveaxis(Hk, 2, 0))
    rad = np.max(np.abs(ref))
    out = {}
    try:
        for algo in (3, 4):
            ctx.set_option(L.OPT_EIG_ALGO, algo)
            ev = R.eigvals()
            assert np.max(np.abs(ev - ref)) <= 1e-13 * rad, (algo, np.max(np.abs(ev - ref)) / rad)
            assert np.all(np.diff(ev, axis=1) >= 0)
            out[algo] = [R.eig_sum(0, (0.0, 1.0), scale=1.0), R.eig_sum(1, (0.1 * rad, 0.05 * rad), scale=1.0),
                         R.eig_sum(3, (0.0, 0.1 * rad), scale=1.0)]
    finally:
        ctx.set_option(L.OPT_EIG_ALGO, 0)
    f = lambda x: 1.0 / (1.0 + np.exp(x))
    exact = [ref.sum(), (ref * f((ref - 0.1 * rad) / (0.05 * rad))).sum(), (np.exp(-(ref / (0.1 * rad)) ** 2) / (0.1 * rad * np.sqrt(np.pi))).sum()]
    for algo in (3, 4):
        for got, want in zip(out[algo], exact):
            assert abs(got - want) <= 1e-11 * max(abs(want), rad), (algo, got, want)


def test_bisection_handles_decoupled_and_scaled_matrices(ctx):
    """diagonal H (every off-diagonal exactly zero), a zero matrix, and spectra scaled by 1e-150 / 1e+150: the power-of-two scaling
    and the periodic rescaling of the Sturm recurrence keep the counts exact"""
    from autobz_b200 import _lib as L
    n = 48
    rng = np.random.default_rng(5)
    try:
        ctx.set_option(L.OPT_EIG_ALGO, 4)
        for scale in (1.0, 1e-150, 1e150, 0.0):
            H = np.zeros((n, n, 3, 3, 3), dtype=np.complex128, order="F")
            dvals = rng.normal(size=n)
            H[:, :, 1, 1, 1] = np.diag(dvals) * scale
            if scale == 1.0:                     # a second case with weak coupling between nearly equal diagonal entries
                A = 1e-9 * (rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n)))
                H[:, :, 1, 1, 1] += A + A.conj().T
            S = L.DeviceSeries(ctx, H, (-1, -1, -1), (1.0,) * 3)
            R = L.DeviceRule(ctx, S, 2)
            ev = R.eigvals()
            ref = np.linalg.eigvalsh(H[:, :, 1, 1, 1])
            assert np.max(np.abs(ev - ref[None])) <= 1e-13 * max(np.max(np.abs(ref)), 1e-300)
            R.close(); S.close()
    finally:
        ctx.set_option(L.OPT_EIG_ALGO, 0)


@pytest.mark.parametrize("W", [2, 3, 8])
def test_serpentine_plane_shares_partition_the_rule(ctx, orc, W):
    """k3_stride = -W deals the k3 planes of a symmetry-reduced rule in serpentine order (rank r: r, 2W-1-r, 2W+r, 4W-1-r, ...):
    the shares are disjoint, cover every plane, follow that sequence, give the same rules through both construction paths
    (device symptr / host wsym array), add up to the whole rule's sums - and balance the node counts of the cubic wedge better than
    the round-robin dealing (src/fourier.jl:246-255 deals its thread chunks round-robin; the node set and the sums are the same)."""
    n, N = 3, 21
    H, lo = ab.synthetic.wannier_hamiltonian(n, 2, cubic=True)
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    syms = np.array(ab.cube_automorphisms(3), dtype=np.int32)
    w_host, nirr = ctx.symptr_rule(N, syms)
    whole = L.DeviceRule(ctx, S, N, syms=syms)
    z = np.array([0.3 + 0.05j, -0.7 + 0.2j])
    ref = orc.symptr_sum(orc.Series(H, lo), N, w_host, z, scale=1.0)[0]
    shares = [L.DeviceRule(ctx, S, N, syms=syms, k3_lo=r, k3_stride=-W, count_all=False) for r in range(W)]
    assert sum(len(s) for s in shares) == nirr == len(whole)
    seen = set()
    for r, sh in enumerate(shares):
        Hk, k, w = sh.copy_out()
        planes = sorted(set(np.rint(k[:, 2] * N).astype(int).tolist()))
        expect = [p for p in range(N) if (p % (2 * W)) in (r, 2 * W - 1 - r)]
        assert set(planes) <= set(expect) and not (seen & set(planes))
        seen |= set(planes)
        via_host = L.DeviceRule(ctx, S, N, wsym=w_host, k3_lo=r, k3_stride=-W)
        Hh, kh, wh = via_host.copy_out()
        assert np.array_equal(k, kh) and np.array_equal(w, wh) and np.array_equal(Hk, Hh)
    tot = sum(s.resolvent_sum(z) for s in shares)
    assert np.max(np.abs(tot - ref) / np.abs(ref)) < 1e-11
    rr = [len(L.DeviceRule(ctx, S, N, syms=syms, k3_lo=r, k3_stride=W, count_all=False)) for r in range(W)]
    assert max(len(s) for s in shares) <= max(rr)
    with pytest.raises(Exception):
        L.DeviceRule(ctx, S, N, syms=syms, k3_lo=W, k3_stride=-W)
