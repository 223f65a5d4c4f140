"""The library's C++ host-side IAI engine (csrc/abz_iai_engine.hpp, behind abz_iai_solve) driven on CPU by an
oracle-backed test double (tests/native/iai_engine_cpu.cpp): it must reproduce the oracle's sequential recursion
(orc_iai) decision for decision - identical numevals, integrals equal to rounding - both with host-driven innermost
panels and with whole innermost integrals handed out as leaf tasks (what the device does with one warp per task)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "native"))

c_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def eng(orc):
    import build
    lib = C.CDLL(build.build())
    return lib


XFN = C.CFUNCTYPE(C.c_int, c_dp, C.c_long, C.c_void_p)


def _solve(lib, S, ndim, lkind, la, lb, fkind, vkind, z, lin, atol, rtol, leaf, maxevals=2 ** 62, cap2=64, cap1=2048, rank=0, nranks=1,
           xfn=None):
    i3 = lambda v: (C.c_int * 3)(*[int(x) for x in v])
    d3 = lambda v: (C.c_double * 3)(*[float(x) for x in v])
    la_ = np.ascontiguousarray(la, dtype=np.float64)
    lb_ = None if lb is None else np.ascontiguousarray(lb, dtype=np.float64)
    zz = np.array([z.real, z.imag])
    ln = None if lin is None else np.ascontiguousarray(lin, dtype=np.float64)
    out = np.zeros(3)
    st = (C.c_long * 6)()
    dp = lambda a: None if a is None else a.ctypes.data_as(c_dp)
    rc = lib.iai_cpu_solve(dp(S.c), C.c_int(S.n), C.c_int(ndim), i3(S.M), i3(S.lo), d3(S.period), C.c_int(lkind), dp(la_), dp(lb_),
                           C.c_int(fkind), C.c_int(vkind), dp(zz), None, dp(ln), C.c_double(atol), C.c_double(rtol),
                           C.c_long(maxevals), C.c_int(int(leaf)), C.c_long(cap2), C.c_long(cap1), C.c_int(rank), C.c_int(nranks),
                           xfn if xfn is not None else XFN(0), dp(out), st)
    _solve.last_stats = [int(x) for x in st]
    return rc, complex(out[0], out[1]), out[2], int(st[0]), int(st[1])


@pytest.mark.parametrize("leaf", [False, True, 2])
@pytest.mark.parametrize("lims", ["cubic", "tetra"])
def test_engine_matches_recursion_svo(orc, eng, svo, lims, leaf):
    H, lo, A = svo
    S = orc.Series(H, lo)
    z = complex(12.5, 0.05)
    if lims == "cubic":
        args, atol = (0, [0.0] * 3, [1.0] * 3), 3e-2
    else:
        args, atol = (1, [0.5] * 3, None), 1e-3
    Io, Eo, neo = orc.iai(S, 3, args[0], args[1], args[2], vkind=1, z=z, atol=atol)
    rc, I, E, ne, rounds = _solve(eng, S, 3, args[0], args[1], args[2], 0, 1, z, None, atol, 0.0, leaf)
    assert rc == 0
    assert ne == neo
    assert abs(I - Io) <= 1e-13 * abs(Io) and abs(E - Eo) <= 1e-9 * Eo
    if leaf:   # leaf tasks collapse the innermost adaptive loops into one device round each
        _, _, _, _, rounds_host = _solve(eng, S, 3, args[0], args[1], args[2], 0, 1, z, None, atol, 0.0, False)
        assert rounds < rounds_host


@pytest.mark.parametrize("leaf", [False, True, 2])
def test_engine_complex_values_and_rtol(orc, eng, svo, leaf):
    """complex-valued integrand (tr G), relative tolerance only, maxevals cut-off"""
    H, lo, A = svo
    S = orc.Series(H, lo)
    z = complex(12.0, 0.1)
    Io, Eo, neo = orc.iai(S, 3, 1, [0.5] * 3, None, vkind=0, z=z, atol=0.0, rtol=1e-3)
    rc, I, E, ne, _ = _solve(eng, S, 3, 1, [0.5] * 3, None, 0, 0, z, None, 0.0, 1e-3, leaf)
    assert rc == 0 and ne == neo and abs(I - Io) <= 1e-13 * abs(Io)
    Io, Eo, neo = orc.iai(S, 3, 1, [0.5] * 3, None, vkind=0, z=z, atol=1e-9, maxevals=45)
    rc, I, E, ne, _ = _solve(eng, S, 3, 1, [0.5] * 3, None, 0, 0, z, None, 1e-9, 0.0, leaf, maxevals=45)
    assert rc == 0 and ne == neo and abs(I - Io) <= 1e-13 * abs(Io)


@pytest.mark.parametrize("ndim", [1, 2])
def test_engine_low_dimensions_and_affine(orc, eng, ndim):
    """docs/src/examples.md:44-60,79-106 shapes (1-D, 2-D cos bands) and the affine integrand of test/fourier.jl:41"""
    shape = (1, 1) + (3,) * ndim + (1,) * (3 - ndim)
    c = np.zeros(shape, dtype=complex)
    for d in range(ndim):
        for sgn in (0, 2):
            idx = [0, 0] + [1] * ndim + [0] * (3 - ndim)
            idx[2 + d] = sgn
            c[tuple(idx)] = 0.5
    S = orc.Series(c, (-1,) * ndim + (0,) * (3 - ndim))
    z = complex(0.0, 0.1)
    Io, Eo, neo = orc.iai(S, ndim, 0, [0.0] * ndim, [1.0] * ndim, vkind=0, z=z, atol=1e-3)
    for leaf in (False, True):
        rc, I, E, ne, _ = _solve(eng, S, ndim, 0, [0.0] * ndim, [1.0] * ndim, 0, 0, z, None, 1e-3, 0.0, leaf)
        assert rc == 0 and ne == neo and abs(I - Io) <= 1e-13 * abs(Io)
    Io, Eo, neo = orc.iai(S, ndim, 0, [0.0] * ndim, [1.0] * ndim, vkind=2, lin=(1.3, 1.0), atol=1e-8)
    rc, I, E, ne, _ = _solve(eng, S, ndim, 0, [0.0] * ndim, [1.0] * ndim, 1, 2, 0j, [1.3, 0.0, 1.0, 0.0], 1e-8, 0.0, False)
    assert rc == 0 and ne == neo and abs(I - Io) <= 1e-13 * abs(Io) and abs(I - 1.0) < 1e-9


def test_engine_arena_exhaustion_is_an_error(orc, eng, svo):
    H, lo, A = svo
    S = orc.Series(H, lo)
    rc, *_ = _solve(eng, S, 3, 1, [0.5] * 3, None, 0, 1, complex(12.5, 0.05), None, 1e-3, 0.0, False, cap2=8, cap1=2048)
    assert rc == -2


_GLOO_IAI = r'''
import ctypes as C, os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle")); sys.path.insert(0, os.path.join({root!r}, "tests"))
sys.path.insert(0, os.path.join({root!r}, "tests", "native"))
import numpy as np, torch, torch.distributed as dist
import orc, build
from test_iai_engine_native import _solve, XFN
nranks = {nranks}
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=nranks)
rank = dist.get_rank()
lib = C.CDLL(build.build())
calls = [0]
def allreduce(buf, n, user):
    a = np.ctypeslib.as_array(buf, shape=(n,))
    t = torch.from_numpy(a.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    a[:] = t.numpy()
    calls[0] += 1
    return 0
xfn = XFN(allreduce)
d = np.load(os.path.join({root!r}, "tests", "golden", "svo_hr.npz"))
S = orc.Series(np.asfortranarray(d["H_R"]), tuple(int(x) for x in d["lo"]))
ok = True
for leaf in (False, True, 2, 6):      # 6: middle integrals as tasks + look-ahead on the outermost integral
    for (lk, la, lb, atol) in ((1, [0.5] * 3, None, 1e-3), (0, [0.0] * 3, [1.0] * 3, 3e-2)):
        z = complex(12.5, 0.05)
        if leaf == 6:
            if lk == 0: continue
            z, atol = complex(12.5, 0.005), 2e-5          # enough refinements of the outermost integral for look-ahead to happen
        rc1, I1, E1, ne1, r1 = _solve(lib, S, 3, lk, la, lb, 0, 1, z, None, atol, 0.0, leaf if leaf != 6 else 2, cap2=160)
        rc2, I2, E2, ne2, r2 = _solve(lib, S, 3, lk, la, lb, 0, 1, z, None, atol, 0.0, leaf, cap2=160, rank=rank, nranks=nranks, xfn=xfn)
        if leaf == 6 and not (_solve.last_stats[5] > 0 and r2 < r1): rc2 = -99      # the look-ahead case must actually look ahead
        good = rc1 == 0 and rc2 == 0 and I1 == I2 and E1 == E2 and ne1 == ne2     # bit-identical, evaluations of all ranks
        ok = ok and good
        print("RANK", rank, "leaf", leaf, "lims", lk, "OK" if good else "FAIL", I1, I2, ne1, ne2, "rounds", r1, r2, flush=True)
# a pole on one rank's share surfaces as an error on every rank (no hang)
c = np.zeros((1, 1, 3, 3, 1), dtype=complex); c[0, 0, 1, 1, 0] = 1.0
S1 = orc.Series(c, (-1, -1, 0))
rc, *_ = _solve(lib, S1, 2, 0, [0.0] * 2, [1.0] * 2, 0, 0, complex(1.0, 0.0), None, 1e-3, 0.0, False, rank=rank, nranks=nranks, xfn=xfn)
ok = ok and rc == -4
print("RANK", rank, "error rc", rc, "allreduce calls", calls[0], flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


@pytest.mark.parametrize("nranks,lanes", [(2, 1), (3, 1), (2, 3)])
def test_engine_sharded_over_ranks_gloo(tmp_path, orc, eng, nranks, lanes):
    """world_size 2 and 3 over gloo: the outermost panels' nodes are dealt round-robin to the ranks, one small allreduce per
    outer refinement step; integral, error estimate and total evaluation count are bit-identical to the single-rank solve,
    and an integrand error on one rank stops every rank.  (2, 3): the same with three rounds in flight per rank."""
    import socket
    import subprocess
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "gloo_iai.py"
    script.write_text(_GLOO_IAI.format(root=root, port=port, nranks=nranks))
    env = dict(os.environ, IAI_CPU_LANES=str(lanes))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env) for r in range(nranks)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_engine_matches_recursion_on_random_problems(orc, eng):
    """property test: on random small Wannier-like series, random complex frequencies, tolerances, limits and dimensions the
    level-synchronous engine (both modes) takes exactly the oracle recursion's decisions"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None, derandomize=True)
    @given(seed=st.integers(0, 10 ** 6), ndim=st.integers(1, 3), n=st.integers(1, 3), lkind=st.integers(0, 1), leaf=st.booleans(),
           vkind=st.integers(0, 1), tol=st.sampled_from([3e-1, 3e-2, 3e-3]), eta=st.sampled_from([0.3, 0.1, 0.05]))
    def check(seed, ndim, n, lkind, leaf, vkind, tol, eta):
        rng = np.random.default_rng(seed)
        M = (3,) * ndim + (1,) * (3 - ndim)
        c = rng.standard_normal((n, n) + M) + 1j * rng.standard_normal((n, n) + M)
        # Hermitian H(k): H_{-R} = H_R^dagger
        c = 0.5 * (c + np.conj(np.transpose(c[:, :, ::-1, ::-1, ::-1], (1, 0, 2, 3, 4))))
        S = orc.Series(c, (-1,) * ndim + (0,) * (3 - ndim))
        z = complex(rng.uniform(-1, 1), eta)
        la = [0.5] * ndim if lkind else [0.0] * ndim
        lb = None if lkind else [1.0] * ndim
        Io, Eo, neo = orc.iai(S, ndim, lkind, la, lb, vkind=vkind, z=z, atol=tol)
        rc, I, E, ne, _ = _solve(eng, S, ndim, lkind, la, lb, 0, vkind, z, None, tol, 0.0, leaf)
        assert rc == 0 and ne == neo
        assert abs(I - Io) <= 1e-12 * max(abs(Io), 1e-12) and abs(E - Eo) <= 1e-9 * max(Eo, 1e-300)

    check()


@pytest.mark.parametrize("leaf", [False, True])
def test_engine_lanes_do_not_change_results(orc, eng, svo, leaf, monkeypatch):
    """several rounds in flight (the children of the outermost panels dealt to lanes, each lane its own sequence of rounds):
    bit-identical integral, error estimate and numevals for 1, 3 and 5 lanes, cubic and tetrahedral limits, and fewer
    engine iterations are NOT required - only that every 1-D integral sees its own evaluations in its own order"""
    H, lo, A = svo
    S = orc.Series(H, lo)
    z = complex(12.5, 0.05)
    for args, atol in (((0, [0.0] * 3, [1.0] * 3), 3e-2), ((1, [0.5] * 3, None), 1e-3), ((0, [0.0] * 2, [1.0] * 2), 1e-3)):
        ndim = len(args[1])
        Sd = S if ndim == 3 else orc.Series(np.ascontiguousarray(H[:, :, :, :, H.shape[4] // 2:H.shape[4] // 2 + 1]), (lo[0], lo[1], 0))
        res = []
        for lanes in (1, 3, 5):
            monkeypatch.setenv("IAI_CPU_LANES", str(lanes))
            rc, I, E, ne, rounds = _solve(eng, Sd, ndim, args[0], args[1], args[2], 0, 1, z, None, atol, 0.0, leaf)
            assert rc == 0
            res.append((I, E, ne))
        assert res[0] == res[1] == res[2]
    monkeypatch.delenv("IAI_CPU_LANES", raising=False)


# ---- general iterated limits: several segments per level, limits served by a callback --------------------------------------
LIMITS_FN = C.CFUNCTYPE(C.c_int32, C.c_int32, c_dp, c_dp, C.c_int32, C.c_void_p)


def _solve_general(lib, S, ndim, lims, fkind, vkind, z, lin, atol, rtol, leaf, maxevals=2 ** 62, cap2=64, cap1=2048):
    import orc as _orc
    i3 = lambda v: (C.c_int * 3)(*[int(x) for x in v])
    d3 = lambda v: (C.c_double * 3)(*[float(x) for x in v])
    zz = np.array([z.real, z.imag])
    ln = None if lin is None else np.ascontiguousarray(lin, dtype=np.float64)
    out = np.zeros(3)
    st = (C.c_long * 6)()
    dp = lambda a: None if a is None else a.ctypes.data_as(c_dp)
    cb = C.cast(_orc.limits_callback(lims), LIMITS_FN)
    cb._keep = lims
    rc = lib.iai_cpu_solve_general(dp(S.c), C.c_int(S.n), C.c_int(ndim), i3(S.M), i3(S.lo), d3(S.period), cb, None,
                                   C.c_int(fkind), C.c_int(vkind), dp(zz), None, dp(ln), C.c_double(atol), C.c_double(rtol),
                                   C.c_long(maxevals), C.c_int(int(leaf)), C.c_long(cap2), C.c_long(cap1), C.c_int(0), C.c_int(1),
                                   XFN(0), dp(out), st)
    return rc, complex(out[0], out[1]), out[2], int(st[0]), int(st[1])


@pytest.mark.parametrize("leaf", [False, True])
def test_engine_general_limits_match_recursion(orc, eng, svo, leaf):
    """Several initial segments per 1-D integral (QuadGK's do_quadgk over a PuncturedInterval, src/fourier.jl:493-500) and limits
    obtained through the segments / fixandeliminate callback: the engine against the oracle's recursion over the SAME limits
    object - identical numevals; TetrahedralLimits served through the callback must reproduce the built-in kind bit for bit."""
    import autobz_b200 as ab
    H, lo, A = svo
    S = orc.Series(H, lo)
    z = complex(12.5, 0.05)
    seg = ab.SegmentedLimits((0.0, 0.2, 0.5), (0.0, 0.25, 0.3, 0.5), (0.0, 0.1, 0.5))
    Io, Eo, neo = orc.iai_general(S, 3, seg, vkind=1, z=z, atol=2e-3)
    rc, I, E, ne, _ = _solve_general(eng, S, 3, seg, 0, 1, z, None, 2e-3, 0.0, leaf)
    assert rc == 0 and ne == neo and abs(I - Io) <= 1e-13 * abs(Io) and abs(E - Eo) <= 1e-9 * Eo
    # one segment per level == CubicLimits
    one = ab.SegmentedLimits((0.0, 0.5), (0.0, 0.5), (0.0, 0.5))
    Ic, Ec, nec = orc.iai(S, 3, 0, [0.0] * 3, [0.5] * 3, vkind=1, z=z, atol=2e-3)
    rc, I, E, ne, _ = _solve_general(eng, S, 3, one, 0, 1, z, None, 2e-3, 0.0, leaf)
    assert rc == 0 and ne == nec and I == Ic and E == Ec
    # the tetrahedron through the callback == the built-in TetrahedralLimits
    tet = ab.TetrahedralLimits([0.5] * 3)
    rc0, I0, E0, ne0, _ = _solve(eng, S, 3, 1, [0.5] * 3, None, 0, 1, z, None, 1e-3, 0.0, leaf)
    rc, I, E, ne, _ = _solve_general(eng, S, 3, tet, 0, 1, z, None, 1e-3, 0.0, leaf)
    assert rc == rc0 == 0 and ne == ne0 and I == I0 and E == E0


def test_engine_polyhedron_volumes(orc, eng):
    """nested_quad(1, lims) over hand-built convex polyhedra = their volume (the reference checks its IBZ limits this way,
    test/test_ibz.jl:121-149): tetrahedron, cube, octahedron, a sheared prism - through PolyhedronLimits' slices and the callback"""
    import autobz_b200 as ab
    S = orc.Series(np.zeros((1, 1, 1, 1, 1), complex), (0, 0, 0))
    shapes = {
        "tetrahedron": ([(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)], 1 / 6),
        "cube": ([(x, y, z) for x in (0, 0.5) for y in (0, 0.5) for z in (0, 0.5)], 0.125),
        "octahedron": ([(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)], 4 / 3),
        "sheared prism": ([(0, 0, 0), (1, 0, 0), (0.3, 1, 0), (0.2, 0.1, 0.7), (1.2, 0.1, 0.7), (0.5, 1.1, 0.7)], 0.35),
    }
    for name, (verts, vol) in shapes.items():
        lims = ab.PolyhedronLimits(np.array(verts, dtype=float))
        rc, I, E, ne, _ = _solve_general(eng, S, 3, lims, 1, 2, 0j, [0.0, 0.0, 1.0, 0.0], 1e-10, 0.0, False)
        Io, Eo, neo = orc.iai_general(S, 3, lims, vkind=2, lin=(0.0, 1.0), atol=1e-10)
        assert rc == 0 and abs(I.real - vol) < 1e-8, (name, I, vol)
        assert ne == neo and abs(I - Io) <= 1e-13 * abs(Io)


@pytest.mark.parametrize("case", ["svo3d_tetra", "svo3d_cubic", "cos2d", "cos2d_rtol", "cos2d_maxevals", "svo3d_maxevals"])
def test_engine_lookahead_keeps_decisions_and_counts(orc, eng, svo, case):
    """Look-ahead on the outermost integral (abz_iai_engine.hpp, ABZ_IAI_SPECULATE): the bisection of the panel next in the heap is
    started together with the current one and parked until QuadGK's order reaches it.  Same integral, same error estimate, same
    numevals as the sequential recursion (what is never reached is not counted), fewer rounds."""
    H, lo, A = svo
    atol, rtol, mx = 0.0, 0.0, 2 ** 62
    if case.startswith("cos2d"):
        c = np.zeros((1, 1, 3, 3, 1), dtype=complex)
        c[0, 0, 0, 1, 0] = c[0, 0, 2, 1, 0] = c[0, 0, 1, 0, 0] = c[0, 0, 1, 2, 0] = 0.5
        S, ndim, z, vk, mode = orc.Series(c, (-1, -1, 0)), 2, complex(0.3, 0.005), 0, 1
        lk, la, lb = 0, [0.0, 0.0], [0.5, 0.5]
        if case == "cos2d":
            atol = 1e-6
        elif case == "cos2d_rtol":
            rtol = 1e-7
        else:
            atol, mx = 1e-9, 100
    else:
        S, ndim, vk, mode = orc.Series(H, lo), 3, 1, 2
        if case == "svo3d_cubic":
            lk, la, lb, z, atol = 0, [0.0] * 3, [0.5] * 3, complex(12.5, 0.02), 3e-5
        elif case == "svo3d_tetra":
            lk, la, lb, z, atol = 1, [0.5] * 3, None, complex(12.5, 0.005), 2e-5
        else:
            lk, la, lb, z, atol, mx = 1, [0.5] * 3, None, complex(12.5, 0.005), 1e-9, 100
    Io, Eo, neo = orc.iai(S, ndim, lk, la, lb, vkind=vk, z=z, atol=atol, rtol=rtol, maxevals=mx)
    rc0, I0, E0, ne0, rounds0 = _solve(eng, S, ndim, lk, la, lb, 0, vk, z, None, atol, rtol, mode, maxevals=mx, cap2=160)
    assert rc0 == 0 and ne0 == neo and abs(I0 - Io) <= 1e-13 * abs(Io)
    for policy in ((16, 32, 48) if ndim == 2 else (48,)):  # the panel next in the heap / the quarters of the bisected panel / both
        rc1, I1, E1, ne1, rounds1 = _solve(eng, S, ndim, lk, la, lb, 0, vk, z, None, atol, rtol, mode | 4 | policy, maxevals=mx, cap2=256)
        started, used = _solve.last_stats[4], _solve.last_stats[5]
        assert rc1 == 0 and ne1 == neo
        assert I1 == I0 and E1 == E0                       # bit-identical: the same values combined in the same order
        if "maxevals" in case:
            assert started > used                          # some were started ahead, never reached, and (ne1 == neo) never counted
        else:
            assert 0 < used <= started and rounds1 < rounds0
    # too small an arena for look-ahead: the engine simply does not speculate
    rc2, I2, E2, ne2, rounds2 = _solve(eng, S, ndim, lk, la, lb, 0, vk, z, None, atol, rtol, mode | 4, maxevals=mx, cap2=64, cap1=64)
    assert rc2 == 0 and ne2 == neo and I2 == I0 and _solve.last_stats[4] == 0
