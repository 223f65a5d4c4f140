"""ctypes binding of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does (it fails loudly without the CUDA library).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_lp = C.POINTER(C.c_long)


def build(force=False):
    """Compile liborc.so with gcc (oracle/Makefile) when missing or stale."""
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "autobz_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_quadgk_test.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.c_double, C.c_long, c_dp, c_lp]
        _LIB.orc_symptr_rule.restype = C.c_long
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


class Series:
    """H_R coefficients: complex array [n, n, M1, M2, M3] (Fortran order), lo = lowest R per dim."""

    def __init__(self, coeffs, lo, period=(1.0, 1.0, 1.0)):
        c = np.asarray(coeffs)
        if c.ndim != 5 or c.shape[0] != c.shape[1]:
            raise ValueError("coeffs must be [n, n, M1, M2, M3]")
        self.c = np.asfortranarray(c, dtype=np.complex128)
        self.n = c.shape[0]
        self.M = tuple(c.shape[2:])
        self.lo = tuple(int(x) for x in lo)
        self.period = tuple(float(x) for x in period)

    def args(self):
        return (_dp(self.c), C.c_int(self.n), _i3(self.M), _i3(self.lo), _d3(self.period))


def _z(zs):
    z = np.ascontiguousarray(np.atleast_1d(np.asarray(zs, dtype=np.complex128)))
    return z


def _sigma(sigma, n, nw):
    if sigma is None:
        return None, None
    s = np.asfortranarray(np.asarray(sigma, dtype=np.complex128))
    assert s.shape == (n, n, nw)
    return s, _dp(s)


def grid_eval_full(s, N, k3_lo=0, k3_hi=None, nthreads=0):
    k3_hi = N if k3_hi is None else k3_hi
    H = np.empty((s.n, s.n, N, N, k3_hi - k3_lo), dtype=np.complex128, order="F")
    rc = lib().orc_grid_eval_full(*s.args(), C.c_int(N), C.c_int(k3_lo), C.c_int(k3_hi), _dp(H), C.c_int(nthreads))
    assert rc == 0, rc
    return H


def eval_points(s, k):
    k = np.ascontiguousarray(np.asarray(k, dtype=np.float64).reshape(-1, 3))
    H = np.empty((s.n, s.n, k.shape[0]), dtype=np.complex128, order="F")
    rc = lib().orc_eval_points(*s.args(), C.c_long(k.shape[0]), _dp(k), _dp(H))
    assert rc == 0, rc
    return H


def ptr_sum(s, N, zs=None, sigma=None, fkind=0, scale=None, k3_lo=0, k3_hi=None, nthreads=0, k2_lo=0, k2_hi=None):
    """scale * sum_k f(H(k)) over the full N^3 PTR grid planes [k3_lo, k3_hi); default scale 1/N^3."""
    k3_hi = N if k3_hi is None else k3_hi
    z = _z(0.0 if zs is None else zs)
    nw = z.size
    sg, sgp = _sigma(sigma, s.n, nw)
    out = np.zeros(nw, dtype=np.complex128)
    scale = 1.0 / N ** 3 if scale is None else scale
    k2_hi = N if k2_hi is None else k2_hi
    rc = lib().orc_ptr_sum_rows(*s.args(), C.c_int(N), C.c_int(k3_lo), C.c_int(k3_hi), C.c_int(k2_lo), C.c_int(k2_hi),
                                C.c_int(fkind), C.c_int(nw), _dp(z), sgp, C.c_double(scale), _dp(out), C.c_int(nthreads))
    if rc:
        raise FloatingPointError("oracle: singular matrix / NaN in integrand")
    return out


def symptr_rule(N, syms):
    """syms: iterable of 3x3 integer matrices.  Returns (wsym int32 [N,N,N] Fortran order (i1 fastest), nirr)."""
    sy = np.ascontiguousarray(np.asarray(syms, dtype=np.int32).reshape(-1, 3, 3))
    w = np.zeros((N, N, N), dtype=np.int32, order="F")
    nirr = lib().orc_symptr_rule(C.c_int(N), C.c_int(sy.shape[0]), sy.ctypes.data_as(C.POINTER(C.c_int32)),
                                 w.ctypes.data_as(C.POINTER(C.c_int32)))
    return w, int(nirr)


def symptr_sum(s, N, wsym, zs=None, sigma=None, fkind=0, scale=1.0, nthreads=0):
    z = _z(0.0 if zs is None else zs)
    nw = z.size
    sg, sgp = _sigma(sigma, s.n, nw)
    out = np.zeros(nw, dtype=np.complex128)
    cnt = (C.c_long * 1)(0)
    w = np.asfortranarray(wsym, dtype=np.int32)
    rc = lib().orc_symptr_sum(*s.args(), C.c_int(N), w.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(fkind),
                              C.c_int(nw), _dp(z), sgp, C.c_double(scale), _dp(out), cnt, C.c_int(nthreads))
    if rc:
        raise FloatingPointError("oracle: singular matrix / NaN in integrand")
    return out, int(cnt[0])


def ptr_eig_sum(s, N, kind, params, wsym=None, scale=1.0, nthreads=0):
    prm = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
    out = np.zeros(1)
    cnt = (C.c_long * 1)(0)
    wp = None
    if wsym is not None:
        w = np.asfortranarray(wsym, dtype=np.int32)
        wp = w.ctypes.data_as(C.POINTER(C.c_int32))
    rc = lib().orc_ptr_eig_sum(*s.args(), C.c_int(N), wp, C.c_int(kind), _dp(prm), C.c_double(scale), _dp(out), cnt,
                               C.c_int(nthreads))
    assert rc == 0
    return float(out[0]), int(cnt[0])


def resolvent_trace_batch(H, zs, sigma=None, lu=False):
    H = np.asfortranarray(H, dtype=np.complex128)
    n = H.shape[0]
    nk = H.size // (n * n)
    z = _z(zs)
    sg, sgp = _sigma(sigma, n, z.size)
    out = np.zeros((nk, z.size), dtype=np.complex128)
    fn = lib().orc_resolvent_trace_lu_batch if lu else lib().orc_resolvent_trace_batch
    rc = fn(_dp(H), C.c_int(n), C.c_long(nk), C.c_int(z.size), _dp(z), sgp, _dp(out))
    if rc:
        raise FloatingPointError("oracle: singular matrix / NaN")
    return out


def eigvals_batch(H):
    H = np.asfortranarray(H, dtype=np.complex128)
    n = H.shape[0]
    nk = H.size // (n * n)
    w = np.zeros((nk, n))
    lib().orc_eigvals_batch(_dp(H), C.c_int(n), C.c_long(nk), _dp(w))
    return w


def quadgk_test(kind, a, b, p0=0.0, p1=0.0, atol=0.0, rtol=None, maxevals=10 ** 7):
    if rtol is None:
        rtol = np.sqrt(np.finfo(float).eps) if atol == 0 else 0.0
    out = np.zeros(3)
    ne = (C.c_long * 1)(0)
    rc = lib().orc_quadgk_test(kind, p0, p1, a, b, atol, rtol, maxevals, _dp(out), ne)
    assert rc == 0, rc
    return complex(out[0], out[1]), out[2], int(ne[0])


def gk15_nodes(a, b):
    xs = np.zeros(15)
    lib().orc_gk15_nodes(C.c_double(a), C.c_double(b), _dp(xs))
    return xs


def iai(s, dim, lkind, la, lb=None, vkind=0, z=0.0, sigma=None, lin=(1.0, 0.0), atol=0.0, rtol=None,
        maxevals=2 ** 62):
    """Nested adaptive GK(7,15) over CubicLimits (lkind 0) / TetrahedralLimits (lkind 1).
    The series must have singleton trailing dims for dim < 3.  Returns (I, E, numevals)."""
    if rtol is None:
        rtol = np.sqrt(np.finfo(float).eps) if atol == 0 else 0.0
    zz = _z(z)
    sgp = None
    if sigma is not None:
        sg = np.asfortranarray(np.asarray(sigma, dtype=np.complex128))
        sgp = _dp(sg)
    la = np.ascontiguousarray(np.asarray(la, dtype=np.float64))
    lbp = None
    if lb is not None:
        lb = np.ascontiguousarray(np.asarray(lb, dtype=np.float64))
        lbp = _dp(lb)
    linv = np.ascontiguousarray(np.asarray(lin, dtype=np.float64))
    out = np.zeros(3)
    ne = (C.c_long * 1)(0)
    rc = lib().orc_iai(_dp(s.c), C.c_int(s.n), C.c_int(dim), _i3(s.M), _i3(s.lo), _d3(s.period), C.c_int(lkind),
                       _dp(la), lbp, C.c_int(vkind), _dp(zz), sgp, _dp(linv), C.c_double(atol), C.c_double(rtol),
                       C.c_long(maxevals), _dp(out), ne)
    if rc:
        raise FloatingPointError("oracle: NaN/Inf in integrand (DomainError)")
    return complex(out[0], out[1]), out[2], int(ne[0])


LIMITS_FN = C.CFUNCTYPE(C.c_int, C.c_int, c_dp, c_dp, C.c_int, C.c_void_p)


def limits_callback(lims):
    """ctypes callback (orc_limits_fn) that serves `segments(fixandeliminate(...))` of a Python limits object with
    .segments() -> breakpoints and .fix(x) -> inner limits (the protocol of autobz_b200.bz)"""
    def fn(dim, xf, segs, maxseg, user):
        try:
            cur = lims
            for k in range(lims.ndim - dim):
                cur = cur.fix(xf[k])
            sg = [float(v) for v in cur.segments()]
            if len(sg) > maxseg or len(sg) < 2:
                return -1
            for i, v in enumerate(sg):
                segs[i] = v
            return len(sg)
        except Exception:
            return -2
    return LIMITS_FN(fn)


def iai_general(s, dim, lims, vkind=0, z=0.0, sigma=None, lin=(1.0, 0.0), atol=0.0, rtol=None, maxevals=2 ** 62):
    """orc_iai over general iterated limits (several breakpoints per level, fixandeliminate through a callback).
    Returns (I, E, numevals)."""
    if rtol is None:
        rtol = np.sqrt(np.finfo(float).eps) if atol == 0 else 0.0
    zz = _z(z)
    sgp = None
    if sigma is not None:
        sg = np.asfortranarray(np.asarray(sigma, dtype=np.complex128))
        sgp = _dp(sg)
    linv = np.ascontiguousarray(np.asarray(lin, dtype=np.float64))
    out = np.zeros(3)
    ne = (C.c_long * 1)(0)
    cb = limits_callback(lims)
    f = lib().orc_iai_general
    f.argtypes = [c_dp, C.c_int, C.c_int, C.c_int * 3, C.c_int * 3, C.c_double * 3, C.c_int, c_dp, c_dp, LIMITS_FN, C.c_void_p,
                  C.c_int, c_dp, c_dp, c_dp, C.c_double, C.c_double, C.c_long, c_dp, C.POINTER(C.c_long)]
    rc = f(_dp(s.c), s.n, dim, _i3(s.M), _i3(s.lo), _d3(s.period), 2, None, None, cb, None, vkind, _dp(zz), sgp, _dp(linv),
           atol, rtol, maxevals, _dp(out), ne)
    if rc:
        raise FloatingPointError(f"oracle: error {rc} (NaN/Inf in integrand or bad limits)")
    return complex(out[0], out[1]), out[2], int(ne[0])


def ggr_data(s, ndim, N, wsym=None):
    """get_ggr_data (src/dos_ggr.jl:14-44): (weights [nnodes], energies [nnodes, n], velocities [nnodes, ndim, n])"""
    nmax = N ** ndim
    wp = None
    if wsym is not None:
        w32 = np.asfortranarray(wsym, dtype=np.int32)
        wp = w32.ctypes.data_as(C.POINTER(C.c_int32))
        nmax = int(np.count_nonzero(w32))
    w = np.zeros(nmax); e = np.zeros((nmax, s.n)); v = np.zeros((nmax, ndim, s.n))
    lib().orc_ggr_data.restype = C.c_long
    cnt = lib().orc_ggr_data(*s.args()[:2], C.c_int(ndim), *s.args()[2:], C.c_int(N), wp, _dp(w), _dp(e), _dp(v))
    assert cnt == nmax, (cnt, nmax)
    return w, e, v


def ggr_sum(ndim, npt, E, w, e, v):
    """sum_ggr (src/dos_ggr.jl:58-65)"""
    Ev = np.ascontiguousarray(np.atleast_1d(np.asarray(E, dtype=np.float64)))
    out = np.zeros(Ev.size)
    rc = lib().orc_ggr_sum(C.c_int(ndim), C.c_int(npt), C.c_int(Ev.size), _dp(Ev), C.c_long(w.size), C.c_int(e.shape[1]),
                           _dp(np.ascontiguousarray(w)), _dp(np.ascontiguousarray(e)), _dp(np.ascontiguousarray(v)), _dp(out))
    assert rc == 0
    return out
