"""ctypes binding of libautobz_cuda.so (the C ABI in include/autobz_cuda.h).

This is the only place where the product touches native code.  There is no CPU fallback: if the
library is missing, cannot be loaded, or no B200 is present, every entry point raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libautobz_cuda.so")

ABZ_OK = 0
ABZ_E_INVALID, ABZ_E_OOM, ABZ_E_CUDA, ABZ_E_SINGULAR, ABZ_E_UNSUPPORTED, ABZ_E_NCCL = -1, -2, -3, -4, -5, -6
F_RESOLVENT_TRACE, F_TRACE_H = 0, 1
EIG_SUM, EIG_FERMI_ENERGY, EIG_FERMI_COUNT, EIG_GAUSS_DOS = 0, 1, 2, 3
OPT_RESOLVENT_ALGO, OPT_MEM_BUDGET_MB, OPT_FUSED_SMALL, OPT_EIG_ALGO, OPT_IAI_LEAF_SPILL, OPT_IAI_LANES = 1, 2, 3, 4, 5, 6
IAI_DEVICE_LEAVES, IAI_DEVICE_MIDDLES, IAI_SPECULATE = 1, 2, 4

# every symbol include/autobz_cuda.h declares (tests check that the library exports all of them)
EXPORTS = [
    "abz_version", "abz_last_error", "abz_ctx_create", "abz_ctx_destroy", "abz_ctx_set_option", "abz_ctx_launch_count",
    "abz_ctx_last_timings", "abz_series_create", "abz_series_destroy", "abz_rule_create_full", "abz_rule_create_sym", "abz_rule_create_nodes",
    "abz_symptr_rule", "abz_rule_create_symptr", "abz_rule_destroy", "abz_rule_info", "abz_rule_materialize", "abz_rule_copy_out",
    "abz_rule_resolvent_sum", "abz_rule_resolvent_matrix_sum", "abz_rule_eig_sum", "abz_rule_eig_sum_batch", "abz_rule_eigvals", "abz_rule_ggr_data", "abz_rule_ggr_sum", "abz_points_eval", "abz_points_resolvent",
    "abz_nest_create", "abz_nest_destroy", "abz_nest_contract3", "abz_nest_contract2", "abz_nest_eval", "abz_nest_eval_h", "abz_nest_eval_matrix", "abz_iai_solve", "abz_iai_solve_sharded", "abz_iai_solve_general",
    "abz_comm_unique_id", "abz_comm_init", "abz_allreduce_sum", "abz_comm_destroy",
]


class AutoBZCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libautobz_cuda error {code}: {msg}")
        self.code = code


class SingularIntegrandError(AutoBZCudaError, FloatingPointError):
    """NaN/Inf or a singular matrix in an integrand (QuadGK throws DomainError in the reference)."""


_lib = None
c_dp = C.POINTER(C.c_double)
EXCHANGE_FN = C.CFUNCTYPE(C.c_int32, C.POINTER(C.c_double), C.c_int64, C.c_void_p)   # abz_exchange_fn
LIMITS_FN = C.CFUNCTYPE(C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32, C.c_void_p)   # abz_limits_fn
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)


def load():
    """Load libautobz_cuda.so; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.abz_last_error.restype = C.c_char_p
    lib.abz_last_error.argtypes = [C.c_void_p]
    lib.abz_ctx_launch_count.restype = C.c_int64
    lib.abz_ctx_launch_count.argtypes = [C.c_void_p]
    lib.abz_ctx_create.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
    lib.abz_ctx_destroy.argtypes = [C.c_void_p]
    lib.abz_ctx_set_option.argtypes = [C.c_void_p, C.c_int32, C.c_int64]
    lib.abz_ctx_last_timings.argtypes = [C.c_void_p, c_dp, c_dp]
    lib.abz_series_create.argtypes = [C.c_void_p, c_dp, C.c_int32, C.c_int32, c_i32p, c_i32p, c_dp, C.POINTER(C.c_uint64)]
    lib.abz_series_destroy.argtypes = [C.c_void_p, C.c_uint64]
    lib.abz_rule_create_full.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]
    lib.abz_rule_create_sym.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, c_i32p, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]
    lib.abz_rule_create_nodes.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_int64, c_i32p, c_dp, C.POINTER(C.c_uint64)]
    lib.abz_symptr_rule.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_i32p, c_i32p, c_i64p]
    lib.abz_rule_create_symptr.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_int32, c_i32p, C.c_int32, C.c_int32,
                                           C.POINTER(C.c_uint64), c_i64p]
    lib.abz_rule_destroy.argtypes = [C.c_void_p, C.c_uint64]
    lib.abz_rule_info.argtypes = [C.c_void_p, C.c_uint64, c_i64p, c_i32p, c_i32p]
    lib.abz_rule_materialize.argtypes = [C.c_void_p, C.c_uint64]
    lib.abz_rule_copy_out.argtypes = [C.c_void_p, C.c_uint64, c_dp, c_dp, c_dp]
    lib.abz_rule_resolvent_sum.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_int32, c_dp, c_dp, C.c_double, c_dp]
    lib.abz_rule_resolvent_matrix_sum.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, c_dp, c_dp, C.c_double, c_dp]
    lib.abz_rule_eig_sum.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, c_dp, C.c_double, c_dp]
    lib.abz_rule_eig_sum_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_int32, c_dp, C.c_double, c_dp]
    lib.abz_rule_eigvals.argtypes = [C.c_void_p, C.c_uint64, c_dp]
    lib.abz_rule_ggr_data.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, c_dp, c_dp]
    lib.abz_rule_ggr_sum.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, c_dp, C.c_double, c_dp]
    lib.abz_points_eval.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, c_dp, c_dp]
    lib.abz_points_resolvent.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, c_dp, C.c_int32, C.c_int32, c_dp, c_dp, c_dp]
    lib.abz_nest_create.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_uint64)]
    lib.abz_nest_destroy.argtypes = [C.c_void_p, C.c_uint64]
    lib.abz_nest_contract3.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, c_dp, c_i64p]
    lib.abz_nest_contract2.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, c_dp, c_i64p, c_i64p]
    lib.abz_nest_eval.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, c_dp, c_i64p, C.c_int32, c_dp, c_dp, c_dp]
    lib.abz_nest_eval_h.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, c_dp, c_i64p, c_dp]
    lib.abz_nest_eval_matrix.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, c_dp, c_i64p, c_dp, c_dp, c_dp]
    lib.abz_iai_solve.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, c_dp, c_dp, C.c_int32, C.c_int32, c_dp, c_dp, c_dp, C.c_double,
                                  C.c_double, C.c_int64, C.c_int32, c_dp, c_i64p]
    lib.abz_iai_solve_sharded.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, c_dp, c_dp, C.c_int32, C.c_int32, c_dp, c_dp, c_dp, C.c_double,
                                          C.c_double, C.c_int64, C.c_int32, C.c_int32, C.c_int32, EXCHANGE_FN, C.c_void_p, c_dp, c_i64p]
    lib.abz_iai_solve_general.argtypes = [C.c_void_p, C.c_uint64, LIMITS_FN, C.c_void_p, C.c_int32, C.c_int32, c_dp, c_dp, c_dp, C.c_double,
                                          C.c_double, C.c_int64, C.c_int32, C.c_int32, C.c_int32, EXCHANGE_FN, C.c_void_p, c_dp, c_i64p]
    lib.abz_comm_unique_id.argtypes = [C.c_void_p]
    lib.abz_comm_init.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.abz_allreduce_sum.argtypes = [C.c_void_p, c_dp, C.c_int64]
    lib.abz_comm_destroy.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _cz(z):
    return np.ascontiguousarray(np.atleast_1d(np.asarray(z, dtype=np.complex128)))


class Context:
    """One CUDA stream on one B200 (abz_ctx).  Single-threaded; create one per Python thread."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = ABZ_OK
        for attempt in range(3):
            rc = self.lib.abz_ctx_create(int(device), C.byref(h))
            msg = "" if rc == ABZ_OK else self.lib.abz_last_error(None).decode()
            # a just-vacated GPU has been seen to refuse driver initialisation once; an absent device stays absent
            if rc != ABZ_E_CUDA or "initialization" not in msg or attempt == 2:
                break
            import time
            time.sleep(3.0)
        if rc != ABZ_OK:
            raise AutoBZCudaError(rc, self.lib.abz_last_error(None).decode())
        self.h = h
        self.device = device

    def check(self, rc):
        if rc == ABZ_OK:
            return
        msg = self.lib.abz_last_error(self.h).decode()
        if rc == ABZ_E_SINGULAR:
            raise SingularIntegrandError(rc, msg)
        if rc == ABZ_E_INVALID:
            raise ValueError(f"libautobz_cuda: {msg}")
        raise AutoBZCudaError(rc, msg)

    def set_option(self, opt, value):
        self.check(self.lib.abz_ctx_set_option(self.h, opt, int(value)))

    @property
    def launch_count(self):
        return int(self.lib.abz_ctx_launch_count(self.h))

    def last_timings(self):
        a, b = C.c_double(), C.c_double()
        self.check(self.lib.abz_ctx_last_timings(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.abz_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- symptr_rule on the device
    def symptr_rule(self, npt, syms):
        sy = np.ascontiguousarray(np.asarray(syms, dtype=np.int32).reshape(-1, 3, 3))
        w = np.zeros((npt, npt, npt), dtype=np.int32, order="F")
        nirr = C.c_int64()
        self.check(self.lib.abz_symptr_rule(self.h, npt, sy.shape[0], sy.ctypes.data_as(c_i32p), w.ctypes.data_as(c_i32p),
                                            C.byref(nirr)))
        return w, int(nirr.value)

    # -- communicator
    def comm_init(self, rank, nranks, uid):
        self.check(self.lib.abz_comm_init(self.h, rank, nranks, uid))

    def allreduce_sum(self, arr):
        a = np.ascontiguousarray(arr).view(np.float64).reshape(-1)
        self.check(self.lib.abz_allreduce_sum(self.h, _dp(a), a.size))
        return arr


def comm_unique_id():
    buf = C.create_string_buffer(128)
    rc = load().abz_comm_unique_id(buf)
    if rc != ABZ_OK:
        raise AutoBZCudaError(rc, load().abz_last_error(None).decode())
    return buf.raw


class DeviceSeries:
    """H_R on the device.  coeffs: [n, n, M1(, M2(, M3))] complex or real, Fortran order semantics
    (numpy array indexed [a, b, i1, i2, i3]); lo = lowest R index per dimension."""

    def __init__(self, ctx, coeffs, lo, period):
        c = np.asarray(coeffs)
        if c.ndim < 3 or c.ndim > 5 or c.shape[0] != c.shape[1]:
            raise ValueError("coeffs must have shape [n, n, M1(, M2(, M3))]")
        self.ndim = c.ndim - 2
        while c.ndim < 5:
            c = c[..., None]
        lo = tuple(int(x) for x in lo) + (0,) * (3 - len(tuple(lo)))
        period = tuple(float(x) for x in np.atleast_1d(period)) + (1.0,) * (3 - len(tuple(np.atleast_1d(period))))
        self.ctx = ctx
        self.n = c.shape[0]
        self.M = tuple(int(m) for m in c.shape[2:])
        self.lo, self.period = lo[:3], period[:3]
        is_complex = np.iscomplexobj(c)
        buf = np.asfortranarray(c, dtype=np.complex128 if is_complex else np.float64)
        h = C.c_uint64()
        M = (C.c_int32 * 3)(*self.M)
        lo_ = (C.c_int32 * 3)(*self.lo)
        per = (C.c_double * 3)(*self.period)
        ctx.check(ctx.lib.abz_series_create(ctx.h, _dp(buf), int(is_complex), self.n, M, lo_, per, C.byref(h)))
        self.h = h.value
        self.h2d_bytes = buf.nbytes

    def close(self):
        if self.h is not None and self.ctx.h:
            self.ctx.lib.abz_series_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def eval_points(self, k):
        k = np.ascontiguousarray(np.asarray(k, dtype=np.float64).reshape(-1, 3))
        H = np.empty((self.n, self.n, k.shape[0]), dtype=np.complex128, order="F")
        self.ctx.check(self.ctx.lib.abz_points_eval(self.ctx.h, self.h, k.shape[0], _dp(k), _dp(H)))
        return H

    def points_resolvent(self, k, z, sigma=None, fkind=F_RESOLVENT_TRACE):
        k = np.ascontiguousarray(np.asarray(k, dtype=np.float64).reshape(-1, 3))
        zz = _cz(z)
        nw = 1 if fkind == F_TRACE_H else zz.size
        sg = None if sigma is None else np.asfortranarray(np.asarray(sigma, dtype=np.complex128).reshape(self.n, self.n, nw))
        y = np.empty((k.shape[0], nw), dtype=np.complex128)
        self.ctx.check(self.ctx.lib.abz_points_resolvent(self.ctx.h, self.h, k.shape[0], _dp(k), fkind, nw, _dp(zz), _dp(sg), _dp(y)))
        return y


class DeviceRule:
    """A quadrature rule on the device: FourierPTR (full grid) or FourierMonkhorstPack (wsym given)."""

    def __init__(self, ctx, series, npt, wsym=None, k3_lo=0, k3_hi=None, k3_stride=1, nodes=None, weights=None, syms=None, count_all=True):
        self.ctx, self.series, self.npt = ctx, series, int(npt)
        h = C.c_uint64()
        self.nirr_total = None
        if syms is not None:
            sy = np.ascontiguousarray(np.asarray(syms, dtype=np.int32).reshape(-1, 3, 3))
            nirr = C.c_int64()
            # count_all = False: nirr_total = NULL, the library then builds the orbit weights of the selected planes only
            ctx.check(ctx.lib.abz_rule_create_symptr(ctx.h, series.h, self.npt, sy.shape[0], sy.ctypes.data_as(c_i32p), int(k3_lo),
                                                     int(k3_stride), C.byref(h), C.byref(nirr) if count_all else None))
            self.nirr_total = int(nirr.value) if count_all else None
        elif nodes is not None:
            idx = np.ascontiguousarray(np.asarray(nodes, dtype=np.int32).reshape(-1, 3))
            wv = None if weights is None else np.ascontiguousarray(np.asarray(weights, dtype=np.float64))
            ctx.check(ctx.lib.abz_rule_create_nodes(ctx.h, series.h, self.npt, idx.shape[0], idx.ctypes.data_as(c_i32p), _dp(wv),
                                                    C.byref(h)))
        elif wsym is None:
            k3_hi = self.npt if k3_hi is None else k3_hi
            ctx.check(ctx.lib.abz_rule_create_full(ctx.h, series.h, self.npt, int(k3_lo), int(k3_hi), C.byref(h)))
        else:
            w = np.asfortranarray(wsym, dtype=np.int32)
            if w.shape != (self.npt,) * 3:
                raise ValueError("wsym must be [npt, npt, npt]")
            ctx.check(ctx.lib.abz_rule_create_sym(ctx.h, series.h, self.npt, w.ctypes.data_as(c_i32p), int(k3_lo), int(k3_stride),
                                                  C.byref(h)))
        self.h = h.value
        nn, no, npt_ = C.c_int64(), C.c_int32(), C.c_int32()
        ctx.check(ctx.lib.abz_rule_info(ctx.h, self.h, C.byref(nn), C.byref(no), C.byref(npt_)))
        self.nnodes = int(nn.value)
        self.materialized = False

    def __len__(self):
        return self.nnodes

    def close(self):
        if self.h is not None and self.ctx.h:
            self.ctx.lib.abz_rule_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def materialize(self):
        self.ctx.check(self.ctx.lib.abz_rule_materialize(self.ctx.h, self.h))
        self.materialized = True

    def copy_out(self, want_h=True):
        n = self.series.n
        H = np.empty((n, n, self.nnodes), dtype=np.complex128, order="F") if want_h else None
        k = np.empty((self.nnodes, 3))
        w = np.empty(self.nnodes)
        self.ctx.check(self.ctx.lib.abz_rule_copy_out(self.ctx.h, self.h, _dp(H), _dp(k), _dp(w)))
        return H, k, w

    def resolvent_sum(self, z, sigma=None, scale=1.0, fkind=F_RESOLVENT_TRACE):
        zz = _cz(z)
        nw = 1 if fkind == F_TRACE_H else zz.size
        n = self.series.n
        sg = None if sigma is None else np.asfortranarray(np.asarray(sigma, dtype=np.complex128).reshape(n, n, nw))
        out = np.empty(nw, dtype=np.complex128)
        self.ctx.check(self.ctx.lib.abz_rule_resolvent_sum(self.ctx.h, self.h, fkind, nw, _dp(zz), _dp(sg), float(scale), _dp(out)))
        return out

    def resolvent_matrix_sum(self, z, sigma=None, scale=1.0):
        """scale * sum_i w_i (z_w - H(k_i) - Sigma_w)^-1 -> [nw, n, n]"""
        zz = _cz(z)
        nw, n = zz.size, self.series.n
        sg = None if sigma is None else np.asfortranarray(np.asarray(sigma, dtype=np.complex128).reshape(n, n, nw))
        out = np.empty((n, n, nw), dtype=np.complex128, order="F")
        self.ctx.check(self.ctx.lib.abz_rule_resolvent_matrix_sum(self.ctx.h, self.h, nw, _dp(zz), _dp(sg), float(scale), _dp(out)))
        return np.ascontiguousarray(np.moveaxis(out, 2, 0))

    def eig_sum(self, kind, params=(0.0, 1.0), scale=1.0):
        prm = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
        out = np.zeros(1)
        self.ctx.check(self.ctx.lib.abz_rule_eig_sum(self.ctx.h, self.h, int(kind), _dp(prm), float(scale), _dp(out)))
        return float(out[0])

    def eig_sum_batch(self, kind, params, scale=1.0):
        """scale * sum_i w_i g_p(eigvals(H(k_i))) for every parameter pair in params [nparams, 2]: one diagonalisation per node"""
        prm = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1, 2))
        out = np.zeros(prm.shape[0])
        self.ctx.check(self.ctx.lib.abz_rule_eig_sum_batch(self.ctx.h, self.h, int(kind), prm.shape[0], _dp(prm), float(scale), _dp(out)))
        return out

    def eigvals(self):
        ev = np.empty((self.nnodes, self.series.n))
        self.ctx.check(self.ctx.lib.abz_rule_eigvals(self.ctx.h, self.h, _dp(ev)))
        return ev

    def ggr_data(self, ndim, copy=True):
        """get_ggr_data on the device: (energies [nnodes, n] ascending, velocities [nnodes, ndim, n]); cached in the rule"""
        n = self.series.n
        e = np.empty((self.nnodes, n)) if copy else None
        v = np.empty((self.nnodes, ndim, n)) if copy else None
        self.ctx.check(self.ctx.lib.abz_rule_ggr_data(self.ctx.h, self.h, int(ndim), _dp(e), _dp(v)))
        return e, v

    def ggr_sum(self, E, scale=1.0):
        """sum_ggr over the rule's nodes for every energy in E (needs ggr_data first)"""
        Ev = np.ascontiguousarray(np.atleast_1d(np.asarray(E, dtype=np.float64)))
        out = np.empty(Ev.size)
        self.ctx.check(self.ctx.lib.abz_rule_ggr_sum(self.ctx.h, self.h, Ev.size, _dp(Ev), float(scale), _dp(out)))
        return out



class DeviceNest:
    """Arena of contracted series for IAI panels (abz_nest_*)."""

    def __init__(self, ctx, series, ndim, cap2, cap1):
        self.ctx, self.series, self.ndim = ctx, series, ndim
        self.cap2, self.cap1 = int(cap2), int(cap1)
        h = C.c_uint64()
        ctx.check(ctx.lib.abz_nest_create(ctx.h, series.h, ndim, self.cap2, self.cap1, C.byref(h)))
        self.h = h.value

    def close(self):
        if self.h is not None and self.ctx.h:
            self.ctx.lib.abz_nest_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def contract3(self, x3, slot2):
        x = np.ascontiguousarray(x3, dtype=np.float64)
        s = np.ascontiguousarray(slot2, dtype=np.int64)
        self.ctx.check(self.ctx.lib.abz_nest_contract3(self.ctx.h, self.h, x.size, _dp(x), s.ctypes.data_as(c_i64p)))

    def contract2(self, x2, parent, slot1):
        x = np.ascontiguousarray(x2, dtype=np.float64)
        s = np.ascontiguousarray(slot1, dtype=np.int64)
        p = None if parent is None else np.ascontiguousarray(parent, dtype=np.int64)
        self.ctx.check(self.ctx.lib.abz_nest_contract2(self.ctx.h, self.h, x.size, _dp(x),
                                                       None if p is None else p.ctypes.data_as(c_i64p), s.ctypes.data_as(c_i64p)))

    def eval(self, x1, slot1, z, sigma=None, fkind=F_RESOLVENT_TRACE):
        x = np.ascontiguousarray(x1, dtype=np.float64)
        s = None if slot1 is None else np.ascontiguousarray(slot1, dtype=np.int64)
        zz = _cz(z)
        n = self.series.n
        sg = None if sigma is None else np.asfortranarray(np.asarray(sigma, dtype=np.complex128).reshape(n, n))
        y = np.empty(x.size, dtype=np.complex128)
        self.ctx.check(self.ctx.lib.abz_nest_eval(self.ctx.h, self.h, x.size, _dp(x), None if s is None else s.ctypes.data_as(c_i64p),
                                                  fkind, _dp(zz), _dp(sg), _dp(y)))
        return y

    def eval_matrix(self, x1, slot1, z, sigma=None):
        """(z - H - Sigma)^-1 at the nodes of innermost panels, [npts, n, n] (matrix-valued gloc_integrand under IAI)"""
        x = np.ascontiguousarray(x1, dtype=np.float64)
        s = None if slot1 is None else np.ascontiguousarray(slot1, dtype=np.int64)
        zz = _cz(z)
        n = self.series.n
        sg = None if sigma is None else np.asfortranarray(np.asarray(sigma, dtype=np.complex128).reshape(n, n))
        Y = np.empty((n, n, x.size), dtype=np.complex128, order="F")
        self.ctx.check(self.ctx.lib.abz_nest_eval_matrix(self.ctx.h, self.h, x.size, _dp(x), None if s is None else s.ctypes.data_as(c_i64p),
                                                         _dp(zz), _dp(sg), _dp(Y)))
        return np.ascontiguousarray(np.moveaxis(Y, 2, 0))

    def eval_h(self, x1, slot1):
        """H at the nodes of innermost panels, [n, n, npts] (for integrands evaluated on the host)"""
        x = np.ascontiguousarray(x1, dtype=np.float64)
        s = None if slot1 is None else np.ascontiguousarray(slot1, dtype=np.int64)
        n = self.series.n
        H = np.empty((n, n, x.size), dtype=np.complex128, order="F")
        self.ctx.check(self.ctx.lib.abz_nest_eval_h(self.ctx.h, self.h, x.size, _dp(x), None if s is None else s.ctypes.data_as(c_i64p),
                                                    _dp(H)))
        return H

    def iai_solve(self, lkind, la, lb, fkind, vkind, z, sigma, lin, atol, rtol, maxevals, device_leaves=True, rank=0, nranks=1,
                  allreduce=None, limits=None, device_middles=True, speculate=True):
        """abz_iai_solve(_sharded): the whole nested adaptive solve with the control flow on the library's host side.
        nranks > 1: outermost panel nodes dealt round-robin to the ranks; `allreduce(np.ndarray) -> np.ndarray` sums over
        the ranks (None = NCCL on the ctx communicator).  Returns (I complex, E, numevals, rounds, launches)."""
        la_ = None if la is None else np.ascontiguousarray(la, dtype=np.float64)
        lb_ = None if lb is None else np.ascontiguousarray(lb, dtype=np.float64)
        zz = _cz(0j if z is None else z)
        n = self.series.n
        sg = None if sigma is None else np.asfortranarray(np.asarray(sigma, dtype=np.complex128).reshape(n, n))
        ln = None
        if lin is not None:
            a, b = complex(lin[0]), complex(lin[1])
            ln = np.array([a.real, a.imag, b.real, b.imag], dtype=np.float64)
        out = np.zeros(3)
        stats = np.zeros(4, dtype=np.int64)
        err = []

        def _xchg(buf, cnt, user):
            try:
                a = np.ctypeslib.as_array(buf, shape=(cnt,))
                a[:] = np.asarray(allreduce(a.copy()), dtype=np.float64).reshape(-1)
                return 0
            except Exception as e:      # never let an exception cross the C boundary
                err.append(e)
                return -1

        xfn = EXCHANGE_FN(_xchg) if (nranks > 1 and allreduce is not None) else EXCHANGE_FN(0)
        flags = (IAI_DEVICE_LEAVES | (IAI_DEVICE_MIDDLES if device_middles else 0) | (IAI_SPECULATE if speculate else 0)) if device_leaves else 0
        if limits is not None:
            # general iterated limits (abz_iai_solve_general): `limits` has .ndim, .segments() -> breakpoints and .fix(x) -> inner limits
            # (segments / fixandeliminate of the reference's AbstractIteratedLimits); the library asks for the breakpoints of one level
            def _lims(dim, xf, segs, maxseg, user):
                try:
                    cur = limits
                    for k in range(limits.ndim - dim):
                        cur = cur.fix(xf[k])
                    sg_ = [float(v) for v in cur.segments()]
                    if not 2 <= len(sg_) <= maxseg:
                        raise ValueError(f"limits returned {len(sg_)} breakpoints (2 ... {maxseg} supported)")
                    for i, v in enumerate(sg_):
                        segs[i] = v
                    return len(sg_)
                except Exception as e:      # never let an exception cross the C boundary
                    err.append(e)
                    return -1

            lfn = LIMITS_FN(_lims)
            rc = self.ctx.lib.abz_iai_solve_general(self.ctx.h, self.h, lfn, None, int(fkind), int(vkind), _dp(zz), _dp(sg), _dp(ln),
                                                    float(atol), float(rtol), int(min(maxevals, 2 ** 62)), flags, int(rank), int(nranks),
                                                    xfn, None, _dp(out), stats.ctypes.data_as(c_i64p))
        else:
            rc = self.ctx.lib.abz_iai_solve_sharded(self.ctx.h, self.h, int(lkind), _dp(la_), _dp(lb_), int(fkind), int(vkind), _dp(zz),
                                                    _dp(sg), _dp(ln), float(atol), float(rtol), int(min(maxevals, 2 ** 62)), flags,
                                                    int(rank), int(nranks), xfn, None, _dp(out), stats.ctypes.data_as(c_i64p))
        if err:
            raise err[0]
        self.ctx.check(rc)
        self.last_exchanges = int(stats[3])
        return complex(out[0], out[1]), float(out[2]), int(stats[0]), int(stats[1]), int(stats[2])
