"""IntegralProblem / init / solve / IntegralSolver / batchsolve — host-side mirror of
src/interfaces.jl:34-243 and of the FourierIntegrand dispatch in src/fourier.jl:323-530 and
src/brillouin.jl:321-499, with the arithmetic handed to libautobz_cuda.so through backend.py.

Control flow kept on the host exactly as in the reference: tolerance rescaling by j = |det B| and
nsyms (src/brillouin.jl:337-355, 429-444), symmetrisation of scalar results (TrivialRep, :98-107),
AutoPTR's additive grid refinement and convergence test (AutoSymPTR.autosymptr, call site
src/algorithms.jl:418-432), IAI's nested GK panels (iai.py)."""
import time

import math

import numpy as np

from . import _lib
from .algorithms import (IAI, PTR, AbsoluteEstimate, AutoPTR, AutoSymPTRJL, AuxQuadGKJL, EvalCounter, MonkhorstPack, NestedQuad,
                         monkhorst_pack_schedule)
from .backend import DeviceBackend
from .bz import CubicLimits, SymmetricBZ, TetrahedralLimits
from .fourier import FourierIntegrand, FourierValue
from .iai import NestedGK


class Basis:
    """AutoSymPTR.Basis(B): the PTR domain (canonical_ptr_basis = identity, src/brillouin.jl:10)."""

    def __init__(self, B):
        self.B = np.array(B, dtype=float)

    @property
    def ndim(self):
        return self.B.shape[0]


class IntegralSolution:
    """IntegralSolution{u, resid, retcode, numevals} (src/interfaces.jl:120-126); numevals = -1 if undefined."""

    def __init__(self, u, resid, retcode=True, numevals=-1):
        self.u, self.resid, self.retcode, self.numevals = u, resid, retcode, numevals

    def __repr__(self):
        return f"IntegralSolution(u={self.u!r}, resid={self.resid!r}, retcode={self.retcode}, numevals={self.numevals})"


class IntegralProblem:
    """IntegralProblem(f, domain, p=NullParameters()) (src/interfaces.jl:34-48)"""

    def __init__(self, f, dom, p=None):
        self.f, self.dom, self.p = f, dom, p


class Shard:
    """This rank's share of a multi-GPU solve: k3 planes are partitioned across ranks (contiguous blocks
    for full grids, round-robin for symmetry-reduced ones, as the reference's thread chunks,
    src/fourier.jl:156, 246-255) and partial sums meet in ONE small allreduce per rule evaluation."""

    def __init__(self, rank=0, nranks=1, allreduce=None):
        self.rank, self.nranks = int(rank), int(nranks)
        self._allreduce = allreduce

    def allreduce(self, arr):
        if self.nranks == 1:
            return arr
        if self._allreduce is None:
            raise RuntimeError("Shard with nranks > 1 needs an allreduce function")
        return self._allreduce(arr)


def _shard_allreduce(shard):
    """the shard's allreduce for rule construction (node counts of a symmetry-reduced rule built plane-wise), None on one rank"""
    return shard.allreduce if shard.nranks > 1 else None


def torch_allreduce(device=None):
    """allreduce over torch.distributed (NCCL over NVLink on GPU ranks, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    def f(arr):
        a = np.ascontiguousarray(arr)
        t = torch.from_numpy(a.view(np.float64).reshape(-1).copy())
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy().view(a.dtype).reshape(a.shape)

    return f


_CHECK = ("abstol", "reltol", "maxiters")


def checkkwargs(kws):
    """src/interfaces.jl:64-69"""
    for k in kws:
        if k not in _CHECK:
            raise ValueError(f"keyword {k} unrecognized")


def _norm(x):
    return float(np.linalg.norm(np.atleast_1d(x)))


def _alg_norm(salg):
    """the algorithm's norm; the default `abs` stands for LinearAlgebra.norm (Frobenius on matrix-valued integrals)"""
    return _norm if salg.norm is abs else salg.norm


# ---------------------------------------------------------------------------------------------------
# parameters (MixedParameters, src/parameters.jl:11-35): ((args...), {kws...})
def _params(p):
    if p is None:
        return ((), {})
    if isinstance(p, MixedParameters):
        return (p.args, p.kwargs)
    if isinstance(p, tuple) and len(p) == 2 and isinstance(p[1], dict):
        return p
    if isinstance(p, dict):
        return ((), p)
    if isinstance(p, (tuple, list)):
        return (tuple(p), {})
    return ((p,), {})


class MixedParameters:
    """MixedParameters(args...; kwargs...) (src/parameters.jl:11-35): positional and keyword parameters of an integrand,
    with the reference's access (`p[i]` positional - 0-based here -, `p.name` / `p["name"]` keyword) and `merge`.
    Everywhere a parameter is expected, a MixedParameters, a plain ((args...), {kwargs}) pair, a dict, a tuple or a scalar
    are accepted (`_params` normalises them)."""

    def __init__(self, *args, **kwargs):
        object.__setattr__(self, "args", tuple(args))
        object.__setattr__(self, "kwargs", dict(kwargs))

    def __getitem__(self, i):
        return self.kwargs[i] if isinstance(i, str) else self.args[i]

    def __getattr__(self, name):
        try:
            return object.__getattribute__(self, "kwargs")[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        raise AttributeError("MixedParameters is immutable")

    def __eq__(self, other):
        return isinstance(other, MixedParameters) and self.args == other.args and self.kwargs == other.kwargs

    def __hash__(self):
        return hash((self.args, tuple(sorted(self.kwargs.items(), key=lambda kv: kv[0]))))

    def merge(self, q):
        """merge(p, q) (src/parameters.jl:26-35): MixedParameters / dicts (NamedTuples) merge keywords (right wins) and
        append positional parameters; tuples append their entries; any other value is appended as one positional parameter"""
        if isinstance(q, MixedParameters):
            return MixedParameters(*(self.args + q.args), **{**self.kwargs, **q.kwargs})
        if isinstance(q, tuple) and len(q) == 2 and isinstance(q[1], dict):
            return MixedParameters(*(self.args + tuple(q[0])), **{**self.kwargs, **q[1]})
        if isinstance(q, dict):
            return MixedParameters(*self.args, **{**self.kwargs, **q})
        if isinstance(q, (tuple, list)):
            return MixedParameters(*(self.args + tuple(q)), **self.kwargs)
        return MixedParameters(*(self.args + (q,)), **self.kwargs)

    def __repr__(self):
        kw = ", ".join(f"{k}={v!r}" for k, v in self.kwargs.items())
        return "MixedParameters(" + ", ".join([*(repr(a) for a in self.args), *([kw] if kw else [])]) + ")"


def merge(p, q):
    """Base.merge on parameters (src/parameters.jl:26-35)"""
    a, k = _params(p)
    return MixedParameters(*a, **k).merge(q)


def paramzip(*args, **kwargs):
    """paramzip(args...; kwargs...) (src/parameters.jl:37-56): result[i] = MixedParameters(args[0][i], ...; name=kwargs[name][i])"""
    cols = [list(a) for a in args] + [list(v) for v in kwargs.values()]
    if not cols:
        return []
    n = min(len(c) for c in cols)                     # zip semantics
    names = list(kwargs)
    return [MixedParameters(*(a[i] for a in (list(x) for x in args)), **{k: list(kwargs[k])[i] for k in names}) for i in range(n)]


def paramproduct(*args, **kwargs):
    """paramproduct(args...; kwargs...) (src/parameters.jl:58-69): the Cartesian product as an object array of shape
    (len(args[0]), ..., len(kwargs[...]), ...), element [i1, ..., in] = MixedParameters(args[0][i1], ...; name=kwargs[name][in])"""
    cols = [list(a) for a in args] + [list(v) for v in kwargs.values()]
    names = list(kwargs)
    na = len(args)
    out = np.empty(tuple(len(c) for c in cols), dtype=object)
    for idx in np.ndindex(*out.shape):
        vals = [c[i] for c, i in zip(cols, idx)]
        out[idx] = MixedParameters(*vals[:na], **dict(zip(names, vals[na:])))
    return out


class _BoundIntegrand:
    """A FourierIntegrand with all parameters known, reduced to what the device needs."""

    def __init__(self, f, plist):
        self.f = f
        self.native = f.native
        self.plist = [_params(p) for p in plist]
        if self.native:
            self.bound = [f.f.bind(*f.merged(p)) for p in self.plist]
            self.is_eig = f.f.is_eig
            self.fkind = f.f.fkind

    def rule_sums(self, rule):
        """sum_i w_i f(x_i) over the rule's local nodes for every parameter -> list of values"""
        f = self.f
        if not self.native:
            # generic / batch user integrand on the host (S3 seam, src/batch.jl:4-20): H(k) comes back from the device
            H, k, w = rule.copy_out()
            return [np.sum(w * f.host_values(H, k[:, :f.s.ndim], p)) if H.shape[2] else 0.0 for p in self.plist]
        if self.is_eig:
            if hasattr(rule, "eig_sum_batch"):      # every H(k) diagonalised once for all parameters
                return list(rule.eig_sum_batch(f.f.kind, [tuple(b) for b in self.bound]))
            return [rule.eig_sum(f.f.kind, b) for b in self.bound]
        if self.fkind == _lib.F_TRACE_H:
            t = rule.resolvent_sum(None, None, _lib.F_TRACE_H)[0]
            return [("affine", t, b) for b in self.bound]
        zs = np.array([b[0] for b in self.bound], dtype=np.complex128)
        sig = None
        if any(b[1] is not None for b in self.bound):
            n = f.s.norb
            sig = np.zeros((n, n, len(self.bound)), dtype=np.complex128, order="F")
            for i, b in enumerate(self.bound):
                if b[1] is not None:
                    sig[:, :, i] = np.asarray(b[1], dtype=np.complex128).reshape(n, n)
        if f.f.is_matrix:
            return list(rule.resolvent_matrix_sum(zs, sig))
        y = rule.resolvent_sum(zs, sig, _lib.F_RESOLVENT_TRACE)
        return list(y)


WARN_UNKNOWN_SYMMETRY = ("A symmetric BZ was used with an integrand whose symmetry representation is unknown.\n"
                         "For correctness, the calculation will be repeated on the full BZ.\n"
                         "However, it is better either to integrate without symmetries or to use symmetries by extending "
                         "SymRep for your type.")     # src/brillouin.jl:332-336


def _rule_apply(rule, bf, shard, ndim):
    """(rule)(f, B, buffer) = quadsum(rule, f, vol/(npt^d nsyms)) (src/fourier.jl:204-207, 289-292), then
    SymmetricRule's symmetrize (TrivialRep: x nsyms, src/brillouin.jl:107,127-130) — i.e. sum_i w_i f_i / npt^d
    for scalar integrands.  Multi-rank: local partial sums + one allreduce."""
    sums = bf.rule_sums(rule)
    npt_d = float(rule.npt) ** ndim
    if bf.native and not bf.is_eig and bf.fkind == _lib.F_TRACE_H:
        # affine integrand a*tr(H)+b: sum_i w_i (a t_i + b) = a * sum_i w_i t_i + b * npt^d (sum w_i = npt^d)
        t = np.array([sums[0][1]], dtype=np.complex128)
        t = shard.allreduce(t)[0]
        vals = [bd[0] * t / npt_d + bd[1] for (_, _, bd) in sums]
        return vals
    arr = np.array(sums)
    if np.iscomplexobj(arr):
        arr = arr.astype(np.complex128)
    else:
        arr = arr.astype(np.float64)
    arr = shard.allreduce(arr)
    vals = arr / npt_d
    if bf.native and not bf.is_eig and bf.f.f.is_matrix:
        # matrix-valued: the rule output is sum_i w_i f_i vol/(npt^d nsyms); the integrand's SymRep maps it to the FBZ
        # (symmetrize(f, bz, x), src/brillouin.jl:86-107,127-130)
        nsym = getattr(rule, "nsyms", 1)
        if nsym > 1:
            return [bf.f.f.symmetrize(bf.bz, v / nsym) for v in vals]
        return list(vals)
    if bf.native and not bf.is_eig:
        vals = bf.f.f.post(vals, None)
    return list(vals)


# ---------------------------------------------------------------------------------------------------
class IntegralCache:
    """IntegralCache (src/interfaces.jl:50-57): problem + algorithm + cacheval + kwargs, reusable across parameters."""

    def __init__(self, f, dom, p, alg, kwargs, backend, shard):
        self.f, self.dom, self.p, self.alg, self.kwargs = f, dom, p, alg, kwargs
        self.backend, self.shard = backend, shard
        self.cacheval = {}

    def close(self):
        """Release the device objects of the cache (rules, arena) now instead of when Python collects the cache: a one-shot
        `solve` must hand its multi-GB rule buffers back to the library's pool before the next solve asks for them."""
        for v in list(self.cacheval.values()):
            for o in (v if isinstance(v, (list, tuple)) else (v,)):
                if isinstance(o, IntegralCache) or (hasattr(o, "close") and not isinstance(o, type)):
                    try:
                        o.close()
                    except Exception:
                        pass
        self.cacheval = {}


def init(prob, alg, backend=None, shard=None, **kwargs):
    """init(prob, alg; kwargs...) (src/interfaces.jl:78-82): build the cache (rules / arena) once.
    DOSProblem + DOSAlgorithm dispatch to dos.dos_init (src/dos_interfaces.jl:84-88)."""
    from . import dos
    if isinstance(prob, dos.DOSProblem):
        return dos.dos_init(prob, alg, backend=backend, shard=shard, **kwargs)
    checkkwargs(kwargs)
    if not isinstance(prob.f, FourierIntegrand):
        raise TypeError("autobz_b200 implements the FourierIntegrand hot path only (SURVEY.md §8)")
    backend = backend if backend is not None else DeviceBackend()
    shard = shard if shard is not None else Shard()
    inner, counter = _unwrap(alg)
    if isinstance(inner, AbsoluteEstimate):
        # init_cacheval(f, dom, p, ::AbsoluteEstimate) (src/algorithms.jl:640-643): one cache per algorithm
        wrap = (lambda a: EvalCounter(a)) if counter else (lambda a: a)
        cache = IntegralCache(prob.f, prob.dom, prob.p, alg, kwargs, backend, shard)
        cache.cacheval["est"] = init(prob, wrap(inner.est_alg), backend=backend, shard=shard, **inner.kws)
        cache.cacheval["abs"] = init(prob, wrap(inner.abs_alg), backend=backend, shard=shard)
        return cache
    cache = IntegralCache(prob.f, prob.dom, prob.p, alg, kwargs, backend, shard)
    _init_cacheval(cache)
    return cache


def solve_(cache, plist=None):
    """solve!(cache) (src/interfaces.jl:116-118).  plist: several parameter sets solved against the same
    cached rule in one device pass (the batchsolve fast path); returns one IntegralSolution per entry."""
    from . import dos
    if isinstance(cache, dos.DOSCache):
        return dos.dos_solve_(cache)
    single = plist is None
    ps = [cache.p] if single else list(plist)
    sols = _do_solve(cache, ps)
    return sols[0] if single else sols


def solve(prob, alg, backend=None, shard=None, **kwargs):
    """solve(prob, alg; kwargs...) = solve!(init(prob, alg; kwargs...)) (src/interfaces.jl:106-109)"""
    cache = init(prob, alg, backend=backend, shard=shard, **kwargs)
    try:
        return solve_(cache)
    finally:
        if hasattr(cache, "close"):
            cache.close()


# ---------------------------------------------------------------------------------------------------
def _unwrap(alg):
    counter = isinstance(alg, EvalCounter)
    return (alg.alg if counter else alg), counter


def _standard(dom, alg):
    """bz_to_standard (src/brillouin.jl:375-377, 392-394, 418-420): (bz, unitless domain, standard algorithm, j, nsyms)"""
    if isinstance(dom, SymmetricBZ):
        j = abs(np.linalg.det(dom.B))
        if isinstance(alg, IAI):
            return dom.lims, NestedQuad(*alg.algs), j, dom.nsyms, dom.ndim
        if isinstance(alg, PTR):
            return Basis(np.eye(dom.ndim)), MonkhorstPack(npt=alg.npt, syms=dom.syms, nthreads=alg.nthreads), j, dom.nsyms, dom.ndim
        if isinstance(alg, AutoPTR):
            return (Basis(np.eye(dom.ndim)),
                    AutoSymPTRJL(norm=alg.norm, a=alg.a, nmin=alg.nmin, nmax=alg.nmax, n0=alg.n0, dn=alg.dn, keepmost=alg.keepmost,
                                 syms=dom.syms, nthreads=alg.nthreads), j, dom.nsyms, dom.ndim)
        raise TypeError("unsupported BZ algorithm (TAI is out of scope, SURVEY.md §2)")
    if isinstance(dom, Basis):
        return dom, alg, None, None, dom.ndim
    if _is_iterated_limits(dom):
        return dom, alg, None, None, dom.ndim
    raise TypeError("unsupported domain")


def _is_iterated_limits(dom):
    """IteratedIntegration.AbstractIteratedLimits by protocol: ndim, segments() -> breakpoints, fix(x) -> inner limits"""
    return all(hasattr(dom, a) for a in ("ndim", "segments", "fix")) and not isinstance(dom, SymmetricBZ)


def _init_cacheval(cache):
    alg, _ = _unwrap(cache.alg)
    dom, salg, j, ns, ndim = _standard(cache.dom, alg)
    f = cache.f
    if ndim != f.s.ndim:
        raise ValueError("variables in Fourier series don't match domain")
    cv = cache.cacheval
    cv["std"] = (dom, salg, j, ns, ndim)
    if isinstance(salg, MonkhorstPack):
        # init_fourier_rule (src/fourier.jl:330-342): the rule (node set + device handles) is built at init
        cv["rule"] = cache.backend.make_rule(f.s, ndim, salg.npt, salg.syms, cache.shard.rank, cache.shard.nranks, allreduce=_shard_allreduce(cache.shard))
    elif isinstance(salg, AutoSymPTRJL):
        # AutoSymPTR.alloc_cache builds the first rule at init (src/fourier.jl:348-360)
        n0, dn = monkhorst_pack_schedule(salg.a, salg.nmin, salg.nmax, salg.n0, salg.dn)
        cv["schedule"] = (n0, dn)
        cv["rules"] = [cache.backend.make_rule(f.s, ndim, n0, salg.syms, cache.shard.rank, cache.shard.nranks, allreduce=_shard_allreduce(cache.shard))]
    elif isinstance(salg, NestedQuad):
        if not _is_iterated_limits(dom):
            raise TypeError("NestedQuad needs iterated limits")
        if f.native and f.f.is_eig:
            raise TypeError("IAI on the device supports the resolvent / affine integrands and host (batch) integrands")
        # arena of contracted series: live level-2 slots <= 2 outstanding outer panels x K nodes (x initial segments), live level-1
        # slots <= that x 2 panels x K nodes; GK(7,15) fits the default 64 / 2048
        K = max(2 * a.order + 1 for a in salg.algs)
        # (small series: 256 level-2 slots leave room for the look-ahead bisection of the native engine, abz_iai_engine.hpp)
        c2 = max(256 if f.s.norb <= 3 else 64, 4 * K)
        cv["nest_caps"] = (c2, max(2048, max(64, 4 * K) * 2 * K))
        cv["nest"] = cache.backend.make_nest(f.s, ndim, *cv["nest_caps"])
    else:
        raise TypeError(f"unsupported algorithm {type(salg).__name__} for FourierIntegrand")


def _do_solve_absolute_estimate(cache, ps, alg, counter):
    """do_solve(f, dom, p, ::AbsoluteEstimate, cacheval) (src/algorithms.jl:645-653)"""
    est, ab_ = cache.cacheval["est"], cache.cacheval["abs"]
    kws = dict(cache.kwargs)
    abstol, reltol = kws.get("abstol"), kws.get("reltol")
    norm = _norm if alg.norm is abs else alg.norm
    sols = []
    for p in ps:
        s_est = _do_solve(est, [p])[0]
        val = norm(s_est.u)
        rtol = math.sqrt(np.finfo(np.float64).eps) if reltol is None else reltol
        atol = max(0.0 if abstol is None else abstol, rtol * val)
        ab_.kwargs = {"abstol": atol, "reltol": 0.0, **({"maxiters": kws["maxiters"]} if "maxiters" in kws else {})}
        s_abs = _do_solve(ab_, [p])[0]
        ne = (s_est.numevals + s_abs.numevals) if counter else -1      # EvalCounter counts the evaluations of both solves
        sols.append(IntegralSolution(s_abs.u, s_abs.resid, s_abs.retcode, ne))
    return sols


def _do_solve(cache, ps):
    alg, counter = _unwrap(cache.alg)
    if isinstance(alg, AbsoluteEstimate):
        return _do_solve_absolute_estimate(cache, ps, alg, counter)
    dom, salg, j, ns, ndim = cache.cacheval["std"]
    kws = dict(cache.kwargs)
    abstol, reltol = kws.get("abstol"), kws.get("reltol")
    maxiters = kws.get("maxiters", 2 ** 62)
    on_bz = j is not None
    bf = _BoundIntegrand(cache.f, ps)
    bf.bz = cache.dom
    shard = cache.shard
    if bf.native and not bf.is_eig and cache.f.f.is_matrix:
        if on_bz and cache.dom.syms is not None and cache.f.f.symmetrize is None:
            # do_solve_autobz (src/brillouin.jl:348-353): unknown SymRep + non-trivial value => repeat on the full BZ
            import warnings
            warnings.warn(WARN_UNKNOWN_SYMMETRY)
            fbz = SymmetricBZ(cache.dom.A, cache.dom.B, CubicLimits([0.0] * ndim, [1.0] * ndim), None)
            sub = init(IntegralProblem(cache.f, fbz, cache.p), cache.alg, backend=cache.backend, shard=cache.shard, **cache.kwargs)
            return _do_solve(sub, ps)
    if isinstance(salg, MonkhorstPack):
        # do_solve_autobz (src/brillouin.jl:337-355): sol = rule(f) (scale vol/(npt^d nsyms)); val = j*nsyms*sol
        rule = cache.cacheval["rule"]
        vals = _rule_apply(rule, bf, shard, ndim)
        sc = j if on_bz else abs(np.linalg.det(dom.B))
        ne = len(rule) if counter else -1
        return [IntegralSolution(sc * v, None, True, ne) for v in vals]
    if isinstance(salg, AutoSymPTRJL):
        sc = j if on_bz else abs(np.linalg.det(dom.B))
        atol = None if abstol is None else abstol / sc     # src/brillouin.jl:433 (no nsyms: rule output is symmetrised)
        sols = []
        for p in ps:
            b1 = _BoundIntegrand(cache.f, [p])
            b1.bz = cache.dom
            val, err, ne = _autosymptr(cache, b1, salg, atol, reltol, maxiters, ndim)
            sols.append(IntegralSolution(val * sc, err * sc, True, ne if counter else -1))
        return sols
    if isinstance(salg, NestedQuad):
        sols = []
        sc = j if on_bz else 1.0
        atol = abstol
        if on_bz and abstol is not None:
            atol = abstol / (j * ns)                        # src/brillouin.jl:342
        # per-level algorithms (NestedQuad(algs...): algs[dim] belongs to variable `dim`, src/algorithms.jl:462-463, 503-505); one
        # algorithm is repeated for every level.  A custom norm or an order other than 7 keeps the solve in the Python engine.
        algs = salg.algs
        if len(algs) == 1:
            algs = algs * ndim
        if len(algs) != ndim:
            raise ValueError(f"NestedQuad got {len(algs)} algorithms for a {ndim}-dimensional domain")
        if any(a.norm is not abs for a in algs):
            raise NotImplementedError("AuxQuadGKJL(norm=...) other than abs / the Frobenius norm is not supported")
        orders = tuple(a.order for a in algs)
        default_gk = all(o == 7 for o in orders)
        caps = cache.cacheval.get("nest_caps", (64, 2048))
        for p in ps:
            b1 = _BoundIntegrand(cache.f, [p])
            if not b1.native:
                # generic / batch host integrand: H at the panel nodes comes from the device (abz_nest_eval_h)
                fi, pp = cache.f, b1.plist[0]
                real = [True]

                def user(H, k, fi=fi, pp=pp, real=real):
                    v = np.asarray(fi.host_values(H, k, pp))
                    real[0] = real[0] and not np.iscomplexobj(v)
                    return v

                eng = NestedGK(cache.cacheval["nest"], ndim, dom, None, None, None, None, np.complex128, atol, reltol, maxiters, user=user,
                               rank=shard.rank, nranks=shard.nranks, allreduce=shard.allreduce if shard.nranks > 1 else None, orders=orders, cap2=caps[0], cap1=caps[1])
                Iv, Ev, ne = eng.run()
                cache.cacheval["iai_rounds"] = eng.rounds
                mult = sc * (ns if on_bz else 1)
                u = Iv * mult
                sols.append(IntegralSolution(float(u.real) if real[0] else complex(u), float(Ev) * mult, True, ne if counter else -1))
                continue
            bound = b1.bound[0]
            ff = cache.f.f
            if getattr(ff, "is_matrix", False):
                # matrix-valued gloc_integrand under IAI (docs/src/examples.md:90-106; the reference's nest is generic in the value
                # type, src/fourier.jl:452-456): (z - H - Sigma)^-1 at the panel nodes comes from the device (abz_nest_eval_matrix), the
                # nested GK recursion with norm = Frobenius runs in the Python engine; SymRep maps the IBZ value to the full BZ
                z, sigma = bound
                n = cache.f.s.norb
                eng = NestedGK(cache.cacheval["nest"], ndim, dom, None, z, sigma, None, np.complex128, atol, reltol, maxiters,
                               rank=shard.rank, nranks=shard.nranks, allreduce=shard.allreduce if shard.nranks > 1 else None,
                               vshape=(n, n), matrix=True, orders=orders, cap2=caps[0], cap1=caps[1])
                Iv, Ev, ne = eng.run()
                cache.cacheval["iai_rounds"] = eng.rounds
                if on_bz and cache.dom.syms is not None and ff.symmetrize is not None:
                    Iv = ff.symmetrize(cache.dom, Iv)
                sols.append(IntegralSolution(sc * Iv, float(Ev) * sc * (ns if on_bz else 1), True, ne if counter else -1))
                continue
            if b1.fkind == _lib.F_TRACE_H:
                z, sigma = 0j, None
                test = ff.post(np.zeros(1, dtype=np.complex128), bound)
            else:
                z, sigma = bound
                test = ff.post(np.zeros(1, dtype=np.complex128), bound)
            dtype = np.complex128 if np.iscomplexobj(test) else np.float64
            nest = cache.cacheval["nest"]
            vkind = _native_vkind(ff)
            if default_gk and vkind is not None and hasattr(nest, "iai_solve") and getattr(cache.backend, "iai_engine", "python") == "native":
                # abz_iai_solve: same control flow, run by the library's C++ host engine (one ccall per solve)
                atol_ = 0.0 if atol is None else atol
                rtol_ = reltol if reltol is not None else (np.sqrt(np.finfo(float).eps) if atol_ == 0 else 0.0)
                general = None
                if isinstance(dom, TetrahedralLimits) and dom.s == 1.0:
                    lkind, la, lb = 1, dom.a, None
                elif isinstance(dom, CubicLimits):
                    lkind, la, lb = 0, dom.a, dom.b
                else:
                    # any other iterated limits (polyhedral IBZ, several segments per level, ...): the library asks this object for
                    # the breakpoints of every 1-D integral it starts (abz_iai_solve_general)
                    lkind, la, lb, general = 2, None, None, dom
                lin = bound if vkind == 2 else None
                Iv, Ev, ne, rounds, launches = nest.iai_solve(lkind, la, lb, b1.fkind, vkind, z, sigma, lin, atol_, rtol_, maxiters,
                                                              device_leaves=getattr(cache.backend, "iai_device_leaves", True),
                                                              device_middles=getattr(cache.backend, "iai_device_middles", True),
                                                              speculate=getattr(cache.backend, "iai_speculate", True),
                                                              rank=shard.rank, nranks=shard.nranks,
                                                              allreduce=shard.allreduce if shard.nranks > 1 else None, limits=general)
                cache.cacheval["iai_rounds"] = rounds
                Iv = Iv if dtype == np.complex128 else Iv.real
            else:
                eng = NestedGK(nest, ndim, dom, b1.fkind, z, sigma, lambda y, ff=ff, bound=bound: ff.post(y, bound),
                               dtype, atol, reltol, maxiters,
                               rank=shard.rank, nranks=shard.nranks, allreduce=shard.allreduce if shard.nranks > 1 else None, orders=orders, cap2=caps[0], cap1=caps[1])
                Iv, Ev, ne = eng.run()
                cache.cacheval["iai_rounds"] = eng.rounds
            mult = sc * (ns if on_bz else 1)                # val = j * symmetrize(f, bz, sol.u) (TrivialRep: x nsyms)
            u = Iv * mult
            u = complex(u) if dtype == np.complex128 else float(u)
            sols.append(IntegralSolution(u, float(Ev) * mult, True, ne if counter else -1))
        return sols
    raise TypeError("unsupported algorithm")


def _native_vkind(ff):
    """abz_iai_solve's value map for the integrand (0 identity, 1 -Im/pi, 2 a*y+b), or None when only its Python
    `post` knows it (subclasses that override post fall back to the Python engine)."""
    from .fourier import AffineTraceIntegrand, DOSIntegrand, TrGlocIntegrand
    for cls, vk in ((DOSIntegrand, 1), (AffineTraceIntegrand, 2), (TrGlocIntegrand, 0)):
        if type(ff) is cls:
            return vk
    return None


def _autosymptr(cache, bf, salg, atol, reltol, maxevals, ndim):
    """AutoSymPTR.autosymptr (call site src/algorithms.jl:429): I1 = rule_1(f), I2 = rule_2(f), err = norm(I1 - I2);
    refine with nextrule (npt + dn, src/fourier.jl:315-321) until err <= max(reltol*norm(I2), abstol) or
    numevals >= maxevals.  Rules already in the cache (from another parameter) are reused; at most `keepmost`
    are kept afterwards.  [restated; SURVEY.md App. A.2]"""
    atol_ = 0.0 if atol is None else atol
    rtol_ = reltol if reltol is not None else (np.sqrt(np.finfo(float).eps) if atol_ == 0 else 0.0)
    rules = cache.cacheval["rules"]
    n0, dn = cache.cacheval["schedule"]
    f, backend, shard = cache.f, cache.backend, cache.shard

    timing = cache.cacheval.setdefault("timing", [])

    def rule_at(i):
        while len(rules) <= i:
            prev = rules[-1]
            t0 = time.perf_counter()
            rules.append(backend.make_rule(f.s, ndim, prev.npt + dn, salg.syms, shard.rank, shard.nranks, allreduce=_shard_allreduce(shard)))
            timing.append(("make_rule", prev.npt + dn, time.perf_counter() - t0))
        return rules[i]

    def apply(r):
        t0 = time.perf_counter()
        v = _rule_apply(r, bf, shard, ndim)[0]
        timing.append(("apply", r.npt, time.perf_counter() - t0))
        return v

    numevals = 0
    r1 = rule_at(0)
    int1 = apply(r1)
    numevals += len(r1)
    if numevals >= maxevals:
        return int1, float("nan"), numevals
    r2 = rule_at(1)
    int2 = apply(r2)
    numevals += len(r2)
    norm = _alg_norm(salg)
    err = norm(int1 - int2)
    i = 1
    while not (err <= max(rtol_ * norm(int2), atol_)) and numevals < maxevals and np.isfinite(err):
        i += 1
        r = rule_at(i)
        int1, int2 = int2, apply(r)
        numevals += len(r)
        err = norm(int1 - int2)
    # keep the `keepmost` most refined rules for the next parameter (src/algorithms.jl:429 keepmost)
    keep = max(1, salg.keepmost)
    used = rules[: i + 1]
    drop = (used[:-keep] if len(used) > keep else []) + rules[i + 1:]      # also rules an earlier parameter refined beyond this one's last grid
    for r in drop:
        r.close()
    cache.cacheval["rules"] = used[-keep:] if len(used) > keep else used
    cache.cacheval["last_npt"] = used[-1].npt
    return int2, err, numevals


# ---------------------------------------------------------------------------------------------------
class IntegralSolver:
    """IntegralSolver(f, dom, alg; abstol, reltol, maxiters) / IntegralSolver(prob, alg; ...)
    (src/interfaces.jl:142-187; FourierIntegrand functor src/fourier.jl:89-93): solver(args...; kws...) -> u."""

    def __init__(self, *a, backend=None, shard=None, **kwargs):
        if isinstance(a[0], IntegralProblem):
            prob, alg = a[0], a[1]
        else:
            f, dom, alg = a[0], a[1], a[2]
            prob = IntegralProblem(f, dom, None)
        checkkwargs(kwargs)
        self.prob, self.alg, self.kwargs = prob, alg, kwargs
        self.cache = init(prob, alg, backend=backend, shard=shard, **kwargs)

    def solve_p(self, p):
        """solve_p(s, p) (src/interfaces.jl:174-182): remake_cache with merged parameters, then solve!"""
        base = _params(self.prob.p)
        q = _params(p)
        merged = (base[0] + q[0], {**base[1], **q[1]})
        return solve_(self.cache, [merged])[0]

    def __call__(self, *args, **kws):
        return self.solve_p((args, kws)).u


def batchsolve(solver, ps, callback=None):
    """batchsolve(solver, ps) (src/interfaces.jl:199-243): solve for every parameter in `ps` against ONE cached
    rule.  The reference threads over parameters with the grid shared; here the parameters are an inner batch
    dimension of the device kernels (one H(k) load, n_omega resolvents) whenever the algorithm is a fixed rule."""
    shape = ps.shape if isinstance(ps, np.ndarray) and ps.dtype == object and ps.ndim > 1 else None   # paramproduct
    plist = list(ps.ravel()) if shape is not None else list(ps)
    base = _params(solver.prob.p)
    merged = []
    for p in plist:
        q = _params(p)
        merged.append((base[0] + q[0], {**base[1], **q[1]}))
    t0 = time.time()
    sols = solve_(solver.cache, merged)
    t = time.time() - t0
    if callback is not None:
        for i, (p, s) in enumerate(zip(plist, sols)):
            callback(solver.prob.f, i, len(plist), p, s, t / max(1, len(plist)))
    out = np.array([s.u for s in sols])
    return out.reshape(shape + out.shape[1:]) if shape is not None else out


def batchsolve_log(path, solver, ps, verb=False):
    """batchsolve(h5, f::IntegralSolver, ps) of ext/HDF5Ext.jl:116-158: solve for every parameter and record, per parameter,
    the datasets of the reference's HDF5 archive - I (result), E (error estimate, NaN if none), t (seconds), retcode,
    numevals, and the parameters (positional "args/<j>", keyword "kwargs/<name>").  h5py is not part of this image, so the
    archive is written as a NumPy .npz with those dataset names; `numpy.load(path)` gives them back.  Returns the results."""
    plist = list(ps)
    n = len(plist)
    rec = {"E": np.full(n, np.nan), "t": np.zeros(n), "retcode": np.zeros(n, dtype=np.int32), "numevals": np.zeros(n, dtype=np.int64)}
    results = [None] * n

    def cb(_, i, ntot, p, sol, t):
        if verb:
            print(f"{i + 1:5d} / {ntot} done in {t:e} (s)")
        results[i] = sol.u
        rec["E"][i] = np.nan if sol.resid is None else float(np.max(np.abs(sol.resid)))
        rec["t"][i] = t
        rec["retcode"][i] = int(bool(sol.retcode))
        rec["numevals"][i] = sol.numevals

    out = batchsolve(solver, plist, callback=cb)
    rec["I"] = np.array(results)
    for i, p in enumerate(plist):
        args, kws = _params(p)
        for j, a in enumerate(args):
            rec.setdefault(f"args/{j + 1}", np.zeros(n, dtype=np.asarray(a).dtype))[i] = a
        for k, v in kws.items():
            rec.setdefault(f"kwargs/{k}", np.zeros(n, dtype=np.asarray(v).dtype))[i] = v
    np.savez(path, **rec)
    return out
