"""FourierSeries / FourierValue / FourierIntegrand and the device-native integrands.

Host-side mirror of src/fourier.jl:22-58, 111-122 (containers) and of the canonical user integrands
(aps_example/aps_example.jl:30, docs/src/examples.md:13-20,90, test/fourier.jl:41, src/dos_ggr.jl:19).
The arithmetic of the named integrands runs in libautobz_cuda.so; a plain Python callable `f` is
also accepted and is evaluated on the host on H(k) copied back from the device (the reference's
generic path for arbitrary user functions)."""
import numpy as np

from . import _lib


class FourierSeries:
    """FourierSeries(C; period, offset) of FourierSeriesEvaluators.jl:
    f(x) = sum_i C[i] exp(2 pi i sum_d x_d (i_d + offset_d) / period_d), i 1-based as in Julia.

    C: array [M1(,M2(,M3))] for a scalar series, or [n, n, M1(,M2(,M3))] with norb=n for a matrix-valued
    one.  `lo` (lowest R index per dimension) may be given instead of the Julia-style `offset`
    (lo = offset + 1); OffsetArray inputs of the reference correspond to passing `lo` directly."""

    def __init__(self, C, period=1.0, offset=0, lo=None, norb=None):
        C = np.asarray(C)
        if norb is None:
            C = C[None, None]
            norb = 1
        if C.shape[0] != norb or C.shape[1] != norb:
            raise ValueError("matrix-valued coefficients must have shape [n, n, M1, ...]")
        self.ndim = C.ndim - 2
        if not 1 <= self.ndim <= 3:
            raise ValueError("only 1-, 2- and 3-dimensional series are supported")
        self.norb = norb
        self.c = C
        if lo is None:
            off = np.broadcast_to(np.asarray(offset, dtype=int), (self.ndim,))
            lo = tuple(int(o) + 1 for o in off)
        self.lo = tuple(int(x) for x in np.broadcast_to(np.asarray(lo, dtype=int), (self.ndim,)))
        self.period = tuple(float(x) for x in np.broadcast_to(np.asarray(period, dtype=float), (self.ndim,)))
        self._dev = {}

    @property
    def M(self):
        return tuple(self.c.shape[2:])

    def device(self, ctx):
        """The series uploaded to `ctx` (cached)."""
        key = id(ctx)
        d = self._dev.get(key)
        if d is None or d.h is None or d.ctx is not ctx:        # (the cached handle keeps its context alive, so an id is not reused under it)
            d = self._dev[key] = _lib.DeviceSeries(ctx, self.c, self.lo, self.period)
        return d

    def drop_device(self):
        for d in self._dev.values():
            d.close()
        self._dev = {}


def period(s):
    return s.period


class FourierValue:
    """FourierValue(x, s) (src/fourier.jl:111-114): point x and series value s = H(x)."""
    __slots__ = ("x", "s")

    def __init__(self, x, s):
        self.x, self.s = x, s


class BatchIntegrand:
    """BatchIntegrand(f!, y, x; max_batch) (src/batch.jl:10-38): f!(y, x, args...; kws...) fills y[i] with the integrand at
    node i of a whole vector of nodes - the reference's documented hook for threads / GPU / distributed evaluation.
    As the integrand of a FourierIntegrand, x is a FourierValue whose fields are arrays: x.x [nb, ndim] the nodes and
    x.s [nb, n, n] (or [nb] for a scalar series) the series values, evaluated on the device; y is a preallocated
    numpy array of length nb and dtype `dtype`.  max_batch is the soft cap on nb (src/batch.jl:17)."""

    def __init__(self, f, dtype=np.complex128, max_batch=None):
        if max_batch is not None and max_batch <= 0:
            raise ValueError("maximum batch size must be positive")
        self.f, self.dtype, self.max_batch = f, np.dtype(dtype), max_batch


class NestedBatchIntegrand:
    """NestedBatchIntegrand(f::Tuple, y, x; max_batch) (src/batch.jl:50-77): the reference's container of per-thread worker
    copies of one integrand for multi-threaded nested evaluation (FourierIntegrand(p, ws, nest), test/fourier.jl:24-37).
    Here the batch is what goes to the device, so the workers collapse to the first one: it is evaluated on whole batches
    of nodes (H(k) from the device), at most max_batch at a time.  Results equal the un-nested integrand's."""

    def __init__(self, f, dtype=np.complex128, max_batch=None):
        workers = tuple(f) if isinstance(f, (tuple, list)) else (f,)
        if not workers:
            raise ValueError("NestedBatchIntegrand needs at least one worker")
        if max_batch is not None and max_batch <= 0:
            raise ValueError("maximum batch size must be positive")
        while isinstance(workers[0], NestedBatchIntegrand):      # nests of nests: same integrand underneath
            workers = workers[0].f
        self.f, self.dtype, self.max_batch = workers, np.dtype(dtype), max_batch


class _NativeIntegrand:
    """Base of the integrands whose arithmetic runs on the device."""
    fkind = _lib.F_RESOLVENT_TRACE
    is_eig = False
    is_matrix = False

    def post(self, y, bound=None):
        """map the device value(s) to the integrand's value (vectorised)"""
        return y


def _getkw(kws, *names, default=None, required=True):
    for nm in names:
        if nm in kws:
            return kws[nm]
    if required and default is None:
        raise TypeError(f"missing integrand parameter {names[0]}")
    return default


class TrGlocIntegrand(_NativeIntegrand):
    """tr[(omega + i eta - H(k) - Sigma)^-1]: trace of gloc_integrand(h_k; eta, omega) of
    docs/src/examples.md:13-20,90.  Parameters: eta (η), omega (ω) as keywords, optional Sigma (Σ):
    complex scalar or n x n matrix."""

    def bind(self, args, kws):
        if len(args) == 2:
            eta, omega = args
        elif len(args) == 1:
            eta = _getkw(kws, "eta", "η")
            omega = args[0]
        else:
            eta = _getkw(kws, "eta", "η")
            omega = _getkw(kws, "omega", "ω")
        sigma = _getkw(kws, "Sigma", "Σ", "sigma", required=False)
        z = complex(omega, eta)
        if sigma is not None and np.ndim(sigma) == 0:
            z = z - complex(sigma)
            sigma = None
        return z, sigma


class GlocIntegrand(TrGlocIntegrand):
    """gloc_integrand(h_k; eta, omega) = inv(complex(omega, eta) I - h_k.s) (docs/src/examples.md:20,90): the matrix-valued
    local Green's function.  PTR / AutoPTR sums run on the device (abz_rule_resolvent_matrix_sum).  Its SymRep is UnknownRep
    unless `symmetrize` is given: on an IBZ the reference then returns the IBZ integral as is (src/brillouin.jl:107);
    symmetrize(bz, G) -> G_FBZ lets the caller supply the representation, e.g. sum_S S G S^H."""
    is_matrix = True
    vkind = None

    def __init__(self, symmetrize=None):
        self.symmetrize = symmetrize


class DOSIntegrand(TrGlocIntegrand):
    """dos_integrand(h_k, eta, omega) = -imag(tr(inv((omega + i eta) I - h_k.s)))/pi
    (aps_example/aps_example.jl:30).  Parameters positional (eta, omega) as in the example, or keywords."""

    def post(self, y, bound=None):
        return -np.imag(y) / np.pi


class AffineTraceIntegrand(_NativeIntegrand):
    """f(x::FourierValue, a; b) = a * tr(x.s) + b  (test/fourier.jl:41 with a scalar series)."""
    fkind = _lib.F_TRACE_H

    def bind(self, args, kws):
        a = args[0] if len(args) >= 1 else _getkw(kws, "a")
        b = args[1] if len(args) >= 2 else _getkw(kws, "b", default=0.0, required=False)
        return (a, 0.0 if b is None else b)

    def post(self, y, bound=None):
        a, b = bound
        return a * y + b


class EigenIntegrand(_NativeIntegrand):
    """g(eigvals(Hermitian(H(k)))) with the eigen-decomposition on the device (src/dos_ggr.jl:19,34).
    kind: 'sum' | 'fermi_energy' | 'fermi_count' | 'gauss_dos'; parameters (mu, T) or (omega, sigma)."""
    is_eig = True
    _kinds = {"sum": _lib.EIG_SUM, "fermi_energy": _lib.EIG_FERMI_ENERGY, "fermi_count": _lib.EIG_FERMI_COUNT,
              "gauss_dos": _lib.EIG_GAUSS_DOS}

    def __init__(self, kind="sum"):
        if kind not in self._kinds:
            raise ValueError(f"unknown eigenvalue integrand {kind}")
        self.kind = self._kinds[kind]

    def bind(self, args, kws):
        if self.kind == _lib.EIG_SUM:
            return (0.0, 1.0)
        if len(args) >= 2:
            return (float(args[0]), float(args[1]))
        if self.kind == _lib.EIG_GAUSS_DOS:
            return (float(_getkw(kws, "omega", "ω")), float(_getkw(kws, "sigma", "σ")))
        return (float(_getkw(kws, "mu", "μ")), float(_getkw(kws, "T", "kT")))


# ready-made instances named as in the reference's examples
dos_integrand = DOSIntegrand()
gloc_trace_integrand = TrGlocIntegrand()
gloc_integrand = GlocIntegrand()


class FourierIntegrand:
    """FourierIntegrand(f, s, args...; kws...) (src/fourier.jl:37-58): integrand f(FourierValue(x, s(x)), args...; kws...)
    with the series evaluated one dimension at a time by the specialised rules."""

    def __init__(self, f, s, *args, nest=None, **kws):
        if not isinstance(s, FourierSeries):
            raise TypeError("s must be a FourierSeries")
        if nest is not None:
            # FourierIntegrand(f, w, nest::NestedBatchIntegrand) (src/fourier.jl:37-46): the nest's workers evaluate f
            if not isinstance(nest, NestedBatchIntegrand):
                raise TypeError("nest must be a NestedBatchIntegrand")
            worker, mb, dt = nest.f[0], nest.max_batch, nest.dtype

            def batch(y, x, *a, **k):
                for i in range(len(y)):
                    y[i] = worker(FourierValue(x.x[i], x.s[i]), *a, **k)

            f = BatchIntegrand(batch, dtype=dt, max_batch=mb)
        self.f, self.s, self.args, self.kws = f, s, tuple(args), dict(kws)

    @property
    def native(self):
        return isinstance(self.f, _NativeIntegrand)

    def merged(self, p):
        """merge(f.f.p, p) of MixedParameters (src/fourier.jl:95-100): stored args first, then the call's."""
        args, kws = p if p is not None else ((), {})
        return self.args + tuple(args), {**self.kws, **kws}

    def __call__(self, x, p=None):
        """fallback evaluator (src/fourier.jl:120-122) — only for plain Python integrands"""
        args, kws = self.merged(p)
        if self.native:
            raise TypeError("device-native integrands are evaluated by the specialised rules")
        return self.f(x, *args, **kws)

    def host_values(self, H, k, p=None):
        """values of a host (non-native) integrand on a batch of nodes: H [n, n, nb] from the device, k [nb, ndim].
        BatchIntegrand: f!(y, x, p) in chunks of max_batch (src/batch.jl:4-20); plain callable: node by node."""
        args, kws = self.merged(p)
        nb = H.shape[2]
        Hm = np.moveaxis(H, 2, 0)
        if self.s.norb == 1:
            Hm = Hm[:, 0, 0]
        if isinstance(self.f, BatchIntegrand):
            y = np.empty(nb, dtype=self.f.dtype)
            step = nb if self.f.max_batch is None else int(self.f.max_batch)
            for a in range(0, nb, max(step, 1)):
                b = min(nb, a + step)
                self.f.f(y[a:b], FourierValue(k[a:b], Hm[a:b]), *args, **kws)
            return y
        return np.array([self.f(FourierValue(k[i], Hm[i]), *args, **kws) for i in range(nb)])
