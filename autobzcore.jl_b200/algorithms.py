"""Algorithm descriptors — mirror of src/brillouin.jl:360-499 (IAI, PTR, AutoPTR, count_bz_to_standard)
and src/algorithms.jl:202-240, 342-432, 450-455, 662-666 (AuxQuadGKJL, MonkhorstPack, AutoSymPTRJL,
NestedQuad, EvalCounter).  Plain data; the control flow lives in interfaces.py / rules.py / iai.py."""
import math


class IntegralAlgorithm:
    pass


class AutoBZAlgorithm(IntegralAlgorithm):
    pass


class AuxQuadGKJL(IntegralAlgorithm):
    """AuxQuadGKJL(; order=7, norm=norm) (src/algorithms.jl:202-208).  Order 7 (GK 7/15) runs in the library's native IAI engine and
    its device-side integrals; any other order >= 2 runs the same state machine in the Python engine (iai.NestedGK) with the device
    evaluating the panel nodes."""

    def __init__(self, order=7, norm=abs):
        if int(order) != order or order < 2:
            raise ValueError("the Gauss-Kronrod order must be an integer >= 2")
        self.order, self.norm = int(order), norm


class NestedQuad(IntegralAlgorithm):
    """NestedQuad(alg) / NestedQuad(algs...) (src/algorithms.jl:450-455)"""

    def __init__(self, *algs):
        if len(algs) == 0:
            algs = (AuxQuadGKJL(),)
        for a in algs:
            if not isinstance(a, AuxQuadGKJL):
                raise TypeError("NestedQuad levels must be AuxQuadGKJL")
        self.algs = algs


class MonkhorstPack(IntegralAlgorithm):
    """MonkhorstPack(; npt=50, syms=nothing, nthreads=1) (src/algorithms.jl:342-347)"""

    def __init__(self, npt=50, syms=None, nthreads=1):
        self.npt, self.syms, self.nthreads = int(npt), syms, nthreads


class AutoSymPTRJL(IntegralAlgorithm):
    """AutoSymPTRJL(; norm, a=1.0, nmin=50, nmax=1000, n0=6.0, dn=log(10), keepmost=2, syms=nothing)
    (src/algorithms.jl:393-406)"""

    def __init__(self, norm=abs, a=1.0, nmin=50, nmax=1000, n0=6.0, dn=math.log(10), keepmost=2, syms=None, nthreads=1,
                 **kw):
        n0 = kw.pop("n₀", n0)
        dn = kw.pop("Δn", dn)
        if kw:
            raise TypeError(f"unknown options {list(kw)}")
        self.norm, self.a, self.nmin, self.nmax, self.n0, self.dn = norm, float(a), int(nmin), int(nmax), float(n0), float(dn)
        self.keepmost, self.syms, self.nthreads = int(keepmost), syms, nthreads


class IAI(AutoBZAlgorithm):
    """IAI(alg=AuxQuadGKJL()) (src/brillouin.jl:368-377): iterated adaptive integration."""

    def __init__(self, *algs):
        self.algs = algs if algs else (AuxQuadGKJL(),)


class PTR(AutoBZAlgorithm):
    """PTR(; npt=50, nthreads=1) (src/brillouin.jl:386-394)"""

    def __init__(self, npt=50, nthreads=1):
        self.npt, self.nthreads = int(npt), nthreads


class AutoPTR(AutoBZAlgorithm):
    """AutoPTR(; norm, a=1.0, nmin=50, nmax=1000, n0=6.0, dn=log(10), keepmost=2) (src/brillouin.jl:405-420)"""

    def __init__(self, norm=abs, a=1.0, nmin=50, nmax=1000, n0=6.0, dn=math.log(10), keepmost=2, nthreads=1, **kw):
        n0 = kw.pop("n₀", n0)
        dn = kw.pop("Δn", dn)
        if kw:
            raise TypeError(f"unknown options {list(kw)}")
        self.norm, self.a, self.nmin, self.nmax, self.n0, self.dn = norm, float(a), int(nmin), int(nmax), float(n0), float(dn)
        self.keepmost, self.nthreads = int(keepmost), nthreads


class AbsoluteEstimate(IntegralAlgorithm):
    """AbsoluteEstimate(est_alg, abs_alg; norm, kws...) (src/algorithms.jl:614-653): a rough estimate I from `est_alg` (solved with
    `kws`: reltol / abstol / maxiters) turns the caller's relative tolerance into an absolute one for `abs_alg`:
    abstol = max(abstol, reltol * norm(I)), reltol = 0."""

    def __init__(self, est_alg, abs_alg, norm=abs, **kws):
        for k in kws:
            if k not in ("abstol", "reltol", "maxiters"):
                raise ValueError(f"keyword {k} unrecognized")
        self.est_alg, self.abs_alg, self.norm, self.kws = est_alg, abs_alg, norm, kws


def PTR_IAI(ptr=None, iai=None, **kws):
    """PTR_IAI(; ptr=PTR(), iai=IAI()) (src/brillouin.jl:466-476): IAI with abstol = reltol * norm(PTR estimate)"""
    return AbsoluteEstimate(ptr if ptr is not None else PTR(), iai if iai is not None else IAI(), **kws)


def AutoPTR_IAI(reltol=1.0, ptr=None, iai=None, **kws):
    """AutoPTR_IAI(; reltol=1.0, ptr=AutoPTR(), iai=IAI()) (src/brillouin.jl:479-490): the estimate comes from AutoPTR at `reltol`"""
    return AbsoluteEstimate(ptr if ptr is not None else AutoPTR(), iai if iai is not None else IAI(), reltol=reltol, **kws)


class TAI(AutoBZAlgorithm):
    """TAI(; norm, initdiv=1) (src/brillouin.jl:446-463): tree-adaptive integration through HCubature.  Named for completeness of the
    reference's algorithm list; it is not on the Fourier hot path and `init` refuses it."""

    def __init__(self, norm=abs, initdiv=1):
        self.norm, self.initdiv = norm, int(initdiv)


class EvalCounter(IntegralAlgorithm):
    """EvalCounter(alg) (src/algorithms.jl:662-666): sol.numevals = number of integrand evaluations."""

    def __init__(self, alg):
        self.alg = alg


def monkhorst_pack_schedule(a, nmin, nmax, n0, dn):
    """AutoSymPTR.MonkhorstPackRule(syms, a, nmin, nmax, n0, dn) (call sites src/fourier.jl:301-304,
    src/algorithms.jl:407-409): integer first grid and increment.  [restated from the published
    algorithm; exact rounding unpinned, SURVEY.md App. A.2]"""
    first = min(max(int(nmin), int(math.ceil(n0 / a))), int(nmax))
    step = max(1, int(math.ceil(dn / a)))
    return first, step
