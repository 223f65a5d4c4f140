"""Density-of-states problems: host-side mirror of src/dos_interfaces.jl (DOSProblem / DOSSolution / DOSCache / init /
solve!), src/dos_algorithms.jl (GGR) and src/dos_ggr.jl (init_cacheval = data pass, dos_solve = sum_ggr), with the data
pass - H(k) and dH/dk on the PTR grid, eigen-decomposition, band velocities - and the Gilat-Raubenheimer sum on the device
(abz_rule_ggr_data / abz_rule_ggr_sum)."""
import numpy as np

from .backend import DeviceBackend
from .bz import SymmetricBZ
from .fourier import FourierSeries
from .interfaces import Shard, checkkwargs


class DOSAlgorithm:
    pass


class GGR(DOSAlgorithm):
    """GGR(; npt=50) (src/dos_algorithms.jl:24-27): generalized Gilat-Raubenheimer method, npt k-points per dimension."""

    def __init__(self, npt=50):
        self.npt = int(npt)


class DOSProblem:
    """DOSProblem(H, domain, p) (src/dos_interfaces.jl:33-38): H a FourierSeries, domain an energy (or array of energies:
    one data pass serves them all), p the SymmetricBZ from load_bz."""

    def __init__(self, H, domain, p=None):
        self.H, self.domain, self.p = H, domain, p


class DOSSolution:
    """DOSSolution(u, err, retcode, numevals) (src/dos_interfaces.jl:40-45)"""

    def __init__(self, u, err=None, retcode=True, numevals=-1):
        self.u, self.err, self.retcode, self.numevals = u, err, retcode, numevals

    def __repr__(self):
        return f"DOSSolution(u={self.u!r}, err={self.err!r}, retcode={self.retcode}, numevals={self.numevals})"


class DOSCache:
    """DOSCache (src/dos_interfaces.jl:49-64): assigning .H marks the cache fresh so that the data pass is redone;
    .domain may be changed freely between solves (the data pass is reused, test/dos.jl:104-108)."""

    def __init__(self, H, domain, p, alg, kwargs, backend, shard):
        object.__setattr__(self, "isfresh", False)
        self.H, self.domain, self.p, self.alg, self.kwargs = H, domain, p, alg, kwargs
        self.backend, self.shard = backend, shard
        self.cacheval = None
        object.__setattr__(self, "isfresh", False)

    def __setattr__(self, name, item):
        if name == "H":
            object.__setattr__(self, "isfresh", True)
        object.__setattr__(self, name, item)


def _init_cacheval(cache):
    """init_cacheval(h, domain, p, alg::GGR) (src/dos_ggr.jl:1-12): rule on the (symmetry-reduced) PTR grid + data pass"""
    h, bz, alg = cache.H, cache.p, cache.alg
    if not isinstance(h, FourierSeries):
        raise TypeError("GGR currently supports Fourier series Hamiltonians")
    if not isinstance(bz, SymmetricBZ):
        raise TypeError("GGR supports BZ parameters from load_bz")
    if bz.ndim != h.ndim:
        raise ValueError("variables in Fourier series don't match domain")
    rule = cache.backend.make_rule(h, h.ndim, alg.npt, bz.syms, cache.shard.rank, cache.shard.nranks,
                                   allreduce=cache.shard.allreduce if cache.shard.nranks > 1 else None)
    if len(rule) == 0:
        raise ValueError("GGR - no data in rule")
    rule.ggr_data(h.ndim, copy=False)
    return rule


def dos_init(prob, alg, backend=None, shard=None, **kwargs):
    """init(::DOSProblem, ::DOSAlgorithm; kwargs...) (src/dos_interfaces.jl:84-88)"""
    checkkwargs(kwargs)
    if not isinstance(alg, GGR):
        raise TypeError("only the GGR algorithm is implemented (SURVEY.md 8f)")
    cache = DOSCache(prob.H, prob.domain, prob.p, alg, kwargs, backend if backend is not None else DeviceBackend(),
                     shard if shard is not None else Shard())
    cache.cacheval = _init_cacheval(cache)
    object.__setattr__(cache, "isfresh", False)
    return cache


def dos_solve_(cache):
    """solve!(::DOSCache) (src/dos_interfaces.jl:107-115) -> dos_solve (src/dos_ggr.jl:46-56)"""
    if cache.isfresh:
        cache.cacheval = _init_cacheval(cache)
        object.__setattr__(cache, "isfresh", False)
    E = cache.domain
    if not isinstance(cache.p, SymmetricBZ):
        raise TypeError("GGR supports BZ parameters from load_bz")
    scalar = np.ndim(E) == 0
    if not scalar and np.ndim(E) != 1:
        raise TypeError("GGR supports domains of individual eigenvalues")
    part = np.asarray(cache.cacheval.ggr_sum(np.atleast_1d(np.asarray(E, dtype=np.float64))), dtype=np.float64)
    A = cache.shard.allreduce(part)
    return DOSSolution(float(A[0]) if scalar else np.array(A), None, True, -1)


def dos_solve(prob, alg, backend=None, shard=None, **kwargs):
    """solve(::DOSProblem, ::DOSAlgorithm; kwargs...) (src/dos_interfaces.jl:96-99)"""
    return dos_solve_(dos_init(prob, alg, backend=backend, shard=shard, **kwargs))
