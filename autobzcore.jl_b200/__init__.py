"""autobz_b200 — B200-native (sm_100a) hot path of AutoBZCore.jl v0.3.8 behind the reference's
IntegralProblem / init / solve API: Wannier/Fourier interpolation of H(k), per-k resolvent trace and
Hermitian eigenvalues, weighted k-sums, for PTR / AutoPTR / IAI on SymmetricBZ domains.

All arithmetic on the path runs in libautobz_cuda.so (hand-written CUDA); there is no CPU fallback."""
from . import _lib, synthetic  # noqa: F401
from ._lib import AutoBZCudaError, SingularIntegrandError  # noqa: F401
from .algorithms import (IAI, PTR, PTR_IAI, TAI, AbsoluteEstimate, AutoPTR, AutoPTR_IAI, AutoSymPTRJL, AuxQuadGKJL, EvalCounter, MonkhorstPack,  # noqa: F401
                         NestedQuad)
from .backend import DeviceBackend, default_context  # noqa: F401
from .bz import (FBZ, IBZ, AbstractSymRep, CubicLimits, CubicSymIBZ, FunctionRep, InversionSymIBZ, PolygonLimits, PolyhedronLimits,  # noqa: F401
                 SegmentedLimits, SymmetricBZ, SymRep, TetrahedralLimits, TrivialRep, UnknownRep, cube_automorphisms, load_bz, nsyms, symmetrize)
from .dos import GGR, DOSCache, DOSProblem, DOSSolution  # noqa: F401
from .fourier import (AffineTraceIntegrand, BatchIntegrand, DOSIntegrand, EigenIntegrand, FourierIntegrand, FourierSeries,  # noqa: F401
                      FourierValue, GlocIntegrand, NestedBatchIntegrand, TrGlocIntegrand, dos_integrand, gloc_integrand, gloc_trace_integrand)
from .interfaces import (Basis, IntegralProblem, IntegralSolution, IntegralSolver, MixedParameters, Shard, batchsolve, batchsolve_log, init,
                         merge, paramproduct, paramzip,  # noqa: F401
                         solve, solve_, torch_allreduce)
from .wannier import read_w90_hrdat, read_wout_lattice  # noqa: F401

# v0.4+ names of the reference API (BASELINE.json north_star) as aliases
FourierIntegralFunction = FourierIntegrand
CommonSolveFourierIntegralFunction = FourierIntegrand
BatchIntegralFunction = BatchIntegrand
