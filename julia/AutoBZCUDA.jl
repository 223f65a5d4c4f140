# AutoBZCUDA.jl — reference-side binding of libautobz_cuda.so for AutoBZCore.jl v0.3.8.
#
# Written blind (no Julia in the build image; syntax checked by eye only — see INTEGRATION.md).
# It shows the exact ccall signatures a maintainer binds and how the device rules plug into the
# reference's dispatch seams:
#   S1  init_fourier_rule -> FourierPTR / FourierMonkhorstPack      (src/fourier.jl:330-337, 166-174, 265-277)
#   S2  (rule)(f, B, buffer) = quadsum(...)                          (src/fourier.jl:204-207, 289-292)
#   S3  BatchIntegrand f!(y, x, p)                                   (src/batch.jl:4-20)
#   S4  IAI innermost / level closures                               (src/fourier.jl:432-486)
# The adaptive control flow (autosymptr, auxquadgk, nested_quad) stays in Julia untouched.
module AutoBZCUDA

using AutoBZCore
using AutoBZCore: FourierIntegrand, FourierValue, IntegralSolution, MonkhorstPack, AutoSymPTRJL
using AutoSymPTR, FourierSeriesEvaluators, StaticArrays, LinearAlgebra
import AutoBZCore: init_fourier_rule, rule_type
import AutoSymPTR: nextrule

const LIB = get(ENV, "AUTOBZ_CUDA_LIB", "libautobz_cuda")

struct AbzError <: Exception
    code::Int32
    msg::String
end
Base.showerror(io::IO, e::AbzError) = print(io, "libautobz_cuda error ", e.code, ": ", e.msg)

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer=0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:abz_ctx_create, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, r)
        rc == 0 || throw(AbzError(rc, unsafe_string(ccall((:abz_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
        ctx = new(r[])
        finalizer(c -> ccall((:abz_ctx_destroy, LIB), Int32, (Ptr{Cvoid},), c.h), ctx)
        return ctx
    end
end

function check(ctx::Context, rc::Int32)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:abz_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h))
    rc == -1 && throw(ArgumentError(msg))                       # ABZ_E_INVALID  (src/fourier.jl:167 style)
    rc == -4 && throw(DomainError(NaN, msg))                    # ABZ_E_SINGULAR (QuadGK's DomainError)
    throw(AbzError(rc, msg))
end

# one context per Julia thread: the reference protects per-thread state by copying (src/interfaces.jl:213)
const CTXS = Dict{Int,Context}()
context() = get!(() -> Context(parse(Int, get(ENV, "LOCAL_RANK", "0"))), CTXS, Threads.threadid())

# ---- series: Array{SMatrix{n,n,T},3} (or OffsetArray) has the layout ComplexF64[n,n,M1,M2,M3] --------
mutable struct DeviceSeries
    ctx::Context
    h::UInt64
    norb::Int
end

function DeviceSeries(ctx::Context, s::FourierSeries{S,3}) where {S}
    C = parent(s.c)                                            # strip OffsetArray axes
    T = eltype(eltype(C))
    n = size(eltype(C), 1)
    M = Int32[size(C)...]
    lo = Int32[(first(axes(s.c, d)) + s.o[d]) for d in 1:3]    # index + offset = R
    per = Float64[s.t...]
    h = Ref{UInt64}(0)
    GC.@preserve C begin
        rc = ccall((:abz_series_create, LIB), Int32,
                   (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ref{UInt64}),
                   ctx.h, Ptr{Float64}(pointer(C)), T <: Complex ? 1 : 0, n, M, lo, per, h)
    end
    check(ctx, rc)
    ds = DeviceSeries(ctx, h[], n)
    finalizer(x -> ccall((:abz_series_destroy, LIB), Int32, (Ptr{Cvoid}, UInt64), x.ctx.h, x.h), ds)
    return ds
end

# ---- S1: device rules with the interface of FourierPTR / FourierMonkhorstPack ------------------------
mutable struct DeviceRule{d}
    ctx::Context
    h::UInt64
    series::DeviceSeries
    npt::Int
    nsyms::Int
    nnodes::Int
end
Base.length(r::DeviceRule) = r.nnodes
rule_type(::DeviceRule{d}) where {d} = FourierValue{SVector{d,Float64},Nothing}

function DeviceRule(ds::DeviceSeries, npt::Integer, syms)
    ctx = ds.ctx
    h = Ref{UInt64}(0)
    if syms === nothing
        check(ctx, ccall((:abz_rule_create_full, LIB), Int32, (Ptr{Cvoid}, UInt64, Int32, Int32, Int32, Ref{UInt64}),
                         ctx.h, ds.h, npt, 0, npt, h))
        nsym = 1
    else
        S = Int32[round(Int32, s[i, j]) for j in 1:3, i in 1:3, s in syms]       # row-major 3x3 per symmetry
        nirr = Ref{Int64}(0)
        # symptr_rule + CSR compaction on the device; the dense Int32[npt,npt,npt] weights never visit the host
        # (abz_symptr_rule + abz_rule_create_sym remain for callers that already hold AutoSymPTR's wsym array)
        check(ctx, ccall((:abz_rule_create_symptr, LIB), Int32,
                         (Ptr{Cvoid}, UInt64, Int32, Int32, Ptr{Int32}, Int32, Int32, Ref{UInt64}, Ref{Int64}),
                         ctx.h, ds.h, npt, length(syms), S, 0, 1, h, nirr))
        nsym = length(syms)
    end
    nn = Ref{Int64}(0); no = Ref{Int32}(0); np = Ref{Int32}(0)
    check(ctx, ccall((:abz_rule_info, LIB), Int32, (Ptr{Cvoid}, UInt64, Ref{Int64}, Ref{Int32}, Ref{Int32}), ctx.h, h[], nn, no, np))
    r = DeviceRule{3}(ctx, h[], ds, npt, nsym, nn[])
    finalizer(x -> ccall((:abz_rule_destroy, LIB), Int32, (Ptr{Cvoid}, UInt64), x.ctx.h, x.h), r)
    return r
end

# A FourierWorkspace that remembers its device copy
struct CudaWorkspace{W}
    w::W
    ds::DeviceSeries
end
cuda_workspace(s::FourierSeries) = CudaWorkspace(AutoBZCore.workspace_allocate_vec(s, period(s)), DeviceSeries(context(), s))

# replaces init_fourier_rule (src/fourier.jl:330-337)
function init_fourier_rule(w::CudaWorkspace, dom::AutoSymPTR.Basis, alg::MonkhorstPack)
    return DeviceRule(w.ds, alg.npt, alg.syms)
end
# AutoPTR: rule definition + nextrule (src/fourier.jl:296-321)
struct DeviceRuleDef{S}
    ds::DeviceSeries
    m::AutoSymPTR.MonkhorstPackRule{S}
end
(r::DeviceRuleDef)(::Type{T}, ::Val{d}) where {T,d} = DeviceRule(r.ds, r.m.n₀, r.m.syms)
nextrule(p::DeviceRule, r::DeviceRuleDef) = DeviceRule(r.ds, p.npt + r.m.Δn, r.m.syms)

# ---- S2: rule application for the named integrands ---------------------------------------------------
"tr[(ω + iη - H(k) - Σ)^-1] integrand marker: `ResolventTrace()(h_k::FourierValue; η, ω)` also works on the CPU path"
struct ResolventTrace end
(::ResolventTrace)(h_k::FourierValue; η, ω, Σ=nothing) =
    tr(inv(complex(ω, η) * I - h_k.s - (Σ === nothing ? zero(h_k.s) : Σ)))

"sum_i w_i tr[(z_w - H(k_i) - Σ_w)^-1] for all frequencies in one device pass; scale as AutoSymPTR.quadsum's"
function resolvent_sum(rule::DeviceRule, z::Vector{ComplexF64}, Σ::Union{Nothing,Array{ComplexF64,3}}, scale::Float64)
    out = Vector{ComplexF64}(undef, length(z))
    GC.@preserve z Σ out begin
        rc = ccall((:abz_rule_resolvent_sum, LIB), Int32,
                   (Ptr{Cvoid}, UInt64, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}),
                   rule.ctx.h, rule.h, 0, length(z), Ptr{Float64}(pointer(z)),
                   Σ === nothing ? Ptr{Float64}(C_NULL) : Ptr{Float64}(pointer(Σ)), scale, Ptr{Float64}(pointer(out)))
    end
    check(rule.ctx, rc)
    return out
end

# (rule)(f, B, buffer): src/fourier.jl:204-207, 289-292.  vol/(npt^d nsyms) exactly as the reference.
function (rule::DeviceRule{d})(f::AutoBZCore.ParameterIntegrand{ResolventTrace}, B::AutoSymPTR.Basis, buffer=nothing) where {d}
    p = f.p
    z = ComplexF64[complex(p.ω, p.η)]
    scale = abs(det(B.B)) / (rule.npt^d * rule.nsyms)
    return resolvent_sum(rule, z, nothing, scale)[1]
end

"sum_i w_i (z_w - H(k_i) - Σ_w)^-1: the matrix-valued gloc_integrand of docs/src/examples.md:20,90 (symmetrise with your SymRep)"
function resolvent_matrix_sum(rule::DeviceRule, z::Vector{ComplexF64}, Σ::Union{Nothing,Array{ComplexF64,3}}, scale::Float64)
    n = rule.series.norb
    out = Array{ComplexF64}(undef, n, n, length(z))
    GC.@preserve z Σ out begin
        rc = ccall((:abz_rule_resolvent_matrix_sum, LIB), Int32,
                   (Ptr{Cvoid}, UInt64, Int32, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}),
                   rule.ctx.h, rule.h, length(z), Ptr{Float64}(pointer(z)),
                   Σ === nothing ? Ptr{Float64}(C_NULL) : Ptr{Float64}(pointer(Σ)), scale, Ptr{Float64}(pointer(out)))
    end
    check(rule.ctx, rc)
    return out
end

# GGR (src/dos_ggr.jl): get_ggr_data on the device (energies + band velocities at every node), then sum_ggr for a list of energies.
# Drop-in for init_cacheval(h, domain, p, alg::GGR) / dos_solve: weights stay on the device with the rule.
function ggr_data!(rule::DeviceRule{d}; copy=false) where {d}
    n = rule.series.norb
    e = copy ? Array{Float64}(undef, n, rule.nnodes) : nothing
    v = copy ? Array{Float64}(undef, n, d, rule.nnodes) : nothing
    check(rule.ctx, ccall((:abz_rule_ggr_data, LIB), Int32, (Ptr{Cvoid}, UInt64, Int32, Ptr{Float64}, Ptr{Float64}),
                          rule.ctx.h, rule.h, d, copy ? e : C_NULL, copy ? v : C_NULL))
    return e, v
end
function ggr_sum(rule::DeviceRule, E::Vector{Float64})
    out = similar(E)
    check(rule.ctx, ccall((:abz_rule_ggr_sum, LIB), Int32, (Ptr{Cvoid}, UInt64, Int32, Ptr{Float64}, Float64, Ptr{Float64}),
                          rule.ctx.h, rule.h, length(E), E, 1.0, out))
    return out
end

# generic user integrands: copy H(k) back and keep the reference's Julia path (iterate protocol, src/fourier.jl:176-202)
function copy_out(rule::DeviceRule{d}) where {d}
    n = rule.series.norb
    Hk = Array{ComplexF64}(undef, n, n, rule.nnodes); k = Array{Float64}(undef, 3, rule.nnodes); w = Vector{Float64}(undef, rule.nnodes)
    check(rule.ctx, ccall((:abz_rule_copy_out, LIB), Int32, (Ptr{Cvoid}, UInt64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                          rule.ctx.h, rule.h, Ptr{Float64}(pointer(Hk)), k, w))
    return Hk, k, w
end
function (rule::DeviceRule{d})(f::F, B::AutoSymPTR.Basis, buffer=nothing) where {d,F}
    Hk, k, w = copy_out(rule)
    n = rule.series.norb
    acc = sum(i -> w[i] * f(FourierValue(SVector{d}(k[1:d, i]), SMatrix{n,n}(view(Hk, :, :, i)))), 1:rule.nnodes)
    return acc * abs(det(B.B)) / (rule.npt^d * rule.nsyms)
end

# ---- S2'': Hermitian eigenvalues on the rule's nodes (eigen(Hermitian(h)), src/dos_ggr.jl:19,34; config C5) -------------
# kind 0: sum of eigenvalues, 1: Fermi-weighted band energy sum_n e_n f((e_n - mu)/T), 2: occupation sum_n f(..), 3: Gaussian DOS
# (ABZ_EIG_* in include/autobz_cuda.h); params = (mu, T) or (w, s)
function eig_sum(rule::DeviceRule, kind::Integer, params::Vector{Float64}, scale::Float64)
    out = zeros(1)
    check(rule.ctx, ccall((:abz_rule_eig_sum, LIB), Int32, (Ptr{Cvoid}, UInt64, Int32, Ptr{Float64}, Float64, Ptr{Float64}),
                          rule.ctx.h, rule.h, kind, params, scale, out))
    return out[1]
end
function eigvals(rule::DeviceRule)
    e = Matrix{Float64}(undef, rule.series.norb, rule.nnodes)
    check(rule.ctx, ccall((:abz_rule_eigvals, LIB), Int32, (Ptr{Cvoid}, UInt64, Ptr{Float64}), rule.ctx.h, rule.h, e))
    return e
end
# keep H(k) of this rule on the device between calls (the reference's cached rule, src/fourier.jl:344-360)
materialize!(rule::DeviceRule) = (check(rule.ctx, ccall((:abz_rule_materialize, LIB), Int32, (Ptr{Cvoid}, UInt64), rule.ctx.h, rule.h)); rule)

# algorithm options (ABZ_OPT_* in the header), e.g. set_option!(ctx, 1, 3): frequency sweep from one tridiagonalisation per k
set_option!(ctx::Context, option::Integer, value::Integer) =
    check(ctx, ccall((:abz_ctx_set_option, LIB), Int32, (Ptr{Cvoid}, Int32, Int64), ctx.h, option, value))

# ---- multi-GPU: one process per GPU, the k3 planes (PTR) or outermost panel nodes (IAI) dealt to ranks, ONE sum-allreduce
# per rule evaluation.  Either pass an MPI.Allreduce! closure as `exchange` to iai_solve / reduce the rule sums with MPI in
# Julia, or let the library own an NCCL communicator: rank 0 creates the id, everybody gets it (e.g. MPI.Bcast!), then
function unique_id()
    id = zeros(UInt8, 128)
    rc = ccall((:abz_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id)
    rc == 0 || throw(AbzError(rc, "abz_comm_unique_id"))
    return id
end
comm_init!(ctx::Context, rank::Integer, nranks::Integer, id::Vector{UInt8}) =
    check(ctx, ccall((:abz_comm_init, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), ctx.h, rank, nranks, id))
allreduce_sum!(ctx::Context, buf::Vector{Float64}) =
    (check(ctx, ccall((:abz_allreduce_sum, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), ctx.h, buf, length(buf))); buf)
comm_destroy!(ctx::Context) = check(ctx, ccall((:abz_comm_destroy, LIB), Int32, (Ptr{Cvoid},), ctx.h))

# ---- S3: BatchIntegrand f!(y, x, p) for scattered k (src/batch.jl:4-20) -------------------------------
function resolvent_batch!(y::Vector{ComplexF64}, x::Vector{SVector{3,Float64}}, ds::DeviceSeries, z::ComplexF64)
    resize!(y, length(x))
    zz = Float64[real(z), imag(z)]
    GC.@preserve y x begin
        rc = ccall((:abz_points_resolvent, LIB), Int32,
                   (Ptr{Cvoid}, UInt64, Int64, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   ds.ctx.h, ds.h, length(x), Ptr{Float64}(pointer(x)), 0, 1, zz, C_NULL, Ptr{Float64}(pointer(y)))
    end
    check(ds.ctx, rc)
    return nothing
end
gpu_batch_integrand(ds::DeviceSeries; η, max_batch=2^16) =
    BatchIntegrand((y, x, p) -> resolvent_batch!(y, x, ds, complex(p, η)), ComplexF64[], SVector{3,Float64}[], max_batch=max_batch)

# ---- S4: IAI levels: arena of contracted series (workspace_contract!, src/fourier.jl:478) --------------
mutable struct DeviceNest
    ctx::Context
    h::UInt64
end
function DeviceNest(ds::DeviceSeries, ndim, cap2, cap1)
    h = Ref{UInt64}(0)
    check(ds.ctx, ccall((:abz_nest_create, LIB), Int32, (Ptr{Cvoid}, UInt64, Int32, Int64, Int64, Ref{UInt64}), ds.ctx.h, ds.h, ndim, cap2, cap1, h))
    nest = DeviceNest(ds.ctx, h[])
    finalizer(x -> ccall((:abz_nest_destroy, LIB), Int32, (Ptr{Cvoid}, UInt64), x.ctx.h, x.h), nest)
    return nest
end
contract3!(n::DeviceNest, x3::Vector{Float64}, slot2::Vector{Int64}) =
    check(n.ctx, ccall((:abz_nest_contract3, LIB), Int32, (Ptr{Cvoid}, UInt64, Int64, Ptr{Float64}, Ptr{Int64}), n.ctx.h, n.h, length(x3), x3, slot2))
contract2!(n::DeviceNest, x2::Vector{Float64}, parent::Vector{Int64}, slot1::Vector{Int64}) =
    check(n.ctx, ccall((:abz_nest_contract2, LIB), Int32, (Ptr{Cvoid}, UInt64, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}), n.ctx.h, n.h, length(x2), x2, parent, slot1))
function eval!(y::Vector{ComplexF64}, n::DeviceNest, x1::Vector{Float64}, slot1::Vector{Int64}, z::ComplexF64)
    resize!(y, length(x1))
    zz = Float64[real(z), imag(z)]      # ComplexF64 as the two doubles the ABI takes (a Ref{ComplexF64} does not convert to Ptr{Float64})
    GC.@preserve y check(n.ctx, ccall((:abz_nest_eval, LIB), Int32, (Ptr{Cvoid}, UInt64, Int64, Ptr{Float64}, Ptr{Int64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       n.ctx.h, n.h, length(x1), x1, slot1, 0, zz, C_NULL, Ptr{Float64}(pointer(y))))
    return y
end

# H at the panel nodes for integrands evaluated in Julia (f.f(FourierValue(k, H), p), src/fourier.jl:452-456)
function eval_h!(Hk::Array{ComplexF64,3}, n::DeviceNest, x1::Vector{Float64}, slot1::Vector{Int64})
    check(n.ctx, ccall((:abz_nest_eval_h, LIB), Int32, (Ptr{Cvoid}, UInt64, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}),
                       n.ctx.h, n.h, length(x1), x1, slot1, Ptr{Float64}(pointer(Hk))))
    return Hk
end

# (z - H(k) - Sigma)^-1 at the nodes of innermost panels: the value type of `gloc_integrand` under IAI (docs/src/examples.md:90-106).
# G is n x n x npts, column-major per node.
function eval_matrix!(G::Array{ComplexF64,3}, n::DeviceNest, x1::Vector{Float64}, slot1::Vector{Int64}, z::ComplexF64,
                      Σ::Union{Nothing,Matrix{ComplexF64}}=nothing)
    zz = Float64[real(z), imag(z)]
    GC.@preserve G Σ check(n.ctx, ccall((:abz_nest_eval_matrix, LIB), Int32,
        (Ptr{Cvoid}, UInt64, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        n.ctx.h, n.h, length(x1), x1, slot1, zz, Σ === nothing ? C_NULL : Ptr{Float64}(pointer(Σ)), Ptr{Float64}(pointer(G))))
    return G
end

# The whole nested solve in one call (do_solve(f::FourierIntegrand, lims, ::NestedQuad), src/fourier.jl:493-510): same
# control flow as IteratedIntegration/QuadGK, run by the library's host engine.  lims::CubicLimits or TetrahedralLimits.
# vkind 0: tr G, 1: -Im(tr G)/pi (aps_example.jl:30).  exchange: @cfunction(allreduce!, Int32, (Ptr{Float64}, Int64, Ptr{Cvoid}))
# for multi-rank solves (outermost panel nodes dealt to ranks), or C_NULL to use the NCCL communicator of abz_comm_init.
# Tolerance defaults follow QuadGK / the reference (src/algorithms.jl:224-237): reltol = sqrt(eps) when abstol == 0, else 0.
function iai_solve(n::DeviceNest, lkind::Integer, la::Vector{Float64}, lb, z::ComplexF64; vkind=0, abstol=0.0,
                   reltol=(abstol == 0 ? sqrt(eps(Float64)) : 0.0),
                   maxiters=typemax(Int64) >> 1, device_leaves=true, rank=0, nranks=1, exchange=C_NULL)
    out = zeros(3); stats = zeros(Int64, 4)
    zz = Float64[real(z), imag(z)]
    check(n.ctx, ccall((:abz_iai_solve_sharded, LIB), Int32,
                       (Ptr{Cvoid}, UInt64, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                        Float64, Float64, Int64, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}),
                       n.ctx.h, n.h, lkind, la, lb === nothing ? C_NULL : lb, 0, vkind, zz, C_NULL, C_NULL,
                       abstol, reltol, maxiters, device_leaves ? 7 : 0,   # ABZ_IAI_DEVICE_LEAVES | _MIDDLES | ABZ_IAI_SPECULATE
                       rank, nranks, exchange, C_NULL, out, stats))
    return IntegralSolution(vkind == 1 ? out[1] : complex(out[1], out[2]), out[3], true, Int(stats[1]))
end

# General iterated limits (any IteratedIntegration.AbstractIteratedLimits, e.g. the Polyhedron3 of ext/SymmetryReduceBZExt.jl:33-58):
# the library asks for the breakpoints of the variable `dim` given the outer variables fixed so far (outermost first).
function _limits_cb(dim::Int32, xfixed::Ptr{Float64}, segs::Ptr{Float64}, maxseg::Int32, user::Ptr{Cvoid})::Int32
    try
        lims = unsafe_pointer_to_objref(user)[]
        nd = ndims(lims)
        cur = lims
        for k in 1:(nd - dim)                      # fixandeliminate from the outermost variable inwards
            cur = fixandeliminate(cur, unsafe_load(xfixed, k), Val(nd - k + 1))
        end
        sg = segments(cur, dim)
        length(sg) > maxseg && return Int32(-1)
        for (i, v) in enumerate(sg)
            unsafe_store!(segs, Float64(v), i)
        end
        return Int32(length(sg))
    catch
        return Int32(-2)
    end
end
function iai_solve_general(n::DeviceNest, lims, z::ComplexF64; vkind=0, abstol=0.0, reltol=(abstol == 0 ? sqrt(eps(Float64)) : 0.0),
                           maxiters=typemax(Int64) >> 1, device_leaves=true, rank=0, nranks=1, exchange=C_NULL)
    out = zeros(3); stats = zeros(Int64, 4)
    zz = Float64[real(z), imag(z)]
    box = Ref(lims)
    cb = @cfunction(_limits_cb, Int32, (Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Cvoid}))
    GC.@preserve box begin
        check(n.ctx, ccall((:abz_iai_solve_general, LIB), Int32,
                           (Ptr{Cvoid}, UInt64, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                            Float64, Float64, Int64, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}),
                           n.ctx.h, n.h, cb, pointer_from_objref(box), 0, vkind, zz, C_NULL, C_NULL,
                           abstol, reltol, maxiters, device_leaves ? 3 : 0, rank, nranks, exchange, C_NULL, out, stats))
    end
    return IntegralSolution(vkind == 1 ? out[1] : complex(out[1], out[2]), out[3], true, Int(stats[1]))
end

# v0.4+ API names (BASELINE.json north_star) as thin aliases
const FourierIntegralFunction = FourierIntegrand

end # module
