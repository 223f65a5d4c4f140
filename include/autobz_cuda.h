/*
 * autobz_cuda.h — C ABI of libautobz_cuda.so, the B200 (sm_100a) implementation of the
 * data-parallel hot path of AutoBZCore.jl v0.3.8:
 *   Wannier/Fourier interpolation H(k) = sum_R H_R exp(2 pi i k.R) on k-batches,
 *   per-k resolvent trace tr[(z - H(k) - Sigma)^-1] / Hermitian eigenvalues,
 *   and the quadrature-weighted k-sum.
 *
 * The reference has no FFI (it is pure Julia); every entry point below names the Julia call
 * site (file:line under the AutoBZCore.jl tree) whose arithmetic it replaces.  Julia binds these
 * with `ccall((:abz_..., "libautobz_cuda"), Int32, (...), ...)` (see INTEGRATION.md and
 * julia/AutoBZCUDA.jl); Python binds them with ctypes (autobzcore.jl_b200/_lib.py).
 *
 * Conventions
 *  - Every function returns int32: 0 = ABZ_OK, negative = ABZ_E_*; text via abz_last_error().
 *  - No exceptions, no callbacks, no exit().  Caller owns all host buffers; the library owns
 *    device memory behind opaque 64-bit handles with explicit *_destroy.
 *  - Complex data is interleaved (re, im) Float64, i.e. Julia ComplexF64 / numpy complex128.
 *  - coeffs layout  = ComplexF64[n, n, M1, M2, M3] column-major (Julia Array{SMatrix{n,n}} 3-d,
 *    src/fourier.jl:127-130); H(k) layout = ComplexF64[n, n, nodes] with nodes in the reference's
 *    iteration order (k1 fastest, src/fourier.jl:132-164, 216-263).
 *  - A ctx is single-threaded (one CUDA stream); distinct ctx objects are independent.
 *    Calls are synchronous with respect to returned host data.
 */
#ifndef AUTOBZ_CUDA_H
#define AUTOBZ_CUDA_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABZ_OK 0
#define ABZ_E_INVALID (-1)      /* invalid argument (ArgumentError in the reference) */
#define ABZ_E_OOM (-2)          /* device allocation failed / budget exceeded */
#define ABZ_E_CUDA (-3)         /* CUDA runtime error */
#define ABZ_E_SINGULAR (-4)     /* singular matrix or NaN/Inf in an integrand (QuadGK DomainError) */
#define ABZ_E_UNSUPPORTED (-5)  /* shape outside the supported range */
#define ABZ_E_NCCL (-6)         /* NCCL unavailable / failed */

typedef struct abz_ctx abz_ctx;
typedef uint64_t abz_series_t;  /* Fourier series (H_R coefficients) on the device */
typedef uint64_t abz_rule_t;    /* quadrature rule: node set + (optionally cached) H(k) */
typedef uint64_t abz_nest_t;    /* IAI arena: contracted series for nested panels */

/* integrand on H(k):  (user integrand f(FourierValue(k, H(k)), p), src/fourier.jl:120) */
#define ABZ_F_RESOLVENT_TRACE 0 /* tr[(z I - H - Sigma)^-1]  (aps_example/aps_example.jl:30, docs/src/examples.md:13-20) */
#define ABZ_F_TRACE_H 1         /* tr H(k)                  (linear test integrand, test/fourier.jl:41) */
/* eigenvalue integrands g(eig(Hermitian(H(k)))) (src/dos_ggr.jl:19,34) */
#define ABZ_EIG_SUM 0           /* sum_n e_n */
#define ABZ_EIG_FERMI_ENERGY 1  /* sum_n e_n f((e_n-mu)/T), params = {mu, T} */
#define ABZ_EIG_FERMI_COUNT 2   /* sum_n f((e_n-mu)/T) */
#define ABZ_EIG_GAUSS_DOS 3     /* sum_n exp(-((e_n-w)/s)^2)/(s sqrt(pi)), params = {w, s} */

/* resolvent algorithm selection (abz_ctx_set_option ABZ_OPT_RESOLVENT_ALGO) */
#define ABZ_OPT_RESOLVENT_ALGO 1   /* 0 auto, 1 generic pivoted Gauss-Jordan (register-resident; 4 = the earlier shared-memory
                                    * formulation, kept for cross-checks), 2 register/DMMA fast path, 3 frequency sweep from one
                                    * Householder tridiagonalisation per k: tr (z-H)^-1 = p'(z)/p(z), O(n) per frequency - opt-in, needs
                                    * Hermitian H(k) and a scalar (or no) self-energy; a matrix Sigma falls back to 0 */
#define ABZ_OPT_MEM_BUDGET_MB 2    /* device workspace budget for streamed chunks (default 4096) */
#define ABZ_OPT_FUSED_SMALL 3      /* 1 (default): fuse evaluation+resolvent for norb<=4 */
#define ABZ_OPT_IAI_LEAF_SPILL 5   /* segments per device-side innermost integral beyond the 63 held in shared memory (default 1024);
                                    * an integral that outgrows them is redone with host-driven panels (single rank); a negative value -c
                                    * (1 <= c <= 63) limits the TOTAL capacity to c segments - a test hook for that fallback */
#define ABZ_OPT_IAI_LANES 6        /* IAI rounds in flight in single-rank solves with norb <= 6 (default 4; 1 = one round at a time);
                                    * results and numevals do not depend on it */
#define ABZ_OPT_EIG_ALGO 4         /* 0 (default): Householder tridiagonalisation (warp-per-matrix in registers for norb <= 32,
                                    * CTA-per-matrix in registers for 33..64) followed by the eigenvalues of the tridiagonal: implicit QL
                                    * of the tridiagonal matrix for large batches, Sturm-count bisection (one warp per matrix) for batches
                                    * that would leave the thread-per-matrix QL kernel latency-bound; 1: cyclic two-sided Jacobi;
                                    * 2: as 0 but always the shared-memory tridiagonalisation (cross-check);
                                    * 3 / 4: as 0 but always QL / always bisection */

int32_t abz_version(void);
const char* abz_last_error(const abz_ctx* ctx);  /* ctx may be NULL: last error of abz_ctx_create */

int32_t abz_ctx_create(int32_t device, abz_ctx** out);
int32_t abz_ctx_destroy(abz_ctx* ctx);
int32_t abz_ctx_set_option(abz_ctx* ctx, int32_t option, int64_t value);
/* number of kernels launched by this ctx so far (bench.py "gpu_launches") */
int64_t abz_ctx_launch_count(const abz_ctx* ctx);
/* device time of the last matrix-function phase / evaluation phase in ms (CUDA events on the ctx stream) */
int32_t abz_ctx_last_timings(const abz_ctx* ctx, double* eval_ms, double* matfun_ms);

/* FourierSeries(C; period, offset): replaces FourierSeriesEvaluators.FourierSeries +
 * workspace_allocate_vec (src/fourier.jl:56-86).  lo[d] = lowest R index per dimension
 * (index + offset); is_complex = 0 means coeffs are Float64[n,n,M1,M2,M3].
 * ndim < 3 series are passed with trailing M = 1, lo = 0. */
int32_t abz_series_create(abz_ctx* ctx, const double* coeffs, int32_t is_complex, int32_t norb,
                          const int32_t M[3], const int32_t lo[3], const double period[3], abz_series_t* out);
/* retires the handle; rules and nests built on the series keep its coefficients alive until they are destroyed themselves */
int32_t abz_series_destroy(abz_ctx* ctx, abz_series_t s);

/* ---- S1: rule construction -------------------------------------------------------------- */
/* FourierPTR(w, T, Val(3), npt) (src/fourier.jl:166-174): full npt^3 grid, planes k3 in
 * [k3_lo, k3_hi) (the multi-GPU shard unit; the reference threads over the same loop, :156). */
int32_t abz_rule_create_full(abz_ctx* ctx, abz_series_t s, int32_t npt, int32_t k3_lo, int32_t k3_hi, abz_rule_t* out);
/* FourierMonkhorstPack(w, T, Val(3), npt, syms) (src/fourier.jl:265-277): symmetry-reduced nodes
 * given as AutoSymPTR.symptr_rule's wsym array Int32[npt,npt,npt] (i1 fastest; 0 = not a node,
 * else orbit size).  Only planes k3 = k3_lo + i*k3_stride < npt are taken (k3_stride = nranks
 * gives the reference's :scatter distribution, src/fourier.jl:246-255).  k3_stride = -nranks (k3_lo = rank) deals the planes
 * in serpentine order instead - rank r takes r, 2W-1-r, 2W+r, 4W-1-r, ... with W = nranks - which balances the monotonically
 * shrinking planes of an irreducible wedge (cubic group, npt = 96, W = 8: round-robin 22 % above the average, serpentine 4 %). */
int32_t abz_rule_create_sym(abz_ctx* ctx, abz_series_t s, int32_t npt, const int32_t* wsym,
                            int32_t k3_lo, int32_t k3_stride, abz_rule_t* out);
/* The same rule from an explicit node list (the reference's rule.wxs vector of (w, x) pairs,
 * src/fourier.jl:210-214): idx = Int32[3, nnodes] grid indices (i1, i2, i3) sorted by (i3, i2, i1),
 * w = Float64[nnodes] or NULL (all 1).  Used for 1-d / 2-d series and custom node sets. */
int32_t abz_rule_create_nodes(abz_ctx* ctx, abz_series_t s, int32_t npt, int64_t nnodes, const int32_t* idx,
                              const double* w, abz_rule_t* out);
/* AutoSymPTR.symptr_rule (call site src/fourier.jl:271) on the device: syms = Int32[3,3,nsyms]
 * row-major per matrix, wsym_out = Int32[npt^3] host buffer; returns the irreducible count.
 * The list must be a group (identity included, closed under products, no duplicates) - every list load_bz produces is;
 * ABZ_E_INVALID otherwise (the reference's sequential scan is order-dependent for non-groups). */
int32_t abz_symptr_rule(abz_ctx* ctx, int32_t npt, int32_t nsyms, const int32_t* syms, int32_t* wsym_out, int64_t* nirr);
/* symptr_rule + FourierMonkhorstPack in one step, entirely on the device (src/fourier.jl:265-277): the dense
 * weight array never visits the host, the CSR node lists are compacted by a warp-per-row kernel.  Same node set,
 * order and weights as abz_symptr_rule followed by abz_rule_create_sym.  nirr_total (may be NULL) = length(rule)
 * over ALL k3 planes, whatever (k3_lo, k3_stride) selects for this rank.  With nirr_total == NULL the orbit weights are
 * computed for the selected planes only (each grid point's orbit test is independent of the others): ranks that split the
 * planes then share the construction work as well and add up their abz_rule_info node counts to get length(rule). */
int32_t abz_rule_create_symptr(abz_ctx* ctx, abz_series_t s, int32_t npt, int32_t nsyms, const int32_t* syms, int32_t k3_lo,
                               int32_t k3_stride, abz_rule_t* out, int64_t* nirr_total);
int32_t abz_rule_destroy(abz_ctx* ctx, abz_rule_t r);
/* length(rule) (src/fourier.jl:177, 284) and norb */
int32_t abz_rule_info(abz_ctx* ctx, abz_rule_t r, int64_t* nnodes, int32_t* norb, int32_t* npt);
/* Evaluate and keep H(k) at every node on the device (the reference's cached rule.s / rule.wxs,
 * src/fourier.jl:127-130, 210-214).  ABZ_E_OOM if it does not fit the budget: then use the sums
 * below, which stream k3 chunks instead. */
int32_t abz_rule_materialize(abz_ctx* ctx, abz_rule_t r);
/* getindex/iterate of the rule (src/fourier.jl:176-202, 279-285) for generic user integrands on
 * the host: Hk = ComplexF64[n,n,nnodes], kfrac = Float64[3,nnodes], w = Float64[nnodes]; any may be NULL */
int32_t abz_rule_copy_out(abz_ctx* ctx, abz_rule_t r, double* Hk, double* kfrac, double* w);

/* ---- S2: rule application = quadsum(rule, f, scale) (src/fourier.jl:204-207, 289-292) ---- */
/* out[2*nw] = scale * sum_i w_i f(H(k_i); z_w, Sigma_w).  fkind = ABZ_F_*.  z = ComplexF64[nw];
 * sigma = ComplexF64[n,n,nw] or NULL.  Uses the cached H(k) when materialised, otherwise
 * evaluates chunk by chunk (fused; nothing of size nnodes*n^2 touches HBM for norb<=4). */
int32_t abz_rule_resolvent_sum(abz_ctx* ctx, abz_rule_t r, int32_t fkind, int32_t nw, const double* z,
                               const double* sigma, double scale, double* out);
/* matrix-valued integrand gloc_integrand(h_k; eta, omega) = inv(complex(omega, eta) I - h_k.s) (docs/src/examples.md:20,90):
 * out = ComplexF64[n,n,nw], out[:,:,w] = scale * sum_i w_i (z_w I - H(k_i) - Sigma_w)^-1 (pivoted Gauss-Jordan inverse).
 * On an IBZ the caller symmetrises the result with its SymRep (src/brillouin.jl:86-107; UnknownRep leaves it as is). */
int32_t abz_rule_resolvent_matrix_sum(abz_ctx* ctx, abz_rule_t r, int32_t nw, const double* z, const double* sigma, double scale,
                                      double* out);
/* out[0] = scale * sum_i w_i g(eigvals(H(k_i))), kind = ABZ_EIG_* */
int32_t abz_rule_eig_sum(abz_ctx* ctx, abz_rule_t r, int32_t kind, const double* params, double scale, double* out);
/* the same for nparams parameter sets at once (params = Float64[2, nparams], out = Float64[nparams]): every H(k) is diagonalised
 * ONCE per call and all parameters are summed over its eigenvalues - the device analogue of the reference's parameter sweep over a
 * shared cached grid (batchsolve, src/interfaces.jl:199-243).  On a materialised rule (abz_rule_materialize: the reference's cached
 * rule) the eigenvalues themselves stay cached on the device, so later calls only re-sum. */
int32_t abz_rule_eig_sum_batch(abz_ctx* ctx, abz_rule_t r, int32_t kind, int32_t nparams, const double* params, double scale,
                               double* out);
/* all eigenvalues, Float64[n, nnodes] ascending per node (GGR data pass, src/dos_ggr.jl:14-44) */
int32_t abz_rule_eigvals(abz_ctx* ctx, abz_rule_t r, double* evals);

/* ---- GGR density of states (src/dos_ggr.jl) ------------------------------------------------ */
/* get_ggr_data (src/dos_ggr.jl:14-44): at every node e = eigen(Hermitian(h)).values and the band velocities
 * v_d = real(diag(U' V_d U)) * period_d, V_d from JacobianSeries(h) (coefficient R times 2 pi i R_d / period_d),
 * d < ndim.  Cached on the device in the rule; optionally copied out: energies = Float64[n, nnodes] ascending,
 * velocities = Float64[n, ndim, nnodes] (either may be NULL). */
int32_t abz_rule_ggr_data(abz_ctx* ctx, abz_rule_t r, int32_t ndim, double* energies, double* velocities);
/* sum_ggr (src/dos_ggr.jl:58-104): out[i] = scale * sum_nodes w sum_bands ggr_formula(1/(2 npt), E_i, e, v...) */
int32_t abz_rule_ggr_sum(abz_ctx* ctx, abz_rule_t r, int32_t nE, const double* E, double scale, double* out);

/* ---- S3/S4: scattered nodes for IAI panels (src/fourier.jl:432-486) ----------------------- */
/* workspace_evaluate(w, x) at npts arbitrary points (x NOT scaled by the period, :454):
 * k = Float64[3,npts]; Hk = ComplexF64[n,n,npts] */
int32_t abz_points_eval(abz_ctx* ctx, abz_series_t s, int64_t npts, const double* k, double* Hk);
/* BatchIntegrand f!(y, x, p) for the resolvent (src/batch.jl:4-20): y[w + nw*i] = f(H(k_i); z_w) */
int32_t abz_points_resolvent(abz_ctx* ctx, abz_series_t s, int64_t npts, const double* k, int32_t fkind,
                             int32_t nw, const double* z, const double* sigma, double* y);
/* Nested panels: an arena of contracted series living on the device.
 * workspace_contract!(f.w, x) (src/fourier.jl:478): slot ids are caller-managed integers in
 * [0, capacity).  Level 2 slots hold the series contracted at x3; level 1 slots at (x3, x2). */
int32_t abz_nest_create(abz_ctx* ctx, abz_series_t s, int32_t ndim, int64_t cap2, int64_t cap1, abz_nest_t* out);
int32_t abz_nest_destroy(abz_ctx* ctx, abz_nest_t nest);
/* contract the root series at x3[i] into level-2 slot slot2[i], i < n (ndim == 3 only) */
int32_t abz_nest_contract3(abz_ctx* ctx, abz_nest_t nest, int64_t n, const double* x3, const int64_t* slot2);
/* contract level-2 slot parent[i] (or the root when ndim == 2) at x2[i] into level-1 slot slot1[i] */
int32_t abz_nest_contract2(abz_ctx* ctx, abz_nest_t nest, int64_t n, const double* x2, const int64_t* parent,
                           const int64_t* slot1);
/* innermost closure (src/fourier.jl:452-456): evaluate level-1 slot slot1[i] (or the root when
 * ndim == 1) at x1[i] and apply the integrand: y = ComplexF64[npts] */
int32_t abz_nest_eval(abz_ctx* ctx, abz_nest_t nest, int64_t npts, const double* x1, const int64_t* slot1,
                      int32_t fkind, const double* z, const double* sigma, double* y);

/* the same closure for a generic user integrand evaluated by the host (f.f(FourierValue(k, H(k)), p), src/fourier.jl:452-456;
 * BatchIntegrand's f!(y, x, p), src/batch.jl:4-20): only H at the nodes, Hk = ComplexF64[n,n,npts] */
int32_t abz_nest_eval_h(abz_ctx* ctx, abz_nest_t nest, int64_t npts, const double* x1, const int64_t* slot1, double* Hk);
/* the matrix-valued innermost closure, gloc_integrand(h_k; eta, omega) = inv(complex(omega, eta) I - h_k.s) under IAI
 * (docs/src/examples.md:90-106; the reference's nest is generic in the value type, src/fourier.jl:452-456): Y = ComplexF64[n,n,npts],
 * Y[:,:,i] = (z I - H(x1_i on level-1 slot slot1_i) - Sigma)^-1 by pivoted Gauss-Jordan; z = ComplexF64[1], sigma = ComplexF64[n,n] or NULL */
int32_t abz_nest_eval_matrix(abz_ctx* ctx, abz_nest_t nest, int64_t npts, const double* x1, const int64_t* slot1, const double* z,
                             const double* sigma, double* Y);
/* The whole nested adaptive solve do_solve(f::FourierIntegrand, lims, ::NestedQuad) (src/fourier.jl:493-510) with
 * GK(7,15) at every level (AuxQuadGKJL/QuadGKJL defaults, src/algorithms.jl:215-240): the adaptive control flow
 * (QuadGK do_quadgk/adapt/refine, DataStructures heap, inner abstol = abstol/len, src/fourier.jl:479-480) runs on
 * the library's host side, every round of live innermost panels is one device batch on `nest`'s arena.
 * For hosts whose own loop is too slow to drive abz_nest_* round by round (Python); a Julia host may use either.
 *   lkind 0: CubicLimits(la, lb);  1: TetrahedralLimits(la)  (load_bz(CubicSymIBZ), src/brillouin.jl:301-307)
 *   fkind = ABZ_F_*;  vkind 0: value = y;  1: -Im(y)/pi (aps_example.jl:30);  2: lin[0:2]*y + lin[2:4] (complex a, b)
 *   atol, rtol, maxevals apply to the outermost integral exactly as abstol/reltol/maxiters of the reference.
 *   flags: ABZ_IAI_DEVICE_LEAVES = run each innermost 1-D adaptive integral entirely on the device (norb <= 6)
 *   out = {Re I, Im I, E};  stats = Int64[4] {numevals (EvalCounter semantics), device rounds, kernel launches,
 *   exchanges} or NULL */
#define ABZ_IAI_DEVICE_LEAVES 1
/* ABZ_IAI_DEVICE_MIDDLES (implies the leaves): in 3-d solves over CubicLimits / TetrahedralLimits every MIDDLE integral (one per node of
 * the outermost panels) runs entirely on the device too - one CTA keeps its segment heap in shared memory, contracts the series at
 * each of its nodes and hands the innermost integrals to its warps - so the host engine only steps the outermost integral:
 * stats[1] (device rounds) drops from one per middle-level refinement to one per outermost refinement.  Same decisions, same numevals. */
#define ABZ_IAI_DEVICE_MIDDLES 2
/* ABZ_IAI_SPECULATE (with device leaves / middles; any number of ranks): when the outermost integral bisects its worst panel the engine also
 * starts the bisections QuadGK is likely to ask for next (the quarters of that panel and the panel next in its heap; environment
 * ABZ_IAI_LOOKAHEAD = 1 heap only, 2 quarters only, 3 both = default) and parks the results until QuadGK's own order reaches them
 * (dropped, and not counted, if it never does).  Decisions, integral, error estimate and numevals are those of the sequential
 * algorithm; a round carries several bisections, so stats[1] drops (C3: 80 -> 32).  If the attempt fails (e.g. a singular point inside a panel the
 * sequential algorithm would not have refined) the solve is repeated without look-ahead and that result / error is returned. */
#define ABZ_IAI_SPECULATE 4
int32_t abz_iai_solve(abz_ctx* ctx, abz_nest_t nest, int32_t lkind, const double* la, const double* lb, int32_t fkind,
                      int32_t vkind, const double* z, const double* sigma, const double* lin, double atol, double rtol,
                      int64_t maxevals, int32_t flags, double* out, int64_t* stats);

/* Multi-GPU IAI: the 15 nodes of every panel of the OUTERMOST integral are dealt round-robin to the
 * ranks; each rank integrates its own nodes and all ranks meet in one small sum-allreduce per outer refinement step,
 * then take the identical accept/refine decision (bit-identical result to the single-rank solve; stats[0] = evaluations
 * of all ranks).  The allreduce is `exchange` (in-place sum of n doubles over the ranks, return 0 on success; e.g. a
 * @cfunction around MPI.Allreduce!) or, when NULL, NCCL on the communicator of abz_comm_init.  Collective: every rank
 * must call it with the same arguments. */
typedef int32_t (*abz_exchange_fn)(double* buf, int64_t n, void* user);
int32_t abz_iai_solve_sharded(abz_ctx* ctx, abz_nest_t nest, int32_t lkind, const double* la, const double* lb, int32_t fkind,
                              int32_t vkind, const double* z, const double* sigma, const double* lin, double atol, double rtol,
                              int64_t maxevals, int32_t flags, int32_t rank, int32_t nranks, abz_exchange_fn exchange,
                              void* exchange_user, double* out, int64_t* stats);

/* The same solve over GENERAL iterated limits - what a caller gets from the reference's IBZ loader (Polyhedron3 / Polygon2 with
 * `segments` and `fixandeliminate`, ext/SymmetryReduceBZExt.jl:33-58; PuncturedInterval(segs), src/fourier.jl:493-500), or any
 * IteratedIntegration.AbstractIteratedLimits.  The geometry stays with the caller: `limits(dim, x_fixed, segs, maxseg, user)` must
 * write the ascending breakpoints of the variable of `dim` (1-based, dim = ndim is the outermost) given the outer variables fixed
 * so far (x_fixed[0] = outermost, ndim - dim values) - i.e. segments(fixandeliminate(...fixandeliminate(lims, x_ndim)..., x_{dim+1}))
 * - and return their number (2 ... maxseg = 64; a negative value aborts the solve with ABZ_E_INVALID).  Every 1-D integral starts
 * from all of its segments as QuadGK does (sum over the initial panels, heapify, then adapt); inner tolerances are abstol / (last
 * breakpoint - first breakpoint) (src/fourier.jl:476-480).  Device-side innermost integrals are used where the innermost range is a
 * single segment (always the case for a convex IBZ).  Other arguments as abz_iai_solve_sharded. */
typedef int32_t (*abz_limits_fn)(int32_t dim, const double* x_fixed, double* segs, int32_t maxseg, void* user);
int32_t abz_iai_solve_general(abz_ctx* ctx, abz_nest_t nest, abz_limits_fn limits, void* limits_user, int32_t fkind, int32_t vkind,
                              const double* z, const double* sigma, const double* lin, double atol, double rtol, int64_t maxevals,
                              int32_t flags, int32_t rank, int32_t nranks, abz_exchange_fn exchange, void* exchange_user, double* out,
                              int64_t* stats);

/* ---- multi-GPU: one small allreduce of partial sums ----------------------- */
/* NCCL is dlopen'ed at first use (libnccl.so.2).  uid = 128-byte ncclUniqueId from rank 0. */
int32_t abz_comm_unique_id(void* uid128);
int32_t abz_comm_init(abz_ctx* ctx, int32_t rank, int32_t nranks, const void* uid128);
int32_t abz_allreduce_sum(abz_ctx* ctx, double* host_buf, int64_t n);
int32_t abz_comm_destroy(abz_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
