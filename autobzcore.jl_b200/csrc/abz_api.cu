// libautobz_cuda.so — C ABI + host-side engine (chunk planning, launches) for the AutoBZCore.jl
// hot path on B200.  See include/autobz_cuda.h for the contract and the reference call sites.
#include "../../include/autobz_cuda.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <map>
#include <vector>

#include "abz_common.cuh"
#include "abz_eig.cuh"
#include "abz_iai.cuh"
#include "abz_kernels.cuh"
#include "abz_resolvent_mma.cuh"
#include "abz_resolvent_mma_team.cuh"
#include "abz_resolvent_gjreg.cuh"

using namespace abz;

namespace {

std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// Device-memory pool of a context: series / rule / arena buffers are recycled by (rounded) size instead of going through
// cudaMalloc / cudaFree, whose cost on a busy device is erratic (measured: 0.7 ms to 256 ms for the same teardown) and
// which would otherwise sit inside every solve that builds a rule (AutoPTR refinements, the end-to-end path).
cudaError_t pool_alloc(abz_ctx* ctx, void** p, size_t bytes);
void pool_free(abz_ctx* ctx, void* p);

struct Series {
    abz_ctx* ctx = nullptr;
    double2* c = nullptr;
    int n = 0, M[3] = {1, 1, 1}, lo[3] = {0, 0, 0};
    double period[3] = {1, 1, 1};
    ~Series() { pool_free(ctx, c); }
};

struct Rule {
    abz_ctx* ctx = nullptr;
    uint64_t series_id = 0;
    Series* s = nullptr;
    std::shared_ptr<Series> keep;   // a rule keeps its series alive: abz_series_destroy only retires the handle
    int N = 0;
    bool full = true;
    bool nodes_on_host = true;   // false: node_k1 / node_w live on the device only (abz_rule_create_symptr)
    long np3 = 0, nrows = 0, nnz = 0;
    std::vector<int> h_plane_k3, h_row_k2, h_node_k1;
    std::vector<long> h_plane_rowptr, h_row_nodeptr;
    std::vector<double> h_node_w;
    int* d_plane_k3 = nullptr; long* d_plane_rowptr = nullptr;
    int* d_row_k2 = nullptr; long* d_row_nodeptr = nullptr;
    int* d_node_k1 = nullptr; double* d_node_w = nullptr;
    double2* d_ptab[3] = {nullptr, nullptr, nullptr};
    double2* d_H = nullptr;   // materialised H(k) [nnz][n*n]
    double* d_eig = nullptr;   // eigenvalues [nnz][n] of a materialised rule (cached across parameters, like the reference's cached grid)
    double* d_ggr_e = nullptr; double* d_ggr_v = nullptr; int ggr_ndim = 0;   // GGR data pass: energies [nnz][n], velocities [nnz][ndim][n]
    ~Rule() {
        pool_free(ctx, d_ggr_e); pool_free(ctx, d_ggr_v); pool_free(ctx, d_eig);
        pool_free(ctx, d_plane_k3); pool_free(ctx, d_plane_rowptr); pool_free(ctx, d_row_k2); pool_free(ctx, d_row_nodeptr);
        pool_free(ctx, d_node_k1); pool_free(ctx, d_node_w); pool_free(ctx, d_H);
        for (auto& p : d_ptab) pool_free(ctx, p);
    }
};

struct Nest {
    abz_ctx* ctx = nullptr;
    Series* s = nullptr;
    std::shared_ptr<Series> keep;   // as Rule::keep
    int ndim = 3;
    long cap2 = 0, cap1 = 0;
    double2* L2 = nullptr; double2* L1 = nullptr;
    ~Nest() { pool_free(ctx, L2); pool_free(ctx, L1); }
};

struct Chunk { long p0, p1, r0, r1; };
// IAI kernels that evaluate the innermost series and the integrand per thread (nest_*_small, iai_leaf, iai_mid): norb <= IAI_SMALL_MAXN
// (closed-form adjugate up to 3, in-register pivoted Gauss-Jordan up to 6); rule sums keep their own n <= 3 fused path
constexpr int IAI_SMALL_MAXN = 6;

// one in-flight round of the IAI engine (abz_iai_engine.hpp "lanes"): its own stream, staging and scratch buffers
struct IaiLane {
    cudaStream_t stream = nullptr;
    void* pin_in = nullptr; size_t pin_in_cap = 0;
    void* pin_out = nullptr; size_t pin_out_cap = 0;
    DevBuf in, out, spill, mid_spill;
    size_t ns = 0, nt = 0, nm = 0;
    ~IaiLane() {
        if (pin_in) cudaFreeHost(pin_in);
        if (pin_out) cudaFreeHost(pin_out);
        if (stream) cudaStreamDestroy(stream);
    }
};
constexpr int ABZ_RETRY_PIVOTED = 1;   // internal: the unpivoted fast path saw a tiny pivot; rerun with the pivoted kernel

}  // namespace

struct abz_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::unordered_map<void*, size_t> pool_size;                    // every live or pooled block -> rounded size
    std::map<size_t, std::vector<void*>> pool_free_list;             // rounded size -> free blocks (ordered: best fit)
    size_t pool_held = 0, pool_cap = (size_t)16 << 30;              // bytes parked in the free lists / their limit
    std::string err;
    uint64_t next_id = 1;
    std::unordered_map<uint64_t, std::shared_ptr<Series>> series;
    std::unordered_map<uint64_t, std::unique_ptr<Rule>> rules;
    std::unordered_map<uint64_t, std::unique_ptr<Nest>> nests;
    int resolvent_algo = 0;
    size_t budget = (size_t)4096 << 20;
    int fused_small = 1;
    int eig_algo = 0;             // 0: tridiagonalisation + QL / bisection, 1: two-sided Jacobi, 2: shared-memory tridiagonalisation, 3 / 4: force QL / bisection
    int iai_lanes_opt = 4;        // IAI rounds in flight (ABZ_OPT_IAI_LANES)
    int leaf_spill = LEAF_SPILL;  // segments per device-side innermost integral beyond the 63 kept in shared memory
    bool force_generic = false;   // set while re-running a call whose fast path asked for pivoting
    DevBuf C2, C1, Hc, partial, acc, zbuf, sigbuf, errflag, tmp_a, tmp_b, tmp_c, tmp_d, iai_in, iai_out, eig_d, eig_e, eig_trail, symw, symlist;
    void* pin_in = nullptr; size_t pin_in_cap = 0;     // pinned staging for the IAI engine's per-round traffic
    void* pin_out = nullptr; size_t pin_out_cap = 0;
    long launches = 0;
    std::vector<std::unique_ptr<IaiLane>> iai_lanes;
    std::vector<cudaEvent_t> events;
    size_t ev_used = 0;
    double eval_ms = 0, matfun_ms = 0;
    int sm_count = 148;
    // nccl (dlopen)
    void* nccl_lib = nullptr; void* nccl_comm = nullptr; int nranks = 1;
};

namespace {

int fail(abz_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}

size_t pool_round(size_t b) { const size_t g = b > ((size_t)1 << 20) ? ((size_t)1 << 20) : 4096; return ((b + g - 1) / g) * g; }

cudaError_t pool_alloc(abz_ctx* ctx, void** p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) return cudaSuccess;
    const size_t r = pool_round(bytes);
    // best fit among the parked blocks, wasting at most the request again (a changed workload reuses what is parked)
    for (auto it = ctx->pool_free_list.lower_bound(r); it != ctx->pool_free_list.end() && it->first <= 2 * r; ++it) {
        if (it->second.empty()) continue;
        *p = it->second.back();
        it->second.pop_back();
        ctx->pool_held -= it->first;
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, r);
    if (e != cudaSuccess) {            // out of memory: give the parked blocks back to the driver and retry once
        cudaGetLastError();
        for (auto& kv : ctx->pool_free_list) { for (void* q : kv.second) { cudaFree(q); ctx->pool_size.erase(q); } kv.second.clear(); }
        ctx->pool_held = 0;
        e = cudaMalloc(p, r);
    }
    if (e == cudaSuccess) ctx->pool_size[*p] = r; else *p = nullptr;
    return e;
}

void pool_free(abz_ctx* ctx, void* p) {
    if (!p) return;
    if (!ctx) { cudaFree(p); return; }
    auto it = ctx->pool_size.find(p);
    if (it == ctx->pool_size.end()) { cudaFree(p); return; }
    const size_t r = it->second;
    if (r > ctx->pool_cap) { cudaFree(p); ctx->pool_size.erase(it); return; }
    // over the limit: the most recently used block stays, the largest parked blocks go back to the driver
    while (ctx->pool_held + r > ctx->pool_cap) {
        auto big = ctx->pool_free_list.end();
        bool evicted = false;
        while (big != ctx->pool_free_list.begin()) {
            --big;
            if (big->second.empty()) continue;
            void* q = big->second.back();
            big->second.pop_back();
            cudaFree(q);
            ctx->pool_size.erase(q);
            ctx->pool_held -= big->first;
            evicted = true;
            break;
        }
        if (!evicted) break;
    }
    ctx->pool_free_list[r].push_back(p);       // work that used p is ordered before any reuse: one stream per context
    ctx->pool_held += r;
}

void pool_release(abz_ctx* ctx) {
    for (auto& kv : ctx->pool_free_list) for (void* q : kv.second) cudaFree(q);
    ctx->pool_free_list.clear(); ctx->pool_size.clear(); ctx->pool_held = 0;
}

#define CU(ctx, expr)                                                                                     \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            cudaGetLastError();                                                                           \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? ABZ_E_OOM : ABZ_E_CUDA,                  \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                             \
        }                                                                                                 \
    } while (0)

#define LAUNCH_CHECK(ctx, name)                                                                           \
    do {                                                                                                  \
        (ctx)->launches++;                                                                                \
        cudaError_t e__ = cudaGetLastError();                                                             \
        if (e__ != cudaSuccess) return fail(ctx, ABZ_E_CUDA, std::string(name) + " launch: " + cudaGetErrorString(e__)); \
    } while (0)

cudaEvent_t next_event(abz_ctx* ctx) {
    if (ctx->ev_used == ctx->events.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->events.push_back(e);
    }
    cudaEvent_t e = ctx->events[ctx->ev_used++];
    cudaEventRecord(e, ctx->stream);
    return e;
}

template <class T>
int upload(abz_ctx* ctx, T** dptr, const std::vector<T>& h) {
    *dptr = nullptr;
    if (h.empty()) return ABZ_OK;
    CU(ctx, pool_alloc(ctx, (void**)dptr, h.size() * sizeof(T)));
    CU(ctx, cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return ABZ_OK;
}

Series* get_series(abz_ctx* ctx, abz_series_t id) {
    auto it = ctx->series.find(id);
    return it == ctx->series.end() ? nullptr : it->second.get();
}
Rule* get_rule(abz_ctx* ctx, abz_rule_t id) {
    auto it = ctx->rules.find(id);
    return it == ctx->rules.end() ? nullptr : it->second.get();
}
Nest* get_nest(abz_ctx* ctx, abz_nest_t id) {
    auto it = ctx->nests.find(id);
    return it == ctx->nests.end() ? nullptr : it->second.get();
}

int finish_rule(abz_ctx* ctx, Rule* r) {
    Series* s = r->s;
    int rc;
    if ((rc = upload(ctx, &r->d_plane_k3, r->h_plane_k3))) return rc;
    if ((rc = upload(ctx, &r->d_plane_rowptr, r->h_plane_rowptr))) return rc;
    if ((rc = upload(ctx, &r->d_row_k2, r->h_row_k2))) return rc;
    if ((rc = upload(ctx, &r->d_row_nodeptr, r->h_row_nodeptr))) return rc;
    if (!r->full && r->nodes_on_host) {
        if ((rc = upload(ctx, &r->d_node_k1, r->h_node_k1))) return rc;
        if ((rc = upload(ctx, &r->d_node_w, r->h_node_w))) return rc;
    }
    for (int d = 0; d < 3; d++) {
        size_t cnt = (size_t)s->M[d] * r->N;
        CU(ctx, pool_alloc(ctx, (void**)&r->d_ptab[d], cnt * sizeof(double2)));
        phase_table_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(r->d_ptab[d], s->M[d], s->lo[d], r->N);
        LAUNCH_CHECK(ctx, "phase_table_kernel");
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ABZ_OK;
}

// number of planes share_plane(p, k3_lo, k3_stride) < N (the sequence increases with p)
static long share_nplanes(long N, int k3_lo, int k3_stride) {
    long n = 0;
    while (share_plane((int)n, k3_lo, k3_stride) < N) n++;
    return n;
}
static bool share_valid(int k3_lo, int k3_stride) { return k3_lo >= 0 && (k3_stride >= 1 || (k3_stride < 0 && k3_lo < -k3_stride)); }

// ---- chunk planning: consecutive planes while they fit; a plane larger than the cap is split by rows
std::vector<Chunk> plan_chunks(const Rule* r, long node_cap, long row_cap, long plane_cap) {
    std::vector<Chunk> out;
    node_cap = std::max<long>(node_cap, 1); row_cap = std::max<long>(row_cap, 1); plane_cap = std::max<long>(plane_cap, 1);
    long p = 0;
    while (p < r->np3) {
        long r0 = r->h_plane_rowptr[p];
        long pr1 = r->h_plane_rowptr[p + 1];
        long pn = r->h_row_nodeptr[pr1] - r->h_row_nodeptr[r0];
        if (pn > node_cap || pr1 - r0 > row_cap) {
            long ra = r0;
            while (ra < pr1) {
                long rb = ra + 1;
                while (rb < pr1 && rb - ra < row_cap && r->h_row_nodeptr[rb + 1] - r->h_row_nodeptr[ra] <= node_cap) rb++;
                out.push_back({p, p + 1, ra, rb});
                ra = rb;
            }
            p++;
            continue;
        }
        long q = p + 1;
        while (q < r->np3 && q - p < plane_cap) {
            long qr1 = r->h_plane_rowptr[q + 1];
            if (qr1 - r0 > row_cap || r->h_row_nodeptr[qr1] - r->h_row_nodeptr[r0] > node_cap) break;
            q++;
        }
        out.push_back({p, q, r0, r->h_plane_rowptr[q]});
        p = q;
    }
    return out;
}

size_t stage_smem(int M) { return (size_t)((M + 1) & ~1) * (ST_RT + 2 * ST_JT) * sizeof(double2) + 64; }

int launch_stage(abz_ctx* ctx, const double2* in, double2* out, const double2* ptab, const long* ptr, long b0,
                 long nbatch, const int* klist, int N, int M, long rows, long in_stride) {
    if (nbatch <= 0 || rows <= 0) return ABZ_OK;
    size_t smem = stage_smem(M);
    if (smem > 200 * 1024) return fail(ctx, ABZ_E_UNSUPPORTED, "series has too many coefficients per dimension");
    dim3 grid((unsigned)nbatch, (unsigned)((rows + ST_RT - 1) / ST_RT));
    contract_stage_kernel<<<grid, 128, smem, ctx->stream>>>(in, out, ptab, ptr, b0, klist, N, M, rows, in_stride);
    LAUNCH_CHECK(ctx, "contract_stage_kernel");
    return ABZ_OK;
}

// evaluate stages for a chunk.  upto = 1: stop after stage 2 (C1 rows; fused small path); 0: also H
int eval_chunk(abz_ctx* ctx, Rule* r, const Chunk& ch, bool need_h, double2* Hdst, const double2* coeffs = nullptr) {
    Series* s = r->s;
    if (!coeffs) coeffs = s->c;
    const long nn = (long)s->n * s->n;
    const long rows2 = nn * s->M[0] * s->M[1], rows1 = nn * s->M[0];
    const long nplanes = ch.p1 - ch.p0, nrows = ch.r1 - ch.r0;
    CU(ctx, ctx->C2.reserve((size_t)nplanes * rows2 * sizeof(double2)));
    CU(ctx, ctx->C1.reserve((size_t)nrows * rows1 * sizeof(double2)));
    // stage 3: one batch over planes [p0,p1): use a 2-entry ptr built on the fly in tmp_d
    long h_ptr3[2] = {ch.p0, ch.p1};
    CU(ctx, ctx->tmp_d.reserve(4 * sizeof(long)));
    CU(ctx, cudaMemcpyAsync(ctx->tmp_d.p, h_ptr3, sizeof(h_ptr3), cudaMemcpyHostToDevice, ctx->stream));
    int rc = launch_stage(ctx, coeffs, ctx->C2.as<double2>(), r->d_ptab[2], ctx->tmp_d.as<long>(), 0, 1, r->d_plane_k3, r->N,
                          s->M[2], rows2, 0);
    if (rc) return rc;
    // stage 2
    bool whole = (ch.r0 == r->h_plane_rowptr[ch.p0] && ch.r1 == r->h_plane_rowptr[ch.p1]);
    if (whole) {
        rc = launch_stage(ctx, ctx->C2.as<double2>(), ctx->C1.as<double2>(), r->d_ptab[1], r->d_plane_rowptr, ch.p0, nplanes,
                          r->d_row_k2, r->N, s->M[1], rows1, rows2);
    } else {
        long h_ptr2[2] = {ch.r0, ch.r1};
        CU(ctx, cudaMemcpyAsync(ctx->tmp_d.as<long>() + 2, h_ptr2, sizeof(h_ptr2), cudaMemcpyHostToDevice, ctx->stream));
        rc = launch_stage(ctx, ctx->C2.as<double2>(), ctx->C1.as<double2>(), r->d_ptab[1], ctx->tmp_d.as<long>() + 2, 0, 1,
                          r->d_row_k2, r->N, s->M[1], rows1, rows2);
    }
    if (rc) return rc;
    if (!need_h) return ABZ_OK;
    // stage 1
    rc = launch_stage(ctx, ctx->C1.as<double2>(), Hdst, r->d_ptab[0], r->d_row_nodeptr, ch.r0, nrows, r->d_node_k1, r->N,
                      s->M[0], nn, rows1);
    return rc;
}

int check_errflag(abz_ctx* ctx, const char* what) {
    int h = 0;
    CU(ctx, cudaMemcpyAsync(&h, ctx->errflag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (h) {
        cudaMemsetAsync(ctx->errflag.p, 0, sizeof(int), ctx->stream);
        if (h & 8) return fail(ctx, ABZ_E_INVALID, std::string(what) + ": H(k) is not Hermitian - the frequency-sweep resolvent "
                                                   "(ABZ_OPT_RESOLVENT_ALGO = 3) does not apply; use algorithm 0");
        if ((h & 2) && !(h & 1) && !ctx->force_generic) return ABZ_RETRY_PIVOTED;
        return fail(ctx, ABZ_E_SINGULAR, std::string(what) + ": singular matrix or NaN/Inf in the integrand");
    }
    return ABZ_OK;
}

// Householder tridiagonalisation of nk materialised matrices into ctx->eig_d / eig_e (structure of arrays):
// n <= 32: one warp per matrix with the rows in registers; else one CTA per matrix in shared memory
static int launch_tridiag(abz_ctx* ctx, const double2* H, long nk, int n, int* herm_flag = nullptr) {
    CU(ctx, ctx->eig_d.reserve((size_t)nk * n * sizeof(double)));
    CU(ctx, ctx->eig_e.reserve((size_t)nk * n * sizeof(double)));
    double* dd = ctx->eig_d.as<double>(); double* ee = ctx->eig_e.as<double>();
    if (n <= 32 && ctx->eig_algo != 2) {
        const long nblk = std::min<long>((nk + 3) / 4, (long)ctx->sm_count * 16);
        static int minb = 0;      // resident CTAs per SM the compiler budgets registers for (2: 255 registers, 3: 168 + a few spills)
        if (!minb) { const char* e = getenv("ABZ_TRIDIAG_MINB"); minb = (e && atoi(e) == 2) ? 2 : 3; }
#define TRIDIAG_WARP(NX)                                                                                    \
    do {                                                                                                    \
        if (minb == 2) eig_tridiag_warp_kernel<NX, 2><<<(unsigned)nblk, 128, 0, ctx->stream>>>(H, nk, n, dd, ee, herm_flag); \
        else eig_tridiag_warp_kernel<NX, 3><<<(unsigned)nblk, 128, 0, ctx->stream>>>(H, nk, n, dd, ee, herm_flag);           \
    } while (0)
        if (n <= 8) TRIDIAG_WARP(8);
        else if (n <= 16) TRIDIAG_WARP(16);
        else if (n <= 24) TRIDIAG_WARP(24);
        else TRIDIAG_WARP(32);
#undef TRIDIAG_WARP
        LAUNCH_CHECK(ctx, "eig_tridiag_warp_kernel");
        return ABZ_OK;
    }
    if (n > 32 && n <= 64 && ctx->eig_algo != 2) {
        const long ncta = std::min<long>(nk, (long)ctx->sm_count * 2);
        static const bool split = getenv("ABZ_TRIDIAG_SPLIT") ? atoi(getenv("ABZ_TRIDIAG_SPLIT")) != 0 : true;
        if (!split) {
            eig_tridiag_reg64_kernel<63><<<(unsigned)ncta, 256, 0, ctx->stream>>>(H, nk, n, dd, ee, herm_flag, nullptr, 0);
            LAUNCH_CHECK(ctx, "eig_tridiag_reg64_kernel");
            return ABZ_OK;
        }
        // the first 32 reflections with the CTA-per-matrix kernel, the trailing (n - 32)^2 block with the warp-per-matrix kernel
        const int n2 = n - 32;
        CU(ctx, ctx->eig_trail.reserve((size_t)nk * n2 * n2 * sizeof(double2)));
        double2* tr = ctx->eig_trail.as<double2>();
        eig_tridiag_reg64_kernel<32><<<(unsigned)ncta, 256, 0, ctx->stream>>>(H, nk, n, dd, ee, herm_flag, tr, n2);
        LAUNCH_CHECK(ctx, "eig_tridiag_reg64_kernel");
        const long nblk2 = std::min<long>((nk + 3) / 4, (long)ctx->sm_count * 16);
        double* d2 = dd + 32 * nk; double* e2 = ee + 32 * nk;
        if (n2 <= 8) eig_tridiag_warp_kernel<8, 3><<<(unsigned)nblk2, 128, 0, ctx->stream>>>(tr, nk, n2, d2, e2, nullptr);
        else if (n2 <= 16) eig_tridiag_warp_kernel<16, 3><<<(unsigned)nblk2, 128, 0, ctx->stream>>>(tr, nk, n2, d2, e2, nullptr);
        else if (n2 <= 24) eig_tridiag_warp_kernel<24, 3><<<(unsigned)nblk2, 128, 0, ctx->stream>>>(tr, nk, n2, d2, e2, nullptr);
        else eig_tridiag_warp_kernel<32, 3><<<(unsigned)nblk2, 128, 0, ctx->stream>>>(tr, nk, n2, d2, e2, nullptr);
        LAUNCH_CHECK(ctx, "eig_tridiag_warp_kernel");
        return ABZ_OK;
    }
    const int RP = n > 32 ? 64 : 32;
    const size_t smem = ((size_t)n * RP + RP + (size_t)(4 * RP / 32) * (RP - 1)) * sizeof(double2);
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / (4 * RP), (size_t)232448 / (smem + 1024)));
    const long ncta = std::min<long>(nk, (long)ctx->sm_count * per_sm);
    if (RP == 32) eig_tridiag_kernel<32><<<(unsigned)ncta, 128, smem, ctx->stream>>>(H, nk, n, dd, ee, herm_flag);
    else eig_tridiag_kernel<64><<<(unsigned)ncta, 256, smem, ctx->stream>>>(H, nk, n, dd, ee, herm_flag);
    LAUNCH_CHECK(ctx, "eig_tridiag_kernel");
    return ABZ_OK;
}

int upload_params(abz_ctx* ctx, int n, int nw, const double* z, const double* sigma) {
    CU(ctx, ctx->zbuf.reserve((size_t)std::max(nw, 1) * sizeof(double2)));
    if (z) CU(ctx, cudaMemcpyAsync(ctx->zbuf.p, z, (size_t)nw * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
    if (sigma) {
        size_t b = (size_t)nw * n * n * sizeof(double2);
        CU(ctx, ctx->sigbuf.reserve(b));
        CU(ctx, cudaMemcpyAsync(ctx->sigbuf.p, sigma, b, cudaMemcpyHostToDevice, ctx->stream));
    }
    return ABZ_OK;
}

// register-resident Gauss-Jordan teams (abz_resolvent_gjreg.cuh): variant 0 = 32 columns per thread, 168 registers (12 warps per SM),
// 1 = 32 columns, 255 registers (8 warps, deeper prefetch of the pivot row), 2 = 16 columns per thread (teams twice as wide, 16 warps)
// Defaults from measurements (profiles/r01_generic_resolvent_timing.log): traces 0, matrix-valued sums 2; ABZ_GJ_VARIANT overrides.
static int gj_variant(bool matrix) {
    static const int env = getenv("ABZ_GJ_VARIANT") ? atoi(getenv("ABZ_GJ_VARIANT")) : -1;
    if (env >= 0 && env <= 2) return env;
    return matrix ? 2 : 0;
}
static int gj_ctas_per_sm(int n, bool matrix) {
    const int v = gj_variant(matrix);
    if (n <= 32) return v == 0 ? 12 : 8;
    return v == 0 ? 3 : 2;
}
#define GJ_DISPATCH(KERNEL, matrix, n, grid, smem, stream, ...)                                                       \
    do {                                                                                                       \
        const int v_ = gj_variant(matrix);                                                                        \
        if ((n) <= 32) {                                                                                       \
            if (v_ == 0) KERNEL<32, 32, 4, 12><<<grid, 32, smem, stream>>>(__VA_ARGS__);                       \
            else if (v_ == 1) KERNEL<32, 32, 16, 8><<<grid, 32, smem, stream>>>(__VA_ARGS__);                  \
            else KERNEL<32, 16, 8, 8><<<grid, 64, smem, stream>>>(__VA_ARGS__);                                \
        } else {                                                                                               \
            if (v_ == 0) KERNEL<64, 32, 4, 3><<<grid, 128, smem, stream>>>(__VA_ARGS__);                       \
            else if (v_ == 1) KERNEL<64, 32, 16, 2><<<grid, 128, smem, stream>>>(__VA_ARGS__);                 \
            else KERNEL<64, 16, 8, 2><<<grid, 256, smem, stream>>>(__VA_ARGS__);                               \
        }                                                                                                      \
    } while (0)

size_t gj_smem_bytes(int n, int nw, int nwarps) {
    size_t per_warp = (size_t)n * (n + 1) + (n + 1) / 2 + 1;
    return ((size_t)n * n + nw + per_warp * nwarps) * sizeof(double2);
}

// matrix function on nk materialised matrices H (device).  mode 0: weighted partial sums -> acc (+=),
// mode 1: values -> yout (device, [nk][nw])
int run_matfun(abz_ctx* ctx, const double2* H, const double* wnode, long nk, int n, int fkind, int nw,
               const double2* z, const double2* sigma, int mode, double2* yout) {
    if (nk <= 0) return ABZ_OK;
    const long sm = ctx->sm_count;
    if (fkind == ABZ_F_TRACE_H) {
        long ncta = std::min<long>((nk + 255) / 256, sm * 8);
        if (mode == 0) {
            CU(ctx, ctx->partial.reserve((size_t)ncta * nw * sizeof(double2)));
            trace_h_kernel<<<(unsigned)ncta, 256, 0, ctx->stream>>>(H, wnode, nk, n, 0, ctx->partial.as<double2>(), nw);
            LAUNCH_CHECK(ctx, "trace_h_kernel");
            reduce_partials_kernel<<<nw, 256, 0, ctx->stream>>>(ctx->partial.as<double2>(), ncta, nw, 1.0, ctx->acc.as<double2>());
            LAUNCH_CHECK(ctx, "reduce_partials_kernel");
        } else {
            trace_h_kernel<<<(unsigned)ncta, 256, 0, ctx->stream>>>(H, wnode, nk, n, 1, yout, nw);
            LAUNCH_CHECK(ctx, "trace_h_kernel");
        }
        return ABZ_OK;
    }
    int* ef = ctx->errflag.as<int>();
    if (n <= 3) {
        if (mode == 1) {
            unsigned g = (unsigned)((nk + SM_THREADS - 1) / SM_THREADS);
            if (n == 1) small_values_kernel<1><<<g, SM_THREADS, 0, ctx->stream>>>(H, nk, fkind, nw, z, sigma, yout, ef);
            else if (n == 2) small_values_kernel<2><<<g, SM_THREADS, 0, ctx->stream>>>(H, nk, fkind, nw, z, sigma, yout, ef);
            else small_values_kernel<3><<<g, SM_THREADS, 0, ctx->stream>>>(H, nk, fkind, nw, z, sigma, yout, ef);
            LAUNCH_CHECK(ctx, "small_values_kernel");
            return ABZ_OK;
        }
        // sums over a materialised H: handled by the caller through small_fused_kernel<_, true>
        return fail(ctx, ABZ_E_INVALID, "internal: small-norb sums go through run_small_fused");
    }
    // frequency sweep from one tridiagonalisation per k (opt-in: Hermitian H(k), scalar self-energy folded into z)
    if (ctx->resolvent_algo == 3 && !sigma && n <= EIG_MAXN && !ctx->force_generic) {
        { int rct = launch_tridiag(ctx, H, nk, n, ef); if (rct) return rct; }
        const long ncx = (nk + TS_THREADS - 1) / TS_THREADS;
        dim3 grid((unsigned)ncx, (unsigned)((nw + TS_WCH - 1) / TS_WCH));
        if (mode == 0) {
            CU(ctx, ctx->partial.reserve((size_t)ncx * nw * sizeof(double2)));
            tridiag_resolvent_kernel<<<grid, TS_THREADS, 0, ctx->stream>>>(ctx->eig_d.as<double>(), ctx->eig_e.as<double>(), wnode, nk, n, nw, z, 0,
                                                                          ctx->partial.as<double2>(), ef);
            LAUNCH_CHECK(ctx, "tridiag_resolvent_kernel");
            reduce_partials_kernel<<<nw, 256, 0, ctx->stream>>>(ctx->partial.as<double2>(), ncx, nw, 1.0, ctx->acc.as<double2>());
            LAUNCH_CHECK(ctx, "reduce_partials_kernel");
        } else {
            tridiag_resolvent_kernel<<<grid, TS_THREADS, 0, ctx->stream>>>(ctx->eig_d.as<double>(), ctx->eig_e.as<double>(), wnode, nk, n, nw, z, 1,
                                                                          yout, ef);
            LAUNCH_CHECK(ctx, "tridiag_resolvent_kernel");
        }
        return ABZ_OK;
    }
    // DMMA register-resident fast path (norb <= 32; unpivoted block elimination with growth monitoring)
    bool use_mma = (ctx->resolvent_algo != 1 && ctx->resolvent_algo != 4) && !ctx->force_generic && mma_resolvent_supported(n) &&
                   !(getenv("ABZ_MMA_TEAM") && atoi(getenv("ABZ_MMA_TEAM")) == 1 && n > 24);
    if (use_mma) {
        long ncta = 0; int kper_m = 1;
        if (mma_resolvent_plan(n, nk, nw, sm, &ncta, &kper_m) == 0) {
            static const int fused_mma_env = getenv("ABZ_FUSED_MMA") ? atoi(getenv("ABZ_FUSED_MMA")) : 1;
            if (mode == 0 && fused_mma_env && nw >= 8 && mma_resolvent_variant() == ABZ_MMA_DEFAULT_VARIANT && mma_resolvent_warps() >= 8 &&
                mma_fused_smem(n, nw) <= 160 * 1024) {
                // K3-fused in direct mode: every node's H(k) is staged once in shared memory for all of its frequencies
                CU(ctx, ctx->partial.reserve((size_t)ncta * nw * sizeof(double2)));
                CU(ctx, mma_fused_launch(H, nullptr, nullptr, 0, 0, nullptr, 0, 0, wnode, 0, nk, n, nw, z, sigma, ctx->partial.as<double2>(), ef,
                                         ncta, kper_m, ctx->stream));
                ctx->launches++;
                reduce_partials_kernel<<<nw, 256, 0, ctx->stream>>>(ctx->partial.as<double2>(), ncta, nw, 1.0, ctx->acc.as<double2>());
                LAUNCH_CHECK(ctx, "reduce_partials_kernel");
            } else if (mode == 0) {
                CU(ctx, ctx->partial.reserve((size_t)ncta * nw * sizeof(double2)));
                CU(ctx, mma_resolvent_launch(H, wnode, nk, n, nw, z, sigma, 0, ctx->partial.as<double2>(), ef, ncta, kper_m, ctx->stream));
                ctx->launches++;
                reduce_partials_kernel<<<nw, 256, 0, ctx->stream>>>(ctx->partial.as<double2>(), ncta, nw, 1.0, ctx->acc.as<double2>());
                LAUNCH_CHECK(ctx, "reduce_partials_kernel");
            } else {
                CU(ctx, mma_resolvent_launch(H, wnode, nk, n, nw, z, sigma, 1, yout, ef, ncta, kper_m, ctx->stream));
                ctx->launches++;
            }
            return ABZ_OK;
        }
    }
    if (n > 64) return fail(ctx, ABZ_E_UNSUPPORTED, "norb > 64 is not supported by the resolvent kernels");
    // DMMA block LU with the matrix shared by a team of 4 warps (32 < norb <= 64; unpivoted between blocks, growth-monitored like the
    // one-warp kernel: the same flag makes the host rerun the call with the pivoted teams below).  ABZ_MMA_TEAM=1 also routes
    // 24 < norb <= 32 through two-warp teams (measurement hook).
    static const int team_env = getenv("ABZ_MMA_TEAM") ? atoi(getenv("ABZ_MMA_TEAM")) : 0;
    const bool team_small = team_env == 1 && n > 24 && n <= 32;
    if ((ctx->resolvent_algo == 0 || ctx->resolvent_algo == 2) && !ctx->force_generic && (mma_team_supported(n) || team_small) && team_env != -1) {
        long ncta = 0; int kper_t = 1;
        const size_t need = mma_team_smem((n + 7) / 8, nw);
        if (need <= 110 * 1024) {
            const long per_sm = team_small ? 8 : 2;
            const long target = sm * per_sm;
            kper_t = (int)std::max<long>(1, (nk + target - 1) / target);
            ncta = (nk + kper_t - 1) / kper_t;
            double2* dst = yout;
            if (mode == 0) {
                CU(ctx, ctx->partial.reserve((size_t)ncta * nw * sizeof(double2)));
                dst = ctx->partial.as<double2>();
            }
            CU(ctx, mma_team_launch(H, wnode, nk, n, nw, z, sigma, mode, dst, ef, ncta, kper_t, ctx->stream));
            ctx->launches++;
            if (mode == 0) {
                reduce_partials_kernel<<<nw, 256, 0, ctx->stream>>>(ctx->partial.as<double2>(), ncta, nw, 1.0, ctx->acc.as<double2>());
                LAUNCH_CHECK(ctx, "reduce_partials_kernel");
            }
            return ABZ_OK;
        }
    }
    if (ctx->resolvent_algo != 4 && !(n <= 16 && ctx->resolvent_algo != 1)) {   // (n <= 16: a padded 32-row team wastes its lanes; measured crossover)
        // register-resident pivoted Gauss-Jordan: one team (1 or 4 warps) per matrix, grid = (node chunks, frequency chunks)
        const long target = (long)sm * gj_ctas_per_sm(n, false);   // one resident wave
        const long chunks_k = std::min<long>(nk, target);
        const int kper = (int)((nk + chunks_k - 1) / chunks_k);
        const long ncx = (nk + kper - 1) / kper;
        const long chunks_w = std::min<long>(nw, std::max<long>(1, target / ncx));
        const int wper = (int)std::min<long>(64, (nw + chunks_w - 1) / chunks_w);
        dim3 grid((unsigned)ncx, (unsigned)((nw + wper - 1) / wper));
        double2* dst = yout;
        if (mode == 0) {
            CU(ctx, ctx->partial.reserve((size_t)ncx * nw * sizeof(double2)));
            dst = ctx->partial.as<double2>();
        }
        GJ_DISPATCH(resolvent_gjreg_kernel, false, n, grid, 0, ctx->stream, H, wnode, nk, n, nw, z, sigma, kper, wper, mode, dst, ef);
        LAUNCH_CHECK(ctx, "resolvent_gjreg_kernel");
        if (mode == 0) {
            reduce_partials_kernel<<<nw, 256, 0, ctx->stream>>>(ctx->partial.as<double2>(), ncx, nw, 1.0, ctx->acc.as<double2>());
            LAUNCH_CHECK(ctx, "reduce_partials_kernel");
        }
        return ABZ_OK;
    }
    int nwarps = 8;
    while (nwarps > 1 && gj_smem_bytes(n, nw, nwarps) > 200 * 1024) nwarps--;
    size_t smem = gj_smem_bytes(n, nw, nwarps);
    if (smem > 220 * 1024) return fail(ctx, ABZ_E_UNSUPPORTED, "too many frequencies per call for the generic resolvent kernel");
    // nodes per CTA: aim at ~4 CTAs per SM worth of CTAs, but keep every warp busy: work per node = nw matrices
    long target_cta = sm * 4;
    int kper = (int)std::max<long>(1, (nk + target_cta - 1) / target_cta);
    long ncta = (nk + kper - 1) / kper;
    if (mode == 0) {
        CU(ctx, ctx->partial.reserve((size_t)ncta * nw * sizeof(double2)));
        resolvent_gj_kernel<<<(unsigned)ncta, nwarps * 32, smem, ctx->stream>>>(H, wnode, nk, n, nw, z, sigma, kper, 0,
                                                                               ctx->partial.as<double2>(), ef);
        LAUNCH_CHECK(ctx, "resolvent_gj_kernel");
        reduce_partials_kernel<<<nw, 256, 0, ctx->stream>>>(ctx->partial.as<double2>(), ncta, nw, 1.0, ctx->acc.as<double2>());
        LAUNCH_CHECK(ctx, "reduce_partials_kernel");
    } else {
        resolvent_gj_kernel<<<(unsigned)ncta, nwarps * 32, smem, ctx->stream>>>(H, wnode, nk, n, nw, z, sigma, kper, 1, yout, ef);
        LAUNCH_CHECK(ctx, "resolvent_gj_kernel");
    }
    return ABZ_OK;
}

template <bool FROM_H>
int run_small_fused(abz_ctx* ctx, Rule* r, const double2* C1, const double2* Hmat, long r0, long nrows, int fkind,
                    int nw, const double2* z, const double2* sigma) {
    Series* s = r->s;
    long nnodes = r->h_row_nodeptr[r0 + nrows] - r->h_row_nodeptr[r0];
    if (nnodes <= 0) return ABZ_OK;
    const long per_cta = (long)SM_THREADS * SM_NB;
    long ncta = std::min<long>((nnodes + per_cta - 1) / per_cta, (long)ctx->sm_count * 8);
    if (fkind == ABZ_F_TRACE_H) nw = 1;
    int nwy = (nw + SM_WMAX - 1) / SM_WMAX;
    const size_t smem = (size_t)9 * std::min(nw, SM_WMAX) * sizeof(double2);
    CU(ctx, ctx->partial.reserve((size_t)ncta * nw * sizeof(double2)));
    dim3 grid((unsigned)ncta, (unsigned)nwy);
    int* ef = ctx->errflag.as<int>();
    double2* part = ctx->partial.as<double2>();
#define SMALL_LAUNCH(NORB)                                                                                         \
    small_fused_kernel<NORB, FROM_H><<<grid, SM_THREADS, smem, ctx->stream>>>(C1, Hmat, r->d_ptab[0], r->d_row_nodeptr, r0, nrows, \
                                                                             r->d_node_k1, r->d_node_w, r->N, s->M[0], fkind, nw, z, \
                                                                             sigma, part, ef)
    if (s->n == 1) SMALL_LAUNCH(1);
    else if (s->n == 2) SMALL_LAUNCH(2);
    else SMALL_LAUNCH(3);
#undef SMALL_LAUNCH
    LAUNCH_CHECK(ctx, "small_fused_kernel");
    reduce_partials_kernel<<<nw, 256, 0, ctx->stream>>>(part, ncta, nw, 1.0, ctx->acc.as<double2>());
    LAUNCH_CHECK(ctx, "reduce_partials_kernel");
    return ABZ_OK;
}

void collect_timings(abz_ctx* ctx, const std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& ev_eval,
                     const std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& ev_mat) {
    ctx->eval_ms = 0; ctx->matfun_ms = 0;
    for (auto& p : ev_eval) { float ms = 0; cudaEventElapsedTime(&ms, p.first, p.second); ctx->eval_ms += ms; }
    for (auto& p : ev_mat) { float ms = 0; cudaEventElapsedTime(&ms, p.first, p.second); ctx->matfun_ms += ms; }
    ctx->ev_used = 0;
}

// Opt every kernel that needs more than 48 KB of dynamic shared memory in on the CURRENT device.  The attribute is per device
// (and per kernel), so it is set whenever a context is created - one process may hold one context per GPU.
cudaError_t opt_in_dynamic_smem() {
    cudaError_t e = cudaSuccess;
    auto set = [&](const void* f, int bytes) {
        cudaError_t r = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (r != cudaSuccess && e == cudaSuccess) e = r;
    };
#define ABZ_OPT_IN(kernel, bytes) set(reinterpret_cast<const void*>(&kernel), bytes)
    ABZ_OPT_IN(contract_stage_kernel, 200 * 1024);
    ABZ_OPT_IN(eig_tridiag_kernel<32>, 100 * 1024);
    ABZ_OPT_IN(eig_tridiag_kernel<64>, 100 * 1024);
    ABZ_OPT_IN(resolvent_gj_kernel, 220 * 1024);
    ABZ_OPT_IN(resolvent_gj_matrix_kernel, 220 * 1024);
    { auto k = small_fused_kernel<1, true>; set((const void*)k, 160 * 1024); }
    { auto k = small_fused_kernel<2, true>; set((const void*)k, 160 * 1024); }
    { auto k = small_fused_kernel<3, true>; set((const void*)k, 160 * 1024); }
    { auto k = small_fused_kernel<1, false>; set((const void*)k, 160 * 1024); }
    { auto k = small_fused_kernel<2, false>; set((const void*)k, 160 * 1024); }
    { auto k = small_fused_kernel<3, false>; set((const void*)k, 160 * 1024); }
    { auto k = resolvent_gjreg_matrix_kernel<64, 32, 4, 3>; set((const void*)k, 64 * 1024); }
    { auto k = resolvent_gjreg_matrix_kernel<64, 32, 16, 2>; set((const void*)k, 64 * 1024); }
    { auto k = resolvent_gjreg_matrix_kernel<64, 16, 8, 2>; set((const void*)k, 64 * 1024); }
    ABZ_OPT_IN(eig_jacobi_kernel, 200 * 1024);
    ABZ_OPT_IN(eig_jacobi_vel_kernel, 200 * 1024);
    ABZ_OPT_IN(nest_panel_small_kernel<4>, 160 * 1024); ABZ_OPT_IN(nest_panel_small_kernel<5>, 160 * 1024); ABZ_OPT_IN(nest_panel_small_kernel<6>, 160 * 1024);
    ABZ_OPT_IN(iai_leaf_kernel<4>, 160 * 1024); ABZ_OPT_IN(iai_leaf_kernel<5>, 160 * 1024); ABZ_OPT_IN(iai_leaf_kernel<6>, 160 * 1024);
    ABZ_OPT_IN(iai_mid_kernel<4>, 200 * 1024); ABZ_OPT_IN(iai_mid_kernel<5>, 200 * 1024); ABZ_OPT_IN(iai_mid_kernel<6>, 200 * 1024);
    ABZ_OPT_IN(nest_panel_small_kernel<1>, 160 * 1024);
    ABZ_OPT_IN(nest_panel_small_kernel<2>, 160 * 1024);
    ABZ_OPT_IN(nest_panel_small_kernel<3>, 160 * 1024);
    ABZ_OPT_IN(iai_leaf_kernel<1>, 160 * 1024);
    ABZ_OPT_IN(iai_leaf_kernel<2>, 160 * 1024);
    ABZ_OPT_IN(iai_leaf_kernel<3>, 160 * 1024);
    ABZ_OPT_IN(iai_mid_kernel<1>, 200 * 1024);
    ABZ_OPT_IN(iai_mid_kernel<2>, 200 * 1024);
    ABZ_OPT_IN(iai_mid_kernel<3>, 200 * 1024);
#undef ABZ_OPT_IN
    cudaError_t r = mma_resolvent_opt_in();
    if (r != cudaSuccess && e == cudaSuccess) e = r;
    r = mma_team_opt_in();
    if (r != cudaSuccess && e == cudaSuccess) e = r;
    return e;
}

long node_cap_for(abz_ctx* ctx, int n) { return (long)std::max<size_t>(1, ctx->budget / ((size_t)n * n * sizeof(double2))); }

}  // namespace

// =================================================================================================
extern "C" {

int32_t abz_version(void) { return 100; }

const char* abz_last_error(const abz_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int32_t abz_ctx_create(int32_t device, abz_ctx** out) {
    if (!out) return fail(nullptr, ABZ_E_INVALID, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, ABZ_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev) return fail(nullptr, ABZ_E_INVALID, "device index out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, ABZ_E_CUDA, cudaGetErrorString(e));
    if (prop.major != 10) return fail(nullptr, ABZ_E_UNSUPPORTED, "libautobz_cuda is built for sm_100a (B200) only");
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, ABZ_E_CUDA, cudaGetErrorString(e));
    if ((e = opt_in_dynamic_smem()) != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, ABZ_E_CUDA, std::string("shared-memory opt-in: ") + cudaGetErrorString(e));
    }
    abz_ctx* ctx = new abz_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, ABZ_E_CUDA, cudaGetErrorString(e));
    }
    if (ctx->errflag.reserve(sizeof(int)) != cudaSuccess) { delete ctx; return fail(nullptr, ABZ_E_OOM, "errflag alloc"); }
    cudaMemsetAsync(ctx->errflag.p, 0, sizeof(int), ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    *out = ctx;
    return ABZ_OK;
}

int32_t abz_ctx_destroy(abz_ctx* ctx) {
    if (!ctx) return ABZ_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    abz_comm_destroy(ctx);
    ctx->rules.clear(); ctx->nests.clear(); ctx->series.clear();
    ctx->iai_lanes.clear();
    pool_release(ctx);
    if (ctx->pin_in) cudaFreeHost(ctx->pin_in);
    if (ctx->pin_out) cudaFreeHost(ctx->pin_out);
    for (auto e : ctx->events) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ABZ_OK;
}

int32_t abz_ctx_set_option(abz_ctx* ctx, int32_t option, int64_t value) {
    if (!ctx) return ABZ_E_INVALID;
    switch (option) {
        case ABZ_OPT_RESOLVENT_ALGO: ctx->resolvent_algo = (int)value; return ABZ_OK;
        case ABZ_OPT_MEM_BUDGET_MB: if (value < 1) return fail(ctx, ABZ_E_INVALID, "budget must be >= 1 MB");
            ctx->budget = (size_t)value << 20; return ABZ_OK;
        case ABZ_OPT_FUSED_SMALL: ctx->fused_small = (int)value; return ABZ_OK;
        case ABZ_OPT_EIG_ALGO: ctx->eig_algo = (int)value; return ABZ_OK;
        case ABZ_OPT_IAI_LANES: if (value < 1 || value > 16) return fail(ctx, ABZ_E_INVALID, "lanes must be in 1..16");
            ctx->iai_lanes_opt = (int)value; return ABZ_OK;
        case ABZ_OPT_IAI_LEAF_SPILL: if (value == 0 || value < -63 || value > (1 << 20)) return fail(ctx, ABZ_E_INVALID, "spill capacity out of range");
            ctx->leaf_spill = (int)value; return ABZ_OK;
    }
    return fail(ctx, ABZ_E_INVALID, "unknown option");
}

int64_t abz_ctx_launch_count(const abz_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t abz_ctx_last_timings(const abz_ctx* ctx, double* eval_ms, double* matfun_ms) {
    if (!ctx) return ABZ_E_INVALID;
    if (eval_ms) *eval_ms = ctx->eval_ms;
    if (matfun_ms) *matfun_ms = ctx->matfun_ms;
    return ABZ_OK;
}

int32_t abz_series_create(abz_ctx* ctx, const double* coeffs, int32_t is_complex, int32_t norb, const int32_t M[3],
                          const int32_t lo[3], const double period[3], abz_series_t* out) {
    if (!ctx) return ABZ_E_INVALID;
    if (!coeffs || !M || !lo || !period || !out) return fail(ctx, ABZ_E_INVALID, "NULL argument");
    if (norb < 1 || norb > 64) return fail(ctx, ABZ_E_UNSUPPORTED, "norb must be in 1..64");
    for (int d = 0; d < 3; d++)
        if (M[d] < 1 || !(period[d] > 0)) return fail(ctx, ABZ_E_INVALID, "M >= 1 and period > 0 required");
    cudaSetDevice(ctx->device);
    auto s = std::make_shared<Series>();
    s->ctx = ctx;
    s->n = norb;
    size_t cnt = (size_t)norb * norb;
    for (int d = 0; d < 3; d++) { s->M[d] = M[d]; s->lo[d] = lo[d]; s->period[d] = period[d]; cnt *= M[d]; }
    CU(ctx, pool_alloc(ctx, (void**)&s->c, cnt * sizeof(double2)));
    if (is_complex) {
        CU(ctx, cudaMemcpyAsync(s->c, coeffs, cnt * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        std::vector<double2> tmp(cnt);
        for (size_t i = 0; i < cnt; i++) tmp[i] = make_double2(coeffs[i], 0.0);
        CU(ctx, cudaMemcpyAsync(s->c, tmp.data(), cnt * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    uint64_t id = ctx->next_id++;
    ctx->series[id] = std::move(s);
    *out = id;
    return ABZ_OK;
}

int32_t abz_series_destroy(abz_ctx* ctx, abz_series_t s) {
    if (!ctx) return ABZ_E_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    return ctx->series.erase(s) ? ABZ_OK : fail(ctx, ABZ_E_INVALID, "unknown series handle");
}

int32_t abz_rule_create_full(abz_ctx* ctx, abz_series_t sid, int32_t npt, int32_t k3_lo, int32_t k3_hi, abz_rule_t* out) {
    if (!ctx) return ABZ_E_INVALID;
    Series* s = get_series(ctx, sid);
    if (!s) return fail(ctx, ABZ_E_INVALID, "unknown series handle");
    if (!out || npt < 1 || k3_lo < 0 || k3_hi > npt || k3_lo > k3_hi) return fail(ctx, ABZ_E_INVALID, "invalid grid range");
    cudaSetDevice(ctx->device);
    auto r = std::make_unique<Rule>();
    r->ctx = ctx;
    r->series_id = sid; r->s = s; r->keep = ctx->series[sid]; r->N = npt; r->full = true;
    const long N = npt;
    r->np3 = k3_hi - k3_lo;
    r->nrows = r->np3 * N;
    r->nnz = r->nrows * N;
    r->h_plane_k3.resize(r->np3); r->h_plane_rowptr.resize(r->np3 + 1);
    r->h_row_k2.resize(r->nrows); r->h_row_nodeptr.resize(r->nrows + 1);
    for (long p = 0; p < r->np3; p++) { r->h_plane_k3[p] = (int)(k3_lo + p); r->h_plane_rowptr[p] = p * N; }
    r->h_plane_rowptr[r->np3] = r->nrows;
    for (long q = 0; q < r->nrows; q++) { r->h_row_k2[q] = (int)(q % N); r->h_row_nodeptr[q] = q * N; }
    r->h_row_nodeptr[r->nrows] = r->nnz;
    int rc = finish_rule(ctx, r.get());
    if (rc) return rc;
    uint64_t id = ctx->next_id++;
    ctx->rules[id] = std::move(r);
    *out = id;
    return ABZ_OK;
}

int32_t abz_rule_create_sym(abz_ctx* ctx, abz_series_t sid, int32_t npt, const int32_t* wsym, int32_t k3_lo,
                            int32_t k3_stride, abz_rule_t* out) {
    if (!ctx) return ABZ_E_INVALID;
    Series* s = get_series(ctx, sid);
    if (!s) return fail(ctx, ABZ_E_INVALID, "unknown series handle");
    if (!out || !wsym || npt < 1 || !share_valid(k3_lo, k3_stride)) return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    cudaSetDevice(ctx->device);
    auto r = std::make_unique<Rule>();
    r->ctx = ctx;
    r->series_id = sid; r->s = s; r->keep = ctx->series[sid]; r->N = npt; r->full = false;
    const long N = npt;
    r->h_plane_rowptr.push_back(0);
    r->h_row_nodeptr.push_back(0);
    for (int psel = 0; share_plane(psel, k3_lo, k3_stride) < N; psel++) {
        const long i3 = share_plane(psel, k3_lo, k3_stride);
        bool plane_open = false;
        for (long i2 = 0; i2 < N; i2++) {
            const int32_t* row = wsym + (i3 * N + i2) * N;
            bool row_open = false;
            for (long i1 = 0; i1 < N; i1++) {
                if (!row[i1]) continue;
                if (row[i1] < 0) return fail(ctx, ABZ_E_INVALID, "negative symmetry weight");
                row_open = true;
                r->h_node_k1.push_back((int)i1);
                r->h_node_w.push_back((double)row[i1]);
            }
            if (row_open) {
                plane_open = true;
                r->h_row_k2.push_back((int)i2);
                r->h_row_nodeptr.push_back((long)r->h_node_k1.size());
            }
        }
        if (plane_open) {
            r->h_plane_k3.push_back((int)i3);
            r->h_plane_rowptr.push_back((long)r->h_row_k2.size());
        }
    }
    r->np3 = (long)r->h_plane_k3.size();
    r->nrows = (long)r->h_row_k2.size();
    r->nnz = (long)r->h_node_k1.size();
    int rc = finish_rule(ctx, r.get());
    if (rc) return rc;
    uint64_t id = ctx->next_id++;
    ctx->rules[id] = std::move(r);
    *out = id;
    return ABZ_OK;
}

int32_t abz_rule_create_nodes(abz_ctx* ctx, abz_series_t sid, int32_t npt, int64_t nnodes, const int32_t* idx,
                              const double* w, abz_rule_t* out) {
    if (!ctx) return ABZ_E_INVALID;
    Series* s = get_series(ctx, sid);
    if (!s) return fail(ctx, ABZ_E_INVALID, "unknown series handle");
    if (!out || npt < 1 || nnodes < 0 || (nnodes > 0 && !idx)) return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    cudaSetDevice(ctx->device);
    auto r = std::make_unique<Rule>();
    r->ctx = ctx;
    r->series_id = sid; r->s = s; r->keep = ctx->series[sid]; r->N = npt; r->full = false;
    r->h_plane_rowptr.push_back(0);
    r->h_row_nodeptr.push_back(0);
    long p3 = -1, p2 = -1, p1 = -1;
    for (int64_t i = 0; i < nnodes; i++) {
        long i1 = idx[3 * i], i2 = idx[3 * i + 1], i3 = idx[3 * i + 2];
        if (i1 < 0 || i1 >= npt || i2 < 0 || i2 >= npt || i3 < 0 || i3 >= npt) return fail(ctx, ABZ_E_INVALID, "node index out of range");
        bool newplane = (i3 != p3), newrow = newplane || (i2 != p2);
        if (i3 < p3 || (!newplane && i2 < p2) || (!newrow && i1 <= p1))
            return fail(ctx, ABZ_E_INVALID, "nodes must be sorted by (k3, k2, k1) without duplicates");
        if (newrow && i > 0) r->h_row_nodeptr.push_back((long)r->h_node_k1.size());
        if (newplane && i > 0) r->h_plane_rowptr.push_back((long)r->h_row_k2.size());
        if (newplane) r->h_plane_k3.push_back((int)i3);
        if (newrow) r->h_row_k2.push_back((int)i2);
        r->h_node_k1.push_back((int)i1);
        r->h_node_w.push_back(w ? w[i] : 1.0);
        p3 = i3; p2 = i2; p1 = i1;
    }
    if (nnodes > 0) {
        r->h_row_nodeptr.push_back((long)r->h_node_k1.size());
        r->h_plane_rowptr.push_back((long)r->h_row_k2.size());
    }
    r->np3 = (long)r->h_plane_k3.size();
    r->nrows = (long)r->h_row_k2.size();
    r->nnz = (long)r->h_node_k1.size();
    int rc = finish_rule(ctx, r.get());
    if (rc) return rc;
    uint64_t id = ctx->next_id++;
    ctx->rules[id] = std::move(r);
    *out = id;
    return ABZ_OK;
}

// symptr_rule_kernel with the fast modular reduction whenever every intermediate |S i| stays below 2^22
// The symmetry list must be a group (identity in it, no duplicates, closed under products), as every list from load_bz is:
// AutoSymPTR.symptr_rule's sequential scan is order-dependent otherwise, and the device version relies on
// "representative = smallest linear index of the orbit" and weight = nsyms / |stabiliser|.
static bool syms_form_group(const int32_t* h_syms, int nsyms) {
    auto eq = [&](const int32_t* A, const int32_t* B) { for (int t = 0; t < 9; t++) if (A[t] != B[t]) return false; return true; };
    const int32_t I9[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    bool has_id = false;
    for (int a = 0; a < nsyms; a++) {
        has_id |= eq(h_syms + 9 * a, I9);
        for (int b = 0; b < a; b++) if (eq(h_syms + 9 * a, h_syms + 9 * b)) return false;
    }
    if (!has_id) return false;
    for (int a = 0; a < nsyms; a++)
        for (int b = 0; b < nsyms; b++) {
            int32_t P[9];
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++)
                    P[3 * i + j] = h_syms[9 * a + 3 * i] * h_syms[9 * b + j] + h_syms[9 * a + 3 * i + 1] * h_syms[9 * b + 3 + j] +
                                   h_syms[9 * a + 3 * i + 2] * h_syms[9 * b + 6 + j];
            bool found = false;
            for (int c = 0; c < nsyms && !found; c++) found = eq(P, h_syms + 9 * c);
            if (!found) return false;
        }
    return true;
}

// Orbit weights of the planes i3 = share_plane(p, k3_lo, k3_stride) (p < nplanes) into the dense array d_w; other planes are not touched.
// Every point's test "am I the smallest index of my orbit" is independent of the others, so a rank needs only its own planes.
static int launch_symptr(abz_ctx* ctx, int npt, int nsyms, const int32_t* h_syms, const int* d_syms, int* d_w, int k3_lo = 0,
                         int k3_stride = 1, long nplanes = -1) {
    if (!syms_form_group(h_syms, nsyms))
        return fail(ctx, ABZ_E_INVALID, "the symmetries must form a group (identity included, closed under products, no duplicates)");
    if (nplanes < 0) nplanes = npt;
    if (nplanes == 0) return ABZ_OK;
    long smax = 0;
    for (int t = 0; t < 9 * nsyms; t++) smax = std::max<long>(smax, std::labs((long)h_syms[t]));
    const size_t plane = (size_t)npt * npt;
    const size_t tot = (size_t)npt * npt * npt;
    const size_t sel = plane * (size_t)nplanes;
    const unsigned grid = (unsigned)((sel + 255) / 256);
    const size_t smem = (size_t)nsyms * 9 * sizeof(int);
    const bool fast = 3 * smax * npt < (1L << 22);
    // large grids, at most 64 symmetries: phases with compaction (see abz_iai.cuh); else the one-kernel version
    if (fast && nsyms <= 64 && nsyms > 8 && sel >= ((size_t)1 << 22) && tot < ((size_t)1 << 32)) {
        const unsigned cap1 = (unsigned)(sel / 2 + 1024), cap2 = 16u;
        DevBuf& lb = ctx->symlist;
        CU(ctx, lb.reserve(((size_t)cap1 + cap2 + 16) * sizeof(unsigned)));
        unsigned* cnt = lb.as<unsigned>();           // [0]: survivors of phase 1, [1]: irreducible points, [2]: overflow flag
        unsigned* l1 = cnt + 16;
        unsigned* l2 = l1 + cap1;
        CU(ctx, cudaMemsetAsync(cnt, 0, 16 * sizeof(unsigned), ctx->stream));
        if (k3_stride == 1) CU(ctx, cudaMemsetAsync(d_w + (size_t)k3_lo * plane, 0, sel * sizeof(int), ctx->stream));
        else if (k3_stride > 0) CU(ctx, cudaMemset2DAsync(d_w + (size_t)k3_lo * plane, (size_t)k3_stride * plane * sizeof(int), 0, plane * sizeof(int),
                                                         (size_t)nplanes, ctx->stream));
        else {      // serpentine share: the even and the odd planes of the sequence are two strided sets
            const size_t pitch = (size_t)(-2 * k3_stride) * plane * sizeof(int);
            CU(ctx, cudaMemset2DAsync(d_w + (size_t)share_plane(0, k3_lo, k3_stride) * plane, pitch, 0, plane * sizeof(int), (size_t)((nplanes + 1) / 2), ctx->stream));
            if (nplanes > 1)
                CU(ctx, cudaMemset2DAsync(d_w + (size_t)share_plane(1, k3_lo, k3_stride) * plane, pitch, 0, plane * sizeof(int), (size_t)(nplanes / 2), ctx->stream));
        }
        dim3 g1((unsigned)((plane + 255) / 256), (unsigned)nplanes);
        symptr_filter_kernel<true, false><<<g1, 256, smem, ctx->stream>>>(npt, nsyms, 0, 8, d_syms, nullptr, nullptr, l1, cnt, cap1,
                                                                         reinterpret_cast<int*>(cnt + 2), d_w, k3_lo, k3_stride);
        LAUNCH_CHECK(ctx, "symptr_filter_kernel");
        unsigned h[3] = {0, 0, 0};
        CU(ctx, cudaMemcpyAsync(h, cnt, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (!h[2] && h[0] > 0) {
            symptr_filter_kernel<true, true><<<(h[0] + 255) / 256, 256, smem, ctx->stream>>>(npt, nsyms, 8, nsyms, d_syms, l1, cnt, l2, cnt + 1,
                                                                                            cap2, reinterpret_cast<int*>(cnt + 2), d_w, 0, 1);
            LAUNCH_CHECK(ctx, "symptr_filter_kernel");
        }
        if (!h[2]) return ABZ_OK;
        // a list overflowed (symmetry list with an unusual order): fall through to the one-kernel version, which overwrites d_w
    }
    if (fast) symptr_rule_kernel<true><<<grid, 256, smem, ctx->stream>>>(npt, nsyms, d_syms, d_w, k3_lo, k3_stride, (int)nplanes);
    else symptr_rule_kernel<false><<<grid, 256, smem, ctx->stream>>>(npt, nsyms, d_syms, d_w, k3_lo, k3_stride, (int)nplanes);
    LAUNCH_CHECK(ctx, "symptr_rule_kernel");
    return ABZ_OK;
}

int32_t abz_symptr_rule(abz_ctx* ctx, int32_t npt, int32_t nsyms, const int32_t* syms, int32_t* wsym_out, int64_t* nirr) {
    if (!ctx) return ABZ_E_INVALID;
    if (npt < 1 || nsyms < 1 || nsyms > 1024 || !syms || !wsym_out) return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    for (int t = 0; t < 9 * nsyms; t++)
        if ((long)std::abs((long)syms[t]) * 3 * npt >= ((long)1 << 31)) return fail(ctx, ABZ_E_INVALID, "symmetry matrix entries too large");
    cudaSetDevice(ctx->device);
    size_t tot = (size_t)npt * npt * npt;
    CU(ctx, ctx->tmp_a.reserve(tot * sizeof(int)));
    CU(ctx, ctx->tmp_b.reserve((size_t)nsyms * 9 * sizeof(int)));
    CU(ctx, cudaMemcpyAsync(ctx->tmp_b.p, syms, (size_t)nsyms * 9 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    { int rcs = launch_symptr(ctx, npt, nsyms, syms, ctx->tmp_b.as<int>(), ctx->tmp_a.as<int>()); if (rcs) return rcs; }
    CU(ctx, cudaMemcpyAsync(wsym_out, ctx->tmp_a.p, tot * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (nirr) {
        int64_t c = 0;
        for (size_t i = 0; i < tot; i++) c += wsym_out[i] != 0;
        *nirr = c;
    }
    return ABZ_OK;
}

int32_t abz_rule_create_symptr(abz_ctx* ctx, abz_series_t sid, int32_t npt, int32_t nsyms, const int32_t* syms, int32_t k3_lo,
                               int32_t k3_stride, abz_rule_t* out, int64_t* nirr_total) {
    if (!ctx) return ABZ_E_INVALID;
    Series* s = get_series(ctx, sid);
    if (!s) return fail(ctx, ABZ_E_INVALID, "unknown series handle");
    if (!out || npt < 1 || nsyms < 1 || nsyms > 1024 || !syms || !share_valid(k3_lo, k3_stride))
        return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    for (int t = 0; t < 9 * nsyms; t++)
        if ((long)std::abs((long)syms[t]) * 3 * npt >= ((long)1 << 31)) return fail(ctx, ABZ_E_INVALID, "symmetry matrix entries too large");
    cudaSetDevice(ctx->device);
    const long N = npt;
    const size_t tot = (size_t)N * N * N;
    // dense orbit weights on the device (never copied to the host)
    DevBuf& wbuf = ctx->symw;      // dense orbit weights: kept across rules (AutoPTR builds one per refinement)
    DevBuf cntbuf, k3buf;
    CU(ctx, wbuf.reserve(tot * sizeof(int)));
    CU(ctx, ctx->tmp_b.reserve((size_t)nsyms * 9 * sizeof(int)));
    CU(ctx, cudaMemcpyAsync(ctx->tmp_b.p, syms, (size_t)nsyms * 9 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    // per-row counts: all planes when the total is wanted, else only this rank's (then the orbit weights of the other ranks'
    // planes are not even computed: the caller sums the local node counts over the ranks)
    const bool want_total = (nirr_total != nullptr) && !(k3_lo == 0 && k3_stride == 1);
    std::vector<int> cnt_all;
    const long nplanes_sel = share_nplanes(N, k3_lo, k3_stride);
    {
        int rcs = want_total ? launch_symptr(ctx, npt, nsyms, syms, ctx->tmp_b.as<int>(), wbuf.as<int>())
                             : launch_symptr(ctx, npt, nsyms, syms, ctx->tmp_b.as<int>(), wbuf.as<int>(), k3_lo, k3_stride, nplanes_sel);
        if (rcs) return rcs;
    }
    const long rows_sel = nplanes_sel * N;
    std::vector<int> cnt(rows_sel);
    CU(ctx, cntbuf.reserve((size_t)std::max<long>(N * N, 1) * sizeof(int)));
    if (want_total) {
        sym_row_count_kernel<<<(unsigned)((N * N * 32 + 255) / 256), 256, 0, ctx->stream>>>(wbuf.as<int>(), npt, 0, 1, N * N, cntbuf.as<int>());
        LAUNCH_CHECK(ctx, "sym_row_count_kernel");
        cnt_all.resize(N * N);
        CU(ctx, cudaMemcpyAsync(cnt_all.data(), cntbuf.p, (size_t)N * N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        for (long p = 0; p < nplanes_sel; p++)
            memcpy(cnt.data() + p * N, cnt_all.data() + (long)share_plane((int)p, k3_lo, k3_stride) * N, (size_t)N * sizeof(int));
    } else if (rows_sel > 0) {
        sym_row_count_kernel<<<(unsigned)((rows_sel * 32 + 255) / 256), 256, 0, ctx->stream>>>(wbuf.as<int>(), npt, k3_lo, k3_stride,
                                                                                             rows_sel, cntbuf.as<int>());
        LAUNCH_CHECK(ctx, "sym_row_count_kernel");
        CU(ctx, cudaMemcpyAsync(cnt.data(), cntbuf.p, (size_t)rows_sel * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (nirr_total) {
        int64_t t = 0;
        if (want_total) for (int c : cnt_all) t += c; else for (int c : cnt) t += c;
        *nirr_total = t;
    }
    // CSR skeleton on the host (N^2 entries), node arrays on the device
    auto r = std::make_unique<Rule>();
    r->ctx = ctx;
    r->series_id = sid; r->s = s; r->keep = ctx->series[sid]; r->N = npt; r->full = false; r->nodes_on_host = false;
    r->h_plane_rowptr.push_back(0);
    r->h_row_nodeptr.push_back(0);
    std::vector<int> row_k3;
    long nnz = 0;
    for (long p = 0; p < nplanes_sel; p++) {
        bool open = false;
        for (long i2 = 0; i2 < N; i2++) {
            const int c = cnt[p * N + i2];
            if (!c) continue;
            open = true;
            nnz += c;
            r->h_row_k2.push_back((int)i2);
            row_k3.push_back(share_plane((int)p, k3_lo, k3_stride));
            r->h_row_nodeptr.push_back(nnz);
        }
        if (open) {
            r->h_plane_k3.push_back(share_plane((int)p, k3_lo, k3_stride));
            r->h_plane_rowptr.push_back((long)r->h_row_k2.size());
        }
    }
    r->np3 = (long)r->h_plane_k3.size();
    r->nrows = (long)r->h_row_k2.size();
    r->nnz = nnz;
    int rc = finish_rule(ctx, r.get());
    if (rc) return rc;
    if (nnz > 0) {
        CU(ctx, pool_alloc(ctx, (void**)&r->d_node_k1, (size_t)nnz * sizeof(int)));
        CU(ctx, pool_alloc(ctx, (void**)&r->d_node_w, (size_t)nnz * sizeof(double)));
        CU(ctx, k3buf.reserve(row_k3.size() * sizeof(int)));
        CU(ctx, cudaMemcpyAsync(k3buf.p, row_k3.data(), row_k3.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        sym_row_fill_kernel<<<(unsigned)((r->nrows * 32 + 255) / 256), 256, 0, ctx->stream>>>(
            wbuf.as<int>(), npt, k3buf.as<int>(), r->d_row_k2, r->d_row_nodeptr, r->nrows, r->d_node_k1, r->d_node_w);
        LAUNCH_CHECK(ctx, "sym_row_fill_kernel");
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    uint64_t id = ctx->next_id++;
    ctx->rules[id] = std::move(r);
    *out = id;
    return ABZ_OK;
}

int32_t abz_rule_destroy(abz_ctx* ctx, abz_rule_t r) {
    if (!ctx) return ABZ_E_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    return ctx->rules.erase(r) ? ABZ_OK : fail(ctx, ABZ_E_INVALID, "unknown rule handle");
}

int32_t abz_rule_info(abz_ctx* ctx, abz_rule_t rid, int64_t* nnodes, int32_t* norb, int32_t* npt) {
    if (!ctx) return ABZ_E_INVALID;
    Rule* r = get_rule(ctx, rid);
    if (!r) return fail(ctx, ABZ_E_INVALID, "unknown rule handle");
    if (nnodes) *nnodes = r->nnz;
    if (norb) *norb = r->s->n;
    if (npt) *npt = r->N;
    return ABZ_OK;
}

int32_t abz_rule_materialize(abz_ctx* ctx, abz_rule_t rid) {
    if (!ctx) return ABZ_E_INVALID;
    Rule* r = get_rule(ctx, rid);
    if (!r) return fail(ctx, ABZ_E_INVALID, "unknown rule handle");
    if (r->d_H || r->nnz == 0) return ABZ_OK;
    cudaSetDevice(ctx->device);
    Series* s = r->s;
    const long nn = (long)s->n * s->n;
    size_t bytes = (size_t)r->nnz * nn * sizeof(double2);
    size_t free_b = 0, total_b = 0;
    CU(ctx, cudaMemGetInfo(&free_b, &total_b));
    if (bytes > free_b - std::min<size_t>(free_b, ctx->budget + ((size_t)1 << 30)))
        return fail(ctx, ABZ_E_OOM, "H(k) on this rule does not fit in device memory; use the streamed sums");
    CU(ctx, pool_alloc(ctx, (void**)&r->d_H, bytes));
    long rows1 = nn * s->M[0], rows2 = rows1 * s->M[1];
    auto chunks = plan_chunks(r, node_cap_for(ctx, s->n), (long)(ctx->budget / (rows1 * sizeof(double2))),
                              (long)(ctx->budget / (rows2 * sizeof(double2))));
    cudaEvent_t e0 = next_event(ctx);
    for (auto& ch : chunks) {
        int rc = eval_chunk(ctx, r, ch, true, r->d_H + r->h_row_nodeptr[ch.r0] * nn);
        if (rc) { pool_free(ctx, r->d_H); r->d_H = nullptr; return rc; }
    }
    cudaEvent_t e1 = next_event(ctx);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    collect_timings(ctx, {{e0, e1}}, {});
    return ABZ_OK;
}

int32_t abz_rule_copy_out(abz_ctx* ctx, abz_rule_t rid, double* Hk, double* kfrac, double* w) {
    if (!ctx) return ABZ_E_INVALID;
    Rule* r = get_rule(ctx, rid);
    if (!r) return fail(ctx, ABZ_E_INVALID, "unknown rule handle");
    cudaSetDevice(ctx->device);
    Series* s = r->s;
    const long nn = (long)s->n * s->n;
    if ((kfrac || w) && !r->full && !r->nodes_on_host && r->nnz > 0) {
        r->h_node_k1.resize(r->nnz); r->h_node_w.resize(r->nnz);
        CU(ctx, cudaMemcpyAsync(r->h_node_k1.data(), r->d_node_k1, (size_t)r->nnz * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaMemcpyAsync(r->h_node_w.data(), r->d_node_w, (size_t)r->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        r->nodes_on_host = true;
    }
    if (kfrac || w) {
        for (long p = 0; p < r->np3; p++)
            for (long q = r->h_plane_rowptr[p]; q < r->h_plane_rowptr[p + 1]; q++)
                for (long i = r->h_row_nodeptr[q]; i < r->h_row_nodeptr[q + 1]; i++) {
                    int k1 = r->full ? (int)(i - r->h_row_nodeptr[q]) : r->h_node_k1[i];
                    if (kfrac) {
                        kfrac[3 * i] = (double)k1 / r->N;
                        kfrac[3 * i + 1] = (double)r->h_row_k2[q] / r->N;
                        kfrac[3 * i + 2] = (double)r->h_plane_k3[p] / r->N;
                    }
                    if (w) w[i] = r->full ? 1.0 : r->h_node_w[i];
                }
    }
    if (!Hk) return ABZ_OK;
    if (r->d_H) {
        CU(ctx, cudaMemcpyAsync(Hk, r->d_H, (size_t)r->nnz * nn * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        return ABZ_OK;
    }
    long rows1 = nn * s->M[0], rows2 = rows1 * s->M[1];
    auto chunks = plan_chunks(r, node_cap_for(ctx, s->n), (long)(ctx->budget / (rows1 * sizeof(double2))),
                              (long)(ctx->budget / (rows2 * sizeof(double2))));
    for (auto& ch : chunks) {
        long n0 = r->h_row_nodeptr[ch.r0], n1 = r->h_row_nodeptr[ch.r1];
        CU(ctx, ctx->Hc.reserve((size_t)(n1 - n0) * nn * sizeof(double2)));
        int rc = eval_chunk(ctx, r, ch, true, ctx->Hc.as<double2>());
        if (rc) return rc;
        CU(ctx, cudaMemcpyAsync(Hk + 2 * n0 * nn, ctx->Hc.p, (size_t)(n1 - n0) * nn * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return ABZ_OK;
}

static int32_t abz_rule_resolvent_sum_impl(abz_ctx* ctx, abz_rule_t rid, int32_t fkind, int32_t nw, const double* z,
                               const double* sigma, double scale, double* out) {
    if (!ctx) return ABZ_E_INVALID;
    Rule* r = get_rule(ctx, rid);
    if (!r) return fail(ctx, ABZ_E_INVALID, "unknown rule handle");
    if (fkind != ABZ_F_RESOLVENT_TRACE && fkind != ABZ_F_TRACE_H) return fail(ctx, ABZ_E_INVALID, "unknown integrand kind");
    if (fkind == ABZ_F_TRACE_H) { nw = 1; sigma = nullptr; }
    if (nw < 1 || !out || (fkind == ABZ_F_RESOLVENT_TRACE && !z)) return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    cudaSetDevice(ctx->device);
    Series* s = r->s;
    const int n = s->n;
    const long nn = (long)n * n;
    int rc = upload_params(ctx, n, nw, fkind == ABZ_F_RESOLVENT_TRACE ? z : nullptr, sigma);
    if (rc) return rc;
    CU(ctx, ctx->acc.reserve((size_t)nw * sizeof(double2)));
    CU(ctx, cudaMemsetAsync(ctx->acc.p, 0, (size_t)nw * sizeof(double2), ctx->stream));
    const double2* dz = ctx->zbuf.as<double2>();
    const double2* dsig = sigma ? ctx->sigbuf.as<double2>() : nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_eval, ev_mat;
    ctx->ev_used = 0;
    const bool small = (n <= 3);
    if (r->d_H) {
        cudaEvent_t e0 = next_event(ctx);
        if (small) rc = run_small_fused<true>(ctx, r, nullptr, r->d_H, 0, r->nrows, fkind, nw, dz, dsig);
        else rc = run_matfun(ctx, r->d_H, r->d_node_w, r->nnz, n, fkind, nw, dz, dsig, 0, nullptr);
        if (rc) return rc;
        ev_mat.push_back({e0, next_event(ctx)});
    } else {
        const bool fused = small && ctx->fused_small;
        // K3-fused: the DMMA one-warp kernel forms H(k) itself from the C1 rows (stage 1 folded in, no H(k) in HBM)
        static const int fused_mma_env = getenv("ABZ_FUSED_MMA") ? atoi(getenv("ABZ_FUSED_MMA")) : 1;
        const bool fused_mma = !small && fused_mma_env && fkind == ABZ_F_RESOLVENT_TRACE && nw >= 8 && mma_resolvent_supported(n) &&
                               (ctx->resolvent_algo == 0 || ctx->resolvent_algo == 2) && !ctx->force_generic &&
                               !(getenv("ABZ_MMA_TEAM") && atoi(getenv("ABZ_MMA_TEAM")) == 1 && n > 24) &&
                               mma_resolvent_variant() == ABZ_MMA_DEFAULT_VARIANT && mma_resolvent_warps() >= 8 &&
                               mma_fused_smem(n, nw) <= 160 * 1024;
        long rows1 = nn * s->M[0], rows2 = rows1 * s->M[1];
        long ncap = (fused || fused_mma) ? ((long)1 << 40) : node_cap_for(ctx, n);
        auto chunks = plan_chunks(r, ncap, (long)(ctx->budget / (rows1 * sizeof(double2))),
                                  (long)(ctx->budget / (rows2 * sizeof(double2))));
        for (auto& ch : chunks) {
            long n0 = r->h_row_nodeptr[ch.r0], n1 = r->h_row_nodeptr[ch.r1];
            cudaEvent_t e0 = next_event(ctx);
            if (!fused && !fused_mma) CU(ctx, ctx->Hc.reserve((size_t)(n1 - n0) * nn * sizeof(double2)));
            rc = eval_chunk(ctx, r, ch, !fused && !fused_mma, ctx->Hc.as<double2>());
            if (rc) return rc;
            cudaEvent_t e1 = next_event(ctx);
            if (fused) rc = run_small_fused<false>(ctx, r, ctx->C1.as<double2>(), nullptr, ch.r0, ch.r1 - ch.r0, fkind, nw, dz, dsig);
            else if (fused_mma) {
                long ncta = 0; int kper_m = 1;
                if (n1 > n0 && mma_resolvent_plan(n, n1 - n0, nw, ctx->sm_count, &ncta, &kper_m) == 0) {
                    CU(ctx, ctx->partial.reserve((size_t)ncta * nw * sizeof(double2)));
                    CU(ctx, mma_fused_launch(ctx->C1.as<double2>(), r->d_ptab[0], r->d_row_nodeptr, ch.r0, ch.r1, r->d_node_k1, r->N, s->M[0],
                                             r->d_node_w, n0, n1 - n0, n, nw, dz, dsig, ctx->partial.as<double2>(), ctx->errflag.as<int>(),
                                             ncta, kper_m, ctx->stream));
                    ctx->launches++;
                    reduce_partials_kernel<<<nw, 256, 0, ctx->stream>>>(ctx->partial.as<double2>(), ncta, nw, 1.0, ctx->acc.as<double2>());
                    LAUNCH_CHECK(ctx, "reduce_partials_kernel");
                } else if (n1 > n0) rc = fail(ctx, ABZ_E_UNSUPPORTED, "internal: no launch plan for the fused DMMA resolvent");
            } else if (small) {
                // unfused small path (option): treat the chunk as a materialised block
                rc = fail(ctx, ABZ_E_UNSUPPORTED, "unfused small-norb streaming is not implemented; materialize the rule");
            } else
                rc = run_matfun(ctx, ctx->Hc.as<double2>(), r->d_node_w ? r->d_node_w + n0 : nullptr, n1 - n0, n, fkind, nw, dz, dsig, 0, nullptr);
            if (rc) return rc;
            cudaEvent_t e2 = next_event(ctx);
            ev_eval.push_back({e0, e1});
            ev_mat.push_back({e1, e2});
        }
    }
    std::vector<double2> h(nw);
    CU(ctx, cudaMemcpyAsync(h.data(), ctx->acc.p, (size_t)nw * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    rc = check_errflag(ctx, "abz_rule_resolvent_sum");
    collect_timings(ctx, ev_eval, ev_mat);
    if (rc) return rc;
    for (int w = 0; w < nw; w++) { out[2 * w] = scale * h[w].x; out[2 * w + 1] = scale * h[w].y; }
    return ABZ_OK;
}

int32_t abz_rule_resolvent_sum(abz_ctx* ctx, abz_rule_t rid, int32_t fkind, int32_t nw, const double* z,
                               const double* sigma, double scale, double* out) {
    int32_t rc = abz_rule_resolvent_sum_impl(ctx, rid, fkind, nw, z, sigma, scale, out);
    if (rc == ABZ_RETRY_PIVOTED) {
        ctx->force_generic = true;
        rc = abz_rule_resolvent_sum_impl(ctx, rid, fkind, nw, z, sigma, scale, out);
        ctx->force_generic = false;
    }
    return rc;
}


// G(w) = scale * sum_i w_i (z_w - H(k_i) - Sigma_w)^-1, matrix-valued (docs/src/examples.md:20,90)
int32_t abz_rule_resolvent_matrix_sum(abz_ctx* ctx, abz_rule_t rid, int32_t nw, const double* z, const double* sigma, double scale,
                                      double* out) {
    if (!ctx) return ABZ_E_INVALID;
    Rule* r = get_rule(ctx, rid);
    if (!r) return fail(ctx, ABZ_E_INVALID, "unknown rule handle");
    if (nw < 1 || !z || !out) return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    cudaSetDevice(ctx->device);
    Series* s = r->s;
    const int n = s->n;
    const long nn = (long)n * n;
    if (n > 64) return fail(ctx, ABZ_E_UNSUPPORTED, "norb > 64 is not supported by the resolvent kernels");
    int rc = upload_params(ctx, n, nw, z, sigma);
    if (rc) return rc;
    const size_t nacc = (size_t)nw * nn;
    CU(ctx, ctx->acc.reserve(nacc * sizeof(double2)));
    CU(ctx, cudaMemsetAsync(ctx->acc.p, 0, nacc * sizeof(double2), ctx->stream));
    const size_t per_warp = ((size_t)n * (n + 1) + (n + 1) / 2 + 1 + nn) * sizeof(double2);
    const int nwarps = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / per_warp));
    const size_t smem = per_warp * nwarps;
    std::vector<Chunk> chunks;
    long rows1 = nn * s->M[0], rows2 = rows1 * s->M[1];
    if (r->d_H) chunks.push_back({0, r->np3, 0, r->nrows});
    else chunks = plan_chunks(r, node_cap_for(ctx, n), (long)(ctx->budget / (rows1 * sizeof(double2))), (long)(ctx->budget / (rows2 * sizeof(double2))));
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_eval, ev_mat;
    ctx->ev_used = 0;
    for (auto& ch : chunks) {
        const long n0 = r->h_row_nodeptr[ch.r0], n1 = r->h_row_nodeptr[ch.r1], nk = n1 - n0;
        if (nk <= 0) continue;
        cudaEvent_t e0 = next_event(ctx);
        const double2* Hd = r->d_H ? r->d_H + n0 * nn : nullptr;
        if (!Hd) {
            CU(ctx, ctx->Hc.reserve((size_t)nk * nn * sizeof(double2)));
            rc = eval_chunk(ctx, r, ch, true, ctx->Hc.as<double2>());
            if (rc) return rc;
            Hd = ctx->Hc.as<double2>();
        }
        cudaEvent_t e1 = next_event(ctx);
        // node chunks per frequency: enough CTAs to fill the chip, few enough that the partials stay small
        const bool legacy = (ctx->resolvent_algo == 4) || (n <= 20 && ctx->resolvent_algo != 1);   // n <= 20: a padded 32-row team wastes its lanes (measured crossover)
        const long per_sm = legacy ? 2 : gj_ctas_per_sm(n, true);
        const long target = std::max<long>(1, ((long)ctx->sm_count * per_sm + nw - 1) / nw);
        const int kper = (int)std::max<long>(legacy ? nwarps : 1, (nk + target - 1) / target);
        const long ncta = (nk + kper - 1) / kper;
        CU(ctx, ctx->partial.reserve((size_t)ncta * nacc * sizeof(double2)));
        dim3 grid((unsigned)ncta, (unsigned)nw);
        const double* wn = r->d_node_w ? r->d_node_w + n0 : nullptr;
        const double2* sgd = sigma ? ctx->sigbuf.as<double2>() : nullptr;
        if (legacy) {
            resolvent_gj_matrix_kernel<<<grid, nwarps * 32, smem, ctx->stream>>>(Hd, wn, nk, n, nw, ctx->zbuf.as<double2>(), sgd, kper,
                                                                                ctx->partial.as<double2>(), ctx->errflag.as<int>());
        } else {
            GJ_DISPATCH(resolvent_gjreg_matrix_kernel, true, n, grid, (size_t)nn * sizeof(double2), ctx->stream, Hd, wn, nk, n, nw,
                        ctx->zbuf.as<double2>(), sgd, kper, ctx->partial.as<double2>(), ctx->errflag.as<int>());
        }
        LAUNCH_CHECK(ctx, "resolvent_gj_matrix_kernel");
        const long nred = (long)nacc;
        for (long off = 0; off < nred; off += 65535) {      // one CTA per output element
            const int cnt = (int)std::min<long>(65535, nred - off);
            reduce_partials_strided_kernel<<<cnt, 256, 0, ctx->stream>>>(ctx->partial.as<double2>() + off, ncta, nred, 1.0, ctx->acc.as<double2>() + off);
            LAUNCH_CHECK(ctx, "reduce_partials_strided_kernel");
        }
        cudaEvent_t e2 = next_event(ctx);
        ev_eval.push_back({e0, e1});
        ev_mat.push_back({e1, e2});
    }
    std::vector<double2> h(nacc);
    CU(ctx, cudaMemcpyAsync(h.data(), ctx->acc.p, nacc * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    rc = check_errflag(ctx, "abz_rule_resolvent_matrix_sum");
    collect_timings(ctx, ev_eval, ev_mat);
    if (rc == ABZ_RETRY_PIVOTED) rc = ABZ_OK;
    if (rc) return rc;
    for (size_t i = 0; i < nacc; i++) { out[2 * i] = scale * h[i].x; out[2 * i + 1] = scale * h[i].y; }
    return ABZ_OK;
}

static size_t eig_smem_bytes(int n, int threads) {
    int np = (n + 1) & ~1, npair = np / 2;
    return (size_t)n * (n + 1) * 16 + (size_t)npair * 16 + (size_t)npair * 8 + (size_t)n * 8 + (size_t)(2 * (threads / 32) + 2) * 8 +
           (size_t)npair * 8 + 64;
}

static int run_eig_jacobi(abz_ctx* ctx, const double2* H, const double* wnode, long nk, int n, int mode, int kind, double p0, double p1,
                          double* evals, double* acc) {
    int threads = std::min(256, std::max(32, ((n * ((n + 1) / 2) + 31) / 32) * 32));
    size_t smem = eig_smem_bytes(n, threads);
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (200 * 1024) / smem));
    long ncta = std::min<long>(nk, (long)ctx->sm_count * per_sm);
    if (mode == 0) CU(ctx, ctx->partial.reserve((size_t)ncta * sizeof(double)));
    eig_jacobi_kernel<<<(unsigned)ncta, threads, smem, ctx->stream>>>(H, wnode, nk, n, mode, kind, p0, p1, evals,
                                                                     ctx->partial.as<double>(), ctx->errflag.as<int>());
    LAUNCH_CHECK(ctx, "eig_jacobi_kernel");
    if (mode == 0) {
        reduce_real_kernel<<<1, 256, 0, ctx->stream>>>(ctx->partial.as<double>(), ncta, 1.0, acc);
        LAUNCH_CHECK(ctx, "reduce_real_kernel");
    }
    return ABZ_OK;
}

// Stage B choice (profiles/r02_eig_stage_b_timing.log, one B200): the warp-per-matrix bisection kernel wins while the batch leaves the
// thread-per-matrix QL kernel latency-bound - up to ~5 k matrices at n = 16, ~12 k at n = 32, ~16 k at n = 48, every measured batch
// (67 k) at n = 64.  ABZ_EIG_BISECT_MAX overrides the batch limit (0 = never bisect).
static bool eig_use_bisect(const abz_ctx* ctx, long nk, int n) {
    static const long lim = [] { const char* e = getenv("ABZ_EIG_BISECT_MAX"); return e ? atol(e) : -1L; }();
    if (ctx->eig_algo == 3 || n > EIG_MAXN) return false;        // ABZ_OPT_EIG_ALGO 3: always QL, 4: always bisection
    if (ctx->eig_algo == 4) return true;
    if (n < 12) return false;
    const long t = lim >= 0 ? lim : (n < 24 ? 5000L : n < 40 ? 12000L : n < 56 ? 16000L : 150000L);
    return nk <= t;
}

// eigenvalues of nk materialised matrices: Householder tridiagonalisation (CTA per matrix) + implicit QL (thread per matrix)
static int run_eig(abz_ctx* ctx, const double2* H, const double* wnode, long nk, int n, int mode, int kind, double p0, double p1,
                   double* evals, double* acc) {
    if (nk <= 0) return ABZ_OK;
    if (ctx->eig_algo == 1 || n > EIG_MAXN) return run_eig_jacobi(ctx, H, wnode, nk, n, mode == 2 ? 1 : mode, kind, p0, p1, evals, acc);
    { int rct = launch_tridiag(ctx, H, nk, n); if (rct) return rct; }
    // stage B: a warp per matrix (Sturm bisection) while the batch is small enough that the thread-per-matrix QL kernel would be bound by
    // the latency of one thread (1.6 ms at n = 64 whatever the batch); QL for large batches, where its lower operation count wins
    if (eig_use_bisect(ctx, nk, n)) {
        const long nb = (nk + BIS_WARPS - 1) / BIS_WARPS;
        if (mode == 0) CU(ctx, ctx->partial.reserve((size_t)nb * sizeof(double)));
        eig_bisect_kernel<<<(unsigned)nb, BIS_WARPS * 32, 0, ctx->stream>>>(ctx->eig_d.as<double>(), ctx->eig_e.as<double>(), wnode, nk, n, mode, kind,
                                                                          p0, p1, evals, ctx->partial.as<double>(), ctx->errflag.as<int>());
        LAUNCH_CHECK(ctx, "eig_bisect_kernel");
        if (mode == 0) {
            reduce_real_kernel<<<1, 256, 0, ctx->stream>>>(ctx->partial.as<double>(), nb, 1.0, acc);
            LAUNCH_CHECK(ctx, "reduce_real_kernel");
        }
        return ABZ_OK;
    }
    const long nblk = (nk + 31) / 32;
    if (mode == 0) CU(ctx, ctx->partial.reserve((size_t)nblk * sizeof(double)));
    eig_tql_kernel<<<(unsigned)nblk, 32, (size_t)2 * n * 32 * sizeof(double), ctx->stream>>>(ctx->eig_d.as<double>(), ctx->eig_e.as<double>(), wnode, nk, n, mode, kind, p0,
                                                          p1, evals, ctx->partial.as<double>(), ctx->errflag.as<int>());
    LAUNCH_CHECK(ctx, "eig_tql_kernel");
    if (mode == 0) {
        reduce_real_kernel<<<1, 256, 0, ctx->stream>>>(ctx->partial.as<double>(), nblk, 1.0, acc);
        LAUNCH_CHECK(ctx, "reduce_real_kernel");
    }
    return ABZ_OK;
}

// sums over the eigenvalues in `evals` ([nk][n]) for every parameter set: acc[p] += sum_k w_k sum_b g_p(e_b)
static int run_eig_cached_sums(abz_ctx* ctx, const double* evals, const double* wnode, long nk, int n, int kind, int nprm,
                               const double* dparams, double* acc) {
    if (nk <= 0) return ABZ_OK;
    const long ncta = std::max<long>(1, std::min<long>((nk + 255) / 256, (long)ctx->sm_count * 4));
    CU(ctx, ctx->partial.reserve((size_t)ncta * nprm * sizeof(double)));
    for (int p0 = 0; p0 < nprm; p0 += 65535) {
        const int cnt = std::min(65535, nprm - p0);
        dim3 grid((unsigned)ncta, (unsigned)cnt);
        eig_cached_sum_kernel<<<grid, 256, 0, ctx->stream>>>(evals, wnode, nk, n, kind, cnt, dparams + 2 * p0, ctx->partial.as<double>());
        LAUNCH_CHECK(ctx, "eig_cached_sum_kernel");
        reduce_real_strided_kernel<<<cnt, 256, 0, ctx->stream>>>(ctx->partial.as<double>(), ncta, cnt, 1.0, acc + p0);
        LAUNCH_CHECK(ctx, "reduce_real_strided_kernel");
    }
    return ABZ_OK;
}

int32_t abz_rule_eig_sum_batch(abz_ctx* ctx, abz_rule_t rid, int32_t kind, int32_t nparams, const double* params, double scale,
                               double* out) {
    if (!ctx) return ABZ_E_INVALID;
    Rule* r = get_rule(ctx, rid);
    if (!r) return fail(ctx, ABZ_E_INVALID, "unknown rule handle");
    if (kind < 0 || kind > 3 || !out || nparams < 1 || (kind != 0 && !params)) return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    cudaSetDevice(ctx->device);
    Series* s = r->s;
    const int n = s->n;
    const long nn = (long)n * n;
    std::vector<double> hp((size_t)2 * nparams);
    for (int p = 0; p < nparams; p++) { hp[2 * p] = params ? params[2 * p] : 0.0; hp[2 * p + 1] = params ? params[2 * p + 1] : 1.0; }
    CU(ctx, ctx->zbuf.reserve(hp.size() * sizeof(double)));
    CU(ctx, cudaMemcpyAsync(ctx->zbuf.p, hp.data(), hp.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, ctx->acc.reserve((size_t)nparams * sizeof(double)));
    CU(ctx, cudaMemsetAsync(ctx->acc.p, 0, (size_t)nparams * sizeof(double), ctx->stream));
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_eval, ev_mat;
    ctx->ev_used = 0;
    int rc;
    const double* dprm = ctx->zbuf.as<double>();
    if (r->d_eig) {
        // a materialised rule that has been diagonalised before: parameters sweep over the cached eigenvalues
        cudaEvent_t e0 = next_event(ctx);
        rc = run_eig_cached_sums(ctx, r->d_eig, r->d_node_w, r->nnz, n, kind, nparams, dprm, ctx->acc.as<double>());
        if (rc) return rc;
        ev_mat.push_back({e0, next_event(ctx)});
    } else {
        std::vector<Chunk> chunks;
        long rows1 = nn * s->M[0], rows2 = rows1 * s->M[1];
        if (r->d_H) chunks.push_back({0, r->np3, 0, r->nrows});
        else chunks = plan_chunks(r, node_cap_for(ctx, n), (long)(ctx->budget / (rows1 * sizeof(double2))),
                                  (long)(ctx->budget / (rows2 * sizeof(double2))));
        double* keep = nullptr;
        if (r->d_H && r->nnz > 0) CU(ctx, pool_alloc(ctx, (void**)&keep, (size_t)r->nnz * n * sizeof(double)));
        for (auto& ch : chunks) {
            const long n0 = r->h_row_nodeptr[ch.r0], n1 = r->h_row_nodeptr[ch.r1], nk = n1 - n0;
            if (nk <= 0) continue;
            cudaEvent_t e0 = next_event(ctx);
            const double2* Hd = r->d_H ? r->d_H + n0 * nn : nullptr;
            if (!Hd) {
                CU(ctx, ctx->Hc.reserve((size_t)nk * nn * sizeof(double2)));
                rc = eval_chunk(ctx, r, ch, true, ctx->Hc.as<double2>());
                if (rc) { pool_free(ctx, keep); return rc; }
                Hd = ctx->Hc.as<double2>();
            }
            cudaEvent_t e1 = next_event(ctx);
            double* ev = keep ? keep + n0 * n : nullptr;
            if (!ev) { CU(ctx, ctx->tmp_a.reserve((size_t)nk * n * sizeof(double))); ev = ctx->tmp_a.as<double>(); }
            rc = run_eig(ctx, Hd, nullptr, nk, n, 2, 0, 0, 1, ev, nullptr);
            if (!rc) rc = run_eig_cached_sums(ctx, ev, r->d_node_w ? r->d_node_w + n0 : nullptr, nk, n, kind, nparams, dprm, ctx->acc.as<double>());
            if (rc) { pool_free(ctx, keep); return rc; }
            cudaEvent_t e2 = next_event(ctx);
            ev_eval.push_back({e0, e1});
            ev_mat.push_back({e1, e2});
        }
        if (keep) {
            rc = check_errflag(ctx, "abz_rule_eig_sum");
            if (rc) { pool_free(ctx, keep); collect_timings(ctx, ev_eval, ev_mat); return rc; }
            r->d_eig = keep;
        }
    }
    std::vector<double> h((size_t)nparams);
    CU(ctx, cudaMemcpyAsync(h.data(), ctx->acc.p, (size_t)nparams * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    rc = check_errflag(ctx, "abz_rule_eig_sum");
    collect_timings(ctx, ev_eval, ev_mat);
    if (rc) return rc;
    for (int p = 0; p < nparams; p++) out[p] = scale * h[p];
    return ABZ_OK;
}

int32_t abz_rule_eig_sum(abz_ctx* ctx, abz_rule_t rid, int32_t kind, const double* params, double scale, double* out) {
    return abz_rule_eig_sum_batch(ctx, rid, kind, 1, params, scale, out);
}

int32_t abz_rule_eigvals(abz_ctx* ctx, abz_rule_t rid, double* evals) {
    if (!ctx) return ABZ_E_INVALID;
    Rule* r = get_rule(ctx, rid);
    if (!r) return fail(ctx, ABZ_E_INVALID, "unknown rule handle");
    if (!evals) return fail(ctx, ABZ_E_INVALID, "NULL output");
    cudaSetDevice(ctx->device);
    Series* s = r->s;
    const int n = s->n;
    const long nn = (long)n * n;
    long rows1 = nn * s->M[0], rows2 = rows1 * s->M[1];
    std::vector<Chunk> chunks;
    if (r->d_H) chunks.push_back({0, r->np3, 0, r->nrows});
    else chunks = plan_chunks(r, node_cap_for(ctx, n), (long)(ctx->budget / (rows1 * sizeof(double2))),
                              (long)(ctx->budget / (rows2 * sizeof(double2))));
    for (auto& ch : chunks) {
        long n0 = r->h_row_nodeptr[ch.r0], n1 = r->h_row_nodeptr[ch.r1];
        const double2* H = r->d_H;
        if (!H) {
            CU(ctx, ctx->Hc.reserve((size_t)(n1 - n0) * nn * sizeof(double2)));
            int rc = eval_chunk(ctx, r, ch, true, ctx->Hc.as<double2>());
            if (rc) return rc;
            H = ctx->Hc.as<double2>();
        }
        CU(ctx, ctx->tmp_a.reserve((size_t)(n1 - n0) * n * sizeof(double)));
        int rc = run_eig(ctx, H, nullptr, n1 - n0, n, 1, 0, 0, 1, ctx->tmp_a.as<double>(), nullptr);
        if (rc) return rc;
        CU(ctx, cudaMemcpyAsync(evals + n0 * n, ctx->tmp_a.p, (size_t)(n1 - n0) * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return check_errflag(ctx, "abz_rule_eigvals");
}

// ---- GGR data pass and sum (src/dos_ggr.jl) ----------------------------------------------------------------------
int32_t abz_rule_ggr_data(abz_ctx* ctx, abz_rule_t rid, int32_t ndim, double* energies, double* velocities) {
    if (!ctx) return ABZ_E_INVALID;
    Rule* r = get_rule(ctx, rid);
    if (!r) return fail(ctx, ABZ_E_INVALID, "unknown rule handle");
    Series* s = r->s;
    if (ndim < 1 || ndim > 3) return fail(ctx, ABZ_E_INVALID, "GGR implemented for up to 3d BZ");
    for (int d = ndim; d < 3; d++)
        if (s->M[d] != 1) return fail(ctx, ABZ_E_INVALID, "variables in Fourier series don't match domain");
    cudaSetDevice(ctx->device);
    const int n = s->n;
    const long nn = (long)n * n;
    if (n > 64) return fail(ctx, ABZ_E_UNSUPPORTED, "norb > 64 is not supported by the GGR data pass");
    if (r->nnz > 0 && (!r->d_ggr_e || r->ggr_ndim != ndim)) {
        pool_free(ctx, r->d_ggr_e); pool_free(ctx, r->d_ggr_v); r->d_ggr_e = r->d_ggr_v = nullptr;
        CU(ctx, pool_alloc(ctx, (void**)&r->d_ggr_e, (size_t)r->nnz * n * sizeof(double)));
        CU(ctx, pool_alloc(ctx, (void**)&r->d_ggr_v, (size_t)r->nnz * n * ndim * sizeof(double)));
        // JacobianSeries coefficients: 2 pi i R_d / period_d * H_R
        const size_t ctot = (size_t)nn * s->M[0] * s->M[1] * s->M[2];
        DevBuf dco[3], Vc[3];
        for (int d = 0; d < ndim; d++) {
            CU(ctx, dco[d].reserve(ctot * sizeof(double2)));
            jacobian_coeff_kernel<<<(unsigned)((ctot + 255) / 256), 256, 0, ctx->stream>>>(s->c, dco[d].as<double2>(), nn, s->M[0], s->M[1],
                                                                                          s->M[2], s->lo[d], d, s->period[d]);
            LAUNCH_CHECK(ctx, "jacobian_coeff_kernel");
        }
        const int np = (n + 1) & ~1, npair = np / 2;
        const int threads = std::min(256, std::max(64, ((n * ((n + 1) / 2) + 31) / 32) * 32));
        const size_t smem = ((size_t)n * (n + 1) + (size_t)n * n + npair) * 16 + (size_t)(npair + 3 * n + 2 * (threads / 32) + 2) * 8 +
                            (size_t)2 * npair * 4 + 64;
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / smem));
        long rows1 = nn * s->M[0], rows2 = rows1 * s->M[1];
        const size_t share = ctx->budget / 4;
        auto chunks = plan_chunks(r, (long)std::max<size_t>(1, share / (nn * sizeof(double2))), (long)(ctx->budget / (rows1 * sizeof(double2))),
                                  (long)(ctx->budget / (rows2 * sizeof(double2))));
        cudaEvent_t e0 = next_event(ctx);
        for (auto& ch : chunks) {
            const long n0 = r->h_row_nodeptr[ch.r0], n1 = r->h_row_nodeptr[ch.r1], nk = n1 - n0;
            if (nk <= 0) continue;
            CU(ctx, ctx->Hc.reserve((size_t)nk * nn * sizeof(double2)));
            int rc = eval_chunk(ctx, r, ch, true, ctx->Hc.as<double2>());
            if (rc) return rc;
            for (int d = 0; d < ndim; d++) {
                CU(ctx, Vc[d].reserve((size_t)nk * nn * sizeof(double2)));
                rc = eval_chunk(ctx, r, ch, true, Vc[d].as<double2>(), dco[d].as<double2>());
                if (rc) return rc;
            }
            const long ncta = std::min<long>(nk, (long)ctx->sm_count * per_sm);
            eig_jacobi_vel_kernel<<<(unsigned)ncta, threads, smem, ctx->stream>>>(
                ctx->Hc.as<double2>(), Vc[0].as<double2>(), Vc[1].as<double2>(), Vc[2].as<double2>(), nk, n, ndim, s->period[0], s->period[1],
                s->period[2], r->d_ggr_e + n0 * n, r->d_ggr_v + n0 * n * ndim, ctx->errflag.as<int>());
            LAUNCH_CHECK(ctx, "eig_jacobi_vel_kernel");
        }
        cudaEvent_t e1 = next_event(ctx);
        int rc = check_errflag(ctx, "abz_rule_ggr_data");
        collect_timings(ctx, {}, {{e0, e1}});
        if (rc) { pool_free(ctx, r->d_ggr_e); pool_free(ctx, r->d_ggr_v); r->d_ggr_e = r->d_ggr_v = nullptr; return rc; }
        r->ggr_ndim = ndim;
    }
    if (energies && r->nnz > 0)
        CU(ctx, cudaMemcpyAsync(energies, r->d_ggr_e, (size_t)r->nnz * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (velocities && r->nnz > 0)
        CU(ctx, cudaMemcpyAsync(velocities, r->d_ggr_v, (size_t)r->nnz * n * ndim * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ABZ_OK;
}

int32_t abz_rule_ggr_sum(abz_ctx* ctx, abz_rule_t rid, int32_t nE, const double* E, double scale, double* out) {
    if (!ctx) return ABZ_E_INVALID;
    Rule* r = get_rule(ctx, rid);
    if (!r) return fail(ctx, ABZ_E_INVALID, "unknown rule handle");
    if (nE < 1 || !E || !out) return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    if (r->nnz > 0 && !r->d_ggr_e) return fail(ctx, ABZ_E_INVALID, "abz_rule_ggr_sum: call abz_rule_ggr_data first");
    cudaSetDevice(ctx->device);
    if (r->nnz == 0) { for (int i = 0; i < nE; i++) out[i] = 0.0; return ABZ_OK; }
    CU(ctx, ctx->tmp_a.reserve((size_t)nE * sizeof(double)));
    CU(ctx, ctx->tmp_c.reserve((size_t)nE * sizeof(double)));
    CU(ctx, cudaMemcpyAsync(ctx->tmp_a.p, E, (size_t)nE * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ggr_sum_kernel<<<(unsigned)nE, 256, 0, ctx->stream>>>(r->d_ggr_e, r->d_ggr_v, r->d_node_w, r->nnz, r->s->n, r->ggr_ndim, 1.0 / (2.0 * r->N),
                                                         ctx->tmp_a.as<double>(), scale, ctx->tmp_c.as<double>());
    LAUNCH_CHECK(ctx, "ggr_sum_kernel");
    CU(ctx, cudaMemcpyAsync(out, ctx->tmp_c.p, (size_t)nE * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ABZ_OK;
}

// ---- scattered points ---------------------------------------------------------------------------
static int points_eval_dev(abz_ctx* ctx, Series* s, int64_t npts, const double* k, double2** Hdev) {
    const long nn = (long)s->n * s->n;
    CU(ctx, ctx->tmp_a.reserve((size_t)npts * 3 * sizeof(double)));
    CU(ctx, ctx->Hc.reserve((size_t)npts * nn * sizeof(double2)));
    CU(ctx, cudaMemcpyAsync(ctx->tmp_a.p, k, (size_t)npts * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    size_t smem = (size_t)(s->M[0] + s->M[1] + s->M[2]) * sizeof(double2);
    points_eval_kernel<<<(unsigned)npts, 128, smem, ctx->stream>>>(s->c, (int)nn, s->M[0], s->M[1], s->M[2], s->lo[0], s->lo[1],
                                                                  s->lo[2], s->period[0], s->period[1], s->period[2],
                                                                  ctx->tmp_a.as<double>(), ctx->Hc.as<double2>());
    LAUNCH_CHECK(ctx, "points_eval_kernel");
    *Hdev = ctx->Hc.as<double2>();
    return ABZ_OK;
}

int32_t abz_points_eval(abz_ctx* ctx, abz_series_t sid, int64_t npts, const double* k, double* Hk) {
    if (!ctx) return ABZ_E_INVALID;
    Series* s = get_series(ctx, sid);
    if (!s) return fail(ctx, ABZ_E_INVALID, "unknown series handle");
    if (npts < 0 || (npts > 0 && (!k || !Hk))) return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    if (npts == 0) return ABZ_OK;
    cudaSetDevice(ctx->device);
    double2* Hd;
    int rc = points_eval_dev(ctx, s, npts, k, &Hd);
    if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(Hk, Hd, (size_t)npts * s->n * s->n * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ABZ_OK;
}

static int32_t abz_points_resolvent_impl(abz_ctx* ctx, abz_series_t sid, int64_t npts, const double* k, int32_t fkind, int32_t nw,
                             const double* z, const double* sigma, double* y) {
    if (!ctx) return ABZ_E_INVALID;
    Series* s = get_series(ctx, sid);
    if (!s) return fail(ctx, ABZ_E_INVALID, "unknown series handle");
    if (fkind == ABZ_F_TRACE_H) { nw = 1; sigma = nullptr; }
    if (npts < 0 || nw < 1 || (npts > 0 && (!k || !y)) || (fkind == ABZ_F_RESOLVENT_TRACE && !z))
        return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    if (npts == 0) return ABZ_OK;
    cudaSetDevice(ctx->device);
    double2* Hd;
    int rc = points_eval_dev(ctx, s, npts, k, &Hd);
    if (rc) return rc;
    rc = upload_params(ctx, s->n, nw, fkind == ABZ_F_RESOLVENT_TRACE ? z : nullptr, sigma);
    if (rc) return rc;
    CU(ctx, ctx->tmp_b.reserve((size_t)npts * nw * sizeof(double2)));
    rc = run_matfun(ctx, Hd, nullptr, npts, s->n, fkind, nw, ctx->zbuf.as<double2>(), sigma ? ctx->sigbuf.as<double2>() : nullptr, 1,
                    ctx->tmp_b.as<double2>());
    if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(y, ctx->tmp_b.p, (size_t)npts * nw * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    return check_errflag(ctx, "abz_points_resolvent");
}

int32_t abz_points_resolvent(abz_ctx* ctx, abz_series_t sid, int64_t npts, const double* k, int32_t fkind, int32_t nw,
                             const double* z, const double* sigma, double* y) {
    int32_t rc = abz_points_resolvent_impl(ctx, sid, npts, k, fkind, nw, z, sigma, y);
    if (rc == ABZ_RETRY_PIVOTED) {
        ctx->force_generic = true;
        rc = abz_points_resolvent_impl(ctx, sid, npts, k, fkind, nw, z, sigma, y);
        ctx->force_generic = false;
    }
    return rc;
}


// ---- IAI nest arena -------------------------------------------------------------------------------
int32_t abz_nest_create(abz_ctx* ctx, abz_series_t sid, int32_t ndim, int64_t cap2, int64_t cap1, abz_nest_t* out) {
    if (!ctx) return ABZ_E_INVALID;
    Series* s = get_series(ctx, sid);
    if (!s) return fail(ctx, ABZ_E_INVALID, "unknown series handle");
    if (!out || ndim < 1 || ndim > 3 || cap2 < 0 || cap1 < 0) return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    for (int d = ndim; d < 3; d++)
        if (s->M[d] != 1) return fail(ctx, ABZ_E_INVALID, "variables in Fourier series don't match domain");
    cudaSetDevice(ctx->device);
    auto nst = std::make_unique<Nest>();
    nst->ctx = ctx;
    nst->s = s; nst->keep = ctx->series[sid]; nst->ndim = ndim; nst->cap2 = cap2; nst->cap1 = cap1;
    const size_t nn = (size_t)s->n * s->n;
    if (ndim == 3 && cap2 > 0) CU(ctx, pool_alloc(ctx, (void**)&nst->L2, (size_t)cap2 * nn * s->M[0] * s->M[1] * sizeof(double2)));
    if (ndim >= 2 && cap1 > 0) CU(ctx, pool_alloc(ctx, (void**)&nst->L1, (size_t)cap1 * nn * s->M[0] * sizeof(double2)));
    uint64_t id = ctx->next_id++;
    ctx->nests[id] = std::move(nst);
    *out = id;
    return ABZ_OK;
}

int32_t abz_nest_destroy(abz_ctx* ctx, abz_nest_t nest) {
    if (!ctx) return ABZ_E_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    return ctx->nests.erase(nest) ? ABZ_OK : fail(ctx, ABZ_E_INVALID, "unknown nest handle");
}

static int check_slots(abz_ctx* ctx, const int64_t* slot, int64_t n, int64_t cap, const char* what) {
    for (int64_t i = 0; i < n; i++)
        if (slot[i] < 0 || slot[i] >= cap) return fail(ctx, ABZ_E_INVALID, std::string(what) + ": slot index out of range");
    return ABZ_OK;
}

int32_t abz_nest_contract3(abz_ctx* ctx, abz_nest_t nid, int64_t n, const double* x3, const int64_t* slot2) {
    if (!ctx) return ABZ_E_INVALID;
    Nest* nst = get_nest(ctx, nid);
    if (!nst) return fail(ctx, ABZ_E_INVALID, "unknown nest handle");
    if (nst->ndim != 3) return fail(ctx, ABZ_E_INVALID, "abz_nest_contract3 needs a 3-d nest");
    if (n < 0 || (n > 0 && (!x3 || !slot2))) return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    if (n == 0) return ABZ_OK;
    int rc = check_slots(ctx, slot2, n, nst->cap2, "abz_nest_contract3");
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    Series* s = nst->s;
    const long rows = (long)s->n * s->n * s->M[0] * s->M[1];
    CU(ctx, ctx->tmp_a.reserve((size_t)n * sizeof(double)));
    CU(ctx, ctx->tmp_b.reserve((size_t)n * sizeof(long)));
    CU(ctx, cudaMemcpyAsync(ctx->tmp_a.p, x3, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->tmp_b.p, slot2, (size_t)n * sizeof(long), cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((unsigned)((rows + 255) / 256), (unsigned)n);
    nest_contract_kernel<<<grid, 256, (size_t)s->M[2] * sizeof(double2), ctx->stream>>>(
        s->c, 0, nullptr, ctx->tmp_a.as<double>(), ctx->tmp_b.as<long>(), nst->L2, rows, s->M[2], s->lo[2], s->period[2]);
    LAUNCH_CHECK(ctx, "nest_contract_kernel");
    return ABZ_OK;
}

int32_t abz_nest_contract2(abz_ctx* ctx, abz_nest_t nid, int64_t n, const double* x2, const int64_t* parent, const int64_t* slot1) {
    if (!ctx) return ABZ_E_INVALID;
    Nest* nst = get_nest(ctx, nid);
    if (!nst) return fail(ctx, ABZ_E_INVALID, "unknown nest handle");
    if (nst->ndim < 2) return fail(ctx, ABZ_E_INVALID, "abz_nest_contract2 needs a nest with ndim >= 2");
    if (n < 0 || (n > 0 && (!x2 || !slot1 || (nst->ndim == 3 && !parent)))) return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    if (n == 0) return ABZ_OK;
    int rc = check_slots(ctx, slot1, n, nst->cap1, "abz_nest_contract2");
    if (rc) return rc;
    if (nst->ndim == 3 && (rc = check_slots(ctx, parent, n, nst->cap2, "abz_nest_contract2(parent)"))) return rc;
    cudaSetDevice(ctx->device);
    Series* s = nst->s;
    const long rows = (long)s->n * s->n * s->M[0];
    CU(ctx, ctx->tmp_a.reserve((size_t)n * sizeof(double)));
    CU(ctx, ctx->tmp_b.reserve((size_t)n * sizeof(long)));
    CU(ctx, ctx->tmp_c.reserve((size_t)n * sizeof(long)));
    CU(ctx, cudaMemcpyAsync(ctx->tmp_a.p, x2, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->tmp_b.p, slot1, (size_t)n * sizeof(long), cudaMemcpyHostToDevice, ctx->stream));
    const double2* src = s->c; const long* par = nullptr; long stride = 0;
    if (nst->ndim == 3) {
        CU(ctx, cudaMemcpyAsync(ctx->tmp_c.p, parent, (size_t)n * sizeof(long), cudaMemcpyHostToDevice, ctx->stream));
        src = nst->L2; par = ctx->tmp_c.as<long>(); stride = rows * s->M[1];
    }
    dim3 grid((unsigned)((rows + 255) / 256), (unsigned)n);
    nest_contract_kernel<<<grid, 256, (size_t)s->M[1] * sizeof(double2), ctx->stream>>>(
        src, stride, par, ctx->tmp_a.as<double>(), ctx->tmp_b.as<long>(), nst->L1, rows, s->M[1], s->lo[1], s->period[1]);
    LAUNCH_CHECK(ctx, "nest_contract_kernel");
    return ABZ_OK;
}

static int32_t abz_nest_eval_impl(abz_ctx* ctx, abz_nest_t nid, int64_t npts, const double* x1, const int64_t* slot1, int32_t fkind,
                      const double* z, const double* sigma, double* y) {
    if (!ctx) return ABZ_E_INVALID;
    Nest* nst = get_nest(ctx, nid);
    if (!nst) return fail(ctx, ABZ_E_INVALID, "unknown nest handle");
    if (npts < 0 || (npts > 0 && (!x1 || !y || (nst->ndim >= 2 && !slot1))) || (fkind == ABZ_F_RESOLVENT_TRACE && !z))
        return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    if (npts == 0) return ABZ_OK;
    int rc;
    if (nst->ndim >= 2 && (rc = check_slots(ctx, slot1, npts, nst->cap1, "abz_nest_eval"))) return rc;
    cudaSetDevice(ctx->device);
    Series* s = nst->s;
    const int n = s->n;
    const long nn = (long)n * n;
    CU(ctx, ctx->tmp_a.reserve((size_t)npts * sizeof(double)));
    CU(ctx, ctx->tmp_b.reserve((size_t)npts * sizeof(long)));
    CU(ctx, ctx->tmp_c.reserve((size_t)npts * sizeof(double2)));
    CU(ctx, cudaMemcpyAsync(ctx->tmp_a.p, x1, (size_t)npts * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const long* dslot = nullptr; const double2* L1 = s->c; long stride = 0;
    if (nst->ndim >= 2) {
        CU(ctx, cudaMemcpyAsync(ctx->tmp_b.p, slot1, (size_t)npts * sizeof(long), cudaMemcpyHostToDevice, ctx->stream));
        dslot = ctx->tmp_b.as<long>(); L1 = nst->L1; stride = nn * s->M[0];
    }
    rc = upload_params(ctx, n, 1, fkind == ABZ_F_RESOLVENT_TRACE ? z : nullptr, sigma);
    if (rc) return rc;
    double2 zz = make_double2(z ? z[0] : 0.0, z ? z[1] : 0.0);
    const double2* dsig = sigma ? ctx->sigbuf.as<double2>() : nullptr;
    double2* yd = ctx->tmp_c.as<double2>();
    int* ef = ctx->errflag.as<int>();
    if (n <= IAI_SMALL_MAXN) {
        unsigned g = (unsigned)((npts + 127) / 128);
#define NEST_LAUNCH(NORB) \
    nest_eval_small_kernel<NORB><<<g, 128, 0, ctx->stream>>>(L1, stride, dslot, ctx->tmp_a.as<double>(), npts, s->M[0], s->lo[0], \
                                                            s->period[0], fkind, zz, dsig, yd, ef)
        switch (s->n) { case 1: NEST_LAUNCH(1); break; case 2: NEST_LAUNCH(2); break; case 3: NEST_LAUNCH(3); break; case 4: NEST_LAUNCH(4); break; case 5: NEST_LAUNCH(5); break; default: NEST_LAUNCH(6); break; }
#undef NEST_LAUNCH
        LAUNCH_CHECK(ctx, "nest_eval_small_kernel");
    } else {
        CU(ctx, ctx->Hc.reserve((size_t)npts * nn * sizeof(double2)));
        dim3 grid((unsigned)((nn + 127) / 128), (unsigned)npts);
        nest_eval_h_kernel<<<grid, 128, (size_t)s->M[0] * sizeof(double2), ctx->stream>>>(
            L1, stride, dslot, ctx->tmp_a.as<double>(), (int)nn, s->M[0], s->lo[0], s->period[0], ctx->Hc.as<double2>());
        LAUNCH_CHECK(ctx, "nest_eval_h_kernel");
        rc = run_matfun(ctx, ctx->Hc.as<double2>(), nullptr, npts, n, fkind, 1, ctx->zbuf.as<double2>(), dsig, 1, yd);
        if (rc) return rc;
    }
    CU(ctx, cudaMemcpyAsync(y, yd, (size_t)npts * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    return check_errflag(ctx, "abz_nest_eval");
}

int32_t abz_nest_eval(abz_ctx* ctx, abz_nest_t nid, int64_t npts, const double* x1, const int64_t* slot1, int32_t fkind,
                      const double* z, const double* sigma, double* y) {
    int32_t rc = abz_nest_eval_impl(ctx, nid, npts, x1, slot1, fkind, z, sigma, y);
    if (rc == ABZ_RETRY_PIVOTED) {
        ctx->force_generic = true;
        rc = abz_nest_eval_impl(ctx, nid, npts, x1, slot1, fkind, z, sigma, y);
        ctx->force_generic = false;
    }
    return rc;
}


int32_t abz_nest_eval_h(abz_ctx* ctx, abz_nest_t nid, int64_t npts, const double* x1, const int64_t* slot1, double* Hk) {
    if (!ctx) return ABZ_E_INVALID;
    Nest* nst = get_nest(ctx, nid);
    if (!nst) return fail(ctx, ABZ_E_INVALID, "unknown nest handle");
    if (npts < 0 || (npts > 0 && (!x1 || !Hk || (nst->ndim >= 2 && !slot1)))) return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    if (npts == 0) return ABZ_OK;
    int rc;
    if (nst->ndim >= 2 && (rc = check_slots(ctx, slot1, npts, nst->cap1, "abz_nest_eval_h"))) return rc;
    cudaSetDevice(ctx->device);
    Series* s = nst->s;
    const long nn = (long)s->n * s->n;
    CU(ctx, ctx->tmp_a.reserve((size_t)npts * sizeof(double)));
    CU(ctx, ctx->tmp_b.reserve((size_t)npts * sizeof(long)));
    CU(ctx, ctx->Hc.reserve((size_t)npts * nn * sizeof(double2)));
    CU(ctx, cudaMemcpyAsync(ctx->tmp_a.p, x1, (size_t)npts * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const long* dslot = nullptr; const double2* L1 = s->c; long stride = 0;
    if (nst->ndim >= 2) {
        CU(ctx, cudaMemcpyAsync(ctx->tmp_b.p, slot1, (size_t)npts * sizeof(long), cudaMemcpyHostToDevice, ctx->stream));
        dslot = ctx->tmp_b.as<long>(); L1 = nst->L1; stride = nn * s->M[0];
    }
    dim3 grid((unsigned)((nn + 127) / 128), (unsigned)npts);
    nest_eval_h_kernel<<<grid, 128, (size_t)s->M[0] * sizeof(double2), ctx->stream>>>(L1, stride, dslot, ctx->tmp_a.as<double>(), (int)nn,
                                                                                      s->M[0], s->lo[0], s->period[0], ctx->Hc.as<double2>());
    LAUNCH_CHECK(ctx, "nest_eval_h_kernel");
    CU(ctx, cudaMemcpyAsync(Hk, ctx->Hc.p, (size_t)npts * nn * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ABZ_OK;
}


// matrix-valued innermost closure: Y[:, :, i] = (z - H(x1_i on slot1_i) - Sigma)^-1 (gloc_integrand under IAI, docs/src/examples.md:90-106)
int32_t abz_nest_eval_matrix(abz_ctx* ctx, abz_nest_t nid, int64_t npts, const double* x1, const int64_t* slot1, const double* z,
                             const double* sigma, double* Y) {
    if (!ctx) return ABZ_E_INVALID;
    Nest* nst = get_nest(ctx, nid);
    if (!nst) return fail(ctx, ABZ_E_INVALID, "unknown nest handle");
    if (npts < 0 || !z || (npts > 0 && (!x1 || !Y || (nst->ndim >= 2 && !slot1)))) return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    if (npts == 0) return ABZ_OK;
    int rc;
    if (nst->ndim >= 2 && (rc = check_slots(ctx, slot1, npts, nst->cap1, "abz_nest_eval_matrix"))) return rc;
    cudaSetDevice(ctx->device);
    Series* s = nst->s;
    const int n = s->n;
    const long nn = (long)n * n;
    CU(ctx, ctx->tmp_a.reserve((size_t)npts * sizeof(double)));
    CU(ctx, ctx->tmp_b.reserve((size_t)npts * sizeof(long)));
    CU(ctx, ctx->Hc.reserve((size_t)npts * nn * sizeof(double2)));
    CU(ctx, ctx->partial.reserve((size_t)npts * nn * sizeof(double2)));
    CU(ctx, cudaMemcpyAsync(ctx->tmp_a.p, x1, (size_t)npts * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const long* dslot = nullptr; const double2* L1 = s->c; long stride = 0;
    if (nst->ndim >= 2) {
        CU(ctx, cudaMemcpyAsync(ctx->tmp_b.p, slot1, (size_t)npts * sizeof(long), cudaMemcpyHostToDevice, ctx->stream));
        dslot = ctx->tmp_b.as<long>(); L1 = nst->L1; stride = nn * s->M[0];
    }
    rc = upload_params(ctx, n, 1, z, sigma);
    if (rc) return rc;
    dim3 gh((unsigned)((nn + 127) / 128), (unsigned)npts);
    nest_eval_h_kernel<<<gh, 128, (size_t)s->M[0] * sizeof(double2), ctx->stream>>>(L1, stride, dslot, ctx->tmp_a.as<double>(), (int)nn,
                                                                                   s->M[0], s->lo[0], s->period[0], ctx->Hc.as<double2>());
    LAUNCH_CHECK(ctx, "nest_eval_h_kernel");
    // one node per team and kper = 1: the matrix-sum kernels then leave every node's inverse in its own partial block
    const double2* sgd = sigma ? ctx->sigbuf.as<double2>() : nullptr;
    dim3 grid((unsigned)npts, 1);
    if (n <= 20) {
        const size_t smem = ((size_t)n * (n + 1) + (n + 1) / 2 + 1 + nn) * sizeof(double2);
        resolvent_gj_matrix_kernel<<<grid, 32, smem, ctx->stream>>>(ctx->Hc.as<double2>(), nullptr, npts, n, 1, ctx->zbuf.as<double2>(), sgd, 1,
                                                                   ctx->partial.as<double2>(), ctx->errflag.as<int>());
    } else {
        GJ_DISPATCH(resolvent_gjreg_matrix_kernel, true, n, grid, (size_t)nn * sizeof(double2), ctx->stream, ctx->Hc.as<double2>(), nullptr, npts, n, 1,
                    ctx->zbuf.as<double2>(), sgd, 1, ctx->partial.as<double2>(), ctx->errflag.as<int>());
    }
    LAUNCH_CHECK(ctx, "resolvent_gj_matrix_kernel");
    CU(ctx, cudaMemcpyAsync(Y, ctx->partial.p, (size_t)npts * nn * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    rc = check_errflag(ctx, "abz_nest_eval_matrix");
    return rc == ABZ_RETRY_PIVOTED ? ABZ_OK : rc;
}

// ---- IAI: the whole nested adaptive solve, control flow on the library's host side -------------------
namespace {

int pin_reserve(abz_ctx* ctx, void** p, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return ABZ_OK;
    if (*p) { cudaFreeHost(*p); *p = nullptr; *cap = 0; }
    size_t want = std::max<size_t>(bytes * 2, (size_t)1 << 16);
    CU(ctx, cudaHostAlloc(p, want, cudaHostAllocDefault));
    *cap = want;
    return ABZ_OK;
}

struct IaiDeviceBackend {
    abz_ctx* ctx; Nest* nst; int fkind, vkind; double2 z; const double2* dsig; abz_iai::cplx la, lb;
    double rtol; int64_t maxevals;
    abz_exchange_fn xfn = nullptr; void* xuser = nullptr;
    bool heap_overflow = false;

    // in-place sum over ranks: the caller's hook (e.g. a host-language MPI / torch.distributed allreduce) or NCCL on the ctx communicator
    int exchange(double* buf, size_t n) {
        if (xfn) {
            if (xfn(buf, (int64_t)n, xuser) != 0) return fail(ctx, ABZ_E_NCCL, "abz_iai_solve_sharded: the exchange callback failed");
            return ABZ_OK;
        }
        return abz_allreduce_sum(ctx, buf, (int64_t)n);
    }

    // ---- lanes (several rounds in flight, one stream each): the fused small-orbital kernels only; the general path below
    //      shares the context's work buffers and runs one round at a time
    int nlanes = 1;
    int lanes() const { return nlanes; }
    int submit(int g, abz_iai::Round& R) {
        if (nlanes == 1) return ABZ_OK;
        IaiLane& ln = *ctx->iai_lanes[g];
        const int rc = enqueue(ln, R);
        if (rc) cudaStreamSynchronize(ln.stream);      // a failed submit is not in the engine's fifo: what it queued must finish before the caller unwinds
        return rc;
    }
    int wait(int g, abz_iai::Round& R) { return nlanes == 1 ? run_round(R) : collect(*ctx->iai_lanes[g], R); }

    int init_lanes(int want) {
        nlanes = want < 1 ? 1 : want;
        if (nlanes == 1) return ABZ_OK;
        while ((int)ctx->iai_lanes.size() < nlanes) {
            std::unique_ptr<IaiLane> ln(new IaiLane());
            CU(ctx, cudaStreamCreateWithFlags(&ln->stream, cudaStreamNonBlocking));
            ctx->iai_lanes.push_back(std::move(ln));
        }
        CU(ctx, cudaStreamSynchronize(ctx->stream));      // parameters uploaded on the context's stream are complete
        return ABZ_OK;
    }

    int enqueue(IaiLane& ln, abz_iai::Round& R) {
        Series* s = nst->s;
        const int n = s->n;
        const long nn = (long)n * n;
        const size_t n3 = R.c3_x.size(), n2 = R.c2_x.size(), ns = R.seg_a.size(), nt = R.task_a.size(), nm = R.mid_a.size();
        if (n3 > 0 && nst->ndim != 3) return fail(ctx, ABZ_E_INVALID, "internal: level-2 contraction on a nest with ndim < 3");
        cudaStream_t st = ln.stream;
        const size_t words = 2 * n3 + 3 * n2 + 3 * ns + 4 * nt + 4 * nm;
        int rc = pin_reserve(ctx, &ln.pin_in, &ln.pin_in_cap, words * 8);
        if (rc) return rc;
        CU(ctx, ln.in.reserve(words * 8 + 8));
        char* h = (char*)ln.pin_in;
        size_t off = 0;
        auto put = [&](const void* src, size_t cnt) { size_t o = off; if (cnt) memcpy(h + 8 * off, src, 8 * cnt); off += cnt; return o; };
        const size_t o_c3x = put(R.c3_x.data(), n3), o_c3s = put(R.c3_slot.data(), n3);
        const size_t o_c2x = put(R.c2_x.data(), n2), o_c2p = put(R.c2_parent.data(), n2), o_c2s = put(R.c2_slot.data(), n2);
        const size_t o_sa = put(R.seg_a.data(), ns), o_sb = put(R.seg_b.data(), ns), o_ss = put(R.seg_slot.data(), ns);
        const size_t o_ta = put(R.task_a.data(), nt), o_tb = put(R.task_b.data(), nt), o_tt = put(R.task_atol.data(), nt),
                     o_ts = put(R.task_slot.data(), nt);
        const size_t o_ma = put(R.mid_a.data(), nm), o_mb = put(R.mid_b.data(), nm), o_mt = put(R.mid_atol.data(), nm),
                     o_ms = put(R.mid_slot.data(), nm);
        if (words) CU(ctx, cudaMemcpyAsync(ln.in.p, h, words * 8, cudaMemcpyHostToDevice, st));
        const double* dD = ln.in.as<double>();
        const long* dL = ln.in.as<long>();
        if (n3) {
            const long rows = nn * s->M[0] * s->M[1];
            dim3 grid((unsigned)((rows + 255) / 256), (unsigned)n3);
            nest_contract_kernel<<<grid, 256, (size_t)s->M[2] * sizeof(double2), st>>>(
                s->c, 0, nullptr, dD + o_c3x, dL + o_c3s, nst->L2, rows, s->M[2], s->lo[2], s->period[2]);
            LAUNCH_CHECK(ctx, "nest_contract_kernel");
        }
        if (n2) {
            const long rows = nn * s->M[0];
            const bool root = (nst->ndim == 2);
            dim3 grid((unsigned)((rows + 255) / 256), (unsigned)n2);
            nest_contract_kernel<<<grid, 256, (size_t)s->M[1] * sizeof(double2), st>>>(
                root ? s->c : nst->L2, root ? 0 : rows * s->M[1], root ? nullptr : dL + o_c2p, dD + o_c2x, dL + o_c2s, nst->L1, rows,
                s->M[1], s->lo[1], s->period[1]);
            LAUNCH_CHECK(ctx, "nest_contract_kernel");
        }
        const size_t owords = 4 * ns + 4 * nt + 5 * nm + 1;
        rc = pin_reserve(ctx, &ln.pin_out, &ln.pin_out_cap, owords * 8);
        if (rc) return rc;
        CU(ctx, ln.out.reserve(owords * 8));
        double* dout = ln.out.as<double>();
        int* ef = ctx->errflag.as<int>();
        const bool has_slots = (nst->ndim >= 2);
        const double2* L1 = has_slots ? nst->L1 : s->c;
        const long stride = has_slots ? nn * s->M[0] : 0;
        if (ns) {
            const long* sslot = has_slots ? dL + o_ss : nullptr;
            unsigned g = (unsigned)((ns + 7) / 8);
            if ((size_t)8 * s->M[0] * n * n * sizeof(double2) > 160 * 1024)
                return fail(ctx, ABZ_E_UNSUPPORTED, "series has too many coefficients per dimension for the IAI panel kernel");
#define PANEL_LAUNCH(NORB)                                                                                               \
    nest_panel_small_kernel<NORB><<<g, 128, (size_t)8 * s->M[0] * NORB * NORB * sizeof(double2), st>>>(L1, stride, dD + o_sa, dD + o_sb, sslot, (long)ns, s->M[0], s->lo[0], \
                                                             s->period[0], fkind, vkind, z, dsig, la, lb, dout, ef)
            switch (s->n) { case 1: PANEL_LAUNCH(1); break; case 2: PANEL_LAUNCH(2); break; case 3: PANEL_LAUNCH(3); break; case 4: PANEL_LAUNCH(4); break; case 5: PANEL_LAUNCH(5); break; default: PANEL_LAUNCH(6); break; }
#undef PANEL_LAUNCH
            LAUNCH_CHECK(ctx, "nest_panel_small_kernel");
        }
        if (nt) {
            if (!has_slots) return fail(ctx, ABZ_E_UNSUPPORTED, "device-side innermost integrals need ndim >= 2");
            rc = launch_leaf(dD + o_ta, dD + o_tb, dD + o_tt, dL + o_ts, (long)nt, dout + 4 * ns, st, ln.spill);
            if (rc) return rc;
        }
        if (nm) {
            rc = launch_mid(dD + o_ma, dD + o_mb, dD + o_mt, dL + o_ms, (long)nm, dout + 4 * ns + 4 * nt, st, ln.mid_spill);
            if (rc) return rc;
        }
        CU(ctx, cudaMemcpyAsync(dout + 4 * ns + 4 * nt + 5 * nm, ef, sizeof(int), cudaMemcpyDeviceToDevice, st));
        CU(ctx, cudaMemcpyAsync(ln.pin_out, dout, owords * 8, cudaMemcpyDeviceToHost, st));
        ln.ns = ns; ln.nt = nt; ln.nm = nm;
        return ABZ_OK;
    }

    int collect(IaiLane& ln, abz_iai::Round& R) {
        CU(ctx, cudaStreamSynchronize(ln.stream));
        return unpack((const double*)ln.pin_out, ln.ns, ln.nt, ln.nm, R);
    }

    int unpack(const double* ho, size_t ns, size_t nt, size_t nm, abz_iai::Round& R) {
        int flag = 0;
        memcpy(&flag, ho + 4 * ns + 4 * nt + 5 * nm, sizeof(int));
        if (flag) {
            cudaMemsetAsync(ctx->errflag.as<int>(), 0, sizeof(int), ctx->stream);
            if ((flag & 2) && !(flag & 1) && !ctx->force_generic) return ABZ_RETRY_PIVOTED;
            if (flag & 4) { heap_overflow = true; return fail(ctx, ABZ_E_UNSUPPORTED, "abz_iai_solve: an innermost integral outgrew the device segment heap"); }
            return fail(ctx, ABZ_E_SINGULAR, "abz_iai_solve: singular matrix or NaN/Inf in the integrand");
        }
        R.seg_I.resize(ns); R.seg_D.resize(ns);
        for (size_t i = 0; i < ns; i++) {
            R.seg_I[i] = abz_iai::cplx{ho[4 * i], ho[4 * i + 1]};
            R.seg_D[i] = abz_iai::cplx{ho[4 * i + 2], ho[4 * i + 3]};
        }
        R.task_I.resize(nt); R.task_E.resize(nt); R.task_ne.resize(nt);
        const double* ht = ho + 4 * ns;
        for (size_t i = 0; i < nt; i++) {
            R.task_I[i] = abz_iai::cplx{ht[4 * i], ht[4 * i + 1]};
            R.task_E[i] = ht[4 * i + 2];
            int64_t ne; memcpy(&ne, ht + 4 * i + 3, 8);
            R.task_ne[i] = ne;
        }
        R.mid_I.resize(nm); R.mid_E.resize(nm); R.mid_ne.resize(nm);
        const double* hm = ho + 4 * ns + 4 * nt;
        for (size_t i = 0; i < nm; i++) {
            R.mid_I[i] = abz_iai::cplx{hm[5 * i], hm[5 * i + 1]};
            R.mid_E[i] = hm[5 * i + 2];
            int64_t ne; memcpy(&ne, hm + 5 * i + 3, 8);
            R.mid_ne[i] = ne;
        }
        return ABZ_OK;
    }

    int run_round(abz_iai::Round& R) {
        int rc = run_once(R);
        if (rc == ABZ_RETRY_PIVOTED) {       // the unpivoted fast resolvent saw a tiny pivot: redo the round pivoted
            ctx->force_generic = true;
            rc = run_once(R);
        }
        return rc;
    }

    int run_once(abz_iai::Round& R) {
        Series* s = nst->s;
        const int n = s->n;
        const long nn = (long)n * n;
        const size_t n3 = R.c3_x.size(), n2 = R.c2_x.size(), ns = R.seg_a.size(), nt = R.task_a.size(), nm = R.mid_a.size();
        if (n3 > 0 && nst->ndim != 3) return fail(ctx, ABZ_E_INVALID, "internal: level-2 contraction on a nest with ndim < 3");
        // ---- pack [c3_x | c3_slot | c2_x | c2_parent | c2_slot | seg_a | seg_b | seg_slot | task_a | task_b | task_atol | task_slot | mid_*]
        const size_t words = 2 * n3 + 3 * n2 + 3 * ns + 4 * nt + 4 * nm;
        int rc = pin_reserve(ctx, &ctx->pin_in, &ctx->pin_in_cap, words * 8);
        if (rc) return rc;
        CU(ctx, ctx->iai_in.reserve(words * 8 + 8));
        char* h = (char*)ctx->pin_in;
        size_t off = 0;
        auto put = [&](const void* src, size_t cnt) { size_t o = off; if (cnt) memcpy(h + 8 * off, src, 8 * cnt); off += cnt; return o; };
        const size_t o_c3x = put(R.c3_x.data(), n3), o_c3s = put(R.c3_slot.data(), n3);
        const size_t o_c2x = put(R.c2_x.data(), n2), o_c2p = put(R.c2_parent.data(), n2), o_c2s = put(R.c2_slot.data(), n2);
        const size_t o_sa = put(R.seg_a.data(), ns), o_sb = put(R.seg_b.data(), ns), o_ss = put(R.seg_slot.data(), ns);
        const size_t o_ta = put(R.task_a.data(), nt), o_tb = put(R.task_b.data(), nt), o_tt = put(R.task_atol.data(), nt),
                     o_ts = put(R.task_slot.data(), nt);
        const size_t o_ma = put(R.mid_a.data(), nm), o_mb = put(R.mid_b.data(), nm), o_mt = put(R.mid_atol.data(), nm),
                     o_ms = put(R.mid_slot.data(), nm);
        if (words) CU(ctx, cudaMemcpyAsync(ctx->iai_in.p, h, words * 8, cudaMemcpyHostToDevice, ctx->stream));
        const double* dD = ctx->iai_in.as<double>();
        const long* dL = ctx->iai_in.as<long>();
        // ---- contractions
        if (n3) {
            const long rows = nn * s->M[0] * s->M[1];
            dim3 grid((unsigned)((rows + 255) / 256), (unsigned)n3);
            nest_contract_kernel<<<grid, 256, (size_t)s->M[2] * sizeof(double2), ctx->stream>>>(
                s->c, 0, nullptr, dD + o_c3x, dL + o_c3s, nst->L2, rows, s->M[2], s->lo[2], s->period[2]);
            LAUNCH_CHECK(ctx, "nest_contract_kernel");
        }
        if (n2) {
            const long rows = nn * s->M[0];
            const bool root = (nst->ndim == 2);
            dim3 grid((unsigned)((rows + 255) / 256), (unsigned)n2);
            nest_contract_kernel<<<grid, 256, (size_t)s->M[1] * sizeof(double2), ctx->stream>>>(
                root ? s->c : nst->L2, root ? 0 : rows * s->M[1], root ? nullptr : dL + o_c2p, dD + o_c2x, dL + o_c2s, nst->L1, rows,
                s->M[1], s->lo[1], s->period[1]);
            LAUNCH_CHECK(ctx, "nest_contract_kernel");
        }
        // ---- innermost panels
        const size_t owords = 4 * ns + 4 * nt + 5 * nm + 1;
        rc = pin_reserve(ctx, &ctx->pin_out, &ctx->pin_out_cap, owords * 8);
        if (rc) return rc;
        CU(ctx, ctx->iai_out.reserve(owords * 8));
        double* dout = ctx->iai_out.as<double>();
        int* ef = ctx->errflag.as<int>();
        const bool has_slots = (nst->ndim >= 2);
        const double2* L1 = has_slots ? nst->L1 : s->c;
        const long stride = has_slots ? nn * s->M[0] : 0;
        if (ns) {
            const long* sslot = has_slots ? dL + o_ss : nullptr;
            if (n <= IAI_SMALL_MAXN && (n <= 3 || (size_t)8 * s->M[0] * n * n * sizeof(double2) <= 160 * 1024)) {
                unsigned g = (unsigned)((ns + 7) / 8);
                if ((size_t)8 * s->M[0] * n * n * sizeof(double2) > 160 * 1024)
                    return fail(ctx, ABZ_E_UNSUPPORTED, "series has too many coefficients per dimension for the IAI panel kernel");
#define PANEL_LAUNCH(NORB)                                                                                               \
    nest_panel_small_kernel<NORB><<<g, 128, (size_t)8 * s->M[0] * NORB * NORB * sizeof(double2), ctx->stream>>>(L1, stride, dD + o_sa, dD + o_sb, sslot, (long)ns, s->M[0], s->lo[0], \
                                                             s->period[0], fkind, vkind, z, dsig, la, lb, dout, ef)
                switch (s->n) { case 1: PANEL_LAUNCH(1); break; case 2: PANEL_LAUNCH(2); break; case 3: PANEL_LAUNCH(3); break; case 4: PANEL_LAUNCH(4); break; case 5: PANEL_LAUNCH(5); break; default: PANEL_LAUNCH(6); break; }
#undef PANEL_LAUNCH
                LAUNCH_CHECK(ctx, "nest_panel_small_kernel");
            } else {
                const long npts = 15 * (long)ns;
                CU(ctx, ctx->tmp_a.reserve((size_t)npts * sizeof(double)));
                CU(ctx, ctx->tmp_b.reserve((size_t)npts * sizeof(long)));
                CU(ctx, ctx->tmp_c.reserve((size_t)npts * sizeof(double2)));
                CU(ctx, ctx->Hc.reserve((size_t)npts * nn * sizeof(double2)));
                panel_nodes_kernel<<<(unsigned)((npts + 127) / 128), 128, 0, ctx->stream>>>(dD + o_sa, dD + o_sb, sslot, (long)ns,
                                                                                          ctx->tmp_a.as<double>(), ctx->tmp_b.as<long>());
                LAUNCH_CHECK(ctx, "panel_nodes_kernel");
                dim3 grid((unsigned)((nn + 127) / 128), (unsigned)npts);
                nest_eval_h_kernel<<<grid, 128, (size_t)s->M[0] * sizeof(double2), ctx->stream>>>(
                    L1, stride, has_slots ? ctx->tmp_b.as<long>() : nullptr, ctx->tmp_a.as<double>(), (int)nn, s->M[0], s->lo[0],
                    s->period[0], ctx->Hc.as<double2>());
                LAUNCH_CHECK(ctx, "nest_eval_h_kernel");
                rc = run_matfun(ctx, ctx->Hc.as<double2>(), nullptr, npts, n, fkind, 1, ctx->zbuf.as<double2>(), dsig, 1,
                                ctx->tmp_c.as<double2>());
                if (rc) return rc;
                panel_combine_kernel<<<(unsigned)((ns + 127) / 128), 128, 0, ctx->stream>>>(ctx->tmp_c.as<double2>(), dD + o_sa, dD + o_sb,
                                                                                          (long)ns, vkind, la, lb, dout);
                LAUNCH_CHECK(ctx, "panel_combine_kernel");
            }
        }
        if (nt) {
            if (n > IAI_SMALL_MAXN || !has_slots) return fail(ctx, ABZ_E_UNSUPPORTED, "device-side innermost integrals need norb <= 6 and ndim >= 2");
            rc = launch_leaf(dD + o_ta, dD + o_tb, dD + o_tt, dL + o_ts, (long)nt, dout + 4 * ns, ctx->stream, ctx->tmp_d);
            if (rc) return rc;
        }
        if (nm) {
            if (n > IAI_SMALL_MAXN) return fail(ctx, ABZ_E_UNSUPPORTED, "device-side middle integrals need norb <= 6");
            rc = launch_mid(dD + o_ma, dD + o_mb, dD + o_mt, dL + o_ms, (long)nm, dout + 4 * ns + 4 * nt, ctx->stream, ctx->tmp_c);
            if (rc) return rc;
        }
        CU(ctx, cudaMemcpyAsync(dout + 4 * ns + 4 * nt + 5 * nm, ef, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(ctx->pin_out, dout, owords * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        return unpack((const double*)ctx->pin_out, ns, nt, nm, R);
    }

    // whole middle integrals (3-d, norb <= 3): one CTA each, see iai_mid_kernel
    int lkind = 0; double lim_a[3] = {0, 0, 0}, lim_b[3] = {0, 0, 0};
    int launch_mid(const double* ma, const double* mb, const double* mt, const long* ms, long nm, double* out, cudaStream_t st,
                   DevBuf& spillbuf) {
        Series* s = nst->s;
        const long stride2 = (long)s->n * s->n * s->M[0] * s->M[1];
        const int spill_cap = ctx->leaf_spill;
        CU(ctx, spillbuf.reserve((size_t)nm * 2 * MID_WARPS * std::max(spill_cap, 1) * sizeof(LeafSeg) + 64));
        const size_t smem = sizeof(MidShared) + (size_t)MID_WARPS * ((size_t)s->M[0] * s->n * s->n + s->M[1]) * sizeof(double2);
        if (smem > 200 * 1024) return fail(ctx, ABZ_E_UNSUPPORTED, "series has too many coefficients per dimension for the IAI middle kernel");
        LeafSeg* spill = reinterpret_cast<LeafSeg*>(spillbuf.as<char>() + 64);
#define MID_LAUNCH(NORB)                                                                                               \
    iai_mid_kernel<NORB><<<(unsigned)(2 * nm), MID_WARPS * 32, smem, st>>>(nst->L2, stride2, ma, mb, mt, ms, lkind, lim_a[0], lim_b[0], lim_a[1], \
                                                                    s->M[0], s->lo[0], s->period[0], s->M[1], s->lo[1], s->period[1], fkind, vkind, z, \
                                                                    dsig, la, lb, rtol, (long long)maxevals, spill, spill_cap, out, \
                                                                    ctx->errflag.as<int>())
        switch (s->n) { case 1: MID_LAUNCH(1); break; case 2: MID_LAUNCH(2); break; case 3: MID_LAUNCH(3); break; case 4: MID_LAUNCH(4); break; case 5: MID_LAUNCH(5); break; default: MID_LAUNCH(6); break; }
#undef MID_LAUNCH
        LAUNCH_CHECK(ctx, "iai_mid_kernel");
        return ABZ_OK;
    }

    int launch_leaf(const double* ta, const double* tb, const double* tt, const long* ts, long nt, double* out, cudaStream_t st,
                    DevBuf& spillbuf) {
        Series* s = nst->s;
        const long stride = (long)s->n * s->n * s->M[0];
        // global spill area for segment heaps deeper than the shared-memory levels
        const int spill_cap = ctx->leaf_spill;
        CU(ctx, spillbuf.reserve((size_t)nt * std::max(spill_cap, 1) * sizeof(LeafSeg) + 64));
        unsigned g = (unsigned)((nt + LEAF_WARPS - 1) / LEAF_WARPS);
        if ((size_t)LEAF_WARPS * s->M[0] * s->n * s->n * sizeof(double2) > 160 * 1024)
            return fail(ctx, ABZ_E_UNSUPPORTED, "series has too many coefficients per dimension for the IAI leaf kernel");
        LeafSeg* spill = reinterpret_cast<LeafSeg*>(spillbuf.as<char>() + 64);
#define LEAF_LAUNCH(NORB)                                                                                              \
    iai_leaf_kernel<NORB><<<g, LEAF_WARPS * 32, (size_t)LEAF_WARPS * s->M[0] * NORB * NORB * sizeof(double2), st>>>(nst->L1, stride, ta, tb, tt, ts, nt, s->M[0], s->lo[0], s->period[0], \
                                                                 fkind, vkind, z, dsig, la, lb, rtol, (long long)maxevals, spill, spill_cap, out, \
                                                                 ctx->errflag.as<int>())
        switch (s->n) { case 1: LEAF_LAUNCH(1); break; case 2: LEAF_LAUNCH(2); break; case 3: LEAF_LAUNCH(3); break; case 4: LEAF_LAUNCH(4); break; case 5: LEAF_LAUNCH(5); break; default: LEAF_LAUNCH(6); break; }
#undef LEAF_LAUNCH
        LAUNCH_CHECK(ctx, "iai_leaf_kernel");
        return ABZ_OK;
    }
};

}  // namespace

int32_t abz_iai_solve(abz_ctx* ctx, abz_nest_t nid, int32_t lkind, const double* la, const double* lb, int32_t fkind,
                      int32_t vkind, const double* z, const double* sigma, const double* lin, double atol, double rtol,
                      int64_t maxevals, int32_t flags, double* out, int64_t* stats) {
    return abz_iai_solve_sharded(ctx, nid, lkind, la, lb, fkind, vkind, z, sigma, lin, atol, rtol, maxevals, flags, 0, 1, nullptr, nullptr,
                                 out, stats);
}

static int32_t iai_solve_impl(abz_ctx* ctx, abz_nest_t nid, int32_t lkind, const double* la, const double* lb, abz_limits_fn lfn,
                              void* luser, int32_t fkind, int32_t vkind, const double* z, const double* sigma, const double* lin,
                              double atol, double rtol, int64_t maxevals, int32_t flags, int32_t rank, int32_t nranks,
                              abz_exchange_fn exchange, void* exchange_user, double* out, int64_t* stats);

int32_t abz_iai_solve_sharded(abz_ctx* ctx, abz_nest_t nid, int32_t lkind, const double* la, const double* lb, int32_t fkind,
                              int32_t vkind, const double* z, const double* sigma, const double* lin, double atol, double rtol,
                              int64_t maxevals, int32_t flags, int32_t rank, int32_t nranks, abz_exchange_fn exchange,
                              void* exchange_user, double* out, int64_t* stats) {
    if (!ctx) return ABZ_E_INVALID;
    if ((lkind != 0 && lkind != 1) || !la || (lkind == 0 && !lb)) return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    return iai_solve_impl(ctx, nid, lkind, la, lb, nullptr, nullptr, fkind, vkind, z, sigma, lin, atol, rtol, maxevals, flags, rank, nranks,
                          exchange, exchange_user, out, stats);
}

int32_t abz_iai_solve_general(abz_ctx* ctx, abz_nest_t nid, abz_limits_fn limits, void* limits_user, int32_t fkind, int32_t vkind,
                              const double* z, const double* sigma, const double* lin, double atol, double rtol, int64_t maxevals,
                              int32_t flags, int32_t rank, int32_t nranks, abz_exchange_fn exchange, void* exchange_user, double* out,
                              int64_t* stats) {
    if (!ctx) return ABZ_E_INVALID;
    if (!limits) return fail(ctx, ABZ_E_INVALID, "abz_iai_solve_general: the limits callback is NULL");
    return iai_solve_impl(ctx, nid, 2, nullptr, nullptr, limits, limits_user, fkind, vkind, z, sigma, lin, atol, rtol, maxevals, flags, rank,
                          nranks, exchange, exchange_user, out, stats);
}

static int32_t iai_solve_impl(abz_ctx* ctx, abz_nest_t nid, int32_t lkind, const double* la, const double* lb, abz_limits_fn lfn,
                              void* luser, int32_t fkind, int32_t vkind, const double* z, const double* sigma, const double* lin,
                              double atol, double rtol, int64_t maxevals, int32_t flags, int32_t rank, int32_t nranks,
                              abz_exchange_fn exchange, void* exchange_user, double* out, int64_t* stats) {
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, ABZ_E_INVALID, "invalid rank/nranks");
    if (nranks > 1 && !exchange && !ctx->nccl_comm)
        return fail(ctx, ABZ_E_NCCL, "abz_iai_solve_sharded: nranks > 1 needs an exchange callback or abz_comm_init");
    Nest* nst = get_nest(ctx, nid);
    if (!nst) return fail(ctx, ABZ_E_INVALID, "unknown nest handle");
    if (!out || vkind < 0 || vkind > 2 ||
        (fkind != ABZ_F_RESOLVENT_TRACE && fkind != ABZ_F_TRACE_H) || (fkind == ABZ_F_RESOLVENT_TRACE && !z) || (vkind == 2 && !lin) ||
        !(atol >= 0) || !(rtol >= 0) || maxevals < 1)
        return fail(ctx, ABZ_E_INVALID, "invalid arguments");
    cudaSetDevice(ctx->device);
    Series* s = nst->s;
    if (fkind == ABZ_F_TRACE_H) sigma = nullptr;
    int rc = upload_params(ctx, s->n, 1, fkind == ABZ_F_RESOLVENT_TRACE ? z : nullptr, sigma);
    if (rc) return rc;
    abz_iai::Limits lims;
    lims.kind = lkind; lims.nd = nst->ndim; lims.s = 1.0;
    lims.fn = lfn; lims.user = luser;
    for (int d = 0; d < nst->ndim; d++) { lims.a[d] = la ? la[d] : 0.0; lims.b[d] = lb ? lb[d] : 0.0; }
    IaiDeviceBackend be;
    be.ctx = ctx; be.nst = nst; be.fkind = fkind; be.vkind = vkind;
    be.z = make_double2(z ? z[0] : 0.0, z ? z[1] : 0.0);
    be.dsig = sigma ? ctx->sigbuf.as<double2>() : nullptr;
    be.la = abz_iai::cplx{lin ? lin[0] : 1.0, lin ? lin[1] : 0.0};
    be.lb = abz_iai::cplx{lin ? lin[2] : 0.0, lin ? lin[3] : 0.0};
    be.rtol = rtol; be.maxevals = maxevals;
    be.xfn = exchange; be.xuser = exchange_user;
    // device-side integrals: norb <= 6 and the per-warp copies of the 1-D series must fit the kernels' shared memory (otherwise the
    // host-driven panels serve the solve: same decisions, more rounds)
    const size_t coef1 = (size_t)s->M[0] * s->n * s->n * sizeof(double2);
    const bool small_ok = s->n <= IAI_SMALL_MAXN && (size_t)8 * coef1 <= 160 * 1024;
    const bool leaf = (flags & (ABZ_IAI_DEVICE_LEAVES | ABZ_IAI_DEVICE_MIDDLES)) && small_ok && nst->ndim >= 2 &&
                      (size_t)LEAF_WARPS * coef1 <= 160 * 1024;
    // (norb = 6: the 36 complex entries of a node's matrix exceed the 128 registers of the 512-thread middle-integral CTAs - measured
    //  0.45 s against 0.20 s with device-side innermost integrals only, profiles/r02_iai_norb6_timing.log)
    const bool mid = leaf && (flags & ABZ_IAI_DEVICE_MIDDLES) && nst->ndim == 3 && lkind != 2 && s->n <= 5 &&
                     sizeof(MidShared) + (size_t)MID_WARPS * (coef1 + (size_t)s->M[1] * sizeof(double2)) <= 200 * 1024;
    be.lkind = lkind;
    for (int d = 0; d < 3; d++) { be.lim_a[d] = lims.a[d]; be.lim_b[d] = lims.b[d]; }
    const long launches0 = ctx->launches;
    {   // rounds in flight: the fused small-orbital kernels can overlap (a round uses a small part of the chip and is
        // bound by the dependent chain of its deepest innermost integral); ABZ_IAI_LANES overrides
        static const int env_lanes = getenv("ABZ_IAI_LANES") ? atoi(getenv("ABZ_IAI_LANES")) : 0;
        const int want = env_lanes > 0 ? std::min(env_lanes, 16) : ctx->iai_lanes_opt;
        rc = be.init_lanes((small_ok && nst->ndim >= 2 && nranks == 1) ? want : 1);   // sharded solves keep one round at a time
        if (rc) return rc;
    }
    int64_t rounds0 = 0;
    if ((flags & ABZ_IAI_SPECULATE) && leaf && lkind != 2) {
        // look-ahead on the outermost integral (abz_iai_engine.hpp): same decisions, fewer rounds.  A failure of this attempt may have
        // come from a panel the sequential algorithm never reaches, so it is not reported: the plain engine below decides.
        abz_iai::Engine<IaiDeviceBackend> eng0(be, nst->ndim, lims, atol, rtol, maxevals, nst->cap2, nst->cap1, leaf, rank, nranks, mid, true);
        { static const int pol = getenv("ABZ_IAI_LOOKAHEAD") ? atoi(getenv("ABZ_IAI_LOOKAHEAD")) : 3; eng0.spec_policy = (pol >= 1 && pol <= 3) ? pol : 3; }
        rc = eng0.run();
        ctx->force_generic = false;
        if (rc == 0) {
            if (stats) { stats[0] = eng0.numevals; stats[1] = eng0.rounds; stats[2] = ctx->launches - launches0; stats[3] = eng0.exchanges; }
            out[0] = eng0.result.re; out[1] = eng0.result.im; out[2] = eng0.result_err;
            return ABZ_OK;
        }
        cudaMemsetAsync(ctx->errflag.p, 0, sizeof(int), ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        be.heap_overflow = false;
        rounds0 = eng0.rounds;
    }
    abz_iai::Engine<IaiDeviceBackend> eng(be, nst->ndim, lims, atol, rtol, maxevals, nst->cap2, nst->cap1, leaf, rank, nranks, mid);
    rc = eng.run();
    ctx->force_generic = false;
    if (rc && be.nlanes > 1) {
        // the engine has drained its lanes; kernels of the other lanes may have raised the flag again after it was read
        cudaMemsetAsync(ctx->errflag.p, 0, sizeof(int), ctx->stream);
        cudaStreamSynchronize(ctx->stream);
    }
    if (rc == ABZ_E_UNSUPPORTED && leaf && be.heap_overflow && nranks == 1) {
        // an innermost integral outgrew the device segment heap (1087 segments): same solve with host-driven innermost panels,
        // whose heaps live in host memory (single rank only: ranks must not diverge in their exchange sequence)
        be.heap_overflow = false;
        abz_iai::Engine<IaiDeviceBackend> eng2(be, nst->ndim, lims, atol, rtol, maxevals, nst->cap2, nst->cap1, false, rank, nranks);
        rc = eng2.run();
        ctx->force_generic = false;
        if (rc && be.nlanes > 1) { cudaMemsetAsync(ctx->errflag.p, 0, sizeof(int), ctx->stream); cudaStreamSynchronize(ctx->stream); }
        if (stats) { stats[0] = eng2.numevals; stats[1] = rounds0 + eng.rounds + eng2.rounds; stats[2] = ctx->launches - launches0; stats[3] = 0; }
        if (rc == abz_iai::IAI_E_NAN) return fail(ctx, ABZ_E_SINGULAR, eng2.error);
        if (rc == abz_iai::IAI_E_ARENA) return fail(ctx, ABZ_E_OOM, eng2.error);
        if (rc == abz_iai::IAI_E_STALL || rc == abz_iai::IAI_E_LIMITS) return fail(ctx, ABZ_E_INVALID, eng2.error);
        if (rc) return rc;
        out[0] = eng2.result.re; out[1] = eng2.result.im; out[2] = eng2.result_err;
        return ABZ_OK;
    }
    if (stats) { stats[0] = eng.numevals; stats[1] = rounds0 + eng.rounds; stats[2] = ctx->launches - launches0; stats[3] = eng.exchanges; }
    if (rc == abz_iai::IAI_E_NAN) return fail(ctx, ABZ_E_SINGULAR, eng.error);
    if (rc == abz_iai::IAI_E_ARENA) return fail(ctx, ABZ_E_OOM, eng.error);
    if (rc == abz_iai::IAI_E_STALL || rc == abz_iai::IAI_E_LIMITS) return fail(ctx, ABZ_E_INVALID, eng.error);
    if (rc) return rc;
    out[0] = eng.result.re; out[1] = eng.result.im; out[2] = eng.result_err;
    return ABZ_OK;
}


// ---- NCCL through dlopen ----------------------------------------------------------------------------
typedef struct { char internal[128]; } abz_nccl_uid;
typedef int (*nccl_get_uid_t)(abz_nccl_uid*);
typedef int (*nccl_init_rank_t)(void**, int, abz_nccl_uid, int);
typedef int (*nccl_allreduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_destroy_t)(void*);

static void* nccl_open(std::string* why) {
    static void* lib = nullptr;
    if (lib) return lib;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (auto nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) return lib; }
    if (why) *why = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
    return nullptr;
}

int32_t abz_comm_unique_id(void* uid128) {
    std::string why;
    void* lib = nccl_open(&why);
    if (!lib) return fail(nullptr, ABZ_E_NCCL, why);
    auto f = (nccl_get_uid_t)dlsym(lib, "ncclGetUniqueId");
    if (!f || !uid128) return fail(nullptr, ABZ_E_NCCL, "ncclGetUniqueId unavailable");
    return f((abz_nccl_uid*)uid128) == 0 ? ABZ_OK : fail(nullptr, ABZ_E_NCCL, "ncclGetUniqueId failed");
}

int32_t abz_comm_init(abz_ctx* ctx, int32_t rank, int32_t nranks, const void* uid128) {
    if (!ctx) return ABZ_E_INVALID;
    if (nranks < 1 || rank < 0 || rank >= nranks || !uid128) return fail(ctx, ABZ_E_INVALID, "invalid rank/nranks");
    std::string why;
    void* lib = nccl_open(&why);
    if (!lib) return fail(ctx, ABZ_E_NCCL, why);
    auto f = (nccl_init_rank_t)dlsym(lib, "ncclCommInitRank");
    if (!f) return fail(ctx, ABZ_E_NCCL, "ncclCommInitRank unavailable");
    cudaSetDevice(ctx->device);
    abz_nccl_uid uid;
    memcpy(&uid, uid128, sizeof(uid));
    void* comm = nullptr;
    if (f(&comm, nranks, uid, rank) != 0) return fail(ctx, ABZ_E_NCCL, "ncclCommInitRank failed");
    ctx->nccl_lib = lib; ctx->nccl_comm = comm; ctx->nranks = nranks;
    return ABZ_OK;
}

int32_t abz_allreduce_sum(abz_ctx* ctx, double* host_buf, int64_t n) {
    if (!ctx) return ABZ_E_INVALID;
    if (!ctx->nccl_comm) return ctx->nranks == 1 ? ABZ_OK : fail(ctx, ABZ_E_NCCL, "communicator not initialised");
    if (n <= 0 || !host_buf) return fail(ctx, ABZ_E_INVALID, "invalid buffer");
    auto f = (nccl_allreduce_t)dlsym(ctx->nccl_lib, "ncclAllReduce");
    if (!f) return fail(ctx, ABZ_E_NCCL, "ncclAllReduce unavailable");
    cudaSetDevice(ctx->device);
    CU(ctx, ctx->tmp_a.reserve((size_t)n * sizeof(double)));
    CU(ctx, cudaMemcpyAsync(ctx->tmp_a.p, host_buf, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (f(ctx->tmp_a.p, ctx->tmp_a.p, (size_t)n, /*ncclDouble*/ 8, /*ncclSum*/ 0, ctx->nccl_comm, ctx->stream) != 0)
        return fail(ctx, ABZ_E_NCCL, "ncclAllReduce failed");
    CU(ctx, cudaMemcpyAsync(host_buf, ctx->tmp_a.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return ABZ_OK;
}

int32_t abz_comm_destroy(abz_ctx* ctx) {
    if (!ctx || !ctx->nccl_comm) return ABZ_OK;
    auto f = (nccl_destroy_t)dlsym(ctx->nccl_lib, "ncclCommDestroy");
    if (f) f(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    return ABZ_OK;
}

}  // extern "C"
