// K2: scattered-node kernels for IAI panels (src/fourier.jl:432-486): the same nested
// one-dimension-at-a-time contraction as the grid path, but at arbitrary x per work item.
// Contracted series live in an arena of slots on the device so that a level-synchronous batch
// of GK panels (15 nodes each) is three launches regardless of how many panels are live.
#pragma once
#include "abz_common.cuh"
#include "abz_kernels.cuh"
#include "abz_iai_engine.hpp"
#include <cooperative_groups.h>

namespace abz {

// out[slot[i]][row] = sum_m src_i[m][row] * exp(2 pi i x[i] (m+lo)/period)
// src_i = base + parent[i]*src_stride (parent == NULL -> the root series, shared by all items)
// grid = (row tiles of 256, items)   (workspace_contract!, src/fourier.jl:478)
__global__ void __launch_bounds__(256)
nest_contract_kernel(const double2* __restrict__ src, long src_stride, const long* __restrict__ parent,
                     const double* __restrict__ x, const long* __restrict__ slot, double2* __restrict__ dst, long rows,
                     int M, int lo, double period) {
    extern __shared__ double2 nc_ph[];   // [M]
    const long i = blockIdx.y;
    const double xi = x[i];
    for (int m = threadIdx.x; m < M; m += 256) nc_ph[m] = cis2pi(xi * (double)(m + lo) / period);
    __syncthreads();
    const long row = (long)blockIdx.x * 256 + threadIdx.x;
    if (row >= rows) return;
    const double2* s = src + (parent ? parent[i] * src_stride : 0);
    double2 acc = make_double2(0.0, 0.0);
    for (int m = 0; m < M; m++) acc = cfma(acc, s[(long)m * rows + row], nc_ph[m]);
    dst[slot[i] * rows + row] = acc;
}

// innermost closure for norb <= 3: evaluate the 1-D series c[M1][NN] at xi and apply the integrand
// (workspace_evaluate! + f.f(v, p), src/fourier.jl:452-456)
template <int NORB>
__device__ __forceinline__ double2 nest_point_small(const double2* __restrict__ c, double xi, int M1, int lo, double period,
                                                    int fkind, double2 z, const double2* __restrict__ sigma,
                                                    int* __restrict__ errflag) {
    constexpr int NN = NORB * NORB;
    double2 h[NN];
#pragma unroll
    for (int e = 0; e < NN; e++) h[e] = make_double2(0.0, 0.0);
    // phases e^{2 pi i x R / period}, R = lo .. lo + M1 - 1, by powers of w = e^{2 pi i x / period} outwards from R = 0
    // (one sincospi per node; |R| <= ~5 multiplications deep, the same few-ulp accuracy as reducing x R directly)
    const double2 w = cis2pi(xi / period);
    const int m0 = -lo;
    if (m0 >= 0 && m0 < M1) {
#pragma unroll
        for (int e = 0; e < NN; e++) h[e] = c[m0 * NN + e];
        double2 p = w;
        for (int m = m0 + 1; m < M1; m++) {
#pragma unroll
            for (int e = 0; e < NN; e++) h[e] = cfma(h[e], c[m * NN + e], p);
            p = cmul(p, w);
        }
        const double2 wc = make_double2(w.x, -w.y);
        p = wc;
        for (int m = m0 - 1; m >= 0; m--) {
#pragma unroll
            for (int e = 0; e < NN; e++) h[e] = cfma(h[e], c[m * NN + e], p);
            p = cmul(p, wc);
        }
    } else {
        double2 p = cis2pi(xi * (double)lo / period);
        for (int m = 0; m < M1; m++) {
#pragma unroll
            for (int e = 0; e < NN; e++) h[e] = cfma(h[e], c[m * NN + e], p);
            p = cmul(p, w);
        }
    }
    double2 t;
    if (fkind == 1) {
        t = make_double2(0.0, 0.0);
#pragma unroll
        for (int d = 0; d < NORB; d++) { t.x += h[d * (NORB + 1)].x; t.y += h[d * (NORB + 1)].y; }
    } else {
        t = small_resolvent_trace<NORB>(h, z, sigma);
        if (!(isfinite(t.x) && isfinite(t.y))) *errflag = 1;
    }
    return t;
}

// One thread per node.
template <int NORB>
__global__ void __launch_bounds__(128)
nest_eval_small_kernel(const double2* __restrict__ L1, long l1_stride, const long* __restrict__ slot1,
                       const double* __restrict__ x1, long npts, int M1, int lo, double period, int fkind, double2 z,
                       const double2* __restrict__ sigma, double2* __restrict__ y, int* __restrict__ errflag) {
    const long i = (long)blockIdx.x * 128 + threadIdx.x;
    if (i >= npts) return;
    const double2* c = L1 + (slot1 ? slot1[i] * l1_stride : 0);
    y[i] = nest_point_small<NORB>(c, x1[i], M1, lo, period, fkind, z, sigma, errflag);
}

// Innermost GK(7,15) panels for norb <= 3, fused: 16 lanes per panel [a,b] evaluate the 15 nodes of
// QuadGK.evalrule on the series in level-1 slot `seg_slot`, apply the integrand, and lane 0 combines them in
// evalrule's operation order (no FMA contraction: bit-identical to the host's gk_combine).
// out[4*seg] = (I.re, I.im, D.re, D.im), D = Kronrod - Gauss.
template <int NORB>
__global__ void __launch_bounds__(128)
nest_panel_small_kernel(const double2* __restrict__ L1, long l1_stride, const double* __restrict__ seg_a,
                        const double* __restrict__ seg_b, const long* __restrict__ seg_slot, long nseg, int M1, int lo,
                        double period, int fkind, int vkind, double2 z, const double2* __restrict__ sigma,
                        abz_iai::cplx la, abz_iai::cplx lb, double* __restrict__ out, int* __restrict__ errflag) {
    constexpr int NN = NORB * NORB;
    __shared__ abz_iai::cplx vals[8][16];
    extern __shared__ double2 np_coef[];              // [8][M1*NN]: the 1-D series of each panel's slot, staged once
    const int ls = threadIdx.x >> 4, j = threadIdx.x & 15;
    const long seg = (long)blockIdx.x * 8 + ls;
    double a = 0.0, b = 0.0;
    double2* cs = np_coef + ls * (M1 * NN);
    if (seg < nseg) {
        a = seg_a[seg]; b = seg_b[seg];
        const double2* c = L1 + (seg_slot ? seg_slot[seg] * l1_stride : 0);
        for (int e = j; e < M1 * NN; e += 16) cs[e] = c[e];       // coalesced 128-bit loads by the panel's 16 lanes
    }
    __syncthreads();
    if (seg < nseg && j < 15) {
        double2 y = nest_point_small<NORB>(cs, abz_iai::gk_node(a, b, j), M1, lo, period, fkind, z, sigma, errflag);
        vals[ls][j] = abz_iai::post_value(vkind, abz_iai::cplx{y.x, y.y}, la, lb);
    }
    __syncthreads();
    if (seg < nseg && j == 0) {
        abz_iai::cplx I, D;
        abz_iai::gk_combine(a, b, vals[ls], &I, &D);
        out[4 * seg] = I.re; out[4 * seg + 1] = I.im; out[4 * seg + 2] = D.re; out[4 * seg + 3] = D.im;
    }
}

// general norb: nodes of the queued panels (then nest_eval_h_kernel + the resolvent kernel + panel_combine_kernel)
__global__ void __launch_bounds__(128)
panel_nodes_kernel(const double* __restrict__ seg_a, const double* __restrict__ seg_b, const long* __restrict__ seg_slot,
                   long nseg, double* __restrict__ x, long* __restrict__ slot) {
    const long t = (long)blockIdx.x * 128 + threadIdx.x;
    if (t >= nseg * 15) return;
    const long seg = t / 15; const int j = (int)(t % 15);
    x[t] = abz_iai::gk_node(seg_a[seg], seg_b[seg], j);
    if (seg_slot) slot[t] = seg_slot[seg];
}
__global__ void __launch_bounds__(128)
panel_combine_kernel(const double2* __restrict__ y, const double* __restrict__ seg_a, const double* __restrict__ seg_b,
                     long nseg, int vkind, abz_iai::cplx la, abz_iai::cplx lb, double* __restrict__ out) {
    const long seg = (long)blockIdx.x * 128 + threadIdx.x;
    if (seg >= nseg) return;
    abz_iai::cplx f[15], I, D;
#pragma unroll
    for (int j = 0; j < 15; j++) { double2 v = y[seg * 15 + j]; f[j] = abz_iai::post_value(vkind, abz_iai::cplx{v.x, v.y}, la, lb); }
    abz_iai::gk_combine(seg_a[seg], seg_b[seg], f, &I, &D);
    out[4 * seg] = I.re; out[4 * seg + 1] = I.im; out[4 * seg + 2] = D.re; out[4 * seg + 3] = D.im;
}

// ---- device-side innermost adaptive integrals ------------------------------------------------------------
// One warp per innermost 1-D integral (QuadGK do_quadgk/adapt/refine on the series of one level-1 slot):
// lanes 0-14 / 16-30 evaluate the 15 nodes of the two halves of the popped segment, lanes 0 and 16 combine them
// (evalrule's order, no FMA contraction), lane 0 keeps the DataStructures.jl binary heap (Reverse on E): entries
// 1..63 in shared memory, deeper ones in a per-task global spill area.  Same arithmetic per node as
// nest_panel_small_kernel, hence the same accept/refine decisions as the host-driven engine.
constexpr int LEAF_WARPS = 4;
constexpr int LEAF_SMEM_SEGS = 63;
constexpr int LEAF_SPILL = 1024;
struct LeafSeg { double E, a, b, Ire, Iim; };

__device__ __forceinline__ double leaf_abs(abz_iai::cplx v) { return v.im == 0.0 ? fabs(v.re) : hypot(v.re, v.im); }

// One innermost adaptive integral by ONE warp (QuadGK do_quadgk / adapt on the 1-D series c[M1][NN] in shared memory): the body of
// iai_leaf_kernel, shared with iai_mid_kernel so that both take bit-identical decisions.  hs: the warp's 63 shared-memory heap
// entries, hg: its global spill area, vals: 32 shared complex slots of the warp.  Lane 0 returns I, E, ne (numevals); *err gets the
// flags (1: non-finite, 4: heap overflow).
template <int NORB>
__device__ __forceinline__ void leaf_integrate(const double2* __restrict__ c, LeafSeg* __restrict__ hs, LeafSeg* __restrict__ hg,
                                               int cap_total, abz_iai::cplx* __restrict__ vals, double a0, double b0, double atol,
                                               int M1, int lo, double period, int fkind, int vkind, double2 z,
                                               const double2* __restrict__ sigma, abz_iai::cplx la, abz_iai::cplx lb, double rtol,
                                               long long maxevals, int* __restrict__ errflag, abz_iai::cplx* Iout, double* Eout,
                                               long long* neout) {
    const int lane = threadIdx.x & 31;
#define LEAF_AT(i) (((i) <= LEAF_SMEM_SEGS) ? hs[(i) - 1] : hg[(i) - 1 - LEAF_SMEM_SEGS])
    const int half = lane >> 4, j = lane & 15;
    // ---- first panel
    double pa = a0, pb = b0;
    if (half == 0 && j < 15) {
        double2 y = nest_point_small<NORB>(c, abz_iai::gk_node(pa, pb, j), M1, lo, period, fkind, z, sigma, errflag);
        vals[lane] = abz_iai::post_value(vkind, abz_iai::cplx{y.x, y.y}, la, lb);
    }
    __syncwarp();
    abz_iai::cplx I{0.0, 0.0};
    double E = 0.0;
    long len = 1;
    long long ne = 15;
    int go = 0;
    if (lane == 0) {
        abz_iai::cplx D;
        abz_iai::gk_combine(pa, pb, vals, &I, &D);
        E = leaf_abs(D);
        hs[0] = LeafSeg{E, pa, pb, I.re, I.im};
        if (!isfinite(E)) { atomicOr(errflag, 1); go = 0; }
        else go = !(ne >= maxevals || E <= atol || E <= rtol * leaf_abs(I));
    }
    go = __shfl_sync(0xffffffffu, go, 0);
    while (go) {
        // ---- pop the segment with the largest error (lane 0), bisect it
        double sa = 0.0, sb = 0.0, sE = 0.0;
        abz_iai::cplx sI{0.0, 0.0};
        if (lane == 0) {
            LeafSeg top = hs[0];
            sa = top.a; sb = top.b; sE = top.E; sI = abz_iai::cplx{top.Ire, top.Iim};
            LeafSeg y = LEAF_AT(len);
            len--;
            if (len > 0) {   // percolate_down(xs, 1, y)
                long i = 1;
                for (;;) {
                    long l = 2 * i;
                    if (l > len) break;
                    long r = l + 1;
                    long jj = l;
                    if (r <= len) { if (!(LEAF_AT(r).E < LEAF_AT(l).E)) jj = r; }
                    if (!(y.E < LEAF_AT(jj).E)) break;
                    LEAF_AT(i) = LEAF_AT(jj);
                    i = jj;
                }
                LEAF_AT(i) = y;
            }
        }
        sa = __shfl_sync(0xffffffffu, sa, 0);
        sb = __shfl_sync(0xffffffffu, sb, 0);
        const double mid = (sa + sb) / 2;
        pa = half ? mid : sa;
        pb = half ? sb : mid;
        __syncwarp();
        if (j < 15) {
            double2 y = nest_point_small<NORB>(c, abz_iai::gk_node(pa, pb, j), M1, lo, period, fkind, z, sigma, errflag);
            vals[lane] = abz_iai::post_value(vkind, abz_iai::cplx{y.x, y.y}, la, lb);
        }
        __syncwarp();
        abz_iai::cplx nI{0.0, 0.0};
        double nE = 0.0;
        if (j == 0) {
            abz_iai::cplx D;
            abz_iai::gk_combine(pa, pb, vals + 16 * half, &nI, &D);
            nE = leaf_abs(D);
        }
        const double I2re = __shfl_sync(0xffffffffu, nI.re, 16), I2im = __shfl_sync(0xffffffffu, nI.im, 16);
        const double E2 = __shfl_sync(0xffffffffu, nE, 16);
        if (lane == 0) {
            I = abz_iai::cplx{(I.re - sI.re) + nI.re + I2re, (I.im - sI.im) + nI.im + I2im};
            E = (E - sE) + nE + E2;
            ne += 30;
            if (!(isfinite(nE) && isfinite(E2))) { atomicOr(errflag, 1); go = 0; }
            else if (len + 2 > cap_total) { atomicOr(errflag, 4); go = 0; }
            else {
                LeafSeg sg[2] = {LeafSeg{nE, sa, mid, nI.re, nI.im}, LeafSeg{E2, mid, sb, I2re, I2im}};
#pragma unroll
                for (int t = 0; t < 2; t++) {   // push!: percolate_up(xs, len, x)
                    len++;
                    long i = len;
                    for (;;) {
                        long p = i / 2;
                        if (p < 1) break;
                        if (!(LEAF_AT(p).E < sg[t].E)) break;
                        LEAF_AT(i) = LEAF_AT(p);
                        i = p;
                    }
                    LEAF_AT(i) = sg[t];
                }
                go = (E > atol && E > rtol * leaf_abs(I) && ne < maxevals);
            }
        }
        go = __shfl_sync(0xffffffffu, go, 0);
    }
    if (lane == 0) {
        abz_iai::cplx Iv{hs[0].Ire, hs[0].Iim};
        double Ev = hs[0].E;
        for (long k = 2; k <= len; k++) { const LeafSeg sgk = LEAF_AT(k); Iv.re += sgk.Ire; Iv.im += sgk.Iim; Ev += sgk.E; }
        *Iout = Iv; *Eout = Ev; *neout = ne;
    }
#undef LEAF_AT
}

template <int NORB>
__global__ void __launch_bounds__(LEAF_WARPS * 32)
iai_leaf_kernel(const double2* __restrict__ L1, long l1_stride, const double* __restrict__ task_a,
                const double* __restrict__ task_b, const double* __restrict__ task_atol, const long* __restrict__ task_slot,
                long ntask, int M1, int lo, double period, int fkind, int vkind, double2 z, const double2* __restrict__ sigma,
                abz_iai::cplx la, abz_iai::cplx lb, double rtol, long long maxevals, LeafSeg* __restrict__ spill, int spill_cap,
                double* __restrict__ out, int* __restrict__ errflag) {
    constexpr int NN = NORB * NORB;
    __shared__ LeafSeg heap_s[LEAF_WARPS][LEAF_SMEM_SEGS];
    __shared__ abz_iai::cplx vals[LEAF_WARPS][32];
    extern __shared__ double2 leaf_coef[];            // [LEAF_WARPS][M1*NN]: this task's 1-D series, staged once per integral
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long task = (long)blockIdx.x * LEAF_WARPS + w;
    if (task >= ntask) return;
    double2* c = leaf_coef + w * (M1 * NN);
    {
        const double2* cg = L1 + task_slot[task] * l1_stride;
        for (int e = lane; e < M1 * NN; e += 32) c[e] = cg[e];        // coalesced 128-bit loads
    }
    __syncwarp();
    LeafSeg* hg = spill + task * (long)(spill_cap > 0 ? spill_cap : 1);
    const int cap_total = spill_cap < 0 ? -spill_cap : LEAF_SMEM_SEGS + spill_cap;   // negative: total capacity (test hook)
    abz_iai::cplx Iv{0.0, 0.0};
    double Ev = 0.0;
    long long ne = 0;
    leaf_integrate<NORB>(c, heap_s[w], hg, cap_total, vals[w], task_a[task], task_b[task], task_atol[task], M1, lo, period, fkind, vkind, z,
                         sigma, la, lb, rtol, maxevals, errflag, &Iv, &Ev, &ne);
    if (lane == 0) {
        out[4 * task] = Iv.re; out[4 * task + 1] = Iv.im; out[4 * task + 2] = Ev;
        reinterpret_cast<long long*>(out)[4 * task + 3] = ne;
    }
}

// ---- device-side MIDDLE integrals (3-d solves) -----------------------------------------------------------------------------
// One CLUSTER of two CTAs (2 x 16 warps: one warp for each of the 30 nodes of a refinement step) per middle-level adaptive integral
// (over x2, for the series contracted at one x3 node in a level-2 slot): CTA 0 of the pair keeps
// that integral's segment heap in shared memory and runs QuadGK's loop itself - pop the worst segment, bisect, and for each of
// the 30 new nodes x2: contract the level-2 series at x2 (workspace_contract!, src/fourier.jl:478), inner abstol = abstol / len
// (:479-480), and run the whole innermost adaptive integral (leaf_integrate, one warp per node, nodes handed out dynamically);
// then the two panels are combined in evalrule's order, pushed, and the convergence test decides.  No host round trip below the
// outermost level: a round of the host engine is one refinement step of the OUTERMOST integral.  Same arithmetic and the same
// decisions as the host-driven engine, hence identical numevals.  The two CTAs meet in two cluster barriers per step; node values
// and the step's panels travel through distributed shared memory.
//   lkind 0: x1 in [la1, lb1];  1: x1 in [0, la1 * x2 / la2]   (CubicLimits / TetrahedralLimits)
// out[5 t ..]: I.re, I.im, E, (int64) evaluations of all innermost integrals, (int64) evaluations of this integral's own nodes
constexpr int MID_WARPS = 16;
constexpr int MID_HEAP = 1023;
struct MidShared {
    LeafSeg heap[MID_HEAP];
    LeafSeg leaf_heap[MID_WARPS][LEAF_SMEM_SEGS];
    abz_iai::cplx leaf_vals[MID_WARPS][32];
    abz_iai::cplx node_vals[32];
    double pa[2], pb[2];
    int counter, ntasks, go;
    unsigned long long ne_leaves;
};

template <int NORB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MID_WARPS * 32)
iai_mid_kernel(const double2* __restrict__ L2, long l2_stride, const double* __restrict__ task_a, const double* __restrict__ task_b,
               const double* __restrict__ task_atol, const long* __restrict__ task_slot, int lkind, double la1, double lb1, double la2,
               int M1, int lo1, double period1, int M2, int lo2, double period2, int fkind, int vkind, double2 z,
               const double2* __restrict__ sigma, abz_iai::cplx la, abz_iai::cplx lb, double rtol, long long maxevals,
               LeafSeg* __restrict__ spill, int spill_cap, double* __restrict__ out, int* __restrict__ errflag) {
    constexpr int NN = NORB * NORB;
    extern __shared__ __align__(16) unsigned char mid_raw[];
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    MidShared& sm = *reinterpret_cast<MidShared*>(mid_raw);
    MidShared* sm0 = cluster.map_shared_rank(&sm, 0);                               // the pair's CTA 0 (owner of the heap)
    MidShared* sm1 = cluster.map_shared_rank(&sm, 1);
    double2* coef = reinterpret_cast<double2*>(mid_raw + sizeof(MidShared));      // [MID_WARPS][M1*NN]
    double2* phase = coef + MID_WARPS * (M1 * NN);                                 // [MID_WARPS][M2]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long task = blockIdx.x >> 1;
    const double2* src = L2 + task_slot[task] * l2_stride;                         // [M2][M1*NN]
    const double atol = task_atol[task];
    const int rows = M1 * NN;
    double2* c = coef + w * rows;
    double2* ph = phase + w * M2;
    LeafSeg* hg = spill + ((task * 2 + crank) * MID_WARPS + w) * (long)(spill_cap > 0 ? spill_cap : 1);
    const int cap_leaf = spill_cap < 0 ? -spill_cap : LEAF_SMEM_SEGS + spill_cap;
    if (threadIdx.x == 0) {
        sm.pa[0] = task_a[task]; sm.pb[0] = task_b[task];
        sm.counter = 0; sm.ntasks = 15; sm.go = 1; sm.ne_leaves = 0ull;
    }
    cluster.sync();
    abz_iai::cplx I{0.0, 0.0};           // thread 0 only
    double E = 0.0;
    long len = 0;
    long long ne_own = 0;
    LeafSeg popped{0, 0, 0, 0, 0};
    bool first = true;
    for (;;) {
        // ---- the nodes of this step's panels, handed to the warps one at a time
        for (;;) {
            int t = 0;
            if (lane == 0) t = 2 * atomicAdd(&sm.counter, 1) + crank;      // this CTA's share: every other node
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= sm.ntasks) break;
            const int p = t / 15, jn = t - 15 * p;
            const double x2 = abz_iai::gk_node(sm.pa[p], sm.pb[p], jn);
            double ca, cb;
            if (lkind == 0) { ca = la1; cb = lb1; }
            else { ca = 0.0; cb = la1 * (x2 / la2); }
            const double catol = atol / (cb - ca);
            // contract the level-2 series at x2 into this warp's 1-D series
            __syncwarp();
            for (int m = lane; m < M2; m += 32) ph[m] = cis2pi(x2 * (double)(m + lo2) / period2);
            __syncwarp();
            for (int e = lane; e < rows; e += 32) {
                double2 acc = make_double2(0.0, 0.0);
                for (int m = 0; m < M2; m++) acc = cfma(acc, src[(long)m * rows + e], ph[m]);
                c[e] = acc;
            }
            __syncwarp();
            abz_iai::cplx Iv{0.0, 0.0};
            double Ev = 0.0;
            long long ne = 0;
            leaf_integrate<NORB>(c, sm.leaf_heap[w], hg, cap_leaf, sm.leaf_vals[w], ca, cb, catol, M1, lo1, period1, fkind, vkind, z, sigma,
                                 la, lb, rtol, maxevals, errflag, &Iv, &Ev, &ne);
            if (lane == 0) {
                sm0->node_vals[16 * p + jn] = Iv;                          // distributed shared memory: CTA 0 combines
                atomicAdd(&sm.ne_leaves, (unsigned long long)ne);
                if (!isfinite(Ev)) atomicOr(errflag, 1);
            }
        }
        cluster.sync();
        // ---- combine, push, decide (thread 0 of CTA 0: the DataStructures heap with Reverse on E, as the host engine)
        if (crank == 0 && threadIdx.x == 0) {
            int go = 1;
#define MID_AT(i) sm.heap[(i) - 1]
            if (first) {
                abz_iai::cplx D;
                abz_iai::gk_combine(sm.pa[0], sm.pb[0], sm.node_vals, &I, &D);
                E = leaf_abs(D);
                MID_AT(1) = LeafSeg{E, sm.pa[0], sm.pb[0], I.re, I.im};
                len = 1; ne_own = 15;
                if (!isfinite(E)) { atomicOr(errflag, 1); go = 0; }
                else go = !(ne_own >= maxevals || E <= atol || E <= rtol * leaf_abs(I));
            } else {
                abz_iai::cplx I1, D1, I2, D2;
                abz_iai::gk_combine(sm.pa[0], sm.pb[0], sm.node_vals, &I1, &D1);
                abz_iai::gk_combine(sm.pa[1], sm.pb[1], sm.node_vals + 16, &I2, &D2);
                const double E1 = leaf_abs(D1), E2 = leaf_abs(D2);
                I = abz_iai::cplx{(I.re - popped.Ire) + I1.re + I2.re, (I.im - popped.Iim) + I1.im + I2.im};
                E = (E - popped.E) + E1 + E2;
                ne_own += 30;
                if (!(isfinite(E1) && isfinite(E2))) { atomicOr(errflag, 1); go = 0; }
                else if (len + 2 > MID_HEAP) { atomicOr(errflag, 4); go = 0; }
                else {
                    LeafSeg sg[2] = {LeafSeg{E1, sm.pa[0], sm.pb[0], I1.re, I1.im}, LeafSeg{E2, sm.pa[1], sm.pb[1], I2.re, I2.im}};
#pragma unroll
                    for (int t = 0; t < 2; t++) {
                        len++;
                        long i = len;
                        for (;;) {
                            long p = i / 2;
                            if (p < 1) break;
                            if (!(MID_AT(p).E < sg[t].E)) break;
                            MID_AT(i) = MID_AT(p);
                            i = p;
                        }
                        MID_AT(i) = sg[t];
                    }
                    go = (E > atol && E > rtol * leaf_abs(I) && ne_own < maxevals);
                }
            }
            if (go) {          // pop the worst segment and bisect it
                popped = MID_AT(1);
                LeafSeg y = MID_AT(len);
                len--;
                if (len > 0) {
                    long i = 1;
                    for (;;) {
                        long l = 2 * i;
                        if (l > len) break;
                        long r = l + 1;
                        long jj = l;
                        if (r <= len) { if (!(MID_AT(r).E < MID_AT(l).E)) jj = r; }
                        if (!(y.E < MID_AT(jj).E)) break;
                        MID_AT(i) = MID_AT(jj);
                        i = jj;
                    }
                    MID_AT(i) = y;
                }
                const double mid = (popped.a + popped.b) / 2;
                sm.pa[0] = popped.a; sm.pb[0] = mid; sm.pa[1] = mid; sm.pb[1] = popped.b;
                sm1->pa[0] = popped.a; sm1->pb[0] = mid; sm1->pa[1] = mid; sm1->pb[1] = popped.b;
                sm.ntasks = 30; sm1->ntasks = 30;
            }
            sm.counter = 0; sm1->counter = 0;
            sm.go = go; sm1->go = go;
        }
        first = false;
        cluster.sync();
        if (!sm.go) break;
    }
    if (crank == 1 && threadIdx.x == 0) atomicAdd(&sm0->ne_leaves, sm.ne_leaves);
    cluster.sync();
    if (crank == 0 && threadIdx.x == 0) {
        abz_iai::cplx Iv{MID_AT(1).Ire, MID_AT(1).Iim};
        double Ev = MID_AT(1).E;
        for (long k = 2; k <= len; k++) { Iv.re += MID_AT(k).Ire; Iv.im += MID_AT(k).Iim; Ev += MID_AT(k).E; }
        out[5 * task] = Iv.re; out[5 * task + 1] = Iv.im; out[5 * task + 2] = Ev;
        reinterpret_cast<long long*>(out)[5 * task + 3] = (long long)sm.ne_leaves;
        reinterpret_cast<long long*>(out)[5 * task + 4] = ne_own;
#undef MID_AT
    }
}

// general norb: evaluate H at the nodes into a buffer (then the generic resolvent kernel runs on it)
// grid = (ceil(n^2/128), npts)
__global__ void __launch_bounds__(128)
nest_eval_h_kernel(const double2* __restrict__ L1, long l1_stride, const long* __restrict__ slot1,
                   const double* __restrict__ x1, int nn, int M1, int lo, double period, double2* __restrict__ Hout) {
    extern __shared__ double2 ne_ph[];
    const long i = blockIdx.y;
    const double xi = x1[i];
    for (int m = threadIdx.x; m < M1; m += 128) ne_ph[m] = cis2pi(xi * (double)(m + lo) / period);
    __syncthreads();
    const int e = blockIdx.x * 128 + threadIdx.x;
    if (e >= nn) return;
    const double2* c = L1 + (slot1 ? slot1[i] * l1_stride : 0);
    double2 acc = make_double2(0.0, 0.0);
    for (int m = 0; m < M1; m++) acc = cfma(acc, c[(long)m * nn + e], ne_ph[m]);
    Hout[i * nn + e] = acc;
}

// full 3-D evaluation at scattered points (workspace_evaluate(w, x), src/fourier.jl:122,170):
// one CTA per point, phases for the three dimensions in shared memory, threads over matrix elements.
__global__ void __launch_bounds__(128)
points_eval_kernel(const double2* __restrict__ C, int nn, int M1, int M2, int M3, int lo1, int lo2, int lo3, double p1,
                   double p2, double p3, const double* __restrict__ k, double2* __restrict__ Hout) {
    extern __shared__ double2 pe_ph[];   // [M1 + M2 + M3]
    const long i = blockIdx.x;
    const double x1 = k[3 * i], x2 = k[3 * i + 1], x3 = k[3 * i + 2];
    for (int m = threadIdx.x; m < M1 + M2 + M3; m += 128) {
        double2 p;
        if (m < M1) p = cis2pi(x1 * (double)(m + lo1) / p1);
        else if (m < M1 + M2) p = cis2pi(x2 * (double)(m - M1 + lo2) / p2);
        else p = cis2pi(x3 * (double)(m - M1 - M2 + lo3) / p3);
        pe_ph[m] = p;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nn; e += 128) {
        double2 a3 = make_double2(0.0, 0.0);
        for (int m3 = 0; m3 < M3; m3++) {
            double2 a2 = make_double2(0.0, 0.0);
            for (int m2 = 0; m2 < M2; m2++) {
                double2 a1 = make_double2(0.0, 0.0);
                const double2* c = C + (((long)m3 * M2 + m2) * M1) * nn + e;
                for (int m1 = 0; m1 < M1; m1++) a1 = cfma(a1, c[(long)m1 * nn], pe_ph[m1]);
                a2 = cfma(a2, a1, pe_ph[M1 + m2]);
            }
            a3 = cfma(a3, a2, pe_ph[M1 + M2 + m3]);
        }
        Hout[i * nn + e] = a3;
    }
}

// AutoSymPTR.symptr_rule on the device (call site src/fourier.jl:271).  Orbit representative =
// the image with the smallest column-major linear index (which is the first node of the orbit met
// by the reference's column-major scan); weight = number of distinct images.
// j mod N in [0, N).  FAST: |j| < 2^22 (checked on the host): quotient from a float reciprocal, one fix-up each way.
template <bool FAST>
__device__ __forceinline__ int symptr_mod(int j, int N, float invN) {
    if (FAST) {
        int r = j - __float2int_rd((float)j * invN) * N;
        r += (r < 0) ? N : 0;
        r -= (r >= N) ? N : 0;
        return r;
    }
    int r = j % N;
    return r + ((r < 0) ? N : 0);
}
// linear index of the image of (i1, i2, i3) under the 3x3 integer matrix S (row-major)
template <bool FAST>
__device__ __forceinline__ long symptr_image(const int* __restrict__ S, int i1, int i2, int i3, int N, float invN, long NN) {
    const int j1 = symptr_mod<FAST>(S[0] * i1 + S[1] * i2 + S[2] * i3, N, invN);
    const int j2 = symptr_mod<FAST>(S[3] * i1 + S[4] * i2 + S[5] * i3, N, invN);
    const int j3 = symptr_mod<FAST>(S[6] * i1 + S[7] * i2 + S[8] * i3, N, invN);
    return (long)j3 * NN + (j1 + N * j2);
}

template <bool FAST>
__global__ void __launch_bounds__(256)
symptr_rule_kernel(int N, int nsyms, const int* __restrict__ syms, int* __restrict__ wsym, int k3_lo, int k3_stride, int nplanes) {
    extern __shared__ int sy[];
    for (int t = threadIdx.x; t < 9 * nsyms; t += 256) sy[t] = syms[t];
    __syncthreads();
    // the selected planes i3 = k3_lo + p * k3_stride, p < nplanes (one rank's share; all planes: 0, 1, N)
    const long NN = (long)N * N;
    const long lidx = (long)blockIdx.x * 256 + threadIdx.x;
    if (lidx >= NN * nplanes) return;
    const long rr = lidx % NN;
    const int i1 = (int)(rr % N), i2 = (int)(rr / N), i3 = share_plane((int)(lidx / NN), k3_lo, k3_stride);
    const long idx = (long)i3 * NN + rr;
    const float invN = 1.0f / (float)N;
    // pass 1: is this node the smallest linear index of its orbit?  (most nodes leave after a few symmetries)
    bool has_self = false;
    for (int s = 0; s < nsyms; s++) {
        const long jdx = symptr_image<FAST>(sy + 9 * s, i1, i2, i3, N, invN, NN);
        if (jdx < idx) { wsym[idx] = 0; return; }
        has_self |= (jdx == idx);
    }
    // pass 2 (irreducible nodes only, ~1/nsyms of the grid): weight = number of distinct images
    int cnt = has_self ? 0 : 1;   // identity absent from the list: the node itself still belongs to its orbit
    constexpr int CACHE = 64;
    if (nsyms <= CACHE) {
        long img[CACHE];
        for (int s = 0; s < nsyms; s++) img[s] = symptr_image<FAST>(sy + 9 * s, i1, i2, i3, N, invN, NN);
        for (int s = 0; s < nsyms; s++) {
            bool dup = false;
            for (int r = 0; r < s; r++) dup |= (img[r] == img[s]);
            cnt += dup ? 0 : 1;
        }
    } else {
        for (int s = 0; s < nsyms; s++) {
            const long jdx = symptr_image<FAST>(sy + 9 * s, i1, i2, i3, N, invN, NN);
            bool dup = false;
            for (int r = 0; r < s && !dup; r++) dup = (symptr_image<FAST>(sy + 9 * r, i1, i2, i3, N, invN, NN) == jdx);
            cnt += dup ? 0 : 1;
        }
    }
    wsym[idx] = cnt;
}

// Two-phase variant for large grids: in symptr_rule_kernel the rare irreducible lanes (1 / nsyms of the grid) run ~100x the
// work of the others and hold their whole warp, so the kernel runs at a few per cent lane utilisation.  Here
//   phase 1  every grid point tries the first few symmetries and the survivors (points not yet beaten) are compacted;
//   phase 2  the survivors try the remaining symmetries; an irreducible point gets weight nsyms / |stabiliser| (the list
//            is a group: checked on the host).
// wsym must be zero on entry; only irreducible points are written.  Lists hold 32-bit linear indices (N^3 < 2^32).
// phase 1 (in_list == NULL): grid = (ceil(N^2 / 256), planes), i3 = k3_lo + blockIdx.y * k3_stride (one rank's planes, or all of
// them with k3_lo = 0, k3_stride = 1) - 32-bit index arithmetic only;
// phase 2 (in_list != NULL): grid = ceil(count / 256).  GROUP: the symmetry list is a group (checked on the host), so the
// weight of an irreducible point is nsyms / |stabiliser| and phase 2 writes wsym itself (no phase 3).
template <bool FAST, bool GROUP>
__global__ void __launch_bounds__(256)
symptr_filter_kernel(int N, int nsyms, int s0, int s1, const int* __restrict__ syms, const unsigned* __restrict__ in_list,
                     const unsigned* __restrict__ in_count, unsigned* __restrict__ out_list, unsigned* __restrict__ out_count,
                     unsigned out_cap, int* __restrict__ overflow, int* __restrict__ wsym, int k3_lo, int k3_stride) {
    extern __shared__ int sy[];
    for (int t = threadIdx.x; t < 9 * nsyms; t += 256) sy[t] = syms[t];
    __syncthreads();
    bool keep = false;
    unsigned idx = 0;
    int i1 = 0, i2 = 0, i3 = 0;
    if (in_list) {
        const unsigned t = blockIdx.x * 256u + threadIdx.x;
        if (t < *in_count) {
            idx = in_list[t];
            const unsigned plane = (unsigned)N * (unsigned)N;
            i3 = (int)(idx / plane);
            const unsigned r = idx - (unsigned)i3 * plane;
            i2 = (int)(r / (unsigned)N); i1 = (int)(r - (unsigned)i2 * (unsigned)N);
            keep = true;
        }
    } else {
        const unsigned r = blockIdx.x * 256u + threadIdx.x;
        if (r < (unsigned)N * (unsigned)N) {
            i3 = share_plane((int)blockIdx.y, k3_lo, k3_stride);
            i2 = (int)(r / (unsigned)N); i1 = (int)(r - (unsigned)i2 * (unsigned)N);
            idx = (unsigned)i3 * (unsigned)N * (unsigned)N + r;
            keep = true;
        }
    }
    int stab = 0;
    if (keep) {
        const long NN = (long)N * N;
        const float invN = 1.0f / (float)N;
        for (int s = s0; s < s1; s++) {
            const long jdx = symptr_image<FAST>(sy + 9 * s, i1, i2, i3, N, invN, NN);
            if (jdx < (long)idx) { keep = false; break; }
            stab += (jdx == (long)idx);
        }
    }
    if (GROUP && in_list) {          // final pass of a group: irreducible points get their weight here
        if (keep) {
            // symmetries s < s0 were all >= idx in phase 1; count the ones that fix the point among them too
            const long NN = (long)N * N;
            const float invN = 1.0f / (float)N;
            for (int s = 0; s < s0; s++) stab += (symptr_image<FAST>(sy + 9 * s, i1, i2, i3, N, invN, NN) == (long)idx);
            wsym[idx] = nsyms / stab;
        }
        return;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m) {
        const int lane = threadIdx.x & 31;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(out_count, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) {
            const unsigned pos = base + __popc(m & ((1u << lane) - 1u));
            if (pos < out_cap) out_list[pos] = idx; else *overflow = 1;
        }
    }
}
// CSR construction of the symmetry-reduced rule on the device (the reference's flags arrays,
// src/fourier.jl:237-243): one warp per (k2, k3) row of the dense weight array.
// pass 1: rowcnt[p*N + i2] = #irreducible nodes in row i2 of selected plane p (i3 = k3_lo + p*k3_stride)
__global__ void __launch_bounds__(256)
sym_row_count_kernel(const int* __restrict__ wsym, int N, int k3_lo, int k3_stride, long nrows_all, int* __restrict__ rowcnt) {
    const long row = ((long)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nrows_all) return;
    const long p = row / N, i2 = row % N;
    const int* w = wsym + ((long)share_plane((int)p, k3_lo, k3_stride) * N + i2) * N;
    int c = 0;
    for (int i1 = lane; i1 < N; i1 += 32) c += (w[i1] != 0);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if (lane == 0) rowcnt[row] = c;
}
// pass 2: node_k1 / node_w of every non-empty row at its CSR offset, k1 ascending
__global__ void __launch_bounds__(256)
sym_row_fill_kernel(const int* __restrict__ wsym, int N, const int* __restrict__ row_k3, const int* __restrict__ row_k2,
                    const long* __restrict__ row_nodeptr, long nrows, int* __restrict__ node_k1, double* __restrict__ node_w) {
    const long row = ((long)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    const int* w = wsym + ((long)row_k3[row] * N + row_k2[row]) * N;
    long off = row_nodeptr[row];
    for (int base = 0; base < N; base += 32) {
        const int i1 = base + lane;
        const int v = (i1 < N) ? w[i1] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, v != 0);
        if (v != 0) {
            const long o = off + __popc(m & ((1u << lane) - 1u));
            node_k1[o] = i1;
            node_w[o] = (double)v;
        }
        off += __popc(m);
    }
}

}  // namespace abz
