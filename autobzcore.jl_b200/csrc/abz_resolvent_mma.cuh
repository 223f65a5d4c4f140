// K3-fast: tr[(z - H(k) - Sigma_w)^-1] with the whole norb x norb complex matrix resident in the
// registers of ONE warp, in the FP64 tensor-core accumulator layout, and every O(n^3) step done by
// DMMA (mma.sync.m8n8k4.f64).  One warp per (k, w) matrix, norb <= 32 (padded to a multiple of 8).
//
// Layout: the matrix is an NB x NB grid of 8x8 complex blocks; lane = 4g + q holds, of every block,
// row g, columns 2q and 2q+1 (re and im): the C/D fragment layout of DMMA.  Two facts make the
// register-only formulation cheap:
//   * a block in C layout IS a valid A operand: contracting over k' = 2q + t (t = 0,1 the two k-slabs)
//     instead of k = q + 4t is just a permutation of the summation index, so the left operand needs no
//     data movement at all;
//   * the right operand (B fragment: row 2q + t, column g) is the same permutation of the block's
//     transpose, reached with 4 64-bit shuffles per block (a perfect matching of the lane exchange
//     graph), and the same fragment also gives tr(X Y) = sum_lanes X_C[t] * Y_Bfrag[t] for free.
//
// Algorithm (block LU without inter-block pivoting + trace of the inverse from the factors):
//   for s: D_s = inv(S_s) (8x8 in-register Gauss-Jordan), L_is = A_is D_s, X_sj = D_s U_sj,
//          A_ij -= L_is U_sj;                      A = L~ U~,  U~ = diag(S)(I + X~)
//   V = U~^-1 by back substitution on blocks, M = L~^-1 by forward substitution,
//   tr A^-1 = sum_s tr D_s + sum_{i<j} tr(V_ij M_ji).
// 40 block products (320 DMMA) for norb = 32 against the 64 of a full inversion.
// Pivoting: A = z - H - Sigma with a positive-definite anti-Hermitian part (Im z > 0, causal Sigma) has
// every leading principal minor nonsingular (numerical range argument), so elimination without row
// exchanges is safe; the kernel monitors the smallest pivot against max|a_ij| and raises a flag when
// the smallest pivot falls below 2e-3 max|a_ij|, on which the host reruns the call with the pivoted Gauss-Jordan kernel.
#pragma once
#include <algorithm>
#include <cstdlib>
#include "abz_common.cuh"

namespace abz {

struct BFrag { double r[2], i[2]; };

// C-layout block (x*0: column 2q, x*1: column 2q+1 of row g) -> B fragment b[t] = X[2q + t][g]
__device__ __forceinline__ BFrag to_bfrag(double xr0, double xr1, double xi0, double xi1, int src0, int src1, bool par) {
    // round a: this lane presents slot (a ^ par); it receives the value for t = a ^ par from lane src_a
    double pr0 = par ? xr1 : xr0, pr1 = par ? xr0 : xr1;
    double pi0 = par ? xi1 : xi0, pi1 = par ? xi0 : xi1;
    double vr0 = __shfl_sync(0xffffffffu, pr0, src0), vr1 = __shfl_sync(0xffffffffu, pr1, src1);
    double vi0 = __shfl_sync(0xffffffffu, pi0, src0), vi1 = __shfl_sync(0xffffffffu, pi1, src1);
    BFrag b;
    b.r[0] = par ? vr1 : vr0; b.r[1] = par ? vr0 : vr1;
    b.i[0] = par ? vi1 : vi0; b.i[1] = par ? vi0 : vi1;
    return b;
}

// C += sgn * A * B (complex 8x8x8): A in C layout, B as fragment.  NEG = true subtracts.
template <bool NEG>
__device__ __forceinline__ void bmm(double& cr0, double& cr1, double& ci0, double& ci1, double ar0, double ar1, double ai0,
                                    double ai1, const BFrag& b) {
    const double pr0 = NEG ? -ar0 : ar0, pr1 = NEG ? -ar1 : ar1;     // +-Ar
    const double pi0 = NEG ? -ai0 : ai0, pi1 = NEG ? -ai1 : ai1;     // +-Ai
    const double mi0 = NEG ? ai0 : -ai0, mi1 = NEG ? ai1 : -ai1;     // -+Ai
    // Cr += (+-Ar) Br + (-+Ai) Bi ; Ci += (+-Ar) Bi + (+-Ai) Br
    dmma884(cr0, cr1, pr0, b.r[0]);
    dmma884(ci0, ci1, pr0, b.i[0]);
    dmma884(cr0, cr1, pr1, b.r[1]);
    dmma884(ci0, ci1, pr1, b.i[1]);
    dmma884(cr0, cr1, mi0, b.i[0]);
    dmma884(ci0, ci1, pi0, b.r[0]);
    dmma884(cr0, cr1, mi1, b.i[1]);
    dmma884(ci0, ci1, pi1, b.r[1]);
}

// Two INDEPENDENT block products issued DMMA by DMMA: four accumulation chains in flight instead of two.  A dependent DMMA may issue
// 32.6 cycles after its predecessor (2 issue slots of 16.3 cycles, profiles/r01_microbench_fp64.log), so the two chains of one product
// leave no slack at all - ptxas pads them with NOPs (12 % of the stall samples of the kernel, profiles/r02_ncu_k3fast_summary.txt) -
// while four chains put every dependent DMMA four slots behind its predecessor.
template <bool NEG1, bool NEG2>
__device__ __forceinline__ void bmm2(double& c1r0, double& c1r1, double& c1i0, double& c1i1, double a1r0, double a1r1, double a1i0, double a1i1,
                                     const BFrag& b1, double& c2r0, double& c2r1, double& c2i0, double& c2i1, double a2r0, double a2r1,
                                     double a2i0, double a2i1, const BFrag& b2) {
    const double p1r0 = NEG1 ? -a1r0 : a1r0, p1r1 = NEG1 ? -a1r1 : a1r1, p1i0 = NEG1 ? -a1i0 : a1i0, p1i1 = NEG1 ? -a1i1 : a1i1;
    const double m1i0 = NEG1 ? a1i0 : -a1i0, m1i1 = NEG1 ? a1i1 : -a1i1;
    const double p2r0 = NEG2 ? -a2r0 : a2r0, p2r1 = NEG2 ? -a2r1 : a2r1, p2i0 = NEG2 ? -a2i0 : a2i0, p2i1 = NEG2 ? -a2i1 : a2i1;
    const double m2i0 = NEG2 ? a2i0 : -a2i0, m2i1 = NEG2 ? a2i1 : -a2i1;
    dmma884(c1r0, c1r1, p1r0, b1.r[0]); dmma884(c1i0, c1i1, p1r0, b1.i[0]);
    dmma884(c2r0, c2r1, p2r0, b2.r[0]); dmma884(c2i0, c2i1, p2r0, b2.i[0]);
    dmma884(c1r0, c1r1, p1r1, b1.r[1]); dmma884(c1i0, c1i1, p1r1, b1.i[1]);
    dmma884(c2r0, c2r1, p2r1, b2.r[1]); dmma884(c2i0, c2i1, p2r1, b2.i[1]);
    dmma884(c1r0, c1r1, m1i0, b1.i[0]); dmma884(c1i0, c1i1, p1i0, b1.r[0]);
    dmma884(c2r0, c2r1, m2i0, b2.i[0]); dmma884(c2i0, c2i1, p2i0, b2.r[0]);
    dmma884(c1r0, c1r1, m1i1, b1.i[1]); dmma884(c1i0, c1i1, p1i1, b1.r[1]);
    dmma884(c2r0, c2r1, m2i1, b2.i[1]); dmma884(c2i0, c2i1, p2i1, b2.r[1]);
}
// one block product on four chains: the two k-slabs accumulate separately and are added at the end (4 DADD)
template <bool NEG>
__device__ __forceinline__ void bmm_split(double& cr0, double& cr1, double& ci0, double& ci1, double ar0, double ar1, double ai0, double ai1,
                                          const BFrag& b) {
    const double pr0 = NEG ? -ar0 : ar0, pr1 = NEG ? -ar1 : ar1, pi0 = NEG ? -ai0 : ai0, pi1 = NEG ? -ai1 : ai1;
    const double mi0 = NEG ? ai0 : -ai0, mi1 = NEG ? ai1 : -ai1;
    double tr0 = 0, tr1 = 0, ti0 = 0, ti1 = 0;
    dmma884(cr0, cr1, pr0, b.r[0]); dmma884(ci0, ci1, pr0, b.i[0]);
    dmma884(tr0, tr1, pr1, b.r[1]); dmma884(ti0, ti1, pr1, b.i[1]);
    dmma884(cr0, cr1, mi0, b.i[0]); dmma884(ci0, ci1, pi0, b.r[0]);
    dmma884(tr0, tr1, mi1, b.i[1]); dmma884(ti0, ti1, pi1, b.r[1]);
    cr0 += tr0; cr1 += tr1; ci0 += ti0; ci1 += ti1;
}

// sign flip on the integer pipe (keeps the FP64 pipe for FMA/DMMA work)
__device__ __forceinline__ double dneg(double x) { return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x)); }

// in-place inverse of an 8x8 complex block in C layout by Gauss-Jordan without row exchanges.
// Per pivot p: m = a_gp / a_pp for every row g (row p itself uses m = 1 - 1/a_pp so that the same update
// x -= m * pivot_row scales it), then the pivot column is overwritten (row p: 1/a_pp, others: -m).
// f * conj(pivot) is formed while the reciprocal of |pivot|^2 is in flight: dependent FP64 depth 8 per step.
// minhi tracks min_p |pivot_p|^2 through the high word (integer pipe).
__device__ __forceinline__ void inv8(double& xr0, double& xr1, double& xi0, double& xi1, int lane, int& minhi) {
    const int g = lane >> 2, q = lane & 3, quad = lane & ~3;
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const int sg = p & 1, qp = p >> 1;
        // pivot row restricted to my two columns
        const double pr0 = __shfl_sync(0xffffffffu, xr0, 4 * p + q), pr1 = __shfl_sync(0xffffffffu, xr1, 4 * p + q);
        const double pi0 = __shfl_sync(0xffffffffu, xi0, 4 * p + q), pi1 = __shfl_sync(0xffffffffu, xi1, 4 * p + q);
        // pivot a_pp and my row's multiplier a_gp
        // straight from the lane that owns a_pp, not through the shuffled pivot row: one shuffle less on the dependent
        // chain of every pivot step (768 -> 782 k k-points/s on the C4 workload)
        const double ppr = __shfl_sync(0xffffffffu, sg ? xr1 : xr0, 4 * p + qp);
        const double ppi = __shfl_sync(0xffffffffu, sg ? xi1 : xi0, 4 * p + qp);
        const double fr = __shfl_sync(0xffffffffu, sg ? xr1 : xr0, quad | qp);
        const double fi = __shfl_sync(0xffffffffu, sg ? xi1 : xi0, quad | qp);
        const double d = fma(ppr, ppr, ppi * ppi);
        minhi = min(minhi, __double2hiint(d));
        const double tr_ = fma(fr, ppr, fi * ppi), ti_ = fma(-fr, ppi, fi * ppr);     // f * conj(pivot)
        const double dinv = fast_rcp(d);
        double mr = tr_ * dinv, mi = ti_ * dinv;                                       // a_gp / a_pp
        const double rr = ppr * dinv, nri = ppi * dinv;                                 // 1 / a_pp = rr - i nri
        const bool prow = (g == p);
        if (prow) { mr = 1.0 - rr; mi = nri; }
        xr0 = fma(-mr, pr0, xr0); xr0 = fma(mi, pi0, xr0);
        xi0 = fma(-mr, pi0, xi0); xi0 = fma(-mi, pr0, xi0);
        xr1 = fma(-mr, pr1, xr1); xr1 = fma(mi, pi1, xr1);
        xi1 = fma(-mr, pi1, xi1); xi1 = fma(-mi, pr1, xi1);
        if (q == qp) {
            const double cr = prow ? rr : dneg(mr), ci = dneg(prow ? nri : mi);
            if (sg) { xr1 = cr; xi1 = ci; } else { xr0 = cr; xi0 = ci; }
        }
    }
}

// Division-free form of inv8 (VAR 4).  The dependent chain of a pivot step of inv8 is  exchange -> |pivot|^2 -> reciprocal (MUFU seed + two
// Newton steps) -> multiplier -> two FMAs, about nine visits of the FP64 pipe that DMMA shares; here a step is the cross-multiplied row
// operation  row_g <- a_pp row_g - a_gp row_p  (g != p): exchange -> four FMAs, no reciprocal at all.  Every row carries its own scale:
// z_g is that scale c_g until the row's own pivot step and the diagonal entry s_g = pivot_g c_g afterwards (both are multiplied by a_pp in every
// other step), the freed column p receives the new column of R with R_pp = c_p, and ONE reciprocal per row at the end gives
// inv = diag(1/s) R.  Row scales square from step to step (c ~ pivot^(2^p - 1)), so the rows are rescaled by an exact power of two
// after four steps (|entries| stay below ~ max|a|^30: overflow needs max|a| > 1e10 and ends as NaN -> flag -> pivoted rerun).  Same FP64
// instruction count as inv8 (8 x 20 + 22 against 8 x 23), less than half its dependent depth.  The pivot monitor works on exponents:
// log2|pivot_p| ~ mag(a_pp) - mag(c_p) with mag = high word of max(|re|, |im|), kept per lane (row) and min-reduced by the caller.
__device__ __forceinline__ int dmag(double re, double im) { return max(__double2hiint(re) & 0x7fffffff, __double2hiint(im) & 0x7fffffff); }
__device__ __forceinline__ void inv8_ff(double& xr0, double& xr1, double& xi0, double& xi1, int lane, int& minhi) {
    const int g = lane >> 2, q = lane & 3, quad = lane & ~3;
    double zr = 1.0, zi = 0.0;
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const int sg = p & 1, qp = p >> 1;
        const bool prow = (g == p), pcol = (q == qp);
        const double ppr = __shfl_sync(0xffffffffu, sg ? xr1 : xr0, 4 * p + qp);       // a_pp
        const double ppi = __shfl_sync(0xffffffffu, sg ? xi1 : xi0, 4 * p + qp);
        const double gfr = __shfl_sync(0xffffffffu, sg ? xr1 : xr0, quad | qp);        // a_gp of my row
        const double gfi = __shfl_sync(0xffffffffu, sg ? xi1 : xi0, quad | qp);
        // column p turns into the new column of R: c_p in row p, zero elsewhere (then updated like every other column)
        const double cr = prow ? zr : 0.0, ci = prow ? zi : 0.0;
        if (sg) { xr1 = pcol ? cr : xr1; xi1 = pcol ? ci : xi1; } else { xr0 = pcol ? cr : xr0; xi0 = pcol ? ci : xi0; }
        const double pr0 = __shfl_sync(0xffffffffu, xr0, 4 * p + q), pi0 = __shfl_sync(0xffffffffu, xi0, 4 * p + q);
        const double pr1 = __shfl_sync(0xffffffffu, xr1, 4 * p + q), pi1 = __shfl_sync(0xffffffffu, xi1, 4 * p + q);
        const int dl = min(max(dmag(ppr, ppi) - dmag(zr, zi), -0x1ff00000), 0x1ff00000);
        minhi = min(minhi, prow ? 0x3ff00000 + 2 * dl : 0x7ff00000);                    // ~ high word of |pivot_p|^2
        // the pivot row itself is only scaled by a_pp (f = 0); its z becomes the diagonal entry
        const double fr = prow ? 0.0 : gfr, fi = prow ? 0.0 : gfi;
        const double wr = prow ? ppr : zr, wi = prow ? ppi : zi;
        double t;
        t = ppr * xr0; t = fma(-ppi, xi0, t); t = fma(-fr, pr0, t); const double nr0 = fma(fi, pi0, t);
        t = ppr * xi0; t = fma(ppi, xr0, t); t = fma(-fr, pi0, t); const double ni0 = fma(-fi, pr0, t);
        t = ppr * xr1; t = fma(-ppi, xi1, t); t = fma(-fr, pr1, t); const double nr1 = fma(fi, pi1, t);
        t = ppr * xi1; t = fma(ppi, xr1, t); t = fma(-fr, pi1, t); const double ni1 = fma(-fi, pr1, t);
        xr0 = nr0; xi0 = ni0; xr1 = nr1; xi1 = ni1;
        zr = fma(-ppi, wi, ppr * wr); zi = fma(ppi, wr, ppr * wi);
        if (p == 3) {                        // exact rescaling of my row by 2^-exponent(z)
            const double r = __hiloint2double(0x7fe00000 - (dmag(zr, zi) & 0x7ff00000), 0);
            xr0 *= r; xi0 *= r; xr1 *= r; xi1 *= r; zr *= r; zi *= r;
        }
    }
    const double dinv = fast_rcp(fma(zr, zr, zi * zi));
    const double ir = zr * dinv, ii = -zi * dinv;                                      // 1 / s_g
    double t;
    t = ir * xr0; const double nr0 = fma(-ii, xi0, t); t = ir * xi0; const double ni0 = fma(ii, xr0, t);
    t = ir * xr1; const double nr1 = fma(-ii, xi1, t); t = ir * xi1; const double ni1 = fma(ii, xr1, t);
    xr0 = nr0; xi0 = ni0; xr1 = nr1; xi1 = ni1;
}

// One pivot step of inv8 in two halves, so that independent DMMA work can be placed BETWEEN them in program order: the exchange
// (eight shuffles: pivot row, pivot, this row's multiplier) and the arithmetic.  The compiler keeps shuffles and mma.sync in source
// order and only floats the scalar FP64 chain, so the block product written between the halves is what fills the chain's latency.
struct PivIn { double pr0, pr1, pi0, pi1, ppr, ppi, fr, fi; };
template <int P>
__device__ __forceinline__ PivIn inv8_exchange(double xr0, double xr1, double xi0, double xi1, int lane) {
    const int q = lane & 3, quad = lane & ~3;
    constexpr int sg = P & 1, qp = P >> 1;
    PivIn v;
    v.pr0 = __shfl_sync(0xffffffffu, xr0, 4 * P + q); v.pr1 = __shfl_sync(0xffffffffu, xr1, 4 * P + q);
    v.pi0 = __shfl_sync(0xffffffffu, xi0, 4 * P + q); v.pi1 = __shfl_sync(0xffffffffu, xi1, 4 * P + q);
    v.ppr = __shfl_sync(0xffffffffu, sg ? xr1 : xr0, 4 * P + qp);
    v.ppi = __shfl_sync(0xffffffffu, sg ? xi1 : xi0, 4 * P + qp);
    v.fr = __shfl_sync(0xffffffffu, sg ? xr1 : xr0, quad | qp);
    v.fi = __shfl_sync(0xffffffffu, sg ? xi1 : xi0, quad | qp);
    return v;
}
template <int P>
__device__ __forceinline__ void inv8_eliminate(double& xr0, double& xr1, double& xi0, double& xi1, const PivIn& v, int lane, int& minhi) {
    const int g = lane >> 2, q = lane & 3;
    constexpr int sg = P & 1, qp = P >> 1;
    const double d = fma(v.ppr, v.ppr, v.ppi * v.ppi);
    minhi = min(minhi, __double2hiint(d));
    const double tr_ = fma(v.fr, v.ppr, v.fi * v.ppi), ti_ = fma(-v.fr, v.ppi, v.fi * v.ppr);
    const double dinv = fast_rcp(d);
    double mr = tr_ * dinv, mi = ti_ * dinv;
    const double rr = v.ppr * dinv, nri = v.ppi * dinv;
    const bool prow = (g == P);
    if (prow) { mr = 1.0 - rr; mi = nri; }
    xr0 = fma(-mr, v.pr0, xr0); xr0 = fma(mi, v.pi0, xr0);
    xi0 = fma(-mr, v.pi0, xi0); xi0 = fma(-mi, v.pr0, xi0);
    xr1 = fma(-mr, v.pr1, xr1); xr1 = fma(mi, v.pi1, xr1);
    xi1 = fma(-mr, v.pi1, xi1); xi1 = fma(-mi, v.pr1, xi1);
    if (q == qp) {
        const double cr = prow ? rr : dneg(mr), ci = dneg(prow ? nri : mi);
        if (sg) { xr1 = cr; xi1 = ci; } else { xr0 = cr; xi0 = ci; }
    }
}

// VAR 2 (norb 25..32, NB = 4): the same block LU + trace from the factors as a STATIC SOFTWARE PIPELINE.  Each of the 8 pivot steps
// of a diagonal-block inversion is issued around one independent block product, and the substitution phase is reformulated so that
// enough independent products exist for every inversion but the first:
//   W = -(I + X~)^-1 + I (strict upper, in place over X, rows top down): W_ij = X_ij - sum_{i<t<j} W_it X_tj      - needs no D at all
//   N = -L~^-1 + I      (strict lower, in place over L, columns left to right): N_ij = L_ij - sum_{j<t<i} L_it N_tj
//   tr A^-1 = sum_s tr D_s + sum_{i<j} tr(W_ij (D_j N_ji))      (U~^-1 = (I + X~)^-1 diag(D); the two sign flips cancel)
// schedule:  D0 | L_i0, column 1 | D1 with {X_01, columns 2, 3 of step 0} | L_i1, column 2 | D2 with {X_03, X_12, column 3 of step 1,
// W_02, N_20, N_30, D1 N_10} | L_32, A_33 | D3 with {X_23, W_03, W_13, N_30, N_31, D2 N_20, D2 N_21} | D3 N_3i and the trace.
__device__ __forceinline__ double2 warp_trace_inverse_pipelined(double (&R0)[4][4], double (&R1)[4][4], double (&I0)[4][4],
                                                               double (&I1)[4][4], int lane, int& minpiv) {
    const int g = lane >> 2, q = lane & 3;
    const bool par = g & 1;
    const int src0 = 4 * (2 * q + (par ? 1 : 0)) + (g >> 1);
    const int src1 = 4 * (2 * q + (par ? 0 : 1)) + (g >> 1);
    double tr = 0.0, ti = 0.0;
#define ABZ_FRAG(i, j) to_bfrag(R0[i][j], R1[i][j], I0[i][j], I1[i][j], src0, src1, par)
    // C_ij -= A_it * b      (trailing update, W and N substitutions)
#define ABZ_UPD(i, j, t, b) bmm<true>(R0[i][j], R1[i][j], I0[i][j], I1[i][j], R0[i][t], R1[i][t], I0[i][t], I1[i][t], b)
    // C_ij = A_it * b       (L_is = A_is D_s in place with t = j; X_sj = D_s U_sj in place with i = t)
#define ABZ_SET(i, j, it, tt, b)                                                                                       \
    {                                                                                                                  \
        double c0_ = 0, c1_ = 0, c2_ = 0, c3_ = 0;                                                                     \
        bmm<false>(c0_, c1_, c2_, c3_, R0[it][tt], R1[it][tt], I0[it][tt], I1[it][tt], b);                             \
        R0[i][j] = c0_; R1[i][j] = c1_; I0[i][j] = c2_; I1[i][j] = c3_;                                                \
    }
    // tr += tr(W_ij * (D_j N_ji)): product in C layout, its fragment against W_ij
#define ABZ_TRACE(ia, ja)                                                                                              \
    {                                                                                                                  \
        const BFrag bn_ = ABZ_FRAG(ja, ia);                                                                            \
        double c0_ = 0, c1_ = 0, c2_ = 0, c3_ = 0;                                                                     \
        bmm<false>(c0_, c1_, c2_, c3_, R0[ja][ja], R1[ja][ja], I0[ja][ja], I1[ja][ja], bn_);                           \
        const BFrag bt_ = to_bfrag(c0_, c1_, c2_, c3_, src0, src1, par);                                               \
        tr += R0[ia][ja] * bt_.r[0] - I0[ia][ja] * bt_.i[0] + R1[ia][ja] * bt_.r[1] - I1[ia][ja] * bt_.i[1];           \
        ti += R0[ia][ja] * bt_.i[0] + I0[ia][ja] * bt_.r[0] + R1[ia][ja] * bt_.i[1] + I1[ia][ja] * bt_.r[1];           \
    }
#define ABZ_PIV(K, P, WORK)                                                                                            \
    {                                                                                                                  \
        const PivIn pv_ = inv8_exchange<P>(R0[K][K], R1[K][K], I0[K][K], I1[K][K], lane);                              \
        WORK;                                                                                                          \
        inv8_eliminate<P>(R0[K][K], R1[K][K], I0[K][K], I1[K][K], pv_, lane, minpiv);                                  \
    }
#define ABZ_DIAG_TRACE(K)                                                                                              \
    {                                                                                                                  \
        if (2 * q == g) { tr += R0[K][K]; ti += I0[K][K]; }                                                            \
        if (2 * q + 1 == g) { tr += R1[K][K]; ti += I1[K][K]; }                                                        \
    }
    // ---- D0 (nothing of this matrix to overlap with)
    inv8(R0[0][0], R1[0][0], I0[0][0], I1[0][0], lane, minpiv);
    ABZ_DIAG_TRACE(0)
    {
        const BFrag bD0 = ABZ_FRAG(0, 0);
        ABZ_SET(1, 0, 1, 0, bD0) ABZ_SET(2, 0, 2, 0, bD0) ABZ_SET(3, 0, 3, 0, bD0)            // L_i0 = A_i0 D_0
    }
    const BFrag bU01 = ABZ_FRAG(0, 1);
    ABZ_UPD(1, 1, 0, bU01); ABZ_UPD(2, 1, 0, bU01); ABZ_UPD(3, 1, 0, bU01);                   // column 1 of step 0
    {   // ---- D1 around X_01 and columns 2, 3 of step 0
        const BFrag bU02 = ABZ_FRAG(0, 2);
        ABZ_PIV(1, 0, ABZ_SET(0, 1, 0, 0, bU01))
        ABZ_PIV(1, 1, ABZ_UPD(1, 2, 0, bU02))
        ABZ_PIV(1, 2, ABZ_UPD(2, 2, 0, bU02))
        ABZ_PIV(1, 3, ABZ_UPD(3, 2, 0, bU02))
        const BFrag bU03 = ABZ_FRAG(0, 3);
        ABZ_PIV(1, 4, ABZ_SET(0, 2, 0, 0, bU02))
        ABZ_PIV(1, 5, ABZ_UPD(1, 3, 0, bU03))
        ABZ_PIV(1, 6, ABZ_UPD(2, 3, 0, bU03))
        ABZ_PIV(1, 7, ABZ_UPD(3, 3, 0, bU03))
        ABZ_SET(0, 3, 0, 0, bU03)                                                               // X_03
    }
    ABZ_DIAG_TRACE(1)
    {
        const BFrag bD1 = ABZ_FRAG(1, 1);
        ABZ_SET(2, 1, 2, 1, bD1) ABZ_SET(3, 1, 3, 1, bD1)                                       // L_i1
    }
    const BFrag bU12 = ABZ_FRAG(1, 2);
    ABZ_UPD(2, 2, 1, bU12); ABZ_UPD(3, 2, 1, bU12);                                            // column 2 of step 1
    {   // ---- D2 around X_12, column 3 of step 1, W_02, N_20, N_30 (first term), D1 N_10
        const BFrag bU13 = ABZ_FRAG(1, 3);
        ABZ_PIV(2, 0, ABZ_SET(1, 2, 1, 1, bU12))                                                // X_12
        ABZ_PIV(2, 1, ABZ_UPD(2, 3, 1, bU13))
        ABZ_PIV(2, 2, ABZ_UPD(3, 3, 1, bU13))
        ABZ_PIV(2, 3, ABZ_SET(1, 3, 1, 1, bU13))                                                // X_13
        const BFrag bN10 = ABZ_FRAG(1, 0);                                                      // N_10 = L_10
        ABZ_PIV(2, 4, ABZ_UPD(2, 0, 1, bN10))                                                   // N_20 = L_20 - L_21 N_10
        ABZ_PIV(2, 5, ABZ_UPD(3, 0, 1, bN10))                                                   // N_30 = L_30 - L_31 N_10 (- L_32 N_20 later)
        const BFrag bX12 = ABZ_FRAG(1, 2);
        ABZ_PIV(2, 6, ABZ_UPD(0, 2, 1, bX12))                                                   // W_02 = X_02 - W_01 X_12
        ABZ_PIV(2, 7, ABZ_TRACE(0, 1))                                                          // tr(W_01 D_1 N_10)
    }
    ABZ_DIAG_TRACE(2)
    {
        const BFrag bD2 = ABZ_FRAG(2, 2);
        ABZ_SET(3, 2, 3, 2, bD2)                                                                // L_32
    }
    const BFrag bU23 = ABZ_FRAG(2, 3);
    ABZ_UPD(3, 3, 2, bU23);                                                                    // A_33 of step 2
    {   // ---- D3 around X_23, W_03, W_13, N_30 (second term), N_31, D2 N_20, D2 N_21
        const BFrag bX13 = ABZ_FRAG(1, 3);                                                      // X_13, still unmodified
        ABZ_PIV(3, 0, ABZ_SET(2, 3, 2, 2, bU23))                                                // X_23
        ABZ_PIV(3, 1, ABZ_UPD(0, 3, 1, bX13))                                                   // W_03 = X_03 - W_01 X_13 ...
        const BFrag bX23 = ABZ_FRAG(2, 3);
        ABZ_PIV(3, 2, ABZ_UPD(0, 3, 2, bX23))                                                   //        ... - W_02 X_23
        ABZ_PIV(3, 3, ABZ_UPD(1, 3, 2, bX23))                                                   // W_13 = X_13 - W_12 X_23
        const BFrag bN20 = ABZ_FRAG(2, 0);
        ABZ_PIV(3, 4, ABZ_UPD(3, 0, 2, bN20))                                                   // N_30 -= L_32 N_20
        const BFrag bN21 = ABZ_FRAG(2, 1);                                                      // N_21 = L_21
        ABZ_PIV(3, 5, ABZ_UPD(3, 1, 2, bN21))                                                   // N_31 = L_31 - L_32 N_21
        ABZ_PIV(3, 6, ABZ_TRACE(0, 2))                                                          // tr(W_02 D_2 N_20)
        ABZ_PIV(3, 7, ABZ_TRACE(1, 2))                                                          // tr(W_12 D_2 N_21)
    }
    ABZ_DIAG_TRACE(3)
    ABZ_TRACE(0, 3) ABZ_TRACE(1, 3) ABZ_TRACE(2, 3)                                             // tr(W_i3 D_3 N_3i)
#undef ABZ_FRAG
#undef ABZ_UPD
#undef ABZ_SET
#undef ABZ_TRACE
#undef ABZ_PIV
#undef ABZ_DIAG_TRACE
    return make_double2(warp_sum(tr), warp_sum(ti));
}

// VAR 3: VAR 1 (right-looking substitutions) with the independent block products of every step issued in PAIRS (bmm2: four DMMA
// accumulation chains in flight), single left-overs on four chains too (bmm_split).  Same block algebra as VAR 0 / 1.
template <int NB>
__device__ __forceinline__ double2 warp_trace_inverse_paired(double (&R0)[NB][NB], double (&R1)[NB][NB], double (&I0)[NB][NB],
                                                            double (&I1)[NB][NB], int lane, int& minpiv) {
    const int g = lane >> 2, q = lane & 3;
    const bool par = g & 1;
    const int src0 = 4 * (2 * q + (par ? 1 : 0)) + (g >> 1);
    const int src1 = 4 * (2 * q + (par ? 0 : 1)) + (g >> 1);
#define ABZ_B(i, j) R0[i][j], R1[i][j], I0[i][j], I1[i][j]
    inv8(ABZ_B(0, 0), lane, minpiv);
#pragma unroll
    for (int s = 0; s < NB - 1; s++) {
        const BFrag bD = to_bfrag(ABZ_B(s, s), src0, src1, par);
        {   // L_is = A_is D_s, rows in pairs
#pragma unroll
            for (int pp = 0; pp < (NB - 1 - s) / 2; pp++) {
                const int i = s + 1 + 2 * pp;
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
                bmm2<false, false>(a0, a1, a2, a3, ABZ_B(i, s), bD, b0, b1, b2, b3, ABZ_B(i + 1, s), bD);
                R0[i][s] = a0; R1[i][s] = a1; I0[i][s] = a2; I1[i][s] = a3;
                R0[i + 1][s] = b0; R1[i + 1][s] = b1; I0[i + 1][s] = b2; I1[i + 1][s] = b3;
            }
            if ((NB - 1 - s) & 1) {
                const int i = NB - 1;
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                bmm_split<false>(a0, a1, a2, a3, ABZ_B(i, s), bD);
                R0[i][s] = a0; R1[i][s] = a1; I0[i][s] = a2; I1[i][s] = a3;
            }
        }
#pragma unroll
        for (int j = s + 1; j < NB; j++) {
            const BFrag bU = to_bfrag(ABZ_B(s, j), src0, src1, par);
            // A_ij -= L_is U_sj (i > s) and X_sj = D_s U_sj: NB - s independent products sharing bU, in pairs
#pragma unroll
            for (int pp = 0; pp < (NB - 1 - s) / 2; pp++) {
                const int i = s + 1 + 2 * pp;
                bmm2<true, true>(ABZ_B(i, j), ABZ_B(i, s), bU, ABZ_B(i + 1, j), ABZ_B(i + 1, s), bU);
            }
            double x0 = 0, x1 = 0, x2 = 0, x3 = 0;
            if ((NB - 1 - s) & 1) bmm2<true, false>(ABZ_B(NB - 1, j), ABZ_B(NB - 1, s), bU, x0, x1, x2, x3, ABZ_B(s, s), bU);
            else bmm_split<false>(x0, x1, x2, x3, ABZ_B(s, s), bU);
            R0[s][j] = x0; R1[s][j] = x1; I0[s][j] = x2; I1[s][j] = x3;
            if (j == s + 1) inv8(ABZ_B(j, j), lane, minpiv);                 // D_{s+1} (look-ahead)
        }
    }
    double tr = 0.0, ti = 0.0;
#pragma unroll
    for (int s = 0; s < NB; s++) {
        if (2 * q == g) { tr += R0[s][s]; ti += I0[s][s]; }
        if (2 * q + 1 == g) { tr += R1[s][s]; ti += I1[s][s]; }
    }
    // V = U~^-1, columns right to left
#pragma unroll
    for (int jv = NB - 1; jv >= 1; jv--) {
        const BFrag bD = to_bfrag(ABZ_B(jv, jv), src0, src1, par);
        {
#pragma unroll
            for (int pp = 0; pp < jv / 2; pp++) {                            // V'_{iv,jv} = -X_{iv,jv} D_jv
                const int iv = 2 * pp;
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
                bmm2<true, true>(a0, a1, a2, a3, ABZ_B(iv, jv), bD, b0, b1, b2, b3, ABZ_B(iv + 1, jv), bD);
                R0[iv][jv] = a0; R1[iv][jv] = a1; I0[iv][jv] = a2; I1[iv][jv] = a3;
                R0[iv + 1][jv] = b0; R1[iv + 1][jv] = b1; I0[iv + 1][jv] = b2; I1[iv + 1][jv] = b3;
            }
            if (jv & 1) {
                const int iv = jv - 1;
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                bmm_split<true>(a0, a1, a2, a3, ABZ_B(iv, jv), bD);
                R0[iv][jv] = a0; R1[iv][jv] = a1; I0[iv][jv] = a2; I1[iv][jv] = a3;
            }
        }
#pragma unroll
        for (int t = jv - 1; t >= 1; t--) {
            const BFrag bV = to_bfrag(ABZ_B(t, jv), src0, src1, par);
#pragma unroll
            for (int pp = 0; pp < t / 2; pp++) bmm2<true, true>(ABZ_B(2 * pp, jv), ABZ_B(2 * pp, t), bV, ABZ_B(2 * pp + 1, jv), ABZ_B(2 * pp + 1, t), bV);
            if (t & 1) bmm_split<true>(ABZ_B(t - 1, jv), ABZ_B(t - 1, t), bV);
        }
    }
    // M = L~^-1, columns left to right; tr += tr(V_{jm,t} M_{t,jm}) as soon as M_{t,jm} is final
#pragma unroll
    for (int jm = 0; jm < NB - 1; jm++) {
#pragma unroll
        for (int im = jm + 1; im < NB; im++) {
            R0[im][jm] = dneg(R0[im][jm]); R1[im][jm] = dneg(R1[im][jm]); I0[im][jm] = dneg(I0[im][jm]); I1[im][jm] = dneg(I1[im][jm]);
        }
#pragma unroll
        for (int t = jm + 1; t < NB; t++) {
            const BFrag b = to_bfrag(ABZ_B(t, jm), src0, src1, par);
            tr += R0[jm][t] * b.r[0] - I0[jm][t] * b.i[0] + R1[jm][t] * b.r[1] - I1[jm][t] * b.i[1];
            ti += R0[jm][t] * b.i[0] + I0[jm][t] * b.r[0] + R1[jm][t] * b.i[1] + I1[jm][t] * b.r[1];
#pragma unroll
            for (int pp = 0; pp < (NB - 1 - t) / 2; pp++) {
                const int im = t + 1 + 2 * pp;
                bmm2<true, true>(ABZ_B(im, jm), ABZ_B(im, t), b, ABZ_B(im + 1, jm), ABZ_B(im + 1, t), b);
            }
            if ((NB - 1 - t) & 1) bmm_split<true>(ABZ_B(NB - 1, jm), ABZ_B(NB - 1, t), b);
        }
    }
#undef ABZ_B
    return make_double2(warp_sum(tr), warp_sum(ti));
}

// VAR 0: left-looking substitutions (V and M interleaved, one accumulation chain per block);
// VAR 1: right-looking substitutions - a finished block of V (M) is turned into its fragment once and immediately applied to all
//        rows above (below) it, so the block products of one step are independent (more DMMA chains in flight, one live fragment
//        instead of the bV[] / bM[] arrays), and the fragments of M feed the trace as they are made.
// PF = 1 / 2: the B fragments of the LU phase are made one product group AHEAD of their use (ptxas keeps shfl.sync and mma.sync in source
// order, so a fragment made right before its products exposes the shuffle + select latency to the warp every time): U_{s,j+1} is
// exchanged before the products of column j (PF >= 1), D_{s+1} right after its inversion, before the rest of step s (PF = 2).
template <int NB, int VAR, bool FF = false, int PF = 0>
__device__ __forceinline__ double2 warp_trace_inverse(double (&R0)[NB][NB], double (&R1)[NB][NB], double (&I0)[NB][NB],
                                                      double (&I1)[NB][NB], int lane, int& minpiv) {
    const int g = lane >> 2, q = lane & 3;
    const bool par = g & 1;
    const int src0 = 4 * (2 * q + (par ? 1 : 0)) + (g >> 1);
    const int src1 = 4 * (2 * q + (par ? 0 : 1)) + (g >> 1);
    // ---- block LU with look-ahead: S_{s+1} is inverted as soon as its update is complete, so that the
    //      latency-bound pivot steps overlap the remaining (independent) trailing-update DMMAs
    if constexpr (FF) inv8_ff(R0[0][0], R1[0][0], I0[0][0], I1[0][0], lane, minpiv);
    else inv8(R0[0][0], R1[0][0], I0[0][0], I1[0][0], lane, minpiv);          // D_0
    BFrag bDn;
    if constexpr (PF >= 2) bDn = to_bfrag(R0[0][0], R1[0][0], I0[0][0], I1[0][0], src0, src1, par);
#pragma unroll
    for (int s = 0; s < NB - 1; s++) {
        const BFrag bD = PF >= 2 ? bDn : to_bfrag(R0[s][s], R1[s][s], I0[s][s], I1[s][s], src0, src1, par);
        BFrag bUn;
        if constexpr (PF >= 1) bUn = to_bfrag(R0[s][s + 1], R1[s][s + 1], I0[s][s + 1], I1[s][s + 1], src0, src1, par);
#pragma unroll
        for (int i = s + 1; i < NB; i++) {                               // L_is = A_is D_s
            double cr0 = 0, cr1 = 0, ci0 = 0, ci1 = 0;
            bmm<false>(cr0, cr1, ci0, ci1, R0[i][s], R1[i][s], I0[i][s], I1[i][s], bD);
            R0[i][s] = cr0; R1[i][s] = cr1; I0[i][s] = ci0; I1[i][s] = ci1;
        }
#pragma unroll
        for (int j = s + 1; j < NB; j++) {
            const BFrag bU = PF >= 1 ? bUn : to_bfrag(R0[s][j], R1[s][j], I0[s][j], I1[s][j], src0, src1, par);
            if constexpr (PF >= 1) {
                if (j + 1 < NB) bUn = to_bfrag(R0[s][j + 1], R1[s][j + 1], I0[s][j + 1], I1[s][j + 1], src0, src1, par);
            }
#pragma unroll
            for (int i = s + 1; i < NB; i++)                             // A_ij -= L_is U_sj
                bmm<true>(R0[i][j], R1[i][j], I0[i][j], I1[i][j], R0[i][s], R1[i][s], I0[i][s], I1[i][s], bU);
            if (j == s + 1) {                                            // D_{s+1} (look-ahead)
                if constexpr (FF) inv8_ff(R0[j][j], R1[j][j], I0[j][j], I1[j][j], lane, minpiv);
                else inv8(R0[j][j], R1[j][j], I0[j][j], I1[j][j], lane, minpiv);
                if constexpr (PF >= 2) bDn = to_bfrag(R0[j][j], R1[j][j], I0[j][j], I1[j][j], src0, src1, par);
            }
            {                                                            // X_sj = D_s U_sj (replaces U_sj)
                double cr0 = 0, cr1 = 0, ci0 = 0, ci1 = 0;
                bmm<false>(cr0, cr1, ci0, ci1, R0[s][s], R1[s][s], I0[s][s], I1[s][s], bU);
                R0[s][j] = cr0; R1[s][j] = cr1; I0[s][j] = ci0; I1[s][j] = ci1;
            }
        }
    }
    // ---- trace of the diagonal blocks D_s
    double tr = 0.0, ti = 0.0;
#pragma unroll
    for (int s = 0; s < NB; s++) {
        if (2 * q == g) { tr += R0[s][s]; ti += I0[s][s]; }
        if (2 * q + 1 == g) { tr += R1[s][s]; ti += I1[s][s]; }
    }
    if constexpr (VAR == 1 && PF >= 3) {
        // The right-looking substitutions of VAR 1 with every fragment made AHEAD of its products: the block that becomes final first
        // (row t-1 of a V column, row t+1 of an M column) is updated first and exchanged while the remaining products of the step issue.
        // Every block still receives its terms in the same order: bit-identical to VAR 1.
#define ABZ_B(i, j) R0[i][j], R1[i][j], I0[i][j], I1[i][j]
#pragma unroll
        for (int jv = NB - 1; jv >= 1; jv--) {
            const BFrag bD = to_bfrag(ABZ_B(jv, jv), src0, src1, par);
            BFrag bV;
            {   // V'_{iv,jv} = -X_{iv,jv} D_jv, row jv-1 first (final at once: no t between jv-1 and jv)
                double cr0 = 0, cr1 = 0, ci0 = 0, ci1 = 0;
                bmm<true>(cr0, cr1, ci0, ci1, ABZ_B(jv - 1, jv), bD);
                R0[jv - 1][jv] = cr0; R1[jv - 1][jv] = cr1; I0[jv - 1][jv] = ci0; I1[jv - 1][jv] = ci1;
                if (jv - 1 >= 1) bV = to_bfrag(ABZ_B(jv - 1, jv), src0, src1, par);
            }
#pragma unroll
            for (int iv = jv - 2; iv >= 0; iv--) {
                double cr0 = 0, cr1 = 0, ci0 = 0, ci1 = 0;
                bmm<true>(cr0, cr1, ci0, ci1, ABZ_B(iv, jv), bD);
                R0[iv][jv] = cr0; R1[iv][jv] = cr1; I0[iv][jv] = ci0; I1[iv][jv] = ci1;
            }
#pragma unroll
            for (int t = jv - 1; t >= 1; t--) {                          // bV = fragment of the final V_{t,jv}
                bmm<true>(ABZ_B(t - 1, jv), ABZ_B(t - 1, t), bV);        // V_{t-1,jv} is final after this term
                BFrag bVn;
                if (t - 1 >= 1) bVn = to_bfrag(ABZ_B(t - 1, jv), src0, src1, par);
#pragma unroll
                for (int iv = t - 2; iv >= 0; iv--) bmm<true>(ABZ_B(iv, jv), ABZ_B(iv, t), bV);
                if (t - 1 >= 1) bV = bVn;
            }
        }
        // M = L~^-1: every strictly lower block starts as -L (all columns negated up front, so the left operands L_{im,t} of the
        // updates below are read with the opposite sign: C += (-L) b)
#pragma unroll
        for (int jm = 0; jm < NB - 1; jm++)
#pragma unroll
            for (int im = jm + 1; im < NB; im++) {
                R0[im][jm] = dneg(R0[im][jm]); R1[im][jm] = dneg(R1[im][jm]); I0[im][jm] = dneg(I0[im][jm]); I1[im][jm] = dneg(I1[im][jm]);
            }
        BFrag bcol = to_bfrag(ABZ_B(1, 0), src0, src1, par);             // first fragment of column 0 (M_{jm+1,jm} = -L is final from the start)
#pragma unroll
        for (int jm = 0; jm < NB - 1; jm++) {
            BFrag b = bcol;
#pragma unroll
            for (int t = jm + 1; t < NB; t++) {
                tr += R0[jm][t] * b.r[0] - I0[jm][t] * b.i[0] + R1[jm][t] * b.r[1] - I1[jm][t] * b.i[1];
                ti += R0[jm][t] * b.i[0] + I0[jm][t] * b.r[0] + R1[jm][t] * b.i[1] + I1[jm][t] * b.r[1];
                BFrag bn;
                if (t + 1 < NB) {
                    bmm<false>(ABZ_B(t + 1, jm), ABZ_B(t + 1, t), b);    // -= L_{t+1,t} b (the array holds -L); M_{t+1,jm} is final after this term
                    bn = to_bfrag(ABZ_B(t + 1, jm), src0, src1, par);
                }
                if (t == jm + 1 && jm + 2 < NB) bcol = to_bfrag(ABZ_B(jm + 2, jm + 1), src0, src1, par);   // next column's first fragment
#pragma unroll
                for (int im = t + 2; im < NB; im++) bmm<false>(ABZ_B(im, jm), ABZ_B(im, t), b);
                if (t + 1 < NB) b = bn;
            }
        }
#undef ABZ_B
        return make_double2(warp_sum(tr), warp_sum(ti));
    }
    if (VAR == 1) {
        // V = U~^-1, columns right to left (column jv only reads X blocks of columns t < jv, still untouched)
#pragma unroll
        for (int jv = NB - 1; jv >= 1; jv--) {
            const BFrag bD = to_bfrag(R0[jv][jv], R1[jv][jv], I0[jv][jv], I1[jv][jv], src0, src1, par);
#pragma unroll
            for (int iv = 0; iv < jv; iv++) {                            // V'_{iv,jv} = -X_{iv,jv} D_jv
                double cr0 = 0, cr1 = 0, ci0 = 0, ci1 = 0;
                bmm<true>(cr0, cr1, ci0, ci1, R0[iv][jv], R1[iv][jv], I0[iv][jv], I1[iv][jv], bD);
                R0[iv][jv] = cr0; R1[iv][jv] = cr1; I0[iv][jv] = ci0; I1[iv][jv] = ci1;
            }
#pragma unroll
            for (int t = jv - 1; t >= 1; t--) {                          // V_{t,jv} is final: apply it to the rows above
                const BFrag bV = to_bfrag(R0[t][jv], R1[t][jv], I0[t][jv], I1[t][jv], src0, src1, par);
#pragma unroll
                for (int iv = 0; iv < t; iv++)
                    bmm<true>(R0[iv][jv], R1[iv][jv], I0[iv][jv], I1[iv][jv], R0[iv][t], R1[iv][t], I0[iv][t], I1[iv][t], bV);
            }
        }
        // M = L~^-1, columns left to right; tr += tr(V_{jm,t} M_{t,jm}) as soon as M_{t,jm} is final
#pragma unroll
        for (int jm = 0; jm < NB - 1; jm++) {
#pragma unroll
            for (int im = jm + 1; im < NB; im++) {
                R0[im][jm] = dneg(R0[im][jm]); R1[im][jm] = dneg(R1[im][jm]); I0[im][jm] = dneg(I0[im][jm]); I1[im][jm] = dneg(I1[im][jm]);
            }
#pragma unroll
            for (int t = jm + 1; t < NB; t++) {
                const BFrag b = to_bfrag(R0[t][jm], R1[t][jm], I0[t][jm], I1[t][jm], src0, src1, par);
                tr += R0[jm][t] * b.r[0] - I0[jm][t] * b.i[0] + R1[jm][t] * b.r[1] - I1[jm][t] * b.i[1];
                ti += R0[jm][t] * b.i[0] + I0[jm][t] * b.r[0] + R1[jm][t] * b.i[1] + I1[jm][t] * b.r[1];
#pragma unroll
                for (int im = t + 1; im < NB; im++)
                    bmm<true>(R0[im][jm], R1[im][jm], I0[im][jm], I1[im][jm], R0[im][t], R1[im][t], I0[im][t], I1[im][t], b);
            }
        }
        return make_double2(warp_sum(tr), warp_sum(ti));
    }
    // ---- V = U~^-1 (strict upper blocks, in place over X; columns right to left, rows bottom up) and
    //      M = L~^-1 (strict lower blocks, in place over L; columns left to right, rows top down).
    //      The two substitutions are independent dependency chains: their steps are interleaved.
    {
        BFrag bDj, bV[NB], bM[NB];
#pragma unroll
        for (int jm = 0; jm < NB - 1; jm++) {
            const int jv = NB - 1 - jm;
            bDj = to_bfrag(R0[jv][jv], R1[jv][jv], I0[jv][jv], I1[jv][jv], src0, src1, par);
#pragma unroll
            for (int im = jm + 1; im < NB; im++) {
                const int iv = NB - 1 - im;
                {   // V_{iv,jv} = -X_{iv,jv} D_jv - sum_{iv<t<jv} X_{iv,t} V_{t,jv}
                    double cr0 = 0, cr1 = 0, ci0 = 0, ci1 = 0;
                    bmm<true>(cr0, cr1, ci0, ci1, R0[iv][jv], R1[iv][jv], I0[iv][jv], I1[iv][jv], bDj);
#pragma unroll
                    for (int t = iv + 1; t < jv; t++)
                        bmm<true>(cr0, cr1, ci0, ci1, R0[iv][t], R1[iv][t], I0[iv][t], I1[iv][t], bV[t]);
                    R0[iv][jv] = cr0; R1[iv][jv] = cr1; I0[iv][jv] = ci0; I1[iv][jv] = ci1;
                    if (iv > 0) bV[iv] = to_bfrag(cr0, cr1, ci0, ci1, src0, src1, par);
                }
                {   // M_{im,jm} = -L_{im,jm} - sum_{jm<t<im} L_{im,t} M_{t,jm}
                    double cr0 = -R0[im][jm], cr1 = -R1[im][jm], ci0 = -I0[im][jm], ci1 = -I1[im][jm];
#pragma unroll
                    for (int t = jm + 1; t < im; t++)
                        bmm<true>(cr0, cr1, ci0, ci1, R0[im][t], R1[im][t], I0[im][t], I1[im][t], bM[t]);
                    R0[im][jm] = cr0; R1[im][jm] = cr1; I0[im][jm] = ci0; I1[im][jm] = ci1;
                    if (im < NB - 1) bM[im] = to_bfrag(cr0, cr1, ci0, ci1, src0, src1, par);
                }
            }
        }
    }
    // ---- tr += sum_{j<i} tr(V_ji M_ij) = sum over lanes of V_ji[g][2q+t] * M_ij[2q+t][g] (B fragment of M_ij)
#pragma unroll
    for (int j = 0; j < NB - 1; j++)
#pragma unroll
        for (int i = j + 1; i < NB; i++) {
            const BFrag b = to_bfrag(R0[i][j], R1[i][j], I0[i][j], I1[i][j], src0, src1, par);
            tr += R0[j][i] * b.r[0] - I0[j][i] * b.i[0] + R1[j][i] * b.r[1] - I1[j][i] * b.i[1];
            ti += R0[j][i] * b.i[0] + I0[j][i] * b.r[0] + R1[j][i] * b.i[1] + I1[j][i] * b.r[1];
        }
    return make_double2(warp_sum(tr), warp_sum(ti));
}

constexpr int MMA_WARPS_MAX = 12;
#ifndef ABZ_MMA_DEFAULT_VARIANT
#define ABZ_MMA_DEFAULT_VARIANT 1    // measured on the C4 workload: 785.7 k vs 781.8 k k-points/s (profiles/r02_k3fast_variants.log)
#endif

// CTA c handles k-points [c*kper, (c+1)*kper) x all nw frequencies; warp w takes the (k, w) pairs
// i = w, w + 8, ... of that chunk.  mode 0: outp[c*nw + w] = sum_k wnode_k tr ; mode 1: outp[k*nw + w] = tr
// shared: acc[MMA_WARPS][nw] double2
template <int NB, int MMA_WARPS, int VAR = 0>
__global__ void __launch_bounds__(MMA_WARPS * 32, 1)
resolvent_mma_kernel(const double2* __restrict__ H, const double* __restrict__ wnode, long nk, int n, int nw,
                     const double2* __restrict__ z, const double2* __restrict__ sigma, int kper, int mode,
                     double2* __restrict__ outp, int* __restrict__ errflag) {
    extern __shared__ double2 mma_acc[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    double2* acc = mma_acc + (long)warp * nw;
    if (mode == 0)
        for (int w = lane; w < nw; w += 32) acc[w] = make_double2(0.0, 0.0);
    __syncwarp();
    const long k0 = (long)blockIdx.x * kper;
    const long k1 = k0 + kper < nk ? k0 + kper : nk;
    const long npairs = (k1 - k0) * nw;
    const int npad = 8 * NB - n;
    for (long it = warp; it < npairs; it += MMA_WARPS) {
        const long k = k0 + it / nw;
        const int w = (int)(it % nw);
        const double2* Hk = H + k * (long)n * n;
        const double2* sg = sigma ? sigma + (long)w * n * n : nullptr;
        const double2 zz = z[w];
        // B = H + Sigma - z (= -A): no FP64 work off the diagonal; tr A^-1 = -tr B^-1
        double R0[NB][NB], R1[NB][NB], I0[NB][NB], I1[NB][NB];
        int amaxhi = 0;     // max over entries of max(|re|, |im|), tracked through the high word (integer pipe)
#pragma unroll
        for (int bj = 0; bj < NB; bj++)
#pragma unroll
            for (int bi = 0; bi < NB; bi++) {
                const int row = 8 * bi + g, c0 = 8 * bj + 2 * q, c1 = c0 + 1;
                double2 a0 = make_double2(0.0, 0.0), a1 = a0;
                if (row < n && c0 < n) {
                    a0 = Hk[row + (long)c0 * n];
                    if (sg) { double2 s0 = sg[row + (long)c0 * n]; a0.x += s0.x; a0.y += s0.y; }
                }
                if (row < n && c1 < n) {
                    a1 = Hk[row + (long)c1 * n];
                    if (sg) { double2 s1 = sg[row + (long)c1 * n]; a1.x += s1.x; a1.y += s1.y; }
                }
                if (bi == bj) {
                    if (row == c0) { if (row < n) { a0.x -= zz.x; a0.y -= zz.y; } else { a0.x = -1.0; a0.y = 0.0; } }
                    if (row == c1) { if (row < n) { a1.x -= zz.x; a1.y -= zz.y; } else { a1.x = -1.0; a1.y = 0.0; } }
                }
                R0[bi][bj] = a0.x; I0[bi][bj] = a0.y; R1[bi][bj] = a1.x; I1[bi][bj] = a1.y;
                amaxhi = max(amaxhi, max(max(__double2hiint(a0.x) & 0x7fffffff, __double2hiint(a0.y) & 0x7fffffff),
                                         max(__double2hiint(a1.x) & 0x7fffffff, __double2hiint(a1.y) & 0x7fffffff)));
            }
        int minhi = 0x7ff00000;
        double2 t;
        if constexpr (NB == 4 && VAR == 2) t = warp_trace_inverse_pipelined(R0, R1, I0, I1, lane, minhi);
        else if constexpr (VAR == 3) t = warp_trace_inverse_paired<NB>(R0, R1, I0, I1, lane, minhi);
        else if constexpr (VAR == 4) {
            t = warp_trace_inverse<NB, 1, true>(R0, R1, I0, I1, lane, minhi);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) minhi = min(minhi, __shfl_xor_sync(0xffffffffu, minhi, off));   // kept per row
        } else t = warp_trace_inverse<NB, (VAR == 2 ? 1 : VAR)>(R0, R1, I0, I1, lane, minhi);
        t.x = -t.x - (double)npad; t.y = -t.y;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) amaxhi = max(amaxhi, __shfl_xor_sync(0xffffffffu, amaxhi, off));
        if (lane == 0) {
            // smallest |pivot|^2 against (2e-3 * max|a|)^2 (both within 2^-20 through their high words)
            const double pmin2 = __hiloint2double(minhi, 0), am = __hiloint2double(amaxhi, 0);
            if (!(pmin2 > 4e-6 * am * am) || !(isfinite(t.x) && isfinite(t.y))) atomicOr(errflag, 2);   // ask for the pivoted path
            if (mode == 0) {
                const double wt = wnode ? wnode[k] : 1.0;
                acc[w].x += wt * t.x; acc[w].y += wt * t.y;
            } else {
                outp[k * nw + w] = t;
            }
        }
    }
    if (mode == 0) {
        __syncthreads();
        for (int w = threadIdx.x; w < nw; w += MMA_WARPS * 32) {
            double sx = 0.0, sy = 0.0;
#pragma unroll
            for (int wp = 0; wp < MMA_WARPS; wp++) { double2 v = mma_acc[(long)wp * nw + w]; sx += v.x; sy += v.y; }
            outp[(long)blockIdx.x * nw + w] = make_double2(sx, sy);
        }
    }
}

// K3-fused: the same one-warp-per-matrix block LU with the innermost contraction stage folded in, so that H(k) never exists in
// HBM (FourierSeriesEvaluators evaluates H(k) and hands it to the integrand, src/fourier.jl:127-164; here the CTA is that hand-over).
// A CTA walks its nodes k; for each node all 256 threads form H(k) = sum_m C1[row(k)][m] e^{2 pi i k1 R_m} (the stage-1 sum, 4 matrix
// entries per thread, C1 and the phase table from L2) into one of two shared-memory buffers (leading dimension n | 1: the C-layout
// loads of a quarter-warp then fall into 8 different 16-byte bank groups), and the 8 warps take the frequencies w = warp, warp + 8, ...
// of that node from the buffer while the next node's H is formed into the other one (one block barrier per node).  Per node the
// contraction costs 4 M complex FMAs per thread - for nw = 128 that is 17 FP64 instructions per matrix against ~1150 - and reads the
// M n^2 coefficients of its row from L2; what disappears is the 16 n^2 B per node written and read back by the separate stage-1 kernel.
// Weighted sums only (mode 0), nw >= MMA_WARPS.
// FULL: norb == 8 NB (no padding rows or columns: the guards of the ragged case are compiled out)
// DIRECT: materialised rule - C1 IS H(k), node-major from node n0 on, and the contraction is a copy into the shared-memory buffer
template <int NB, int MMA_WARPS, int VAR, bool FULL, int PF = 0, bool DIRECT = false>
__global__ void __launch_bounds__(MMA_WARPS * 32, 1)
resolvent_mma_fused_kernel(const double2* __restrict__ C1, const double2* __restrict__ ptab, const long* __restrict__ row_nodeptr,
                           long r0, long r1, const int* __restrict__ klist, int N, int M, const double* __restrict__ wnode, long n0,
                           long nk, int n, int nw, const double2* __restrict__ z, const double2* __restrict__ sigma, int kper,
                           double2* __restrict__ outp, int* __restrict__ errflag) {
    extern __shared__ double2 mma_acc[];                 // acc[MMA_WARPS][nw] | sH[2][n * LDH] | amax[2][MMA_WARPS] (int)
    constexpr int NT = MMA_WARPS * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const int LDH = n | 1;
    const int nn = n * n;
    double2* acc = mma_acc + (long)warp * nw;
    double2* sH = mma_acc + (long)MMA_WARPS * nw;
    int* sAmax = reinterpret_cast<int*>(sH + 2L * n * LDH);     // per buffer and warp: max over H(k) of max(|re|, |im|), high words
    for (int w = lane; w < nw; w += 32) acc[w] = make_double2(0.0, 0.0);
    const long ka = n0 + (long)blockIdx.x * kper;
    const long kb = ka + kper < n0 + nk ? ka + kper : n0 + nk;
    long row = r0;
    if constexpr (!DIRECT) {   // row of the first node: row_nodeptr[row] <= ka < row_nodeptr[row + 1]
        long lo = r0, hi = r1;
        while (hi - lo > 1) { const long mid = (lo + hi) >> 1; if (row_nodeptr[mid] <= ka) lo = mid; else hi = mid; }
        row = lo;
    }
    const int npad = FULL ? 0 : 8 * NB - n;
    constexpr int EPT = (64 * NB * NB + NT - 1) / NT;      // matrix entries per thread in the contraction
    auto build = [&](long k, int buf) {
        double2 h[EPT];
#pragma unroll
        for (int i = 0; i < EPT; i++) h[i] = make_double2(0.0, 0.0);
        if constexpr (DIRECT) {
            const double2* src = C1 + (k - n0) * (long)nn;
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                const int e = threadIdx.x + NT * i;
                if (e < nn) h[i] = src[e];
            }
        } else {
        while (row_nodeptr[row + 1] <= k) row++;
        const int k1 = klist ? klist[k] : (int)(k - row_nodeptr[row]);
        const double2* c1 = C1 + (row - r0) * (long)M * nn;
        for (int m = 0; m < M; m++) {
            const double2 ph = ptab[(long)m * N + k1];
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                const int e = threadIdx.x + NT * i;
                if (e < nn) {
                    const double2 a = c1[(long)m * nn + e];
                    h[i].x = fma(a.x, ph.x, h[i].x); h[i].x = fma(-a.y, ph.y, h[i].x);
                    h[i].y = fma(a.x, ph.y, h[i].y); h[i].y = fma(a.y, ph.x, h[i].y);
                }
            }
        }
        }
        double2* dst = sH + (long)buf * n * LDH;
        int hmax = 0;
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int e = threadIdx.x + NT * i;
            if (e < nn) {
                dst[(e % n) + (e / n) * LDH] = h[i];
                hmax = max(hmax, max(__double2hiint(h[i].x) & 0x7fffffff, __double2hiint(h[i].y) & 0x7fffffff));
            }
        }
        hmax = __reduce_max_sync(0xffffffffu, hmax);
        if (lane == 0) sAmax[buf * MMA_WARPS + warp] = hmax;
    };
    if (ka < kb) build(ka, 0);
    __syncthreads();
    for (long k = ka; k < kb; k++) {
        const int buf = (int)((k - ka) & 1);
        if (k + 1 < kb) build(k + 1, buf ^ 1);
        const double2* Hk = sH + (long)buf * n * LDH;
        const double wt = wnode ? wnode[k] : 1.0;
        // the pivot monitor's scale max|a_ij|: once per node from H(k) (the matrix entries are H + Sigma - z: Sigma is added per matrix below)
        const int hmaxk = __reduce_max_sync(0xffffffffu, sAmax[buf * MMA_WARPS + (lane % MMA_WARPS)]);
        for (int w = warp; w < nw; w += MMA_WARPS) {
            const double2* sg = sigma ? sigma + (long)w * nn : nullptr;
            const double2 zz = z[w];
            double R0[NB][NB], R1[NB][NB], I0[NB][NB], I1[NB][NB];
            int amaxhi = max(hmaxk, max(__double2hiint(zz.x) & 0x7fffffff, __double2hiint(zz.y) & 0x7fffffff));
#pragma unroll
            for (int bj = 0; bj < NB; bj++)
#pragma unroll
                for (int bi = 0; bi < NB; bi++) {
                    const int rw = 8 * bi + g, c0 = 8 * bj + 2 * q, c1i = c0 + 1;
                    double2 a0 = make_double2(0.0, 0.0), a1 = a0;
                    if (FULL || (rw < n && c0 < n)) {
                        a0 = Hk[rw + c0 * LDH];
                        if (sg) {
                            double2 s0 = sg[rw + (long)c0 * n]; a0.x += s0.x; a0.y += s0.y;
                            amaxhi = max(amaxhi, max(__double2hiint(a0.x) & 0x7fffffff, __double2hiint(a0.y) & 0x7fffffff));
                        }
                    }
                    if (FULL || (rw < n && c1i < n)) {
                        a1 = Hk[rw + c1i * LDH];
                        if (sg) {
                            double2 s1 = sg[rw + (long)c1i * n]; a1.x += s1.x; a1.y += s1.y;
                            amaxhi = max(amaxhi, max(__double2hiint(a1.x) & 0x7fffffff, __double2hiint(a1.y) & 0x7fffffff));
                        }
                    }
                    if (bi == bj) {
                        if (rw == c0) { if (FULL || rw < n) { a0.x -= zz.x; a0.y -= zz.y; } else { a0.x = -1.0; a0.y = 0.0; } }
                        if (rw == c1i) { if (FULL || rw < n) { a1.x -= zz.x; a1.y -= zz.y; } else { a1.x = -1.0; a1.y = 0.0; } }
                    }
                    R0[bi][bj] = a0.x; I0[bi][bj] = a0.y; R1[bi][bj] = a1.x; I1[bi][bj] = a1.y;
                }
            int minhi = 0x7ff00000;
            double2 t = warp_trace_inverse<NB, VAR, false, PF>(R0, R1, I0, I1, lane, minhi);
            t.x = -t.x - (double)npad; t.y = -t.y;
            if (sg) amaxhi = __reduce_max_sync(0xffffffffu, amaxhi);
            if (lane == 0) {
                const double pmin2 = __hiloint2double(minhi, 0), am = __hiloint2double(amaxhi, 0);
                if (!(pmin2 > 4e-6 * am * am) || !(isfinite(t.x) && isfinite(t.y))) atomicOr(errflag, 2);
                acc[w].x += wt * t.x; acc[w].y += wt * t.y;
            }
        }
        __syncthreads();        // buffer buf may be refilled (node k + 2), buffer buf ^ 1 is complete
    }
    for (int w = threadIdx.x; w < nw; w += NT) {
        double sx = 0.0, sy = 0.0;
#pragma unroll
        for (int wp = 0; wp < MMA_WARPS; wp++) { double2 v = mma_acc[(long)wp * nw + w]; sx += v.x; sy += v.y; }
        outp[(long)blockIdx.x * nw + w] = make_double2(sx, sy);
    }
}

inline bool mma_resolvent_supported(int n) { return n >= 4 && n <= 32; }

// warps per CTA (= per SM): 8 (255 registers) and 12 (168 registers, ~150 spill accesses) measure the same on B200;
// 10 warps cannot have more than 168 registers either (warps are allocated four at a time: a 192-register, 320-thread
// launch is refused with "too many resources requested"), and rebuilding the B fragments of the substitution phase at every
// use instead of keeping bV[] / bM[] does not lower the 12-warp variant's spills (the pressure peak is in the LU phase);
// (the kernel is bound by the FP64 pipe that DMMA and DFMA share, not by occupancy); 8 is the default
inline int mma_resolvent_warps() {
    static int w = 0;
    if (!w) {
        const char* e = getenv("ABZ_MMA_WARPS");
        w = (e && atoi(e) == 12) ? 12 : (e && atoi(e) == 4) ? 4 : 8;      // 4: one warp per sub-partition (measurement hook, norb 25..32)
    }
    return w;
}

inline int mma_resolvent_plan(int n, long nk, int nw, long sm, long* ncta, int* kper) {
    if (!mma_resolvent_supported(n)) return -1;
    const int W = mma_resolvent_warps();
    if ((size_t)nw * W * sizeof(double2) > 160 * 1024) return -1;
    long target = sm * 4;
    long kp = (nk + target - 1) / target;
    // keep every warp busy: at least ~4 matrices per warp per CTA
    while (kp * nw < 4L * W && kp < nk) kp++;
    if (kp < 1) kp = 1;
    *kper = (int)kp;
    *ncta = (nk + kp - 1) / kp;
    return 0;
}

// substitution variant of the norb = 25..32 kernel (ABZ_MMA_VARIANT = 0 / 1, see warp_trace_inverse); smaller matrices use 0
inline int mma_resolvent_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("ABZ_MMA_VARIANT"); v = e ? std::min(4, std::max(0, atoi(e))) : ABZ_MMA_DEFAULT_VARIANT; }
    return v;
}

template <int NB, int W>
inline void mma_launch_one(const double2* H, const double* wnode, long nk, int n, int nw, const double2* z, const double2* sigma,
                           int mode, double2* outp, int* errflag, long ncta, int kper, cudaStream_t stream) {
    size_t smem = (size_t)nw * W * sizeof(double2);
    if (W == 4) smem = std::max<size_t>(smem, 120 * 1024);      // measurement hook: keeps a second 4-warp CTA off the SM
    if (NB == 4 && mma_resolvent_variant() == 4)
        resolvent_mma_kernel<NB, W, (NB == 4 ? 4 : 0)><<<(unsigned)ncta, W * 32, smem, stream>>>(H, wnode, nk, n, nw, z, sigma, kper, mode, outp, errflag);
    else if (NB == 4 && mma_resolvent_variant() == 3)
        resolvent_mma_kernel<NB, W, (NB == 4 ? 3 : 0)><<<(unsigned)ncta, W * 32, smem, stream>>>(H, wnode, nk, n, nw, z, sigma, kper, mode, outp, errflag);
    else if (NB == 4 && mma_resolvent_variant() == 2)
        resolvent_mma_kernel<NB, W, (NB == 4 ? 2 : 0)><<<(unsigned)ncta, W * 32, smem, stream>>>(H, wnode, nk, n, nw, z, sigma, kper, mode, outp, errflag);
    else if (NB == 4 && mma_resolvent_variant() == 1)
        resolvent_mma_kernel<NB, W, (NB == 4 ? 1 : 0)><<<(unsigned)ncta, W * 32, smem, stream>>>(H, wnode, nk, n, nw, z, sigma, kper, mode, outp, errflag);
    else
        resolvent_mma_kernel<NB, W, 0><<<(unsigned)ncta, W * 32, smem, stream>>>(H, wnode, nk, n, nw, z, sigma, kper, mode, outp, errflag);
}

inline size_t mma_fused_smem(int n, int nw, int W = 12) { return ((size_t)nw * W + 2 * (size_t)n * (n | 1)) * sizeof(double2) + 2 * W * sizeof(int); }
// nodes [n0, n0 + nk) of rows [r0, r1) of a rule; C1 holds the contracted series of those rows (row r0 first)
inline cudaError_t mma_fused_launch(const double2* C1, const double2* ptab, const long* row_nodeptr, long r0, long r1, const int* klist,
                                    int N, int M, const double* wnode, long n0, long nk, int n, int nw, const double2* z,
                                    const double2* sigma, double2* outp, int* errflag, long ncta, int kper, cudaStream_t stream) {
    const int W = mma_resolvent_warps() == 12 && n > 24 && ptab ? 12 : 8;
    const size_t smem = mma_fused_smem(n, nw, W);
    if (!ptab) {            // direct mode (materialised H)
#define ABZ_DIRECT_GO(NBX, V, F, P)                                                                                                  \
    resolvent_mma_fused_kernel<NBX, 8, V, F, P, true><<<(unsigned)ncta, 256, smem, stream>>>(C1, ptab, row_nodeptr, r0, r1, klist, N, M, \
                                                                                            wnode, n0, nk, n, nw, z, sigma, kper,   \
                                                                                            outp, errflag)
        switch ((n + 7) / 8) {
            case 1: if (n == 8) ABZ_DIRECT_GO(1, 0, true, 0); else ABZ_DIRECT_GO(1, 0, false, 0); break;
            case 2: if (n == 16) ABZ_DIRECT_GO(2, 0, true, 0); else ABZ_DIRECT_GO(2, 0, false, 0); break;
            case 3: if (n == 24) ABZ_DIRECT_GO(3, 0, true, 0); else ABZ_DIRECT_GO(3, 0, false, 0); break;
            case 4: if (n == 32) ABZ_DIRECT_GO(4, 1, true, 1); else ABZ_DIRECT_GO(4, 1, false, 0); break;
            default: return cudaErrorInvalidValue;
        }
#undef ABZ_DIRECT_GO
        return cudaGetLastError();
    }
#define ABZ_FUSED_GO(NBX, WX, V, F)                                                                                                 \
    resolvent_mma_fused_kernel<NBX, WX, V, F><<<(unsigned)ncta, WX * 32, smem, stream>>>(C1, ptab, row_nodeptr, r0, r1, klist, N, M, \
                                                                                        wnode, n0, nk, n, nw, z, sigma, kper, outp, \
                                                                                        errflag)
    switch ((n + 7) / 8) {
        case 1: if (n == 8) ABZ_FUSED_GO(1, 8, 0, true); else ABZ_FUSED_GO(1, 8, 0, false); break;
        case 2: if (n == 16) ABZ_FUSED_GO(2, 8, 0, true); else ABZ_FUSED_GO(2, 8, 0, false); break;
        case 3: if (n == 24) ABZ_FUSED_GO(3, 8, 0, true); else ABZ_FUSED_GO(3, 8, 0, false); break;
        case 4:
            if (W == 12) { if (n == 32) ABZ_FUSED_GO(4, 12, 1, true); else ABZ_FUSED_GO(4, 12, 1, false); }
            else if (n == 32) {
                // U fragments one product group ahead (PF = 1): 908.6 -> 919.3 k k-points/s on the C4 workload; also prefetching D_{s+1}
                // (PF = 2) gives the gain back (907.1 k) - profiles/r02_k3fused_timing.log.  ABZ_MMA_PREFETCH = 0 / 2 select the others.
                static const int pf = getenv("ABZ_MMA_PREFETCH") ? atoi(getenv("ABZ_MMA_PREFETCH")) : 1;
                if (pf == 0) ABZ_FUSED_GO(4, 8, 1, true);
                else if (pf == 3) resolvent_mma_fused_kernel<4, 8, 1, true, 3><<<(unsigned)ncta, 256, smem, stream>>>(C1, ptab, row_nodeptr, r0, r1, klist, N, M, wnode, n0, nk, n, nw, z, sigma, kper, outp, errflag);
                else if (pf == 2) resolvent_mma_fused_kernel<4, 8, 1, true, 2><<<(unsigned)ncta, 256, smem, stream>>>(C1, ptab, row_nodeptr, r0, r1, klist, N, M, wnode, n0, nk, n, nw, z, sigma, kper, outp, errflag);
                else resolvent_mma_fused_kernel<4, 8, 1, true, 1><<<(unsigned)ncta, 256, smem, stream>>>(C1, ptab, row_nodeptr, r0, r1, klist, N, M, wnode, n0, nk, n, nw, z, sigma, kper, outp, errflag);
            } else ABZ_FUSED_GO(4, 8, 1, false);
            break;
        default: return cudaErrorInvalidValue;
    }
#undef ABZ_FUSED_GO
    return cudaGetLastError();
}

// dynamic shared memory above 48 KB is a per-device opt-in: called for the current device at context creation
inline cudaError_t mma_resolvent_opt_in() {
    cudaError_t e = cudaSuccess;
    auto set = [&](const void* f) {
        cudaError_t r = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (r != cudaSuccess && e == cudaSuccess) e = r;
    };
#define ABZ_MMA_OPT(NBX) { auto k8 = resolvent_mma_kernel<NBX, 8, 0>; set((const void*)k8); auto k12 = resolvent_mma_kernel<NBX, 12, 0>; set((const void*)k12); }
    ABZ_MMA_OPT(1) ABZ_MMA_OPT(2) ABZ_MMA_OPT(3) ABZ_MMA_OPT(4)
#undef ABZ_MMA_OPT
    { auto k8 = resolvent_mma_kernel<4, 8, 1>; set((const void*)k8); auto k12 = resolvent_mma_kernel<4, 12, 1>; set((const void*)k12); }
    { auto k8 = resolvent_mma_kernel<4, 8, 2>; set((const void*)k8); auto k12 = resolvent_mma_kernel<4, 12, 2>; set((const void*)k12); }
    { auto k8 = resolvent_mma_kernel<4, 8, 3>; set((const void*)k8); auto k12 = resolvent_mma_kernel<4, 12, 3>; set((const void*)k12); }
    { auto k8 = resolvent_mma_kernel<4, 8, 4>; set((const void*)k8); auto k12 = resolvent_mma_kernel<4, 12, 4>; set((const void*)k12); auto k4 = resolvent_mma_kernel<4, 4, 4>; set((const void*)k4); }
#define ABZ_FUSED_OPT(NBX, WX, V) { auto f = resolvent_mma_fused_kernel<NBX, WX, V, false>; set((const void*)f); auto t = resolvent_mma_fused_kernel<NBX, WX, V, true>; set((const void*)t); }
    ABZ_FUSED_OPT(1, 8, 0) ABZ_FUSED_OPT(2, 8, 0) ABZ_FUSED_OPT(3, 8, 0) ABZ_FUSED_OPT(4, 8, 1) ABZ_FUSED_OPT(4, 12, 1)
#undef ABZ_FUSED_OPT
    { auto f = resolvent_mma_fused_kernel<4, 8, 1, true, 1>; set((const void*)f); auto t = resolvent_mma_fused_kernel<4, 8, 1, true, 2>; set((const void*)t); }
    { auto f = resolvent_mma_fused_kernel<4, 8, 1, true, 3>; set((const void*)f); }
#define ABZ_DIRECT_OPT(NBX, V, P) { auto f = resolvent_mma_fused_kernel<NBX, 8, V, false, 0, true>; set((const void*)f); auto t = resolvent_mma_fused_kernel<NBX, 8, V, true, P, true>; set((const void*)t); }
    ABZ_DIRECT_OPT(1, 0, 0) ABZ_DIRECT_OPT(2, 0, 0) ABZ_DIRECT_OPT(3, 0, 0) ABZ_DIRECT_OPT(4, 1, 1)
#undef ABZ_DIRECT_OPT
    { auto k4 = resolvent_mma_kernel<4, 4, 0>; set((const void*)k4); auto k41 = resolvent_mma_kernel<4, 4, 1>; set((const void*)k41); }
    { auto k4 = resolvent_mma_kernel<4, 4, 2>; set((const void*)k4); auto k41 = resolvent_mma_kernel<4, 4, 3>; set((const void*)k41); }
    return e;
}

inline cudaError_t mma_resolvent_launch(const double2* H, const double* wnode, long nk, int n, int nw, const double2* z,
                                        const double2* sigma, int mode, double2* outp, int* errflag, long ncta, int kper,
                                        cudaStream_t stream) {
    const int NBv = (n + 7) / 8;
    const int W = mma_resolvent_warps();
#define ABZ_MMA_CASE(NBX)                                                                                            \
    if (W == 8) mma_launch_one<NBX, 8>(H, wnode, nk, n, nw, z, sigma, mode, outp, errflag, ncta, kper, stream);      \
    else mma_launch_one<NBX, 12>(H, wnode, nk, n, nw, z, sigma, mode, outp, errflag, ncta, kper, stream);
    switch (NBv) {
        case 1: ABZ_MMA_CASE(1) break;
        case 2: ABZ_MMA_CASE(2) break;
        case 3: ABZ_MMA_CASE(3) break;
        default:
            if (W == 4) mma_launch_one<4, 4>(H, wnode, nk, n, nw, z, sigma, mode, outp, errflag, ncta, kper, stream);
            else { ABZ_MMA_CASE(4) }
            break;
    }
#undef ABZ_MMA_CASE
    return cudaGetLastError();
}

}  // namespace abz
