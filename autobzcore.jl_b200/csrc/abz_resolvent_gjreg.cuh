// K3-generic: (z - H(k) - Sigma_w)^-1 by in-place Gauss-Jordan elimination with partial pivoting (the robustness of
// LAPACK getrf/getri, i.e. Julia's `inv(::Matrix)`), with the matrix resident in REGISTERS: a team of NP x (NP/CW)
// threads owns one matrix, thread (r, h) holding row r, columns CW h .. CW h + CW - 1.  Pivoting is implicit: the pivot
// row of step p is chosen among the rows not used yet (one REDUX over keys made of the high words of |re| + |im|) and stays where it is (no row
// exchange), the permutation is undone when the trace / the inverse is read out.  Per step the team exchanges one
// column (multipliers and pivot candidates) and one row (the pivot row) through double-buffered shared memory: two
// barriers, CW broadcast 128-bit shared loads (fetched LD entries ahead of their FMAs) and CW complex FMAs per thread.
// NP = 32 serves norb <= 32 (the fallback of the unpivoted DMMA kernel, and the matrix-valued sums), NP = 64 serves
// 32 < norb <= 64.  Replaces the shared-memory formulation (one warp per matrix, three shared accesses per complex
// FMA, 64 KB of shared memory per matrix at norb = 64), kept as ABZ_OPT_RESOLVENT_ALGO = 4.
#pragma once
#include "abz_common.cuh"

namespace abz {

template <int NP>
struct GjShared {
    unsigned key[2][NP];     // pivot candidates: high word of the double |re| + |im| (low 6 bits: 63 - row), 0 = row already used / padding
    double2 col[2][NP];      // the current column (multipliers)
    double2 prow[2][NP];     // the pivot row (in the register order of its segment)
    int piv[NP];             // piv[p] = physical row of logical row p
    int inv[NP];             // inverse permutation
    double2 red[8];          // cross-warp trace reduction
};

template <int T>
__device__ __forceinline__ void gj_team_sync() {
    if (T == 32) __syncwarp(); else __syncthreads();
}

// 1/a: branch-free reciprocal of |a|^2, Smith's algorithm when |a|^2 leaves the normal range
__device__ __forceinline__ double2 gj_crecip(double2 a) {
    const double d = fma(a.x, a.x, a.y * a.y);
    if (d > 1e-280 && d < 1e280) {
        const double s = fast_rcp(d);
        return make_double2(a.x * s, -a.y * s);
    }
    return crecip(a);
}

// register j of segment h holds column CW h + (j + gj_rot(n, h)) % CW after the elimination (see gj_reg_invert)
template <int CW>
__device__ __forceinline__ int gj_rot(int n, int h) {
    const int left = n - CW * h;                 // real columns in this segment
    return left >= CW ? 0 : (left > 0 ? left : 0);
}

// In-place inverse of the team's matrix.  The p loop is rolled: the segment that holds column p keeps it in register 0
// and rotates its registers by one per step (the update writes x[j-1] = x[j] - f * pivot_row[j]; the finished column of the
// inverse enters at register CW-1), so every register index is static without unrolling NP steps.
// On return x holds B = (P A)^-1 row-wise in the physical rows (physical row piv[i] = logical row i) with
// A^-1[i][piv[m]] = B[i][m], columns rotated by gj_rot.  Returns false for a singular / non-finite matrix.
template <int NP, int CW, int LD>
__device__ __forceinline__ bool gj_reg_invert(double2 (&x)[CW], int n, int r, int h, GjShared<NP>& sh) {
    constexpr int T = NP * (NP / CW);
    const int lane = threadIdx.x & 31;
    bool used = (r >= n);                    // padding rows (identity) are never pivots
    if (h == 0) { sh.piv[r] = r; sh.inv[r] = r; }
    for (int p = 0; p < n; p++) {
        const int b = p & 1;
        const bool cur = (NP == CW) || (h == p / CW);
        if (cur) {
            const double2 a = x[0];
            const unsigned hi = (unsigned)__double2hiint(fabs(a.x) + fabs(a.y));   // sign 0: ordered like the value
            sh.key[b][r] = used ? 0u : ((hi & ~63u) | (unsigned)(63 - r));
            sh.col[b][r] = a;
        }
        gj_team_sync<T>();
        // every warp finds the pivot row redundantly: largest candidate, ties to the smaller row (one REDUX)
        unsigned key = sh.key[b][lane];
        if (NP == 64) key = max(key, sh.key[b][lane + 32]);
        key = __reduce_max_sync(0xffffffffu, key);
        if ((key & ~63u) == 0u || key >= 0x7ff00000u) return false;            // uniform: singular, Inf or NaN
        const int bi = 63 - (int)(key & 63u);
        const bool me = (r == bi);
        if (me) {
#pragma unroll
            for (int j = 0; j < CW; j++) sh.prow[b][h * CW + j] = x[j];
            if (cur) { sh.piv[p] = bi; sh.inv[bi] = p; }
#pragma unroll
            for (int j = 0; j < CW; j++) x[j] = make_double2(0.0, 0.0);
            used = true;
        }
        const double2 rp = gj_crecip(sh.col[b][bi]);                           // every thread: 1 / pivot
        const double2 m = sh.col[b][r];
        gj_team_sync<T>();
        // rows other than the pivot row: x -= (m / pivot) * pivot_row; the pivot row (zeroed above): x = pivot_row / pivot
        const double2 f = me ? make_double2(-rp.x, -rp.y) : cmul(m, rp);
        const double2* pr = sh.prow[b] + h * CW;
        double2 t[LD];
        if (cur) {
#pragma unroll
            for (int j0 = 0; j0 < CW; j0 += LD) {
#pragma unroll
                for (int j = 0; j < LD; j++) t[j] = pr[j0 + j];
#pragma unroll
                for (int j = 0; j < LD; j++)
                    if (j0 + j > 0) x[j0 + j - 1] = cfnma(x[j0 + j], f, t[j]);
            }
            x[CW - 1] = me ? rp : make_double2(-f.x, -f.y);                    // column p of the inverse so far
        } else {
#pragma unroll
            for (int j0 = 0; j0 < CW; j0 += LD) {
#pragma unroll
                for (int j = 0; j < LD; j++) t[j] = pr[j0 + j];
#pragma unroll
                for (int j = 0; j < LD; j++) x[j0 + j] = cfnma(x[j0 + j], f, t[j]);
            }
        }
    }
    gj_team_sync<T>();
    return true;
}

// team's matrix A = z - H(k) - Sigma, identity padding; all loads are issued unconditionally (clamped addresses) so that
// they are in flight together
template <int CW>
__device__ __forceinline__ void gj_reg_load(double2 (&x)[CW], const double2* __restrict__ Hk, const double2* __restrict__ sg, double2 zz,
                                            int n, int r, int h) {
    const int rc = r < n ? r : n - 1;
#pragma unroll
    for (int c = 0; c < CW; c++) {
        const int j = h * CW + c;
        const int jc = j < n ? j : n - 1;
        x[c] = __ldg(Hk + rc + jc * n);
    }
    if (sg) {
#pragma unroll
        for (int c = 0; c < CW; c++) {
            const int j = h * CW + c;
            const int jc = j < n ? j : n - 1;
            const double2 s = __ldg(sg + rc + jc * n);
            x[c].x += s.x; x[c].y += s.y;
        }
    }
#pragma unroll
    for (int c = 0; c < CW; c++) {
        const int j = h * CW + c;
        const bool real = (r < n) && (j < n);
        const double dr = (r == j) ? (real ? zz.x : 1.0) : 0.0, di = (r == j && real) ? zz.y : 0.0;
        x[c] = make_double2(real ? dr - x[c].x : dr, real ? di - x[c].y : di);
    }
}

// tr A^-1 = sum_m B[piv[m]][m]: physical row rho holds logical row i = inv[rho]; it owns the term of column m = inv[i]
template <int NP, int CW>
__device__ __forceinline__ double2 gj_reg_trace(const double2 (&x)[CW], int n, int r, int h, GjShared<NP>& sh) {
    constexpr int T = NP * (NP / CW);
    double2 t = make_double2(0.0, 0.0);
    if (r < n) {
        const int m = sh.inv[sh.inv[r]];
        if (m / CW == h) {
            const int jc = ((m % CW) - gj_rot<CW>(n, h) + CW) % CW;
#pragma unroll
            for (int j = 0; j < CW; j++) if (j == jc) t = x[j];
        }
    }
    t.x = warp_sum(t.x); t.y = warp_sum(t.y);
    if (T > 32) {
        const int warp = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) sh.red[warp] = t;
        __syncthreads();
        t = make_double2(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < T / 32; i++) { t.x += sh.red[i].x; t.y += sh.red[i].y; }
    }
    return t;
}

// Trace kernel.  grid = (node chunks of kper, frequency chunks of wper <= 64), block = one team.
// mode 0: partial[chunk * nw + w] = sum over the chunk's nodes of wnode * trace; mode 1: y[k * nw + w] = trace.
template <int NP, int CW, int LD, int MINB>
__global__ void __launch_bounds__(NP * (NP / CW), MINB)
resolvent_gjreg_kernel(const double2* __restrict__ H, const double* __restrict__ wnode, long nk, int n, int nw,
                       const double2* __restrict__ z, const double2* __restrict__ sigma, int kper, int wper, int mode,
                       double2* __restrict__ outp, int* __restrict__ errflag) {
    constexpr int T = NP * (NP / CW);
    __shared__ GjShared<NP> sh;
    __shared__ double2 sAcc[64];
    const int r = threadIdx.x % NP, h = threadIdx.x / NP;
    const long nn = (long)n * n;
    const long k0 = (long)blockIdx.x * kper;
    const long k1 = k0 + kper < nk ? k0 + kper : nk;
    const int w0 = blockIdx.y * wper;
    const int w1 = w0 + wper < nw ? w0 + wper : nw;
    for (int i = threadIdx.x; i < 64; i += T) sAcc[i] = make_double2(0.0, 0.0);
    gj_team_sync<T>();
    double2 x[CW];
    for (long k = k0; k < k1; k++) {
        const double wt = wnode ? wnode[k] : 1.0;
        for (int w = w0; w < w1; w++) {
            gj_reg_load<CW>(x, H + k * nn, sigma ? sigma + (long)w * nn : nullptr, z[w], n, r, h);
            const bool ok = gj_reg_invert<NP, CW, LD>(x, n, r, h, sh);
            double2 t = make_double2(nan(""), nan(""));
            if (ok) t = gj_reg_trace<NP, CW>(x, n, r, h, sh);
            if (threadIdx.x == 0) {
                if (!ok || !(isfinite(t.x) && isfinite(t.y))) *errflag = 1;
                if (mode == 0) { sAcc[w - w0].x += wt * t.x; sAcc[w - w0].y += wt * t.y; }
                else outp[k * nw + w] = t;
            }
            gj_team_sync<T>();
        }
    }
    if (mode == 0)
        for (int i = threadIdx.x; i < w1 - w0; i += T) outp[(long)blockIdx.x * nw + w0 + i] = sAcc[i];
}

// Matrix-valued sum: partial[(chunk * nw + w) * n * n + e] = sum over the chunk's nodes of wnode * [(z_w - H(k) - Sigma_w)^-1]_e
// (the docs' gloc_integrand, docs/src/examples.md:20,90).  grid = (node chunks, nw); the team accumulates in shared memory
// (dynamic: n * n complex), in node order: deterministic.
template <int NP, int CW, int LD, int MINB>
__global__ void __launch_bounds__(NP * (NP / CW), MINB)
resolvent_gjreg_matrix_kernel(const double2* __restrict__ H, const double* __restrict__ wnode, long nk, int n, int nw,
                              const double2* __restrict__ z, const double2* __restrict__ sigma, int kper,
                              double2* __restrict__ partial, int* __restrict__ errflag) {
    constexpr int T = NP * (NP / CW);
    __shared__ GjShared<NP> sh;
    extern __shared__ double2 gj_acc[];
    const int r = threadIdx.x % NP, h = threadIdx.x / NP;
    const int nn = n * n;
    const int w = blockIdx.y;
    const double2 zz = z[w];
    const double2* sg = sigma ? sigma + (long)w * nn : nullptr;
    for (int e = threadIdx.x; e < nn; e += T) gj_acc[e] = make_double2(0.0, 0.0);
    gj_team_sync<T>();
    const long k0 = (long)blockIdx.x * kper;
    const long k1 = k0 + kper < nk ? k0 + kper : nk;
    double2 x[CW];
    for (long k = k0; k < k1; k++) {
        gj_reg_load<CW>(x, H + k * nn, sg, zz, n, r, h);
        const bool ok = gj_reg_invert<NP, CW, LD>(x, n, r, h, sh);
        if (!ok) { if (threadIdx.x == 0) *errflag = 1; gj_team_sync<T>(); continue; }
        const double wt = wnode ? wnode[k] : 1.0;
        if (r < n) {
            const int i = sh.inv[r];
            const int rot = gj_rot<CW>(n, h);
#pragma unroll
            for (int c = 0; c < CW; c++) {
                const int m = h * CW + (c + rot) % CW;
                if (m < n) {
                    double2* dst = gj_acc + i + sh.piv[m] * n;       // a permutation of a column across the lanes: conflict-free
                    double2 a = *dst;
                    a.x += wt * x[c].x; a.y += wt * x[c].y;
                    *dst = a;
                }
            }
        }
        gj_team_sync<T>();
    }
    double2* out = partial + ((long)blockIdx.x * nw + w) * nn;
    for (int e = threadIdx.x; e < nn; e += T) out[e] = gj_acc[e];
}

}  // namespace abz
