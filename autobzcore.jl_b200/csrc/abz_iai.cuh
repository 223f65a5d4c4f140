// K2: scattered-node kernels for IAI panels (src/fourier.jl:432-486): the same nested
// one-dimension-at-a-time contraction as the grid path, but at arbitrary x per work item.
// Contracted series live in an arena of slots on the device so that a level-synchronous batch
// of GK panels (15 nodes each) is three launches regardless of how many panels are live.
#pragma once
#include "abz_common.cuh"
#include "abz_kernels.cuh"

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

// innermost closure for norb <= 3: evaluate the 1-D series of slot1[i] at x1[i] and apply the integrand
// (workspace_evaluate! + f.f(v, p), src/fourier.jl:452-456).  One thread per node.
template <int NORB>
__global__ void __launch_bounds__(128)
nest_eval_small_kernel(const double2* __restrict__ L1, long l1_stride, const long* __restrict__ slot1,
                       const double* __restrict__ x1, long npts, int M1, int lo, double period, int fkind, double2 z,
                       const double2* __restrict__ sigma, double2* __restrict__ y, int* __restrict__ errflag) {
    constexpr int NN = NORB * NORB;
    const long i = (long)blockIdx.x * 128 + threadIdx.x;
    if (i >= npts) return;
    const double2* c = L1 + (slot1 ? slot1[i] * l1_stride : 0);
    const double xi = x1[i];
    double2 h[NN];
#pragma unroll
    for (int e = 0; e < NN; e++) h[e] = make_double2(0.0, 0.0);
    for (int m = 0; m < M1; m++) {
        double2 p = cis2pi(xi * (double)(m + lo) / period);
#pragma unroll
        for (int e = 0; e < NN; e++) h[e] = cfma(h[e], c[m * NN + e], p);
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
    y[i] = t;
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
__global__ void __launch_bounds__(256)
symptr_rule_kernel(int N, int nsyms, const int* __restrict__ syms, int* __restrict__ wsym) {
    extern __shared__ int sy[];
    for (int t = threadIdx.x; t < 9 * nsyms; t += 256) sy[t] = syms[t];
    __syncthreads();
    const long tot = (long)N * N * N;
    const long idx = (long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= tot) return;
    const int i1 = (int)(idx % N), i2 = (int)((idx / N) % N), i3 = (int)(idx / ((long)N * N));
    // images (deduplicated by counting only the first occurrence among the symmetry list)
    int cnt = 0;
    bool is_min = true, has_self = false;
    for (int s = 0; s < nsyms && is_min; s++) {
        const int* S = sy + 9 * s;
        long j1 = (long)S[0] * i1 + (long)S[1] * i2 + (long)S[2] * i3;
        long j2 = (long)S[3] * i1 + (long)S[4] * i2 + (long)S[5] * i3;
        long j3 = (long)S[6] * i1 + (long)S[7] * i2 + (long)S[8] * i3;
        j1 = ((j1 % N) + N) % N; j2 = ((j2 % N) + N) % N; j3 = ((j3 % N) + N) % N;
        const long jdx = (j3 * N + j2) * N + j1;
        if (jdx < idx) { is_min = false; break; }
        if (jdx == idx) has_self = true;
        bool dup = false;
        for (int r = 0; r < s; r++) {
            const int* R = sy + 9 * r;
            long q1 = (long)R[0] * i1 + (long)R[1] * i2 + (long)R[2] * i3;
            long q2 = (long)R[3] * i1 + (long)R[4] * i2 + (long)R[5] * i3;
            long q3 = (long)R[6] * i1 + (long)R[7] * i2 + (long)R[8] * i3;
            q1 = ((q1 % N) + N) % N; q2 = ((q2 % N) + N) % N; q3 = ((q3 % N) + N) % N;
            if ((q3 * N + q2) * N + q1 == jdx) { dup = true; break; }
        }
        if (!dup) cnt++;
    }
    if (is_min && !has_self) cnt++;   // identity absent from the list
    wsym[idx] = is_min ? cnt : 0;
}

}  // namespace abz
