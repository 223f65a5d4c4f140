// K4: batched Hermitian eigenvalues by parallel cyclic two-sided Jacobi, one CTA per k-point,
// matrix resident in shared memory (eigen(Hermitian(H(k))) of src/dos_ggr.jl:19,34).
#pragma once
#include "abz_common.cuh"

namespace abz {

__device__ __forceinline__ double eig_fermi(double x) {
    return x > 0 ? exp(-x) / (1.0 + exp(-x)) : 1.0 / (1.0 + exp(x));
}
__device__ __forceinline__ double eig_kernel_value(double e, int kind, double p0, double p1) {
    switch (kind) {
        case 0: return e;
        case 1: return e * eig_fermi((e - p0) / p1);
        case 2: return eig_fermi((e - p0) / p1);
        default: { double u = (e - p0) / p1; return exp(-u * u) / (p1 * 1.7724538509055160273); }
    }
}

// round-robin tournament ("circle method"): np even players, step s in [0, np-1), pair t in [0, np/2)
__device__ __forceinline__ void rr_pair(int np, int s, int t, int& p, int& q) {
    int m = np - 1;
    if (t == 0) { p = m; q = s; }
    else { p = (s + t) % m; q = (s - t + m) % m; }
    if (p > q) { int tmp = p; p = q; q = tmp; }
}

// mode 0: partial[cta] = sum over this CTA's nodes of wnode * sum_n g(e_n)
// mode 1: evals[k*n + i] ascending
// shared: A[n*lda] double2 | rc[np/2] double | rs[np/2] double2 | d[n] double | red[blockDim/32 * 2] double | flag
__global__ void eig_jacobi_kernel(const double2* __restrict__ H, const double* __restrict__ wnode, long nk, int n, int mode,
                                  int kind, double p0, double p1, double* __restrict__ evals, double* __restrict__ partial,
                                  int* __restrict__ errflag) {
    extern __shared__ double2 eg_smem[];
    const int lda = n + 1;
    const int np = (n + 1) & ~1;
    const int npair = np / 2;
    double2* A = eg_smem;
    double2* rs = A + (long)n * lda;
    double* rc = reinterpret_cast<double*>(rs + npair);
    double* d = rc + npair;
    double* red = d + n;
    int* pq = reinterpret_cast<int*>(red + 2 * (blockDim.x / 32) + 2);   // [2*npair]
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    double my_acc = 0.0;
    for (long k = blockIdx.x; k < nk; k += gridDim.x) {
        __syncthreads();
        const double2* Hk = H + k * (long)n * n;
        for (int e = tid; e < n * n; e += nthr) {
            int i = e % n, j = e / n;
            double2 a = Hk[i + (long)j * n], b = Hk[j + (long)i * n];
            A[i + j * lda] = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
        }
        __syncthreads();
        for (int sweep = 0; sweep < 40; sweep++) {
            // convergence test: off^2 <= 1e-28 tot^2
            double off = 0.0, tot = 0.0;
            for (int e = tid; e < n * n; e += nthr) {
                int i = e % n, j = e / n;
                double2 a = A[i + j * lda];
                double v = a.x * a.x + a.y * a.y;
                tot += v;
                if (i != j) off += v;
            }
            off = warp_sum(off); tot = warp_sum(tot);
            if (lane == 0) { red[2 * warp] = off; red[2 * warp + 1] = tot; }
            __syncthreads();
            if (tid == 0) {
                double o = 0.0, t = 0.0;
                for (int w = 0; w < nwarp; w++) { o += red[2 * w]; t += red[2 * w + 1]; }
                red[2 * nwarp] = o; red[2 * nwarp + 1] = t;
            }
            __syncthreads();
            const double o = red[2 * nwarp], t = red[2 * nwarp + 1];
            __syncthreads();
            if (!(t == t) || !isfinite(t)) { if (tid == 0) *errflag = 1; break; }
            if (o <= 1e-28 * t) break;
            for (int s = 0; s < np - 1; s++) {
                if (tid < npair) {
                    int p, q;
                    rr_pair(np, s, tid, p, q);
                    double c = 1.0; double2 sp = make_double2(0.0, 0.0);
                    if (q < n) {
                        double2 apq = A[p + q * lda];
                        double g = hypot(apq.x, apq.y);
                        if (g > 0.0) {
                            double app = A[p + p * lda].x, aqq = A[q + q * lda].x;
                            double tau = (aqq - app) / (2.0 * g);
                            double tt = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                            c = 1.0 / sqrt(1.0 + tt * tt);
                            double sn = tt * c;
                            sp = make_double2(sn * apq.x / g, sn * apq.y / g);   // s * e^{i phi}
                        }
                    } else { q = -1; }
                    pq[2 * tid] = p; pq[2 * tid + 1] = q;
                    rc[tid] = c; rs[tid] = sp;
                }
                __syncthreads();
                // columns: A <- A J
                for (int item = tid; item < npair * n; item += nthr) {
                    int t2 = item / n, i = item % n;
                    int p = pq[2 * t2], q = pq[2 * t2 + 1];
                    if (q < 0) continue;
                    double c = rc[t2]; double2 sp = rs[t2];
                    double2 akp = A[i + p * lda], akq = A[i + q * lda];
                    // new col p = c akp - conj(sp) akq ; new col q = sp akp + c akq
                    double2 np_ = make_double2(c * akp.x - (sp.x * akq.x + sp.y * akq.y), c * akp.y - (sp.x * akq.y - sp.y * akq.x));
                    double2 nq_ = make_double2(sp.x * akp.x - sp.y * akp.y + c * akq.x, sp.x * akp.y + sp.y * akp.x + c * akq.y);
                    A[i + p * lda] = np_; A[i + q * lda] = nq_;
                }
                __syncthreads();
                // rows: A <- J^H A
                for (int item = tid; item < npair * n; item += nthr) {
                    int t2 = item / n, j = item % n;
                    int p = pq[2 * t2], q = pq[2 * t2 + 1];
                    if (q < 0) continue;
                    double c = rc[t2]; double2 sp = rs[t2];
                    double2 apk = A[p + j * lda], aqk = A[q + j * lda];
                    // new row p = c apk - sp aqk ; new row q = conj(sp) apk + c aqk
                    double2 np_ = make_double2(c * apk.x - (sp.x * aqk.x - sp.y * aqk.y), c * apk.y - (sp.x * aqk.y + sp.y * aqk.x));
                    double2 nq_ = make_double2(sp.x * apk.x + sp.y * apk.y + c * aqk.x, sp.x * apk.y - sp.y * apk.x + c * aqk.y);
                    A[p + j * lda] = np_; A[q + j * lda] = nq_;
                }
                __syncthreads();
            }
        }
        for (int i = tid; i < n; i += nthr) d[i] = A[i + i * lda].x;
        __syncthreads();
        if (mode == 1) {
            for (int i = tid; i < n; i += nthr) {
                double di = d[i];
                int rank = 0;
                for (int j = 0; j < n; j++) { double dj = d[j]; rank += (dj < di) || (dj == di && j < i); }
                evals[k * n + rank] = di;
            }
        } else {
            double v = 0.0;
            for (int i = tid; i < n; i += nthr) v += eig_kernel_value(d[i], kind, p0, p1);
            v = warp_sum(v);
            if (lane == 0) red[warp] = v;
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int w = 0; w < nwarp; w++) s += red[w];
                my_acc += (wnode ? wnode[k] : 1.0) * s;
            }
        }
    }
    if (mode == 0 && tid == 0) partial[blockIdx.x] = my_acc;
}

// deterministic reduction of real partials: acc[0] += scale * sum_c partial[c]
__global__ void __launch_bounds__(256) reduce_real_kernel(const double* __restrict__ partial, long n, double scale, double* __restrict__ acc) {
    __shared__ double sx[256];
    double x = 0.0;
    for (long c = threadIdx.x; c < n; c += 256) x += partial[c];
    sx[threadIdx.x] = x;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sx[threadIdx.x] += sx[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) acc[0] += scale * sx[0];
}

}  // namespace abz
