// Shared device helpers for libautobz_cuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace abz {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// a - b*c
__device__ __forceinline__ double2 cfnma(double2 a, double2 b, double2 c) {
    double x = fma(-b.x, c.x, a.x);
    x = fma(b.y, c.y, x);
    double y = fma(-b.x, c.y, a.y);
    y = fma(-b.y, c.x, y);
    return make_double2(x, y);
}
// a + b*c
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {
    double x = fma(b.x, c.x, a.x);
    x = fma(-b.y, c.y, x);
    double y = fma(b.x, c.y, a.y);
    y = fma(b.y, c.x, y);
    return make_double2(x, y);
}
// 1/a (Smith's algorithm: no spurious overflow)
__device__ __forceinline__ double2 crecip(double2 a) {
    if (fabs(a.x) >= fabs(a.y)) {
        double r = a.y / a.x;
        double d = 1.0 / (a.x + a.y * r);
        return make_double2(d, -r * d);
    } else {
        double r = a.x / a.y;
        double d = 1.0 / (a.x * r + a.y);
        return make_double2(r * d, -d);
    }
}
__device__ __forceinline__ double2 cdiv(double2 a, double2 b) { return cmul(a, crecip(b)); }

// exp(2 pi i frac) with the argument reduced to [-1/2, 1/2]
__device__ __forceinline__ double2 cis2pi(double frac) {
    frac -= rint(frac);
    double s, c;
    sincospi(2.0 * frac, &s, &c);
    return make_double2(c, s);
}

// FP64 tensor-core MMA (SASS DMMA.8x8x4): D[8x8] += A[8x4] * B[4x8]
// lane = 4*g + q:  a = A[g][q], b = B[q][g], (d0,d1) = D[g][2q], D[g][2q+1]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

}  // namespace abz
