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

// branch-free reciprocal: MUFU.RCP64H seed (~2^-20) + two Newton steps (keeps the whole elimination in one
// basic block so that ptxas can interleave the DMMA chains with the latency-bound pivot steps)
__device__ __forceinline__ double fast_rcp(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    const double e = fma(-d, x, 1.0);      // 1 - d x0
    const double e2 = e * e;               // = 1 - d x1 up to rounding (in parallel with x1)
    x = fma(x, e, x);                      // x1
    x = fma(x, e2, x);                     // x2: relative error ~ e^4 (seed 2^-20 -> 2^-80)
    return x;
}
// 1/sqrt(x) for x well inside the normal range: MUFU.RSQ64H seed (~2^-22) + two Newton steps y += (y/2)(1 - x y^2) (error ~1e-26 before
// the final rounding); sqrt(x) = x * fast_rsqrt(x) to ~1.5 ulp.  A third of the dependent depth of DSQRT followed by a division.
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double xy = x * y;
    double e = fma(-xy, y, 1.0);
    y = fma(0.5 * y, e, y);
    xy = x * y;
    e = fma(-xy, y, 1.0);
    y = fma(0.5 * y, e, y);
    return y;
}
// 1/a = conj(a)/|a|^2 with the branch-free reciprocal: 1 MUFU + 8 FP64 instructions instead of Smith's two IEEE divisions.
// For the determinants of z - H (|a|^2 far from the overflow / underflow thresholds); ~2 ulp.
__device__ __forceinline__ double2 crecip_fast(double2 a) {
    const double s = fast_rcp(fma(a.x, a.x, a.y * a.y));
    return make_double2(a.x * s, -a.y * s);
}
__device__ __forceinline__ double2 cdiv_fast(double2 a, double2 b) { return cmul(a, crecip_fast(b)); }

// Plane p of one rank's share of the k3 planes.  k3_stride > 0: k3_lo + p * k3_stride (round-robin among k3_stride ranks).
// k3_stride < 0: "serpentine" dealing among W = -k3_stride ranks, rank k3_lo < W taking planes r, 2W-1-r, 2W+r, 4W-1-r, ...:
// the planes of a symmetry-reduced grid shrink monotonically with k3 (the irreducible wedge), so round-robin hands rank 0 up to
// ~22 % more nodes than the average at W = 8 (cubic group, npt = 96), the serpentine ~4 %.
__host__ __device__ __forceinline__ int share_plane(int p, int k3_lo, int k3_stride) {
    if (k3_stride > 0) return k3_lo + p * k3_stride;
    const int w2 = -2 * k3_stride;
    return (p >> 1) * w2 + ((p & 1) ? w2 - 1 - k3_lo : k3_lo);
}

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

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on a shared-memory mbarrier -----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t addr = smem_u32(bar);
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

}  // namespace abz
