// Microbenchmark: FP64 pipe ceilings on sm_100a (DFMA, DMMA shapes, SHFL), used to pick the
// resolvent kernel formulation and as roofline denominators next to cuBLAS DGEMM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C 2 regs
template <int ILP>
__global__ void k_dmma884(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x; c1[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k4: A 2 regs, B 1, C 4
template <int ILP>
__global__ void k_dmma1684(double* out, int iters, double a, double b) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k8: A 4 regs, B 2, C 4
template <int ILP>
__global__ void k_dmma1688(double* out, int iters, double a, double b) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k16: A 8 regs, B 4, C 4
template <int ILP>
__global__ void k_dmma16816(double* out, int iters, double a, double b) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(a), "d"(b), "d"(b), "d"(a), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_shfl(double* out, int iters) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) v[i] = __shfl_sync(0xffffffffu, v[i], (threadIdx.x + 1) & 31);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix: DMMA m8n8k4 with interleaved 64-bit shuffles, to see whether SHFL issue overlaps the FP64 pipe
template <int NSH>
__global__ void k_mix(double* out, int iters, double a, double b) {
    double c0[4], c1[4], v[NSH > 0 ? NSH : 1];
#pragma unroll
    for (int i = 0; i < 4; i++) { c0[i] = threadIdx.x; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < NSH; i++) v[i] = threadIdx.x * 3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
#pragma unroll
        for (int i = 0; i < NSH; i++) v[i] = __shfl_sync(0xffffffffu, v[i], (threadIdx.x + 5) & 31);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NSH; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
    const int iters = 20000;
    for (int wps = 4; wps <= 32; wps *= 2) {   // warps per SM
        int threads = 256, blocks = sms * wps * 32 / threads;
        printf("--- warps/SM %d ---\n", wps);
        {   float ms = timeit([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * 8 * iters * (double)blocks * threads;
            printf("DFMA ilp8        %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9); }
        {   float ms = timeit([&] { k_dmma884<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)blocks * threads / 32;
            printf("DMMA m8n8k4 ilp8 %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9); }
        {   float ms = timeit([&] { k_dmma1684<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * 16 * 8 * 4 * 4 * iters * (double)blocks * threads / 32;
            printf("DMMA m16n8k4 ilp4 %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9); }
        {   float ms = timeit([&] { k_dmma1688<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * 16 * 8 * 8 * 4 * iters * (double)blocks * threads / 32;
            printf("DMMA m16n8k8 ilp4 %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9); }
        {   float ms = timeit([&] { k_dmma16816<4><<<blocks, threads>>>(out, iters / 2, 1.0000001, 1e-9); });
            double fl = 2.0 * 16 * 8 * 16 * 4 * (iters / 2) * (double)blocks * threads / 32;
            printf("DMMA m16n8k16 ilp4 %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9); }
        {   float ms = timeit([&] { k_dmma884<1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * 8 * 8 * 4 * 1 * iters * (double)blocks * threads / 32;
            printf("DMMA m8n8k4 ilp1 (latency chain) %8.3f ms  %7.2f TFLOP/s  => %.1f ns per dependent DMMA\n", ms, fl / ms * 1e-9, ms * 1e6 / iters); }
        {   float ms = timeit([&] { k_shfl<8><<<blocks, threads>>>(out, iters); });
            double n = 8.0 * iters * (double)blocks * threads / 32;
            printf("SHFL.64 ilp8     %8.3f ms  %7.2f G warp-shfl64/s (%.2f per SM per ns)\n", ms, n / ms * 1e-6, n / ms * 1e-6 / sms); }
        {   float m0 = timeit([&] { k_mix<0><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            float m2 = timeit([&] { k_mix<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            float m4 = timeit([&] { k_mix<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            float m8 = timeit([&] { k_mix<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            printf("mix 4 DMMA + {0,2,4,8} shfl64: %8.3f %8.3f %8.3f %8.3f ms\n", m0, m2, m4, m8); }
    }
    CK(cudaFree(out));
    return 0;
}
