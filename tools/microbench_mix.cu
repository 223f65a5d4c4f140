// Do DMMA and scalar FP64 (DFMA) share an issue pipe on sm_100a, and is there a switch penalty?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NMMA, int NFMA, bool INTERLEAVE>
__global__ void k_mix(double* out, int iters, double a, double b) {
    double c0[8], c1[8], f[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { c0[i] = threadIdx.x; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < 16; i++) f[i] = threadIdx.x * 0.5 + i;
    for (int it = 0; it < iters; it++) {
        if (INTERLEAVE) {
#pragma unroll
            for (int i = 0; i < (NMMA > NFMA ? NMMA : NFMA); i++) {
                if (i < NMMA)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i % 8]), "+d"(c1[i % 8]) : "d"(a), "d"(b));
                if (i < NFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i % 16]) : "d"(a), "d"(b));
            }
        } else {
#pragma unroll
            for (int i = 0; i < NMMA; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i % 8]), "+d"(c1[i % 8]) : "d"(a), "d"(b));
#pragma unroll
            for (int i = 0; i < NFMA; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i % 16]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < 16; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float timeit(F launch) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    return best;
}

template <int NMMA, int NFMA, bool IL> void run(double* out, int sms, int wps) {
    const int iters = 4000;
    int threads = 32 * (wps / 1), blocks = sms;   // one CTA per SM with wps warps
    float ms = timeit([&] { k_mix<NMMA, NFMA, IL><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double cyc = ms * 1e-3 * 1.965e9 / iters;           // cycles per iteration (all warps of an SMSP run concurrently)
    double per_smsp_warps = wps / 4.0;
    printf("warps/SM %2d  %2d DMMA + %2d DFMA %s: %8.3f ms  %7.1f cyc/iter  => %6.1f cyc per warp-iter on the pipe (model shared: %d)\n", wps, NMMA, NFMA,
           IL ? "interleaved" : "blocked    ", ms, cyc, cyc / per_smsp_warps, NMMA * 16 + NFMA * 2);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 1024));
    for (int wps : {4, 8, 12}) {
        if (wps == 4) { run<8, 0, false>(out, sms, 4); run<0, 32, false>(out, sms, 4); run<8, 16, false>(out, sms, 4); run<8, 16, true>(out, sms, 4); run<8, 32, false>(out, sms, 4); run<8, 32, true>(out, sms, 4); run<8, 64, true>(out, sms, 4); }
        if (wps == 8) { run<8, 0, false>(out, sms, 8); run<0, 32, false>(out, sms, 8); run<8, 16, false>(out, sms, 8); run<8, 16, true>(out, sms, 8); run<8, 32, false>(out, sms, 8); run<8, 32, true>(out, sms, 8); run<8, 64, true>(out, sms, 8); }
        if (wps == 12) { run<8, 0, false>(out, sms, 12); run<0, 32, false>(out, sms, 12); run<8, 16, false>(out, sms, 12); run<8, 16, true>(out, sms, 12); run<8, 32, false>(out, sms, 12); run<8, 32, true>(out, sms, 12); run<8, 64, true>(out, sms, 12); }
    }
    return 0;
}
