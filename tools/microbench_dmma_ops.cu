// DMMA issue cost with distinct A/B operand registers (as in the real kernel) vs identical operands
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
#define DMMA(c0, c1, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b))

template <int MODE>
__global__ void k(double* out, const double* in, int iters) {
    double a[8], b[8], c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * (i + 8)]; c0[i] = 0; c1[i] = 0; }
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {            // same A, B for all
#pragma unroll
            for (int i = 0; i < 8; i++) DMMA(c0[i], c1[i], a[0], b[0]);
        } else if (MODE == 1) {     // distinct A and B per instruction
#pragma unroll
            for (int i = 0; i < 8; i++) DMMA(c0[i], c1[i], a[i], b[i]);
        } else if (MODE == 2) {     // complex bmm pattern: 2 accumulators, 4 A values (2 negated), 4 B values
            DMMA(c0[0], c1[0], a[0], b[0]); DMMA(c0[1], c1[1], a[0], b[2]);
            DMMA(c0[0], c1[0], a[1], b[1]); DMMA(c0[1], c1[1], a[1], b[3]);
            DMMA(c0[0], c1[0], -a[2], b[2]); DMMA(c0[1], c1[1], a[2], b[0]);
            DMMA(c0[0], c1[0], -a[3], b[3]); DMMA(c0[1], c1[1], a[3], b[1]);
        } else {                    // 4 independent accumulators pairs (two bmm interleaved)
            DMMA(c0[0], c1[0], a[0], b[0]); DMMA(c0[1], c1[1], a[0], b[2]); DMMA(c0[2], c1[2], a[4], b[0]); DMMA(c0[3], c1[3], a[4], b[2]);
            DMMA(c0[0], c1[0], a[1], b[1]); DMMA(c0[1], c1[1], a[1], b[3]); DMMA(c0[2], c1[2], a[5], b[1]); DMMA(c0[3], c1[3], a[5], b[3]);
            DMMA(c0[0], c1[0], -a[2], b[2]); DMMA(c0[1], c1[1], a[2], b[0]); DMMA(c0[2], c1[2], -a[6], b[2]); DMMA(c0[3], c1[3], a[6], b[0]);
            DMMA(c0[0], c1[0], -a[3], b[3]); DMMA(c0[1], c1[1], a[3], b[1]); DMMA(c0[2], c1[2], -a[7], b[3]); DMMA(c0[3], c1[3], a[7], b[1]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(double* out, double* in, int sms, int wps, int ndmma) {
    const int iters = 4000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<MODE><<<sms, 32 * wps>>>(out, in, iters); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { CK(cudaEventRecord(e0)); k<MODE><<<sms, 32 * wps>>>(out, in, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    double cyc = best * 1e-3 * 1.965e9 / iters / (wps / 4.0) / ndmma;
    printf("mode %d warps/SM %2d: %7.3f ms -> %5.2f cycles per DMMA per SMSP\n", MODE, wps, best, cyc);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double *out, *in; CK(cudaMalloc(&out, sizeof(double) * sms * 1024)); CK(cudaMalloc(&in, sizeof(double) * 1024)); CK(cudaMemset(in, 0, sizeof(double) * 1024));
    for (int wps : {4, 8, 12}) { run<0>(out, in, sms, wps, 8); run<1>(out, in, sms, wps, 8); run<2>(out, in, sms, wps, 8); run<3>(out, in, sms, wps, 16); }
    return 0;
}
