// K3-team: tr[(z - H(k) - Sigma_w)^-1] by the same DMMA block LU + trace-from-the-factors as abz_resolvent_mma.cuh, with ONE
// matrix shared by a TEAM of NW warps (norb <= 64: NB = 8 block rows/columns of 8 x 8 complex blocks, NW = 4 warps).
//
// Distribution: block column j belongs to warp j mod NW, which keeps its NB x ceil(NB/NW) blocks in registers in the DMMA
// accumulator layout (lane 4g+q: row g, columns 2q, 2q+1, re and im: 8 registers per block).  Left operands of every block
// product travel through shared memory in that same layout (written once by their producer, read back lane-for-lane with two
// conflict-free 128-bit loads: a block in C layout IS a valid A operand, see abz_resolvent_mma.cuh), right operands are
// B fragments built by the consumer from its own registers (4 shuffles) or published as fragments:
//   panel s (owner = warp s mod NW):  D_s = inv(S_s) (in-register Gauss-Jordan), L_is = A_is D_s (i > s)      -> smem (write once)
//   every warp, own columns j > s:    A_ij -= L_is U_sj (i > s),  X_sj = D_s U_sj                              -> X_sj to smem
//   look-ahead: the owner of column s+1 updates it first and factors panel s+1 before touching its other columns; the panel
//   hand-over is a producer/consumer named barrier (bar.arrive by the owner, bar.sync by the others), one per step.
//   V = U~^-1 and M = L~^-1 column by column (each column entirely inside its owner: right-looking substitutions, one live
//   fragment at a time), fragments of M to smem, then tr A^-1 = sum_s tr D_s + sum_{i<j} tr(V_ij M_ji) by the owners of V.
// Pivoting: none between blocks (A = z - H - Sigma with Im z > 0 and a causal Sigma has nonsingular leading minors); the
// smallest pivot is monitored against max|a_ij| exactly as in the one-warp kernel and the same flag asks the host to rerun the
// call with the pivoted Gauss-Jordan teams (abz_resolvent_gjreg.cuh).
#pragma once
#include "abz_common.cuh"
#include "abz_resolvent_mma.cuh"

#ifndef ABZ_TEAM_PIVOT_THR
#define ABZ_TEAM_PIVOT_THR 2e-3
#endif

namespace abz {

struct Blk8 { double r0, r1, i0, i1; int minhi; };

// inv8 behind a call: NB call sites per matrix instead of NB inlined copies of the 8-step elimination
__device__ __noinline__ Blk8 inv8_call(Blk8 x, int lane) {
    inv8(x.r0, x.r1, x.i0, x.i1, lane, x.minhi);
    return x;
}

// shared-memory image of a block (C layout or B fragment): two 128-bit words per lane, lane-major per word => conflict-free
typedef double2 TeamBlk[2][32];

__device__ __forceinline__ void tb_store(TeamBlk& b, int lane, double r0, double r1, double i0, double i1) {
    b[0][lane] = make_double2(r0, r1);
    b[1][lane] = make_double2(i0, i1);
}
__device__ __forceinline__ void tb_load(const TeamBlk& b, int lane, double& r0, double& r1, double& i0, double& i1) {
    const double2 a = b[0][lane], c = b[1][lane];
    r0 = a.x; r1 = a.y; i0 = c.x; i1 = c.y;
}
__device__ __forceinline__ BFrag tb_load_frag(const TeamBlk& b, int lane) {
    const double2 a = b[0][lane], c = b[1][lane];
    BFrag f; f.r[0] = a.x; f.r[1] = a.y; f.i[0] = c.x; f.i[1] = c.y;
    return f;
}
__device__ __forceinline__ void tb_store_frag(TeamBlk& b, int lane, const BFrag& f) {
    b[0][lane] = make_double2(f.r[0], f.r[1]);
    b[1][lane] = make_double2(f.i[0], f.i[1]);
}

// producer / consumer named barriers of the team (ids 1, 2 alternate with the step; id 0 is __syncthreads)
__device__ __forceinline__ void team_bar_arrive(int id, int nthreads) {
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void team_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int NB>
struct TeamSmem {
    TeamBlk L[NB * (NB + 1) / 2];     // L_is (i > s) and D_s (i == s) in C layout: index i (i + 1) / 2 + s
    TeamBlk X[NB * (NB - 1) / 2 + 1]; // X_sj (s < j) in C layout: index j (j - 1) / 2 + s
    TeamBlk M[NB * (NB - 1) / 2 + 1]; // B fragments of M_ij (i > j): index i (i - 1) / 2 + j
    TeamBlk Df[NB];                   // B fragments of D_s
    double red[8][2];
    int redi[8][2];
};

// CTA = one team.  CTA c handles k-points [c*kper, (c+1)*kper) x all nw frequencies, one (k, w) matrix after the other.
// mode 0: outp[c*nw + w] = sum_k wnode_k tr ; mode 1: outp[k*nw + w] = tr.   dynamic shared: TeamSmem<NB> | acc[nw]
template <int NB, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
resolvent_mma_team_kernel(const double2* __restrict__ H, const double* __restrict__ wnode, long nk, int n, int nw,
                          const double2* __restrict__ z, const double2* __restrict__ sigma, int kper, int mode,
                          double2* __restrict__ outp, int* __restrict__ errflag, double piv_thr2) {
    constexpr int NC = (NB + NW - 1) / NW;        // block columns per warp
    constexpr int NT = NW * 32;
    extern __shared__ __align__(16) unsigned char team_smem_raw[];
    TeamSmem<NB>& sm = *reinterpret_cast<TeamSmem<NB>*>(team_smem_raw);
    double2* acc = reinterpret_cast<double2*>(team_smem_raw + sizeof(TeamSmem<NB>));
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const bool par = g & 1;
    const int src0 = 4 * (2 * q + (par ? 1 : 0)) + (g >> 1);
    const int src1 = 4 * (2 * q + (par ? 0 : 1)) + (g >> 1);
    if (mode == 0)
        for (int w = threadIdx.x; w < nw; w += NT) acc[w] = make_double2(0.0, 0.0);
    __syncthreads();
    const long k0 = (long)blockIdx.x * kper;
    const long k1 = k0 + kper < nk ? k0 + kper : nk;
    const long npairs = (k1 - k0) * nw;
    const int npad = 8 * NB - n;
    for (long it = 0; it < npairs; it++) {
        const long k = k0 + it / nw;
        const int w = (int)(it % nw);
        const double2* Hk = H + k * (long)n * n;
        const double2* sg = sigma ? sigma + (long)w * n * n : nullptr;
        const double2 zz = z[w];
        // ---- this warp's block columns of B = H + Sigma - z (= -A; tr A^-1 = -tr B^-1), identity padding
        double R0[NB][NC], R1[NB][NC], I0[NB][NC], I1[NB][NC];
        int amaxhi = 0;
#pragma unroll
        for (int jl = 0; jl < NC; jl++) {
            const int bj = jl * NW + wid;
#pragma unroll
            for (int bi = 0; bi < NB; bi++) {
                const int row = 8 * bi + g, c0 = 8 * bj + 2 * q, c1 = c0 + 1;
                double2 a0 = make_double2(0.0, 0.0), a1 = a0;
                if (bj < NB) {
                    if (row < n && c0 < n) {
                        a0 = Hk[row + (long)c0 * n];
                        if (sg) { const double2 s0 = sg[row + (long)c0 * n]; a0.x += s0.x; a0.y += s0.y; }
                    }
                    if (row < n && c1 < n) {
                        a1 = Hk[row + (long)c1 * n];
                        if (sg) { const double2 s1 = sg[row + (long)c1 * n]; a1.x += s1.x; a1.y += s1.y; }
                    }
                    if (bi == bj) {
                        if (row == c0) { if (row < n) { a0.x -= zz.x; a0.y -= zz.y; } else { a0.x = -1.0; a0.y = 0.0; } }
                        if (row == c1) { if (row < n) { a1.x -= zz.x; a1.y -= zz.y; } else { a1.x = -1.0; a1.y = 0.0; } }
                    }
                }
                R0[bi][jl] = a0.x; I0[bi][jl] = a0.y; R1[bi][jl] = a1.x; I1[bi][jl] = a1.y;
                amaxhi = max(amaxhi, max(max(__double2hiint(a0.x) & 0x7fffffff, __double2hiint(a0.y) & 0x7fffffff),
                                         max(__double2hiint(a1.x) & 0x7fffffff, __double2hiint(a1.y) & 0x7fffffff)));
            }
        }
        int minhi = 0x7ff00000;
        double tr = 0.0, ti = 0.0;

        // panel s: D_s, its fragment, L_is (i > s) - executed by the owner of column s only (s, hence s / NW, is static)
#define ABZ_TEAM_FACTOR(S)                                                                                              \
        {                                                                                                               \
            constexpr int sl = (S) / NW;                                                                                \
            Blk8 d{R0[S][sl], R1[S][sl], I0[S][sl], I1[S][sl], minhi};                                                  \
            d = inv8_call(d, lane);                                                                                     \
            R0[S][sl] = d.r0; R1[S][sl] = d.r1; I0[S][sl] = d.i0; I1[S][sl] = d.i1; minhi = d.minhi;                    \
            if (2 * q == g) { tr += d.r0; ti += d.i0; }                                                                 \
            if (2 * q + 1 == g) { tr += d.r1; ti += d.i1; }                                                             \
            const BFrag bD = to_bfrag(d.r0, d.r1, d.i0, d.i1, src0, src1, par);                                         \
            tb_store(sm.L[(S) * ((S) + 1) / 2 + (S)], lane, d.r0, d.r1, d.i0, d.i1);                                    \
            tb_store_frag(sm.Df[S], lane, bD);                                                                          \
            _Pragma("unroll")                                                                                           \
            for (int i = (S) + 1; i < NB; i++) {                                                                        \
                double cr0 = 0, cr1 = 0, ci0 = 0, ci1 = 0;                                                              \
                bmm<false>(cr0, cr1, ci0, ci1, R0[i][sl], R1[i][sl], I0[i][sl], I1[i][sl], bD);                         \
                R0[i][sl] = cr0; R1[i][sl] = cr1; I0[i][sl] = ci0; I1[i][sl] = ci1;                                     \
                tb_store(sm.L[i * (i + 1) / 2 + (S)], lane, cr0, cr1, ci0, ci1);                                        \
            }                                                                                                           \
        }

        if (wid == 0) {
            ABZ_TEAM_FACTOR(0)
            if (NB > 1) team_bar_arrive(1, NT);
        }
        // ---- block LU, right-looking, with look-ahead on the next panel
#pragma unroll
        for (int s = 0; s < NB - 1; s++) {
            if (wid != s % NW) team_bar_sync(1 + (s & 1), NT);
            else __threadfence_block();
#pragma unroll
            for (int jl = 0; jl < NC; jl++) {
                const int j = jl * NW + wid;
                if (j > s && j < NB) {
                    const BFrag bU = to_bfrag(R0[s][jl], R1[s][jl], I0[s][jl], I1[s][jl], src0, src1, par);
#pragma unroll
                    for (int i = s + 1; i < NB; i++) {                 // A_ij -= L_is U_sj
                        double ar0, ar1, ai0, ai1;
                        tb_load(sm.L[i * (i + 1) / 2 + s], lane, ar0, ar1, ai0, ai1);
                        bmm<true>(R0[i][jl], R1[i][jl], I0[i][jl], I1[i][jl], ar0, ar1, ai0, ai1, bU);
                    }
                    {                                                  // X_sj = D_s U_sj (replaces U_sj), published for the V phase
                        double ar0, ar1, ai0, ai1;
                        tb_load(sm.L[s * (s + 1) / 2 + s], lane, ar0, ar1, ai0, ai1);
                        double cr0 = 0, cr1 = 0, ci0 = 0, ci1 = 0;
                        bmm<false>(cr0, cr1, ci0, ci1, ar0, ar1, ai0, ai1, bU);
                        R0[s][jl] = cr0; R1[s][jl] = cr1; I0[s][jl] = ci0; I1[s][jl] = ci1;
                        tb_store(sm.X[j * (j - 1) / 2 + s], lane, cr0, cr1, ci0, ci1);
                    }
                    // look-ahead: column s+1 is complete - its owner factors panel s+1 before its other columns
                    if (jl == (s + 1) / NW && wid == (s + 1) % NW) {
                        switch (s + 1) {       // s is a compile-time constant after unrolling; the switch keeps the macro argument literal
                            case 1: if (NB > 1) ABZ_TEAM_FACTOR(1 < NB ? 1 : 0) break;
                            case 2: if (NB > 2) ABZ_TEAM_FACTOR(2 < NB ? 2 : 0) break;
                            case 3: if (NB > 3) ABZ_TEAM_FACTOR(3 < NB ? 3 : 0) break;
                            case 4: if (NB > 4) ABZ_TEAM_FACTOR(4 < NB ? 4 : 0) break;
                            case 5: if (NB > 5) ABZ_TEAM_FACTOR(5 < NB ? 5 : 0) break;
                            case 6: if (NB > 6) ABZ_TEAM_FACTOR(6 < NB ? 6 : 0) break;
                            case 7: if (NB > 7) ABZ_TEAM_FACTOR(7 < NB ? 7 : 0) break;
                        }
                        if (s + 1 < NB - 1) team_bar_arrive(1 + ((s + 1) & 1), NT);
                    }
                }
            }
        }
#undef ABZ_TEAM_FACTOR
        __threadfence_block();
        __syncthreads();          // every X_sj, L_is, D_s is published
        // ---- V = U~^-1: column jv inside its owner.  V_ij = -X_ij D_j - sum_{i<t<j} X_it V_tj, right-looking: once V_tj is final its
        //      fragment updates the rows above it
#pragma unroll
        for (int jl = 0; jl < NC; jl++) {
            const int jv = jl * NW + wid;
            if (jv < NB && jv > 0) {
                const BFrag bD = tb_load_frag(sm.Df[jv], lane);
#pragma unroll
                for (int iv = 0; iv < NB - 1; iv++)
                    if (iv < jv) {
                        double cr0 = 0, cr1 = 0, ci0 = 0, ci1 = 0;
                        bmm<true>(cr0, cr1, ci0, ci1, R0[iv][jl], R1[iv][jl], I0[iv][jl], I1[iv][jl], bD);
                        R0[iv][jl] = cr0; R1[iv][jl] = cr1; I0[iv][jl] = ci0; I1[iv][jl] = ci1;
                    }
#pragma unroll
                for (int t = NB - 2; t >= 1; t--)
                    if (t < jv) {
                        const BFrag bV = to_bfrag(R0[t][jl], R1[t][jl], I0[t][jl], I1[t][jl], src0, src1, par);
#pragma unroll
                        for (int iv = 0; iv < t; iv++) {
                            double ar0, ar1, ai0, ai1;
                            tb_load(sm.X[t * (t - 1) / 2 + iv], lane, ar0, ar1, ai0, ai1);
                            bmm<true>(R0[iv][jl], R1[iv][jl], I0[iv][jl], I1[iv][jl], ar0, ar1, ai0, ai1, bV);
                        }
                    }
            }
        }
        // ---- M = L~^-1: column jm inside its owner (rows below the diagonal).  M_ij = -L_ij - sum_{j<t<i} L_it M_tj; fragments of the
        //      finished entries are published for the trace
#pragma unroll
        for (int jl = 0; jl < NC; jl++) {
            const int jm = jl * NW + wid;
            if (jm < NB - 1) {
#pragma unroll
                for (int im = 1; im < NB; im++)
                    if (im > jm) { R0[im][jl] = dneg(R0[im][jl]); R1[im][jl] = dneg(R1[im][jl]); I0[im][jl] = dneg(I0[im][jl]); I1[im][jl] = dneg(I1[im][jl]); }
#pragma unroll
                for (int t = 1; t < NB; t++)
                    if (t > jm) {
                        const BFrag bM = to_bfrag(R0[t][jl], R1[t][jl], I0[t][jl], I1[t][jl], src0, src1, par);
                        tb_store_frag(sm.M[t * (t - 1) / 2 + jm], lane, bM);
#pragma unroll
                        for (int im = t + 1; im < NB; im++) {
                            double ar0, ar1, ai0, ai1;
                            tb_load(sm.L[im * (im + 1) / 2 + t], lane, ar0, ar1, ai0, ai1);
                            bmm<true>(R0[im][jl], R1[im][jl], I0[im][jl], I1[im][jl], ar0, ar1, ai0, ai1, bM);
                        }
                    }
            }
        }
        __threadfence_block();
        __syncthreads();          // every fragment of M is published
        // ---- tr += sum_{i<j} tr(V_ij M_ji) by the owner of V's column j
#pragma unroll
        for (int jl = 0; jl < NC; jl++) {
            const int j = jl * NW + wid;
            if (j < NB) {
#pragma unroll
                for (int i = 0; i < NB - 1; i++)
                    if (i < j) {
                        const BFrag b = tb_load_frag(sm.M[j * (j - 1) / 2 + i], lane);
                        tr += R0[i][jl] * b.r[0] - I0[i][jl] * b.i[0] + R1[i][jl] * b.r[1] - I1[i][jl] * b.i[1];
                        ti += R0[i][jl] * b.i[0] + I0[i][jl] * b.r[0] + R1[i][jl] * b.i[1] + I1[i][jl] * b.r[1];
                    }
            }
        }
        tr = warp_sum(tr); ti = warp_sum(ti);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            amaxhi = max(amaxhi, __shfl_xor_sync(0xffffffffu, amaxhi, off));
            minhi = min(minhi, __shfl_xor_sync(0xffffffffu, minhi, off));
        }
        if (lane == 0) { sm.red[wid][0] = tr; sm.red[wid][1] = ti; sm.redi[wid][0] = amaxhi; sm.redi[wid][1] = minhi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double sx = 0.0, sy = 0.0;
            int am = 0, mn = 0x7ff00000;
#pragma unroll
            for (int wp = 0; wp < NW; wp++) { sx += sm.red[wp][0]; sy += sm.red[wp][1]; am = max(am, sm.redi[wp][0]); mn = min(mn, sm.redi[wp][1]); }
            const double2 t = make_double2(-sx - (double)npad, -sy);
            const double pmin2 = __hiloint2double(mn, 0), amx = __hiloint2double(am, 0);
            if (!(pmin2 > piv_thr2 * amx * amx) || !(isfinite(t.x) && isfinite(t.y))) atomicOr(errflag, 2);   // ask for the pivoted path
            if (mode == 0) {
                const double wt = wnode ? wnode[k] : 1.0;
                acc[w].x += wt * t.x; acc[w].y += wt * t.y;
            } else {
                outp[k * nw + w] = t;
            }
        }
        // the next matrix's panel 0 is written after this barrier-separated read of sm.red: one more sync keeps red / M / X / L
        // of this matrix from being overwritten while a slower warp still reads them
        __syncthreads();
    }
    if (mode == 0)
        for (int w = threadIdx.x; w < nw; w += NT) outp[(long)blockIdx.x * nw + w] = acc[w];
}

inline bool mma_team_supported(int n) { return n > 32 && n <= 64; }

// smallest |pivot| allowed relative to max|a_ij| before the call is rerun with the pivoted teams (squared); ABZ_MMA_TEAM_PIVOT_THR
// overrides (measurement hook, profiles/r02_team_resolvent_accuracy.log)
inline double mma_team_pivot_threshold2() {
    static double t = -1.0;
    if (t < 0.0) { const char* e = getenv("ABZ_MMA_TEAM_PIVOT_THR"); const double v = e ? atof(e) : ABZ_TEAM_PIVOT_THR; t = v * v; }
    return t;
}

template <int NB, int NW, int MINB>
inline cudaError_t mma_team_launch_one(const double2* H, const double* wnode, long nk, int n, int nw, const double2* z, const double2* sigma,
                                       int mode, double2* outp, int* errflag, long ncta, int kper, cudaStream_t stream) {
    const size_t smem = sizeof(TeamSmem<NB>) + (size_t)nw * sizeof(double2);
    resolvent_mma_team_kernel<NB, NW, MINB><<<(unsigned)ncta, NW * 32, smem, stream>>>(H, wnode, nk, n, nw, z, sigma, kper, mode, outp, errflag,
                                                                                      mma_team_pivot_threshold2());
    return cudaGetLastError();
}

inline size_t mma_team_smem(int NBv, int nw) {
    const size_t blk = sizeof(TeamBlk);
    return blk * (size_t)(NBv * (NBv + 1) / 2 + 2 * (NBv * (NBv - 1) / 2 + 1) + NBv) + 8 * 2 * sizeof(double) + 8 * 2 * sizeof(int) + (size_t)nw * sizeof(double2);
}

// grid: one resident wave of teams (2 per SM at NB = 8), each walking its k-chunk x all frequencies
inline int mma_team_plan(int n, long nk, int nw, long sm, long* ncta, int* kper) {
    if (!mma_team_supported(n)) return -1;
    if (mma_team_smem(8, nw) > 110 * 1024) return -1;
    const long target = sm * 2;
    long kp = (nk + target - 1) / target;
    if (kp < 1) kp = 1;
    *kper = (int)kp;
    *ncta = (nk + kp - 1) / kp;
    return 0;
}

inline cudaError_t mma_team_opt_in() {
    cudaError_t e = cudaSuccess;
    auto set = [&](const void* f) {
        cudaError_t r = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (r != cudaSuccess && e == cudaSuccess) e = r;
    };
    { auto k = resolvent_mma_team_kernel<5, 4, 2>; set((const void*)k); }
    { auto k = resolvent_mma_team_kernel<6, 4, 2>; set((const void*)k); }
    { auto k = resolvent_mma_team_kernel<7, 4, 2>; set((const void*)k); }
    { auto k = resolvent_mma_team_kernel<8, 4, 2>; set((const void*)k); }
    { auto k = resolvent_mma_team_kernel<4, 2, 8>; set((const void*)k); }
    return e;
}

inline cudaError_t mma_team_launch(const double2* H, const double* wnode, long nk, int n, int nw, const double2* z, const double2* sigma,
                                   int mode, double2* outp, int* errflag, long ncta, int kper, cudaStream_t stream) {
    switch ((n + 7) / 8) {
        case 5: return mma_team_launch_one<5, 4, 2>(H, wnode, nk, n, nw, z, sigma, mode, outp, errflag, ncta, kper, stream);
        case 6: return mma_team_launch_one<6, 4, 2>(H, wnode, nk, n, nw, z, sigma, mode, outp, errflag, ncta, kper, stream);
        case 7: return mma_team_launch_one<7, 4, 2>(H, wnode, nk, n, nw, z, sigma, mode, outp, errflag, ncta, kper, stream);
        case 8: return mma_team_launch_one<8, 4, 2>(H, wnode, nk, n, nw, z, sigma, mode, outp, errflag, ncta, kper, stream);
        default: return mma_team_launch_one<4, 2, 8>(H, wnode, nk, n, nw, z, sigma, mode, outp, errflag, ncta, kper, stream);   // norb <= 32 (experiment)
    }
}

}  // namespace abz
