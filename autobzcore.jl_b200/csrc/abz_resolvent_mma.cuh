// K3-fast: register-resident blocked resolvent trace on the FP64 tensor cores (DMMA).
// (placeholder interface; implemented below the generic path)
#pragma once
#include "abz_common.cuh"

namespace abz {
inline bool mma_resolvent_supported(int n) { (void)n; return false; }
inline int mma_resolvent_partial_count(int n, long nk, int nw, long sm, long* ncta) { (void)n; (void)nk; (void)nw; (void)sm; *ncta = 0; return -1; }
inline void mma_resolvent_launch(const double2*, const double*, long, int, int, const double2*, const double2*, int, double2*, int*, long, cudaStream_t) {}
}  // namespace abz
