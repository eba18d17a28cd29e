// vcb_traj.h -- shared between the two trajectory-solver translation units.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vcb {

struct TrajParams {
    const double* P;        // [M][D2*D2] symmetric precision blocks
    const int32_t* mhat;    // [total] 0-based
    const double* Gv;       // [total][D2]  g_t = P_t E_t
    const int64_t* chunk_off;
    double* Lst;            // per-frame factor blocks Linv_tt, L[t][t-1], L[t][t-2] (layout owned by the solver)
    double* Z;              // [total][Ds]
    double* Y; int64_t ldy;
    const double* Xpow; int64_t ldx; int copy_power;
    int Ds;
    int* err;
    int role_rule;          // experiment switch of the two-warp solver (0 = default)
    int64_t nchunks;
};

// Warp-per-chunk solver (vcb_traj_warp.cu): Ds <= 24.  Bytes of factor scratch per frame, 0 if the
// dimension is not covered.
size_t traj_warp_factor_bytes(int Ds);
int32_t traj_warp_launch(const TrajParams& p, int64_t nchunks, cudaStream_t st);

}  // namespace vcb
