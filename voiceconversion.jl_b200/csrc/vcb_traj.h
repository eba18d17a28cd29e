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

#ifdef __CUDACC__
// 2^-100 <= d < 2^100, tested on the exponent field with integer compares (negative numbers, NaN and infinities fall
// outside as unsigned values): the check feeds only the error flag and should not take FP64-pipe slots next to the
// factorisation's serial chain.
__device__ __forceinline__ bool sane_pivot(double d) {
    const unsigned hi = (unsigned)__double2hiint(d);
    return hi - 0x39B00000u < 0x46300000u - 0x39B00000u;          // biased exponents 1023 - 100 .. 1023 + 99
}

// 1/sqrt(d) to about one ulp from the hardware's 20-bit approximation (rsqrt.approx.f64 = MUFU.RSQ64H, one instruction
// on the high word -- the fp32 route costs two conversions more on the serial chain) and one third-order correction
// (e = 1 - d y^2,  y += y e (1/2 + 3/8 e); |e| < 2^-19, so the remainder 5/16 e^3 is below 2^-58): 6 instructions and a
// ~55-cycle chain instead of the library's 14 / 110 -- this sits 24 times per frame on the serial chain of the
// factorisation.
__device__ __forceinline__ double fast_rsqrt(double d) {
    double y;                                                   // (the caller flags pivots outside 2^-100 .. 2^100)
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double t = d * y;
    const double e = fma(-t, y, 1.0);
    const double c = fma(0.375, e, 0.5);
    return fma(y * e, c, y);
}
#endif

}  // namespace vcb
