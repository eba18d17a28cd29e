// vcb_model.h -- host-side model preprocessing (K0), see vcb_model.cpp.
#pragma once
#include "vcb_common.h"

namespace vcb {

int32_t build_gmmmap(const double* weights, const double* mu, const double* sigma, int twoD, int M,
                     int swap, vcb_gmmmap& g);
int32_t build_traj(const vcb_gmmmap& g, vcb_traj& t);
void tf32_split(double v, float& hi, float& lo);
// dense inverse like Julia's inv / ^-1 (false: singular); a is n x n column-major, inverted in place
bool invert_matrix(std::vector<double>& a, int n);
// diffgmm (src/diffgmm.jl:9-25) on the joint parameters mu (2D,M), sigma (2D,2D,M)
void diffgmm_params(const double* mu, const double* sigma, int twoD, int M, double* mu_out, double* sigma_out);

// Tile plan of the tcgen05 kernels (shared by the packer and the launcher; vcb_fbf_tc.cu).
struct TcPlan {
    int G = 0;       // mixtures per MMA chunk (0 = shape not supported by the tensor-core kernel)
    int N = 0;       // MMA N = G * rows_per_mixture
    int stages = 0;  // B-operand pipeline depth
    int abufs = 0;   // A-operand (frame tile) buffers
    size_t smem = 0; // dynamic shared memory bytes
};
// rows_per_mixture = DP (whitening only) or 2*DP (whitening + regression); part_rows = floats per
// frame of the epilogue groups' merge buffer.
TcPlan tc_plan(int M, int KP, int rows_per_mixture, int part_rows, bool pair = false, int force_g = 0);

}  // namespace vcb
