// Conversion instantiations of the CUDA-core kernel (see vcb_gmm_simt.cuh).
#include "vcb_gmm_simt.cuh"
namespace vcb { namespace simt {
int32_t dispatch_convert(int DS, const SimtParams& p, cudaStream_t st) { return dispatch_simt<true>(DS, p, st); }
}}
