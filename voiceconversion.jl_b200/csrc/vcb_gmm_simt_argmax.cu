// Arg-max instantiations of the CUDA-core kernel (see vcb_gmm_simt.cuh).
#include "vcb_gmm_simt.cuh"
namespace vcb { namespace simt {
int32_t dispatch_argmax(int DS, const SimtParams& p, cudaStream_t st) { return dispatch_simt<false>(DS, p, st); }
}}
