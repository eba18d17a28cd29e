// vcb_aux.cu -- callers either side of the hot path (SURVEY.md section 8f "next" rows):
// push_delta (reference src/datasets.jl:6-13) and the post-processing of align
// (src/align.jl:19-34) on the device, so aligned pairs never leave the GPU.
#include "vcb_kernels.h"

namespace vcb {

namespace {

// out (2D, total): static copy on top, delta below; interior frames of each utterance get
// -0.5 x[t-1] + 0.5 x[t+1], the first and last frame keep delta = static (the reference's quirk).
__global__ void push_delta_kernel(const double* __restrict__ src, int64_t lds, int D, const int64_t* __restrict__ off,
                                  int64_t nseq, double* __restrict__ out, int64_t ldo) {
    const int64_t s = blockIdx.x;
    if (s >= nseq) return;
    const int64_t b = off[s], T = off[s + 1] - b;
    for (int64_t e = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < T * D;
         e += (int64_t)gridDim.y * blockDim.x) {
        const int64_t t = e / D;
        const int k = (int)(e - t * D);
        const double x = src[(b + t) * lds + k];
        double dl = x;
        if (t >= 1 && t + 1 < T) dl = -0.5 * src[(b + t - 1) * lds + k] + 0.5 * src[(b + t + 1) * lds + k];
        out[(b + t) * ldo + k] = x;
        out[(b + t) * ldo + D + k] = dl;
    }
}

// One block per pair.  newtgt[:, path[t]] = tgt[:, t] with the later duplicate winning
// (src/align.jl:21), then every state skipped by a 2-step is the mean of its neighbours
// (src/align.jl:25-32; with DTW(fstep=0, bstep=2) holes are isolated and interior, so the
// reference's in-place sequential pass has no carried dependence).
__global__ void align_post_kernel(const double* __restrict__ tgt, const int64_t* __restrict__ soff,
                                  const int64_t* __restrict__ toff, const int64_t* __restrict__ paths,
                                  int D, double* __restrict__ newtgt) {
    const int p = blockIdx.x;
    const int64_t sb = soff[p], S = soff[p + 1] - sb;
    const int64_t tb = toff[p], T = toff[p + 1] - tb;
    const int64_t* path = paths + tb;
    double* nt = newtgt + sb * D;
    for (int64_t e = threadIdx.x; e < S * D; e += blockDim.x) nt[e] = 0.0;  // zeros(size(src))
    __syncthreads();
    for (int64_t e = threadIdx.x; e < T * D; e += blockDim.x) {
        const int64_t t = e / D;
        const int k = (int)(e - t * D);
        const int64_t st = path[t] - 1;
        if (t == T - 1 || path[t + 1] - 1 != st) nt[st * D + k] = tgt[(tb + t) * D + k];
    }
    __syncthreads();
    for (int64_t e = threadIdx.x; e < (T - 1) * D; e += blockDim.x) {
        const int64_t t = e / D;
        const int k = (int)(e - t * D);
        const int64_t a = path[t] - 1, b = path[t + 1] - 1;
        if (b - a == 2) nt[(a + 1) * D + k] = (nt[a * D + k] + nt[b * D + k]) / 2.0;
    }
}

}  // namespace

int32_t push_delta_device(const double* d_src, int D, const int64_t* d_off, int64_t nseq,
                          int64_t total, double* d_out, cudaStream_t st) {
    return push_delta_strided_device(d_src, D, D, d_off, nseq, total, d_out, 2 * D, st);
}

int32_t push_delta_strided_device(const double* d_src, int64_t lds, int D, const int64_t* d_off, int64_t nseq,
                                  int64_t total, double* d_out, int64_t ldo, cudaStream_t st) {
    if (total == 0 || nseq == 0) return VCB_OK;
    const int64_t avg = (total * D + nseq - 1) / nseq;
    dim3 grid((unsigned)nseq, (unsigned)std::max<int64_t>(1, std::min<int64_t>((avg + 255) / 256, 64)));
    push_delta_kernel<<<grid, 256, 0, st>>>(d_src, lds, D, d_off, nseq, d_out, ldo);
    count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

int32_t align_post_device(const double* d_tgt, const int64_t* d_soff_src, const int64_t* d_soff_tgt,
                          const int64_t* d_paths, int64_t npairs, int D, double* d_newtgt,
                          cudaStream_t st) {
    if (npairs == 0) return VCB_OK;
    align_post_kernel<<<(unsigned)npairs, 256, 0, st>>>(d_tgt, d_soff_src, d_soff_tgt, d_paths, D, d_newtgt);
    count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

}  // namespace vcb
