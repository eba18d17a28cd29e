// vcb_gmm_simt.cu -- K1/K2 on CUDA cores (fp32 FMA) + the Float64 exact-path helpers.
//
// Replaces, per frame x_t (reference src/gmmmap.jl:101-118, src/gmm.jl:24-58):
//   posterior P(m|x_t) from  lpr_m = log w_m + logpdf(N(mux_m, Sxx_m), x_t)   (src/gmm.jl:26-29)
//   Ey_m = muy_m + A_m (x_t - mux_m)                                          (src/gmmmap.jl:109-111)
//   y_t  = sum_m P(m|x_t) Ey_m                                                (src/gmmmap.jl:117)
// and the arg-max mixture sequence of the trajectory converter (src/gmm.jl:44-58).
//
// Formulation: with Linv_m = chol(Sxx_m)^-1 and the centred frame xc = x - xbar (Float64
// subtraction, then fp32), z_m = Linv_m xc + o_m and Ey_m = A_m xc + b_m are rows of one operand
// matrix per mixture, [offset | coefficients]; the log-likelihood is c_m - |z_m|^2 / 2.  One
// thread owns F frames (their [1 | xc] vectors live in registers); the rows of mixture m are
// staged in shared memory with cp.async (double-buffered) and read as warp-uniform broadcasts.
// The posterior-weighted sum is an online soft-max over mixtures, so neither the per-mixture
// means nor the log-likelihoods ever leave registers.  The triangular structure of Linv halves
// the whitening work.
//
// This is the general-shape path; vcb_gmm_tc.cu holds the tcgen05 (3xTF32) kernel.
#include "vcb_gmm_simt.cuh"

namespace vcb {

namespace {

// Float64 log-likelihoods of one frame for mixtures m = tid, tid + nthreads, ...
//   lpr_m = c_m - |Linv_m (x - mux_m)|^2 / 2
__device__ __forceinline__ double loglik_fp64(const double* __restrict__ linv, const double* __restrict__ mux,
                                              double cm, const double* xs /*smem x*/, int D) {
    double q = 0.0;
    for (int r = 0; r < D; ++r) {
        const double* row = linv + (size_t)r * D;
        double z = 0.0;
        for (int k = 0; k <= r; ++k) z = fma(row[k], xs[k] - mux[k], z);
        q = fma(z, z, q);
    }
    return cm - 0.5 * q;
}

constexpr int kExactThreads = 128;

// Exact arg-max for the frames the fp32/tf32 kernels flagged as near ties.  One block per flagged
// frame (grid-stride); warp w takes mixtures w, w+8, ...; lane = row r of z = Linv_m (x - mux_m),
// read from the column-major copy of Linv so every k step is one coalesced load.
constexpr int kRecheckThreads = 256;
__global__ void __launch_bounds__(kRecheckThreads)
recheck_argmax_kernel(const double* __restrict__ X, int64_t ldx, const int* __restrict__ flag_count,
                      const int64_t* __restrict__ flag_list, const double* __restrict__ linv_cm,
                      const double* __restrict__ mux, const double* __restrict__ c, int D, int M,
                      int32_t* __restrict__ mhat, int start) {
    extern __shared__ double xs[];  // [D]
    __shared__ double rv[kRecheckThreads / 32];
    __shared__ int ri[kRecheckThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kRecheckThreads / 32;
    const int n = *flag_count;
    for (int e = start + blockIdx.x; e < n; e += gridDim.x) {
        const int64_t t = flag_list[e];
        __syncthreads();
        for (int k = threadIdx.x; k < D; k += blockDim.x) xs[k] = X[t * ldx + k];
        __syncthreads();
        double bv = -INFINITY;
        int bi = 0x7FFFFFFF;
        for (int m = warp; m < M; m += nwarps) {
            const double* lm = linv_cm + (size_t)m * D * D;  // lm[k*D + r] = Linv[r][k]
            const double* mu = mux + (size_t)m * D;
            double q = 0.0;
            for (int r0 = 0; r0 < D; r0 += 32) {
                const int r = r0 + lane;
                double z = 0.0;
                if (r < D) {
                    // four independent partial sums: the loop is a dependent-FMA latency chain otherwise
                    const int kend = min(r0 + 32, D);
                    double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
                    int k = 0;
                    for (; k + 3 < kend; k += 4) {
                        z0 = fma(lm[(size_t)k * D + r], xs[k] - mu[k], z0);
                        z1 = fma(lm[(size_t)(k + 1) * D + r], xs[k + 1] - mu[k + 1], z1);
                        z2 = fma(lm[(size_t)(k + 2) * D + r], xs[k + 2] - mu[k + 2], z2);
                        z3 = fma(lm[(size_t)(k + 3) * D + r], xs[k + 3] - mu[k + 3], z3);
                    }
                    for (; k < kend; ++k) z0 = fma(lm[(size_t)k * D + r], xs[k] - mu[k], z0);
                    z = (z0 + z1) + (z2 + z3);
                }
                q = fma(z, z, q);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
            const double l = c[m] - 0.5 * q;
            if (l > bv) { bv = l; bi = m; }  // ascending m per warp: first max kept
        }
        if (lane == 0) { rv[warp] = bv; ri[warp] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < nwarps; ++w)
                if (rv[w] > bv || (rv[w] == bv && ri[w] < bi)) { bv = rv[w]; bi = ri[w]; }
            if (bi != 0x7FFFFFFF) mhat[t] = bi;
        }
    }
}

__global__ void __launch_bounds__(kExactThreads)
proba_fp64_kernel(const double* __restrict__ X, int64_t T, int64_t ldx, const double* __restrict__ linv,
                  const double* __restrict__ mux, const double* __restrict__ c, int D, int M,
                  double* __restrict__ post) {
    extern __shared__ double sm[];  // [D] x, then [M] lpr
    double* xs = sm;
    double* lpr = sm + D;
    __shared__ double red[kExactThreads / 32];
    __shared__ double bcast;
    for (int64_t t = blockIdx.x; t < T; t += gridDim.x) {
        __syncthreads();
        for (int k = threadIdx.x; k < D; k += blockDim.x) xs[k] = X[t * ldx + k];
        __syncthreads();
        double u = -INFINITY;
        for (int m = threadIdx.x; m < M; m += blockDim.x) {
            const double l = loglik_fp64(linv + (size_t)m * D * D, mux + (size_t)m * D, c[m], xs, D);
            lpr[m] = l;
            u = fmax(u, l);
        }
        for (int o = 16; o > 0; o >>= 1) u = fmax(u, __shfl_xor_sync(0xFFFFFFFFu, u, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = u;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < kExactThreads / 32; ++w) u = fmax(u, red[w]);
            // (ext) StatsFuns.logsumexp, summed in index order like the reference
            double s = 0.0;
            for (int m = 0; m < M; ++m) s += exp(lpr[m] - u);
            bcast = isinf(u) ? u : (log(s) + u);
        }
        __syncthreads();
        const double logprob = bcast;
        for (int m = threadIdx.x; m < M; m += blockDim.x) post[t * M + m] = exp(lpr[m] - logprob);
    }
}

__global__ void widen_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int64_t)in[i] + 1;  // 1-based like Julia
}

}  // namespace

int simt_padded_dim(int D) {
    static const int sizes[] = {8, 16, 24, 32, 40, 48, 64, 80, 96};
    for (int s : sizes)
        if (D <= s) return s;
    return 0;
}

int32_t simt_convert(const vcb_gmmmap& g, const double* dX, int64_t T, int64_t ldx, double* dY,
                     int64_t ldy, bool copy_power, cudaStream_t st) {
    if (T == 0) return VCB_OK;
    simt::SimtParams p{};
    p.X = dX; p.T = T; p.ldx = ldx; p.w32 = g.d_w32.p; p.c32 = g.d_c32.p; p.xbar = g.d_xbar.p;
    p.D = g.D; p.M = g.M; p.Y = dY; p.ldy = ldy; p.copy_power = copy_power ? 1 : 0;
    if (!g.DS) return fail(VCB_EUNSUPPORTED, "feature dimension %d exceeds the CUDA-core kernel's limit (96)", g.D);
    return simt::dispatch_convert(g.DS, p, st);
}

int32_t recheck_argmax_fp64(const vcb_gmmmap& g, const double* dX, int64_t ldx, const int* d_flag_count,
                            const int64_t* d_flag_list, int32_t* d_mhat, cudaStream_t st) {
    // the first kPanelCap flagged frames go through the panel kernels (Linv_m staged once per 64 frames);
    // whatever is beyond -- pathological inputs where most frames are near ties -- through the per-frame one
    constexpr int kPanelCap = 32768;
    static const bool per_frame = [] { const char* e = getenv("VCB_RECHECK"); return e && e[0] == 'f'; }();
    int start = 0;
    double* d_lout = nullptr;
    if (!per_frame && g.D <= 256) {
        VCB_CUDA(cudaMallocAsync((void**)&d_lout, (size_t)kPanelCap * g.M * sizeof(double), st));
        int32_t rc = recheck_argmax_panels(g, dX, ldx, d_flag_count, d_flag_list, d_mhat, kPanelCap, d_lout, st);
        if (rc != VCB_OK) { cudaFreeAsync(d_lout, st); return rc; }
        start = kPanelCap;
    }
    recheck_argmax_kernel<<<148 * 4, kRecheckThreads, g.D * sizeof(double), st>>>(
        dX, ldx, d_flag_count, d_flag_list, g.d_linv_cm.p, g.d_mux.p, g.d_c.p, g.D, g.M, d_mhat, start);
    count_launch();
    if (d_lout) cudaFreeAsync(d_lout, st);
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

int32_t simt_argmax(const vcb_gmmmap& g, const double* dX, int64_t T, int64_t ldx, int32_t* d_mhat,
                    cudaStream_t st) {
    if (T == 0) return VCB_OK;
    if (!g.DS) return fail(VCB_EUNSUPPORTED, "feature dimension %d exceeds the CUDA-core kernel's limit (96)", g.D);
    int* d_count = nullptr;
    int64_t* d_list = nullptr;
    VCB_CUDA(cudaMallocAsync((void**)&d_count, sizeof(int), st));
    VCB_CUDA(cudaMallocAsync((void**)&d_list, (size_t)T * sizeof(int64_t), st));
    VCB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), st));
    simt::SimtParams p{};
    p.X = dX; p.T = T; p.ldx = ldx; p.w32 = g.d_w32.p; p.c32 = g.d_c32.p; p.xbar = g.d_xbar.p;
    p.D = g.D; p.M = g.M; p.mhat = d_mhat; p.flag_count = d_count; p.flag_list = d_list;
    int32_t rc = simt::dispatch_argmax(g.DS, p, st);
    if (rc == VCB_OK) rc = recheck_argmax_fp64(g, dX, ldx, d_count, d_list, d_mhat, st);
    cudaFreeAsync(d_count, st);
    cudaFreeAsync(d_list, st);
    return rc;
}

int32_t proba_fp64(const vcb_gmmmap& g, const double* dX, int64_t T, int64_t ldx, double* d_post,
                   cudaStream_t st) {
    if (T == 0) return VCB_OK;
    const unsigned grid = (unsigned)std::min<int64_t>(T, 148 * 16);
    proba_fp64_kernel<<<grid, kExactThreads, (g.D + g.M) * sizeof(double), st>>>(
        dX, T, ldx, g.d_linv.p, g.d_mux.p, g.d_c.p, g.D, g.M, d_post);
    count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

int32_t widen_mhat(const int32_t* d_mhat, int64_t T, int64_t* d_out, cudaStream_t st) {
    if (T == 0) return VCB_OK;
    widen_kernel<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(d_mhat, T, d_out);
    count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

}  // namespace vcb
