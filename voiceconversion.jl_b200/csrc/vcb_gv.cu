// vcb_gv.cu -- global-variance helpers, Float64.
//
//   * VarianceScaling post-filter, reference src/gv.jl:10-15:
//       src = sqrt(s2 ./ var(src, 2)) .* (src .- mean(src, 2)) .+ mean(src, 2)      per utterance
//   * TrajectoryGVGMMMap gradient ascent, reference src/trajectory_gmmmap.jl:140-189:
//       y0 = fvconvert(tgmm, X);  y0 <- eq. (58) (the filter above with s2 = mu_v)
//       repeat epochs:  dy = w (-(W'D^-1 W) y + W'D^-1 E) + gvgrad(y);  y += alpha dy,  w = 1/(2T)
//       gvgrad(y)[:,t] = -2/T (pv' (var(y,2) - mu_v)) .* (y[:,t] - mean(y,2))
//     W and D^-1 are never materialised: -(W'D^-1 W) y + W'D^-1 E = W' (P_t (E_t - (W y)_t))_t with
//     (W y)_t = [y_t; 1/2 (y_{t+1} - y_{t-1})] inside the chunk, so one epoch is a per-frame
//     2Ds x 2Ds mat-vec (gv_h_kernel, all frames of the batch in one launch) followed by a
//     per-chunk reduction + update (gv_update_kernel).
#include "vcb_kernels.h"

namespace vcb {

namespace {

// Block-wide sum of one double per thread (blockDim.x <= 1024), result broadcast to all threads.
__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += red[w];
    return t;
}

// One CTA per (sequence, dimension): mean, corrected variance, affine rescale.
__global__ void variance_scaling_kernel(const double* __restrict__ s2, int D, const double* __restrict__ X,
                                        int64_t ldx, const int64_t* __restrict__ off, double* __restrict__ Y,
                                        int64_t ldy) {
    __shared__ double red[32];
    const int64_t b = off[blockIdx.x], T = off[blockIdx.x + 1] - b;
    const int i = blockIdx.y;
    if (T <= 0 || i >= D) return;
    double s = 0.0;
    for (int64_t t = threadIdx.x; t < T; t += blockDim.x) s += X[(b + t) * ldx + i];
    const double mu = block_sum(s, red) / (double)T;
    double q = 0.0;
    for (int64_t t = threadIdx.x; t < T; t += blockDim.x) {
        const double d = X[(b + t) * ldx + i] - mu;
        q = fma(d, d, q);
    }
    const double var = block_sum(q, red) / (double)(T - 1);
    const double sc = sqrt(s2[i] / var);
    for (int64_t t = threadIdx.x; t < T; t += blockDim.x) Y[(b + t) * ldy + i] = sc * (X[(b + t) * ldx + i] - mu) + mu;
}

// h_t = P_{m_t} (E_t - (W y)_t) for every frame of the batch.  blockDim = (tpf >= D2, frames per block).
// edge[t]: bit 0 = first frame of its chunk, bit 1 = last frame.
__global__ void gv_h_kernel(const double* __restrict__ Y, int64_t ldy, const double* __restrict__ E,
                            const int64_t* __restrict__ mhat, const unsigned char* __restrict__ edge,
                            const double* __restrict__ P, int Ds, int64_t total, double* __restrict__ H) {
    extern __shared__ double sm[];   // per frame: d[D2]
    const int D2 = 2 * Ds, f = threadIdx.y, i = threadIdx.x;
    const int64_t t = (int64_t)blockIdx.x * blockDim.y + f;
    double* d = sm + (size_t)f * D2;
    const bool live = t < total;
    if (live && i < D2) {
        double wy;
        if (i < Ds) {
            wy = Y[t * ldy + i];
        } else {
            const unsigned char e = edge[t];
            const int k = i - Ds;
            wy = 0.0;
            if (!(e & 1)) wy = -0.5 * Y[(t - 1) * ldy + k];
            if (!(e & 2)) wy = fma(0.5, Y[(t + 1) * ldy + k], wy);
        }
        d[i] = E[t * D2 + i] - wy;
    }
    __syncthreads();
    if (live && i < D2) {
        const double* pm = P + (size_t)(mhat[t] - 1) * D2 * D2 + i;   // symmetric: row i == column i
        double s0 = 0.0, s1 = 0.0;
        int k = 0;
        for (; k + 1 < D2; k += 2) {
            s0 = fma(pm[(size_t)k * D2], d[k], s0);
            s1 = fma(pm[(size_t)(k + 1) * D2], d[k + 1], s1);
        }
        if (k < D2) s0 = fma(pm[(size_t)k * D2], d[k], s0);
        H[t * D2 + i] = s0 + s1;
    }
}

// One CTA per chunk: global variance / mean of the current y, GV gradient coefficient, update.
__global__ void gv_update_kernel(double* __restrict__ Y, int64_t ldy, const double* __restrict__ H,
                                 const int64_t* __restrict__ chunk_off, const double* __restrict__ muv,
                                 const double* __restrict__ pv, int Ds, double alpha, int* __restrict__ err) {
    extern __shared__ double sm[];   // mu[Ds] | gvd[Ds] | coef[Ds] | part[nwarps][Ds]
    const int64_t b = chunk_off[blockIdx.x];
    const int T = (int)(chunk_off[blockIdx.x + 1] - b);
    if (T <= 0) return;
    const int D2 = 2 * Ds, nth = blockDim.x;
    double* mu = sm;
    double* gvd = sm + Ds;
    double* coef = sm + 2 * Ds;
    double* part = sm + 3 * Ds;
    // thread -> (dimension i, frame phase): consecutive threads read consecutive dimensions
    const int i = threadIdx.x % Ds, ph = threadIdx.x / Ds, nph = nth / Ds;
    const bool act = ph < nph;
    auto reduce_dim = [&](double v, double* out, double scale) {
        __syncthreads();
        if (act) part[ph * Ds + i] = v;
        __syncthreads();
        if (threadIdx.x < Ds) {
            double s = 0.0;
            for (int p = 0; p < nph; ++p) s += part[p * Ds + threadIdx.x];
            out[threadIdx.x] = s * scale;
        }
        __syncthreads();
    };
    double s = 0.0;
    if (act) for (int t = ph; t < T; t += nph) s += Y[(b + t) * ldy + i];
    reduce_dim(s, mu, 1.0 / (double)T);
    double q = 0.0;
    if (act) for (int t = ph; t < T; t += nph) { const double d = Y[(b + t) * ldy + i] - mu[i]; q = fma(d, d, q); }
    reduce_dim(q, gvd, 1.0 / (double)(T - 1));
    if (threadIdx.x < Ds) {
        double c = 0.0;
        for (int k = 0; k < Ds; ++k) c = fma(pv[k + (size_t)threadIdx.x * Ds], gvd[k] - muv[k], c);   // pv'
        coef[threadIdx.x] = -2.0 / (double)T * c;
    }
    __syncthreads();
    const double om = 1.0 / (2.0 * (double)T);
    bool bad = false;
    if (act) for (int t = ph; t < T; t += nph) {
        const double* h = H + (b + t) * D2;
        double g = h[i];
        if (t >= 1) g = fma(0.5, h[-D2 + Ds + i], g);
        if (t + 1 < T) g = fma(-0.5, h[D2 + Ds + i], g);
        const double y = Y[(b + t) * ldy + i];
        const double dy = om * g + coef[i] * (y - mu[i]);
        bad |= (dy != dy);
        Y[(b + t) * ldy + i] = y + alpha * dy;
    }
    if (bad) atomicExch(err, 1);   // @assert !any(isnan(dy))  src/trajectory_gmmmap.jl:165
}

__global__ void chunk_edges_kernel(const int64_t* __restrict__ chunk_off, int64_t nchunks, unsigned char* __restrict__ edge) {
    const int64_t c = blockIdx.x;
    const int64_t b = chunk_off[c], e = chunk_off[c + 1];
    for (int64_t t = b + threadIdx.x; t < e; t += blockDim.x) edge[t] = (unsigned char)((t == b ? 1 : 0) | (t == e - 1 ? 2 : 0));
}

}  // namespace

int32_t variance_scaling_device(const double* d_s2, int D, const double* dX, int64_t ldx, const int64_t* d_off,
                                int64_t nseq, double* dY, int64_t ldy, cudaStream_t st) {
    if (nseq <= 0 || D <= 0) return VCB_OK;
    dim3 grid((unsigned)nseq, (unsigned)D);
    variance_scaling_kernel<<<grid, 128, 0, st>>>(d_s2, D, dX, ldx, d_off, dY, ldy);
    count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

int32_t trajgv_ascent_device(const vcb_trajgv& v, double* dY, int64_t ldy, const double* dE, const int64_t* d_mhat,
                             const int64_t* d_chunk_off, int64_t nchunks, int64_t total, int epochs, double alpha,
                             cudaStream_t st, const int* ws, int64_t npanels) {
    const vcb_traj& tr = *v.t;
    const int Ds = tr.Ds, D2 = 2 * Ds;
    if (total == 0 || nchunks == 0) return VCB_OK;
    double* dH = nullptr;
    unsigned char* d_edge = nullptr;
    int* derr = nullptr;
    VCB_CUDA(cudaMallocAsync((void**)&dH, (size_t)total * D2 * sizeof(double), st));
    VCB_CUDA(cudaMallocAsync((void**)&d_edge, (size_t)total, st));
    VCB_CUDA(cudaMallocAsync((void**)&derr, sizeof(int), st));
    VCB_CUDA(cudaMemsetAsync(derr, 0, sizeof(int), st));
    chunk_edges_kernel<<<(unsigned)nchunks, 128, 0, st>>>(d_chunk_off, nchunks, d_edge);
    count_launch();
    // eq. (58): better initial value
    int32_t rc = variance_scaling_device(v.d_muv.p, Ds, dY, ldy, d_chunk_off, nchunks, dY, ldy, st);
    const int tpf = round_up(D2, 32), fpb = std::max(1, 256 / tpf);
    const dim3 hblock(tpf, fpb), hgrid((unsigned)((total + fpb - 1) / fpb));
    const int uth = std::min(1024, round_up(std::max(Ds * 8, 256), 32));
    const size_t usm = (size_t)(3 * Ds + (uth / Ds) * Ds) * sizeof(double);
    for (int e = 0; e < epochs && rc == VCB_OK; ++e) {
        if (ws) {
            rc = group_gv_step(tr, ws, npanels, dY, ldy, dE, d_edge, dH, st);
            if (rc != VCB_OK) break;
        } else {
            gv_h_kernel<<<hgrid, hblock, (size_t)fpb * D2 * sizeof(double), st>>>(dY, ldy, dE, d_mhat, d_edge, tr.d_P.p, Ds, total, dH);
            count_launch();
        }
        gv_update_kernel<<<(unsigned)nchunks, uth, usm, st>>>(dY, ldy, dH, d_chunk_off, v.d_muv.p, v.d_pv.p, Ds, alpha, derr);
        count_launch();
    }
    if (rc == VCB_OK && cudaGetLastError() != cudaSuccess) rc = fail(VCB_ECUDA, "GV ascent launch failed");
    int herr = 0;
    if (rc == VCB_OK) {
        cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        if (herr) rc = fail(VCB_EARG, "GV gradient became NaN (the reference asserts, src/trajectory_gmmmap.jl:165)");
    }
    cudaFreeAsync(dH, st);
    cudaFreeAsync(d_edge, st);
    cudaFreeAsync(derr, st);
    return rc;
}

}  // namespace vcb
