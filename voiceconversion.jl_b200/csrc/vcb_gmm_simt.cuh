// vcb_gmm_simt.cuh -- the CUDA-core (fp32 FMA) posterior / conditional-mean kernel template.
// See vcb_gmm_simt.cu for the formulation; instantiated for conversion and arg-max in two
// translation units (vcb_gmm_simt_conv.cu / vcb_gmm_simt_argmax.cu) to keep builds parallel.
#pragma once
#include <cfloat>

#include "vcb_kernels.h"

namespace vcb {
namespace simt {


__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct SimtParams {
    const double* X; int64_t T; int64_t ldx;
    const float* w32; const float* c32; const double* xbar;
    int D, M;
    double* Y; int64_t ldy; int copy_power;        // convert
    int32_t* mhat; int* flag_count; int64_t* flag_list;  // arg-max
};

constexpr int kSimtThreads = 128;

template <int DP, int F, bool CONVERT>
__global__ void __launch_bounds__(kSimtThreads)
gmm_simt_kernel(const SimtParams p) {
    constexpr int KS = (DP + 1 + 3) / 4 * 4;
    constexpr int ROWS = CONVERT ? 2 * DP : DP;
    constexpr int MIX_STRIDE = 2 * DP * KS;   // floats between mixtures in w32
    extern __shared__ __align__(16) float wsm_dyn[];
    float* wsm[2] = {wsm_dyn, wsm_dyn + ROWS * KS};

    const int tid = threadIdx.x;
    const int64_t tbase = (int64_t)blockIdx.x * (kSimtThreads * F);

    // ---- [1 | xc | 0] per owned frame, centred in Float64
    float xs[F][KS];
    int64_t tf[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
        tf[f] = tbase + f * kSimtThreads + tid;
        const int64_t tl = tf[f] < p.T ? tf[f] : p.T - 1;
        const double* x = p.X + tl * p.ldx;
        xs[f][0] = 1.0f;
#pragma unroll
        for (int k = 0; k < KS - 1; ++k) xs[f][1 + k] = (k < p.D) ? (float)(x[k] - p.xbar[k]) : 0.0f;
    }

    float mx[F], sum[F], y[CONVERT ? F : 1][CONVERT ? DP : 1];
    float second[F], qbest[F];
    int best[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
        mx[f] = -INFINITY; sum[f] = 0.f; second[f] = -INFINITY; qbest[f] = 0.f; best[f] = 0;
        if (CONVERT) {
#pragma unroll
            for (int r = 0; r < DP; ++r) y[f][r] = 0.f;
        }
    }

    auto stage = [&](int m, int buf) {
        const float* src = p.w32 + (size_t)m * MIX_STRIDE;
        for (int c = tid; c < ROWS * KS / 4; c += kSimtThreads) cp_async16(wsm[buf] + c * 4, src + c * 4);
    };
    stage(0, 0);
    cp_async_commit();

    for (int m = 0; m < p.M; ++m) {
        const int buf = m & 1;
        if (m + 1 < p.M) stage(m + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const float* w = wsm[buf];

        // ---- whitening rows: q = |Linv xc + o|^2, lower-triangular => k <= r (+1 for the offset)
        float q[F];
#pragma unroll
        for (int f = 0; f < F; ++f) q[f] = 0.f;
#pragma unroll
        for (int r = 0; r < DP; ++r) {
            const float4* wr = reinterpret_cast<const float4*>(w + r * KS);
            float acc[F];
#pragma unroll
            for (int f = 0; f < F; ++f) acc[f] = 0.f;
#pragma unroll
            for (int k4 = 0; k4 < (r + 2 + 3) / 4; ++k4) {
                const float4 wv = wr[k4];
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    acc[f] = fmaf(wv.x, xs[f][4 * k4 + 0], acc[f]);
                    acc[f] = fmaf(wv.y, xs[f][4 * k4 + 1], acc[f]);
                    acc[f] = fmaf(wv.z, xs[f][4 * k4 + 2], acc[f]);
                    acc[f] = fmaf(wv.w, xs[f][4 * k4 + 3], acc[f]);
                }
            }
#pragma unroll
            for (int f = 0; f < F; ++f) q[f] = fmaf(acc[f], acc[f], q[f]);
        }
        const float cm = p.c32[m];

        if (CONVERT) {
            float wgt[F];
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const float l = fmaf(-0.5f, q[f], cm);
                if (l > mx[f]) {  // new running maximum: rescale what has been accumulated
                    const float a = expf(mx[f] - l);
                    sum[f] *= a;
#pragma unroll
                    for (int r = 0; r < DP; ++r) y[f][r] *= a;
                    mx[f] = l;
                }
                wgt[f] = expf(l - mx[f]);
                sum[f] += wgt[f];
            }
            // ---- regression rows: Ey = A xc + b, accumulated with the soft-max weight
#pragma unroll
            for (int r = 0; r < DP; ++r) {
                const float4* wr = reinterpret_cast<const float4*>(w + (DP + r) * KS);
                float acc[F];
#pragma unroll
                for (int f = 0; f < F; ++f) acc[f] = 0.f;
#pragma unroll
                for (int k4 = 0; k4 < KS / 4; ++k4) {
                    const float4 wv = wr[k4];
#pragma unroll
                    for (int f = 0; f < F; ++f) {
                        acc[f] = fmaf(wv.x, xs[f][4 * k4 + 0], acc[f]);
                        acc[f] = fmaf(wv.y, xs[f][4 * k4 + 1], acc[f]);
                        acc[f] = fmaf(wv.z, xs[f][4 * k4 + 2], acc[f]);
                        acc[f] = fmaf(wv.w, xs[f][4 * k4 + 3], acc[f]);
                    }
                }
#pragma unroll
                for (int f = 0; f < F; ++f) y[f][r] = fmaf(wgt[f], acc[f], y[f][r]);
            }
        } else {
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const float l = fmaf(-0.5f, q[f], cm);
                if (l > mx[f]) {  // strict: first maximum wins, like indmax (src/gmm.jl:46)
                    second[f] = mx[f]; mx[f] = l; best[f] = m; qbest[f] = q[f];
                } else if (l > second[f]) {
                    second[f] = l;
                }
            }
        }
        __syncthreads();
    }

#pragma unroll
    for (int f = 0; f < F; ++f) {
        if (tf[f] >= p.T) continue;
        if (CONVERT) {
            const float inv = 1.0f / sum[f];
            double* yo = p.Y + tf[f] * p.ldy;
#pragma unroll
            for (int r = 0; r < DP; ++r)
                if (r < p.D) yo[r] = (double)(y[f][r] * inv);
            if (p.copy_power) yo[-1] = p.X[tf[f] * p.ldx - 1];  // src/common.jl:23
        } else {
            p.mhat[tf[f]] = best[f];
            // near tie at fp32 accuracy -> exact Float64 re-check
            if (mx[f] - second[f] < 1e-3f * (1.0f + qbest[f]) || !(mx[f] == mx[f])) {
                const int slot = atomicAdd(p.flag_count, 1);
                p.flag_list[slot] = tf[f];
            }
        }
    }
}

template <int DP, bool CONVERT>
int32_t launch_simt(const SimtParams& p, cudaStream_t st) {
    constexpr int F = (DP <= 32) ? 2 : 1;
    constexpr int KS = (DP + 1 + 3) / 4 * 4;
    constexpr int ROWS = CONVERT ? 2 * DP : DP;
    constexpr size_t smem = (size_t)2 * ROWS * KS * sizeof(float);
    const int64_t per_block = (int64_t)kSimtThreads * F;
    const unsigned grid = (unsigned)((p.T + per_block - 1) / per_block);
    auto k = gmm_simt_kernel<DP, F, CONVERT>;
    if (smem > 48 * 1024) VCB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kSimtThreads, smem, st>>>(p);
    count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

template <bool CONVERT>
int32_t dispatch_simt(int DS, const SimtParams& p, cudaStream_t st) {
    switch (DS) {
        case 8: return launch_simt<8, CONVERT>(p, st);
        case 16: return launch_simt<16, CONVERT>(p, st);
        case 24: return launch_simt<24, CONVERT>(p, st);
        case 32: return launch_simt<32, CONVERT>(p, st);
        case 40: return launch_simt<40, CONVERT>(p, st);
        case 48: return launch_simt<48, CONVERT>(p, st);
        case 64: return launch_simt<64, CONVERT>(p, st);
        case 80: return launch_simt<80, CONVERT>(p, st);
        case 96: return launch_simt<96, CONVERT>(p, st);
        default: return fail(VCB_EUNSUPPORTED, "feature dimension %d exceeds the CUDA-core kernel's limit (96)", p.D);
    }
}

int32_t dispatch_convert(int DS, const SimtParams& p, cudaStream_t st);
int32_t dispatch_argmax(int DS, const SimtParams& p, cudaStream_t st);

}  // namespace simt
}  // namespace vcb
