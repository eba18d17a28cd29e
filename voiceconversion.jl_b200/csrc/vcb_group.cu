// vcb_group.cu -- per-mixture grouped Float64 products for the trajectory path.
//
// E_t = muy_m + A_m (x_t - mux_m) and g_t = P_m E_t (reference src/trajectory_gmmmap.jl:85-89 and the
// right-hand side of :103-105) use the matrices of the arg-max mixture m = mhat_t of every frame.  A
// per-frame mat-vec streams 2 x 18 KB of A_m / P_m from L2 for 48 outputs; instead the frames are
// bucketed by mixture (counting sort, order inside a bucket irrelevant) and every CTA multiplies one
// mixture's matrices, held in shared memory, with a 64-frame panel: a Float64 GEMM per mixture.
// The same kernel serves the GV gradient (h_t = P_m (E_t - (W y)_t), src/trajectory_gmmmap.jl:161).
#include <cstdlib>

#include "vcb_kernels.h"

namespace vcb {

namespace {

constexpr int kFT = 64;          // frames per CTA panel
constexpr int kFP = kFT + 2;     // padded panel row (keeps 16-byte alignment, spreads banks)

// Bucket counts: per-block histogram in shared memory, one global atomic per (block, mixture) --
// 500 k same-address global atomics on 64 counters cost 0.3 ms at C2.
constexpr int kHistBlock = 256, kHistPer = 16;     // frames per block = 4096
__global__ void group_hist_kernel(const int32_t* __restrict__ mhat, int64_t total, int M, int* __restrict__ hist) {
    extern __shared__ int sh[];
    for (int m = threadIdx.x; m < M; m += blockDim.x) sh[m] = 0;
    __syncthreads();
    const int64_t b0 = (int64_t)blockIdx.x * kHistBlock * kHistPer;
    for (int i = 0; i < kHistPer; ++i) {
        const int64_t t = b0 + (int64_t)i * kHistBlock + threadIdx.x;
        if (t < total) atomicAdd(&sh[mhat[t]], 1);
    }
    __syncthreads();
    for (int m = threadIdx.x; m < M; m += blockDim.x)
        if (sh[m]) atomicAdd(&hist[m], sh[m]);
}

// one block: bucket starts, panel starts, scatter cursors
__global__ void group_scan_kernel(const int* __restrict__ hist, int M, int* __restrict__ mstart,
                                  int* __restrict__ tstart, int* __restrict__ cursor) {
    if (threadIdx.x == 0) {
        int a = 0, b = 0;
        for (int m = 0; m < M; ++m) {
            mstart[m] = a; tstart[m] = b; cursor[m] = a;
            a += hist[m];
            b += (hist[m] + kFT - 1) / kFT;
        }
        mstart[M] = a; tstart[M] = b;
    }
}

// Scatter: the block counts its frames per mixture, reserves one range per mixture with a single
// global atomic, then places its frames with shared-memory atomics.
__global__ void group_scatter_kernel(const int32_t* __restrict__ mhat, int64_t total, int M, int* __restrict__ cursor,
                                     int32_t* __restrict__ perm) {
    extern __shared__ int sh[];          // count[M] | base[M]
    int* cnt = sh;
    int* base = sh + M;
    for (int m = threadIdx.x; m < M; m += blockDim.x) cnt[m] = 0;
    __syncthreads();
    const int64_t b0 = (int64_t)blockIdx.x * kHistBlock * kHistPer;
    int slot[kHistPer];
#pragma unroll
    for (int i = 0; i < kHistPer; ++i) {
        const int64_t t = b0 + (int64_t)i * kHistBlock + threadIdx.x;
        slot[i] = (t < total) ? atomicAdd(&cnt[mhat[t]], 1) : 0;
    }
    __syncthreads();
    for (int m = threadIdx.x; m < M; m += blockDim.x) base[m] = cnt[m] ? atomicAdd(&cursor[m], cnt[m]) : 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kHistPer; ++i) {
        const int64_t t = b0 + (int64_t)i * kHistBlock + threadIdx.x;
        if (t < total) perm[base[mhat[t]] + slot[i]] = (int32_t)t;
    }
}

struct GroupParams {
    const int32_t* perm; const int* mstart; const int* tstart; int M;
    const double* A; const double* mux; const double* muy; const double* P;   // per mixture, D2 x D2 column-major
    int D2;
    // MODE 0: panel = x - mux;  E = muy + A panel (-> E, Eout);  G = P E (-> G)
    const double* X; int64_t ldx; double* E; double* Eout; double* G;
    // MODE 1: panel = E - (W y);  H = P panel (-> G)
    const double* Y; int64_t ldy; const unsigned char* edge;
};

// out[4][4] += Mat[i0..i0+3][k] * panel[k][f0..f0+3] over k
__device__ __forceinline__ void panel_product(const double* __restrict__ mat, int ldm, const double* __restrict__ pan,
                                              int D2, int i0, int f0, double (&o)[4][4]) {
#pragma unroll 2
    for (int k = 0; k < D2; ++k) {
        const double2 a01 = *reinterpret_cast<const double2*>(mat + (size_t)k * ldm + i0);
        const double2 a23 = *reinterpret_cast<const double2*>(mat + (size_t)k * ldm + i0 + 2);
        const double2 b01 = *reinterpret_cast<const double2*>(pan + (size_t)k * kFP + f0);
        const double2 b23 = *reinterpret_cast<const double2*>(pan + (size_t)k * kFP + f0 + 2);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) o[r][c] = fma(a[r], b[c], o[r][c]);
    }
}

template <int MODE>
__global__ void group_panel_kernel(const GroupParams p) {
    extern __shared__ __align__(16) double sm[];
    const int D2 = p.D2, D2p = (D2 + 3) & ~3;
    double* mat = sm;                          // [D2][D2p]  (column k at mat + k*D2p)
    double* pan = mat + (size_t)D2 * D2p;      // [D2][kFP]  panel, row k = reduction index; MODE 0 overwrites it
    double* mid = pan;                         //            with the E panel once every thread has read it
    __shared__ int s_m, s_first, s_n;
    if (threadIdx.x == 0) {
        const int tile = blockIdx.x;
        int lo = 0, hi = p.M;                  // largest m with tstart[m] <= tile
        while (hi - lo > 1) { const int mid_m = (lo + hi) >> 1; if (p.tstart[mid_m] <= tile) lo = mid_m; else hi = mid_m; }
        const bool live = tile < p.tstart[p.M];
        const int first = p.mstart[lo] + (tile - p.tstart[lo]) * kFT;
        s_m = lo; s_first = first; s_n = live ? min(kFT, p.mstart[lo + 1] - first) : 0;
    }
    __syncthreads();
    const int m = s_m, nf = s_n;
    if (nf <= 0) return;
    const int32_t* frames = p.perm + s_first;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int TI = D2p >> 2, ti = tid % TI, tf = tid / TI;     // thread tile: rows 4 ti.., frames 4 tf..
    const int i0 = 4 * ti, f0 = 4 * tf;
    const bool worker = tf < kFT / 4;
    // element walks e -> e + nth without divisions, four loads in flight per thread
    auto load_mat = [&](const double* src) {
        const double* sm_ = src + (size_t)m * D2 * D2;
        const int dk = nth / D2p, di = nth - dk * D2p;
        int k = tid / D2p, i = tid - k * D2p;
        for (int e = tid; e < D2 * D2p; e += 4 * nth) {
            double v[4];
            int kk[4], ii[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                kk[u] = k; ii[u] = i;
                v[u] = (e + u * nth < D2 * D2p && i < D2) ? sm_[i + (size_t)k * D2] : 0.0;
                k += dk; i += di;
                if (i >= D2p) { i -= D2p; ++k; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (e + u * nth < D2 * D2p) mat[(size_t)kk[u] * D2p + ii[u]] = v[u];
        }
    };
    // ---- panel
    {
        const int df = nth / D2, dk = nth - df * D2;
        int f = tid / D2, k = tid - f * D2;
        for (int e = tid; e < kFT * D2; e += 4 * nth) {
            double v[4];
            int ff[4], kk[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ff[u] = f; kk[u] = k;
                v[u] = 0.0;
                if (e + u * nth < kFT * D2 && f < nf) {
                    const int64_t t = frames[f];
                    if (MODE == 0) {
                        v[u] = p.X[t * p.ldx + k] - p.mux[(size_t)m * D2 + k];
                    } else {
                        const int Ds = D2 >> 1;
                        double wy;
                        if (k < Ds) {
                            wy = p.Y[t * p.ldy + k];
                        } else {
                            const unsigned char ed = p.edge[t];
                            wy = 0.0;
                            if (!(ed & 1)) wy = -0.5 * p.Y[(t - 1) * p.ldy + k - Ds];
                            if (!(ed & 2)) wy = fma(0.5, p.Y[(t + 1) * p.ldy + k - Ds], wy);
                        }
                        v[u] = p.E[t * D2 + k] - wy;
                    }
                }
                f += df; k += dk;
                if (k >= D2) { k -= D2; ++f; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (e + u * nth < kFT * D2) pan[(size_t)kk[u] * kFP + ff[u]] = v[u];
        }
    }
    load_mat(MODE == 0 ? p.A : p.P);
    __syncthreads();
    double o[4][4];
    if (MODE == 0) {
        // ---- E = muy + A panel
        if (worker) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double b = (i0 + r < D2) ? p.muy[(size_t)m * D2 + i0 + r] : 0.0;
#pragma unroll
                for (int c = 0; c < 4; ++c) o[r][c] = b;
            }
            panel_product(mat, D2p, pan, D2, i0, f0, o);
        }
        __syncthreads();     // the x - mux panel is dead: it becomes the E panel
        if (worker) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int f = f0 + c;
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (i0 + r < D2) mid[(size_t)(i0 + r) * kFP + f] = o[r][c];
                if (f < nf) {
                    const int64_t t = frames[f];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (i0 + r < D2) {
                            p.E[t * D2 + i0 + r] = o[r][c];
                            if (p.Eout) p.Eout[t * D2 + i0 + r] = o[r][c];
                        }
                }
            }
        }
        __syncthreads();
        load_mat(p.P);
        __syncthreads();
    }
    // ---- G = P (E panel | GV panel)
    if (worker) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) o[r][c] = 0.0;
        panel_product(mat, D2p, MODE == 0 ? mid : pan, D2, i0, f0, o);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int f = f0 + c;
            if (f < nf) {
                const int64_t t = frames[f];
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (i0 + r < D2) p.G[t * D2 + i0 + r] = o[r][c];
            }
        }
    }
}

// ---- the same products on the FP64 tensor path (D2 <= 96) ---------------------------------------
// One CTA = one 64-frame panel of one mixture, 8 warps; warp w owns the 8 frames [8w, 8w+8) and all
// NI = ceil(D2/8) row tiles of the output, i.e. NI accumulator tiles of mma.sync.m8n8k4.f64.  Per
// k-step a warp reads one B fragment (panel) and NI A fragments (matrix) from shared memory for NI DMMAs
// (the register-tiled DFMA version needed 4 LDS.128 per 16 DFMA and sat at 25 % of the FP64 pipe).
// Both matrices of the mixture arrive by cp.async while the panel is being gathered.  The E panel of
// a warp's frames is produced and consumed by that warp alone, so no CTA barrier separates the two GEMMs.
// Leading dimensions are = 8 (mod 16) doubles: the four k-columns of a fragment fall on distinct halves of
// the banks (two wavefronts per 64-bit fragment load, the minimum).
__device__ __forceinline__ void dmma_acc(double2& c, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c.x), "+d"(c.y) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16_g(void* smem, const void* gmem) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(gmem) : "memory");
}

template <int MODE, int NI>
__global__ void __launch_bounds__(256) group_panel_dmma_kernel(const GroupParams p) {
    constexpr int DP = 8 * NI, LDM = DP + 8, LDP = kFT + 8;
    extern __shared__ __align__(16) double sm[];
    double* matA = sm;                         // [DP][LDM]  A_m: column k at matA + k*LDM (MODE 0)
    double* matP = matA + (MODE == 0 ? DP * LDM : 0);      // P_m
    double* pan = matP + DP * LDM;             // [DP][LDP]  panel: row k = reduction index, column = frame
    __shared__ int s_m, s_first, s_n;
    const int D2 = p.D2;
    if (threadIdx.x == 0) {
        const int tile = blockIdx.x;
        int lo = 0, hi = p.M;                  // largest m with tstart[m] <= tile
        while (hi - lo > 1) { const int mid_m = (lo + hi) >> 1; if (p.tstart[mid_m] <= tile) lo = mid_m; else hi = mid_m; }
        const bool live = tile < p.tstart[p.M];
        const int first = p.mstart[lo] + (tile - p.tstart[lo]) * kFT;
        s_m = lo; s_first = first; s_n = live ? min(kFT, p.mstart[lo + 1] - first) : 0;
    }
    __syncthreads();
    const int m = s_m, nf = s_n;
    if (nf <= 0) return;
    const int32_t* frames = p.perm + s_first;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, r = lane >> 2, q = lane & 3;

    // ---- matrices: cp.async when every column is a whole number of 16-byte pieces, plain loads otherwise
    auto stage = [&](double* dst, const double* src) {
        const double* sm_ = src + (size_t)m * D2 * D2;
        if ((D2 & 1) == 0) {
            const int pieces = D2 / 2;
            for (int e = tid; e < D2 * pieces; e += 256) {
                const int k = e / pieces, i2 = e - k * pieces;
                cp_async16_g(dst + (size_t)k * LDM + 2 * i2, sm_ + (size_t)k * D2 + 2 * i2);
            }
        } else {
            for (int e = tid; e < D2 * D2; e += 256) {
                const int k = e / D2, i = e - k * D2;
                dst[(size_t)k * LDM + i] = sm_[e];
            }
        }
        // zero the padding rows / columns the fragments touch
        for (int e = tid; e < DP * (DP - D2); e += 256) {
            const int k = e / (DP - D2), i = D2 + e - k * (DP - D2);
            dst[(size_t)k * LDM + i] = 0.0;
        }
        for (int e = tid; e < (DP - D2) * LDM; e += 256) dst[(size_t)D2 * LDM + e] = 0.0;
    };
    if (MODE == 0) stage(matA, p.A);
    stage(matP, p.P);
    asm volatile("cp.async.commit_group;" ::: "memory");

    // ---- panel: every warp gathers its own 8 frames (lane: frame c = lane/4, reduction indices k = q, q+4, ..)
    {
        const int f = 8 * w + r;
        const bool live = f < nf;
        const int64_t t = live ? frames[f] : 0;
        if (MODE == 0) {
            // all loads of the lane first (2 NI independent requests in flight), then the subtractions
            double xv[2 * NI], mv[2 * NI];
#pragma unroll
            for (int j = 0; j < 2 * NI; ++j) {
                const int k = q + 4 * j;
                const bool ok = live && k < D2;
                xv[j] = ok ? p.X[t * p.ldx + k] : 0.0;
                mv[j] = ok ? p.mux[(size_t)m * D2 + k] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 2 * NI; ++j) pan[(size_t)(q + 4 * j) * LDP + f] = xv[j] - mv[j];
        } else {
            for (int k = q; k < DP; k += 4) {
                double v = 0.0;
                if (live && k < D2) {
                    const int Ds = D2 >> 1;
                    double wy;
                    if (k < Ds) {
                        wy = p.Y[t * p.ldy + k];
                    } else {
                        const unsigned char ed = p.edge[t];
                        wy = 0.0;
                        if (!(ed & 1)) wy = -0.5 * p.Y[(t - 1) * p.ldy + k - Ds];
                        if (!(ed & 2)) wy = fma(0.5, p.Y[(t + 1) * p.ldy + k - Ds], wy);
                    }
                    v = p.E[t * D2 + k] - wy;
                }
                pan[(size_t)k * LDP + f] = v;
            }
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    // C[8i + r][8w + 2q + {0,1}] += sum_k Mat[8i + r][k] * Pan[k][8w + ..]:  a = Mat[8i + r][k0 + q],  b = Pan[k0 + q][8w + r]
    auto product = [&](const double* mat, double2 (&acc)[NI]) {
        const double* ap = mat + (size_t)q * LDM + r;
        const double* bp = pan + (size_t)q * LDP + 8 * w + r;
#pragma unroll 2
        for (int k0 = 0; k0 < DP; k0 += 4) {
            const double b = bp[(size_t)k0 * LDP];
#pragma unroll
            for (int i = 0; i < NI; ++i) dmma_acc(acc[i], ap[(size_t)k0 * LDM + 8 * i], b);
        }
    };
    const int f0 = 8 * w + 2 * q;
    const int64_t t0 = (f0 < nf) ? frames[f0] : -1, t1 = (f0 + 1 < nf) ? frames[f0 + 1] : -1;
    double2 acc[NI];
    if (MODE == 0) {
        // ---- E = muy + A panel
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const double b = (8 * i + r < D2) ? p.muy[(size_t)m * D2 + 8 * i + r] : 0.0;
            acc[i] = make_double2(b, b);
        }
        product(matA, acc);
        __syncwarp();            // every lane of the warp has read the x - mux columns of its frames
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int row = 8 * i + r;
            *reinterpret_cast<double2*>(pan + (size_t)row * LDP + f0) = acc[i];      // the E panel of this warp's frames
            if (row < D2) {
                if (t0 >= 0) { p.E[t0 * D2 + row] = acc[i].x; if (p.Eout) p.Eout[t0 * D2 + row] = acc[i].x; }
                if (t1 >= 0) { p.E[t1 * D2 + row] = acc[i].y; if (p.Eout) p.Eout[t1 * D2 + row] = acc[i].y; }
            }
        }
        __syncwarp();
    }
    // ---- G = P (E panel | GV panel)
#pragma unroll
    for (int i = 0; i < NI; ++i) acc[i] = make_double2(0.0, 0.0);
    product(matP, acc);
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int row = 8 * i + r;
        if (row < D2) {
            if (t0 >= 0) p.G[t0 * D2 + row] = acc[i].x;
            if (t1 >= 0) p.G[t1 * D2 + row] = acc[i].y;
        }
    }
}

template <int MODE>
static int32_t launch_group_dmma(const GroupParams& p, int64_t npanels, cudaStream_t st) {
    const int NI = (p.D2 + 7) / 8, DP = 8 * NI;
    const size_t smem = ((size_t)(MODE == 0 ? 2 : 1) * DP * (DP + 8) + (size_t)DP * (kFT + 8)) * sizeof(double);
#define VCB_GROUP_CASE(N)                                                                                             \
    case N:                                                                                                           \
        VCB_CUDA(cudaFuncSetAttribute(group_panel_dmma_kernel<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        group_panel_dmma_kernel<MODE, N><<<(unsigned)npanels, 256, smem, st>>>(p);                                    \
        break;
    switch (NI) {
        VCB_GROUP_CASE(1) VCB_GROUP_CASE(2) VCB_GROUP_CASE(3) VCB_GROUP_CASE(4) VCB_GROUP_CASE(5) VCB_GROUP_CASE(6)
        VCB_GROUP_CASE(7) VCB_GROUP_CASE(8) VCB_GROUP_CASE(9) VCB_GROUP_CASE(10) VCB_GROUP_CASE(11) VCB_GROUP_CASE(12)
        default: return fail(VCB_EUNSUPPORTED, "grouped DMMA product covers dimensions <= 96 (got %d)", p.D2);
    }
#undef VCB_GROUP_CASE
    count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

// ---- exact arg-max for the frames flagged as near ties, as panels -------------------------------
// The flagged frames (a few per thousand) need c_m - 1/2 |Linv_m (x - mux_m)|^2 in Float64 for EVERY
// mixture.  A block per frame streamed 64 x 18 KB of Linv from L2 per frame (L2-bandwidth bound,
// 0.32 ms at C2); here a task is (64-frame panel of the flag list, mixture): Linv_m is staged once per
// 64 frames.  Persistent grid, the flag count is read on the device (no host synchronisation).
__global__ void recheck_panel_kernel(const double* __restrict__ X, int64_t ldx, const int* __restrict__ flag_count,
                                     const int64_t* __restrict__ flag_list, const double* __restrict__ linv_cm,
                                     const double* __restrict__ mux, const double* __restrict__ c, int D, int M,
                                     double* __restrict__ lout, int cap) {
    extern __shared__ __align__(16) double sm[];
    const int Dp = (D + 3) & ~3;
    double* mat = sm;                         // [D][Dp]  column k of Linv_m at mat + k*Dp
    double* pan = mat + (size_t)D * Dp;       // [D][kFP]
    double* qs = pan + (size_t)D * kFP;       // [TI][kFT] partial |z|^2 per row group (summed in a fixed order)
    const int n = min(*flag_count, cap);
    const int npanel = (n + kFT - 1) / kFT;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int TI = Dp >> 2, ti = tid % TI, tf = tid / TI, i0 = 4 * ti, f0 = 4 * tf;
    const bool worker = tf < kFT / 4;
    for (int64_t task = blockIdx.x; task < (int64_t)npanel * M; task += gridDim.x) {
        const int pnl = (int)(task / M), m = (int)(task - (int64_t)pnl * M);
        const int first = pnl * kFT, nf = min(kFT, n - first);
        __syncthreads();
        for (int e = tid; e < D * Dp; e += nth) {
            const int k = e / Dp, i = e - k * Dp;
            mat[e] = (i < D) ? linv_cm[(size_t)m * D * D + (size_t)k * D + i] : 0.0;
        }
        for (int e = tid; e < kFT * D; e += nth) {
            const int f = e / D, k = e - f * D;
            pan[(size_t)k * kFP + f] = (f < nf) ? X[flag_list[first + f] * ldx + k] - mux[(size_t)m * D + k] : 0.0;
        }
        __syncthreads();
        if (worker) {
            double o[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) o[r][cc] = 0.0;
            panel_product(mat, Dp, pan, D, i0, f0, o);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                double q = 0.0;
#pragma unroll
                for (int r = 0; r < 4; ++r) q = fma(o[r][cc], o[r][cc], q);
                qs[ti * kFT + f0 + cc] = q;
            }
        }
        __syncthreads();
        if (tid < nf) {
            double q = 0.0;
            for (int g = 0; g < TI; ++g) q += qs[g * kFT + tid];
            lout[(size_t)(first + tid) * M + m] = c[m] - 0.5 * q;
        }
    }
}

// first maximum over the mixtures (indmax, src/gmm.jl:46); one warp per flagged frame
__global__ void recheck_pick_kernel(const int* __restrict__ flag_count, const int64_t* __restrict__ flag_list,
                                    const double* __restrict__ lout, int M, int cap, int32_t* __restrict__ mhat) {
    const int n = min(*flag_count, cap);
    const int lane = threadIdx.x & 31;
    for (int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < n; e += gridDim.x * (blockDim.x >> 5)) {
        double bv = -INFINITY;
        int bi = 0x7FFFFFFF;
        for (int m = lane; m < M; m += 32) {
            const double l = lout[(size_t)e * M + m];
            if (l > bv) { bv = l; bi = m; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xFFFFFFFFu, bv, o);
            const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0 && bi != 0x7FFFFFFF) mhat[flag_list[e]] = bi;
    }
}

}  // namespace

size_t group_workspace_ints(int M, int64_t total) { return (size_t)total + 4 * (size_t)(M + 1); }

// Buckets the frames by mixture.  ws: group_workspace_ints() ints.  Returns the number of panels in
// *npanels_bound (an upper bound that needs no device read-back: total/kFT + M).
int32_t group_frames_by_mixture(const int32_t* d_mhat, int64_t total, int M, int* ws, int64_t* npanels_bound,
                                cudaStream_t st) {
    int* hist = ws;
    int* mstart = ws + (M + 1);
    int* tstart = ws + 2 * (M + 1);
    int* cursor = ws + 3 * (M + 1);
    int32_t* perm = ws + 4 * (M + 1);
    VCB_CUDA(cudaMemsetAsync(hist, 0, (size_t)(M + 1) * sizeof(int), st));
    const int64_t per = (int64_t)kHistBlock * kHistPer;
    const unsigned g = (unsigned)((total + per - 1) / per);
    if ((size_t)M * 2 * sizeof(int) > 48 * 1024) return fail(VCB_EUNSUPPORTED, "too many mixtures (%d) for the bucket kernels", M);
    group_hist_kernel<<<g, kHistBlock, (size_t)M * sizeof(int), st>>>(d_mhat, total, M, hist);
    group_scan_kernel<<<1, 32, 0, st>>>(hist, M, mstart, tstart, cursor);
    group_scatter_kernel<<<g, kHistBlock, (size_t)2 * M * sizeof(int), st>>>(d_mhat, total, M, cursor, perm);
    count_launch(); count_launch(); count_launch();
    VCB_CUDA(cudaGetLastError());
    *npanels_bound = total / kFT + M;
    return VCB_OK;
}

static int32_t launch_group(int mode, GroupParams& p, const int* ws, int M, int64_t npanels, cudaStream_t st) {
    p.mstart = ws + (M + 1);
    p.tstart = ws + 2 * (M + 1);
    p.perm = ws + 4 * (M + 1);
    p.M = M;
    // FP64 tensor-core panels for D2 <= 96 (VCB_GROUP=simt keeps the register-tiled DFMA kernel)
    static const bool simt = [] { const char* e = getenv("VCB_GROUP"); return e && e[0] == 's'; }();
    if (p.D2 <= 96 && !simt) return mode == 0 ? launch_group_dmma<0>(p, npanels, st) : launch_group_dmma<1>(p, npanels, st);
    const int D2p = (p.D2 + 3) & ~3, TI = D2p / 4;
    const int threads = std::min(1024, round_up(TI * (kFT / 4), 32));
    if (TI * (kFT / 4) > 1024) return fail(VCB_EUNSUPPORTED, "feature dimension %d too large for the grouped product", p.D2);
    const size_t smem = ((size_t)p.D2 * D2p + (size_t)p.D2 * kFP) * sizeof(double);
    if (mode == 0) {
        VCB_CUDA(cudaFuncSetAttribute(group_panel_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        group_panel_kernel<0><<<(unsigned)npanels, threads, smem, st>>>(p);
    } else {
        VCB_CUDA(cudaFuncSetAttribute(group_panel_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        group_panel_kernel<1><<<(unsigned)npanels, threads, smem, st>>>(p);
    }
    count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

int32_t group_e_step(const vcb_traj& tr, const int* ws, int64_t npanels, const double* dX, int64_t ldx, double* dE,
                     double* dEout, double* dG, cudaStream_t st) {
    const vcb_gmmmap& g = *tr.g;
    GroupParams p{};
    p.A = g.d_A.p; p.mux = g.d_mux.p; p.muy = g.d_muy.p; p.P = tr.d_P.p; p.D2 = g.D;
    p.X = dX; p.ldx = ldx; p.E = dE; p.Eout = dEout; p.G = dG;
    return launch_group(0, p, ws, g.M, npanels, st);
}

int32_t group_gv_step(const vcb_traj& tr, const int* ws, int64_t npanels, const double* dY, int64_t ldy,
                      const double* dE, const unsigned char* d_edge, double* dH, cudaStream_t st) {
    const vcb_gmmmap& g = *tr.g;
    GroupParams p{};
    p.P = tr.d_P.p; p.D2 = g.D; p.E = const_cast<double*>(dE); p.G = dH;
    p.Y = dY; p.ldy = ldy; p.edge = d_edge;
    return launch_group(1, p, ws, g.M, npanels, st);
}

}  // namespace vcb

namespace vcb {

// Panel re-check of the first `cap` flagged frames; returns the scratch it used through *scratch_out so the
// caller can free it on the stream.  Frames beyond `cap` are left to the per-frame kernel (start = cap).
int32_t recheck_argmax_panels(const vcb_gmmmap& g, const double* dX, int64_t ldx, const int* d_flag_count,
                              const int64_t* d_flag_list, int32_t* d_mhat, int cap, double* d_lout, cudaStream_t st) {
    const int D = g.D, Dp = (D + 3) & ~3, TI = Dp / 4;
    const int threads = std::min(1024, round_up(TI * (kFT / 4), 32));
    const size_t smem = ((size_t)D * Dp + (size_t)D * kFP + (size_t)TI * kFT) * sizeof(double);
    VCB_CUDA(cudaFuncSetAttribute(recheck_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    recheck_panel_kernel<<<148 * 3, threads, smem, st>>>(dX, ldx, d_flag_count, d_flag_list, g.d_linv_cm.p, g.d_mux.p,
                                                        g.d_c.p, D, g.M, d_lout, cap);
    recheck_pick_kernel<<<148, 256, 0, st>>>(d_flag_count, d_flag_list, d_lout, g.M, cap, d_mhat);
    count_launch(); count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

}  // namespace vcb
