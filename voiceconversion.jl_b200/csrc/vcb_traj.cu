// vcb_traj.cu -- K3: maximum-likelihood trajectory solve (Toda 2007), Float64.
//
// Replaces fvconvert(tgmm, X) steps (iv)-(vii) (reference src/trajectory_gmmmap.jl:85-109):
//   E_t = muy_m + A_m (x_t - mux_m)  for the arg-max mixture m = mhat_t          (:85-89)
//   D^-1 = blkdiag(Dy[:,:,mhat_t])                                               (:95)
//   y = (W' D^-1 W) \ (W' D^-1 E)                                                (:103-105)
// W (constructW, :39-61) and the block-diagonal D^-1 are never materialised.  With
// P_t = Dy[:,:,mhat_t] = [Pss Psd; Pds Pdd] (Ds x Ds blocks) the normal matrix R = W' D^-1 W is
// block-pentadiagonal with dense Ds x Ds blocks:
//   R[t][t]   = Pss_t + 1/4 Pdd_{t-1} + 1/4 Pdd_{t+1}
//   R[t][t-1] = -1/2 Psd_t + 1/2 Pds_{t-1}
//   R[t][t-2] = -1/4 Pdd_{t-1}
// (terms that reach outside 0..T-1 vanish), and with g_t = P_t E_t = [u_t; v_t]:
//   r_t = u_t + 1/2 v_{t-1} - 1/2 v_{t+1}.
// One CTA factorises one chunk by block Cholesky, sequential in time and parallel inside the
// Ds x Ds block operations, holding the three-block-row window in shared memory; per frame it
// streams L[t][t]^-1, L[t][t-1], L[t][t-2] (3 Ds^2 doubles) to HBM once and reads them once in the
// back substitution (cp.async double-buffered).
#include <cstdlib>

#include "vcb_kernels.h"

namespace vcb {

namespace {

// ---- E_t and g_t = P_t E_t ---------------------------------------------------------------------
// One block handles FR frames; 64.. threads per frame, thread i owns output row i.
__global__ void traj_e_kernel(const double* __restrict__ X, int64_t ldx, int64_t total,
                              const int32_t* __restrict__ mhat, const double* __restrict__ A,
                              const double* __restrict__ mux, const double* __restrict__ muy,
                              const double* __restrict__ P, int D2, double* __restrict__ E,
                              double* __restrict__ Gv, double* __restrict__ Eout) {
    extern __shared__ double sm[];  // per frame: dx[D2], e[D2]
    const int tpf = blockDim.x;     // threads per frame (>= D2)
    const int f = threadIdx.y;
    const int64_t t = (int64_t)blockIdx.x * blockDim.y + f;
    double* dx = sm + (size_t)f * 2 * D2;
    double* ev = dx + D2;
    const bool live = t < total;
    const int m = live ? mhat[t] : 0;
    const int i = threadIdx.x;
    if (live && i < D2) dx[i] = X[t * ldx + i] - mux[(size_t)m * D2 + i];
    __syncthreads();
    if (live && i < D2) {
        const double* a = A + (size_t)m * D2 * D2 + i;  // A[i + k*D2]
        double s = 0.0;
        for (int k = 0; k < D2; ++k) s = fma(a[(size_t)k * D2], dx[k], s);
        s += muy[(size_t)m * D2 + i];
        ev[i] = s;
        E[t * D2 + i] = s;
        if (Eout) Eout[t * D2 + i] = s;
    }
    __syncthreads();
    if (live && i < D2) {
        const double* pm = P + (size_t)m * D2 * D2 + i;  // symmetric: row i == column i
        double s = 0.0;
        for (int k = 0; k < D2; ++k) s = fma(pm[(size_t)k * D2], ev[k], s);
        Gv[t * D2 + i] = s;
    }
    (void)tpf;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

struct TrajParams {
    const double* P;        // [M][D2*D2] symmetric precision blocks
    const int32_t* mhat;    // [total] 0-based
    const double* Gv;       // [total][D2]  g_t = P_t E_t
    const int64_t* chunk_off;
    double* Lst;            // [total][3][Ds*Ds]  transposes of Linv_tt, L[t][t-1], L[t][t-2] (compact)
    double* Z;              // [total][Ds]
    double* Y; int64_t ldy;
    const double* Xpow; int64_t ldx; int copy_power;
    int Ds;
    int* err;
};

// Shared-memory block (row-major, leading dimension LD = Ds|1 to spread banks).
#define BLK(b, i, j) (b)[(i) * LD + (j)]

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
traj_solve_kernel(const TrajParams p) {
    const int Ds = p.Ds, D2 = 2 * Ds, LD = Ds | 1, BB = Ds * Ds;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    const int64_t c0 = p.chunk_off[blockIdx.x];
    const int T = (int)(p.chunk_off[blockIdx.x + 1] - c0);
    if (T <= 0) return;

    extern __shared__ __align__(16) double sm[];
    const int BS = Ds * LD;  // doubles per block buffer
    double* Rtt = sm;
    double* Rt1 = Rtt + BS;
    double* Rt2 = Rt1 + BS;
    double* G2 = Rt2 + BS;
    double* Tm = G2 + BS;
    double* S = Tm + BS;
    double* gbuf[2] = {S + BS, S + 2 * BS};                    // G1 / L[t-1][t-2]
    double* ibuf[3] = {S + 3 * BS, S + 4 * BS, S + 5 * BS};    // Linv_t, Linv_{t-1}, Linv_{t-2}
    double* vec = S + 6 * BS;  // r[Ds], tmp[Ds], z0[Ds], z1[Ds], z2[Ds]
    double* rv = vec;
    double* tmpv = vec + Ds;
    double* zb[3] = {vec + 2 * Ds, vec + 3 * Ds, vec + 4 * Ds};

    const int32_t* mh = p.mhat + c0;
    const double* gv = p.Gv + c0 * D2;
    double* Lst = p.Lst + c0 * 3 * BB;
    double* Zg = p.Z + c0 * Ds;

    // =========================== forward: block Cholesky + L z = r ===========================
    for (int t = 0; t < T; ++t) {
        double* G1 = gbuf[t & 1];
        double* Lt1t2 = gbuf[(t & 1) ^ 1];
        double* W = ibuf[t % 3];
        double* Lm1inv = ibuf[(t + 2) % 3];
        double* Lm2inv = ibuf[(t + 1) % 3];
        double* zt = zb[t % 3];
        const double* z1 = zb[(t + 2) % 3];
        const double* z2 = zb[(t + 1) % 3];

        // ---- 0. assemble R[t][t], R[t][t-1], R[t][t-2] and r_t
        {
            const double* Pt = p.P + (size_t)mh[t] * D2 * D2;
            const double* Pm = (t >= 1) ? p.P + (size_t)mh[t - 1] * D2 * D2 : nullptr;
            const double* Pp = (t + 1 < T) ? p.P + (size_t)mh[t + 1] * D2 * D2 : nullptr;
            for (int e = tid; e < BB; e += NT) {
                // e = j*Ds + i so that consecutive threads read consecutive rows of a column
                const int j = e / Ds, i = e - j * Ds;
                double rtt = Pt[i + (size_t)j * D2];
                double rt1 = 0.0, rt2 = 0.0;
                if (Pm) {
                    const double pdd = Pm[(Ds + i) + (size_t)(Ds + j) * D2];
                    rtt = fma(0.25, pdd, rtt);
                    rt1 = 0.5 * Pm[(Ds + i) + (size_t)j * D2] - 0.5 * Pt[i + (size_t)(Ds + j) * D2];
                    if (t >= 2) rt2 = -0.25 * pdd;
                }
                if (Pp) rtt = fma(0.25, Pp[(Ds + i) + (size_t)(Ds + j) * D2], rtt);
                BLK(Rtt, i, j) = rtt;
                BLK(Rt1, i, j) = rt1;
                BLK(Rt2, i, j) = rt2;
            }
            if (tid < Ds) {
                double r = gv[(size_t)t * D2 + tid];
                if (t >= 1) r = fma(0.5, gv[(size_t)(t - 1) * D2 + Ds + tid], r);
                if (t + 1 < T) r = fma(-0.5, gv[(size_t)(t + 1) * D2 + Ds + tid], r);
                rv[tid] = r;
            }
        }
        __syncthreads();
        // ---- 1. G2 = L[t][t-2] = R[t][t-2] * Linv_{t-2}'   (Linv lower: k <= j)
        for (int e = tid; e < BB; e += NT) {
            const int i = e / Ds, j = e - i * Ds;
            double s = 0.0;
            if (t >= 2)
                for (int k = 0; k <= j; ++k) s = fma(BLK(Rt2, i, k), BLK(Lm2inv, j, k), s);
            BLK(G2, i, j) = s;
        }
        __syncthreads();
        // ---- 2. Tm = R[t][t-1] - G2 * L[t-1][t-2]'
        for (int e = tid; e < BB; e += NT) {
            const int i = e / Ds, j = e - i * Ds;
            double s = BLK(Rt1, i, j);
            if (t >= 2)
                for (int k = 0; k < Ds; ++k) s = fma(-BLK(G2, i, k), BLK(Lt1t2, j, k), s);
            BLK(Tm, i, j) = s;
        }
        __syncthreads();
        // ---- 3. G1 = L[t][t-1] = Tm * Linv_{t-1}'
        for (int e = tid; e < BB; e += NT) {
            const int i = e / Ds, j = e - i * Ds;
            double s = 0.0;
            if (t >= 1)
                for (int k = 0; k <= j; ++k) s = fma(BLK(Tm, i, k), BLK(Lm1inv, j, k), s);
            BLK(G1, i, j) = s;
        }
        __syncthreads();
        // ---- 4. S = R[t][t] - G2 G2' - G1 G1'  (lower triangle), W = I
        for (int e = tid; e < BB; e += NT) {
            const int i = e / Ds, j = e - i * Ds;
            if (j <= i) {
                double s = BLK(Rtt, i, j);
                if (t >= 1)
                    for (int k = 0; k < Ds; ++k) s = fma(-BLK(G1, i, k), BLK(G1, j, k), s);
                if (t >= 2)
                    for (int k = 0; k < Ds; ++k) s = fma(-BLK(G2, i, k), BLK(G2, j, k), s);
                BLK(S, i, j) = s;
            } else {
                BLK(S, i, j) = 0.0;
            }
            BLK(W, i, j) = (i == j) ? 1.0 : 0.0;
        }
        __syncthreads();
        // ---- 5. right-looking Cholesky of S fused with W <- L^-1 (forward substitution on I).
        //         The diagonal of L is never written back (only L^-1, G1, G2 are kept), so two
        //         barriers per column suffice.
        for (int k = 0; k < Ds; ++k) {
            const double skk = BLK(S, k, k);
            if (!(skk > 0.0) && tid == 0) atomicExch(p.err, 1);
            const double dinv = 1.0 / sqrt(skk);
            for (int e = tid; e < 2 * Ds; e += NT) {
                if (e < Ds) {              // column k of L (below the diagonal)
                    if (e > k) BLK(S, e, k) *= dinv;
                } else {                   // row k of L^-1
                    const int j = e - Ds;
                    if (j <= k) BLK(W, k, j) *= dinv;
                }
            }
            __syncthreads();
            for (int e = tid; e < BB; e += NT) {
                const int i = e / Ds, j = e - i * Ds;
                if (i > k) {
                    const double lik = BLK(S, i, k);
                    if (j > k && j <= i) BLK(S, i, j) = fma(-lik, BLK(S, j, k), BLK(S, i, j));
                    else if (j <= k) BLK(W, i, j) = fma(-lik, BLK(W, k, j), BLK(W, i, j));
                }
            }
            __syncthreads();
        }
        // ---- 6. z_t = Linv (r - G1 z_{t-1} - G2 z_{t-2})
        for (int j = warp; j < Ds; j += nwarps) {
            double s = 0.0;
            for (int k = lane; k < Ds; k += 32) {
                if (t >= 1) s = fma(BLK(G1, j, k), z1[k], s);
                if (t >= 2) s = fma(BLK(G2, j, k), z2[k], s);
            }
            s = warp_sum(s);
            if (lane == 0) tmpv[j] = rv[j] - s;
        }
        __syncthreads();
        for (int i = warp; i < Ds; i += nwarps) {
            double s = 0.0;
            for (int j = lane; j <= i; j += 32) s = fma(BLK(W, i, j), tmpv[j], s);
            s = warp_sum(s);
            if (lane == 0) { zt[i] = s; Zg[(size_t)t * Ds + i] = s; }
        }
        // ---- 7. stream the three blocks of block-row t to HBM, TRANSPOSED (the back
        //         substitution multiplies by L', so it reads rows)
        {
            double* dst = Lst + (size_t)t * 3 * BB;
            for (int e = tid; e < BB; e += NT) {
                const int j = e / Ds, i = e - j * Ds;   // dst[j][i] = src[i][j]
                dst[e] = BLK(W, i, j);
                dst[BB + e] = BLK(G1, i, j);
                dst[2 * BB + e] = BLK(G2, i, j);
            }
        }
        __syncthreads();
    }

    // =========================== backward: L' y = z ===========================================
    // y_t = Linv_t' (z_t - L[t+1][t]' y_{t+1} - L[t+2][t]' y_{t+2})
    // Needs Linv_t (slot 0 of row t), L[t+1][t] (slot 1 of row t+1), L[t+2][t] (slot 2 of row t+2).
    double* bb[2][3] = {{Rtt, Rt1, Rt2}, {G2, Tm, S}};  // compact Ds*Ds images, double-buffered
    double* yb[3] = {zb[0], zb[1], zb[2]};
    auto prefetch = [&](int t, int buf) {
        const int n16 = BB / 2;  // 16-byte pieces per block
        for (int e = tid; e < 3 * n16; e += NT) {
            const int b = e / n16, o = e - b * n16;
            const int row = t + b;
            if (row < T) cp_async16(bb[buf][b] + 2 * o, Lst + ((size_t)row * 3 + b) * BB + 2 * o);
        }
    };
    const bool vec_ok = (BB % 2) == 0;
    if (vec_ok) { prefetch(T - 1, (T - 1) & 1); cp_async_commit(); }
    for (int t = T - 1; t >= 0; --t) {
        const int buf = t & 1;
        double* yt = yb[t % 3];
        const double* y1 = yb[(t + 1) % 3];
        const double* y2 = yb[(t + 2) % 3];
        if (vec_ok) {
            if (t >= 1) prefetch(t - 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            for (int e = tid; e < 3 * BB; e += NT) {
                const int b = e / BB, o = e - b * BB;
                if (t + b < T) bb[buf][b][o] = Lst[((size_t)(t + b) * 3 + b) * BB + o];
            }
        }
        if (tid < Ds) tmpv[tid] = Zg[(size_t)t * Ds + tid];
        __syncthreads();
        const double* Li = bb[buf][0];
        const double* L1 = bb[buf][1];
        const double* L2 = bb[buf][2];
        for (int j = warp; j < Ds; j += nwarps) {
            double s = 0.0;
            for (int i = lane; i < Ds; i += 32) {
                if (t + 1 < T) s = fma(L1[j * Ds + i], y1[i], s);
                if (t + 2 < T) s = fma(L2[j * Ds + i], y2[i], s);
            }
            s = warp_sum(s);
            if (lane == 0) rv[j] = tmpv[j] - s;
        }
        __syncthreads();
        for (int j = warp; j < Ds; j += nwarps) {
            double s = 0.0;
            for (int i = j + lane; i < Ds; i += 32) s = fma(Li[j * Ds + i], rv[i], s);
            s = warp_sum(s);
            if (lane == 0) {
                yt[j] = s;
                p.Y[(c0 + t) * p.ldy + j] = s;  // reshape(y, D, T)  src/trajectory_gmmmap.jl:109
            }
        }
        if (p.copy_power && tid == 0) p.Y[(c0 + t) * p.ldy - 1] = p.Xpow[(c0 + t) * p.ldx - 1];  // src/common.jl:60
        __syncthreads();
    }
}

}  // namespace

int32_t traj_solve_device(const vcb_traj& tr, const double* dX, int64_t ldx, const int32_t* d_mhat,
                          const int64_t* d_chunk_off, int64_t nchunks, int max_chunk_len,
                          int64_t total, double* dY, int64_t ldy, double* dEy_out, bool copy_power,
                          cudaStream_t st) {
    (void)max_chunk_len;
    if (total == 0 || nchunks == 0) return VCB_OK;
    const vcb_gmmmap& g = *tr.g;
    const int Ds = tr.Ds, D2 = 2 * Ds, BB = Ds * Ds, LD = Ds | 1;
    const size_t smem = ((size_t)12 * Ds * LD + 5 * Ds) * sizeof(double);
    if (smem > 227 * 1024) return fail(VCB_EUNSUPPORTED, "static dimension %d too large for the trajectory solver (shared memory)", Ds);

    double *dE = nullptr, *dG = nullptr, *dL = nullptr, *dZ = nullptr;
    int* derr = nullptr;
    VCB_CUDA(cudaMallocAsync((void**)&dE, (size_t)total * D2 * sizeof(double), st));
    VCB_CUDA(cudaMallocAsync((void**)&dG, (size_t)total * D2 * sizeof(double), st));
    VCB_CUDA(cudaMallocAsync((void**)&dL, (size_t)total * 3 * BB * sizeof(double), st));
    VCB_CUDA(cudaMallocAsync((void**)&dZ, (size_t)total * Ds * sizeof(double), st));
    VCB_CUDA(cudaMallocAsync((void**)&derr, sizeof(int), st));
    VCB_CUDA(cudaMemsetAsync(derr, 0, sizeof(int), st));
    {
        const int tpf = round_up(D2, 32), fpb = std::max(1, 256 / tpf);
        dim3 block(tpf, fpb), grid((unsigned)((total + fpb - 1) / fpb));
        traj_e_kernel<<<grid, block, (size_t)fpb * 2 * D2 * sizeof(double), st>>>(
            dX, ldx, total, d_mhat, g.d_A.p, g.d_mux.p, g.d_muy.p, tr.d_P.p, D2, dE, dG, dEy_out);
        count_launch();
        VCB_CUDA(cudaGetLastError());
    }
    {
        TrajParams p{};
        p.P = tr.d_P.p; p.mhat = d_mhat; p.Gv = dG; p.chunk_off = d_chunk_off; p.Lst = dL; p.Z = dZ;
        p.Y = dY; p.ldy = ldy; p.Xpow = dX; p.ldx = ldx; p.copy_power = copy_power ? 1 : 0;
        p.Ds = Ds; p.err = derr;
        int nt = round_up(BB, 32);
        if (nt > 1024) nt = round_up((BB + 1) / 2, 32);
        if (nt > 1024) nt = 1024;
        // Resident CTAs per SM trade registers (spills) for latency hiding; VCB_TRAJ_MINB=1|2|3
        // overrides the default for tuning runs.
        static const int minb = [] { const char* e = getenv("VCB_TRAJ_MINB"); return e ? atoi(e) : 3; }();
        void (*k)(const TrajParams) = traj_solve_kernel<1024, 1>;
        if (nt <= 640) k = minb >= 3 ? traj_solve_kernel<640, 3> : (minb == 2 ? traj_solve_kernel<640, 2> : traj_solve_kernel<640, 1>);
        VCB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<(unsigned)nchunks, nt, smem, st>>>(p);
        count_launch();
        VCB_CUDA(cudaGetLastError());
    }
    cudaFreeAsync(dE, st);
    cudaFreeAsync(dG, st);
    cudaFreeAsync(dL, st);
    cudaFreeAsync(dZ, st);
    cudaFreeAsync(derr, st);
    return VCB_OK;
}

}  // namespace vcb
