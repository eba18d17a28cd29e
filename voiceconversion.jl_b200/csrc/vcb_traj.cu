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
// One CTA of 64 threads factorises one chunk by block Cholesky, sequential in time.  The threads
// form an 8 x 8 grid of TS x TS register tiles (Ds <= 8*TS), so the Ds x Ds block products run
// register-blocked out of column-major shared-memory operands (the FP64 pipe, not shared-memory
// bandwidth, is the limiter), and the Cholesky of the diagonal block keeps its S and L^-1 tiles in
// registers, broadcasting one column / row per elimination step.  The three-block-row window lives
// in shared memory; per frame L[t][t]^-1, L[t][t-1], L[t][t-2] (3 Ds^2 doubles) are streamed to
// HBM once and read once by the cp.async double-buffered back substitution.  Several CTAs share
// an SM so one chunk's latency-bound Cholesky overlaps another chunk's products.
#include <cstdlib>

#include "vcb_kernels.h"
#include "vcb_traj.h"

namespace vcb {

namespace {

// ---- E_t and g_t = P_t E_t ---------------------------------------------------------------------
// One block handles FR frames; 64.. threads per frame, thread i owns output row i.
__global__ void traj_e_kernel(const double* __restrict__ X, int64_t ldx, int64_t total,
                              const int32_t* __restrict__ mhat, const double* __restrict__ A,
                              const double* __restrict__ mux, const double* __restrict__ muy,
                              const double* __restrict__ P, int D2, int groups, double* __restrict__ E,
                              double* __restrict__ Gv, double* __restrict__ Eout) {
    extern __shared__ double sm[];  // per frame: dx[D2], e[D2]
    const int f = threadIdx.y;
    double* dx = sm + (size_t)f * 2 * D2;
    double* ev = dx + D2;
    const int i = threadIdx.x;
    // a block walks `groups` consecutive groups of blockDim.y frames: neighbouring frames mostly
    // share their mixture, so A_m and P_m are re-read from L1 instead of L2
    for (int gi = 0; gi < groups; ++gi) {
        const int64_t t = ((int64_t)blockIdx.x * groups + gi) * blockDim.y + f;
        const bool live = t < total;
        const int m = live ? mhat[t] : 0;
        __syncthreads();
        if (live && i < D2) dx[i] = X[t * ldx + i] - mux[(size_t)m * D2 + i];
        __syncthreads();
        if (live && i < D2) {
            const double* a = A + (size_t)m * D2 * D2 + i;  // A[i + k*D2]
            double s0 = 0.0, s1 = 0.0;
            int k = 0;
            for (; k + 1 < D2; k += 2) {
                s0 = fma(a[(size_t)k * D2], dx[k], s0);
                s1 = fma(a[(size_t)(k + 1) * D2], dx[k + 1], s1);
            }
            if (k < D2) s0 = fma(a[(size_t)k * D2], dx[k], s0);
            const double s = (s0 + s1) + muy[(size_t)m * D2 + i];
            ev[i] = s;
            E[t * D2 + i] = s;
            if (Eout) Eout[t * D2 + i] = s;
        }
        __syncthreads();
        if (live && i < D2) {
            const double* pm = P + (size_t)m * D2 * D2 + i;  // symmetric: row i == column i
            double s0 = 0.0, s1 = 0.0;
            int k = 0;
            for (; k + 1 < D2; k += 2) {
                s0 = fma(pm[(size_t)k * D2], ev[k], s0);
                s1 = fma(pm[(size_t)(k + 1) * D2], ev[k + 1], s1);
            }
            if (k < D2) s0 = fma(pm[(size_t)k * D2], ev[k], s0);
            Gv[t * D2 + i] = s0 + s1;
        }
    }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }


// FP64 tensor-core tile product (mma.sync m8n8k4): acc (8x8 tile at rows 8*i8, cols 8*j8) +=
// sum_{k < kmax} X[i][k] * Y[j][k], X and Y column-major in shared memory (XT[k*LD + i] = X[i][k]).
// Fragment layout: A a0 = X[8*i8 + lane/4][k + lane%4], B b0 = Y[8*j8 + lane/4][k + lane%4],
// C c0,c1 = [8*i8 + lane/4][8*j8 + 2*(lane%4) + {0,1}].  One warp instruction performs 256 FMAs,
// so the block products cost ~8x fewer issue slots than per-thread DFMA tiles.
//
// NI output tiles of one block column j8 (rows 8 * i8[e]) at once: the B fragment is loaded once per k-step and the
// NI accumulator chains are independent, so neither the shared-memory latency nor the 26-cycle DMMA latency of a
// single dependent chain is exposed (dmma_tile, one tile at a time, spent 56 cycles per k-step on both).
template <int NI>
__device__ __forceinline__ void dmma_col(double (&c)[3][2], const double* __restrict__ XT, const double* __restrict__ YT,
                                         const int (&i8)[3], int j8, int kmax, int LD, int lane) {
    const int r = lane >> 2, q = lane & 3;
    const double* yb = YT + q * LD + j8 * 8 + r;
    const double* xa[NI];
#pragma unroll
    for (int e = 0; e < NI; ++e) xa[e] = XT + q * LD + i8[e] * 8 + r;
#pragma unroll 2
    for (int k = 0; k < kmax; k += 4) {
        const double b = yb[k * LD];
        double a[NI];
#pragma unroll
        for (int e = 0; e < NI; ++e) a[e] = xa[e][k * LD];
#pragma unroll
        for (int e = 0; e < NI; ++e)
            asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(c[e][0]), "+d"(c[e][1]) : "d"(a[e]), "d"(b));
    }
}

// acc = X Y' (+ X2 Y2') over the TS x TS output tiles, dealt to the two warps by (i8 + j8) parity so that a warp owns
// two or three tiles of every block column; tri: Y is lower triangular (k < 8 (j8 + 1)); store(i8, j8, c0, c1).
template <int TS, class Store>
__device__ __forceinline__ void block_product(const double* XT, const double* YT, const double* XT2, const double* YT2, bool tri,
                                              bool enable, int warp, int lane, int LD, Store store) {
    for (int j8 = 0; j8 < TS; ++j8) {
        const int ib = (j8 + warp) & 1, ni = (TS - ib + 1) / 2, kmax = tri ? 8 * (j8 + 1) : 8 * TS;
        const int i8[3] = {ib, ib + 2, ib + 4};
        double c[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
        if (enable) {
            if (ni == 3) {
                dmma_col<3>(c, XT, YT, i8, j8, kmax, LD, lane);
                if (XT2) dmma_col<3>(c, XT2, YT2, i8, j8, kmax, LD, lane);
            } else if (ni == 2) {
                dmma_col<2>(c, XT, YT, i8, j8, kmax, LD, lane);
                if (XT2) dmma_col<2>(c, XT2, YT2, i8, j8, kmax, LD, lane);
            } else if (ni == 1) {
                dmma_col<1>(c, XT, YT, i8, j8, kmax, LD, lane);
                if (XT2) dmma_col<1>(c, XT2, YT2, i8, j8, kmax, LD, lane);
            }
        }
        for (int e = 0; e < ni; ++e) store(i8[e], j8, c[e][0], c[e][1]);
    }
}

// One CTA (64 threads = 8 x 8 grid of TS x TS register tiles) per chunk; Ds <= DSP = 8*TS, the
// padding rows/columns carry a unit diagonal so the factorisation is unaffected.
// Shared-memory plan (column-major blocks, LD = DSP+1): G2 | G1[2] | W[3] where the W buffer of
// step t first stages R[t][t-2], then holds Tm, and finally Linv_t.
template <int TS>
__global__ void __launch_bounds__(64, (TS <= 3) ? 7 : ((TS == 4) ? 4 : 2))
traj_solve_tiled(const TrajParams p) {
    constexpr int DSP = 8 * TS, LD = DSP + 1, BS = DSP * LD;
    constexpr int VEC = (4 * TS + 5) * DSP + 2 * TS * TS;
    const int Ds = p.Ds, D2 = 2 * Ds, BB = Ds * Ds;
    const int tid = threadIdx.x;
    const int ti = tid & 7, tj = tid >> 3, i0 = ti * TS, j0 = tj * TS;
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t c0 = p.chunk_off[blockIdx.x];
    const int T = (int)(p.chunk_off[blockIdx.x + 1] - c0);
    if (T <= 0) return;

    extern __shared__ __align__(16) double sm[];
    double* const G2 = sm;
    double* const vec = sm + 6 * BS;
    double* const lcol = vec;                       // [2][DSP][TS]   panel of L, by block parity
    double* const xrow = vec + 2 * DSP * TS;        // [2][TS][DSP]   rows of L^-1
    double* const dinvb = vec + 4 * DSP * TS;       // [2][TS][TS]    inverse of the diagonal tile
    double* const rv = dinvb + 2 * TS * TS;
    double* const tmpv = rv + DSP;
    double* const zb = tmpv + DSP;                  // [3][DSP]
    for (int e = tid; e < 6 * BS + VEC; e += 64) sm[e] = 0.0;

    const int32_t* mh = p.mhat + c0;
    const double* gv = p.Gv + c0 * D2;
    double* Lst = p.Lst + c0 * 3 * BB;
    double* Zg = p.Z + c0 * Ds;
    const int estep_i = 64 % Ds, estep_j = 64 / Ds;   // element walk e -> e + 64 without divisions
    __syncthreads();

    // =========================== forward: block Cholesky + L z = r ===========================
    for (int t = 0; t < T; ++t) {
        double* const G1 = sm + (1 + (t & 1)) * BS;
        const double* const Lt1t2 = sm + (1 + ((t & 1) ^ 1)) * BS;
        double* const W = sm + (3 + t % 3) * BS;
        const double* const Lm1inv = sm + (3 + (t + 2) % 3) * BS;
        const double* const Lm2inv = sm + (3 + (t + 1) % 3) * BS;
        double* const zt = zb + (t % 3) * DSP;
        const double* const z1 = zb + ((t + 2) % 3) * DSP;
        const double* const z2 = zb + ((t + 1) % 3) * DSP;
        const double* Pt = p.P + (size_t)mh[t] * D2 * D2;
        const double* Pm = (t >= 1) ? p.P + (size_t)mh[t - 1] * D2 * D2 : nullptr;
        const double* Pp = (t + 1 < T) ? p.P + (size_t)mh[t + 1] * D2 * D2 : nullptr;

        // ---- 0. own tiles of R[t][t] (-> S accumulators) and R[t][t-1] straight from L2 into
        //         registers; R[t][t-2] and r_t staged in shared memory
        double s[TS][TS], r1[TS][TS];
#pragma unroll
        for (int x = 0; x < TS; ++x)
#pragma unroll
            for (int y = 0; y < TS; ++y) {
                const int i = i0 + x, j = j0 + y;
                double v = (i == j) ? 1.0 : 0.0, u = 0.0;
                if (i < Ds && j < Ds) {
                    v = Pt[i + (size_t)j * D2];
                    if (Pm) {
                        v = fma(0.25, Pm[(Ds + i) + (size_t)(Ds + j) * D2], v);
                        u = 0.5 * Pm[(Ds + i) + (size_t)j * D2] - 0.5 * Pt[i + (size_t)(Ds + j) * D2];
                    }
                    if (Pp) v = fma(0.25, Pp[(Ds + i) + (size_t)(Ds + j) * D2], v);
                }
                s[x][y] = v;
                r1[x][y] = u;
            }
        if (t >= 2) {
            // the W buffer still holds Linv_{t-3}, whose padding diagonal is 1: clear the padding
            if (Ds < DSP) {
#pragma unroll
                for (int x = 0; x < TS; ++x)
#pragma unroll
                    for (int y = 0; y < TS; ++y)
                        if (i0 + x >= Ds || j0 + y >= Ds) W[(j0 + y) * LD + i0 + x] = 0.0;
            }
            int i = tid % Ds, j = tid / Ds;
            for (int e = tid; e < BB; e += 64) {
                W[j * LD + i] = -0.25 * Pm[(Ds + i) + (size_t)(Ds + j) * D2];
                i += estep_i; j += estep_j;
                if (i >= Ds) { i -= Ds; ++j; }
            }
        }
        if (tid < DSP) {
            double r = 0.0;
            if (tid < Ds) {
                r = gv[(size_t)t * D2 + tid];
                if (t >= 1) r = fma(0.5, gv[(size_t)(t - 1) * D2 + Ds + tid], r);
                if (t + 1 < T) r = fma(-0.5, gv[(size_t)(t + 1) * D2 + Ds + tid], r);
            }
            rv[tid] = r;
        }
        __syncthreads();
        // Block products on the FP64 tensor cores: the TS x TS output tiles (8 x 8 each) are dealt
        // to the two warps; fragment element (row, col pair) of this lane inside a tile:
        const int fr = lane >> 2, fc = 2 * (lane & 3);
        // ---- 1. G2 = L[t][t-2] = R[t][t-2] * Linv_{t-2}'   (Linv lower triangular: k < 8*(j8+1))
        block_product<TS>(W, Lm2inv, nullptr, nullptr, true, t >= 2, warp, lane, LD, [&](int i8, int j8, double c0, double c1) {
            G2[(8 * j8 + fc) * LD + 8 * i8 + fr] = c0;
            G2[(8 * j8 + fc + 1) * LD + 8 * i8 + fr] = c1;
        });
        __syncthreads();
        // ---- 2. Tm = R[t][t-1] - G2 * L[t-1][t-2]'  (into the W buffer; R[t][t-2] is dead)
#pragma unroll
        for (int y = 0; y < TS; ++y)
#pragma unroll
            for (int x = 0; x < TS; ++x) W[(j0 + y) * LD + i0 + x] = r1[x][y];
        __syncthreads();
        if (t >= 2) {
            block_product<TS>(G2, Lt1t2, nullptr, nullptr, false, true, warp, lane, LD, [&](int i8, int j8, double c0, double c1) {
                W[(8 * j8 + fc) * LD + 8 * i8 + fr] -= c0;
                W[(8 * j8 + fc + 1) * LD + 8 * i8 + fr] -= c1;
            });
        }
        __syncthreads();
        // ---- 3. G1 = L[t][t-1] = Tm * Linv_{t-1}'
        block_product<TS>(W, Lm1inv, nullptr, nullptr, true, t >= 1, warp, lane, LD, [&](int i8, int j8, double c0, double c1) {
            G1[(8 * j8 + fc) * LD + 8 * i8 + fr] = c0;
            G1[(8 * j8 + fc + 1) * LD + 8 * i8 + fr] = c1;
        });
        __syncthreads();
        // ---- 4. S = R[t][t] - G2 G2' - G1 G1'  (products into the W buffer -- Tm is dead -- then
        //         every thread subtracts its own register tile)
        if (t >= 1) {
            block_product<TS>(G1, G1, t >= 2 ? G2 : nullptr, G2, false, true, warp, lane, LD, [&](int i8, int j8, double c0, double c1) {
                W[(8 * j8 + fc) * LD + 8 * i8 + fr] = c0;
                W[(8 * j8 + fc + 1) * LD + 8 * i8 + fr] = c1;
            });
            __syncthreads();
#pragma unroll
            for (int y = 0; y < TS; ++y)
#pragma unroll
                for (int x = 0; x < TS; ++x) s[x][y] -= W[(j0 + y) * LD + i0 + x];
        }
        // ---- 5. blocked right-looking Cholesky of S fused with W <- L^-1.  S and W tiles stay in
        //         registers; per block column: the diagonal tile is factorised and inverted by its
        //         owner, the panel and the matching rows of L^-1 are formed with that inverse and
        //         broadcast through shared memory, then every trailing tile takes a TS-rank update.
        double w[TS][TS];
#pragma unroll
        for (int x = 0; x < TS; ++x)
#pragma unroll
            for (int y = 0; y < TS; ++y) w[x][y] = (i0 + x == j0 + y) ? 1.0 : 0.0;
#pragma unroll
        for (int kb = 0; kb < 8; ++kb) {
            double* const dv = dinvb + (kb & 1) * TS * TS;
            double* const lc = lcol + (kb & 1) * DSP * TS;
            double* const xr = xrow + (kb & 1) * TS * DSP;
            if (ti == kb && tj == kb) {
                double di[TS];
#pragma unroll
                for (int c = 0; c < TS; ++c) {
                    const double d = s[c][c];
                    if (!sane_pivot(d)) atomicExch(p.err, 1);        // not positive (or not a sane pivot at all), as in the warp solver
                    di[c] = fast_rsqrt(d);
                    s[c][c] = d * di[c];
#pragma unroll
                    for (int r = c + 1; r < TS; ++r) s[r][c] *= di[c];
#pragma unroll
                    for (int r = c + 1; r < TS; ++r)
#pragma unroll
                        for (int c2 = c + 1; c2 <= r; ++c2) s[r][c2] = fma(-s[r][c], s[c2][c], s[r][c2]);
                }
                double inv[TS][TS];
#pragma unroll
                for (int c = 0; c < TS; ++c) {
#pragma unroll
                    for (int r = 0; r < TS; ++r) inv[r][c] = 0.0;
                    inv[c][c] = di[c];
#pragma unroll
                    for (int r = c + 1; r < TS; ++r) {
                        double a = 0.0;
#pragma unroll
                        for (int k = c; k < r; ++k) a = fma(s[r][k], inv[k][c], a);
                        inv[r][c] = -a * di[r];
                    }
                }
#pragma unroll
                for (int r = 0; r < TS; ++r)
#pragma unroll
                    for (int c = 0; c < TS; ++c) dv[r * TS + c] = inv[r][c];
            }
            __syncthreads();
            if (tj == kb && ti > kb) {        // panel: L = S_tile * Dinv'
                double dl[TS][TS], o[TS][TS];
#pragma unroll
                for (int r = 0; r < TS; ++r)
#pragma unroll
                    for (int c = 0; c < TS; ++c) dl[r][c] = dv[r * TS + c];
#pragma unroll
                for (int x = 0; x < TS; ++x)
#pragma unroll
                    for (int y = 0; y < TS; ++y) {
                        double a = 0.0;
#pragma unroll
                        for (int c = 0; c <= y; ++c) a = fma(s[x][c], dl[y][c], a);
                        o[x][y] = a;
                    }
#pragma unroll
                for (int x = 0; x < TS; ++x)
#pragma unroll
                    for (int y = 0; y < TS; ++y) { s[x][y] = o[x][y]; lc[(i0 + x) * TS + y] = o[x][y]; }
            }
            if (ti == kb && tj <= kb) {       // rows of L^-1: X = Dinv * W_tile
                double dl[TS][TS], o[TS][TS];
#pragma unroll
                for (int r = 0; r < TS; ++r)
#pragma unroll
                    for (int c = 0; c < TS; ++c) dl[r][c] = dv[r * TS + c];
#pragma unroll
                for (int x = 0; x < TS; ++x)
#pragma unroll
                    for (int y = 0; y < TS; ++y) {
                        double a = 0.0;
#pragma unroll
                        for (int c = 0; c <= x; ++c) a = fma(dl[x][c], w[c][y], a);
                        o[x][y] = a;
                    }
#pragma unroll
                for (int x = 0; x < TS; ++x)
#pragma unroll
                    for (int y = 0; y < TS; ++y) { w[x][y] = o[x][y]; xr[x * DSP + j0 + y] = o[x][y]; }
            }
            __syncthreads();
            if (ti > kb) {
                double li[TS][TS];
#pragma unroll
                for (int x = 0; x < TS; ++x)
#pragma unroll
                    for (int c = 0; c < TS; ++c) li[x][c] = lc[(i0 + x) * TS + c];
                if (tj > kb) {
#pragma unroll
                    for (int y = 0; y < TS; ++y)
#pragma unroll
                        for (int c = 0; c < TS; ++c) {
                            const double ljv = lc[(j0 + y) * TS + c];
#pragma unroll
                            for (int x = 0; x < TS; ++x) s[x][y] = fma(-li[x][c], ljv, s[x][y]);
                        }
                } else {
#pragma unroll
                    for (int y = 0; y < TS; ++y)
#pragma unroll
                        for (int c = 0; c < TS; ++c) {
                            const double xv = xr[c * DSP + j0 + y];
#pragma unroll
                            for (int x = 0; x < TS; ++x) w[x][y] = fma(-li[x][c], xv, w[x][y]);
                        }
                }
            }
        }
#pragma unroll
        for (int y = 0; y < TS; ++y)
#pragma unroll
            for (int x = 0; x < TS; ++x) W[(j0 + y) * LD + i0 + x] = w[x][y];
        __syncthreads();
        // ---- 6. z_t = Linv (r - G1 z_{t-1} - G2 z_{t-2})   (thread = row)
        if (tid < DSP) {
            double s0 = 0.0, s1 = 0.0;
            if (t >= 1) {
#pragma unroll 4
                for (int k = 0; k < DSP; ++k) s0 = fma(G1[k * LD + tid], z1[k], s0);
            }
            if (t >= 2) {
#pragma unroll 4
                for (int k = 0; k < DSP; ++k) s1 = fma(G2[k * LD + tid], z2[k], s1);
            }
            tmpv[tid] = rv[tid] - (s0 + s1);
        }
        __syncthreads();
        if (tid < DSP) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll 2
            for (int k = 0; k + 1 < DSP; k += 2) {
                s0 = fma(W[k * LD + tid], tmpv[k], s0);
                s1 = fma(W[(k + 1) * LD + tid], tmpv[k + 1], s1);
            }
            const double z = s0 + s1;
            zt[tid] = z;
            if (tid < Ds) Zg[(size_t)t * Ds + tid] = z;
        }
        // ---- 7. stream the three blocks of block-row t to HBM (row-major, compact)
        {
            double* dst = Lst + (size_t)t * 3 * BB;
            int j = tid % Ds, i = tid / Ds;
            for (int e = tid; e < BB; e += 64) {
                dst[e] = W[j * LD + i];
                dst[BB + e] = G1[j * LD + i];
                dst[2 * BB + e] = G2[j * LD + i];
                j += estep_i; i += estep_j;
                if (j >= Ds) { j -= Ds; ++i; }
            }
        }
        __syncthreads();
    }

    // =========================== backward: L' y = z ===========================================
    // y_t = Linv_t' (z_t - L[t+1][t]' y_{t+1} - L[t+2][t]' y_{t+2});  thread = output row j
    const bool vec_ok = (BB % 2) == 0;
    auto prefetch = [&](int t, int buf) {
        const int n16 = BB / 2;  // 16-byte pieces per block
        for (int e = tid; e < 3 * n16; e += 64) {
            const int b = e / n16, o = e - b * n16;
            if (t + b < T) cp_async16(sm + (buf * 3 + b) * BB + 2 * o, Lst + ((size_t)(t + b) * 3 + b) * BB + 2 * o);
        }
    };
    if (vec_ok) { prefetch(T - 1, (T - 1) & 1); cp_async_commit(); }
    for (int t = T - 1; t >= 0; --t) {
        const int buf = t & 1;
        double* const yt = zb + (t % 3) * DSP;
        const double* const y1 = zb + ((t + 1) % 3) * DSP;
        const double* const y2 = zb + ((t + 2) % 3) * DSP;
        if (vec_ok) {
            if (t >= 1) prefetch(t - 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            for (int e = tid; e < 3 * BB; e += 64) {
                const int b = e / BB, o = e - b * BB;
                if (t + b < T) sm[(buf * 3 + b) * BB + o] = Lst[((size_t)(t + b) * 3 + b) * BB + o];
            }
        }
        __syncthreads();
        const double* const Li = sm + (buf * 3 + 0) * BB;
        const double* const L1 = sm + (buf * 3 + 1) * BB;
        const double* const L2 = sm + (buf * 3 + 2) * BB;
        if (tid < Ds) {
            double s0 = 0.0, s1 = 0.0;
            if (t + 1 < T) {
#pragma unroll 4
                for (int i = 0; i < Ds; ++i) s0 = fma(L1[i * Ds + tid], y1[i], s0);
            }
            if (t + 2 < T) {
#pragma unroll 4
                for (int i = 0; i < Ds; ++i) s1 = fma(L2[i * Ds + tid], y2[i], s1);
            }
            rv[tid] = Zg[(size_t)t * Ds + tid] - (s0 + s1);
        }
        __syncthreads();
        if (tid < Ds) {
            double s0 = 0.0;
            for (int i = tid; i < Ds; ++i) s0 = fma(Li[i * Ds + tid], rv[i], s0);
            yt[tid] = s0;
            p.Y[(c0 + t) * p.ldy + tid] = s0;  // reshape(y, D, T)  src/trajectory_gmmmap.jl:109
        }
        if (p.copy_power && tid == 63) p.Y[(c0 + t) * p.ldy - 1] = p.Xpow[(c0 + t) * p.ldx - 1];  // src/common.jl:60
        __syncthreads();
    }
}

}  // namespace

template <int TS>
static int32_t launch_tiled(const TrajParams& p, int64_t nchunks, cudaStream_t st) {
    constexpr int DSP = 8 * TS, LD = DSP + 1;
    const size_t smem = ((size_t)6 * DSP * LD + (4 * TS + 5) * DSP + 2 * TS * TS) * sizeof(double);
    auto k = traj_solve_tiled<TS>;
    VCB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)nchunks, 64, smem, st>>>(p);
    count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

int32_t traj_solve_device(const vcb_traj& tr, const double* dX, int64_t ldx, const int32_t* d_mhat,
                          const int64_t* d_chunk_off, int64_t nchunks, int max_chunk_len,
                          int64_t total, double* dY, int64_t ldy, double* dEy_out, bool copy_power,
                          cudaStream_t st, const int* ws, int64_t npanels) {
    (void)max_chunk_len;
    if (total == 0 || nchunks == 0) return VCB_OK;
    const vcb_gmmmap& g = *tr.g;
    const int Ds = tr.Ds, D2 = 2 * Ds, BB = Ds * Ds;
    if (Ds > 48) return fail(VCB_EUNSUPPORTED, "static dimension %d exceeds the trajectory solver's limit (48)", Ds);

    double *dE = nullptr, *dG = nullptr, *dL = nullptr, *dZ = nullptr;
    int* const derr = tr.d_err.p;   // sticky, read by the host entry points / vcb_traj_status
    VCB_CUDA(cudaMallocAsync((void**)&dE, (size_t)total * D2 * sizeof(double), st));
    VCB_CUDA(cudaMallocAsync((void**)&dG, (size_t)total * D2 * sizeof(double), st));
    // VCB_TRAJ_SOLVER=tiled keeps the 64-thread CTA solver for dimensions the warp solver covers
    static const bool force_tiled = [] { const char* e = getenv("VCB_TRAJ_SOLVER"); return e && e[0] == 't'; }();
    const size_t warp_bytes = force_tiled ? 0 : traj_warp_factor_bytes(Ds);
    const size_t factor_bytes = std::max((size_t)3 * BB * sizeof(double), warp_bytes);
    VCB_CUDA(cudaMallocAsync((void**)&dL, (size_t)total * factor_bytes, st));
    VCB_CUDA(cudaMallocAsync((void**)&dZ, (size_t)total * Ds * sizeof(double), st));
    if (ws) {
        // frames bucketed by mixture: one Float64 GEMM per mixture panel (vcb_group.cu)
        VCB_TRY(group_e_step(tr, ws, npanels, dX, ldx, dE, dEy_out, dG, st));
    } else {
        const int tpf = round_up(D2, 32), fpb = std::max(1, 256 / tpf);
        const int groups = 1;   // more groups trade parallelism for L1 reuse of A_m / P_m; 1 measured fastest on B200
        const int64_t per_block = (int64_t)fpb * groups;
        dim3 block(tpf, fpb), grid((unsigned)((total + per_block - 1) / per_block));
        traj_e_kernel<<<grid, block, (size_t)fpb * 2 * D2 * sizeof(double), st>>>(
            dX, ldx, total, d_mhat, g.d_A.p, g.d_mux.p, g.d_muy.p, tr.d_P.p, D2, groups, dE, dG, dEy_out);
        count_launch();
        VCB_CUDA(cudaGetLastError());
    }
    stage_mark(st);
    int32_t rc;
    {
        TrajParams p{};
        p.P = tr.d_P.p; p.mhat = d_mhat; p.Gv = dG; p.chunk_off = d_chunk_off; p.Lst = dL; p.Z = dZ;
        p.Y = dY; p.ldy = ldy; p.Xpow = dX; p.ldx = ldx; p.copy_power = copy_power ? 1 : 0;
        p.Ds = Ds; p.err = derr; p.nchunks = nchunks;
        if (warp_bytes) rc = traj_warp_launch(p, nchunks, st);
        else switch ((Ds + 7) / 8) {
            case 1: rc = launch_tiled<1>(p, nchunks, st); break;
            case 2: rc = launch_tiled<2>(p, nchunks, st); break;
            case 3: rc = launch_tiled<3>(p, nchunks, st); break;
            case 4: rc = launch_tiled<4>(p, nchunks, st); break;
            case 5: rc = launch_tiled<5>(p, nchunks, st); break;
            default: rc = launch_tiled<6>(p, nchunks, st); break;
        }
    }
    stage_mark(st);
    cudaFreeAsync(dE, st);
    cudaFreeAsync(dG, st);
    cudaFreeAsync(dL, st);
    cudaFreeAsync(dZ, st);
    return rc;
}

}  // namespace vcb
