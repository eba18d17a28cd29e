// vcb_traj_warp.cu -- K3w: trajectory ML solve, one WARP per chunk, blocks kept as FP64
// tensor-core fragments (Ds <= 24).
//
// Same algorithm as vcb_traj.cu (block Cholesky of the block-pentadiagonal normal matrix
// R = W' D^-1 W of reference src/trajectory_gmmmap.jl:95-105, sequential in time), laid out for
// mma.sync.m8n8k4.f64:
//   * every Ds x Ds block is NT x NT tiles of 8 x 8 (Ds <= 8 NT, padding carries a unit diagonal);
//     a tile lives in the accumulator layout of the instruction: lane (r = lane/4, q = lane%4)
//     holds X[r][2q], X[r][2q+1];
//   * all block products of the factorisation have the form X * Y' (G2 = R2 Linv2', Tm = R1 - G2 Lp',
//     G1 = Tm Linv1', S = R0 - G2 G2' - G1 G1', the Cholesky panel / trailing updates and the
//     recursion for L^-1).  Because the contraction index may be visited in any order, two
//     accumulator-layout tiles feed one product directly: DMMA #1 contracts the even columns
//     (a = x.x, b = y.x), DMMA #2 the odd ones -- no shuffles, no shared-memory round trip;
//   * the 8 x 8 diagonal tiles are factorised redundantly by every lane from a shared-memory copy
//     (a warp has no cheaper way through that serial chain) and each lane back-substitutes the row
//     of the inverse it owns;
//   * no __syncthreads anywhere: a CTA is one warp, so chunks never wait for each other.
// Per frame the kernel issues ~240 DMMA (61 k FP64 FMA) -- the FP64 pipe is the roofline -- and
// streams 15 tiles (7.5 KB at NT = 3: Linv_t and L[t][t-1]) of factors to HBM for the back
// substitution, which re-reads them once in fragment order (each lane reads exactly the bytes it
// wrote); L[t][t-2] = -1/4 Pdd Linv' is not stored, its product is rebuilt there.
#include <cstdlib>

#include "vcb_kernels.h"
#include "vcb_traj.h"

namespace vcb {

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void mma2(double2& c, const double2 x, const double2 y) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c.x), "+d"(c.y) : "d"(x.x), "d"(y.x));
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c.x), "+d"(c.y) : "d"(x.y), "d"(y.y));
}
__device__ __forceinline__ double2 neg2(const double2 v) { return make_double2(-v.x, -v.y); }
__device__ __forceinline__ double2 zero2() { return make_double2(0.0, 0.0); }

// transpose of an 8 x 8 tile in accumulator layout
__device__ __forceinline__ double2 tile_transpose(const double2 x, int r, int q) {
    const int s0 = (8 * q) | (r >> 1), s1 = (8 * q + 4) | (r >> 1);
    const double a0 = __shfl_sync(kFull, x.x, s0), a1 = __shfl_sync(kFull, x.y, s0);
    const double b0 = __shfl_sync(kFull, x.x, s1), b1 = __shfl_sync(kFull, x.y, s1);
    return (r & 1) ? make_double2(a1, b1) : make_double2(a0, b0);
}

// Cholesky of the symmetric 8 x 8 tile d (lower triangle used) and the inverse X = L^-1 of its
// factor, both in accumulator layout (lane (r, q) owns columns 2q, 2q+1 of row r) and computed
// COOPERATIVELY: right-looking elimination, one column per step.  Per column the pivot, the scaled
// column (the lane's row entry and its two column entries) and the finished row of X travel by
// shuffle; every lane then updates only its own two elements of the trailing matrix and of the
// forward substitution L X = I.  12 FP64 instructions per column instead of the ~28 of a
// factorisation repeated in every lane -- DMMA and DFMA share one pipe, and it is the bound.
__device__ __forceinline__ double2 diag_inverse(const double2 d, int lane, int r, int q, int* err) {
    double2 a = d;
    double2 w = make_double2(r == 2 * q ? 1.0 : 0.0, r == 2 * q + 1 ? 1.0 : 0.0);
    bool bad = false;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int qc = c >> 1;
        const double mine = (c & 1) ? a.y : a.x;                   // a[r][c] in the lanes with q == qc
        const double piv = __shfl_sync(kFull, mine, 4 * c + qc);
        bad |= !sane_pivot(piv);                                    // not positive (or not a sane pivot at all)
        const double di = fast_rsqrt(piv);
        const double lcol = mine * di;                              // l[r][c], rows >= c, lanes q == qc
        const double lr = __shfl_sync(kFull, lcol, (lane & ~3) | qc);
        if (c < 7) {
            const double l0 = __shfl_sync(kFull, lcol, 8 * q + qc), l1 = __shfl_sync(kFull, lcol, 8 * q + 4 + qc);
            if (r > c) {                                            // trailing update, columns > c
                if (2 * q > c) a.x = fma(-lr, l0, a.x);
                if (2 * q + 1 > c) a.y = fma(-lr, l1, a.y);
            }
        }
        // row c of X is complete: X[c][:] = w[c][:] / l[c][c]; rows below subtract l[r][c] X[c][:]
        const double2 xr = make_double2(w.x * di, w.y * di);
        const double x0 = __shfl_sync(kFull, xr.x, 4 * c + q), x1 = __shfl_sync(kFull, xr.y, 4 * c + q);
        if (r == c) w = xr;
        if (r > c) {
            w.x = fma(-lr, x0, w.x);
            w.y = fma(-lr, x1, w.y);
        }
    }
    if (bad) atomicExch(err, 1);
    return w;
}

// The same, REDUNDANTLY: every lane factorises the whole tile from a shared-memory copy, then solves
// L' y = e_r for the row r of L^-1 it owns.  ~2.2x the FP64 instructions of the cooperative form but
// no shuffle on the serial chain (8 x (rsqrt + 2 FP64 ops) per tile): with one or two warps per
// scheduler the chain, not the pipe, sets the pace, and this form measured 12 % faster (4.68 vs
// 5.24 ms for 1000 x 500 frames).
__device__ __forceinline__ double2 diag_inverse_redundant(const double2 d, double* sd, int r, int q, int* err) {
    __syncwarp();
    *reinterpret_cast<double2*>(sd + r * 8 + 2 * q) = d;
    __syncwarp();
    double a[8][8], di[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c <= i; c += 2) {
            const double2 v = *reinterpret_cast<const double2*>(sd + i * 8 + c);
            a[i][c] = v.x;
            if (c + 1 <= i) a[i][c + 1] = v.y;
        }
    bool bad = false;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const double dd = a[c][c];
        bad |= !sane_pivot(dd);                        // not positive (or not a sane pivot at all)
        di[c] = fast_rsqrt(dd);
#pragma unroll
        for (int i = c + 1; i < 8; ++i) a[i][c] *= di[c];
#pragma unroll
        for (int i = c + 1; i < 8; ++i)
#pragma unroll
            for (int c2 = c + 1; c2 <= i; ++c2) a[i][c2] = fma(-a[i][c], a[c2][c], a[i][c2]);
    }
    if (bad) atomicExch(err, 1);
    double y[8], s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = (r == k) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 7; k >= 0; --k) {
        y[k] = s[k] * di[k];
#pragma unroll
        for (int m = 0; m < k; ++m) s[m] = fma(-a[k][m], y[k], s[m]);
    }
    double2 o;
    o.x = (q == 0) ? y[0] : (q == 1) ? y[2] : (q == 2) ? y[4] : y[6];
    o.y = (q == 0) ? y[1] : (q == 1) ? y[3] : (q == 2) ? y[5] : y[7];
    return o;
}

template <int NT>
struct WarpLayout {
    static constexpr int DSP = 8 * NT, NL = NT * (NT + 1) / 2, NF = NT * NT, FT = NL + NF;   // stored: Linv_t, L[t][t-1]
    __host__ __device__ static constexpr int low(int i, int j) { return i * (i + 1) / 2 + j; }
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async16_cg(void* smem, const void* gmem) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(gmem) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Blocked Cholesky of the symmetric block S (lower tiles, destroyed) fused with the inverse of its
// factor: Li = L^-1 (lower tiles), all tiles in accumulator layout (U[j][k] = Linv[k][j]').
template <int NT, bool COOP>
__device__ __forceinline__ void block_cholesky_inverse(double2* S, double2* Li, double* sdiag, const int lane, int* err) {
    using LY = WarpLayout<NT>;
    constexpr int NL = LY::NL;
    const int r = lane >> 2, q = lane & 3;
    double2 Lpan[NL], U[NL];
#pragma unroll
    for (int kb = 0; kb < NT; ++kb) {
        const double2 Dinv = COOP ? diag_inverse(S[LY::low(kb, kb)], lane, r, q, err)
                                 : diag_inverse_redundant(S[LY::low(kb, kb)], sdiag, r, q, err);
        Li[LY::low(kb, kb)] = Dinv;
        if (kb + 1 < NT) U[LY::low(kb, kb)] = tile_transpose(Dinv, r, q);
#pragma unroll
        for (int i = kb + 1; i < NT; ++i) {        // panel L[i][kb] = S[i][kb] Dinv'
            double2 c = zero2();
            mma2(c, S[LY::low(i, kb)], Dinv);
            Lpan[LY::low(i, kb)] = c;
        }
#pragma unroll
        for (int j = kb + 1; j < NT; ++j) {        // trailing update
            const double2 nj = neg2(Lpan[LY::low(j, kb)]);
#pragma unroll
            for (int i = j; i < NT; ++i) mma2(S[LY::low(i, j)], Lpan[LY::low(i, kb)], nj);
        }
#pragma unroll
        for (int j = 0; j < kb; ++j) {             // Linv[kb][j] = -Dinv sum_k L[kb][k] Linv[k][j]
            double2 mt = zero2();                  // M' = sum_k U[j][k] L[kb][k]'
#pragma unroll
            for (int k = j; k < kb; ++k) mma2(mt, U[LY::low(k, j)], Lpan[LY::low(kb, k)]);
            const double2 nmt = neg2(mt);
            double2 c = zero2();
            mma2(c, Dinv, nmt);
            Li[LY::low(kb, j)] = c;
            if (kb + 1 < NT) {
                double2 u = zero2();
                mma2(u, nmt, Dinv);
                U[LY::low(kb, j)] = u;             // stored at (max, min): U[j][kb]
            }
        }
    }
}

// Back substitution L' y = z of one chunk for Ds == 8 NT (see the comment at its call sites):
// y_t = Linv_t' (z_t - L[t+1][t]' y_{t+1} - L[t+2][t]' y_{t+2}).  Run by one warp.
template <int NT>
__device__ __forceinline__ void back_sweep_ring(const TrajParams& p, double2 (*const pool)[32], int npool, double* const sz0,
                                                double* const sdiag, double* const sw, const double2* const Fb,
                                                const double* const Zg, const int32_t* const mh, const int64_t c0,
                                                const int T, const int lane) {
    using LY = WarpLayout<NT>;
    constexpr int DSP = LY::DSP, NL = LY::NL, NF = LY::NF, FT = LY::FT;
    constexpr bool FULL = true;
    const int r = lane >> 2, q = lane & 3;
    const int Ds = DSP, D2 = 2 * DSP;
    const size_t lofs = (size_t)r * D2 + 2 * q;
    (void)npool;
    // The serial chain of a step is short (two mat-vecs and their shuffle reductions), so the sweep
    // runs at the speed its factor tiles arrive: they come through a ring of three steps of
    // cp.async groups in the tile pool (Linv_t, L[t+1][t], z_t), requested three steps ahead.
    double* const sy = sz0;   // ring of three, row-layout reads
    double* const su = sdiag;       // u_t = R[t+2][t] y_{t+2}, prepared one step ahead
    for (int e = lane; e < 3 * DSP; e += 32) sy[e] = 0.0;
    for (int e = lane; e < DSP; e += 32) su[e] = 0.0;
    cp_async_wait_all();
    __syncwarp();
    double2(*const ring)[32] = pool;
    double* const zring = reinterpret_cast<double*>(pool + 3 * FT);
    
    auto bfetch = [&](int t) {
        if (t >= 0) {
            const double2* F = Fb + (size_t)t * FT * 32;
            double2(*const slot)[32] = ring + (t % 3) * FT;
#pragma unroll
            for (int e = 0; e < NL; ++e) cp_async16_cg(&slot[e][lane], F + e * 32);
            if (t + 1 < T) {
#pragma unroll
                for (int e = 0; e < NF; ++e) cp_async16_cg(&slot[NL + e][lane], F + (FT + NL + e) * 32);
            }
            if (lane < DSP / 2) cp_async16_cg(zring + (t % 3) * DSP + 2 * lane, Zg + (size_t)t * DSP + 2 * lane);
        }
        cp_async_commit();
    };
    bfetch(T - 1); bfetch(T - 2); bfetch(T - 3);
    int mcur = mh[T - 1], mnext = mh[T >= 2 ? T - 2 : 0];
    for (int t = T - 1; t >= 0; --t) {
        // P tiles behind u_{t-1} = -1/4 Pdd(mhat_t) y_{t+1}: requested now, consumed at the end of the step
        const bool need_u = t >= 1 && t + 1 < T;
        double2 Pd[NF];
        if (need_u) {
            const double* Pq = p.P + (size_t)mcur * D2 * D2;
#pragma unroll
            for (int i = 0; i < NT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) Pd[i * NT + j] = __ldg(reinterpret_cast<const double2*>(Pq + lofs + (Ds + 8 * j) + (size_t)(Ds + 8 * i) * D2));
        }
        const int mfar = mh[t >= 2 ? t - 2 : 0];
        cp_async_wait_group<2>();       // the group of step t has landed (two younger ones may be in flight)
        __syncwarp();                   // z_t was requested by other lanes
        const double2(*const slot)[32] = ring + (t % 3) * FT;
        const double* const zt = zring + (t % 3) * DSP;
        const double* const y1 = sy + ((t + 1) % 3) * DSP;
        double* const yt = sy + (t % 3) * DSP;
        // v = Linv_t u_t
        double v[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) v[i] = 0.0;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const double2 uc = *reinterpret_cast<const double2*>(su + 8 * j + 2 * q);
#pragma unroll
            for (int i = j; i < NT; ++i) {
                const double2 L = slot[LY::low(i, j)][lane];
                v[i] = fma(L.x, uc.x, fma(L.y, uc.y, v[i]));
            }
        }
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            v[i] += __shfl_xor_sync(kFull, v[i], 1);
            v[i] += __shfl_xor_sync(kFull, v[i], 2);
        }
        // a = L[t+1][t]' y_{t+1}
        double2 acc[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[j] = zero2();
        if (t + 1 < T) {
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                const double a = y1[8 * i + r];
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const double2 A1 = slot[NL + i * NT + j][lane];
                    acc[j].x = fma(A1.x, a, acc[j].x);
                    acc[j].y = fma(A1.y, a, acc[j].y);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int s = 4; s < 32; s <<= 1) {
                acc[j].x += __shfl_xor_sync(kFull, acc[j].x, s);
                acc[j].y += __shfl_xor_sync(kFull, acc[j].y, s);
            }
            if (r == 0) *reinterpret_cast<double2*>(sw + 8 * j + 2 * q) = acc[j];
        }
        __syncwarp();
        double2 out[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) out[j] = zero2();
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            const double w = zt[8 * i + r] - sw[8 * i + r] - v[i];
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const double2 L = slot[LY::low(i, j)][lane];
                out[j].x = fma(L.x, w, out[j].x);
                out[j].y = fma(L.y, w, out[j].y);
            }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int s = 4; s < 32; s <<= 1) {
                out[j].x += __shfl_xor_sync(kFull, out[j].x, s);
                out[j].y += __shfl_xor_sync(kFull, out[j].y, s);
            }
            if (r == 0) {
                *reinterpret_cast<double2*>(yt + 8 * j + 2 * q) = out[j];
                double* yg = p.Y + (c0 + t) * p.ldy + 8 * j + 2 * q;   // reshape(y, D, T)  src/trajectory_gmmmap.jl:109
                yg[0] = out[j].x;       // (rows of ldy = Ds + 1 doubles are not 16-byte aligned)
                yg[1] = out[j].y;
            }
        }
        // u_{t-1} = -1/4 Pdd(mhat_t) y_{t+1}  (off the chain: y_{t+1} is one step old)
        if (t >= 1) {
            double u[NT];
#pragma unroll
            for (int i = 0; i < NT; ++i) u[i] = 0.0;
            if (need_u) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const double2 yc = *reinterpret_cast<const double2*>(y1 + 8 * j + 2 * q);
#pragma unroll
                    for (int i = 0; i < NT; ++i)
                        u[i] = fma(-0.25 * Pd[i * NT + j].x, yc.x, fma(-0.25 * Pd[i * NT + j].y, yc.y, u[i]));
                }
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    u[i] += __shfl_xor_sync(kFull, u[i], 1);
                    u[i] += __shfl_xor_sync(kFull, u[i], 2);
                }
            }
#pragma unroll
            for (int i = 0; i < NT; ++i)
                if (q == 0) su[8 * i + r] = u[i];
        }
        bfetch(t - 3);      // the slot of step t is free: all its reads are complete (warp-synchronous)
        mcur = mnext;
        mnext = mfar;
        __syncwarp();
    }
}

// FULL: Ds == 8 NT (no padding, every 16-byte piece aligned): the 39 P tiles a step assembles its
// R blocks from are prefetched into shared memory with cp.async during the previous step -- one
// piece per lane and tile, read back only by the lane that requested it.
// WPC (warps = chunks per CTA): 1, or 7 with PHASE-LOCKED warp pairs.  DMMA and DFMA share the FP64 pipe of an SM
// sub-partition (tools/micro/dmma_bench.cu), and with seven one-warp CTAs per SM three sub-partitions carry two
// unrelated warps whose product phases (a DMMA stream that saturates the pipe) and factorisation phases (a
// latency-bound DFMA chain) meet at random: both in the chain, the pipe idles.  With WPC = 7 the SM holds ONE CTA;
// warps w and w + 4 sit on the same sub-partition (hardware warp slots are handed out in order) and exchange one
// named barrier per half step, so that one of them is always in its product phase while the other factorises.
template <int NT, bool FULL, int MINB = 1, bool COOP = false, int WPC = 1>
__global__ void __launch_bounds__(32 * WPC, WPC == 1 ? MINB : 1) traj_solve_warp(const TrajParams p) {
    using LY = WarpLayout<NT>;
    constexpr int DSP = LY::DSP, NL = LY::NL, NF = LY::NF, FT = LY::FT;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const int Ds = p.Ds, D2 = 2 * Ds;
    const bool even = FULL || (Ds & 1) == 0;
    const int64_t chunk = (int64_t)blockIdx.x * WPC + wib;
    if (chunk >= p.nchunks) return;
    const int64_t c0 = p.chunk_off[chunk];
    const int T = (int)(p.chunk_off[chunk + 1] - c0);
    // steps this warp walks in lock step with its partner on the same sub-partition (0: none)
    int Tc = 0;
    const bool is_b = wib >= 4;
    if (WPC > 1) {
        const int wp = wib ^ 4;
        const int64_t cp = (int64_t)blockIdx.x * WPC + wp;
        if (wp < WPC && cp < p.nchunks) Tc = min(T, (int)(p.chunk_off[cp + 1] - p.chunk_off[cp]));
    }
    const int pair_bar = 1 + (wib & 3);
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory"); };
    if (T <= 0) return;

    // one pool of 512-byte tiles: Linv_{t-1}, Linv_{t-2} (by parity of t) | L[t-1][t-2] | FULL: the P
    // tiles of the step (Pdd_{t-1}, Pds_{t-1}, Psd_t | Pss_t, Pdd_{t+1} lower); the back substitution
    // reuses the pool as a ring of factor tiles
    constexpr int NPT = 3 * NF + 2 * NL;
    constexpr int NPOOL = 2 * NL + NF + (FULL ? NPT : 0);
    constexpr int kSmallDoubles = 64 + 3 * DSP + DSP + (FULL ? DSP : 2);      // sdiag | sz | sw | srow
    constexpr size_t kWarpBytes = (size_t)NPOOL * 32 * sizeof(double2) + (size_t)((kSmallDoubles + 1) & ~1) * sizeof(double);
    __shared__ __align__(16) double2 pool_s[WPC == 1 ? NPOOL : 1][32];
    __shared__ __align__(16) double small_s[WPC == 1 ? kSmallDoubles : 2];
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    double2(*const pool)[32] = WPC == 1 ? pool_s : reinterpret_cast<double2(*)[32]>(dyn_smem + (size_t)wib * kWarpBytes);
    double* const small = WPC == 1 ? small_s : reinterpret_cast<double*>(dyn_smem + (size_t)wib * kWarpBytes + (size_t)NPOOL * 32 * sizeof(double2));
    double2(*const sLinv0)[32] = pool;
    double2(*const sLp)[32] = pool + 2 * NL;
    double2(*const sP)[32] = pool + (FULL ? 2 * NL + NF : 0);
    double* const sdiag = small;
    double(*const sz)[DSP] = reinterpret_cast<double(*)[DSP]>(small + 64);
    double* const sw = small + 64 + 3 * DSP;
    // FULL: r_t (the right-hand side row of the coming step), fetched by the same cp.async group as the P tiles
    double* const srow = small + 64 + 4 * DSP;

    const int32_t* mh = p.mhat + c0;
    const double* gv = p.Gv + c0 * D2;
    double2* const Fb = reinterpret_cast<double2*>(p.Lst) + (size_t)c0 * FT * 32 + lane;
    double* Zg = p.Z + c0 * Ds;

    for (int e = lane; e < 3 * DSP; e += 32) (&sz[0][0])[e] = 0.0;
    __syncwarp();
    // power row of the chunk (src/common.jl:60): independent loads, all lanes, before the serial sweeps
    if (p.copy_power)
        for (int e = lane; e < T; e += 32) p.Y[(c0 + e) * p.ldy - 1] = p.Xpow[(c0 + e) * p.ldx - 1];
    // FULL: the rows of Z hold r_t (traj_rhs_kernel) until the forward sweep replaces them with z_t
    bool rok[NT], c0ok[NT], c1ok[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        rok[i] = FULL || 8 * i + r < Ds;
        c0ok[i] = FULL || 8 * i + 2 * q < Ds;
        c1ok[i] = FULL || 8 * i + 2 * q + 1 < Ds;
    }
    // elements [Aoff + 8i + r][Boff + 8j + 2q + {0,1}] of a symmetric P: adjacent in memory when
    // read through the transposed position
    const size_t lofs = (size_t)r * D2 + 2 * q;
    auto ldq = [&](const double* Pm, int Aoff, int Boff, int i, int j) -> double2 {
        double2 o = zero2();
        const double* s = Pm + lofs + (Boff + 8 * j) + (size_t)(Aoff + 8 * i) * D2;
        if (even) {
            if (rok[i] && c0ok[j]) o = __ldg(reinterpret_cast<const double2*>(s));
        } else {
            if (rok[i] && c0ok[j]) o.x = __ldg(s);
            if (rok[i] && c1ok[j]) o.y = __ldg(s + 1);
        }
        return o;
    };

    // mixture indices of frames tn-1, tn, tn+1 come from registers (rolling window, loaded a step ahead)
    auto prefetch = [&](int tn, int ma, int mb, int mc) {
        const double* Pa = p.P + (size_t)ma * D2 * D2 + lofs;
        const double* Pb = p.P + (size_t)mb * D2 * D2 + lofs;
        const double* Pc = p.P + (size_t)mc * D2 * D2 + lofs;
        const size_t dd = (size_t)Ds * D2 + Ds, ds = (size_t)Ds * D2, sd = Ds;
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const size_t o = 8 * j + (size_t)(8 * i) * D2;
                cp_async16(&sP[i * NT + j][lane], Pa + dd + o);
                cp_async16(&sP[NF + i * NT + j][lane], Pa + ds + o);
                cp_async16(&sP[2 * NF + i * NT + j][lane], Pb + sd + o);
                if (j <= i) {
                    cp_async16(&sP[3 * NF + LY::low(i, j)][lane], Pb + o);
                    cp_async16(&sP[3 * NF + NL + LY::low(i, j)][lane], Pc + dd + o);
                }
            }
        if (lane < DSP / 2) cp_async16(&srow[2 * lane], Zg + (size_t)tn * DSP + 2 * lane);      // r_tn
        cp_async_commit();
    };
    // m0 = mhat[t], m1 = mhat[t+1], m2 = mhat[t+2] (clamped to the chunk)
    int m0 = mh[0], m1 = mh[T > 1 ? 1 : 0], m2 = mh[T > 2 ? 2 : T - 1];
    if (FULL) prefetch(0, m0, m0, m1);

    // =========================== forward: block Cholesky + L z = r ===========================
    for (int t = 0; t < T; ++t) {
        const double* Pt = p.P + (size_t)mh[t] * D2 * D2;
        const double* Pm = (t >= 1) ? p.P + (size_t)mh[t - 1] * D2 * D2 : Pt;
        const double* Pp = (t + 1 < T) ? p.P + (size_t)mh[t + 1] * D2 * D2 : Pt;
        const double2(*const Linv1)[32] = sLinv0 + ((t + 1) & 1) * NL;
        const double2(*const Linv2)[32] = sLinv0 + (t & 1) * NL;
        double2(*const LinvT)[32] = sLinv0 + (t & 1) * NL;          // Linv_t replaces Linv_{t-2}
        double* const zt = sz[t % 3];
        const double* const z1 = sz[(t + 2) % 3];
        const double* const z2 = sz[(t + 1) % 3];
        double2* const F = Fb + (size_t)t * FT * 32;
        const int m3 = mh[t + 3 < T ? t + 3 : T - 1];     // consumed by the next step's prefetch
        if (WPC > 1 && p.role_rule == 0 && is_b && t < Tc) pair_sync();       // anti-phase: B enters its product phase when A leaves its own
        double rr[NT];          // r_t = u_t + 1/2 v_{t-1} - 1/2 v_{t+1}, rows 8i + r
        if (!FULL) {
            const double wm = (t >= 1) ? 0.5 : 0.0, wp = (t + 1 < T) ? -0.5 : 0.0;
            const double* g0 = gv + (size_t)t * D2, *gm = gv + (size_t)(t >= 1 ? t - 1 : t) * D2 + Ds;
            const double* gp = gv + (size_t)(t + 1 < T ? t + 1 : t) * D2 + Ds;
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                rr[i] = 0.0;
                if (rok[i]) rr[i] = fma(wp, gp[8 * i + r], fma(wm, gm[8 * i + r], g0[8 * i + r]));
            }
        }
        if (FULL) {
            cp_async_wait_all();
            __syncwarp();       // srow was requested by other lanes
#pragma unroll
            for (int i = 0; i < NT; ++i) rr[i] = srow[8 * i + r];
        }
        auto Pddm = [&](int i, int j) { return FULL ? sP[i * NT + j][lane] : ldq(Pm, Ds, Ds, i, j); };
        auto Pdsm = [&](int i, int j) { return FULL ? sP[NF + i * NT + j][lane] : ldq(Pm, Ds, 0, i, j); };
        auto Psdt = [&](int i, int j) { return FULL ? sP[2 * NF + i * NT + j][lane] : ldq(Pt, 0, Ds, i, j); };
        auto Psst = [&](int i, int j) { return FULL ? sP[3 * NF + LY::low(i, j)][lane] : ldq(Pt, 0, 0, i, j); };
        auto Pddp = [&](int i, int j) { return FULL ? sP[3 * NF + NL + LY::low(i, j)][lane] : ldq(Pp, Ds, Ds, i, j); };

        // ---- R[t][t-2] = -1/4 Pdd_{t-1}
        double2 R2[NF];
#pragma unroll
        for (int e = 0; e < NF; ++e) R2[e] = zero2();
        if (t >= 1) {
#pragma unroll
            for (int i = 0; i < NT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const double2 v = Pddm(i, j);
                    R2[i * NT + j] = make_double2(-0.25 * v.x, -0.25 * v.y);
                }
        }
        // ---- 1. G2 = L[t][t-2] = R2 Linv_{t-2}'
        double2 G2[NF];
#pragma unroll
        for (int e = 0; e < NF; ++e) G2[e] = zero2();
        if (t >= 2) {
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int k = 0; k <= j; ++k) {
                    const double2 y = Linv2[LY::low(j, k)][lane];
#pragma unroll
                    for (int i = 0; i < NT; ++i) mma2(G2[i * NT + j], R2[i * NT + k], y);
                }
        }
        // ---- 2. Tm = R[t][t-1] - G2 L[t-1][t-2]',  R[t][t-1] = 1/2 Pds_{t-1} - 1/2 Psd_t
        double2 Tm[NF];
#pragma unroll
        for (int e = 0; e < NF; ++e) Tm[e] = zero2();
        if (t >= 1) {
#pragma unroll
            for (int i = 0; i < NT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const double2 a = Pdsm(i, j), b = Psdt(i, j);
                    Tm[i * NT + j] = make_double2(0.5 * a.x - 0.5 * b.x, 0.5 * a.y - 0.5 * b.y);
                }
        }
        if (t >= 2) {
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    const double2 ny = neg2(sLp[j * NT + k][lane]);
#pragma unroll
                    for (int i = 0; i < NT; ++i) mma2(Tm[i * NT + j], G2[i * NT + k], ny);
                }
        }
        // ---- 3. G1 = L[t][t-1] = Tm Linv_{t-1}'
        double2 G1[NF];
#pragma unroll
        for (int e = 0; e < NF; ++e) G1[e] = zero2();
        if (t >= 1) {
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int k = 0; k <= j; ++k) {
                    const double2 y = Linv1[LY::low(j, k)][lane];
#pragma unroll
                    for (int i = 0; i < NT; ++i) mma2(G1[i * NT + j], Tm[i * NT + k], y);
                }
        }
        // ---- 4. S = R[t][t] - G2 G2' - G1 G1'   (lower tiles),
        //         R[t][t] = Pss_t + 1/4 Pdd_{t-1} + 1/4 Pdd_{t+1};  1/4 Pdd_{t-1} = -R2
        double2 S[NL];
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                double2 v = Psst(i, j);
                if (FULL) {       // R2 is dead by now: re-read the tile rather than keep it in registers
                    if (t >= 1) {
                        const double2 u = Pddm(i, j);
                        v.x = fma(0.25, u.x, v.x);
                        v.y = fma(0.25, u.y, v.y);
                    }
                } else {
                    v.x -= R2[i * NT + j].x;
                    v.y -= R2[i * NT + j].y;
                }
                if (t + 1 < T) {
                    const double2 u = Pddp(i, j);
                    v.x = fma(0.25, u.x, v.x);
                    v.y = fma(0.25, u.y, v.y);
                }
                if (i == j && !rok[i]) {      // unit diagonal in the padding
                    if (r == 2 * q) v.x = 1.0;
                    if (r == 2 * q + 1) v.y = 1.0;
                }
                S[LY::low(i, j)] = v;
            }
        if (FULL && t + 1 < T) prefetch(t + 1, m0, m1, m2);     // every tile of this step has been consumed
        if (t >= 1) {
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    const double2 n1 = neg2(G1[j * NT + k]), n2 = neg2(G2[j * NT + k]);
#pragma unroll
                    for (int i = j; i < NT; ++i) {
                        mma2(S[LY::low(i, j)], G1[i * NT + k], n1);
                        if (t >= 2) mma2(S[LY::low(i, j)], G2[i * NT + k], n2);
                    }
                }
        }
        if (WPC > 1 && t < Tc) pair_sync();               // end of the product phase (A: lets B start; B: lets A start)
        // ---- 4b. w = r_t - G1 z_{t-1} - G2 z_{t-2};  stream G1, G2 out; G1 becomes L[t][t-1] of
        //          the next step
        {
            double acc[NT];
#pragma unroll
            for (int i = 0; i < NT; ++i) acc[i] = 0.0;
            if (t >= 1) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const double2 a = *reinterpret_cast<const double2*>(z1 + 8 * j + 2 * q);
                    const double2 b = *reinterpret_cast<const double2*>(z2 + 8 * j + 2 * q);
#pragma unroll
                    for (int i = 0; i < NT; ++i) {
                        acc[i] = fma(G1[i * NT + j].x, a.x, acc[i]);
                        acc[i] = fma(G1[i * NT + j].y, a.y, acc[i]);
                        acc[i] = fma(G2[i * NT + j].x, b.x, acc[i]);
                        acc[i] = fma(G2[i * NT + j].y, b.y, acc[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                acc[i] += __shfl_xor_sync(kFull, acc[i], 1);
                acc[i] += __shfl_xor_sync(kFull, acc[i], 2);
                if (q == 0) sw[8 * i + r] = rr[i] - acc[i];
            }
#pragma unroll
            for (int e = 0; e < NF; ++e) {
                __stcs(&F[(NL + e) * 32], G1[e]);      // streamed: 3.7 GB of factors pass through once     // L[t][t-2] is not stored: the back substitution rebuilds its
                                              // product from P (L2-resident) and Linv_{t-2}
                sLp[e][lane] = G1[e];       // all reads of the old L[t-1][t-2] are complete (warp-synchronous)
            }
        }
        // ---- 5. blocked Cholesky of S fused with Linv_t = L^-1
        double2 Li[NL];
        block_cholesky_inverse<NT, COOP>(S, Li, sdiag, lane, p.err);
        // ---- 6. z_t = Linv_t w;  publish Linv_t
        __syncwarp();
        {
            double acc[NT];
#pragma unroll
            for (int i = 0; i < NT; ++i) acc[i] = 0.0;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double2 w = *reinterpret_cast<const double2*>(sw + 8 * j + 2 * q);
#pragma unroll
                for (int i = j; i < NT; ++i) {
                    acc[i] = fma(Li[LY::low(i, j)].x, w.x, acc[i]);
                    acc[i] = fma(Li[LY::low(i, j)].y, w.y, acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                acc[i] += __shfl_xor_sync(kFull, acc[i], 1);
                acc[i] += __shfl_xor_sync(kFull, acc[i], 2);
                if (q == 0) {
                    zt[8 * i + r] = acc[i];
                    if (rok[i]) Zg[(size_t)t * Ds + 8 * i + r] = acc[i];
                }
            }
#pragma unroll
            for (int e = 0; e < NL; ++e) {
                __stcs(&F[e * 32], Li[e]);
                LinvT[e][lane] = Li[e];
            }
        }
        m0 = m1; m1 = m2; m2 = m3;
        __syncwarp();
        if (WPC > 1 && (p.role_rule != 0 || !is_b) && t < Tc) pair_sync();      // end of the factorisation phase (anti-phase: A only; in-phase: both)
    }

    // =========================== backward: L' y = z ===========================================
    // y_t = Linv_t' (z_t - L[t+1][t]' y_{t+1} - L[t+2][t]' y_{t+2}); the transposed products reduce
    // over the row index r (lanes 4, 8, 16 apart) and leave their result in column layout.
    // L[t+2][t] = R[t+2][t] Linv_t' with R[t+2][t] = -1/4 Pdd_{t+1} symmetric, so
    // L[t+2][t]' y_{t+2} = Linv_t (R[t+2][t] y_{t+2}): two small mat-vecs with tiles that are already
    // here (Linv_t) or L2-resident (P) instead of 9 more factor tiles per frame from HBM.
    if constexpr (FULL) {
        static_assert(3 * FT + 1 <= NPOOL, "factor ring does not fit the tile pool");
        back_sweep_ring<NT>(p, pool, NPOOL, &sz[0][0], sdiag, sw, Fb, Zg, mh, c0, T, lane);
        return;
    }
    // padded dimensions: tiles fetched one step ahead into registers
    double* const sy = &sz[0][0];   // ring of three, row-layout reads
#pragma unroll
    for (int e = 0; e < 3 * DSP; e += 32)
        if (e + lane < 3 * DSP) sy[e + lane] = 0.0;
    __syncwarp();
    auto load_step = [&](int t, double2* L, double2* A1, double2* A2, double* zr) {
        const double2* F = Fb + (size_t)t * FT * 32;
#pragma unroll
        for (int e = 0; e < NL; ++e) L[e] = F[e * 32];
#pragma unroll
        for (int e = 0; e < NF; ++e) {
            A1[e] = (t + 1 < T) ? F[(FT + NL + e) * 32] : zero2();
            A2[e] = zero2();
        }
        if (t + 2 < T) {
            const double* Pq = p.P + (size_t)mh[t + 1] * D2 * D2;
#pragma unroll
            for (int i = 0; i < NT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const double2 v = ldq(Pq, Ds, Ds, i, j);
                    A2[i * NT + j] = make_double2(-0.25 * v.x, -0.25 * v.y);
                }
        }
#pragma unroll
        for (int i = 0; i < NT; ++i) zr[i] = rok[i] ? Zg[(size_t)t * Ds + 8 * i + r] : 0.0;
    };
    auto back_step = [&](int t, const double2* L, const double2* A1, const double2* A2, const double* zr) {
        const double* const y1 = sy + ((t + 1) % 3) * DSP;
        const double* const y2 = sy + ((t + 2) % 3) * DSP;
        double* const yt = sy + (t % 3) * DSP;
        // u = R[t+2][t] y_{t+2} (row layout after the reduction over q), then v = Linv_t u
        double v[NT];
        {
            double u[NT];
#pragma unroll
            for (int i = 0; i < NT; ++i) u[i] = 0.0;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double2 yc = *reinterpret_cast<const double2*>(y2 + 8 * j + 2 * q);
#pragma unroll
                for (int i = 0; i < NT; ++i) u[i] = fma(A2[i * NT + j].x, yc.x, fma(A2[i * NT + j].y, yc.y, u[i]));
            }
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                u[i] += __shfl_xor_sync(kFull, u[i], 1);
                u[i] += __shfl_xor_sync(kFull, u[i], 2);
                if (q == 0) sdiag[8 * i + r] = u[i];
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < NT; ++i) v[i] = 0.0;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double2 uc = *reinterpret_cast<const double2*>(sdiag + 8 * j + 2 * q);
#pragma unroll
                for (int i = j; i < NT; ++i) v[i] = fma(L[LY::low(i, j)].x, uc.x, fma(L[LY::low(i, j)].y, uc.y, v[i]));
            }
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                v[i] += __shfl_xor_sync(kFull, v[i], 1);
                v[i] += __shfl_xor_sync(kFull, v[i], 2);
            }
        }
        double2 acc[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[j] = zero2();
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            const double a = y1[8 * i + r];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                acc[j].x = fma(A1[i * NT + j].x, a, acc[j].x);
                acc[j].y = fma(A1[i * NT + j].y, a, acc[j].y);
            }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int s = 4; s < 32; s <<= 1) {
                acc[j].x += __shfl_xor_sync(kFull, acc[j].x, s);
                acc[j].y += __shfl_xor_sync(kFull, acc[j].y, s);
            }
            if (r == 0) *reinterpret_cast<double2*>(sw + 8 * j + 2 * q) = acc[j];
        }
        __syncwarp();
        double2 out[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) out[j] = zero2();
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            const double w = zr[i] - sw[8 * i + r] - v[i];
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                out[j].x = fma(L[LY::low(i, j)].x, w, out[j].x);
                out[j].y = fma(L[LY::low(i, j)].y, w, out[j].y);
            }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int s = 4; s < 32; s <<= 1) {
                out[j].x += __shfl_xor_sync(kFull, out[j].x, s);
                out[j].y += __shfl_xor_sync(kFull, out[j].y, s);
            }
            if (r == 0) {
                *reinterpret_cast<double2*>(yt + 8 * j + 2 * q) = out[j];
                double* yg = p.Y + (c0 + t) * p.ldy + 8 * j + 2 * q;   // reshape(y, D, T)  src/trajectory_gmmmap.jl:109
                if (c0ok[j]) yg[0] = out[j].x;
                if (c1ok[j]) yg[1] = out[j].y;
            }
        }
        __syncwarp();
    };
    {
        double2 La[NL], A1a[NF], A2a[NF], Lb[NL], A1b[NF], A2b[NF];
        double za[NT], zb[NT];
        int t = T - 1;
        load_step(t, La, A1a, A2a, za);
        for (; t >= 1; t -= 2) {
            load_step(t - 1, Lb, A1b, A2b, zb);
            back_step(t, La, A1a, A2a, za);
            if (t >= 2) load_step(t - 2, La, A1a, A2a, za);
            back_step(t - 1, Lb, A1b, A2b, zb);
        }
        if (t == 0) back_step(0, La, A1a, A2a, za);
    }
}

// ------------------------------------------------------------------------------------------------
// Two warps per chunk (Ds == 8 NT): the products and the diagonal-block factorisation of the forward
// sweep run on different warps -- and therefore on different FP64 pipes (DMMA and DFMA share the
// pipe of their SM sub-partition; tools/micro/dmma_bench.cu).
//   warp G  per step t: G2_t, Tm_t (need only step t-2 / t-1 products: they overlap the other warp's
//           factorisation of step t-1) | wait for Linv_{t-1} | G1_t, S_t, w_t -> hands S_t, w_t over;
//   warp C  per step t: Cholesky of S_t fused with Linv_t = L^-1, z_t = Linv_t w_t -> hands Linv_t, z_t back.
// The two hand-overs are named barriers (bar.arrive by the producer, bar.sync by the consumer); S_t
// travels through the shared-memory slot of Linv_{t-2}, which Linv_t then replaces.  Which warp of the
// CTA is G alternates with the hardware warp slot so that every sub-partition of the SM gets the same
// mix of product and factorisation warps.  Warp G runs the back substitution alone.
// ------------------------------------------------------------------------------------------------
constexpr int kBarG2C = 1, kBarC2G = 2;
__device__ __forceinline__ void bar_sync64(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_arrive64(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

template <int NT, bool COOP>
__global__ void __maxnreg__(128) traj_solve_pair(const TrajParams p) {
    using LY = WarpLayout<NT>;
    constexpr int DSP = LY::DSP, NL = LY::NL, NF = LY::NF, FT = LY::FT;
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5, r = lane >> 2, q = lane & 3;
    constexpr int Ds = DSP, D2 = 2 * DSP;
    const int64_t c0 = p.chunk_off[blockIdx.x];
    const int T = (int)(p.chunk_off[blockIdx.x + 1] - c0);
    if (T <= 0) return;

    // tile pool: Linv_{t-1}, Linv_{t-2} (by parity of t; the slot of Linv_{t-2} also carries S_t) |
    // L[t-1][t-2] | G2_t | the P tiles of the step (Pdd_{t-1}, Pds_{t-1}, Psd_t)
    constexpr int NPOOL = 2 * NL + 2 * NF + 3 * NF;
    __shared__ __align__(16) double2 pool[NPOOL][32];
    double2(*const sLinv0)[32] = pool;
    double2(*const sLp)[32] = pool + 2 * NL;
    double2(*const sG2)[32] = pool + 2 * NL + NF;
    double2(*const sP)[32] = pool + 2 * NL + 2 * NF;
    __shared__ __align__(16) double sdiag[64];
    __shared__ __align__(16) double sz[3][DSP];
    __shared__ __align__(16) double sw[DSP];
    __shared__ __align__(16) double srow[DSP];
    __shared__ int s_slot[2];

    const int32_t* mh = p.mhat + c0;
    const double* gv = p.Gv + c0 * D2;
    double2* const Fb = reinterpret_cast<double2*>(p.Lst) + (size_t)c0 * FT * 32 + lane;
    double* Zg = p.Z + c0 * Ds;

    // ---- prologue (both warps): power row (src/common.jl:60), warp slots; the rows of Z hold r_t (traj_rhs_kernel)
    for (int e = tid; e < 3 * DSP; e += 64) (&sz[0][0])[e] = 0.0;
    if (p.copy_power)
        for (int e = tid; e < T; e += 64) p.Y[(c0 + e) * p.ldy - 1] = p.Xpow[(c0 + e) * p.ldx - 1];
    if (lane == 0) {
        unsigned hw;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw));
        s_slot[wib] = (int)hw;
    }
    __syncthreads();
    // consecutive hardware slots sit on consecutive sub-partitions: pairs (0,1) (2,3) (4,5) (6,7) put their
    // G warp on sub-partitions 0, 2, 1, 3
    const int pair_id = min(s_slot[0], s_slot[1]) >> 1;
    bool is_g = (wib == 0) == ((pair_id & 2) == 0);
    // VCB_TRAJ_ROLE: other assignments, kept to reproduce the measurement (1/2: fixed, 3: by CTA parity, 4: by pair parity)
    if (p.role_rule == 1) is_g = wib == 0;
    if (p.role_rule == 2) is_g = wib == 1;
    if (p.role_rule == 3) is_g = (wib == 0) == ((blockIdx.x & 1) == 0);
    if (p.role_rule == 4) is_g = (wib == 0) == ((pair_id & 1) == 0);

    if (!is_g) {
        // =========================== warp C: factorisation ===========================================
        for (int t = 0; t < T; ++t) {
            bar_sync64(kBarG2C);                                   // S_t and w_t are in shared memory
            double2(*const slot)[32] = sLinv0 + (t & 1) * NL;
            double* const zt = sz[t % 3];
            double2* const F = Fb + (size_t)t * FT * 32;
            double2 S[NL], Li[NL];
#pragma unroll
            for (int e = 0; e < NL; ++e) S[e] = slot[e][lane];
            block_cholesky_inverse<NT, COOP>(S, Li, sdiag, lane, p.err);
            // z_t = Linv_t w;  publish Linv_t
            double acc[NT];
#pragma unroll
            for (int i = 0; i < NT; ++i) acc[i] = 0.0;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double2 w = *reinterpret_cast<const double2*>(sw + 8 * j + 2 * q);
#pragma unroll
                for (int i = j; i < NT; ++i) {
                    acc[i] = fma(Li[LY::low(i, j)].x, w.x, acc[i]);
                    acc[i] = fma(Li[LY::low(i, j)].y, w.y, acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                acc[i] += __shfl_xor_sync(kFull, acc[i], 1);
                acc[i] += __shfl_xor_sync(kFull, acc[i], 2);
                if (q == 0) {
                    zt[8 * i + r] = acc[i];
                    Zg[(size_t)t * Ds + 8 * i + r] = acc[i];
                }
            }
#pragma unroll
            for (int e = 0; e < NL; ++e) {
                __stcs(&F[e * 32], Li[e]);
                slot[e][lane] = Li[e];
            }
            __threadfence_block();
            bar_arrive64(kBarC2G);                                 // Linv_t and z_t are in shared memory
        }
        return;
    }

    // =========================== warp G: products, right-hand side ===================================
    const size_t lofs = (size_t)r * D2 + 2 * q;
    auto ldt = [&](const double* Pm, int Aoff, int Boff, int i, int j) -> double2 {
        return __ldg(reinterpret_cast<const double2*>(Pm + lofs + (Boff + 8 * j) + (size_t)(Aoff + 8 * i) * D2));
    };
    auto prefetch = [&](int tn, int ma, int mb) {          // ma = mhat[tn-1] (clamped), mb = mhat[tn]
        const double* Pa = p.P + (size_t)ma * D2 * D2 + lofs;
        const double* Pb = p.P + (size_t)mb * D2 * D2 + lofs;
        const size_t dd = (size_t)Ds * D2 + Ds, ds = (size_t)Ds * D2, sd = Ds;
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const size_t o = 8 * j + (size_t)(8 * i) * D2;
                cp_async16(&sP[i * NT + j][lane], Pa + dd + o);
                cp_async16(&sP[NF + i * NT + j][lane], Pa + ds + o);
                cp_async16(&sP[2 * NF + i * NT + j][lane], Pb + sd + o);
            }
        if (lane < DSP / 2) cp_async16(&srow[2 * lane], Zg + (size_t)tn * DSP + 2 * lane);      // r_tn
        cp_async_commit();
    };
    int m0 = mh[0], m1 = mh[T > 1 ? 1 : 0], m2 = mh[T > 2 ? 2 : T - 1];
    prefetch(0, m0, m0);
    int mprev = m0;

    for (int t = 0; t < T; ++t) {
        const double2(*const Linv1)[32] = sLinv0 + ((t + 1) & 1) * NL;
        const double2(*const Linv2)[32] = sLinv0 + (t & 1) * NL;
        const double* const z1 = sz[(t + 2) % 3];
        const double* const z2 = sz[(t + 1) % 3];
        double2* const F = Fb + (size_t)t * FT * 32;
        const int m3 = mh[t + 3 < T ? t + 3 : T - 1];
        cp_async_wait_all();
        __syncwarp();       // srow was requested by other lanes
        double rr[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) rr[i] = srow[8 * i + r];
        // ---- 1. G2 = L[t][t-2] = R2 Linv_{t-2}',  R2 = R[t][t-2] = -1/4 Pdd_{t-1}
        double2 G2[NF];
#pragma unroll
        for (int e = 0; e < NF; ++e) G2[e] = zero2();
        if (t >= 1) {
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                double2 R2k[NT];
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    const double2 v = sP[i * NT + k][lane];
                    R2k[i] = make_double2(-0.25 * v.x, -0.25 * v.y);
                }
                if (t >= 2) {
#pragma unroll
                    for (int j = k; j < NT; ++j) {
                        const double2 y = Linv2[LY::low(j, k)][lane];
#pragma unroll
                        for (int i = 0; i < NT; ++i) mma2(G2[i * NT + j], R2k[i], y);
                    }
                }
            }
        }
        // ---- 2. Tm = R[t][t-1] - G2 L[t-1][t-2]',  R[t][t-1] = 1/2 Pds_{t-1} - 1/2 Psd_t
        double2 Tm[NF];
#pragma unroll
        for (int e = 0; e < NF; ++e) Tm[e] = zero2();
        if (t >= 1) {
#pragma unroll
            for (int e = 0; e < NF; ++e) {
                const double2 a = sP[NF + e][lane], b = sP[2 * NF + e][lane];
                Tm[e] = make_double2(0.5 * a.x - 0.5 * b.x, 0.5 * a.y - 0.5 * b.y);
            }
        }
        if (t >= 2) {
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    const double2 ny = neg2(sLp[j * NT + k][lane]);
#pragma unroll
                    for (int i = 0; i < NT; ++i) mma2(Tm[i * NT + j], G2[i * NT + k], ny);
                }
        }
        if (t + 1 < T) prefetch(t + 1, m0, m1);     // every staged tile of this step has been consumed
#pragma unroll
        for (int e = 0; e < NF; ++e) sG2[e][lane] = G2[e];      // read back by this lane only
        // R[t][t] = Pss_t + 1/4 Pdd_{t-1} + 1/4 Pdd_{t+1} (lower tiles) straight from L2: the latency passes
        // while this warp waits for the factorisation of step t-1 anyway
        double2 S[NL];
        {
            const double* Pt = p.P + (size_t)m0 * D2 * D2;
            const double* Pm = p.P + (size_t)mprev * D2 * D2;
            const double* Pp = p.P + (size_t)m1 * D2 * D2;
            double2 Sb[NL], Sc[NL];
#pragma unroll
            for (int i = 0; i < NT; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    S[LY::low(i, j)] = ldt(Pt, 0, 0, i, j);
                    Sb[LY::low(i, j)] = (t >= 1) ? ldt(Pm, Ds, Ds, i, j) : zero2();
                    Sc[LY::low(i, j)] = (t + 1 < T) ? ldt(Pp, Ds, Ds, i, j) : zero2();
                }
            // ---- wait for the factorisation of step t-1
            if (t >= 1) bar_sync64(kBarC2G);
#pragma unroll
            for (int e = 0; e < NL; ++e) {
                S[e].x = fma(0.25, Sb[e].x + Sc[e].x, S[e].x);
                S[e].y = fma(0.25, Sb[e].y + Sc[e].y, S[e].y);
            }
        }
        // ---- 3. G1 = L[t][t-1] = Tm Linv_{t-1}'
        double2 G1[NF];
#pragma unroll
        for (int e = 0; e < NF; ++e) G1[e] = zero2();
        if (t >= 1) {
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int k = 0; k <= j; ++k) {
                    const double2 y = Linv1[LY::low(j, k)][lane];
#pragma unroll
                    for (int i = 0; i < NT; ++i) mma2(G1[i * NT + j], Tm[i * NT + k], y);
                }
        }
        // ---- 4. S = R[t][t] - G2 G2' - G1 G1'   (lower tiles)
        if (t >= 1) {
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    const double2 n1 = neg2(G1[j * NT + k]);
#pragma unroll
                    for (int i = j; i < NT; ++i) mma2(S[LY::low(i, j)], G1[i * NT + k], n1);
                    if (t >= 2) {
                        const double2 n2 = neg2(sG2[j * NT + k][lane]);
#pragma unroll
                        for (int i = j; i < NT; ++i) mma2(S[LY::low(i, j)], sG2[i * NT + k][lane], n2);
                    }
                }
        }
        {
            double2(*const slot)[32] = sLinv0 + (t & 1) * NL;      // Linv_{t-2} is dead: its slot carries S_t
#pragma unroll
            for (int e = 0; e < NL; ++e) slot[e][lane] = S[e];
        }
        // ---- 4b. w = r_t - G1 z_{t-1} - G2 z_{t-2}
        {
            double acc[NT];
#pragma unroll
            for (int i = 0; i < NT; ++i) acc[i] = 0.0;
            if (t >= 1) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const double2 a = *reinterpret_cast<const double2*>(z1 + 8 * j + 2 * q);
                    const double2 b = *reinterpret_cast<const double2*>(z2 + 8 * j + 2 * q);
#pragma unroll
                    for (int i = 0; i < NT; ++i) {
                        const double2 g2 = sG2[i * NT + j][lane];
                        acc[i] = fma(G1[i * NT + j].x, a.x, acc[i]);
                        acc[i] = fma(G1[i * NT + j].y, a.y, acc[i]);
                        acc[i] = fma(g2.x, b.x, acc[i]);
                        acc[i] = fma(g2.y, b.y, acc[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                acc[i] += __shfl_xor_sync(kFull, acc[i], 1);
                acc[i] += __shfl_xor_sync(kFull, acc[i], 2);
                if (q == 0) sw[8 * i + r] = rr[i] - acc[i];
            }
        }
        __threadfence_block();
        bar_arrive64(kBarG2C);                                     // S_t and w_t are in shared memory
        // ---- off the chain: stream G1 out (L[t][t-2] is not stored, see the back substitution); it is
        //      L[t][t-1] of the next step; request the R[t+1][t+1] tiles
#pragma unroll
        for (int e = 0; e < NF; ++e) {
            __stcs(&F[(NL + e) * 32], G1[e]);      // streamed: 3.7 GB of factors pass through once
            sLp[e][lane] = G1[e];
        }
        mprev = m0; m0 = m1; m1 = m2; m2 = m3;
    }
    bar_sync64(kBarC2G);           // the last Linv and z are in memory (and every store of warp C is visible)
    static_assert(3 * FT + 1 <= NPOOL, "factor ring does not fit the tile pool");
    back_sweep_ring<NT>(p, pool, NPOOL, &sz[0][0], sdiag, sw, Fb, Zg, mh, c0, T, lane);
}

// r_t = u_t + 1/2 v_{t-1} - 1/2 v_{t+1} (terms outside the chunk vanish) for every frame of every chunk, with
// g_t = P_t E_t = [u_t; v_t]: written into the rows of Z, which the forward sweep of the Ds == 8 NT solvers
// reads one step ahead and then overwrites with z_t.  One CTA per chunk, fully parallel.
__global__ void __launch_bounds__(256) traj_rhs_kernel(const double* __restrict__ Gv, const int64_t* __restrict__ chunk_off,
                                                       double* __restrict__ Z, int Ds) {
    const int64_t c0 = chunk_off[blockIdx.x];
    const int T = (int)(chunk_off[blockIdx.x + 1] - c0);
    const int D2 = 2 * Ds;
    const double* gv = Gv + c0 * D2;
    double* z = Z + c0 * Ds;
    for (int e = threadIdx.x; e < T * Ds; e += blockDim.x) {
        const int t = e / Ds, k = e - t * Ds;
        double v = gv[(size_t)t * D2 + k];
        if (t >= 1) v = fma(0.5, gv[(size_t)(t - 1) * D2 + Ds + k], v);
        if (t + 1 < T) v = fma(-0.5, gv[(size_t)(t + 1) * D2 + Ds + k], v);
        z[e] = v;
    }
}

template <int NT>
int32_t launch_warp(const TrajParams& p, int64_t nchunks, cudaStream_t st) {
    static const int occ = [] { const char* e = getenv("VCB_TRAJ_MINB"); return e ? atoi(e) : 0; }();
    if (NT == 3 && occ) {      // occupancy experiment: direct P loads (11.5 KB shared memory), capped registers
        auto k = occ >= 16 ? traj_solve_warp<NT, false, 16> : occ >= 12 ? traj_solve_warp<NT, false, 12> : traj_solve_warp<NT, false, 8>;
        VCB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        k<<<(unsigned)nchunks, 32, 0, st>>>(p);
        count_launch();
        VCB_CUDA(cudaGetLastError());
        return VCB_OK;
    }
    // VCB_TRAJ_PAIR: 0 = one warp per chunk, 1 = two warps (redundant diagonal tiles), 2 = two warps (cooperative)
    static const int pair = [] { const char* e = getenv("VCB_TRAJ_PAIR"); return e ? atoi(e) : 0; }();
    static const int role_rule = [] { const char* e = getenv("VCB_TRAJ_ROLE"); return e ? atoi(e) : 0; }();
    if (p.Ds == 8 * NT && !(NT == 3 && occ)) {
        traj_rhs_kernel<<<(unsigned)nchunks, 256, 0, st>>>(p.Gv, p.chunk_off, p.Z, p.Ds);
        count_launch();
    }
    if (p.Ds == 8 * NT && pair) {
        TrajParams p2 = p;
        p2.role_rule = role_rule;
        auto k = pair == 2 ? traj_solve_pair<NT, true> : traj_solve_pair<NT, false>;
        VCB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        static const bool dbg = [&] {
            if (!getenv("VCB_TRAJ_DEBUG")) return false;
            int nb = 0;
            cudaFuncAttributes fa{};
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, 64, 0);
            cudaFuncGetAttributes(&fa, k);
            fprintf(stderr, "[vcb] traj_solve_pair: %d CTAs/SM, %d regs, %zu B static smem\n", nb, fa.numRegs, fa.sharedSizeBytes);
            return true;
        }();
        (void)dbg;
        k<<<(unsigned)nchunks, 64, 0, st>>>(p2);
        count_launch();
        VCB_CUDA(cudaGetLastError());
        return VCB_OK;
    }
    // VCB_TRAJ_WPC=7: one CTA of seven chunks per SM with phase-locked warp pairs (see traj_solve_warp)
    static const int wpc = [] { const char* e = getenv("VCB_TRAJ_WPC"); return e ? atoi(e) : 1; }();
    if (p.Ds == 8 * NT && wpc == 7) {
        using LY = WarpLayout<NT>;
        constexpr int NPOOL = 4 * LY::NL + 4 * LY::NF, kSmall = 64 + 5 * LY::DSP;
        constexpr size_t kWarpBytes = (size_t)NPOOL * 512 + (size_t)((kSmall + 1) & ~1) * sizeof(double);
        auto k = traj_solve_warp<NT, true, 1, false, 7>;
        VCB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(7 * kWarpBytes)));
        static const int phase = [] { const char* e = getenv("VCB_TRAJ_PHASE"); return e ? atoi(e) : 0; }();
        TrajParams p3 = p;
        p3.role_rule = phase;        // 0: partners in opposite phases, 1: partners in the same phase
        k<<<(unsigned)((nchunks + 6) / 7), 224, 7 * kWarpBytes, st>>>(p3);
        count_launch();
        VCB_CUDA(cudaGetLastError());
        return VCB_OK;
    }
    if (p.Ds == 8 * NT) {
        auto k = traj_solve_warp<NT, true>;
        VCB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        k<<<(unsigned)nchunks, 32, 0, st>>>(p);
    } else {
        traj_solve_warp<NT, false><<<(unsigned)nchunks, 32, 0, st>>>(p);
    }
    count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

}  // namespace

size_t traj_warp_factor_bytes(int Ds) {
    const int nt = (Ds + 7) / 8;
    if (nt < 1 || nt > 3) return 0;
    return (size_t)(nt * (nt + 1) / 2 + nt * nt) * 32 * sizeof(double2);
}

int32_t traj_warp_launch(const TrajParams& p, int64_t nchunks, cudaStream_t st) {
    switch ((p.Ds + 7) / 8) {
        case 1: return launch_warp<1>(p, nchunks, st);
        case 2: return launch_warp<2>(p, nchunks, st);
        case 3: return launch_warp<3>(p, nchunks, st);
        default: return fail(VCB_EUNSUPPORTED, "warp trajectory solver covers static dimension <= 24 (got %d)", p.Ds);
    }
}

}  // namespace vcb
