// vcb_dtw.cu -- K4: batched DTW, bit-exact with DTWs.fit! + backward (reference src/dtw.jl:93-145).
//
// The reference recurrence is column-recursive: column t+1 of the cost table depends only on
// column t (src/dtw.jl:113-121), so all template states of a column are independent.  One CTA
// aligns one (template, sequence) pair with one thread per template state:
//
//   * observation costs (src/dtw.jl:33-35, sum_k (v_k - tmpl_k)^2 in strict left-to-right Float64
//     without FMA contraction) are computed for a tile of TT sequence frames at once -- this is
//     the FP64-pipe bound part and has TT independent dependency chains per thread; the costs of
//     tile n+1 are computed piecewise inside the column loop of tile n (software pipeline), so
//     the FP64 pipe keeps working across the per-column barriers;
//   * the TT column updates run on a double-buffered shared-memory cost column (one
//     __syncthreads per column); candidates are visited in the reference order
//     i, i-bstep, ..., i+fstep with a strict `<`, sums associate as ((cost + ocost) + transition);
//   * back-pointers are stored as (j - i + bstep) in BITS bits, packed 32/BITS columns per word in
//     an L2-resident scratch (0.25 B/cell for the usual windows); the local-cost matrix and the
//     cost table never exist in memory;
//   * warp 0 back-tracks from the first minimum of the last column (src/dtw.jl:137), fetching
//     back-pointer words 32 states at a time.
//
// The template is read through a k-major ("transposed") copy so the per-k loads are coalesced.
#include <cstdlib>
#include <type_traits>
#include <utility>

#include "vcb_kernels.h"

namespace vcb {

// 1024 threads x 8 states; 41 s of speech at the reference's 5 ms frame shift
constexpr int kDtwMaxStates = 8192;

// tmpl (D, S) column-major per pair  ->  tmplT[k * S + i]
__global__ void dtw_transpose_kernel(const double* __restrict__ tmpl, const int64_t* __restrict__ toff,
                                     double* __restrict__ tmplT, int D) {
    const int p = blockIdx.x;
    const int64_t b = toff[p];
    const int S = (int)(toff[p + 1] - b);
    __shared__ double tile[32][33];
    const int i0 = blockIdx.y * 32;
    if (i0 >= S) return;
    for (int k0 = 0; k0 < D; k0 += 32) {
        // read: consecutive threads walk k (contiguous in the source)
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            int i = i0 + r, k = k0 + threadIdx.x;
            if (i < S && k < D) tile[r][threadIdx.x] = tmpl[(b + i) * D + k];
        }
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            int k = k0 + r, i = i0 + threadIdx.x;
            if (i < S && k < D) tmplT[b * D + (int64_t)k * S + i] = tile[threadIdx.x][r];
        }
        __syncthreads();
    }
}

// DT > 0: feature dimension known at compile time (observation loop fully unrolled);
// BS >= 0: window (bstep = BS, fstep = FS) known at compile time (candidate scan unrolled);
// SPT: consecutive template states owned by one thread (thread i: states i*SPT .. i*SPT+SPT-1), so a
// CTA of <= 1024 threads covers templates of up to 1024*SPT frames and a thread exchanges only its
// window edges with its neighbours.
template <int BITS, int TT, int MAXT, int MINB, int DT, int BS, int FS, int SPT>
__global__ void __launch_bounds__(MAXT, MINB)
dtw_fused_kernel(const double* __restrict__ tmplT, const int64_t* __restrict__ toff,
                 const double* __restrict__ seq, const int64_t* __restrict__ soff,
                 const int64_t* __restrict__ bpoff, const int64_t* __restrict__ order, uint32_t* __restrict__ bp,
                 int Drt, int fstep_rt, int bstep_rt, int64_t* __restrict__ paths, double* __restrict__ final_cost) {
    constexpr int PER = 32 / BITS;
    constexpr uint32_t MASK = (BITS == 32) ? 0xFFFFFFFFu : ((1u << BITS) - 1u);
    static_assert(TT % 2 == 0, "sequence tiles are read as double2");
    const int D = DT > 0 ? DT : Drt;
    const int bstep = BS >= 0 ? BS : bstep_rt, fstep = BS >= 0 ? FS : fstep_rt;
    const int p = (int)order[blockIdx.x];      // pairs are launched by decreasing cost (the long ones first)
    const int64_t tb = toff[p], sb = soff[p];
    const int S = (int)(toff[p + 1] - tb);
    const int T = (int)(soff[p + 1] - sb);
    const int s0 = threadIdx.x * SPT;            // first state of this thread
    const int Spad = (S + 31) & ~31;
    uint32_t* bpp = bp + bpoff[p];  // [ceil(T/PER)][Spad]

    // cost columns carry `bstep` sentinels (+inf) on the left and `fstep` on the right, so the
    // candidate scan needs no range checks: an infinite candidate never passes the strict `<`
    extern __shared__ double smem[];
    const int colw = (bstep + (int)blockDim.x * SPT + fstep + 1) & ~1;
    double* cur = smem + bstep;                  // [-bstep, blockDim*SPT + fstep)
    double* nxt = smem + colw + bstep;
    double* vt = smem + 2 * colw;                // [D][TT]  sequence tile, k-major (even offset: 16-byte aligned)
    __shared__ double red_v[32];
    __shared__ int red_i[32];
    __shared__ int s_best;
    const double kInf = __longlong_as_double(0x7FF0000000000000LL);

    const double* tcol = tmplT + tb * D + s0;  // element (k, state s0 + j) at tcol[k * S + j]
    for (int e = threadIdx.x; e < 2 * colw; e += blockDim.x) smem[e] = kInf;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SPT; ++j)
        if (s0 + j < S) cur[s0 + j] = (double)(s0 + j + 1);      // src/dtw.jl:49  costtable[:,1] = 1:S
    uint32_t word[SPT];
#pragma unroll
    for (int j = 0; j < SPT; ++j) word[j] = 0;

    for (int t0 = 0; t0 < T; t0 += TT) {
        const int ncols = min(TT, T - t0);
        __syncthreads();  // previous tile's readers of vt are done; column init visible
        for (int e = threadIdx.x; e < TT * D; e += blockDim.x) {
            const int c = e / D, k = e - c * D;
            vt[k * TT + c] = (c < ncols) ? seq[(sb + t0 + c) * D + k] : 0.0;
        }
        __syncthreads();

        // ---- observation costs for TT frames: acc[j][c] = sum_k (v[c][k] - tmpl[k][s0+j])^2 with k
        //      in ascending order, no FMA (src/dtw.jl:33-35)
        double acc[SPT][TT];
#pragma unroll
        for (int j = 0; j < SPT; ++j)
#pragma unroll
            for (int c = 0; c < TT; ++c) acc[j][c] = 0.0;
        if (s0 < S) {
            const double* tp = tcol;
            auto kstep = [&](int k) {
                double tk[SPT];
#pragma unroll
                for (int j = 0; j < SPT; ++j) tk[j] = (SPT == 1 || s0 + j < S) ? tp[j] : 0.0;
                tp += S;
                const double2* v2 = reinterpret_cast<const double2*>(vt + k * TT);
#pragma unroll
                for (int c = 0; c < TT; c += 2) {
                    const double2 v = v2[c >> 1];
#pragma unroll
                    for (int j = 0; j < SPT; ++j) {
                        const double d0 = __dsub_rn(v.x, tk[j]), d1 = __dsub_rn(v.y, tk[j]);
                        acc[j][c] = __dadd_rn(acc[j][c], __dmul_rn(d0, d0));
                        acc[j][c + 1] = __dadd_rn(acc[j][c + 1], __dmul_rn(d1, d1));
                    }
                }
            };
            if (DT > 0) {
#pragma unroll 4
                for (int k = 0; k < DT; ++k) kstep(k);
            } else {
                for (int k = 0; k < D; ++k) kstep(k);
            }
        }

        // ---- column recurrence  (src/dtw.jl:104-125): candidates i (stay), then i-bstep .. i+fstep,
        //      strict `<`; ((cost + ocost) + transition); transition 0 for j = i-1 (adding +0.0 to a
        //      non-negative sum is the identity), 1 for j = i, 2 otherwise (src/dtw.jl:23-31)
#pragma unroll
        for (int c = 0; c < TT; ++c) {
            if (c < ncols) {
                const int t = t0 + c;
                if constexpr (BS >= 0) {
                    // the thread's window of the previous column, read once
                    constexpr int WN = SPT + BS + FS;
                    double win[WN];
#pragma unroll
                    for (int e = 0; e < WN; ++e) win[e] = cur[s0 - BS + e];
#pragma unroll
                    for (int j = 0; j < SPT; ++j) {
                        if (s0 + j < S) {
                            const double oc = acc[j][c];
                            int code = BS;  // minindex - i + bstep
                            double minc = __dadd_rn(__dadd_rn(win[BS + j], oc), 1.0);
#pragma unroll
                            for (int dj = -BS; dj <= FS; ++dj) {
                                if (dj == 0) continue;  // same value as the initial candidate: never `<`
                                double cand = __dadd_rn(win[BS + j + dj], oc);
                                if (dj != -1) cand = __dadd_rn(cand, 2.0);
                                if (cand < minc) { minc = cand; code = dj + BS; }
                            }
                            nxt[s0 + j] = minc;
                            word[j] |= (uint32_t)code << (BITS * (t % PER));
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < SPT; ++j) {
                        const int i = s0 + j;
                        if (i < S) {
                            const double oc = acc[j][c];
                            int code = bstep;
                            double minc = __dadd_rn(__dadd_rn(cur[i], oc), 1.0);
                            for (int dj = -bstep; dj <= fstep; ++dj) {
                                double cand = __dadd_rn(cur[i + dj], oc);
                                if (dj != -1) cand = __dadd_rn(cand, dj == 0 ? 1.0 : 2.0);
                                if (cand < minc) { minc = cand; code = dj + bstep; }
                            }
                            nxt[i] = minc;
                            word[j] |= (uint32_t)code << (BITS * (t % PER));
                        }
                    }
                }
                if ((t % PER) == PER - 1 || t == T - 1) {
#pragma unroll
                    for (int j = 0; j < SPT; ++j) {
                        if (s0 + j < Spad) bpp[(int64_t)(t / PER) * Spad + s0 + j] = word[j];
                        word[j] = 0;
                    }
                }
                __syncthreads();
                double* tmp = cur; cur = nxt; nxt = tmp;
            }
        }
    }

    // ---- indmin(costtable[:, T+1]) -- first minimum  (src/dtw.jl:137)
    {
        double v = kInf;
        int idx = 0x7FFFFFFF;
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
            if (s0 + j < S) {
                double vj = cur[s0 + j];
                // NaN never wins a `<` in the reference's scan unless it is first; keep it simple: treat
                // NaN as +inf except at index 0 (cannot occur for finite inputs).
                if (vj != vj && s0 + j != 0) vj = kInf;
                if (vj < v || idx == 0x7FFFFFFF) { v = vj; idx = s0 + j; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_down_sync(0xFFFFFFFFu, v, o);
            int oi = __shfl_down_sync(0xFFFFFFFFu, idx, o);
            if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        if (l == 0) { red_v[w] = v; red_i[w] = idx; }
        __syncthreads();
        if (w == 0) {
            const int nw = (blockDim.x + 31) >> 5;
            v = (l < nw) ? red_v[l] : kInf;
            idx = (l < nw) ? red_i[l] : 0x7FFFFFFF;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ov = __shfl_down_sync(0xFFFFFFFFu, v, o);
                int oi = __shfl_down_sync(0xFFFFFFFFu, idx, o);
                if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
            }
            if (l == 0) {
                s_best = idx;
                if (final_cost) final_cost[p] = v;
            }
        }
        __syncthreads();
    }

    // ---- backward  (src/dtw.jl:139-142); back-pointer stores above are visible after the barrier
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int st = s_best;
        int64_t* path = paths + sb;
        if (lane == 0) path[T - 1] = st + 1;
        int cur_wi = -1, wbase = 0;
        uint32_t w = 0;
        for (int t = T - 1; t >= 1; --t) {
            const int wi = t / PER;
            if (wi != cur_wi || st < wbase || st >= wbase + 32) {
                wbase = (fstep == 0) ? st - 31 : st - 16;
                wbase = max(0, min(wbase, Spad - 32));
                w = bpp[(int64_t)wi * Spad + wbase + lane];
                cur_wi = wi;
            }
            const uint32_t ww = __shfl_sync(0xFFFFFFFFu, w, st - wbase);
            const int code = (int)((ww >> (BITS * (t % PER))) & MASK);
            st = st + code - bstep;
            if (lane == 0) path[t - 1] = st + 1;
        }
    }
}

template <int BITS, int TT, int MAXT, int MINB, int DT, int BS, int FS, int SPT>
static int32_t launch_dtw_cfg(const double* tmplT, const int64_t* d_toff, const double* seq,
                              const int64_t* d_soff, const int64_t* d_bpoff, const int64_t* d_order, uint32_t* bp, int D,
                              int fstep, int bstep, int64_t npairs, int maxS, int64_t* paths,
                              double* final_cost, cudaStream_t st) {
    const int nt = round_up((maxS + SPT - 1) / SPT, 32);
    const int colw = round_up(bstep + nt * SPT + fstep, 2);
    const size_t smem = (size_t)(2 * colw + D * TT) * sizeof(double);
    auto k = dtw_fused_kernel<BITS, TT, MAXT, MINB, DT, BS, FS, SPT>;
    VCB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)npairs, nt, smem, st>>>(tmplT, d_toff, seq, d_soff, d_bpoff, d_order, bp, D, fstep, bstep, paths, final_cost);
    count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

// ------------------------------------------------------------------------------------------------
// STREAM kernel -- the default for the reference's own windows (fstep = 0, bstep = 1 or 2), D = 24 / 40 and
// templates whose warp slices fit in shared memory.  Dependencies then run one way only, from lower to
// higher template states, so the warps of a CTA form a PIPELINE whose unit of hand-over is a TILE of TT
// columns: warp w owns states 32w .. 32w+31 (one per lane, cost in a register); inside a tile the two
// neighbours below come by shuffle, and lanes 0/1 take them from a shared-memory ring in which the warp
// below publishes the costs of its last two states for every column of a tile, followed by ONE mbarrier
// arrival.  Warp w starts the recurrence of tile n when warp w-1 has finished it and computes the
// observation costs of the tile -- the FP64-bound part, which depends on nobody -- before it asks, so the
// warps sit one recurrence apart and the FP64 pipe always finds warps in their observation phase; there is
// no CTA-wide barrier to drain it (the per-column __syncthreads of dtw_fused_kernel is its largest stall).
// The kernel is PERSISTENT: one CTA per SM walks a list of pairs (balanced on the host by decreasing cost),
// and a warp that has finished the last tile of one pair starts on the next pair at once, so the pipeline
// fills and drains once per LAUNCH, not once per pair.  Hand-overs are mbarriers: per warp interface a ring
// of RT "full" (tile published) and RT "empty" (tile consumed) barriers.  Sequence tiles arrive through a
// private cp.async double buffer one tile ahead; the warp's 32 template frames sit in shared memory for
// the whole pair, copied straight from the caller's (D, S) matrix (no transposed copy on this path).  The
// final-minimum search and the back-tracking of pair q run in a SERVICE warp while the compute warps are
// already in pair q+1 (final costs double-buffered in shared memory).  Bit-exact: same candidate order,
// same strict `<`, same left-to-right sums.
// History (round 2, all bit-exact, C3 = 1000 pairs of ~600 x 600): hand-over per COLUMN through flags
// 4.41 ms; per tile through flags polled with __nanosleep, one CTA per pair 6.0 ms (a 21-warp CTA at 48
// registers does not fit twice on an SM -- six warps land on one sub-partition -- and the sleeps overshoot);
// persistent + mbarriers + template by global loads 2.80 ms; template slices in shared memory 2.49 ms
// (barrier kernel 2.72 ms).  What remains is the integer number of warps per SM sub-partition: a pair with
// 21 (18) active warps loads the four FP64 pipes 6/5/5/5 (5/5/4/4) and the chain runs at the pace of the
// fullest one (17 % of the warp time is spent waiting for the warp below; a one-tile lookahead with
// 4-column tiles did not change that and cost 8 % in hand-overs).
// ------------------------------------------------------------------------------------------------
// f(integral_constant<int, 0>) ... f(integral_constant<int, N-1>): loop indices usable as asm immediates
template <class F, int... I>
__device__ __forceinline__ void static_for(F&& f, std::integer_sequence<int, I...>) { (f(std::integral_constant<int, I>{}), ...); }

__device__ __forceinline__ void dtw_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void dtw_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void dtw_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "DTW_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DTW_DONE;\n"
        "bra DTW_WAIT;\n"
        "DTW_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

template <int BITS, int TT, int MAXT, int DT, int BS>
__global__ void __launch_bounds__(MAXT, 1)
dtw_stream_kernel(const double* __restrict__ tmpl, const int64_t* __restrict__ toff, const double* __restrict__ seq,
                  const int64_t* __restrict__ soff, const int64_t* __restrict__ bpoff, const int32_t* __restrict__ cta_first,
                  const int32_t* __restrict__ cta_list, uint32_t* __restrict__ bp, int64_t* __restrict__ paths,
                  double* __restrict__ final_cost) {
    constexpr int PER = 32 / BITS;
    constexpr int RT = 4, R = RT * TT;           // ring: RT tiles of TT columns a warp may run ahead of its consumer
    // ring[w][pos][4], pos = 1..R: the costs of the warp's last two states ENTERING column t sit at position
    // ((t' - 1) mod R) + 1 (t' counts the columns of the warp interface over all pairs) as [b1, b0, b1, b1]
    // (b0: lane 30, b1: lane 31), so that lane 0 of the warp above reads (c1, c2) = (b1, b0) and lane 1 reads
    // (-, c2) = (b1, b1) with one 16-byte load each; the positions a tile WRITES are contiguous and so are the
    // ones it READS, except the first.
    constexpr int RS = 4 * (R + 1);              // doubles per warp: positions 1..R of 4 doubles (position 0 unused)
    constexpr int WPT = PER / TT;                // tiles per back-pointer word
    constexpr uint32_t MASK = (1u << BITS) - 1u;
    constexpr int D = DT;
    static_assert(DT % 4 == 0 && (BS == 1 || BS == 2) && PER % TT == 0, "stream kernel: dimension 4n, bstep 1 or 2, fstep 0");
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nwc = ((int)blockDim.x >> 5) - 1;  // compute warps; warp nwc is the service warp
    const int nst = nwc * 32;
    const double kInf = __longlong_as_double(0x7FF0000000000000LL);

    extern __shared__ double smem[];
    double* ring = smem;                                   // [nwc][RS]
    double* tiles = ring + (size_t)nwc * RS;               // [nwc][2][TT][D] private double-buffered sequence tiles
    double* tsl = tiles + (size_t)nwc * 2 * TT * D;        // [nwc][D/2][32][2] the warp's states of the current template
    double* fin = tsl + (size_t)nwc * 32 * D;              // [2][nst] final cost columns of pairs q, q+1
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(fin + 2 * (size_t)nst);
    // full[nwc][RT] | empty[nwc][RT] | pairdone[2] | finfree[2]
    const uint32_t full0 = bar0, empty0 = bar0 + 8u * nwc * RT, pairdone0 = bar0 + 16u * nwc * RT, finfree0 = pairdone0 + 16u;
    if (threadIdx.x == 0) {
        for (int e = 0; e < 2 * nwc * RT; ++e) dtw_mbar_init(bar0 + 8u * e, 1);
        dtw_mbar_init(pairdone0, nwc);
        dtw_mbar_init(pairdone0 + 8u, nwc);
        dtw_mbar_init(finfree0, 1);
        dtw_mbar_init(finfree0 + 8u, 1);
    }
    __syncthreads();
    const int lbeg = cta_first[blockIdx.x], npl = cta_first[blockIdx.x + 1] - lbeg;

    if (w == nwc) {
        // ================= service warp: indmin + backward of finished pairs (src/dtw.jl:137-142) =================
        for (int q = 0; q < npl; ++q) {
            const int p = cta_list[lbeg + q];
            const int64_t sb = soff[p];
            const int S = (int)(toff[p + 1] - toff[p]);
            const int T = (int)(soff[p + 1] - sb);
            const int Spad = (S + 31) & ~31;
            const uint32_t* bpp = bp + bpoff[p];
            dtw_mbar_wait(pairdone0 + 8u * (q & 1), (q >> 1) & 1);
            const double* f = fin + (size_t)(q & 1) * nst;
            double v = kInf;
            int idx = 0x7FFFFFFF;
            for (int j = lane; j < S; j += 32) {
                double vj = f[j];
                if (vj != vj && j != 0) vj = kInf;       // see dtw_fused_kernel
                if (vj < v || idx == 0x7FFFFFFF) { v = vj; idx = j; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ov = __shfl_down_sync(0xFFFFFFFFu, v, o);
                int oi = __shfl_down_sync(0xFFFFFFFFu, idx, o);
                if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
            }
            int st = __shfl_sync(0xFFFFFFFFu, idx, 0);
            if (lane == 0 && final_cost) final_cost[p] = v;
            __syncwarp();
            if (lane == 0) dtw_mbar_arrive(finfree0 + 8u * (q & 1));      // the final costs are consumed
            int64_t* path = paths + sb;
            if (lane == 0) path[T - 1] = st + 1;
            int cur_wi = -1, wbase = 0;
            uint32_t wd = 0;
            for (int t = T - 1; t >= 1; --t) {
                const int wi = t / PER;
                if (wi != cur_wi || st < wbase || st >= wbase + 32) {
                    wbase = max(0, min(st - 31, Spad - 32));
                    wd = bpp[(int64_t)wi * Spad + wbase + lane];
                    cur_wi = wi;
                }
                const uint32_t ww = __shfl_sync(0xFFFFFFFFu, wd, st - wbase);
                const int code = (int)((ww >> (BITS * (t % PER))) & MASK);
                st = st + code - BS;
                if (lane == 0) path[t - 1] = st + 1;
            }
        }
        return;
    }

    // ======================================= compute warps =======================================
    const int i = w * 32 + lane;
    uint32_t g_in = 0, g_out = 0;                // tiles taken from the warp below / handed to the warp above so far
    uint32_t rp32 = (uint32_t)__cvta_generic_to_shared(ring + (size_t)(w > 0 ? w - 1 : 0) * RS + 2 * lane);    // lanes 0, 1 read
    uint32_t wp32 = (uint32_t)__cvta_generic_to_shared(ring + (size_t)w * RS + (lane == 30 ? 1 : 0));
    asm volatile("" : "+r"(rp32), "+r"(wp32));
    double* const mytiles = tiles + (size_t)w * 2 * TT * D;
    const double2* const mytmpl = reinterpret_cast<const double2*>(tsl + (size_t)w * 32 * D) + lane;     // dims 2j, 2j+1 at [32 j]
    for (int q = 0; q < npl; ++q) {
        const int p = cta_list[lbeg + q];
        const int64_t tb = toff[p], sb = soff[p];
        const int S = (int)(toff[p + 1] - tb);
        const int T = (int)(soff[p + 1] - sb);
        if (q >= 2) dtw_mbar_wait(finfree0 + 8u * (q & 1), ((q >> 1) - 1) & 1);     // the service warp is done with pair q-2
        if (32 * w >= S) {                       // no state of this pair in this warp
            if (lane == 0) dtw_mbar_arrive(pairdone0 + 8u * (q & 1));
            continue;
        }
        const bool below = w > 0, above = 32 * (w + 1) < S;
        const bool rd = below && lane < 2, pub = above && lane >= 30, pub31 = above && lane == 31;
        const int Spad = (S + 31) & ~31;
        uint32_t* const bpp = bp + bpoff[p] + i;
        {   // this warp's 32 template frames, straight from the caller's (D, S) matrix (lanes past the template
            // repeat its last state); lands with the first sequence tile
            const double* src = tmpl + (tb + min(i, S - 1)) * D;
#pragma unroll
            for (int j = 0; j < D / 2; ++j) {
                const unsigned a = (unsigned)__cvta_generic_to_shared(mytmpl + 32 * j);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(src + 2 * j) : "memory");
            }
        }
        const double* const srow = seq + sb * D;
        double c = (double)(i + 1);              // src/dtw.jl:49  costtable[:,1] = 1:S
        uint32_t word = 0;
        auto fetch_tile = [&](int t0n, int buf) {
            if (t0n < T) {
                const int pieces = min(TT, T - t0n) * (D / 2);
                const double* src = srow + (size_t)t0n * D;
                for (int e = lane; e < pieces; e += 32) {
                    const unsigned a = (unsigned)__cvta_generic_to_shared(mytiles + (size_t)buf * TT * D + 2 * e);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(src + 2 * e) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto tile_body = [&](auto full_tag, const int t0, const int tile) {
            constexpr bool FULL = decltype(full_tag)::value;
            const int ncols = FULL ? TT : T - t0;
            fetch_tile(t0 + TT, (tile + 1) & 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
            // ---- observation costs for TT frames: strict left-to-right Float64 sum, no FMA (src/dtw.jl:33-35)
            double acc[TT];
#pragma unroll
            for (int cc = 0; cc < TT; ++cc) acc[cc] = 0.0;
            {
                // two dimensions at a time, all TT columns side by side: every stage (subtract, square, add) is
                // TT or 2 TT independent operations, so the FP64 pipe never waits for a dependent result
                const double2* vb = reinterpret_cast<const double2*>(mytiles + (size_t)(tile & 1) * TT * D);
                double2 tk = mytmpl[0];
#pragma unroll
                for (int k2 = 0; k2 < D / 2; ++k2) {
                    double2 tn = tk;
                    if (k2 + 1 < D / 2) tn = mytmpl[32 * (k2 + 1)];     // a step ahead
                    double2 v[TT];
#pragma unroll
                    for (int cc = 0; cc < TT; ++cc) v[cc] = vb[cc * (D / 2) + k2];      // columns past the sequence: stale data, unused
#pragma unroll
                    for (int cc = 0; cc < TT; ++cc) { v[cc].x = __dsub_rn(v[cc].x, tk.x); v[cc].y = __dsub_rn(v[cc].y, tk.y); }
#pragma unroll
                    for (int cc = 0; cc < TT; ++cc) { v[cc].x = __dmul_rn(v[cc].x, v[cc].x); v[cc].y = __dmul_rn(v[cc].y, v[cc].y); }
#pragma unroll
                    for (int cc = 0; cc < TT; ++cc) acc[cc] = k2 == 0 ? v[cc].x : __dadd_rn(acc[cc], v[cc].x);   // (0.0 + d^2 is d^2: a square is never -0.0)
#pragma unroll
                    for (int cc = 0; cc < TT; ++cc) acc[cc] = __dadd_rn(acc[cc], v[cc].y);
                    tk = tn;
                }
            }
            // ---- hand-over, once per tile: the warp below has published the costs entering every column of
            //      this tile; the warp above has consumed the ring positions this tile overwrites
            const uint32_t sin = g_in & (RT - 1), sout = g_out & (RT - 1);
            if (below) dtw_mbar_wait(full0 + 8u * ((w - 1) * RT + sin), (g_in / RT) & 1);
            if (above && g_out >= RT) dtw_mbar_wait(empty0 + 8u * (w * RT + sout), (g_out / RT - 1) & 1);
            // (c1, c2) of lanes 0 / 1 entering the first column: the initial column of a pair is known
            // (src/dtw.jl:49), later ones sit in the last position of the previous ring stage
            double2 xn = make_double2(kInf, kInf);
            if (below) {
                if (tile == 0) xn = make_double2((double)(32 * w), lane == 0 ? (double)(32 * w - 1) : (double)(32 * w));
                else if (rd) {
                    const unsigned r0 = rp32 + 32u * (sin == 0 ? R : TT * sin);
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xn.x), "=d"(xn.y) : "r"(r0));
                }
                if (g_in >= 1) {                 // the previous stage of the ring is consumed
                    __syncwarp();
                    if (lane == 0) dtw_mbar_arrive(empty0 + 8u * ((w - 1) * RT + ((g_in - 1) & (RT - 1))));
                }
            }
            // ---- column recurrence (src/dtw.jl:104-125): candidates i (stay), then i-bstep .. i-1, strict `<`.
            //      Lanes past the template (last warp only) run along on finite garbage: costs only travel upwards.
            unsigned rp = rp32 + 32u * (TT * sin);                                   // entering t0 + cc at rp + 32 cc, cc >= 1
            unsigned wp = wp32 + 32u * (TT * sout + 1);                              // entering t0 + 1 + cc at wp + 32 cc
            asm volatile("" : "+r"(rp), "+r"(wp));
            uint32_t wt = 0;
            static_for([&](auto cc_tag) {
                constexpr int cc = decltype(cc_tag)::value;
                const double2 x = xn;
                if (cc + 1 < TT && rd)                     // ahead of this column's ring store
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(xn.x), "=d"(xn.y) : "r"(rp), "n"(32 * (cc + 1)));
                double c1 = __shfl_up_sync(0xFFFFFFFFu, c, 1);
                if (lane == 0) c1 = x.x;
                const double oc = acc[cc];
                uint32_t cd = (uint32_t)BS << (BITS * cc);
                double minc = __dadd_rn(__dadd_rn(c, oc), 1.0);
                if (BS == 2) {
                    double c2 = __shfl_up_sync(0xFFFFFFFFu, c, 2);
                    if (lane < 2) c2 = x.y;
                    const double cand = __dadd_rn(__dadd_rn(c2, oc), 2.0);
                    if (cand < minc) { minc = cand; cd = 0; }
                }
                {
                    const double cand = __dadd_rn(c1, oc);         // transition 0: adding +0.0 is the identity
                    if (cand < minc) { minc = cand; cd = (uint32_t)(BS - 1) << (BITS * cc); }
                }
                if (FULL || cc < ncols) {
                    c = minc;
                    wt |= cd;
                    if (pub) asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(wp), "n"(32 * cc), "d"(c) : "memory");
                    if (pub31) asm volatile("st.shared.v2.f64 [%0+%1], {%2, %2};" ::"r"(wp), "n"(32 * cc + 16), "d"(c) : "memory");
                }
            }, std::make_integer_sequence<int, TT>{});
            // the tile's exit costs are published
            if (above) {
                __syncwarp();
                if (lane == 0) dtw_mbar_arrive(full0 + 8u * (w * RT + sout));
            }
            g_in += below ? 1u : 0u;
            g_out += above ? 1u : 0u;
            word |= wt << (BITS * TT * (tile % WPT));
            if ((tile % WPT) == WPT - 1 || t0 + ncols == T) {
                bpp[(int64_t)(tile / WPT) * Spad] = word;
                word = 0;
            }
        };
        fetch_tile(0, 0);
        int t0 = 0, tile = 0;
        for (; t0 + TT <= T; t0 += TT, ++tile) tile_body(std::true_type{}, t0, tile);
        if (t0 < T) tile_body(std::false_type{}, t0, tile);
        asm volatile("cp.async.wait_all;" ::: "memory");
        fin[(size_t)(q & 1) * nst + i] = (i < S) ? c : kInf;
        __syncwarp();
        if (lane == 0) dtw_mbar_arrive(pairdone0 + 8u * (q & 1));         // after the back-pointer stores of the whole warp
    }
}

// pairs -> CTAs: longest processing time first (pairs sorted by decreasing cost, each to the least loaded CTA)
static void dtw_balance(std::vector<int64_t> order, const int64_t* h_toff, const int64_t* h_soff, int nb,
                        std::vector<int32_t>& first, std::vector<int32_t>& list) {
    // cost of a pair in the stream kernel: its T columns move at the pace of the fullest SM sub-partition, which
    // holds ceil(warps / 4) of the pair's warps (measured, S = T = 576 / 608 / 640 / 672: 1 : 1.00 : 1.03 : 1.17)
    auto cost = [&](int64_t p) {
        const int64_t nw = (h_toff[p + 1] - h_toff[p] + 31) / 32;
        return (h_soff[p + 1] - h_soff[p]) * (100 * ((nw + 3) / 4) + 3 * nw);
    };
    std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return cost(x) > cost(y); });
    std::vector<std::vector<int32_t>> bins(nb);
    std::vector<std::pair<int64_t, int>> heap;      // (-load, bin): max-heap on -load = min-heap on load
    for (int b = 0; b < nb; ++b) heap.push_back({0, -b});
    std::make_heap(heap.begin(), heap.end());
    for (int64_t p : order) {
        std::pop_heap(heap.begin(), heap.end());
        auto& top = heap.back();
        bins[-top.second].push_back((int32_t)p);
        top.first -= cost(p);
        std::push_heap(heap.begin(), heap.end());
    }
    first.assign(nb + 1, 0);
    list.clear();
    for (int b = 0; b < nb; ++b) {
        list.insert(list.end(), bins[b].begin(), bins[b].end());
        first[b + 1] = (int32_t)list.size();
    }
}

// shared memory of the stream kernel with nwc compute warps: hand-over ring, private sequence tiles, template
// slices, two final cost columns, mbarriers
static size_t dtw_stream_smem(int nwc, int D, int TT) {
    return ((size_t)nwc * 4 * (4 * TT + 1) + (size_t)nwc * 2 * TT * D + (size_t)nwc * 32 * D + 2 * (size_t)nwc * 32) * sizeof(double) +
           (size_t)(2 * nwc * 4 + 4) * sizeof(uint64_t);
}
static bool dtw_stream_fits(int maxS, int D, int TT) {
    const int nwc = (maxS + 31) / 32;
    return nwc + 1 <= 22 && dtw_stream_smem(nwc, D, TT) <= 227 * 1024;
}

template <int BITS, int TT, int MAXT, int DT, int BS>
static int32_t launch_dtw_stream(const double* tmpl, const int64_t* d_toff, const double* seq, const int64_t* d_soff,
                                 const int64_t* d_bpoff, const int32_t* d_first, const int32_t* d_list, int nb, uint32_t* bp,
                                 int maxS, int64_t* paths, double* final_cost, cudaStream_t st) {
    const int nwc = (maxS + 31) / 32, nt = (nwc + 1) * 32;
    const size_t smem = dtw_stream_smem(nwc, DT, TT);
    auto k = dtw_stream_kernel<BITS, TT, MAXT, DT, BS>;
    VCB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)nb, nt, smem, st>>>(tmpl, d_toff, seq, d_soff, d_bpoff, d_first, d_list, bp, paths, final_cost);
    count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

template <int BITS, int DT, int BS, int FS>
static int32_t launch_dtw(const double* tmplT, const int64_t* d_toff, const double* seq,
                          const int64_t* d_soff, const int64_t* d_bpoff, const int64_t* d_order, uint32_t* bp, int D,
                          int fstep, int bstep, int64_t npairs, int maxS, int64_t* paths,
                          double* final_cost, cudaStream_t st) {
#define VCB_DTW_ARGS tmplT, d_toff, seq, d_soff, d_bpoff, d_order, bp, D, fstep, bstep, npairs, maxS, paths, final_cost, st
    const int nt = round_up(maxS, 32);
    static const int two_ctas = [] { const char* e = getenv("VCB_DTW_2CTA"); return e ? atoi(e) : 1; }();
    // <= 672 states and a compile-time dimension: two CTAs per SM (8-column tiles, 48 registers): one
    // CTA's barrier-paced recurrence overlaps the other's FP64-bound observation costs (the
    // runtime-dimension build would spill at 48 registers)
    if (nt <= 672 && two_ctas && DT > 0) {
        // (two states per thread with 3-4 CTAs/SM measured 3.1-7.6 ms against 2.78: profiles/r02_experiments.txt)
        return launch_dtw_cfg<BITS, 8, 672, 2, DT, BS, FS, 1>(VCB_DTW_ARGS);
    }
    // <= 768 states: 16-column tiles (80-register budget); up to 1024: 8-column tiles
    if (nt <= 768) return launch_dtw_cfg<BITS, 16, 768, 1, DT, BS, FS, 1>(VCB_DTW_ARGS);
    if (nt <= 1024) return launch_dtw_cfg<BITS, 8, 1024, 1, DT, BS, FS, 1>(VCB_DTW_ARGS);
    // longer templates: several consecutive states per thread (the reference has no length limit,
    // src/dtw.jl:93-98); the tile shrinks so that SPT * TT accumulators stay in registers
    if (maxS <= 2048) return launch_dtw_cfg<BITS, 4, 1024, 1, DT, BS, FS, 2>(VCB_DTW_ARGS);
    if (maxS <= 4096) return launch_dtw_cfg<BITS, 2, 1024, 1, DT, BS, FS, 4>(VCB_DTW_ARGS);
    return launch_dtw_cfg<BITS, 2, 1024, 1, DT, BS, FS, 8>(VCB_DTW_ARGS);
#undef VCB_DTW_ARGS
}

// Compile-time specialisations for the windows the reference uses (tests: bstep=1; align: bstep=2,
// both fstep=0) and common mel-cepstrum orders; everything else takes the runtime-parameter build.
template <int BITS, int BS, int FS>
static int32_t launch_dtw_dim(const double* tmplT, const int64_t* d_toff, const double* seq,
                              const int64_t* d_soff, const int64_t* d_bpoff, const int64_t* d_order, uint32_t* bp, int D,
                              int fstep, int bstep, int64_t npairs, int maxS, int64_t* paths,
                              double* final_cost, cudaStream_t st) {
#define VCB_DTW_ARGS tmplT, d_toff, seq, d_soff, d_bpoff, d_order, bp, D, fstep, bstep, npairs, maxS, paths, final_cost, st
    if constexpr (BS >= 0) {
        switch (D) {
            case 24: return launch_dtw<BITS, 24, BS, FS>(VCB_DTW_ARGS);
            case 25: return launch_dtw<BITS, 25, BS, FS>(VCB_DTW_ARGS);
            case 40: return launch_dtw<BITS, 40, BS, FS>(VCB_DTW_ARGS);
            default: break;
        }
    }
    return launch_dtw<BITS, 0, BS, FS>(VCB_DTW_ARGS);
}

int32_t dtw_fit_batch_device(const double* d_tmpl, const int64_t* h_toff, const double* d_seq,
                             const int64_t* h_soff, int64_t npairs, int D, int fstep, int bstep,
                             int64_t* d_paths, double* d_final_cost, cudaStream_t st) {
    if (npairs == 0) return VCB_OK;
    if (D < 1 || fstep < 0 || bstep < 0) return fail(VCB_EARG, "bad DTW arguments (D=%d fstep=%d bstep=%d)", D, fstep, bstep);
    const int window = fstep + bstep + 1;
    if (window > 256) return fail(VCB_EUNSUPPORTED, "DTW window fstep+bstep+1 = %d > 256", window);
    const int bits = window <= 4 ? 2 : (window <= 16 ? 4 : 8);
    const int per = 32 / bits;
    std::vector<int64_t> bpoff(npairs + 1, 0);
    for (int64_t p = 0; p < npairs; ++p) {
        const int64_t S = h_toff[p + 1] - h_toff[p], T = h_soff[p + 1] - h_soff[p];
        if (S < 1 || T < 1) return fail(VCB_EARG, "pair %lld has an empty template or sequence", (long long)p);
        if (S > kDtwMaxStates)
            return fail(VCB_EUNSUPPORTED, "template of %lld frames: one CTA holds the cost column of up to %d states", (long long)S, kDtwMaxStates);
        bpoff[p + 1] = bpoff[p] + ((T + per - 1) / per) * ((S + 31) / 32 * 32);
    }
    // Two kernels share a batch.  The persistent stream kernel takes the pairs it is built for: the reference's own
    // windows, the common dimensions, templates whose warp slices fit in shared memory (D = 24: up to 672 frames,
    // D = 40: 416), 16-byte aligned matrices; the barrier kernel takes the rest (VCB_DTW_STREAM=0 or
    // vcb_set_kernel_variant(1): everything -- the cross-check of the tests).  Inside each group the launch order is
    // decreasing cost S*T, so the short pairs fill the tail.
    static const int stream_on = [] { const char* e = getenv("VCB_DTW_STREAM"); return e ? atoi(e) : 1; }();
    int stream_maxS = 0;
    if (stream_on && g_variant.load() != 1 && fstep == 0 && (bstep == 1 || bstep == 2) && (D == 24 || D == 40) &&
        ((reinterpret_cast<uintptr_t>(d_seq) | reinterpret_cast<uintptr_t>(d_tmpl)) & 15) == 0)
        while (dtw_stream_fits(stream_maxS + 32, D, 8)) stream_maxS += 32;
    auto cost = [&](int64_t p) { return (h_toff[p + 1] - h_toff[p]) * (h_soff[p + 1] - h_soff[p]); };
    std::vector<int64_t> order(npairs);          // [pairs of the barrier kernel | pairs of the stream kernel]
    for (int64_t p = 0; p < npairs; ++p) order[p] = p;
    const int64_t nlong = std::stable_partition(order.begin(), order.end(), [&](int64_t p) { return h_toff[p + 1] - h_toff[p] > stream_maxS; }) - order.begin();
    std::stable_sort(order.begin(), order.begin() + nlong, [&](int64_t x, int64_t y) { return cost(x) > cost(y); });
    std::stable_sort(order.begin() + nlong, order.end(), [&](int64_t x, int64_t y) { return cost(x) > cost(y); });
    const int64_t nshort = npairs - nlong;
    int maxS_long = 0, maxS_short = 0;
    for (int64_t e = 0; e < npairs; ++e) {
        const int S = (int)(h_toff[order[e] + 1] - h_toff[order[e]]);
        if (e < nlong) maxS_long = std::max(maxS_long, S);
        else maxS_short = std::max(maxS_short, S);
    }
    const int64_t totalS = h_toff[npairs];
    // scratch: offsets, transposed templates, packed back-pointers (stream-ordered allocation)
    int64_t* d_off = nullptr;
    double* d_tmplT = nullptr;
    uint32_t* d_bp = nullptr;
    int32_t* d_lists = nullptr;
    const size_t noff = (size_t)(npairs + 1);
    VCB_CUDA(cudaMallocAsync((void**)&d_off, 4 * noff * sizeof(int64_t), st));
    VCB_CUDA(cudaMallocAsync((void**)&d_bp, (size_t)std::max<int64_t>(bpoff[npairs], 1) * sizeof(uint32_t), st));
    // offsets are tiny; pageable async copies complete before return of the call for the host
    // buffers involved (they are staged by the driver), bpoff lives until we synchronise below.
    VCB_CUDA(cudaMemcpyAsync(d_off, h_toff, noff * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    VCB_CUDA(cudaMemcpyAsync(d_off + noff, h_soff, noff * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    VCB_CUDA(cudaMemcpyAsync(d_off + 2 * noff, bpoff.data(), noff * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    VCB_CUDA(cudaMemcpyAsync(d_off + 3 * noff, order.data(), (size_t)npairs * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    int32_t rc = VCB_OK;
    stage_begin(st);
    if (nlong > 0) {
        VCB_CUDA(cudaMallocAsync((void**)&d_tmplT, (size_t)totalS * D * sizeof(double), st));
        dim3 grid((unsigned)npairs, (maxS_long + 31) / 32), block(32, 8);      // (short pairs are transposed along: 3 % of a step)
        dtw_transpose_kernel<<<grid, block, 0, st>>>(d_tmpl, d_off, d_tmplT, D);
        count_launch();
        VCB_CUDA(cudaGetLastError());
    }
    stage_mark(st);      // [0] template transpose (barrier kernel only); [1] the DTW kernels
    if (nshort > 0) {
        int dev = 0, nsm = 0;
        VCB_CUDA(cudaGetDevice(&dev));
        VCB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
        const int nb = (int)std::min<int64_t>(nshort, nsm);
        std::vector<int32_t> first, list;
        dtw_balance(std::vector<int64_t>(order.begin() + nlong, order.end()), h_toff, h_soff, nb, first, list);
        VCB_CUDA(cudaMallocAsync((void**)&d_lists, (size_t)(nb + 1 + nshort) * sizeof(int32_t), st));
        VCB_CUDA(cudaMemcpyAsync(d_lists, first.data(), (size_t)(nb + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        VCB_CUDA(cudaMemcpyAsync(d_lists + nb + 1, list.data(), (size_t)nshort * sizeof(int32_t), cudaMemcpyHostToDevice, st));
#define VCB_DTW_STREAM(MAXT, DT, BS) launch_dtw_stream<2, 8, MAXT, DT, BS>(d_tmpl, d_off, d_seq, d_off + noff, d_off + 2 * noff, d_lists, d_lists + nb + 1, nb, d_bp, maxS_short, d_paths, d_final_cost, st)
        // at most 21 compute warps + the service warp fit (shared memory): 704 threads, 80 registers each
        if (D == 24 && bstep == 2) rc = VCB_DTW_STREAM(704, 24, 2);
        else if (D == 24) rc = VCB_DTW_STREAM(704, 24, 1);
        else if (bstep == 2) rc = VCB_DTW_STREAM(704, 40, 2);
        else rc = VCB_DTW_STREAM(704, 40, 1);
#undef VCB_DTW_STREAM
    }
    if (nlong > 0 && rc == VCB_OK) {
#define VCB_DTW_CALL d_tmplT, d_off, d_seq, d_off + noff, d_off + 2 * noff, d_off + 3 * noff, d_bp, D, fstep, bstep, nlong, maxS_long, d_paths, d_final_cost, st
        if (fstep == 0 && bstep == 1) rc = launch_dtw_dim<2, 1, 0>(VCB_DTW_CALL);
        else if (fstep == 0 && bstep == 2) rc = launch_dtw_dim<2, 2, 0>(VCB_DTW_CALL);
        else if (bits == 2) rc = launch_dtw_dim<2, -1, 0>(VCB_DTW_CALL);
        else if (bits == 4) rc = launch_dtw_dim<4, -1, 0>(VCB_DTW_CALL);
        else rc = launch_dtw_dim<8, -1, 0>(VCB_DTW_CALL);
#undef VCB_DTW_CALL
    }
    stage_mark(st);
    if (d_lists) cudaFreeAsync(d_lists, st);
    cudaFreeAsync(d_off, st);
    if (d_tmplT) cudaFreeAsync(d_tmplT, st);
    cudaFreeAsync(d_bp, st);
    return rc;
}

// One column of the recurrence: update!(d, v)  (src/dtw.jl:61-90)
__global__ void dtw_update_kernel(const double* __restrict__ tmpl, int D, int S,
                                  const double* __restrict__ last, const double* __restrict__ v,
                                  int fstep, int bstep, double* __restrict__ newcost,
                                  int64_t* __restrict__ newbp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    double oc = 0.0;
    for (int k = 0; k < D; ++k) {
        const double d = __dsub_rn(v[k], tmpl[(int64_t)i * D + k]);
        oc = __dadd_rn(oc, __dmul_rn(d, d));
    }
    int minidx = i;
    double minc = __dadd_rn(__dadd_rn(last[i], oc), 1.0);
    const int jlo = max(i - bstep, 0), jhi = min(i + fstep, S - 1);
    for (int j = jlo; j <= jhi; ++j) {
        double cand = __dadd_rn(last[j], oc);
        if (i != j + 1) cand = __dadd_rn(cand, (i == j) ? 1.0 : 2.0);
        if (cand < minc) { minc = cand; minidx = j; }
    }
    newcost[i] = minc;
    newbp[i] = minidx + 1;
}

int32_t dtw_update_device(const double* d_tmpl, int D, int S, const double* d_last,
                          const double* d_v, int fstep, int bstep, double* d_newcost,
                          int64_t* d_newbp, cudaStream_t st) {
    dtw_update_kernel<<<(S + 127) / 128, 128, 0, st>>>(d_tmpl, D, S, d_last, d_v, fstep, bstep, d_newcost, d_newbp);
    count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

}  // namespace vcb
