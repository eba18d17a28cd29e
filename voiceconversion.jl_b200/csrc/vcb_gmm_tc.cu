// vcb_gmm_tc.cu -- K1/K2 on the 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32.
//
// Same mathematics as vcb_gmm_simt.cu (reference src/gmmmap.jl:101-118, src/gmm.jl:24-58), laid
// out as one GEMM per 128-frame tile:
//
//     [128 frames x KP]  .  [KP x N]      KP = [xc (D) | 0-pad | 1 1],  N = G mixtures x rows
//        A = [x - xbar | 1 1]   B = per mixture the rows of Linv_m (whitening) and, for conversion,
//                                   of A_m (regression); the offsets o_m / b_m ride on the two
//                                   "ones" columns as tf32 hi and lo parts
//
// so that TMEM receives z_m = Linv_m (x - mux_m) and Ey_m = muy_m + A_m (x - mux_m) for G mixtures
// per MMA chunk.  When D is a multiple of 8 the ones columns form their own k-step, which needs a
// single MMA (hi*hi is exact for them) instead of three: 10 instead of 12 MMAs per chunk at D = 24.
// fp32 accuracy is needed (real models have cond(Sxx) ~ 1e7; plain TF32 misses the
// 1e-4 parity bar by three orders of magnitude), so both operands are split into tf32 hi + lo
// and each k-step issues three MMAs (hi.hi, hi.lo, lo.hi) into the same fp32 accumulator.
//
// Warp roles (one persistent CTA per SM, 384 threads, clusters of two CTAs):
//   warps 0-7  two epilogue groups (0-3 / 4-7), each taking every other mixture of a chunk: tcgen05.ld
//              of their 32 TMEM lanes (one frame per thread) for both of the group's mixtures, then
//              the accumulator stage is handed back (one mbarrier arrival per warp) BEFORE the
//              arithmetic: |z|^2 -> log-lik, soft-max with a lazy running maximum, posterior-weighted
//              accumulation of Ey (conversion) or running top-2 (arg-max).  The two partial states of
//              a tile merge through shared memory; per-mixture means never leave the SM.
//   warp 8     B producer: cp.async.bulk (TMA 1-D) of pre-packed operand images into a ring of
//              stages; each CTA of the cluster fetches half of every chunk from L2 and multicasts
//              it to both (pair mode: each CTA keeps only its own row half, no multicast).
//   warp 9     MMA issuer: one elected lane issues a whole chunk's tcgen05.mma and the commits from
//              one elected region; owns the TMEM allocation.  Pair mode (template PAIR, cta_group::2,
//              M = 256): rank 0 of the cluster issues for both SMs, rank 1's warp 9 relays the
//              arrival of its half of B to rank 0.
//   warps 10-11 loaders / storers: the Float64 frames of a tile arrive by one bulk copy (conversion,
//              contiguous input) or by direct loads, are centred in Float64, split into tf32 hi/lo
//              with a bit mask and written as the UMMA K-major (no-swizzle) image, double-buffered
//              across tiles; the same warps convert the merged fp32 result tile to Float64 and store
//              it with coalesced rows, so the epilogue never waits on global stores.
// Pipelines: smem B ring (full/empty), A double buffer (full/empty), TMEM accumulator double
// buffer (full/empty), merge buffer (part_full / out_full / part_empty), staged input (x_full) --
// all mbarriers; tcgen05.commit signals the "empty"/"full" transitions of the MMA side.
#include <cstdlib>

#include "vcb_kernels.h"

namespace vcb {

namespace {

constexpr int kTileM = 128;
// Two epilogue groups split the mixtures of every chunk between them (with N = 192 only two
// accumulator stages fit in TMEM, so a stage must come back quickly: each group reads its columns
// and releases the stage before reducing them).  The groups' partial soft-max states merge through
// shared memory at the end of a tile.  VCB_EPI_GROUPS / VCB_LOADER_WARPS other than 2 are for
// timing experiments only (the 4-group merge is not implemented).
#ifndef VCB_EPI_GROUPS
#define VCB_EPI_GROUPS 2
#endif
constexpr int kEpiGroups = VCB_EPI_GROUPS;
constexpr bool kEpiPairs = kEpiGroups < 4;   // a group reduces two mixtures at a time
constexpr int kProducerWarp = 4 * kEpiGroups, kMmaWarp = kProducerWarp + 1, kLoaderWarp0 = kProducerWarp + 2;
#ifndef VCB_LOADER_WARPS
#define VCB_LOADER_WARPS 2
#endif
constexpr int kLoaderWarps = VCB_LOADER_WARPS;      // A loaders / result storers: 128 / (32 kLoaderWarps) rows per thread
constexpr int kLoaderThreads = 32 * kLoaderWarps;
constexpr int kThreads = (kLoaderWarp0 + kLoaderWarps) * 32;
constexpr int kMaxStages = 4;
constexpr size_t kBarBytes = 1024;

struct TcParams {
    const double* X; int64_t T; int64_t ldx;
    const double* xbar;
    const float* B;        // [NCH][2][N*KP] operand images (hi, lo)
    const float* Bpair;    // the same, split by row halves for CTA-pair MMAs
    const float* cst;      // [NCH*G]
    int D, KP, G, NCH, N, stages, abufs;
    int c1, koff;          // ones columns (c1, c1+1); koff = 1: they form a last k-step that needs one MMA
    int64_t ntiles;
    double* Y; int64_t ldy; int copy_power;
    int32_t* mhat; int* flag_count; int64_t* flag_list;
    int xtma;     // 1: whole input tiles are staged in shared memory by one bulk copy (xoff = doubles before X it starts at)
    int xoff;
    int xslack;   // rows at the end of the matrix that must not be staged (their padding may lie outside the buffer)
    int pair;     // 1: CTA-pair MMAs (cta_group::2, M = 256): rank 0 issues for both SMs, each CTA holds N/2 rows of B
    int cluster;  // CTAs per cluster sharing the B operand stream by TMA multicast (1 or 2)
    long long* prof;   // VCB_TC_DEBUG=9: per-role wait/work cycle counters of CTA 0
    int debug;   // timing experiments only (VCB_TC_DEBUG): 1 = B loads shrunk to 16 B, 2 = one k-step of MMAs,
                 // 3 = both, 4 = epilogue does not read TMEM, 5 = A loaders skip the global loads
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// Polite wait for roles that are usually ahead of the pipeline (producer, loaders): back off between
// polls so the spinning warp does not take issue slots from the epilogue warps on its scheduler.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(200);
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// The same copy delivered to the same shared-memory offset of every CTA in `mask`; each
// destination CTA's mbarrier (same offset) receives the complete_tx.
__device__ __forceinline__ void bulk_g2s_mcast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- CTA-pair (cta_group::2) variants and cluster-scope barrier operations
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// arrive on a barrier of another CTA of the cluster (address from mapa_rank)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// same without release semantics: for hand-offs that order tensor-memory accesses only (the
// tcgen05 fences do that); a cluster-scope release costs several hundred cycles per arrival
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a local barrier that receives arrivals from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP_C:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_C;\n"
        "bra WAIT_LOOP_C;\n"
        "DONE_C:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA], M = 256
__device__ __forceinline__ void umma_tf32_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// One lane of a fully active warp (the issuing code stays warp-uniform so descriptors and
// addresses live in uniform registers; a lane-0 branch instead makes the compiler wrap every
// tcgen05.mma in an R2UR "waterfall" loop, ~60 extra cycles per instruction).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, single CTA
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// arrive on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// NC consecutive columns (multiple of 8) with the fewest instructions: a tcgen05.ld costs ~45-60
// cycles of issue per warp whatever its width (tools/micro/tmem_ld_bench.cu), so wide loads win.
template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
    static_assert(NC % 8 == 0, "column count must be a multiple of 8");
    if constexpr (NC >= 32) {
        tmem_ld32(taddr, v);
        if constexpr (NC > 32) tmem_ld_cols<NC - 32>(taddr + 32, v + 32);
    } else if constexpr (NC >= 16) {
        tmem_ld16(taddr, v);
        if constexpr (NC > 16) tmem_ld_cols<NC - 16>(taddr + 16, v + 16);
    } else if constexpr (NC == 8) {
        tmem_ld8(taddr, v);
    }
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// packed fp32x2 arithmetic (FFMA2 / FMUL2): one issue slot for two lanes of the epilogue's FMAs
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

[[maybe_unused]] __device__ __forceinline__ float to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// K-major, non-swizzled operand: 8-row x 16-byte core matrices; LBO = byte distance between
// consecutive 16-byte K slices, SBO = byte distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4)                      // D format: F32
           | (2u << 7) | (2u << 10)       // A, B format: TF32
           | (0u << 15) | (0u << 16)      // A, B K-major
           | ((uint32_t)(N >> 3) << 17)   // N / 8
           | ((uint32_t)(M >> 4) << 24);  // M / 16
}

// ---------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------
#define TIMED_WAIT(bar, par, acc_var)                               \
    do {                                                            \
        const long long _t0 = clock64();                           \
        mbar_wait(bar, par);                                        \
        acc_var += clock64() - _t0;                                 \
    } while (0)
#define TIMED_WAIT_SLEEP(bar, par, acc_var)                         \
    do {                                                            \
        const long long _t0 = clock64();                           \
        mbar_wait_sleep(bar, par);                                  \
        acc_var += clock64() - _t0;                                 \
    } while (0)

// PAIR: CTA-pair MMAs (cta_group::2).  A separate instantiation because a kernel that contains
// cta_group::2 instructions can only be launched in clusters of two.
template <int DP, bool CONVERT, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
gmm_tc_kernel(const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // warp index through a shuffle: provably warp-uniform, so the role branches are uniform branches
    // and the issuing warp's descriptors can live in uniform registers (CUTLASS canonical_warp_idx_sync)
    const int warp = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int KP = p.KP, N = p.N, NCH = p.NCH, AB = p.abufs;
    constexpr bool pair = PAIR;
    const uint32_t crank = (p.cluster > 1) ? cluster_ctarank() : 0u;
    const bool leader = crank == 0;
    const int NB = pair ? N / 2 : N;                 // rows of B held by this CTA
    const int S = p.stages;                          // sized for this mode by tc_plan
    constexpr int ROWS = CONVERT ? 2 * DP : DP;  // TMEM columns per mixture

    // ---- shared memory carve-up
    const uint32_t a_half = kTileM * KP * 4;      // one of hi / lo
    const uint32_t a_bytes = 2 * a_half;
    const uint32_t b_half = (uint32_t)NB * KP * 4;
    const uint32_t b_bytes = 2 * b_half;
    uint8_t* a_smem = smem_raw;
    uint8_t* b_smem = a_smem + (size_t)AB * a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + (size_t)S * b_bytes);
    // bars: a_full[2], a_empty[2], acc_full[2], acc_empty[2], b_full[kMaxStages], b_empty[kMaxStages],
    //       part_full, part_empty
    const uint32_t bar0 = smem_u32(bars);
    auto a_full = [&](int i) { return bar0 + 8u * i; };
    auto a_empty = [&](int i) { return bar0 + 8u * (2 + i); };
    auto acc_full = [&](int i) { return bar0 + 8u * (4 + i); };
    auto acc_empty = [&](int i) { return bar0 + 8u * (6 + i); };
    auto b_full = [&](int i) { return bar0 + 8u * (8 + i); };
    auto b_empty = [&](int i) { return bar0 + 8u * (8 + kMaxStages + i); };
    const uint32_t part_full = bar0 + 8u * (8 + 2 * kMaxStages);
    const uint32_t part_empty = bar0 + 8u * (9 + 2 * kMaxStages);
    const uint32_t out_full = bar0 + 8u * (10 + 2 * kMaxStages);
    const uint32_t x_full = bar0 + 8u * (11 + 2 * kMaxStages);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12 + 2 * kMaxStages);
    float* part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes);  // [PART_ROWS][128]
    // raw Float64 input tile (conversion only, p.xtma): kTileM * ldx doubles behind the merge buffer
    double* xraw = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(part) + (size_t)(CONVERT ? DP + 2 : 4) * 128 * sizeof(float));
    // column constants of the A operand: xbar[k] for data columns, -1 for the two ones columns, 0 for
    // padding, so that every entry is x[k] - colc[k] without branches (x = 0 outside the data)
    double* colc = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(bars) + 192);   // [KP] <= 104 doubles
    for (int k = threadIdx.x; k < KP; k += blockDim.x)
        colc[k] = (k < p.D) ? p.xbar[k] : ((k == p.c1 || k == p.c1 + 1) ? -1.0 : 0.0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(a_full(i), pair ? 2 * kLoaderThreads : kLoaderThreads);    // pair: both CTAs' loaders arrive at rank 0
            mbar_init(a_empty(i), 1);
            mbar_init(acc_full(i), 1);
            mbar_init(acc_empty(i), (pair ? 8 : 4) * kEpiGroups);
        }
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(b_full(i), (pair && leader) ? 2 : 1);      // pair: own expect_tx + the peer's relay
            mbar_init(b_empty(i), pair ? 1u : (uint32_t)p.cluster);
        }
        mbar_init(part_full, 128);
        mbar_init(part_empty, CONVERT ? kLoaderWarps : 128);   // conversion: released by the two storing warps
        mbar_init(out_full, 4);
        mbar_init(x_full, 1);
        fence_barrier_init();
    }
    if (warp == kMmaWarp) {
        if constexpr (PAIR) tmem_alloc2(smem_u32(tmem_slot), 512);
        else tmem_alloc(smem_u32(tmem_slot), 512);
    }
    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();   // peers' barriers are initialised before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // every CTA runs the same number of tile iterations (tiles past the end are computed on zeros
    // and not stored), so the CTAs of a cluster walk the shared B stream in lock step
    const int64_t my_tiles = (p.ntiles + gridDim.x - 1) / gridDim.x;
    const uint16_t cmask = (uint16_t)((1u << p.cluster) - 1u);

    if (warp == kProducerWarp) {
        // ======================= B producer =======================
        if (lane == 0) {
            const int64_t total = my_tiles * NCH;
            long long w_prod = 0;
            const long long t_begin = clock64();
            for (int64_t it = 0; it < total; ++it) {
                const int s = (int)(it % S);
                const uint32_t ph = (uint32_t)((it / S) & 1);
                const int c = (int)(it % NCH);
                TIMED_WAIT_SLEEP(b_empty(s), ph ^ 1, w_prod);
                const uint32_t nbytes = (p.debug == 1 || p.debug == 3) ? 32u : b_bytes;
                mbar_expect_tx(b_full(s), nbytes);
                const uint32_t dst = smem_u32(b_smem + (size_t)s * b_bytes);
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.B + (size_t)c * 2 * N * KP);
                if (pair) {
                    // this CTA's row half of the chunk (hi and lo images are adjacent): one copy, no multicast
                    bulk_g2s(dst, src + (size_t)crank * b_bytes, nbytes, b_full(s));
                } else if (p.cluster > 1) {
                    // each CTA fetches 1/cluster of the chunk from L2 and multicasts it to all
                    const uint32_t slice = nbytes / (uint32_t)p.cluster;
                    const uint32_t off = cluster_ctarank() * slice;
                    bulk_g2s_mcast(dst + off, src + off, slice, b_full(s), cmask);
                } else {
                    bulk_g2s(dst, src, nbytes, b_full(s));
                }
            }
            if (p.prof && blockIdx.x == 0) { p.prof[0] = w_prod; p.prof[1] = clock64() - t_begin; }
        }
    } else if (warp == kMmaWarp && pair && !leader) {
        // ======================= peer CTA of a pair: relays "my half of B has landed" to rank 0 ====
        const int64_t total = my_tiles * NCH;
        for (int64_t it = 0; it < total; ++it) {
            const int s = (int)(it % S);
            mbar_wait(b_full(s), (uint32_t)((it / S) & 1));
            if (lane == 0) mbar_arrive_remote(mapa_rank(b_full(s), 0));
            __syncwarp();
        }
    } else if (warp == kMmaWarp) {
        // ======================= MMA issuer (whole warp runs the loop; one elected lane issues) ===
        {
            const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
            const uint32_t idesc = make_idesc_tf32(pair ? 2 * kTileM : kTileM, N);
            const int ksteps = (p.debug == 2 || p.debug == 3) ? 1 : KP / 8 - p.koff;   // three-pass k-steps
            const uint32_t a_base = smem_u32(a_smem), b_base = smem_u32(b_smem);
            const uint64_t astep = (2u * (kTileM * 16u)) >> 4, bstep = (2u * ((uint32_t)NB * 16u)) >> 4;
            int64_t it = 0;
            long long w_a = 0, w_b = 0, w_acc = 0;
            const long long t_begin = clock64();
            for (int64_t tl = 0; tl < my_tiles; ++tl) {
                const int ab = (int)(tl % AB);
                const uint32_t aph = (uint32_t)((tl / AB) & 1);
                if (pair) { const long long _t0 = clock64(); mbar_wait_cluster(a_full(ab), aph); w_a += clock64() - _t0; }
                else TIMED_WAIT(a_full(ab), aph, w_a);
                const uint32_t a_hi = a_base + (uint32_t)ab * a_bytes;
                for (int c = 0; c < NCH; ++c, ++it) {
                    const int s = (int)(it % S);
                    const uint32_t ph = (uint32_t)((it / S) & 1);
                    const int acc = (int)(it & 1);
                    const uint32_t accph = (uint32_t)((it >> 1) & 1);
                    if (pair) {
                        long long _t0 = clock64();
                        mbar_wait_cluster(b_full(s), ph);
                        long long _t1 = clock64();
                        mbar_wait_cluster(acc_empty(acc), accph ^ 1);
                        w_b += _t1 - _t0; w_acc += clock64() - _t1;
                    } else {
                        TIMED_WAIT(b_full(s), ph, w_b);
                        TIMED_WAIT(acc_empty(acc), accph ^ 1, w_acc);
                    }
                    tc_fence_after();
                    // Descriptors differ from their k-step-0 value only in the start-address field
                    // (low word), which advances by two 16-byte K slices per k-step.
                    const uint32_t b_hi = b_base + (uint32_t)s * b_bytes;
                    const uint32_t d_tmem = tmem_u + (uint32_t)(acc * N);
                    uint64_t dah = make_desc(a_hi, kTileM * 16u, 128u);
                    uint64_t dal = make_desc(a_hi + a_half, kTileM * 16u, 128u);
                    uint64_t dbh = make_desc(b_hi, (uint32_t)NB * 16u, 128u);
                    uint64_t dbl = make_desc(b_hi + b_half, (uint32_t)NB * 16u, 128u);
                    // one elected region per chunk: the descriptor bases move to uniform registers
                    // once and the k-step offsets are uniform adds (an elected region per k-step
                    // re-moved all twelve operands each time, ~40 issue cycles per MMA)
                    if (!pair && elect_one()) {
                        umma_tf32(d_tmem, dal, dbh, idesc, 0u);  // small terms first; first MMA overwrites
                        umma_tf32(d_tmem, dah, dbl, idesc, 1u);
                        umma_tf32(d_tmem, dah, dbh, idesc, 1u);
                        for (int kk = 1; kk < ksteps; ++kk) {
                            dah += astep; dal += astep; dbh += bstep; dbl += bstep;
                            umma_tf32(d_tmem, dal, dbh, idesc, 1u);
                            umma_tf32(d_tmem, dah, dbl, idesc, 1u);
                            umma_tf32(d_tmem, dah, dbh, idesc, 1u);
                        }
                        if (p.koff) {   // offset k-step: ones (exact in tf32) x [o_hi, o_lo]: one pass is exact
                            dah += astep; dbh += bstep;
                            umma_tf32(d_tmem, dah, dbh, idesc, 1u);
                        }
                        // B stage may be refilled once these MMAs retire -- in every CTA of the cluster
                        if (p.cluster > 1) umma_commit_mcast(b_empty(s), cmask);
                        else umma_commit(b_empty(s));
                        umma_commit(acc_full(acc));  // accumulator ready for the epilogue
                        if (c == NCH - 1) umma_commit(a_empty(ab));
                    }
                    if constexpr (PAIR) if (elect_one()) {
                        umma_tf32_2sm(d_tmem, dal, dbh, idesc, 0u);
                        umma_tf32_2sm(d_tmem, dah, dbl, idesc, 1u);
                        umma_tf32_2sm(d_tmem, dah, dbh, idesc, 1u);
                        for (int kk = 1; kk < ksteps; ++kk) {
                            dah += astep; dal += astep; dbh += bstep; dbl += bstep;
                            umma_tf32_2sm(d_tmem, dal, dbh, idesc, 1u);
                            umma_tf32_2sm(d_tmem, dah, dbl, idesc, 1u);
                            umma_tf32_2sm(d_tmem, dah, dbh, idesc, 1u);
                        }
                        if (p.koff) {
                            dah += astep; dbh += bstep;
                            umma_tf32_2sm(d_tmem, dah, dbh, idesc, 1u);
                        }
                        // every release / ready signal goes to the same barrier of both CTAs of the pair
                        umma_commit_2sm(b_empty(s), cmask);
                        umma_commit_2sm(acc_full(acc), cmask);
                        if (c == NCH - 1) umma_commit_2sm(a_empty(ab), cmask);
                    }
                    __syncwarp();
                }
            }
            if (p.prof && blockIdx.x == 0 && lane == 0) { p.prof[2] = w_a; p.prof[3] = w_b; p.prof[4] = w_acc; p.prof[5] = clock64() - t_begin; }
        }
    } else if (warp >= kLoaderWarp0) {
        // ======================= A loaders (64 threads, two frame rows each) =======================
        const int row0 = threadIdx.x - kLoaderWarp0 * 32;
        long long w_load = 0;
        const long long t_begin = clock64();
        // These warps also write the converted tile to global memory: the merging epilogue group
        // leaves the normalised fp32 result in shared memory and goes straight back to draining
        // accumulators (a row-per-thread store from the epilogue cost ~5 k cycles per tile during
        // which the MMA pipe ran dry); from here the store is coalesced and off the critical path.
        long long w_of = 0, w_st = 0;
        auto store_tile = [&](int64_t tls) {
            TIMED_WAIT_SLEEP(out_full, (uint32_t)(tls & 1), w_of);
            const long long ts0 = p.prof ? clock64() : 0;
            const int64_t t0 = (blockIdx.x + tls * gridDim.x) * kTileM;
            {
                constexpr int dR = kLoaderThreads / DP, dC = kLoaderThreads % DP;
                int rr = row0 / DP, cc = row0 % DP;
                constexpr int NE = kTileM * DP / kLoaderThreads;     // exact: kTileM is a multiple of kLoaderThreads
                static_assert(kTileM % kLoaderThreads == 0, "store loop assumes an exact split");
#pragma unroll 8
                for (int i = 0; i < NE; ++i) {
                    const int64_t t = t0 + rr;
                    const float f = part[rr * (DP + 1) + cc];
                    if (t < p.T && cc < p.D) p.Y[t * p.ldy + cc] = (double)f;
                    rr += dR; cc += dC;
                    if (cc >= DP) { cc -= DP; ++rr; }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(part_empty);
            if (p.prof) w_st += clock64() - ts0;
        };
        // Input staging (p.xtma): a full tile of frames is one contiguous block of the caller's
        // matrix, fetched by a single bulk copy one tile ahead; the conversion below then reads it
        // from shared memory instead of issuing 25 strided global loads per row.
        const uint32_t x_bytes = (uint32_t)(kTileM * p.ldx * sizeof(double));
        auto tile_is_staged = [&](int64_t tli) {
            const int64_t tile_i = blockIdx.x + tli * gridDim.x;
            return p.xtma && tli < my_tiles && (tile_i + 1) * kTileM <= p.T - p.xslack;
        };
        auto stage_tile = [&](int64_t tli) {      // one thread
            const int64_t tile_i = blockIdx.x + tli * gridDim.x;
            mbar_expect_tx(x_full, x_bytes);
            bulk_g2s(smem_u32(xraw), p.X - p.xoff + tile_i * kTileM * p.ldx, x_bytes, x_full);
        };
        if (CONVERT && row0 == 0 && tile_is_staged(0)) stage_tile(0);
        uint32_t xph = 0;
        for (int64_t tl = 0; tl < my_tiles; ++tl) {
            const int ab = (int)(tl % AB);
            const uint32_t aph = (uint32_t)((tl / AB) & 1);
            const int64_t tile = blockIdx.x + tl * gridDim.x;
            TIMED_WAIT_SLEEP(a_empty(ab), aph ^ 1, w_load);
            const bool staged = CONVERT && tile_is_staged(tl);
            if (staged) { mbar_wait_sleep(x_full, xph); xph ^= 1; }
            uint8_t* hi = a_smem + (size_t)ab * a_bytes;
            uint8_t* lo = hi + a_half;
            constexpr int RPT = kTileM / kLoaderThreads;      // rows per thread
            // all loads of (up to 32 columns of) every row of this thread are issued before any is
            // used: the frames stream from HBM, so memory-level parallelism is what keeps this role
            // off the critical path
#pragma unroll
            for (int k0 = 0; k0 < DP + 8; k0 += 32) {
                constexpr int CW = (DP + 8 < 32) ? DP + 8 : 32;
                double xv[RPT][CW];
#pragma unroll
                for (int h = 0; h < RPT; ++h) {
                    const int64_t t = tile * kTileM + row0 + kLoaderThreads * h;
                    const bool live = t < p.T && p.debug != 5;
                    const double* x = staged ? xraw + (size_t)(row0 + kLoaderThreads * h) * p.ldx + p.xoff
                                             : p.X + (live ? t : 0) * p.ldx;
#pragma unroll
                    for (int j = 0; j < CW; ++j) xv[h][j] = (live && k0 + j < p.D) ? x[k0 + j] : 0.0;
                    // the power / c0 column in front of the frame passes through (src/common.jl:23);
                    // copied here because a staged tile already holds it
                    if (CONVERT && k0 == 0 && p.copy_power && live) p.Y[t * p.ldy - 1] = (staged && p.xoff) ? x[-1] : p.X[t * p.ldx - 1];
                }
#pragma unroll
                for (int h = 0; h < RPT; ++h) {
                    const int row = row0 + kLoaderThreads * h;
                    const uint32_t row_off = (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
#pragma unroll
                    for (int j4 = 0; j4 < CW; j4 += 4) {
                        if (k0 + j4 < KP) {
                            float4 hv, lv;
                            float* hp = &hv.x;
                            float* lp = &lv.x;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                // hi = v truncated to tf32's 10 mantissa bits (a mask on the Float64
                                // bits, exact in fp32), lo = the remainder; the tensor core reads the
                                // top 10 mantissa bits of lo, so hi + lo carries >= 21 bits of v
                                const double v = xv[h][j4 + j] - colc[k0 + j4 + j];
                                const double vh = __longlong_as_double(__double_as_longlong(v) & 0xFFFFFC0000000000ll);
                                hp[j] = (float)vh;
                                lp[j] = (float)(v - vh);
                            }
                            const uint32_t off = (uint32_t)((k0 + j4) >> 2) * (kTileM * 16u) + row_off;
                            *reinterpret_cast<float4*>(hi + off) = hv;
                            *reinterpret_cast<float4*>(lo + off) = lv;
                        }
                    }
                }
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
            if (pair && !leader) mbar_arrive_remote(mapa_rank(a_full(ab), 0));   // rank 0 issues for both tiles
            else mbar_arrive(a_full(ab));
            if (CONVERT && p.xtma) {
                // every loader thread has consumed the staged tile: refill the buffer for the next one
                asm volatile("bar.sync 2, %0;" ::"n"(kLoaderThreads) : "memory");
                if (row0 == 0 && tile_is_staged(tl + 1)) stage_tile(tl + 1);
            }
            if (CONVERT && tl >= 1) store_tile(tl - 1);
        }
        if (CONVERT) store_tile(my_tiles - 1);
        if (p.prof && blockIdx.x == 0 && row0 == 0) { p.prof[12] = w_load; p.prof[13] = clock64() - t_begin; p.prof[18] = w_of; p.prof[19] = w_st; }
    } else {
        // ======================= epilogue (warps 0-7; thread = frame = TMEM lane) =======================
        const int group = warp >> 2;          // with two groups: the accumulator stage this group drains
        const int row = threadIdx.x & 127;    // 0..127
        const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
        // Fast path (all of a mixture's columns in registers, stage released before the arithmetic):
        // up to 80 columns per mixture; two mixtures at a time while both fit in the register budget.
        constexpr bool kFast = ROWS <= 80;
        constexpr int LOADW = kFast ? ROWS : 32;          // TMEM columns fetched per wait
        constexpr bool kPairsHere = kEpiPairs && LOADW <= 48;
        int64_t it = 0;
        long long w_full = 0, w_part = 0, w_ld = 0, w_rel = 0, w_cmp = 0;
        const long long t_begin = clock64();
        for (int64_t tl = 0; tl < my_tiles; ++tl) {
            const int64_t tile = blockIdx.x + tl * gridDim.x;
            const int64_t t = tile * kTileM + row;
            float mx = -INFINITY, sum = 0.f, second = -INFINITY, qbest = 0.f;
            int best = 0x7FFFFFFF;
            float y[CONVERT ? DP : 1];
            if (CONVERT) {
#pragma unroll
                for (int r = 0; r < DP; ++r) y[r] = 0.f;
            }
            // one arrival per warp: 256 per-thread arrivals on one mbarrier serialise in the shared-
            // memory atomic unit and sat on the hand-off path of every chunk
            auto release_stage = [&](int acc) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (pair && !leader) mbar_arrive_remote_relaxed(mapa_rank(acc_empty(acc), 0));
                    else mbar_arrive(acc_empty(acc));
                }
            };
            for (int c = 0; c < NCH; ++c, ++it) {
                const int acc = (int)(it & 1);
                const uint32_t accph = (uint32_t)((it >> 1) & 1);
                const float* const cst_c = p.cst + c * p.G;
                TIMED_WAIT(acc_full(acc), accph, w_full);
                tc_fence_after();
                const uint32_t tcol = tmem_base + lane_base + (uint32_t)(acc * N);
                if (p.debug == 4) {
                    // timing experiment: accumulators are not read
                } else if (kFast) {
                    // Software pipeline over the mixtures of this chunk: the TMEM loads of mixture
                    // g+1 are in flight while mixture g is reduced (tcgen05.wait::ld waits for all
                    // outstanding loads, so the wait sits after the compute).
                    auto fetch = [&](float* v, int g) {
                        const uint32_t mcol = tcol + (uint32_t)(g * ROWS);
                        tmem_ld_cols<LOADW>(mcol, v);
                    };
                    // Two mixtures are reduced together so their dependent chains (|z|^2, the
                    // exponentials) overlap.  The running maximum is lazy: it moves only when a
                    // log-likelihood exceeds it by more than kLazy, so the 24-wide rescale of the
                    // partial sums is rare instead of once per mixture (weights up to e^kLazy are
                    // far inside fp32 range, and the relative accuracy of a weight does not depend
                    // on the reference point).
                    constexpr float kLazy = 40.0f;
                    auto reduce2 = [&](const float* va, const float* vb, int g, bool two, float cma, float cmb) {
                        // |z|^2 with packed FMAs: two 2-wide partial sums per mixture
                        uint64_t qa01 = 0ull, qa23 = 0ull, qb01 = 0ull, qb23 = 0ull;
#pragma unroll
                        for (int r = 0; r < DP; r += 4) {
                            const uint64_t a0 = pack2(va[r], va[r + 1]), a1 = pack2(va[r + 2], va[r + 3]);
                            const uint64_t b0 = pack2(vb[r], vb[r + 1]), b1 = pack2(vb[r + 2], vb[r + 3]);
                            qa01 = ffma2(a0, a0, qa01);
                            qa23 = ffma2(a1, a1, qa23);
                            qb01 = ffma2(b0, b0, qb01);
                            qb23 = ffma2(b1, b1, qb23);
                        }
                        float qa0, qa1, qa2, qa3, qb0, qb1, qb2, qb3;
                        unpack2(qa01, qa0, qa1); unpack2(qa23, qa2, qa3);
                        unpack2(qb01, qb0, qb1); unpack2(qb23, qb2, qb3);
                        const float qa = (qa0 + qa1) + (qa2 + qa3), qb = (qb0 + qb1) + (qb2 + qb3);
                        const float la = fmaf(-0.5f, qa, cma);
                        const float lb = two ? fmaf(-0.5f, qb, cmb) : -INFINITY;
                        if (CONVERT) {
                            const float lm = fmaxf(la, lb);
                            if (lm > mx + kLazy) {
                                const float a = __expf(mx - lm);   // 0 on the first visit (mx = -inf)
                                sum *= a;
                                const uint64_t a2 = pack2(a, a);
#pragma unroll
                                for (int r = 0; r < DP; r += 2) {
                                    const uint64_t yy = fmul2(pack2(y[r], y[r + 1]), a2);
                                    unpack2(yy, y[r], y[r + 1]);
                                }
                                mx = lm;
                            }
                            const float wa = (la == -INFINITY) ? 0.f : __expf(la - mx);
                            const float wb = (lb == -INFINITY) ? 0.f : __expf(lb - mx);
                            sum += wa + wb;
                            const uint64_t wa2 = pack2(wa, wa), wb2 = pack2(wb, wb);
#pragma unroll
                            for (int r = 0; r < DP; r += 2) {
                                constexpr int EO = CONVERT ? DP : 0;
                                uint64_t yy = pack2(y[r], y[r + 1]);
                                // (without a second mixture vb is not loaded: 0 x garbage could be NaN)
                                if (two) yy = ffma2(wb2, pack2(vb[EO + r], vb[EO + r + 1]), yy);
                                yy = ffma2(wa2, pack2(va[EO + r], va[EO + r + 1]), yy);
                                unpack2(yy, y[r], y[r + 1]);
                            }
                        } else {
                            const int m = c * p.G + g;
                            if (la > mx) { second = mx; mx = la; best = m; qbest = qa; }
                            else if (la > second) second = la;
                            if (lb > mx) { second = mx; mx = lb; best = m + kEpiGroups; qbest = qb; }
                            else if (lb > second) second = lb;
                        }
                    };
                    // both groups drain the same chunk: group e takes mixtures e, e + kEpiGroups, ...
                    constexpr int GS = kEpiGroups;
                    float v0[LOADW], v1[kPairsHere ? LOADW : 1];
                    bool released = false;
                    for (int g = group; g < p.G; g += (kPairsHere ? 2 : 1) * GS) {
                        const bool two = kPairsHere && g + GS < p.G;
                        const float cma = cst_c[g], cmb = two ? cst_c[g + GS] : -INFINITY;   // in flight under the TMEM loads
                        const long long tl0 = p.prof ? clock64() : 0;
                        fetch(v0, g);
                        if (two) fetch(v1, g + GS);
                        tmem_ld_wait();
                        if (p.prof) w_ld += clock64() - tl0;
                        if (g + (kPairsHere ? 2 : 1) * GS >= p.G) {
                            // this group's last columns of the stage are in registers: hand the
                            // accumulator back before reducing them, so the stage is held for the
                            // TMEM read only and the MMA of chunk c+2 overlaps this arithmetic
                            release_stage(acc);
                            released = true;
                        }
                        const long long tc0 = p.prof ? clock64() : 0;
                        reduce2(v0, v1, g, two, cma, cmb);
                        if (p.prof) w_cmp += clock64() - tc0;
                    }
                    if (!released) release_stage(acc);
                    continue;
                } else {
                for (int g = group; g < p.G; g += kEpiGroups) {
                    const int m = c * p.G + g;
                    const uint32_t mcol = tcol + (uint32_t)(g * ROWS);
                    const float cm = p.cst[m];  // -inf for padding mixtures
                    {
                        float q = 0.f;
#pragma unroll
                        for (int r0 = 0; r0 < DP; r0 += 32) {
                            float v[32];
                            constexpr int W0 = (DP >= 32) ? 32 : DP;
                            if (r0 + 32 <= DP) tmem_ld_cols<W0>(mcol + r0, v);
                            else tmem_ld_cols<(DP % 32) ? (DP % 32) : 32>(mcol + r0, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int r = 0; r < 32; ++r)
                                if (r0 + r < DP) q = fmaf(v[r], v[r], q);
                        }
                        const float l = fmaf(-0.5f, q, cm);
                        if (CONVERT) {
                            if (l > mx) {
                                const float a = __expf(mx - l);
                                sum *= a;
#pragma unroll
                                for (int r = 0; r < DP; ++r) y[r] *= a;
                                mx = l;
                            }
                            const float w = (l == -INFINITY) ? 0.f : __expf(l - mx);
                            sum += w;
#pragma unroll
                            for (int r0 = 0; r0 < DP; r0 += 32) {
                                float v[32];
                                constexpr int W0 = (DP >= 32) ? 32 : DP;
                                if (r0 + 32 <= DP) tmem_ld_cols<W0>(mcol + DP + r0, v);
                                else tmem_ld_cols<(DP % 32) ? (DP % 32) : 32>(mcol + DP + r0, v);
                                tmem_ld_wait();
#pragma unroll
                                for (int r = 0; r < 32; ++r)
                                    if (r0 + r < DP) y[r0 + r] = fmaf(w, v[r], y[r0 + r]);
                            }
                        } else {
                            if (l > mx) { second = mx; mx = l; best = m; qbest = q; }
                            else if (l > second) second = l;
                        }
                    }
                }
                }
                release_stage(acc);
            }
            // ---- merge the two groups' partial states of this tile (group 1 -> smem -> group 0)
            const uint32_t pph = (uint32_t)(tl & 1);
            // the merging/storing group alternates per tile so both groups carry the same load
            const int merger = (kEpiGroups == 2) ? (int)(tl & 1) : 0;
            if (kEpiGroups > 2 && group != 0) {
                // (timing experiments with more epilogue groups only: their partial states are dropped)
            } else if (kEpiGroups == 2 && group != merger) {
                mbar_wait(part_empty, pph ^ 1);
                part[0 * 128 + row] = mx;
                if (CONVERT) {
                    part[1 * 128 + row] = sum;
#pragma unroll
                    for (int r = 0; r < DP; ++r) part[(2 + r) * 128 + row] = y[r];
                } else {
                    part[1 * 128 + row] = second;
                    part[2 * 128 + row] = qbest;
                    part[3 * 128 + row] = __int_as_float(best);
                }
                mbar_arrive(part_full);
            } else {
                const long long ts0 = p.prof ? clock64() : 0;
                if (kEpiGroups == 2) TIMED_WAIT(part_full, pph, w_part);
                const float mxb = (kEpiGroups == 2) ? part[0 * 128 + row] : -INFINITY;
                if (CONVERT) {
                    const float m2 = fmaxf(mx, mxb);
                    const float fa = (mx == -INFINITY) ? 0.f : __expf(mx - m2);
                    const float fb = (mxb == -INFINITY) ? 0.f : __expf(mxb - m2);
                    const float sumb = (kEpiGroups == 2) ? part[1 * 128 + row] : 0.f;
                    const float inv = 1.0f / (sum * fa + sumb * fb);
                    float fin[DP];
#pragma unroll
                    for (int r = 0; r < DP; ++r)
                        fin[r] = (y[r] * fa + ((kEpiGroups == 2) ? part[(2 + r) * 128 + row] : 0.f) * fb) * inv;
                    // every thread of this group has read its partial: the buffer becomes the
                    // tile's result [row][DP + 1] for the storing warps
                    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                    for (int r = 0; r < DP; ++r) part[row * (DP + 1) + r] = fin[r];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(out_full);
                } else {
                    const float secb = (kEpiGroups == 2) ? part[1 * 128 + row] : -INFINITY, qb = (kEpiGroups == 2) ? part[2 * 128 + row] : 0.f;
                    const int bestb = (kEpiGroups == 2) ? __float_as_int(part[3 * 128 + row]) : 0x7FFFFFFF;
                    // first maximum wins on ties, like indmax (src/gmm.jl:46)
                    const bool a_wins = (mx > mxb) || (mx == mxb && best < bestb);
                    const float top = a_wins ? mx : mxb;
                    const float sec = a_wins ? fmaxf(second, mxb) : fmaxf(secb, mx);
                    const float qtop = a_wins ? qbest : qb;
                    const int btop = a_wins ? best : bestb;
                    if (t < p.T) {
                        p.mhat[t] = btop;
                        if (top - sec < 1e-3f * (1.0f + qtop) || !(top == top)) {
                            const int slot = atomicAdd(p.flag_count, 1);
                            p.flag_list[slot] = t;
                        }
                    }
                }
                if (kEpiGroups == 2 && !CONVERT) mbar_arrive(part_empty);
                if (p.prof) w_rel += clock64() - ts0;     // (profiling: tile-end merge + store)
            }
        }
        if (p.prof && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 128)) {
            const int o = threadIdx.x == 0 ? 6 : 9;
            p.prof[o] = w_full; p.prof[o + 1] = w_part; p.prof[o + 2] = clock64() - t_begin;
            p.prof[threadIdx.x == 0 ? 14 : 15] = w_ld;
            if (threadIdx.x == 0) { p.prof[16] = w_rel; p.prof[17] = w_cmp; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into it
    if (warp == kMmaWarp) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc2(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

template <int DP, bool CONVERT>
int32_t launch_tc(const TcParams& p_in, size_t smem, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    TcParams p = p_in;
    unsigned grid = (unsigned)std::min<int64_t>(p.ntiles, sms);
    if (p.cluster == 2) grid &= ~1u;
    constexpr bool kHasPair = (CONVERT && DP == 24) || (!CONVERT && DP == 48);   // keep in step with fill_common
    auto k = gmm_tc_kernel<DP, CONVERT, false>;
    if constexpr (kHasPair) { if (p.pair) k = gmm_tc_kernel<DP, CONVERT, true>; }
    VCB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    long long* d_prof = nullptr;
    if (p.debug == 9) { cudaMalloc((void**)&d_prof, 24 * sizeof(long long)); cudaMemset(d_prof, 0, 24 * sizeof(long long)); p.prof = d_prof; }
    VCB_CUDA(cudaLaunchKernelEx(&cfg, k, p));
    if (d_prof) {
        long long h[24];
        cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
        cudaFree(d_prof);
        fprintf(stderr, "[tc prof CTA0] producer: wait b_empty %lld of %lld | mma: wait a_full %lld b_full %lld acc_empty %lld of %lld | epi0: wait acc_full %lld part %lld of %lld | epi1: wait acc_full %lld part %lld of %lld | loader: wait a_empty %lld of %lld | tmem ld epi0 %lld epi1 %lld | epi0 tile-end %lld reduce %lld | storer: wait out_full %lld store %lld\n",
                h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10], h[11], h[12], h[13], h[14], h[15], h[16], h[17], h[18], h[19]);
    }
    count_launch();
    VCB_CUDA(cudaGetLastError());
    return VCB_OK;
}

template <bool CONVERT>
int32_t dispatch_tc(int DP, const TcParams& p, size_t smem, cudaStream_t st) {
    switch (DP) {
        case 8: return launch_tc<8, CONVERT>(p, smem, st);
        case 16: return launch_tc<16, CONVERT>(p, smem, st);
        case 24: return launch_tc<24, CONVERT>(p, smem, st);
        case 32: return launch_tc<32, CONVERT>(p, smem, st);
        case 40: return launch_tc<40, CONVERT>(p, smem, st);
        case 48: return launch_tc<48, CONVERT>(p, smem, st);
        case 56: return launch_tc<56, CONVERT>(p, smem, st);
        case 64: return launch_tc<64, CONVERT>(p, smem, st);
        default: break;
    }
    if (!CONVERT) {
        switch (DP) {
            case 72: return launch_tc<72, false>(p, smem, st);
            case 80: return launch_tc<80, false>(p, smem, st);
            case 88: return launch_tc<88, false>(p, smem, st);
            case 96: return launch_tc<96, false>(p, smem, st);
            default: break;
        }
    }
    return fail(VCB_EUNSUPPORTED, "tcgen05 kernel: unsupported padded dimension %d", DP);
}

constexpr size_t kSmemLimit = 227 * 1024;

}  // namespace

// pair: the plan is for CTA-pair MMAs, where a CTA stores only its row half of every B stage.
// force_g: keep the packed number of mixtures per chunk and only size the buffers (launch time).
TcPlan tc_plan(int M, int KP, int rows_per_mixture, int part_rows, bool pair, int force_g) {
    // Pick (mixtures per chunk, A buffers) by estimated tensor-pipe time per tile.  Measured on
    // B200 (tools/micro/umma_bench.cu): one thread issues a tcgen05.mma every ~84 cycles at best,
    // an M=128 x N x K=8 tf32 MMA executes in ~N/2 + 11 cycles.
    TcPlan best;
    // conversion (part_rows > 4): room for one raw Float64 input tile of up to DP + 1 columns
    const size_t xraw_bytes = part_rows > 4 ? (size_t)kTileM * (part_rows - 1) * sizeof(double) : 0;
    const size_t extra = kBarBytes + (size_t)part_rows * 128 * sizeof(float) + xraw_bytes;
    const int gmax = 256 / rows_per_mixture;
    if (gmax < 1) return best;
    const int ksteps = KP / 8;
    double best_cost = 1e30;
    for (int abufs = 2; abufs >= 1; --abufs) {
        for (int g = gmax; g >= 1; --g) {
            if (force_g && g != force_g) continue;
            const int n = g * rows_per_mixture;
            if (n % 16) continue;
            const size_t a = (size_t)abufs * 2 * kTileM * KP * 4;
            const size_t bst = ((size_t)2 * n * KP * 4) / (pair ? 2 : 1);
            // a forced shape (the fallback of a pair-planned model to single-CTA MMAs) may run on one stage
            if (a + (force_g ? 1 : 2) * bst + extra > kSmemLimit) continue;
            const int nch = (M + g - 1) / g;
            const double per_mma = std::max(84.0, n / 2.0 + 11.0);
            const double cost = (double)nch * ksteps * 3 * per_mma + (abufs == 1 ? 4000.0 : 0.0);
            if (cost < best_cost - 1e-9) {
                best_cost = cost;
                int stages = (int)((kSmemLimit - extra - a) / bst);
                if (stages > kMaxStages) stages = kMaxStages;
                best.G = g; best.N = n; best.stages = stages; best.abufs = abufs;
                best.smem = a + (size_t)stages * bst + extra;
            }
        }
    }
    return best;
}

bool tc_supported(const vcb_gmmmap& g, bool convert) {
    if (convert) return g.tc.GC > 0 && g.DP <= 64;
    return g.tc.GW > 0 && g.DP <= 96;
}

static void fill_common(const vcb_gmmmap& g, const double* dX, int64_t T, int64_t ldx, bool convert, TcParams& p,
                        size_t& smem) {
    // cluster / pair decision (needs only the problem size), then the buffer plan for that mode
    const int64_t ntiles = (T + kTileM - 1) / kTileM;
    const int packedG = convert ? g.tc.GC : g.tc.GW, packedN = convert ? g.tc.NC : g.tc.NW;
    static const int want_cluster = [] { const char* e = getenv("VCB_TC_CLUSTER"); return e ? atoi(e) : 2; }();
    // clusters of two share the B stream when there are at least two tiles per cluster to amortise it
    p.cluster = (want_cluster >= 2 && ntiles >= 2 && (packedN * g.tc.KP * 8) % 32 == 0) ? 2 : 1;
    // CTA-pair MMAs (cta_group::2): rank 0 of every cluster issues M = 256 instructions for both SMs.
    // Default: on for the arg-max kernel (light epilogue, MMA-bound), off for the conversion kernel
    // (epilogue-bound once the MMA stream is cheaper: no gain).  VCB_TC_PAIR=0/1 forces both.
    static const int want_pair = [] { const char* e = getenv("VCB_TC_PAIR"); return e ? atoi(e) : -1; }();
    const bool use_pair = want_pair < 0 ? !convert : want_pair != 0;
    const bool has_pair_kernel = (convert && g.DP == 24) || (!convert && g.DP == 48);   // instantiated shapes
    p.pair = (use_pair && has_pair_kernel && p.cluster == 2 && packedN % 16 == 0) ? 1 : 0;
    const TcPlan plan = tc_plan(g.M, g.tc.KP, convert ? 2 * g.DP : g.DP, convert ? g.DP + 2 : 4, p.pair != 0, packedG);
    p.X = dX; p.T = T; p.ldx = ldx; p.xbar = g.d_xbar.p;
    p.B = convert ? g.tc.Bc.p : g.tc.Bw.p;
    p.Bpair = convert ? g.tc.Bc2.p : g.tc.Bw2.p;
    if (p.pair) p.B = p.Bpair;
    p.cst = g.tc.cst.p;
    p.c1 = g.tc.c1; p.koff = g.tc.koff;
    p.D = g.D; p.KP = g.tc.KP; p.G = plan.G; p.N = plan.N;
    p.NCH = convert ? g.tc.NCHC : g.tc.NCHW;
    p.stages = plan.stages; p.abufs = plan.abufs;
    p.xtma = 0; p.xoff = 0; p.xslack = 0;
    if (convert && ldx <= g.DP + 1) {
        // bulk copies need 16-byte aligned global addresses; in vc() layout X = fm + 1 is odd-aligned
        // and the power column in front of it belongs to the same matrix, so start one double early
        const bool odd = (reinterpret_cast<uintptr_t>(dX) >> 3) & 1;
        if ((reinterpret_cast<uintptr_t>(dX) & 7) == 0 && (!odd || ldx == g.D + 1)) {
            static const bool off = [] { const char* e = getenv("VCB_TC_XTMA"); return e && e[0] == '0'; }();
            p.xtma = off ? 0 : 1;
            p.xoff = odd ? 1 : 0;
            p.xslack = (ldx == g.D + p.xoff) ? 0 : 1;
        }
    }
    static const int force_stages = [] { const char* e = getenv("VCB_TC_STAGES"); return e ? atoi(e) : 0; }();
    if (force_stages >= 1 && force_stages < p.stages) p.stages = force_stages;     // experiments only
    p.ntiles = (T + kTileM - 1) / kTileM;
    static const int dbg = [] { const char* e = getenv("VCB_TC_DEBUG"); return e ? atoi(e) : 0; }();
    p.debug = dbg;
    smem = plan.smem;
}

int32_t tc_convert(const vcb_gmmmap& g, const double* dX, int64_t T, int64_t ldx, double* dY, int64_t ldy,
                   bool copy_power, cudaStream_t st) {
    if (T == 0) return VCB_OK;
    TcParams p{};
    size_t smem = 0;
    fill_common(g, dX, T, ldx, true, p, smem);
    p.Y = dY; p.ldy = ldy; p.copy_power = copy_power ? 1 : 0;
    return dispatch_tc<true>(g.DP, p, smem, st);
}

int32_t tc_argmax(const vcb_gmmmap& g, const double* dX, int64_t T, int64_t ldx, int32_t* d_mhat,
                  cudaStream_t st) {
    if (T == 0) return VCB_OK;
    int* d_count = nullptr;
    int64_t* d_list = nullptr;
    VCB_CUDA(cudaMallocAsync((void**)&d_count, sizeof(int), st));
    VCB_CUDA(cudaMallocAsync((void**)&d_list, (size_t)T * sizeof(int64_t), st));
    VCB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), st));
    TcParams p{};
    size_t smem = 0;
    fill_common(g, dX, T, ldx, false, p, smem);
    p.mhat = d_mhat; p.flag_count = d_count; p.flag_list = d_list;
    int32_t rc = dispatch_tc<false>(g.DP, p, smem, st);
    if (rc == VCB_OK) rc = recheck_argmax_fp64(g, dX, ldx, d_count, d_list, d_mhat, st);
    cudaFreeAsync(d_count, st);
    cudaFreeAsync(d_list, st);
    return rc;
}

}  // namespace vcb
