// Microbenchmark: cycles per tcgen05.mma (cta_group::1, M=128) for kind::tf32 / kind::f16,
// several N, K-major non-swizzled operands in shared memory.  Build & run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_bench tools/micro/umma_bench.cu && /tmp/umma_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
template <int KIND>  // 0 tf32, 1 f16
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND>
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 4 = product-kernel pattern from shared memory: 4 k-steps x 3 MMAs per chunk, descriptors
// advance per k-step, B cycles through 3 stage buffers; 5 = same with A read from TMEM
// mode: 0 = same A/B every MMA, one accumulator; 1 = alternate two accumulators; 2 = 3 MMAs per k-step
// with hi/lo operand alternation (like the product kernel); 3 = like 0 but 128B-swizzle descriptors
template <int KIND>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int mode, long long* out, int randomize) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint64_t dummy[4];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 220 * 1024 / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 97u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        // random tf32 values in [1, 2) with random sign when randomize, else zeros
        reinterpret_cast<uint32_t*>(smem)[i] = randomize ? (0x3F800000u | (h & 0x007FE000u) | ((h & 1u) << 31)) : 0u;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        for (int q = 0; q < 4; ++q) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy[q])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t fmt = KIND == 0 ? 2u : 0u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t a0 = smem_u32(smem), a1 = a0 + 16384, b0 = a0 + 65536, b1 = b0 + 49152;
        const uint32_t lboA = 128 * 16, lboB = N * 16;
        uint64_t dA0, dA1, dB0, dB1;
        if (mode == 3) {  // 128B swizzle, K-major: 8 rows x 128 B atoms, SBO = 1024
            dA0 = make_desc(a0, 16, 1024, 2); dA1 = make_desc(a1, 16, 1024, 2);
            dB0 = make_desc(b0, 16, 1024, 2); dB1 = make_desc(b1, 16, 1024, 2);
        } else {
            dA0 = make_desc(a0, lboA, 128, 0); dA1 = make_desc(a1, lboA, 128, 0);
            dB0 = make_desc(b0, lboB, 128, 0); dB1 = make_desc(b1, lboB, 128, 0);
        }
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (mode == 0 || mode == 3) umma<KIND>(tm, dA0, dB0, idesc, i > 0);
            else if (mode == 1) umma<KIND>(tm + (i & 1) * N, dA0, dB0, idesc, i > 1);
            else if (mode == 2) {
                umma<KIND>(tm, dA1, dB0, idesc, i > 0);
                umma<KIND>(tm, dA0, dB1, idesc, 1);
                umma<KIND>(tm, dA0, dB0, idesc, 1);
            } else {
                // chunk i: stage buffer (i % 3), 4 k-steps; A hi at a0, lo at a1 (smem) or TMEM cols 384.. / 416..
                const uint32_t tmd = (mode == 7 || mode == 8) ? tm + (i & 1) * N : tm;
                if (mode >= 8) {   // two waits on barriers whose current phase is already complete (parity trick)
                    asm volatile("{.reg .pred P1; W1: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 1; @P1 bra D1; bra W1; D1: }" ::"r"(smem_u32(&dummy[2])) : "memory");
                    asm volatile("{.reg .pred P1; W2: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 1; @P1 bra D2; bra W2; D2: }" ::"r"(smem_u32(&dummy[3])) : "memory");
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t bst = b0 + (i % 3) * 49152u;   // 3 stages x 48 KB (64 KB A + 144 KB B = 208 KB)
                uint64_t bh = make_desc(bst, lboB, 128, 0), bl = make_desc(bst + N * 32 * 4, lboB, 128, 0);
                uint64_t ah = dA0, al = dA1;
                const uint64_t astep = (2u * 128u * 16u) >> 4, bstep = (2u * (uint32_t)N * 16u) >> 4;
                for (int kk = 0; kk < 4; ++kk) {
                    if (mode == 4 || mode >= 7) {
                        umma<KIND>(tmd, al, bh, idesc, kk > 0);
                        umma<KIND>(tmd, ah, bl, idesc, 1);
                        umma<KIND>(tmd, ah, bh, idesc, 1);
                    } else if (mode == 6) {   // every consecutive pair shares one operand
                        umma<KIND>(tm, al, bh, idesc, kk > 0);
                        umma<KIND>(tm, ah, bh, idesc, 1);
                        umma<KIND>(tm, ah, bl, idesc, 1);
                    } else {
                        umma_ts<KIND>(tm, tm + 416 + kk * 8, bh, idesc, kk > 0);
                        umma_ts<KIND>(tm, tm + 384 + kk * 8, bl, idesc, 1);
                        umma_ts<KIND>(tm, tm + 384 + kk * 8, bh, idesc, 1);
                    }
                    ah += astep; al += astep; bh += bstep; bl += bstep;
                }
                if (mode >= 7) {   // two commits per chunk, as the product kernel does (barriers never waited on)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[0])) : "memory");
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[1])) : "memory");
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{.reg .pred P1; W: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0; @P1 bra D; bra W; D: }" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 8);
    const int smem = 220 * 1024;
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 4000;
    for (int kind = 0; kind < 2; ++kind)
        for (int mode : {4, 8})
            for (int N : {192}) {
                if (kind == 1) continue;
                for (int grid : {148}) {
                    *out = 0;
                    for (int rnd = 0; rnd < 2; ++rnd) {
                    *out = 0;
                    if (kind == 0) bench<0><<<grid, 128, smem>>>(N, iters, mode, out, rnd);
                    else bench<1><<<grid, 128, smem>>>(N, iters, mode, out, rnd);
                    cudaError_t e = cudaDeviceSynchronize();
                    const int nmma = (mode == 2 ? 3 : (mode >= 4 ? 12 : 1)) * iters;
                    printf("kind=%s mode=%d N=%3d grid=%3d data=%s : %8.1f cycles/MMA  (%s)\n", kind ? "f16 " : "tf32", mode, N, grid,
                           rnd ? "random" : "zeros", (double)*out / nmma, cudaGetErrorString(e));
                    }
                }
            }
    return 0;
}
