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

// mode: 0 = same A/B every MMA, one accumulator; 1 = alternate two accumulators; 2 = 3 MMAs per k-step
// with hi/lo operand alternation (like the product kernel); 3 = like 0 but 128B-swizzle descriptors
template <int KIND>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int mode, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
            else {
                umma<KIND>(tm, dA1, dB0, idesc, i > 0);
                umma<KIND>(tm, dA0, dB1, idesc, 1);
                umma<KIND>(tm, dA0, dB0, idesc, 1);
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
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    for (int kind = 0; kind < 2; ++kind)
        for (int mode = 0; mode < 4; ++mode)
            for (int N : {64, 96, 128, 192, 256}) {
                if (mode == 1 && N > 256) continue;
                for (int grid : {1, 148}) {
                    *out = 0;
                    if (kind == 0) bench<0><<<grid, 128, smem>>>(N, iters, mode, out);
                    else bench<1><<<grid, 128, smem>>>(N, iters, mode, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    const int nmma = (mode == 2 ? 3 : 1) * iters;
                    printf("kind=%s mode=%d N=%3d grid=%3d : %8.1f cycles/MMA  (%s)\n", kind ? "f16 " : "tf32", mode, N, grid,
                           (double)*out / nmma, cudaGetErrorString(e));
                }
            }
    return 0;
}
