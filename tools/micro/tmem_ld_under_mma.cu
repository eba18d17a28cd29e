// Microbenchmark: latency of tcgen05.ld (TMEM -> registers) while another warp keeps the tensor core
// busy with tcgen05.mma into a DIFFERENT accumulator stage.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/t tools/micro/tmem_ld_under_mma.cu && /tmp/t
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void ld32(uint32_t t, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(t));
}

// warps 0-3: loaders of TMEM columns [256, 448); warp 4 lane 0: MMA issuer into columns [0, 192)
__global__ void __launch_bounds__(160, 1) bench(int mma_on, int nld, long long* out, uint32_t* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3F800000u;
    if (threadIdx.x == 0) {
        stop = 0;
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
    if (warp == 4) {
        if (threadIdx.x == 128 && mma_on) {
            const int N = 192;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
            const uint64_t da = make_desc(smem_u32(smem), 128 * 16, 128), db = make_desc(smem_u32(smem) + 65536, N * 16, 128);
            int i = 0;
            while (!stop) {
                for (int k = 0; k < 12; ++k)
                    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(k > 0 ? 1 : 0) : "memory");
                if (mma_on == 2) {   // wait for the batch (queue never deeper than 12 MMAs)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                    asm volatile("{.reg .pred P1; W: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1; @P1 bra D; bra W; D: }" ::"r"(smem_u32(&bar)), "r"(i & 1) : "memory");
                }
                ++i;
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
    } else {
        const uint32_t ta = tm + (((uint32_t)warp * 32) << 16) + 256;
        uint32_t acc = 0;
        long long total = 0, worst = 0;
        for (int it = 0; it < 2000; ++it) {
            uint32_t v[2][32];
            const long long t0 = clock64();
            ld32(ta, v[0]);
            if (nld > 1) ld32(ta + 32, v[1]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const long long dt = clock64() - t0;
            total += dt; if (dt > worst) worst = dt;
            acc ^= v[0][0] ^ v[0][31] ^ (nld > 1 ? v[1][7] : 0);
            for (int spin = 0; spin < 20; ++spin) acc = acc * 1664525u + 1013904223u;   // ~ a little compute
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = total / 2000; out[1] = worst; }
        if (acc == 0x1234567u) sink[threadIdx.x] = acc;
        __syncwarp();
        if (threadIdx.x == 0) stop = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
    long long* out; uint32_t* sink;
    cudaMallocManaged(&out, 16); cudaMalloc(&sink, 4096);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int mma_on = 0; mma_on < 3; ++mma_on)
        for (int nld : {1, 2}) {
            out[0] = out[1] = 0;
            bench<<<148, 160, 160 * 1024>>>(mma_on, nld, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            printf("mma=%s loads_in_flight=%d : mean latency %lld cycles, worst %lld  (%s)\n",
                   mma_on == 0 ? "off" : (mma_on == 1 ? "continuous" : "batches-of-12"), nld, out[0], out[1], cudaGetErrorString(e));
        }
    return 0;
}
