// Microbenchmark: tcgen05.ld (TMEM -> registers) bandwidth per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_ld_bench tools/micro/tmem_ld_bench.cu && /tmp/tmem_ld_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>  // columns per load: 8, 16 or 32
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* v);
template <> __device__ __forceinline__ void ld<8>(uint32_t t, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(t));
}
template <> __device__ __forceinline__ void ld<32>(uint32_t t, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(t));
}

template <int X>
__global__ void bench(int iters, int inflight, long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot + (((uint32_t)(warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t v[4][X];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < inflight) ld<X>(tm + ((i * 4 + j) * X) % (512 - X), v[j]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < inflight) acc ^= v[j][0] ^ v[j][X - 1];
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
    long long* out; uint32_t* sink;
    cudaMallocManaged(&out, 8); cudaMalloc(&sink, 4096);
    const int iters = 4000;
    for (int threads : {32, 128, 256})
        for (int inflight : {1, 2, 4}) {
            *out = 0; bench<8><<<148, threads>>>(iters, inflight, out, sink); cudaDeviceSynchronize();
            double bytes = (double)iters * inflight * 8 * 4 * threads;
            printf("x8  threads=%3d inflight=%d : %7.1f B/clk/SM  (%s)\n", threads, inflight, bytes / *out, cudaGetErrorString(cudaGetLastError()));
            *out = 0; bench<32><<<148, threads>>>(iters, inflight, out, sink); cudaDeviceSynchronize();
            bytes = (double)iters * inflight * 32 * 4 * threads;
            printf("x32 threads=%3d inflight=%d : %7.1f B/clk/SM  (%s)\n", threads, inflight, bytes / *out, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
