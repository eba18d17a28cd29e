// dmma_bench.cu -- throughput / latency of mma.sync.m8n8k4.f64 and DFMA on sm_100a, alone and mixed.
// usage: dmma_bench   (prints cycles per instruction per SM sub-partition for several occupancies)
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int MODE>   // MODE 0: DMMA only, 1: DFMA only, 2: 1 DMMA : 4 DFMA mixed
__global__ void k(double* out, int iters, long long* cyc) {
    double c[ILP][2], f[ILP];
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; f[i] = i + threadIdx.x; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE != 1)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            if (MODE == 1) f[i] = fma(f[i], a, b);
            if (MODE == 2) {
#pragma unroll
                for (int r = 0; r < 4; ++r) f[i] = fma(f[i], a, b);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP, int MODE>
void run(int warps_per_sm, const char* name) {
    double* out; long long* cyc; long long h;
    int sms = 148, iters = 2000;
    cudaMalloc(&out, sizeof(double) * sms * warps_per_sm * 32);
    cudaMalloc(&cyc, 8);
    k<ILP, MODE><<<sms, warps_per_sm * 32>>>(out, iters, cyc);
    k<ILP, MODE><<<sms, warps_per_sm * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_sched_warps = warps_per_sm / 4.0;
    double n = (double)iters * ILP * (MODE == 2 ? 1 : 1);
    printf("%-10s ILP=%d warps/SM=%2d: %.2f cycles per %s per warp -> %.2f cycles per instr per scheduler\n", name, ILP,
           warps_per_sm, h / n, MODE == 2 ? "(1 DMMA + 4 DFMA)" : "instr", h / n / (per_sched_warps < 1 ? 1 : per_sched_warps));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<1, 0>(4, "dmma"); run<2, 0>(4, "dmma"); run<4, 0>(4, "dmma"); run<8, 0>(4, "dmma");
    run<8, 0>(8, "dmma"); run<8, 0>(16, "dmma");
    run<1, 1>(4, "dfma"); run<4, 1>(4, "dfma"); run<8, 1>(4, "dfma"); run<8, 1>(8, "dfma"); run<8, 1>(16, "dfma");
    run<4, 2>(4, "mixed"); run<8, 2>(4, "mixed"); run<8, 2>(8, "mixed"); run<8, 2>(16, "mixed");
    return 0;
}
