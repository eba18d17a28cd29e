// copy_probe.cu -- platform ceiling of the end-to-end paths: concurrent pinned H2D + D2H on 1/2/4/8 GPUs
// from ONE process (one host thread per device, like libvcb200's multi-device mode).
//   copy_probe [MB per direction per device = 200] [reps = 10] [numa = 0|1]
// numa = 1: every worker thread first binds itself to the CPUs local to its GPU
// (/sys/bus/pci/devices/<bdf>/local_cpulist), so that its page-locked buffers are first-touched there.
// Prints one JSON object per device count.
#include <cuda_runtime.h>
#include <sched.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static bool bind_local(int dev) {
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof(bdf), dev) != cudaSuccess) return false;
    for (char* c = bdf; *c; ++c) *c = (char)tolower(*c);
    std::string path = std::string("/sys/bus/pci/devices/") + bdf + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return false;
    char buf[512] = {0};
    const bool ok = fgets(buf, sizeof(buf), f) != nullptr;
    fclose(f);
    if (!ok) return false;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b; ++c) CPU_SET(c, &set); }
        else if (sscanf(tok, "%d", &a) == 1) CPU_SET(a, &set);
    }
    return sched_setaffinity(0, sizeof(set), &set) == 0;
}

int main(int argc, char** argv) {
    const size_t mb = argc > 1 ? atol(argv[1]) : 200;
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    const int numa = argc > 3 ? atoi(argv[3]) : 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { fprintf(stderr, "no CUDA device\n"); return 1; }
    const size_t bytes = mb << 20;
    for (int n = 1; n <= ndev; n *= 2) {
        std::atomic<int> ready{0}, go{0};
        std::vector<double> secs(n, 0.0);
        std::vector<int> bound(n, 0);
        std::vector<std::thread> th;
        for (int d = 0; d < n; ++d)
            th.emplace_back([&, d] {
                cudaSetDevice(d);
                if (numa) bound[d] = bind_local(d) ? 1 : 0;
                void *hin, *hout, *din, *dout;
                cudaHostAlloc(&hin, bytes, cudaHostAllocPortable);
                cudaHostAlloc(&hout, bytes, cudaHostAllocPortable);
                memset(hin, 1, bytes);
                memset(hout, 0, bytes);
                cudaMalloc(&din, bytes);
                cudaMalloc(&dout, bytes);
                cudaStream_t s0, s1;
                cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking);
                cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
                cudaMemcpyAsync(din, hin, bytes, cudaMemcpyHostToDevice, s0);
                cudaMemcpyAsync(hout, dout, bytes, cudaMemcpyDeviceToHost, s1);
                cudaDeviceSynchronize();
                ready.fetch_add(1);
                while (go.load() == 0) std::this_thread::yield();
                const auto t0 = std::chrono::steady_clock::now();
                for (int r = 0; r < reps; ++r) {
                    cudaMemcpyAsync(din, hin, bytes, cudaMemcpyHostToDevice, s0);
                    cudaMemcpyAsync(hout, dout, bytes, cudaMemcpyDeviceToHost, s1);
                }
                cudaStreamSynchronize(s0);
                cudaStreamSynchronize(s1);
                secs[d] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                cudaFreeHost(hin); cudaFreeHost(hout); cudaFree(din); cudaFree(dout);
            });
        while (ready.load() < n) std::this_thread::yield();
        go.store(1);
        for (auto& t : th) t.join();
        double worst = 0, sum = 0;
        for (int d = 0; d < n; ++d) { worst = secs[d] > worst ? secs[d] : worst; sum += (double)bytes * reps / secs[d]; }
        printf("{\"gpus\": %d, \"mb_per_direction\": %zu, \"reps\": %d, \"numa_bind\": %d, \"per_gpu_gbs_each_direction\": [", n, mb, reps, numa);
        for (int d = 0; d < n; ++d) printf("%s%.1f", d ? ", " : "", (double)bytes * reps / secs[d] / 1e9);
        printf("], \"aggregate_gbs_each_direction\": %.1f, \"aggregate_gbs_slowest_rank\": %.1f}\n", sum / 1e9,
               (double)bytes * reps * n / worst / 1e9);
        fflush(stdout);
    }
    return 0;
}
