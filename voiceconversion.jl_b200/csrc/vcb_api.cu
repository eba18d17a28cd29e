// vcb_api.cu -- the C ABI of libvcb200.so (include/vcb200.h): argument checks, handle lifetime,
// host<->device staging and kernel selection.  No exception leaves this file.
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>

#include "vcb_kernels.h"

namespace vcb {

static thread_local std::string t_last_error;
std::atomic<int64_t> g_launches{0};
std::atomic<int> g_variant{0};

// multi-device mode (vcb_init): the devices the host batch entry points shard over
static std::mutex g_devset_mu;
static std::vector<int> g_devset;       // empty or one entry: single-device mode
static std::vector<int> device_set() {
    std::lock_guard<std::mutex> lk(g_devset_mu);
    return g_devset;
}

namespace {
// Ring of per-call event sets: the marks of up to kCalls consecutive calls stay readable, so a
// benchmark loop can enqueue its steps back to back and read the stage times afterwards.
struct StageTimer {
    std::atomic<bool> on{false};
    static constexpr int kMax = 8, kCalls = 64;
    cudaEvent_t ev[kCalls][kMax] = {};
    int n[kCalls] = {};
    int64_t calls = 0;      // calls recorded since timing was enabled
    int device = -1;
} g_stages;
}  // namespace

void stage_begin(cudaStream_t st) {
    if (!g_stages.on.load(std::memory_order_relaxed)) return;
    int dev = 0;
    cudaGetDevice(&dev);
    if (g_stages.device != dev) {      // events belong to a device: (re)create them here
        for (auto& set : g_stages.ev)
            for (auto& e : set) {
                if (e) cudaEventDestroy(e);
                cudaEventCreate(&e);
            }
        g_stages.device = dev;
    }
    const int slot = (int)(g_stages.calls++ % StageTimer::kCalls);
    g_stages.n[slot] = 0;
    stage_mark(st);
}
void stage_mark(cudaStream_t st) {
    if (!g_stages.on.load(std::memory_order_relaxed) || g_stages.calls == 0 || g_stages.device < 0) return;
    const int slot = (int)((g_stages.calls - 1) % StageTimer::kCalls);
    if (g_stages.n[slot] >= StageTimer::kMax) return;
    cudaEventRecord(g_stages.ev[slot][g_stages.n[slot]++], st);
}

int32_t fail(int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_last_error = buf;
    return code;
}

// One-time per-device setup: require sm_100, keep the stream-ordered pool's memory.
static int32_t ensure_device(int* dev_out = nullptr) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return fail(VCB_ECUDA, "no CUDA device available (%s); libvcb200 has no CPU path", cudaGetErrorString(e));
    static std::mutex mu;
    static bool ready[64] = {false};
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 64 && !ready[dev]) {
        cudaDeviceProp prop;
        VCB_CUDA(cudaGetDeviceProperties(&prop, dev));
        if (prop.major != 10)
            return fail(VCB_ECUDA, "device %d (%s, sm_%d%d) is not a Blackwell sm_100 GPU; libvcb200 is built for sm_100a only",
                        dev, prop.name, prop.major, prop.minor);
        cudaMemPool_t pool;
        VCB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = UINT64_MAX;
        VCB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        ready[dev] = true;
    }
    if (dev_out) *dev_out = dev;
    return VCB_OK;
}

// Makes the handle's device current for the duration of one library call and puts the caller's
// device back on return (a destroy running inside a garbage collector must not move torch's or
// Julia's current device).
struct DeviceGuard {
    int prev = -1;
    bool moved = false;
    int32_t enter(int handle_dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != handle_dev) {
            VCB_CUDA(cudaSetDevice(handle_dev));
            moved = true;
        }
        return VCB_OK;
    }
    ~DeviceGuard() {
        if (moved && prev >= 0) cudaSetDevice(prev);
    }
};
#define VCB_ON_DEVICE(dev) DeviceGuard _dg; VCB_TRY(_dg.enter(dev))

// Reads and clears the trajectory handle's pivot flag (stream already synchronised or ordered).
static int32_t traj_take_status(const vcb_traj& t, cudaStream_t st, bool* not_pd) {
    int h = 0;
    VCB_CUDA(cudaMemcpyAsync(&h, t.d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    VCB_CUDA(cudaStreamSynchronize(st));
    if (h) VCB_CUDA(cudaMemsetAsync(t.d_err.p, 0, sizeof(int), st));
    *not_pd = h != 0;
    return VCB_OK;
}
static int32_t traj_not_pd() {
    return fail(VCB_ENOTPD, "W' D^-1 W is not positive definite: some Dy[:,:,m] = (Syy - A Sxy)^-1 is indefinite "
                            "(src/trajectory_gmmmap.jl:24-28, :105); the band Cholesky cannot solve it");
}

// Scratch that lives for one host call.
struct Scratch {
    std::vector<void*> ptrs;
    cudaStream_t st;
    explicit Scratch(cudaStream_t s) : st(s) {}
    ~Scratch() { for (void* p : ptrs) cudaFreeAsync(p, st); }
    template <class T>
    cudaError_t get(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, std::max<size_t>(count, 1) * sizeof(T), st);
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = static_cast<T*>(p);
        return e;
    }
};

static bool use_tc(const vcb_gmmmap& g, bool convert) {
    const int v = g_variant.load();
    if (v == 1) return false;
    return tc_supported(g, convert);
}

static int32_t convert_device(const vcb_gmmmap& g, const double* dX, int64_t T, int64_t ldx,
                              double* dY, int64_t ldy, bool copy_power, cudaStream_t st) {
    if (use_tc(g, true)) return tc_convert(g, dX, T, ldx, dY, ldy, copy_power, st);
    if (g_variant.load() == 2) return fail(VCB_EUNSUPPORTED, "tcgen05 kernel does not support this model shape (D=%d, M=%d)", g.D, g.M);
    return simt_convert(g, dX, T, ldx, dY, ldy, copy_power, st);
}

static int32_t argmax_device(const vcb_gmmmap& g, const double* dX, int64_t T, int64_t ldx,
                             int32_t* d_mhat, cudaStream_t st) {
    if (use_tc(g, false)) return tc_argmax(g, dX, T, ldx, d_mhat, st);
    if (g_variant.load() == 2) return fail(VCB_EUNSUPPORTED, "tcgen05 kernel does not support this model shape (D=%d, M=%d)", g.D, g.M);
    return simt_argmax(g, dX, T, ldx, d_mhat, st);
}

// Chunk list of vc(c::TrajectoryConverter, fm) (src/common.jl:42-57) for a ragged batch.
static void build_chunks(const int64_t* offsets, int64_t nseq, int chunk_limit, std::vector<int64_t>& chunks) {
    chunks.clear();
    chunks.push_back(offsets[0]);
    for (int64_t s = 0; s < nseq; ++s) {
        const int64_t b = offsets[s], e = offsets[s + 1];
        if (chunk_limit <= 0) {
            if (e > b) chunks.push_back(e);
        } else {
            for (int64_t c = b + chunk_limit; c < e; c += chunk_limit) chunks.push_back(c);
            if (e > b) chunks.push_back(e);
        }
    }
}

struct GvRun { const vcb_trajgv* gv = nullptr; int epochs = 0; double alpha = 0.0; };

static int32_t traj_device(const vcb_traj& t, const double* dX, int64_t ldx, const int64_t* offsets,
                           int64_t nseq, int chunk_limit, double* dY, int64_t ldy, int64_t* dmhat,
                           double* dEy, bool copy_power, cudaStream_t st, GvRun gvr = GvRun()) {
    const vcb_gmmmap& g = *t.g;
    if (nseq <= 0) return VCB_OK;
    const int64_t base = offsets[0], total = offsets[nseq] - base;
    for (int64_t s = 0; s < nseq; ++s)
        if (offsets[s + 1] < offsets[s]) return fail(VCB_EARG, "offsets must be non-decreasing");
    if (total == 0) return VCB_OK;
    std::vector<int64_t> chunks;
    build_chunks(offsets, nseq, chunk_limit, chunks);
    for (auto& c : chunks) c -= base;
    const int64_t nchunks = (int64_t)chunks.size() - 1;
    int maxlen = 0;
    for (int64_t c = 0; c < nchunks; ++c) maxlen = (int)std::max<int64_t>(maxlen, chunks[c + 1] - chunks[c]);
    Scratch sc(st);
    int32_t* d_mhat = nullptr;
    int64_t* d_chunks = nullptr;
    if (gvr.gv) {
        // var(y, 2) of a one-frame chunk is NaN in the reference, whose @assert then fires (:165)
        for (int64_t c = 0; c < nchunks; ++c)
            if (chunks[c + 1] - chunks[c] < 2) return fail(VCB_EARG, "GV conversion needs at least 2 frames per chunk");
        if (!dEy) { VCB_CUDA(sc.get(&dEy, (size_t)total * g.D)); dEy -= base * g.D; }
        if (!dmhat) { VCB_CUDA(sc.get(&dmhat, (size_t)total)); dmhat -= base; }
    }
    VCB_CUDA(sc.get(&d_mhat, (size_t)total));
    VCB_CUDA(sc.get(&d_chunks, chunks.size()));
    VCB_CUDA(cudaMemcpyAsync(d_chunks, chunks.data(), chunks.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    const double* X0 = dX + base * ldx;
    stage_begin(st);
    VCB_TRY(argmax_device(g, X0, total, ldx, d_mhat, st));                       // src/trajectory_gmmmap.jl:82
    stage_mark(st);      // [0] arg-max (+ Float64 re-check)
    // bucket the frames by arg-max mixture once; E_t / P_t E_t and the GV gradient run as per-mixture GEMMs
    static const bool per_frame = [] { const char* e = getenv("VCB_TRAJ_E"); return e && e[0] == 'f'; }();
    int* ws = nullptr;
    int64_t npanels = 0;
    if (!per_frame && total < (int64_t)1 << 31 && g.D <= 256 && g.M <= 4096) {
        VCB_CUDA(sc.get(&ws, group_workspace_ints(g.M, total)));
        VCB_TRY(group_frames_by_mixture(d_mhat, total, g.M, ws, &npanels, st));
    }
    stage_mark(st);      // [1] bucketing by mixture; traj_solve_device marks [2] E / PE and [3] the band solver
    VCB_TRY(traj_solve_device(t, X0, ldx, d_mhat, d_chunks, nchunks, maxlen, total, dY + base * ldy, ldy,
                              dEy ? dEy + base * g.D : nullptr, copy_power, st, ws, npanels));
    if (dmhat) VCB_TRY(widen_mhat(d_mhat, total, dmhat + base, st));
    if (gvr.gv)      // src/trajectory_gmmmap.jl:150-171
        VCB_TRY(trajgv_ascent_device(*gvr.gv, dY + base * ldy, ldy, dEy + base * g.D, dmhat + base, d_chunks, nchunks,
                                     total, gvr.epochs, gvr.alpha, st, ws, npanels));
    return VCB_OK;
}

}  // namespace vcb

using namespace vcb;

#define VCB_GUARD_BEGIN try {
#define VCB_GUARD_END                                                        \
    }                                                                        \
    catch (const std::bad_alloc&) { return fail(VCB_ENOMEM, "out of host memory"); } \
    catch (...) { return fail(VCB_EARG, "unexpected internal error"); }

extern "C" {

int32_t vcb_version(void) { return VCB_VERSION; }

int32_t vcb_last_error(char* buf, size_t buflen) {
    if (!buf || buflen == 0) return VCB_EARG;
    std::strncpy(buf, t_last_error.c_str(), buflen - 1);
    buf[buflen - 1] = '\0';
    return VCB_OK;
}

int32_t vcb_device_count(int32_t* count) {
    if (!count) return fail(VCB_EARG, "null count");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return fail(VCB_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *count = n;
    return VCB_OK;
}

int32_t vcb_init(int32_t ndev) {
    VCB_GUARD_BEGIN
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1)
        return fail(VCB_ECUDA, "no CUDA device available (%s); libvcb200 has no CPU path", cudaGetErrorString(e));
    if (ndev < 0 || ndev > n) return fail(VCB_EARG, "vcb_init(%d): %d device(s) visible", ndev, n);
    if (ndev == 0) ndev = n;
    int prev = -1;
    cudaGetDevice(&prev);
    std::vector<int> set;
    for (int d = 0; d < ndev; ++d) {
        VCB_CUDA(cudaSetDevice(d));
        VCB_TRY(ensure_device());
        set.push_back(d);
    }
    if (prev >= 0) cudaSetDevice(prev);
    std::lock_guard<std::mutex> lk(g_devset_mu);
    g_devset = set;
    return VCB_OK;
    VCB_GUARD_END
}

int32_t vcb_num_devices(int32_t* n) {
    if (!n) return fail(VCB_EARG, "null argument");
    const size_t k = device_set().size();
    *n = k > 1 ? (int32_t)k : 1;
    return VCB_OK;
}

int32_t vcb_set_device(int32_t device) {
    VCB_CUDA(cudaSetDevice(device));
    return ensure_device();
}

int32_t vcb_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(VCB_EARG, "null ptr");
    VCB_CUDA(cudaMallocHost(ptr, bytes));
    return VCB_OK;
}
int32_t vcb_host_free(void* ptr) { VCB_CUDA(cudaFreeHost(ptr)); return VCB_OK; }
int32_t vcb_host_register(void* ptr, size_t bytes) { VCB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault)); return VCB_OK; }
int32_t vcb_host_unregister(void* ptr) { VCB_CUDA(cudaHostUnregister(ptr)); return VCB_OK; }

int32_t vcb_set_kernel_variant(int32_t variant) {
    if (variant < 0 || variant > 2) return fail(VCB_EARG, "variant must be 0, 1 or 2");
    g_variant.store(variant);
    return VCB_OK;
}
int64_t vcb_launch_count(void) { return g_launches.load(); }

int32_t vcb_stage_timing(int32_t enable) {
    g_stages.on.store(enable != 0);
    g_stages.calls = 0;
    return VCB_OK;
}
int32_t vcb_stage_times(double* ms, int32_t capacity, int32_t* count) {
    if (!ms || !count || capacity < 0) return fail(VCB_EARG, "null argument");
    *count = 0;
    const int64_t ncalls = std::min<int64_t>(g_stages.calls, StageTimer::kCalls);
    if (ncalls == 0) return VCB_OK;
    const int last = (int)((g_stages.calls - 1) % StageTimer::kCalls);
    const int marks = g_stages.n[last];
    if (marks < 2) return VCB_OK;
    VCB_CUDA(cudaEventSynchronize(g_stages.ev[last][marks - 1]));
    const int nst = std::min(marks - 1, (int)capacity);
    for (int i = 0; i < nst; ++i) ms[i] = 0.0;
    int used = 0;
    for (int64_t c = 0; c < ncalls; ++c) {
        const int slot = (int)((g_stages.calls - 1 - c) % StageTimer::kCalls);
        if (g_stages.n[slot] != marks) continue;     // a call of another kind (e.g. DTW between trajectories)
        for (int i = 0; i < nst; ++i) {
            float f = 0.f;
            VCB_CUDA(cudaEventElapsedTime(&f, g_stages.ev[slot][i], g_stages.ev[slot][i + 1]));
            ms[i] += f;
        }
        ++used;
    }
    for (int i = 0; i < nst; ++i) ms[i] /= used;
    *count = nst;
    return VCB_OK;
}

// ------------------------------------------------------------------------------------------------
// GMMMap
// ------------------------------------------------------------------------------------------------
int32_t vcb_gmmmap_create(const double* weights, const double* mu, const double* sigma, int32_t twoD,
                          int32_t M, int32_t swap, vcb_gmmmap** out) {
    VCB_GUARD_BEGIN
    if (!out) return fail(VCB_EARG, "null out");
    *out = nullptr;
    int dev = 0;
    VCB_TRY(ensure_device(&dev));
    vcb_gmmmap* g = new vcb_gmmmap();
    g->device = dev;
    int32_t rc = build_gmmmap(weights, mu, sigma, twoD, M, swap, *g);
    if (rc != VCB_OK) { delete g; return rc; }
    *out = g;
    return VCB_OK;
    VCB_GUARD_END
}

int32_t vcb_gmmmap_destroy(vcb_gmmmap* g) {
    if (!g) return VCB_OK;
    DeviceGuard dg;
    dg.enter(g->device);
    delete g;
    return VCB_OK;
}

int32_t vcb_gmmmap_dim(const vcb_gmmmap* g, int32_t* dim) {
    if (!g || !dim) return fail(VCB_EARG, "null argument");
    *dim = g->D;
    return VCB_OK;
}
int32_t vcb_gmmmap_ncomponents(const vcb_gmmmap* g, int32_t* M) {
    if (!g || !M) return fail(VCB_EARG, "null argument");
    *M = g->M;
    return VCB_OK;
}

int32_t vcb_gmmmap_get_param(const vcb_gmmmap* g, int32_t which, double* out) {
    if (!g || !out) return fail(VCB_EARG, "null argument");
    const std::vector<double>* src = nullptr;
    std::vector<double> tmp;
    const int D = g->D, M = g->M;
    switch (which) {
        case 0: case 1: {  // stored [M][D] == column-major (D, M)
            src = which == 0 ? &g->mux : &g->muy;
            break;
        }
        case 2: src = &g->A; break;
        case 3: src = &g->Sxx; break;
        case 4: src = &g->Sxy; break;
        case 5: src = &g->Syx; break;
        case 6: src = &g->Syy; break;
        case 7: src = &g->w; break;
        default: return fail(VCB_EARG, "unknown parameter id %d", which);
    }
    (void)D; (void)M;
    std::memcpy(out, src->data(), src->size() * sizeof(double));
    return VCB_OK;
}

int32_t vcb_gmmmap_convert_dev(const vcb_gmmmap* g, const double* dX, int32_t xrows, int64_t T,
                               int64_t ldx, double* dY, int64_t ldy, void* stream) {
    VCB_GUARD_BEGIN
    if (!g || (T > 0 && (!dX || !dY))) return fail(VCB_EARG, "null argument");
    if (xrows != g->D) return fail(VCB_EDIM, "Inconsistent dimentions. (frame has %d rows, dim(g) = %d)", xrows, g->D);
    if (T < 0 || ldx < xrows || ldy < xrows) return fail(VCB_EARG, "bad T/ld (T=%lld ldx=%lld ldy=%lld)", (long long)T, (long long)ldx, (long long)ldy);
    VCB_ON_DEVICE(g->device);
    return convert_device(*g, dX, T, ldx, dY, ldy, false, (cudaStream_t)stream);
    VCB_GUARD_END
}

int32_t vcb_gmmmap_vc_dev(const vcb_gmmmap* g, const double* dfm, int32_t rows, int64_t T, double* dout,
                          void* stream) {
    VCB_GUARD_BEGIN
    if (!g || (T > 0 && (!dfm || !dout))) return fail(VCB_EARG, "null argument");
    if (rows != g->D + 1) return fail(VCB_EDIM, "Inconsistent dimentions. (feature matrix has %d rows, expected 1 + dim(g) = %d)", rows, g->D + 1);
    if (T < 0) return fail(VCB_EARG, "negative T");
    VCB_ON_DEVICE(g->device);
    return convert_device(*g, dfm + 1, T, rows, dout + 1, rows, true, (cudaStream_t)stream);
    VCB_GUARD_END
}

// Host pipeline shared by convert / vc: frames are cut into slices that rotate through NSLOT
// (stream, device-in, device-out) slots so H2D, kernel and D2H of neighbouring slices overlap.
// The transfer is the whole cost (1 M frames: 200 MB each way against a 0.7 ms kernel), so the
// schedule is built around PCIe: the first and last slices are small and double / halve (the
// pipeline fills and drains in the time of a 16 Ki-frame copy instead of a 128 Ki-frame one), and
// streams and staging buffers are kept between calls (one context per device and concurrent caller).
namespace {
constexpr int kSlots = 4;
struct HostPipe {
    int device = -1;
    size_t in_elems = 0, out_elems = 0;
    cudaStream_t st[kSlots] = {};
    double* din[kSlots] = {};
    double* dout[kSlots] = {};
    // page-locked bounce buffer (grow-only) for small results bound for pageable caller memory: an async
    // copy into pageable memory blocks the calling thread until the stream has drained, which would
    // serialise the slices of the host pipelines
    void* bounce = nullptr;
    size_t bounce_bytes = 0;
    void* bounce_for(size_t bytes) {
        if (bytes > bounce_bytes) {
            if (bounce) cudaFreeHost(bounce);
            bounce = nullptr;
            bounce_bytes = 0;
            if (cudaHostAlloc(&bounce, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            bounce_bytes = bytes;
        }
        return bounce;
    }
};
// true if `p` is page-locked (allocated or registered with CUDA): async copies to it do not block the caller
static bool is_pinned(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
std::mutex g_pipe_mu;
std::vector<HostPipe*> g_pipe_pool;

HostPipe* pipe_acquire(int device, size_t in_elems, size_t out_elems) {
    HostPipe* p = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_pipe_mu);
        for (size_t i = 0; i < g_pipe_pool.size(); ++i)
            if (g_pipe_pool[i]->device == device) { p = g_pipe_pool[i]; g_pipe_pool.erase(g_pipe_pool.begin() + i); break; }
    }
    if (!p) {
        p = new HostPipe();
        p->device = device;
        for (int s = 0; s < kSlots; ++s)
            if (cudaStreamCreateWithFlags(&p->st[s], cudaStreamNonBlocking) != cudaSuccess) { delete p; return nullptr; }
    }
    if (p->in_elems < in_elems || p->out_elems < out_elems) {
        for (int s = 0; s < kSlots; ++s) {
            if (p->din[s]) cudaFree(p->din[s]);
            if (p->dout[s]) cudaFree(p->dout[s]);
            p->din[s] = p->dout[s] = nullptr;
        }
        p->in_elems = p->out_elems = 0;
        for (int s = 0; s < kSlots; ++s)
            if (cudaMalloc((void**)&p->din[s], in_elems * sizeof(double)) != cudaSuccess ||
                cudaMalloc((void**)&p->dout[s], out_elems * sizeof(double)) != cudaSuccess) {
                for (int q = 0; q < kSlots; ++q) { if (p->din[q]) cudaFree(p->din[q]); if (p->dout[q]) cudaFree(p->dout[q]); p->din[q] = p->dout[q] = nullptr; }
                std::lock_guard<std::mutex> lk(g_pipe_mu);
                g_pipe_pool.push_back(p);
                return nullptr;
            }
        p->in_elems = in_elems;
        p->out_elems = out_elems;
    }
    return p;
}
void pipe_release(HostPipe* p) {
    std::lock_guard<std::mutex> lk(g_pipe_mu);
    g_pipe_pool.push_back(p);
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// Multi-device mode (vcb_init): the HOST-pointer batch entry points shard their batch over the
// devices of the set -- contiguous ranges balanced by cost, one host thread + pipeline per device,
// the model replicated on every device by K0, no data-path exchange.
// ------------------------------------------------------------------------------------------------

// Runs job(part, device) for every part on its own host thread with that device current (part 0 on the
// calling thread) and returns the first failure; worker threads hand their error message to the caller.
static int32_t run_on_devices(const std::vector<int>& devs, const std::function<int32_t(int, int)>& job) {
    const int n = (int)devs.size();
    std::vector<int32_t> rc(n, VCB_OK);
    std::vector<std::string> msg(n);
    auto body = [&](int part) {
        if (cudaSetDevice(devs[part]) != cudaSuccess) {
            rc[part] = VCB_ECUDA;
            msg[part] = "cudaSetDevice failed in multi-device mode";
            return;
        }
        rc[part] = ensure_device();
        if (rc[part] == VCB_OK) {
            try {
                rc[part] = job(part, devs[part]);
            } catch (const std::bad_alloc&) {
                rc[part] = fail(VCB_ENOMEM, "out of host memory");
            } catch (...) {
                rc[part] = fail(VCB_EARG, "unexpected internal error");
            }
        }
        if (rc[part] != VCB_OK) msg[part] = t_last_error;
    };
    int prev = -1;
    cudaGetDevice(&prev);
    std::vector<std::thread> th;
    for (int part = 1; part < n; ++part) th.emplace_back(body, part);
    body(0);
    for (auto& t : th) t.join();
    if (prev >= 0) cudaSetDevice(prev);
    for (int part = 0; part < n; ++part)
        if (rc[part] != VCB_OK) {
            t_last_error = msg[part];
            return rc[part];
        }
    return VCB_OK;
}

// The handle to use on `device`: the handle itself on its own device, else its replica (built once, by the
// calling worker thread, with `device` current).
static const vcb_gmmmap* gmmmap_on_device(const vcb_gmmmap& g, int device, int32_t* rc) {
    *rc = VCB_OK;
    if (device == g.device) return &g;
    std::lock_guard<std::mutex> lk(g.rep_mu);
    if ((int)g.replicas.size() <= device) g.replicas.resize(device + 1, nullptr);
    if (!g.replicas[device]) {
        vcb_gmmmap* r = new vcb_gmmmap();
        r->device = device;
        *rc = build_gmmmap(g.src_w.data(), g.src_mu.data(), g.src_sigma.data(), 2 * g.D, g.M, g.src_swap, *r);
        if (*rc != VCB_OK) { delete r; return nullptr; }
        r->src_w.clear(); r->src_mu.clear(); r->src_sigma.clear();
        r->src_w.shrink_to_fit(); r->src_mu.shrink_to_fit(); r->src_sigma.shrink_to_fit();
        g.replicas[device] = r;
    }
    return g.replicas[device];
}
static const vcb_traj* traj_on_device(const vcb_traj& t, int device, int32_t* rc) {
    *rc = VCB_OK;
    if (device == t.device) return &t;
    const vcb_gmmmap* gr = gmmmap_on_device(*t.g, device, rc);
    if (!gr) return nullptr;
    std::lock_guard<std::mutex> lk(t.rep_mu);
    if ((int)t.replicas.size() <= device) t.replicas.resize(device + 1, nullptr);
    if (!t.replicas[device]) {
        vcb_traj* r = new vcb_traj();
        *rc = build_traj(*gr, *r);
        if (*rc != VCB_OK) { delete r; return nullptr; }
        t.replicas[device] = r;
    }
    return t.replicas[device];
}

// Contiguous split of n units into `parts` ranges of near-equal total cost (cost(i) >= 0).
static std::vector<int64_t> split_by_cost(int64_t n, int parts, const std::function<double(int64_t)>& cost) {
    std::vector<double> pre(n + 1, 0.0);
    for (int64_t i = 0; i < n; ++i) pre[i + 1] = pre[i] + cost(i);
    std::vector<int64_t> b(parts + 1, 0);
    b[parts] = n;
    for (int r = 1; r < parts; ++r) {
        const double target = pre[n] * r / parts;
        int64_t i = std::lower_bound(pre.begin(), pre.end(), target) - pre.begin();
        i = std::min<int64_t>(std::max<int64_t>(i, b[r - 1]), n);
        if (i > b[r - 1] && std::fabs(pre[i - 1] - target) <= std::fabs(pre[i] - target)) --i;
        b[r] = i;
    }
    return b;
}

static int32_t fbf_host(const vcb_gmmmap& g, const double* X, int64_t T, int64_t ldx, double* Y,
                        int64_t ldy, bool whole_rows) {
    if (T == 0) return VCB_OK;
    // 128 Ki-frame slices measured best on B200 in steady state (VCB_SLICE sweep); VCB_SLICE_MIN is
    // the size the ramps start from / end at.
    static const int64_t slice_frames = [] { const char* e = getenv("VCB_SLICE"); return e ? atoll(e) : 131072LL; }();
    static const int64_t ramp_frames = [] { const char* e = getenv("VCB_SLICE_MIN"); return e ? atoll(e) : 16384LL; }();
    const int64_t slice = std::min<int64_t>(T, std::max<int64_t>(slice_frames, 128));
    const int64_t ramp = std::min<int64_t>(slice, std::max<int64_t>(ramp_frames, 128));
    // whole_rows (vc): X/Y point at row 1 of (rows, T) matrices with ld == rows; the copies move
    // complete columns starting one double earlier (the power row).
    const int64_t pre = whole_rows ? 1 : 0;
    HostPipe* hp = pipe_acquire(g.device, (size_t)slice * ldx, (size_t)slice * ldy);
    if (!hp) return fail(VCB_ECUDA, "staging allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    int32_t rc = VCB_OK;
    int idx = 0;
    int64_t next = ramp;        // size of the next slice while ramping up
    for (int64_t b = 0; b < T && rc == VCB_OK; idx = (idx + 1) % kSlots) {
        const int64_t left = T - b;
        // ramp up by doubling; ramp down by halving what is left once less than two full slices remain
        int64_t n = std::min(next, slice);
        if (left <= 2 * n) n = (left > 2 * ramp || left > slice) ? (left + 1) / 2 : left;
        n = std::min(n, std::min(left, slice));
        next = std::min(slice, next * 2);
        const size_t in_elems = (size_t)(n - 1) * ldx + g.D + pre;
        const size_t out_elems = (size_t)(n - 1) * ldy + g.D + pre;
        cudaStream_t s = hp->st[idx];
        double *din = hp->din[idx], *dout = hp->dout[idx];
        if (cudaMemcpyAsync(din, X + b * ldx - pre, in_elems * sizeof(double), cudaMemcpyHostToDevice, s) != cudaSuccess) {
            rc = fail(VCB_ECUDA, "H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        rc = convert_device(g, din + pre, n, ldx, dout + pre, ldy, whole_rows, s);
        if (rc != VCB_OK) break;
        if (whole_rows || ldy == g.D) {
            if (cudaMemcpyAsync(Y + b * ldy - pre, dout, out_elems * sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess)
                rc = fail(VCB_ECUDA, "D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        } else {
            // strided output: only the D converted rows of each column belong to the caller
            if (cudaMemcpy2DAsync(Y + b * ldy, ldy * sizeof(double), dout, ldy * sizeof(double),
                                  g.D * sizeof(double), n, cudaMemcpyDeviceToHost, s) != cudaSuccess)
                rc = fail(VCB_ECUDA, "D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
        b += n;
    }
    for (int s = 0; s < kSlots; ++s)
        if (cudaStreamSynchronize(hp->st[s]) != cudaSuccess && rc == VCB_OK)
            rc = fail(VCB_ECUDA, "conversion failed: %s", cudaGetErrorString(cudaGetLastError()));
    pipe_release(hp);
    return rc;
}

// Frames shard contiguously over the device set (src/common.jl:17-19: no state is carried from frame to
// frame); small calls stay on the handle's device.
static int32_t fbf_host_multi(const vcb_gmmmap& g, const double* X, int64_t T, int64_t ldx, double* Y, int64_t ldy,
                              bool whole_rows) {
    const std::vector<int> devs = device_set();
    const int n = (int)devs.size();
    if (n <= 1 || T < (int64_t)n * 65536) {
        VCB_ON_DEVICE(g.device);
        return fbf_host(g, X, T, ldx, Y, ldy, whole_rows);
    }
    return run_on_devices(devs, [&](int part, int dev) -> int32_t {
        const int64_t b = T * part / n, e = T * (part + 1) / n;
        int32_t rc;
        const vcb_gmmmap* gd = gmmmap_on_device(g, dev, &rc);
        if (!gd) return rc;
        return fbf_host(*gd, X + b * ldx, e - b, ldx, Y + b * ldy, ldy, whole_rows);
    });
}

int32_t vcb_gmmmap_convert(const vcb_gmmmap* g, const double* X, int32_t xrows, int64_t T, int64_t ldx,
                           double* Y, int64_t ldy) {
    VCB_GUARD_BEGIN
    if (!g || (T > 0 && (!X || !Y))) return fail(VCB_EARG, "null argument");
    if (xrows != g->D) return fail(VCB_EDIM, "Inconsistent dimentions. (frame has %d rows, dim(g) = %d)", xrows, g->D);
    if (T < 0 || ldx < xrows || ldy < xrows) return fail(VCB_EARG, "bad T/ld");
    return fbf_host_multi(*g, X, T, ldx, Y, ldy, false);
    VCB_GUARD_END
}

int32_t vcb_gmmmap_vc(const vcb_gmmmap* g, const double* fm, int32_t rows, int64_t T, double* out) {
    VCB_GUARD_BEGIN
    if (!g || (T > 0 && (!fm || !out))) return fail(VCB_EARG, "null argument");
    if (rows != g->D + 1) return fail(VCB_EDIM, "Inconsistent dimentions. (feature matrix has %d rows, expected 1 + dim(g) = %d)", rows, g->D + 1);
    if (T < 0) return fail(VCB_EARG, "negative T");
    return fbf_host_multi(*g, fm + 1, T, rows, out + 1, rows, true);
    VCB_GUARD_END
}

int32_t vcb_gmmmap_predict_proba(const vcb_gmmmap* g, const double* X, int32_t xrows, int64_t T,
                                 int64_t ldx, double* post) {
    VCB_GUARD_BEGIN
    if (!g || (T > 0 && (!X || !post))) return fail(VCB_EARG, "null argument");
    if (xrows != g->D) return fail(VCB_EDIM, "Inconsistent dimentions. (frame has %d rows, dim(g) = %d)", xrows, g->D);
    if (T < 0 || ldx < xrows) return fail(VCB_EARG, "bad T/ld");
    if (T == 0) return VCB_OK;
    VCB_ON_DEVICE(g->device);
    cudaStream_t st = nullptr;
    Scratch sc(st);
    double *dX = nullptr, *dP = nullptr;
    const size_t in_elems = (size_t)(T - 1) * ldx + xrows;
    VCB_CUDA(sc.get(&dX, in_elems));
    VCB_CUDA(sc.get(&dP, (size_t)T * g->M));
    VCB_CUDA(cudaMemcpyAsync(dX, X, in_elems * sizeof(double), cudaMemcpyHostToDevice, st));
    VCB_TRY(proba_fp64(*g, dX, T, ldx, dP, st));
    VCB_CUDA(cudaMemcpyAsync(post, dP, (size_t)T * g->M * sizeof(double), cudaMemcpyDeviceToHost, st));
    VCB_CUDA(cudaStreamSynchronize(st));
    return VCB_OK;
    VCB_GUARD_END
}

int32_t vcb_gmmmap_predict(const vcb_gmmmap* g, const double* X, int32_t xrows, int64_t T, int64_t ldx,
                           int64_t* mhat) {
    VCB_GUARD_BEGIN
    if (!g || (T > 0 && (!X || !mhat))) return fail(VCB_EARG, "null argument");
    if (xrows != g->D) return fail(VCB_EDIM, "Inconsistent dimentions. (frame has %d rows, dim(g) = %d)", xrows, g->D);
    if (T < 0 || ldx < xrows) return fail(VCB_EARG, "bad T/ld");
    if (T == 0) return VCB_OK;
    VCB_ON_DEVICE(g->device);
    cudaStream_t st = nullptr;
    Scratch sc(st);
    double* dX = nullptr;
    int32_t* dm = nullptr;
    int64_t* dm64 = nullptr;
    const size_t in_elems = (size_t)(T - 1) * ldx + xrows;
    VCB_CUDA(sc.get(&dX, in_elems));
    VCB_CUDA(sc.get(&dm, (size_t)T));
    VCB_CUDA(sc.get(&dm64, (size_t)T));
    VCB_CUDA(cudaMemcpyAsync(dX, X, in_elems * sizeof(double), cudaMemcpyHostToDevice, st));
    VCB_TRY(argmax_device(*g, dX, T, ldx, dm, st));
    VCB_TRY(widen_mhat(dm, T, dm64, st));
    VCB_CUDA(cudaMemcpyAsync(mhat, dm64, (size_t)T * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    VCB_CUDA(cudaStreamSynchronize(st));
    return VCB_OK;
    VCB_GUARD_END
}

// ------------------------------------------------------------------------------------------------
// TrajectoryGMMMap
// ------------------------------------------------------------------------------------------------
int32_t vcb_traj_create(const vcb_gmmmap* g, vcb_traj** out) {
    VCB_GUARD_BEGIN
    if (!g || !out) return fail(VCB_EARG, "null argument");
    *out = nullptr;
    VCB_ON_DEVICE(g->device);
    vcb_traj* t = new vcb_traj();
    int32_t rc = build_traj(*g, *t);
    if (rc != VCB_OK) { delete t; return rc; }
    *out = t;
    return VCB_OK;
    VCB_GUARD_END
}

int32_t vcb_traj_destroy(vcb_traj* t) {
    if (!t) return VCB_OK;
    DeviceGuard dg;
    dg.enter(t->device);     // not t->g->device: the parent may already be gone
    delete t;
    return VCB_OK;
}

int32_t vcb_traj_status(const vcb_traj* t, void* stream) {
    VCB_GUARD_BEGIN
    if (!t) return fail(VCB_EARG, "null argument");
    VCB_ON_DEVICE(t->device);
    bool not_pd = false;
    VCB_TRY(traj_take_status(*t, (cudaStream_t)stream, &not_pd));
    return not_pd ? traj_not_pd() : VCB_OK;
    VCB_GUARD_END
}

int32_t vcb_traj_get_Dy(const vcb_traj* t, double* out) {
    if (!t || !out) return fail(VCB_EARG, "null argument");
    std::memcpy(out, t->Dy.data(), t->Dy.size() * sizeof(double));
    return VCB_OK;
}

static int32_t traj_check(const vcb_traj* t, int32_t xrows, int64_t ldx, const int64_t* offsets, int64_t nseq) {
    if (!t || !offsets) return fail(VCB_EARG, "null argument");
    if (nseq < 0) return fail(VCB_EARG, "negative nseq");
    // src/trajectory_gmmmap.jl:66-68
    if (xrows != t->g->D) return fail(VCB_EDIM, "Inconsistent dimentions. (frame has %d rows, dim(t) = %d)", xrows, t->g->D);
    if (ldx < xrows) return fail(VCB_EARG, "ldx < rows");
    return VCB_OK;
}

int32_t vcb_traj_convert_batch_dev(const vcb_traj* t, const double* dX, int32_t xrows, int64_t ldx,
                                   const int64_t* offsets, int64_t nseq, int32_t chunk_limit, double* dY,
                                   int64_t ldy, int64_t* dmhat, double* dEy, void* stream) {
    VCB_GUARD_BEGIN
    VCB_TRY(traj_check(t, xrows, ldx, offsets, nseq));
    if (!dX || !dY) return fail(VCB_EARG, "null argument");
    if (ldy < t->Ds) return fail(VCB_EARG, "ldy < dim/2");
    VCB_ON_DEVICE(t->device);
    return traj_device(*t, dX, ldx, offsets, nseq, chunk_limit, dY, ldy, dmhat, dEy, false, (cudaStream_t)stream);
    VCB_GUARD_END
}

int32_t vcb_traj_vc_batch_dev(const vcb_traj* t, const double* dfm, int32_t rows, const int64_t* offsets,
                              int64_t nseq, int32_t chunk_limit, double* dout, void* stream) {
    VCB_GUARD_BEGIN
    VCB_TRY(traj_check(t, rows - 1, rows, offsets, nseq));
    if (!dfm || !dout) return fail(VCB_EARG, "null argument");
    VCB_ON_DEVICE(t->device);
    return traj_device(*t, dfm + 1, rows, offsets, nseq, chunk_limit, dout + 1, t->Ds + 1, nullptr, nullptr, true,
                       (cudaStream_t)stream);
    VCB_GUARD_END
}

// One slice of utterances [s0, s1) of a host batch on one stream: H2D, conversion, D2H; no
// synchronisation (the per-call scratch is released in stream order).
static int32_t traj_host_slice(const vcb_traj& t, const double* X, int64_t ldx, const int64_t* offsets, int64_t s0,
                               int64_t s1, int chunk_limit, double* Y, int64_t ldy, int64_t* mhat, double* Ey,
                               bool whole_rows, GvRun gvr, cudaStream_t st) {
    const int64_t base = offsets[s0], total = offsets[s1] - base;
    if (total <= 0) return VCB_OK;
    const int D2 = t.g->D, Ds = t.Ds;
    const int64_t pre = whole_rows ? 1 : 0;
    Scratch sc(st);
    double *dX = nullptr, *dY = nullptr, *dE = nullptr;
    int64_t* dm = nullptr;
    const size_t in_elems = (size_t)(total - 1) * ldx + D2 + pre;
    const size_t out_elems = (size_t)(total - 1) * ldy + Ds + pre;
    VCB_CUDA(sc.get(&dX, in_elems));
    VCB_CUDA(sc.get(&dY, out_elems));
    if (mhat) VCB_CUDA(sc.get(&dm, (size_t)total));
    if (Ey) VCB_CUDA(sc.get(&dE, (size_t)total * D2));
    VCB_CUDA(cudaMemcpyAsync(dX, X + base * ldx - pre, in_elems * sizeof(double), cudaMemcpyHostToDevice, st));
    std::vector<int64_t> rel(offsets + s0, offsets + s1 + 1);
    for (auto& o : rel) o -= base;
    VCB_TRY(traj_device(t, dX + pre, ldx, rel.data(), s1 - s0, chunk_limit, dY + pre, ldy, dm, dE, whole_rows, st, gvr));
    if (whole_rows || ldy == Ds) {
        VCB_CUDA(cudaMemcpyAsync(Y + base * ldy - pre, dY, out_elems * sizeof(double), cudaMemcpyDeviceToHost, st));
    } else {
        VCB_CUDA(cudaMemcpy2DAsync(Y + base * ldy, ldy * sizeof(double), dY, ldy * sizeof(double), Ds * sizeof(double),
                                   total, cudaMemcpyDeviceToHost, st));
    }
    if (mhat) VCB_CUDA(cudaMemcpyAsync(mhat + base, dm, (size_t)total * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (Ey) VCB_CUDA(cudaMemcpyAsync(Ey + base * D2, dE, (size_t)total * D2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    return VCB_OK;
}

// Host batch: the utterances are cut into slices of about one solver wave (one warp per chunk, ~7
// chunks per SM) that rotate through the cached streams, so the H2D copy of slice i+1 and the D2H
// copy of slice i-1 run under the kernels of slice i.  The band solver is latency-bound (its time
// hardly depends on the number of chunks up to a full wave), so smaller slices would not pay.
static int32_t traj_host_one(const vcb_traj& t, const double* X, int64_t ldx, const int64_t* offsets, int64_t nseq,
                             int chunk_limit, double* Y, int64_t ldy, int64_t* mhat, double* Ey, bool whole_rows,
                             GvRun gvr = GvRun()) {
    if (nseq == 0) return VCB_OK;
    if (offsets[0] < 0) return fail(VCB_EARG, "offsets must start at a non-negative frame");
    for (int64_t s = 0; s < nseq; ++s)
        if (offsets[s + 1] < offsets[s]) return fail(VCB_EARG, "offsets must be non-decreasing");
    if (offsets[nseq] - offsets[0] <= 0) return VCB_OK;
    static const int64_t wave = [] { const char* e = getenv("VCB_TRAJ_SLICE_CHUNKS"); return e ? atoll(e) : 1036LL; }();
    HostPipe* hp = pipe_acquire(t.device, 0, 0);
    if (!hp) return fail(VCB_ECUDA, "stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    int32_t rc = VCB_OK;
    int idx = 0;
    for (int64_t s0 = 0; s0 < nseq && rc == VCB_OK; idx = (idx + 1) % kSlots) {
        int64_t s1 = s0, chunks = 0;
        while (s1 < nseq && (chunks < wave || s1 == s0)) {
            const int64_t len = offsets[s1 + 1] - offsets[s1];
            chunks += chunk_limit > 0 ? (len + chunk_limit - 1) / chunk_limit : (len > 0);
            ++s1;
        }
        // do not leave a tail of less than half a wave for a slice of its own
        int64_t rest = 0;
        for (int64_t s = s1; s < nseq && rest < wave; ++s) {
            const int64_t len = offsets[s + 1] - offsets[s];
            rest += chunk_limit > 0 ? (len + chunk_limit - 1) / chunk_limit : (len > 0);
        }
        if (2 * rest < wave) s1 = nseq;
        rc = traj_host_slice(t, X, ldx, offsets, s0, s1, chunk_limit, Y, ldy, mhat, Ey, whole_rows, gvr, hp->st[idx]);
        s0 = s1;
    }
    for (int s = 0; s < kSlots; ++s)
        if (cudaStreamSynchronize(hp->st[s]) != cudaSuccess && rc == VCB_OK)
            rc = fail(VCB_ECUDA, "trajectory conversion failed: %s", cudaGetErrorString(cudaGetLastError()));
    bool not_pd = false;
    if (rc == VCB_OK) rc = traj_take_status(t, hp->st[0], &not_pd);
    pipe_release(hp);
    if (rc == VCB_OK && not_pd) rc = traj_not_pd();
    return rc;
}

// Utterances shard contiguously over the device set, balanced by frames (every chunk is its own linear
// system, src/common.jl:44-57); the GV variant and small batches stay on the handle's device.
static int32_t traj_host(const vcb_traj& t, const double* X, int64_t ldx, const int64_t* offsets, int64_t nseq,
                         int chunk_limit, double* Y, int64_t ldy, int64_t* mhat, double* Ey, bool whole_rows,
                         GvRun gvr = GvRun()) {
    const std::vector<int> devs = device_set();
    const int n = (int)devs.size();
    if (n <= 1 || gvr.gv || nseq < 2 * n || offsets[nseq] - offsets[0] < (int64_t)n * 16384) {
        VCB_ON_DEVICE(t.device);
        return traj_host_one(t, X, ldx, offsets, nseq, chunk_limit, Y, ldy, mhat, Ey, whole_rows, gvr);
    }
    for (int64_t s = 0; s < nseq; ++s)
        if (offsets[s + 1] < offsets[s]) return fail(VCB_EARG, "offsets must be non-decreasing");
    const std::vector<int64_t> cut = split_by_cost(nseq, n, [&](int64_t i) { return (double)(offsets[i + 1] - offsets[i]); });
    return run_on_devices(devs, [&](int part, int dev) -> int32_t {
        const int64_t b = cut[part], e = cut[part + 1];
        if (e <= b) return VCB_OK;
        int32_t rc;
        const vcb_traj* td = traj_on_device(t, dev, &rc);
        if (!td) return rc;
        return traj_host_one(*td, X, ldx, offsets + b, e - b, chunk_limit, Y, ldy, mhat, Ey, whole_rows);
    });
}

int32_t vcb_traj_convert_batch(const vcb_traj* t, const double* X, int32_t xrows, int64_t ldx,
                               const int64_t* offsets, int64_t nseq, int32_t chunk_limit, double* Y,
                               int64_t ldy, int64_t* mhat, double* Ey) {
    VCB_GUARD_BEGIN
    VCB_TRY(traj_check(t, xrows, ldx, offsets, nseq));
    if (!X || !Y) return fail(VCB_EARG, "null argument");
    if (ldy < t->Ds) return fail(VCB_EARG, "ldy < dim/2");
    return traj_host(*t, X, ldx, offsets, nseq, chunk_limit, Y, ldy, mhat, Ey, false);
    VCB_GUARD_END
}

int32_t vcb_traj_vc_batch(const vcb_traj* t, const double* fm, int32_t rows, const int64_t* offsets,
                          int64_t nseq, int32_t chunk_limit, double* out) {
    VCB_GUARD_BEGIN
    VCB_TRY(traj_check(t, rows - 1, rows, offsets, nseq));
    if (!fm || !out) return fail(VCB_EARG, "null argument");
    return traj_host(*t, fm + 1, rows, offsets, nseq, chunk_limit, out + 1, t->Ds + 1, nullptr, nullptr, true);
    VCB_GUARD_END
}

// vc(mapper, [fm[1,:]; push_delta(fm[2:end,:])]) in one call (bin/vc.jl:76-82): the
// delta features are appended on the device, per utterance, before the chunked trajectory conversion.
static int32_t traj_static_device(const vcb_traj& t, const double* dfm, const int64_t* offsets, int64_t nseq,
                                  int chunk_limit, double* dout, cudaStream_t st) {
    const int Ds = t.Ds;
    if (nseq <= 0) return VCB_OK;
    const int64_t base = offsets[0], total = offsets[nseq] - base;
    if (total <= 0) return VCB_OK;
    Scratch sc(st);
    double* dfull = nullptr;
    int64_t* dOff = nullptr;
    VCB_CUDA(sc.get(&dfull, (size_t)total * (2 * Ds + 1)));
    VCB_CUDA(sc.get(&dOff, (size_t)nseq + 1));
    std::vector<int64_t> rel(offsets, offsets + nseq + 1);
    for (auto& o : rel) o -= base;
    VCB_CUDA(cudaMemcpyAsync(dOff, rel.data(), rel.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    const double* src = dfm + base * (Ds + 1);
    VCB_TRY(push_delta_strided_device(src + 1, Ds + 1, Ds, dOff, nseq, total, dfull + 1, 2 * Ds + 1, st));
    VCB_CUDA(cudaMemcpy2DAsync(dfull, (size_t)(2 * Ds + 1) * sizeof(double), src, (size_t)(Ds + 1) * sizeof(double),
                               sizeof(double), (size_t)total, cudaMemcpyDeviceToDevice, st));   // power row
    return traj_device(t, dfull + 1, 2 * Ds + 1, rel.data(), nseq, chunk_limit, dout + base * (Ds + 1) + 1, Ds + 1, nullptr,
                       nullptr, true, st);
}

int32_t vcb_traj_vc_static_batch_dev(const vcb_traj* t, const double* dfm, int32_t rows, const int64_t* offsets,
                                     int64_t nseq, int32_t chunk_limit, double* dout, void* stream) {
    VCB_GUARD_BEGIN
    if (!t || !offsets || !dfm || !dout) return fail(VCB_EARG, "null argument");
    if (nseq < 0) return fail(VCB_EARG, "negative nseq");
    if (rows != t->Ds + 1) return fail(VCB_EDIM, "Inconsistent dimentions. (static frame has %d rows, 1 + dim(t)/2 = %d)", rows, t->Ds + 1);
    VCB_ON_DEVICE(t->device);
    return traj_static_device(*t, dfm, offsets, nseq, chunk_limit, dout, (cudaStream_t)stream);
    VCB_GUARD_END
}

int32_t vcb_traj_vc_static_batch(const vcb_traj* t, const double* fm, int32_t rows, const int64_t* offsets,
                                 int64_t nseq, int32_t chunk_limit, double* out) {
    VCB_GUARD_BEGIN
    if (!t || !offsets || !fm || !out) return fail(VCB_EARG, "null argument");
    if (nseq < 0) return fail(VCB_EARG, "negative nseq");
    if (rows != t->Ds + 1) return fail(VCB_EDIM, "Inconsistent dimentions. (static frame has %d rows, 1 + dim(t)/2 = %d)", rows, t->Ds + 1);
    if (nseq == 0) return VCB_OK;
    VCB_ON_DEVICE(t->device);
    const int64_t base = offsets[0], total = offsets[nseq] - base;
    if (total <= 0) return VCB_OK;
    cudaStream_t st = nullptr;
    Scratch sc(st);
    double *dI = nullptr, *dO = nullptr;
    const size_t n = (size_t)total * rows;
    VCB_CUDA(sc.get(&dI, n));
    VCB_CUDA(sc.get(&dO, n));
    VCB_CUDA(cudaMemcpyAsync(dI, fm + base * rows, n * sizeof(double), cudaMemcpyHostToDevice, st));
    std::vector<int64_t> rel(offsets, offsets + nseq + 1);
    for (auto& o : rel) o -= base;
    VCB_TRY(traj_static_device(*t, dI, rel.data(), nseq, chunk_limit, dO, st));
    VCB_CUDA(cudaMemcpyAsync(out + base * rows, dO, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    bool not_pd = false;
    VCB_TRY(traj_take_status(*t, st, &not_pd));
    return not_pd ? traj_not_pd() : VCB_OK;
    VCB_GUARD_END
}

// ------------------------------------------------------------------------------------------------
// TrajectoryGVGMMMap, VarianceScaling, diffgmm
// ------------------------------------------------------------------------------------------------
int32_t vcb_trajgv_create(const vcb_traj* t, const double* mu_v, const double* sigma_vv, vcb_trajgv** out) {
    VCB_GUARD_BEGIN
    if (!t || !mu_v || !sigma_vv || !out) return fail(VCB_EARG, "null argument");
    *out = nullptr;
    const int Ds = t->Ds;
    for (int i = 0; i < Ds; ++i)      // @assert sum(mu_v .< 0) == 0   src/trajectory_gmmmap.jl:124
        if (mu_v[i] < 0.0) return fail(VCB_EARG, "GV mean %d is negative", i + 1);
    std::vector<double> pv(sigma_vv, sigma_vv + (size_t)Ds * Ds);
    if (!invert_matrix(pv, Ds)) return fail(VCB_ESINGULAR, "GV covariance is singular");   // inv(S_vv)  :125
    VCB_ON_DEVICE(t->device);
    vcb_trajgv* v = new vcb_trajgv();
    v->t = t;
    v->device = t->device;
    v->muv.assign(mu_v, mu_v + Ds);
    v->pv = pv;
    if (v->d_muv.upload(v->muv) != cudaSuccess || v->d_pv.upload(v->pv) != cudaSuccess) {
        delete v;
        return fail(VCB_ECUDA, "device upload of the GV parameters failed");
    }
    *out = v;
    return VCB_OK;
    VCB_GUARD_END
}

int32_t vcb_trajgv_destroy(vcb_trajgv* v) {
    if (!v) return VCB_OK;
    DeviceGuard dg;
    dg.enter(v->device);
    delete v;
    return VCB_OK;
}

static int32_t gv_args(const vcb_trajgv* v, int32_t epochs) {
    if (!v) return fail(VCB_EARG, "null argument");
    if (epochs < 0) return fail(VCB_EARG, "negative epochs");
    return VCB_OK;
}

int32_t vcb_trajgv_convert_batch_dev(const vcb_trajgv* v, const double* dX, int32_t xrows, int64_t ldx,
                                     const int64_t* offsets, int64_t nseq, int32_t chunk_limit, int32_t epochs,
                                     double alpha, double* dY, int64_t ldy, void* stream) {
    VCB_GUARD_BEGIN
    VCB_TRY(gv_args(v, epochs));
    VCB_TRY(traj_check(v->t, xrows, ldx, offsets, nseq));
    if (!dX || !dY) return fail(VCB_EARG, "null argument");
    if (ldy < v->t->Ds) return fail(VCB_EARG, "ldy < dim/2");
    VCB_ON_DEVICE(v->device);
    return traj_device(*v->t, dX, ldx, offsets, nseq, chunk_limit, dY, ldy, nullptr, nullptr, false,
                       (cudaStream_t)stream, GvRun{v, epochs, alpha});
    VCB_GUARD_END
}

int32_t vcb_trajgv_vc_batch_dev(const vcb_trajgv* v, const double* dfm, int32_t rows, const int64_t* offsets,
                                int64_t nseq, int32_t chunk_limit, int32_t epochs, double alpha, double* dout,
                                void* stream) {
    VCB_GUARD_BEGIN
    VCB_TRY(gv_args(v, epochs));
    VCB_TRY(traj_check(v->t, rows - 1, rows, offsets, nseq));
    if (!dfm || !dout) return fail(VCB_EARG, "null argument");
    VCB_ON_DEVICE(v->device);
    return traj_device(*v->t, dfm + 1, rows, offsets, nseq, chunk_limit, dout + 1, v->t->Ds + 1, nullptr, nullptr, true,
                       (cudaStream_t)stream, GvRun{v, epochs, alpha});
    VCB_GUARD_END
}

int32_t vcb_trajgv_convert_batch(const vcb_trajgv* v, const double* X, int32_t xrows, int64_t ldx,
                                 const int64_t* offsets, int64_t nseq, int32_t chunk_limit, int32_t epochs,
                                 double alpha, double* Y, int64_t ldy) {
    VCB_GUARD_BEGIN
    VCB_TRY(gv_args(v, epochs));
    VCB_TRY(traj_check(v->t, xrows, ldx, offsets, nseq));
    if (!X || !Y) return fail(VCB_EARG, "null argument");
    if (ldy < v->t->Ds) return fail(VCB_EARG, "ldy < dim/2");
    VCB_ON_DEVICE(v->device);
    return traj_host(*v->t, X, ldx, offsets, nseq, chunk_limit, Y, ldy, nullptr, nullptr, false, GvRun{v, epochs, alpha});
    VCB_GUARD_END
}

int32_t vcb_trajgv_vc_batch(const vcb_trajgv* v, const double* fm, int32_t rows, const int64_t* offsets,
                            int64_t nseq, int32_t chunk_limit, int32_t epochs, double alpha, double* out) {
    VCB_GUARD_BEGIN
    VCB_TRY(gv_args(v, epochs));
    VCB_TRY(traj_check(v->t, rows - 1, rows, offsets, nseq));
    if (!fm || !out) return fail(VCB_EARG, "null argument");
    VCB_ON_DEVICE(v->device);
    return traj_host(*v->t, fm + 1, rows, offsets, nseq, chunk_limit, out + 1, v->t->Ds + 1, nullptr, nullptr, true,
                     GvRun{v, epochs, alpha});
    VCB_GUARD_END
}

int32_t vcb_variance_scaling_batch_dev(const double* d_sigma2, int32_t D, const double* dX, int64_t ldx,
                                       const int64_t* offsets, int64_t nseq, double* dY, int64_t ldy, void* stream) {
    VCB_GUARD_BEGIN
    if (!d_sigma2 || !dX || !dY || !offsets) return fail(VCB_EARG, "null argument");
    if (D < 1 || nseq < 0 || ldx < D || ldy < D) return fail(VCB_EARG, "bad arguments");
    if (nseq == 0) return VCB_OK;
    VCB_TRY(ensure_device());
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc(st);
    int64_t* dOff = nullptr;
    VCB_CUDA(sc.get(&dOff, (size_t)nseq + 1));
    VCB_CUDA(cudaMemcpyAsync(dOff, offsets, ((size_t)nseq + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    return variance_scaling_device(d_sigma2, D, dX, ldx, dOff, nseq, dY, ldy, st);
    VCB_GUARD_END
}

int32_t vcb_variance_scaling_batch(const double* sigma2, int32_t D, const double* X, int64_t ldx,
                                   const int64_t* offsets, int64_t nseq, double* Y, int64_t ldy) {
    VCB_GUARD_BEGIN
    if (!sigma2 || !X || !Y || !offsets) return fail(VCB_EARG, "null argument");
    if (D < 1 || nseq < 0 || ldx < D || ldy < D) return fail(VCB_EARG, "bad arguments");
    if (nseq == 0) return VCB_OK;
    VCB_TRY(ensure_device());
    const int64_t base = offsets[0], total = offsets[nseq] - base;
    if (total <= 0) return VCB_OK;
    cudaStream_t st = nullptr;
    Scratch sc(st);
    double *dS = nullptr, *dI = nullptr, *dO = nullptr;
    int64_t* dOff = nullptr;
    VCB_CUDA(sc.get(&dS, (size_t)D));
    VCB_CUDA(sc.get(&dI, (size_t)total * D));
    VCB_CUDA(sc.get(&dO, (size_t)total * D));
    VCB_CUDA(sc.get(&dOff, (size_t)nseq + 1));
    std::vector<int64_t> rel(offsets, offsets + nseq + 1);
    for (auto& o : rel) o -= base;
    VCB_CUDA(cudaMemcpyAsync(dS, sigma2, (size_t)D * sizeof(double), cudaMemcpyHostToDevice, st));
    VCB_CUDA(cudaMemcpy2DAsync(dI, D * sizeof(double), X + base * ldx, ldx * sizeof(double), D * sizeof(double), total,
                               cudaMemcpyHostToDevice, st));
    VCB_CUDA(cudaMemcpyAsync(dOff, rel.data(), rel.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    VCB_TRY(variance_scaling_device(dS, D, dI, D, dOff, nseq, dO, D, st));
    VCB_CUDA(cudaMemcpy2DAsync(Y + base * ldy, ldy * sizeof(double), dO, D * sizeof(double), D * sizeof(double), total,
                               cudaMemcpyDeviceToHost, st));
    VCB_CUDA(cudaStreamSynchronize(st));
    return VCB_OK;
    VCB_GUARD_END
}

int32_t vcb_diffgmm(const double* mu, const double* sigma, int32_t twoD, int32_t M, double* mu_out, double* sigma_out) {
    VCB_GUARD_BEGIN
    if (!mu || !sigma || !mu_out || !sigma_out) return fail(VCB_EARG, "null argument");
    if (twoD < 2 || (twoD & 1) || M < 1) return fail(VCB_EARG, "bad joint dimension %d or M %d", twoD, M);
    diffgmm_params(mu, sigma, twoD, M, mu_out, sigma_out);
    return VCB_OK;
    VCB_GUARD_END
}

// ------------------------------------------------------------------------------------------------
// DTW
// ------------------------------------------------------------------------------------------------
static int32_t dtw_check(const int64_t* toff, const int64_t* soff, int64_t npairs, int D) {
    if (npairs < 0 || D < 1) return fail(VCB_EARG, "bad DTW batch (npairs=%lld, D=%d)", (long long)npairs, D);
    if (npairs > 0 && (!toff || !soff)) return fail(VCB_EARG, "null offsets");
    return VCB_OK;
}

int32_t vcb_dtw_fit_batch_dev(const double* dtmpl, const int64_t* tmpl_off, const double* dseq,
                              const int64_t* seq_off, int64_t npairs, int32_t D, int32_t fstep,
                              int32_t bstep, int64_t* dpaths, double* dfinal_cost, void* stream) {
    VCB_GUARD_BEGIN
    VCB_TRY(dtw_check(tmpl_off, seq_off, npairs, D));
    if (npairs == 0) return VCB_OK;
    if (!dtmpl || !dseq || !dpaths) return fail(VCB_EARG, "null argument");
    VCB_TRY(ensure_device());
    return dtw_fit_batch_device(dtmpl, tmpl_off, dseq, seq_off, npairs, D, fstep, bstep, dpaths, dfinal_cost,
                                (cudaStream_t)stream);
    VCB_GUARD_END
}

// One device: pairs [0, npairs) described by tmpl_off / seq_off (absolute frame offsets into tmpl / seq).
// Pairs are processed in slices on rotating streams, so the H2D copy of slice i+1 overlaps the kernel of slice i.
static int32_t dtw_host_one(const double* tmpl, const int64_t* tmpl_off, const double* seq, const int64_t* seq_off,
                            int64_t npairs, int D, int fstep, int bstep, int64_t* paths, double* final_cost) {
    static const int64_t slice_pairs = [] { const char* e = getenv("VCB_DTW_SLICE"); return e ? atoll(e) : 148LL; }();      // one pair per SM of the persistent DTW kernel; 74 .. 148 measured equal (copy-bound), 296: +10 %
    int dev = 0;
    VCB_TRY(ensure_device(&dev));
    HostPipe* hp = pipe_acquire(dev, 0, 0);
    if (!hp) return fail(VCB_ECUDA, "stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    // results go through the pipe's page-locked bounce buffer unless the caller's arrays are page-locked
    const int64_t sb0 = seq_off[0], nT_all = seq_off[npairs] - sb0;
    const bool direct = is_pinned(paths + sb0) && (!final_cost || is_pinned(final_cost));
    int64_t* bpaths = nullptr;
    double* bcost = nullptr;
    if (!direct) {
        void* b = hp->bounce_for((size_t)nT_all * sizeof(int64_t) + (size_t)npairs * sizeof(double));
        if (b) {
            bpaths = static_cast<int64_t*>(b);
            bcost = reinterpret_cast<double*>(bpaths + nT_all);
        }
    }
    int32_t rc = VCB_OK;
    int idx = 0;
    for (int64_t p0 = 0; p0 < npairs && rc == VCB_OK; idx = (idx + 1) % kSlots) {
        int64_t p1 = std::min(npairs, p0 + std::max<int64_t>(slice_pairs, 1));
        if (npairs - p1 < slice_pairs / 2) p1 = npairs;          // no tiny tail slice
        cudaStream_t st = hp->st[idx];
        rc = [&]() -> int32_t {
            const int64_t n = p1 - p0;
            const int64_t tb = tmpl_off[p0], sb = seq_off[p0];
            const int64_t nS = tmpl_off[p1] - tb, nT = seq_off[p1] - sb;
            Scratch sc(st);
            double *dT = nullptr, *dS = nullptr, *dC = nullptr;
            int64_t* dP = nullptr;
            VCB_CUDA(sc.get(&dT, (size_t)nS * D));
            VCB_CUDA(sc.get(&dS, (size_t)nT * D));
            VCB_CUDA(sc.get(&dP, (size_t)nT));
            VCB_CUDA(sc.get(&dC, (size_t)n));
            VCB_CUDA(cudaMemcpyAsync(dT, tmpl + tb * D, (size_t)nS * D * sizeof(double), cudaMemcpyHostToDevice, st));
            VCB_CUDA(cudaMemcpyAsync(dS, seq + sb * D, (size_t)nT * D * sizeof(double), cudaMemcpyHostToDevice, st));
            std::vector<int64_t> to(tmpl_off + p0, tmpl_off + p1 + 1), so(seq_off + p0, seq_off + p1 + 1);
            for (auto& o : to) o -= tb;
            for (auto& o : so) o -= sb;
            VCB_TRY(dtw_fit_batch_device(dT, to.data(), dS, so.data(), n, D, fstep, bstep, dP, dC, st));
            int64_t* hpaths = bpaths ? bpaths + (sb - sb0) : paths + sb;
            VCB_CUDA(cudaMemcpyAsync(hpaths, dP, (size_t)nT * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
            if (final_cost)
                VCB_CUDA(cudaMemcpyAsync(bcost ? bcost + p0 : final_cost + p0, dC, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
            return VCB_OK;
        }();
        p0 = p1;
    }
    for (int s2 = 0; s2 < kSlots; ++s2)
        if (cudaStreamSynchronize(hp->st[s2]) != cudaSuccess && rc == VCB_OK)
            rc = fail(VCB_ECUDA, "DTW failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc == VCB_OK && bpaths) {
        std::memcpy(paths + sb0, bpaths, (size_t)nT_all * sizeof(int64_t));
        if (final_cost) std::memcpy(final_cost, bcost, (size_t)npairs * sizeof(double));
    }
    pipe_release(hp);
    return rc;
}

int32_t vcb_dtw_fit_batch(const double* tmpl, const int64_t* tmpl_off, const double* seq,
                          const int64_t* seq_off, int64_t npairs, int32_t D, int32_t fstep, int32_t bstep,
                          int64_t* paths, double* final_cost) {
    VCB_GUARD_BEGIN
    VCB_TRY(dtw_check(tmpl_off, seq_off, npairs, D));
    if (npairs == 0) return VCB_OK;
    if (!tmpl || !seq || !paths) return fail(VCB_EARG, "null argument");
    for (int64_t p = 0; p < npairs; ++p)
        if (tmpl_off[p + 1] < tmpl_off[p] || seq_off[p + 1] < seq_off[p]) return fail(VCB_EARG, "offsets must be non-decreasing");
    const std::vector<int> devs = device_set();
    const int n = (int)devs.size();
    if (n <= 1 || npairs < 4 * n) return dtw_host_one(tmpl, tmpl_off, seq, seq_off, npairs, D, fstep, bstep, paths, final_cost);
    // pairs shard contiguously over the device set, balanced by cells S*T (one DTW object per pair, src/align.jl:16)
    const std::vector<int64_t> cut = split_by_cost(npairs, n, [&](int64_t i) {
        return (double)(tmpl_off[i + 1] - tmpl_off[i]) * (double)(seq_off[i + 1] - seq_off[i]);
    });
    return run_on_devices(devs, [&](int part, int) -> int32_t {
        const int64_t b = cut[part], e = cut[part + 1];
        if (e <= b) return VCB_OK;
        return dtw_host_one(tmpl, tmpl_off + b, seq, seq_off + b, e - b, D, fstep, bstep, paths, final_cost ? final_cost + b : nullptr);
    });
    VCB_GUARD_END
}

int32_t vcb_dtw_update(const double* tmpl, int32_t D, int32_t S, const double* lastcost, const double* v,
                       int32_t fstep, int32_t bstep, double* newcost, int64_t* newbp) {
    VCB_GUARD_BEGIN
    if (!tmpl || !lastcost || !v || !newcost || !newbp) return fail(VCB_EARG, "null argument");
    if (D < 1 || S < 1 || fstep < 0 || bstep < 0) return fail(VCB_EARG, "bad DTW arguments");
    VCB_TRY(ensure_device());
    cudaStream_t st = nullptr;
    Scratch sc(st);
    double *dT = nullptr, *dL = nullptr, *dV = nullptr, *dN = nullptr;
    int64_t* dB = nullptr;
    VCB_CUDA(sc.get(&dT, (size_t)S * D));
    VCB_CUDA(sc.get(&dL, (size_t)S));
    VCB_CUDA(sc.get(&dV, (size_t)D));
    VCB_CUDA(sc.get(&dN, (size_t)S));
    VCB_CUDA(sc.get(&dB, (size_t)S));
    VCB_CUDA(cudaMemcpyAsync(dT, tmpl, (size_t)S * D * sizeof(double), cudaMemcpyHostToDevice, st));
    VCB_CUDA(cudaMemcpyAsync(dL, lastcost, (size_t)S * sizeof(double), cudaMemcpyHostToDevice, st));
    VCB_CUDA(cudaMemcpyAsync(dV, v, (size_t)D * sizeof(double), cudaMemcpyHostToDevice, st));
    VCB_TRY(dtw_update_device(dT, D, S, dL, dV, fstep, bstep, dN, dB, st));
    VCB_CUDA(cudaMemcpyAsync(newcost, dN, (size_t)S * sizeof(double), cudaMemcpyDeviceToHost, st));
    VCB_CUDA(cudaMemcpyAsync(newbp, dB, (size_t)S * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    VCB_CUDA(cudaStreamSynchronize(st));
    return VCB_OK;
    VCB_GUARD_END
}

// ------------------------------------------------------------------------------------------------
// Callers either side of the path
// ------------------------------------------------------------------------------------------------
int32_t vcb_push_delta_batch(const double* src, int32_t D, const int64_t* offsets, int64_t nseq, double* out) {
    VCB_GUARD_BEGIN
    if (!src || !offsets || !out) return fail(VCB_EARG, "null argument");
    if (D < 1 || nseq < 0) return fail(VCB_EARG, "bad arguments");
    if (nseq == 0) return VCB_OK;
    VCB_TRY(ensure_device());
    const int64_t base = offsets[0], total = offsets[nseq] - base;
    if (total <= 0) return VCB_OK;
    cudaStream_t st = nullptr;
    Scratch sc(st);
    double *dI = nullptr, *dO = nullptr;
    int64_t* dOff = nullptr;
    VCB_CUDA(sc.get(&dI, (size_t)total * D));
    VCB_CUDA(sc.get(&dO, (size_t)total * 2 * D));
    VCB_CUDA(sc.get(&dOff, (size_t)nseq + 1));
    std::vector<int64_t> rel(offsets, offsets + nseq + 1);
    for (auto& o : rel) o -= base;
    VCB_CUDA(cudaMemcpyAsync(dI, src + base * D, (size_t)total * D * sizeof(double), cudaMemcpyHostToDevice, st));
    VCB_CUDA(cudaMemcpyAsync(dOff, rel.data(), rel.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    VCB_TRY(push_delta_device(dI, D, dOff, nseq, total, dO, st));
    VCB_CUDA(cudaMemcpyAsync(out + base * 2 * D, dO, (size_t)total * 2 * D * sizeof(double), cudaMemcpyDeviceToHost, st));
    VCB_CUDA(cudaStreamSynchronize(st));
    return VCB_OK;
    VCB_GUARD_END
}

static int32_t align_host_one(const double* src, const int64_t* src_off, const double* tgt, const int64_t* tgt_off,
                              int64_t npairs, int D, double* newtgt, int64_t* paths) {
    VCB_TRY(ensure_device());
    const int64_t sb = src_off[0], tb = tgt_off[0];
    const int64_t nS = src_off[npairs] - sb, nT = tgt_off[npairs] - tb;
    cudaStream_t st = nullptr;
    Scratch sc(st);
    double *dS = nullptr, *dT = nullptr, *dN = nullptr;
    int64_t *dP = nullptr, *dOff = nullptr;
    VCB_CUDA(sc.get(&dS, (size_t)nS * D));
    VCB_CUDA(sc.get(&dT, (size_t)nT * D));
    VCB_CUDA(sc.get(&dN, (size_t)nS * D));
    VCB_CUDA(sc.get(&dP, (size_t)nT));
    VCB_CUDA(sc.get(&dOff, 2 * ((size_t)npairs + 1)));
    std::vector<int64_t> so(src_off, src_off + npairs + 1), to(tgt_off, tgt_off + npairs + 1);
    for (auto& o : so) o -= sb;
    for (auto& o : to) o -= tb;
    VCB_CUDA(cudaMemcpyAsync(dS, src + sb * D, (size_t)nS * D * sizeof(double), cudaMemcpyHostToDevice, st));
    VCB_CUDA(cudaMemcpyAsync(dT, tgt + tb * D, (size_t)nT * D * sizeof(double), cudaMemcpyHostToDevice, st));
    VCB_CUDA(cudaMemcpyAsync(dOff, so.data(), so.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    VCB_CUDA(cudaMemcpyAsync(dOff + npairs + 1, to.data(), to.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    // align: template = source, sequence = target, DTW(fstep=0, bstep=2)  (src/align.jl:16-17)
    VCB_TRY(dtw_fit_batch_device(dS, so.data(), dT, to.data(), npairs, D, 0, 2, dP, nullptr, st));
    VCB_TRY(align_post_device(dT, dOff, dOff + npairs + 1, dP, npairs, D, dN, st));
    VCB_CUDA(cudaMemcpyAsync(newtgt + sb * D, dN, (size_t)nS * D * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (paths) VCB_CUDA(cudaMemcpyAsync(paths + tb, dP, (size_t)nT * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    VCB_CUDA(cudaStreamSynchronize(st));
    return VCB_OK;
}

int32_t vcb_align_batch(const double* src, const int64_t* src_off, const double* tgt, const int64_t* tgt_off,
                        int64_t npairs, int32_t D, double* newtgt, int64_t* paths) {
    VCB_GUARD_BEGIN
    VCB_TRY(dtw_check(src_off, tgt_off, npairs, D));
    if (npairs == 0) return VCB_OK;
    if (!src || !tgt || !newtgt) return fail(VCB_EARG, "null argument");
    for (int64_t p = 0; p < npairs; ++p)
        if (src_off[p + 1] < src_off[p] || tgt_off[p + 1] < tgt_off[p]) return fail(VCB_EARG, "offsets must be non-decreasing");
    const std::vector<int> devs = device_set();
    const int n = (int)devs.size();
    if (n <= 1 || npairs < 4 * n) return align_host_one(src, src_off, tgt, tgt_off, npairs, D, newtgt, paths);
    // the batch loop of bin/align.jl:84-113: pairs shard over the device set, balanced by cells
    const std::vector<int64_t> cut = split_by_cost(npairs, n, [&](int64_t i) {
        return (double)(src_off[i + 1] - src_off[i]) * (double)(tgt_off[i + 1] - tgt_off[i]);
    });
    return run_on_devices(devs, [&](int part, int) -> int32_t {
        const int64_t b = cut[part], e = cut[part + 1];
        if (e <= b) return VCB_OK;
        return align_host_one(src, src_off + b, tgt, tgt_off + b, e - b, D, newtgt, paths);
    });
    VCB_GUARD_END
}

}  // extern "C"
