// vcb_common.h -- shared internals of libvcb200 (error reporting, CUDA checks, handle layouts).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <vector>

#include "vcb200.h"

namespace vcb {

// Thread-local last-error message (returned by vcb_last_error).
int32_t fail(int32_t code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern std::atomic<int> g_variant;  // 0 auto, 1 simt, 2 tcgen05

#define VCB_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return ::vcb::fail(VCB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                             \
    } while (0)

#define VCB_TRY(expr)                      \
    do {                                   \
        int32_t _rc = (expr);              \
        if (_rc != VCB_OK) return _rc;     \
    } while (0)

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Optional per-stage CUDA-event marks of the last trajectory / DTW device call (vcb_stage_timing):
// bench.py uses them to time the dominant kernel alone inside its timed region.  Process-wide,
// profiling only (not meant for concurrent callers).
void stage_begin(cudaStream_t st);
void stage_mark(cudaStream_t st);

// Device buffer with RAII (handles and per-call scratch).
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count) {
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t upload(const std::vector<T>& h) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
};

}  // namespace vcb

// ------------------------------------------------------------------------------------------------
// Handles
// ------------------------------------------------------------------------------------------------

// Packed operands of the tcgen05 (3xTF32) posterior + conditional-mean kernel, see vcb_fbf_tc.cu.
struct vcb_tc_pack {
    int KP = 0;             // reduction length (multiple of 8): [xc (D) | 0-pad | ones (2)], see vcb_model.cpp
    int c1 = 0;             // column of the first "ones" entry (the second is c1 + 1)
    int koff = 0;           // 1 if the ones columns live in an extra, single-pass k-step
    int GC = 0, GW = 0;     // mixtures per MMA chunk: conversion kernel / whitening-only kernel
    int NC = 0, NW = 0;     // MMA N = GC*2*DP / GW*DP
    int NCHC = 0, NCHW = 0; // number of chunks = ceil(M / G)
    vcb::DevBuf<float> Bc;  // [NCHC][2 (hi,lo)][image NC x KP]   whitening + regression rows
    vcb::DevBuf<float> Bw;  // [NCHW][2 (hi,lo)][image NW x KP]   whitening rows only
    // the same operands split by row halves for CTA-pair MMAs (cta_group::2): per chunk
    // [CTA rank 0: hi, lo images of rows 0..N/2) | [rank 1: hi, lo of rows N/2..N)]
    vcb::DevBuf<float> Bc2, Bw2;
    vcb::DevBuf<float> cst; // c_m = log w - (D log 2pi + logdet)/2, padded with -inf
};

struct vcb_gmmmap {
    int device = 0;
    int D = 0, M = 0;
    int DP = 0;   // D rounded up to a multiple of 8 (tensor-core kernel)
    int DS = 0;   // padded dimension of the CUDA-core kernel (one of its instantiated sizes; 0 = none)
    int KS = 0;   // CUDA-core operand row length: [1 | xc (DS) | 0-pad] rounded up to a multiple of 4
    // host Float64 parameters (GMMMapParam, src/gmmmap.jl:10-21)
    std::vector<double> w, mux, muy, A, Sxx, Sxy, Syx, Syy;
    std::vector<double> xbar;  // centring vector (weighted mean of mux), applied in Float64
    // device Float64 (exact-path operands: arg-max re-check, E_t, posterior output)
    vcb::DevBuf<double> d_linv;  // [M][D][D] row-major inverse Cholesky factor (lower)
    vcb::DevBuf<double> d_linv_cm;  // the same, column-major ([k][r]) for coalesced row-parallel reads
    vcb::DevBuf<double> d_mux, d_muy;  // [M][D]
    vcb::DevBuf<double> d_A;     // [M][D*D] column-major (A[i + k*D])
    vcb::DevBuf<double> d_c;     // [M]
    vcb::DevBuf<double> d_xbar;  // [D]
    // device fp32 operands of the CUDA-core kernel: [M][2*DS][KS]
    vcb::DevBuf<float> d_w32;
    vcb::DevBuf<float> d_c32;    // [M]
    vcb_tc_pack tc;
    // multi-device mode (vcb_init): the joint parameters this handle was built from, and its replicas on
    // the other devices (index = device ordinal; built on first use by that device's worker, owned here)
    std::vector<double> src_w, src_mu, src_sigma;
    int src_swap = 0;
    mutable std::mutex rep_mu;
    mutable std::vector<vcb_gmmmap*> replicas;
    vcb_gmmmap() = default;
    vcb_gmmmap(const vcb_gmmmap&) = delete;
    vcb_gmmmap& operator=(const vcb_gmmmap&) = delete;
    ~vcb_gmmmap() {
        for (vcb_gmmmap* r : replicas) delete r;
    }
};

struct vcb_traj;
struct vcb_trajgv {
    int device = 0;                   // own copy: the parent may be finalised first (unordered GC finalisers)
    const vcb_traj* t = nullptr;      // borrowed
    std::vector<double> muv, pv;      // GV mean (Ds), inv(S_vv) (Ds,Ds) column-major
    vcb::DevBuf<double> d_muv, d_pv;
};

struct vcb_traj {
    int device = 0;             // own copy of g->device (see vcb_trajgv)
    const vcb_gmmmap* g = nullptr;
    int Ds = 0;                 // static dimension = dim(g)/2
    std::vector<double> Dy;     // (2Ds,2Ds,M) as the reference computes it (LU inverse)
    vcb::DevBuf<double> d_P;    // [M][2Ds*2Ds] symmetrised precision, column-major
    // sticky device flag: a solver met a non-positive pivot of W' D^-1 W (some Dy_m is not positive
    // definite).  Host entry points read and clear it after their synchronisation; *_dev callers
    // query it with vcb_traj_status.
    vcb::DevBuf<int> d_err;
    mutable std::mutex rep_mu;                    // multi-device mode: replicas on the other devices
    mutable std::vector<vcb_traj*> replicas;
    vcb_traj() = default;
    vcb_traj(const vcb_traj&) = delete;
    vcb_traj& operator=(const vcb_traj&) = delete;
    ~vcb_traj() {
        for (vcb_traj* r : replicas) delete r;
    }
};
