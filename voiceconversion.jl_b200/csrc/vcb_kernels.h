// vcb_kernels.h -- device-level entry points (all take device pointers + a stream, never sync).
#pragma once
#include <algorithm>

#include "vcb_common.h"
#include "vcb_model.h"

namespace vcb {

// ---- K4 DTW (vcb_dtw.cu)
int32_t dtw_fit_batch_device(const double* d_tmpl, const int64_t* h_toff, const double* d_seq,
                             const int64_t* h_soff, int64_t npairs, int D, int fstep, int bstep,
                             int64_t* d_paths, double* d_final_cost, cudaStream_t st);
int32_t dtw_update_device(const double* d_tmpl, int D, int S, const double* d_last,
                          const double* d_v, int fstep, int bstep, double* d_newcost,
                          int64_t* d_newbp, cudaStream_t st);

// ---- K1/K2 on CUDA cores, fp32 (vcb_gmm_simt.cu)
// Supported padded dimensions of the CUDA-core kernel; returns 0 if D is too large.
int simt_padded_dim(int D);
// Y[:,t] = sum_m P(m|x_t)(muy_m + A_m (x_t - mux_m)); optional copy of the row above X (power).
int32_t simt_convert(const vcb_gmmmap& g, const double* dX, int64_t T, int64_t ldx, double* dY,
                     int64_t ldy, bool copy_power, cudaStream_t st);
// arg-max mixture per frame (0-based, int32) with an FP64 re-check of near ties.
int32_t simt_argmax(const vcb_gmmmap& g, const double* dX, int64_t T, int64_t ldx, int32_t* d_mhat,
                    cudaStream_t st);
// FP64 re-check of flagged frames (shared with the tensor-core arg-max kernel).
//   d_flag_count: one int (device), d_flag_list: frame indices.
int32_t recheck_argmax_fp64(const vcb_gmmmap& g, const double* dX, int64_t ldx, const int* d_flag_count,
                            const int64_t* d_flag_list, int32_t* d_mhat, cudaStream_t st);
// predict_proba in Float64 (src/gmm.jl:24-41): post (M, T)
int32_t proba_fp64(const vcb_gmmmap& g, const double* dX, int64_t T, int64_t ldx, double* d_post,
                   cudaStream_t st);
int32_t widen_mhat(const int32_t* d_mhat, int64_t T, int64_t* d_out, cudaStream_t st);

// ---- K1/K2 on tcgen05 tensor cores, 3xTF32 (vcb_gmm_tc.cu)
bool tc_supported(const vcb_gmmmap& g, bool convert);
int32_t tc_convert(const vcb_gmmmap& g, const double* dX, int64_t T, int64_t ldx, double* dY,
                   int64_t ldy, bool copy_power, cudaStream_t st);
int32_t tc_argmax(const vcb_gmmmap& g, const double* dX, int64_t T, int64_t ldx, int32_t* d_mhat,
                  cudaStream_t st);

// ---- K3 trajectory (vcb_traj.cu)
// Chunks: c_begin[nchunks+1] frame offsets (device), each chunk solved independently.
int32_t traj_solve_device(const vcb_traj& t, const double* dX, int64_t ldx, const int32_t* d_mhat,
                          const int64_t* d_chunk_off, int64_t nchunks, int max_chunk_len,
                          int64_t total_frames, double* dY, int64_t ldy, double* dEy_out,
                          bool copy_power, cudaStream_t st, const int* ws = nullptr, int64_t npanels = 0);

// ---- per-mixture grouped Float64 products (vcb_group.cu)
size_t group_workspace_ints(int M, int64_t total);
int32_t group_frames_by_mixture(const int32_t* d_mhat, int64_t total, int M, int* ws, int64_t* npanels_bound,
                                cudaStream_t st);
// E_t = muy_m + A_m (x_t - mux_m), g_t = P_m E_t for all frames (d_mhat 0-based buckets in ws)
int32_t group_e_step(const vcb_traj& tr, const int* ws, int64_t npanels, const double* dX, int64_t ldx, double* dE,
                     double* dEout, double* dG, cudaStream_t st);
// h_t = P_m (E_t - (W y)_t)
int32_t group_gv_step(const vcb_traj& tr, const int* ws, int64_t npanels, const double* dY, int64_t ldy,
                      const double* dE, const unsigned char* d_edge, double* dH, cudaStream_t st);

// exact arg-max of the first `cap` flagged frames as (panel, mixture) tasks; d_lout: cap * M doubles
int32_t recheck_argmax_panels(const vcb_gmmmap& g, const double* dX, int64_t ldx, const int* d_flag_count,
                              const int64_t* d_flag_list, int32_t* d_mhat, int cap, double* d_lout, cudaStream_t st);

// ---- GV helpers (vcb_gv.cu)
int32_t variance_scaling_device(const double* d_s2, int D, const double* dX, int64_t ldx, const int64_t* d_off,
                                int64_t nseq, double* dY, int64_t ldy, cudaStream_t st);
// gradient ascent of fvconvert(tgv, X) on the solved trajectories dY (in place); d_mhat 1-based
int32_t trajgv_ascent_device(const vcb_trajgv& v, double* dY, int64_t ldy, const double* dE, const int64_t* d_mhat,
                             const int64_t* d_chunk_off, int64_t nchunks, int64_t total, int epochs, double alpha,
                             cudaStream_t st, const int* ws = nullptr, int64_t npanels = 0);

// ---- callers either side (vcb_aux.cu)
int32_t push_delta_device(const double* d_src, int D, const int64_t* d_off, int64_t nseq,
                          int64_t total, double* d_out, cudaStream_t st);
// same with leading dimensions: src rows of length lds (D used), out rows of length ldo (2D written)
int32_t push_delta_strided_device(const double* d_src, int64_t lds, int D, const int64_t* d_off, int64_t nseq,
                                  int64_t total, double* d_out, int64_t ldo, cudaStream_t st);
int32_t align_post_device(const double* d_tgt, const int64_t* d_soff_src, const int64_t* d_soff_tgt,
                          const int64_t* d_paths, int64_t npairs, int D, double* d_newtgt,
                          cudaStream_t st);

}  // namespace vcb
