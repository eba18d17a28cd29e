/*
 * vcb200.h -- C ABI of libvcb200.so: the B200 (sm_100a) implementation of the spectral-parameter
 * conversion hot path of r9y9/VoiceConversion.jl.
 *
 * The reference has no FFI layer (it is plain Julia); these entry points are what a Julia `ccall`
 * shim binds to keep the package's API (see INTEGRATION.md and julia/VoiceConversionB200.jl).
 * Each entry point names the reference interface it replaces (paths relative to the reference).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.  Every function returns an
 *     int32 status (VCB_OK == 0); no exception crosses the boundary.  vcb_last_error() returns
 *     the calling thread's last message.
 *   - All matrices are column-major Float64, exactly the memory of a Julia Matrix{Float64}:
 *     a (D, T) feature matrix is T contiguous frames of `ld` doubles (ld >= D).
 *   - Caller owns all buffers; the library never keeps a caller pointer after returning.
 *   - Handles are bound to the CUDA device that was current when they were created
 *     (vcb_set_device); they are immutable, so convert/fit calls are re-entrant.
 *   - `*_dev` variants take DEVICE pointers and a cudaStream_t (passed as void*); they enqueue
 *     work and return without synchronising.  The host variants stage through the device
 *     (pipelined H2D / kernel / D2H) and return when the result is in the caller's buffer.
 *   - There is no CPU fallback: without a usable sm_100 device every compute call fails with
 *     VCB_ECUDA.
 */
#ifndef VCB200_H
#define VCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VCB_VERSION 100 /* 0.1.0 */

enum vcb_status {
    VCB_OK = 0,
    VCB_EDIM = 1,      /* DimensionMismatch("Inconsistent dimentions.")  src/gmmmap.jl:102, src/trajectory_gmmmap.jl:68 */
    VCB_ENOTPD = 2,    /* PosDefException from MvNormal(mu, Hermitian(Sxx))   src/gmm.jl:16-17 */
    VCB_ESINGULAR = 3, /* SingularException from Sxx^-1 / Dy^-1               src/gmmmap.jl:35, src/trajectory_gmmmap.jl:27 */
    VCB_EARG = 4,      /* ArgumentError (null pointer, bad size, weights not a probability vector) */
    VCB_ENOMEM = 5,
    VCB_ECUDA = 6,     /* CUDA runtime failure or no sm_100 device */
    VCB_EUNSUPPORTED = 7
};

typedef struct vcb_gmmmap vcb_gmmmap; /* GMMMap            src/gmmmap.jl:57-91 */
typedef struct vcb_traj vcb_traj;     /* TrajectoryGMMMap  src/trajectory_gmmmap.jl:3-32 */

/* ---------------------------------------------------------------------------------------------
 * Library / device
 * ------------------------------------------------------------------------------------------- */
int32_t vcb_version(void);
/* Copies the calling thread's last error message (NUL-terminated) into buf. */
int32_t vcb_last_error(char* buf, size_t buflen);
int32_t vcb_device_count(int32_t* count);
/* Makes `device` current for this host thread (one process per GPU: call once with LOCAL_RANK). */
int32_t vcb_set_device(int32_t device);
/* Multi-device mode for ONE process driving several GPUs -- what a Julia caller of vc(mapper, fms)
 * (bin/vc.jl:82) or of the batch align loop (bin/align.jl:84-113) gets: after vcb_init(ndev) (0 = all
 * visible devices, 1 = back to single-device) the HOST-pointer batch entry points -- vcb_gmmmap_convert,
 * vcb_gmmmap_vc, vcb_traj_convert_batch, vcb_traj_vc_batch, vcb_dtw_fit_batch, vcb_align_batch -- shard
 * their batch over devices 0..ndev-1: contiguous ranges balanced by cost (frames / frames per utterance
 * / cells per pair), one host thread and one copy/compute pipeline per device, the model replicated on
 * each device on first use, no exchange between devices (every unit is independent: src/common.jl:17-19,
 * :44-57, src/align.jl:16).  Small batches, the GV variant and all *_dev entry points stay on the
 * handle's device.  Page-lock the caller's buffers (vcb_host_alloc / vcb_host_register) so that every
 * device copies at PCIe speed. */
int32_t vcb_init(int32_t ndev);
/* Number of devices the host batch entry points currently shard over (1 = single-device mode). */
int32_t vcb_num_devices(int32_t* n);
/* Pinned host memory for zero-staging H2D/D2H (optional; any host pointer is accepted). */
int32_t vcb_host_alloc(void** ptr, size_t bytes);
int32_t vcb_host_free(void* ptr);
/* Page-locks / unlocks a caller-owned range (e.g. a Julia Array) so host calls copy at PCIe speed. */
int32_t vcb_host_register(void* ptr, size_t bytes);
int32_t vcb_host_unregister(void* ptr);
/* Kernel selection for the posterior/conditional-mean stage: 0 = auto, 1 = CUDA-core fp32 kernel,
 * 2 = tcgen05 3xTF32 kernel.  Variant 1 also routes DTW through the per-column barrier kernel instead of
 * the persistent warp-pipeline kernel.  Process-wide; intended for tests and profiling. */
int32_t vcb_set_kernel_variant(int32_t variant);
/* Number of kernel launches issued by this library since process start (bench bookkeeping). */
int64_t vcb_launch_count(void);

/* Profiling aid for bench.py: when enabled, the trajectory and DTW device paths record CUDA events
 * between their stages on the caller's stream (trajectory: arg-max + re-check | bucketing | E, PE |
 * band solver; DTW: template transpose | fused kernel).  vcb_stage_times waits for the last mark of
 * the most recent call and returns the stage durations in milliseconds (count <= capacity),
 * averaged over the calls of that kind recorded since timing was enabled (the last 64 at most), so
 * a loop may enqueue its steps back to back and read the times once at the end.
 * Process-wide and not meant for concurrent callers. */
int32_t vcb_stage_timing(int32_t enable);
int32_t vcb_stage_times(double* ms, int32_t capacity, int32_t* count);

/* ---------------------------------------------------------------------------------------------
 * GMMMap -- replaces GMMMap(weights, mu, Sigma; swap) (src/gmmmap.jl:62-90), GMMMapParam
 * (src/gmmmap.jl:10-39), split_joint_gmm (:41-52) and GaussianMixtureModel (src/gmm.jl:8-20).
 *   weights (M), mu (twoD, M), sigma (twoD, twoD, M), column-major.
 * Model preprocessing (split/swap, A = Syx Sxx^-1 by LU inverse, Cholesky of Hermitian(Sxx,:U),
 * log-determinants, whitening operands) runs once on the host in Float64 and is uploaded.
 * ------------------------------------------------------------------------------------------- */
int32_t vcb_gmmmap_create(const double* weights, const double* mu, const double* sigma,
                          int32_t twoD, int32_t M, int32_t swap, vcb_gmmmap** out);
int32_t vcb_gmmmap_destroy(vcb_gmmmap* g);
/* dim(g), ncomponents(g)  src/gmmmap.jl:94-95 (length(g) == 1 is a constant of the shim) */
int32_t vcb_gmmmap_dim(const vcb_gmmmap* g, int32_t* dim);
int32_t vcb_gmmmap_ncomponents(const vcb_gmmmap* g, int32_t* M);
/* Fields of GMMMapParam for the shim's `params` (src/gmmmap.jl:10-21).  which: 0 mux (D,M),
 * 1 muy (D,M), 2 SyxSxx^-1 (D,D,M), 3 Sxx, 4 Sxy, 5 Syx, 6 Syy (D,D,M), 7 weights (M). */
int32_t vcb_gmmmap_get_param(const vcb_gmmmap* g, int32_t which, double* out);

/* fvconvert(g, x) for T frames at once (src/gmmmap.jl:101-118; T == 1 is the reference call):
 *   Y[:,t] = sum_m P(m | X[:,t]) * (muy_m + A_m (X[:,t] - mux_m)).
 * xrows must equal dim(g) (else VCB_EDIM).  X (xrows, T) with leading dimension ldx; Y likewise. */
int32_t vcb_gmmmap_convert(const vcb_gmmmap* g, const double* X, int32_t xrows, int64_t T,
                           int64_t ldx, double* Y, int64_t ldy);
int32_t vcb_gmmmap_convert_dev(const vcb_gmmmap* g, const double* dX, int32_t xrows, int64_t T,
                               int64_t ldx, double* dY, int64_t ldy, void* stream);
/* vc(c::FrameByFrameConverter, fm) (src/common.jl:7-26): fm and out are (rows, T), row 0 (power)
 * is copied through, rows 1.. are converted.  rows must equal 1 + dim(g). */
int32_t vcb_gmmmap_vc(const vcb_gmmmap* g, const double* fm, int32_t rows, int64_t T, double* out);
int32_t vcb_gmmmap_vc_dev(const vcb_gmmmap* g, const double* dfm, int32_t rows, int64_t T,
                          double* dout, void* stream);
/* predict_proba(g.px, X) (src/gmm.jl:24-41): post (M, T).  predict (src/gmm.jl:44-58): mhat (T),
 * 1-based index of the first maximum. */
int32_t vcb_gmmmap_predict_proba(const vcb_gmmmap* g, const double* X, int32_t xrows, int64_t T,
                                 int64_t ldx, double* post);
int32_t vcb_gmmmap_predict(const vcb_gmmmap* g, const double* X, int32_t xrows, int64_t T,
                           int64_t ldx, int64_t* mhat);

/* ---------------------------------------------------------------------------------------------
 * TrajectoryGMMMap -- replaces TrajectoryGMMMap(g, T) (src/trajectory_gmmmap.jl:11-31),
 * constructW (:39-61; W is never materialised), fvconvert(tgmm, X) (:65-110) and
 * vc(c::TrajectoryConverter, fm) (src/common.jl:31-63).
 * The mutable `length(t)` state (W is rebuilt when T changes, :70-72) lives in the shim.
 * ------------------------------------------------------------------------------------------- */
/* Precomputes Dy_m = (Syy_m - A_m Sxy_m)^-1 (LU inverse).  dim(g) must be even (static+delta). */
int32_t vcb_traj_create(const vcb_gmmmap* g, vcb_traj** out);
int32_t vcb_traj_destroy(vcb_traj* t);
/* The band solver is a Cholesky of W' D^-1 W (src/trajectory_gmmmap.jl:105, where the reference's
 * sparse `\` falls back to LU).  If some Dy_m is not positive definite a pivot fails: the host entry
 * points then return VCB_ENOTPD; after *_dev calls, vcb_traj_status(t, stream) synchronises the
 * stream and returns VCB_ENOTPD if any conversion enqueued on this handle since the last query met
 * such a pivot (the flag is cleared by the query), VCB_OK otherwise. */
int32_t vcb_traj_status(const vcb_traj* t, void* stream);
/* Dy (dim, dim, M) as the reference stores it (src/trajectory_gmmmap.jl:24-28). */
int32_t vcb_traj_get_Dy(const vcb_traj* t, double* out);
/* Batched fvconvert / vc.  X holds the frames of nseq utterances back to back: (xrows, total)
 * with leading dimension ldx; offsets (nseq+1, in frames).  Each utterance is cut into chunks of
 * `chunk_limit` frames (last chunk shorter; <= 0 means one chunk) and each chunk is solved
 * independently:  (W' D^-1 W) y = W' D^-1 E  with the arg-max mixture sequence.
 * Y: static output (xrows/2, total) with leading dimension ldy.
 * Optional side outputs (may be NULL): mhat (total, 1-based), Ey (xrows, total) -- the
 * tgmm.Ey the GV variant reuses (src/trajectory_gmmmap.jl:90-91). */
int32_t vcb_traj_convert_batch(const vcb_traj* t, const double* X, int32_t xrows, int64_t ldx,
                               const int64_t* offsets, int64_t nseq, int32_t chunk_limit,
                               double* Y, int64_t ldy, int64_t* mhat, double* Ey);
/* Device variant: dX, dY, dmhat, dEy are device pointers; offsets stays a HOST array. */
int32_t vcb_traj_convert_batch_dev(const vcb_traj* t, const double* dX, int32_t xrows,
                                   int64_t ldx, const int64_t* offsets, int64_t nseq,
                                   int32_t chunk_limit, double* dY, int64_t ldy, int64_t* dmhat,
                                   double* dEy, void* stream);
/* vc(c::TrajectoryConverter, fm) for a batch: fm (rows, total) with rows = 1 + dim; out
 * (1 + dim/2, total); row 0 passthrough (src/common.jl:35,60). */
int32_t vcb_traj_vc_batch(const vcb_traj* t, const double* fm, int32_t rows,
                          const int64_t* offsets, int64_t nseq, int32_t chunk_limit, double* out);
int32_t vcb_traj_vc_batch_dev(const vcb_traj* t, const double* dfm, int32_t rows,
                              const int64_t* offsets, int64_t nseq, int32_t chunk_limit,
                              double* dout, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DTW -- replaces DTWs.fit!(d, template, sequence) + backward(d) (src/dtw.jl:93-145) for a batch
 * of pairs, and update!(d, v) (src/dtw.jl:61-90) as a single-column step.
 *   tmpl: templates back to back (D, sum S_p); seq likewise; offsets in frames (npairs+1).
 *   paths: (sum T_p) 1-based template indices, pair p at paths[seq_off[p] ...].
 *   final_cost (npairs, may be NULL): costtable[path[end], T+1].
 * Bit-exact with the reference recurrence: Float64, ((cost + ocost) + transition), candidates in
 * the order i, i-bstep .. i+fstep with strict `<`, first minimum at the free end point.
 * ------------------------------------------------------------------------------------------- */
int32_t vcb_dtw_fit_batch(const double* tmpl, const int64_t* tmpl_off, const double* seq,
                          const int64_t* seq_off, int64_t npairs, int32_t D, int32_t fstep,
                          int32_t bstep, int64_t* paths, double* final_cost);
/* Device variant: dtmpl/dseq/dpaths/dfinal_cost are device pointers; offsets are HOST arrays. */
int32_t vcb_dtw_fit_batch_dev(const double* dtmpl, const int64_t* tmpl_off, const double* dseq,
                              const int64_t* seq_off, int64_t npairs, int32_t D, int32_t fstep,
                              int32_t bstep, int64_t* dpaths, double* dfinal_cost, void* stream);
/* One column of the recurrence (update!, src/dtw.jl:61-90): lastcost (S) -> newcost (S),
 * newbp (S, 1-based). Host pointers. */
int32_t vcb_dtw_update(const double* tmpl, int32_t D, int32_t S, const double* lastcost,
                       const double* v, int32_t fstep, int32_t bstep, double* newcost,
                       int64_t* newbp);

/* ---------------------------------------------------------------------------------------------
 * Callers either side of the path (SURVEY.md section 8f "next" rows)
 * ------------------------------------------------------------------------------------------- */
/* push_delta (src/datasets.jl:6-13): src (D, total) -> out (2D, total), per utterance. */
int32_t vcb_push_delta_batch(const double* src, int32_t D, const int64_t* offsets, int64_t nseq,
                             double* out);
/* align(src, tgt) (src/align.jl:8-35) for a batch: DTW(fstep=0, bstep=2), scatter, one-pass
 * hole interpolation.  newtgt (D, sum S_p); paths as in vcb_dtw_fit_batch (may be NULL). */
int32_t vcb_align_batch(const double* src, const int64_t* src_off, const double* tgt,
                        const int64_t* tgt_off, int64_t npairs, int32_t D, double* newtgt,
                        int64_t* paths);

/* vc(mapper, [fm[1,:]; push_delta(fm[2:end,:])]) fused (bin/vc.jl:76-82): fm holds the power row and
 * the STATIC features only, (1+Ds, total); the delta rows are appended per utterance on the device
 * (boundary quirk of src/datasets.jl:8-11 included) before the chunked conversion. out (1+Ds, total). */
int32_t vcb_traj_vc_static_batch(const vcb_traj* t, const double* fm, int32_t rows, const int64_t* offsets,
                                 int64_t nseq, int32_t chunk_limit, double* out);
int32_t vcb_traj_vc_static_batch_dev(const vcb_traj* t, const double* dfm, int32_t rows, const int64_t* offsets,
                                     int64_t nseq, int32_t chunk_limit, double* dout, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Global variance and differential models
 * ------------------------------------------------------------------------------------------- */
typedef struct vcb_trajgv vcb_trajgv; /* TrajectoryGVGMMMap  src/trajectory_gmmmap.jl:114-127 */
/* TrajectoryGVGMMMap(tgmm, mu_v (Ds), S_vv (Ds,Ds)): asserts mu_v >= 0 (:124, VCB_EARG), stores
 * inv(S_vv) (:125, VCB_ESINGULAR).  The trajectory handle is borrowed and must outlive this one. */
int32_t vcb_trajgv_create(const vcb_traj* t, const double* mu_v, const double* sigma_vv, vcb_trajgv** out);
int32_t vcb_trajgv_destroy(vcb_trajgv* v);
/* fvconvert(tgv, X; epochs, alpha) (src/trajectory_gmmmap.jl:140-172) for a ragged batch, chunked
 * like vcb_traj_convert_batch.  The reference's defaults are epochs = 100, alpha = 1.0e-5.  Chunks
 * shorter than 2 frames are rejected (their global variance is NaN and the reference asserts). */
int32_t vcb_trajgv_convert_batch(const vcb_trajgv* v, const double* X, int32_t xrows, int64_t ldx,
                                 const int64_t* offsets, int64_t nseq, int32_t chunk_limit, int32_t epochs,
                                 double alpha, double* Y, int64_t ldy);
int32_t vcb_trajgv_convert_batch_dev(const vcb_trajgv* v, const double* dX, int32_t xrows, int64_t ldx,
                                     const int64_t* offsets, int64_t nseq, int32_t chunk_limit, int32_t epochs,
                                     double alpha, double* dY, int64_t ldy, void* stream);
/* vc(c::TrajectoryConverter, fm) with the GV converter (src/common.jl:31-63): fm (1+2Ds, total) ->
 * out (1+Ds, total), power row copied. */
int32_t vcb_trajgv_vc_batch(const vcb_trajgv* v, const double* fm, int32_t rows, const int64_t* offsets,
                            int64_t nseq, int32_t chunk_limit, int32_t epochs, double alpha, double* out);
int32_t vcb_trajgv_vc_batch_dev(const vcb_trajgv* v, const double* dfm, int32_t rows, const int64_t* offsets,
                                int64_t nseq, int32_t chunk_limit, int32_t epochs, double alpha, double* dout,
                                void* stream);
/* fvpostf(VarianceScaling(sigma2), src) (src/gv.jl:10-21) per utterance of a ragged batch:
 * Y = sqrt(sigma2 ./ var(X, 2)) .* (X .- mean(X, 2)) .+ mean(X, 2); X, Y (D, total) with leading
 * dimensions ldx, ldy; Y may alias X (fvpostf!). */
int32_t vcb_variance_scaling_batch(const double* sigma2, int32_t D, const double* X, int64_t ldx,
                                   const int64_t* offsets, int64_t nseq, double* Y, int64_t ldy);
int32_t vcb_variance_scaling_batch_dev(const double* d_sigma2, int32_t D, const double* dX, int64_t ldx,
                                       const int64_t* offsets, int64_t nseq, double* dY, int64_t ldy, void* stream);
/* diffgmm (src/diffgmm.jl:9-25) on the joint parameters: mu (2D,M), sigma (2D,2D,M) -> the joint
 * parameters of the differential model; feed them to vcb_gmmmap_create. Host-side, no GPU work. */
int32_t vcb_diffgmm(const double* mu, const double* sigma, int32_t twoD, int32_t M, double* mu_out,
                    double* sigma_out);

#ifdef __cplusplus
}
#endif
#endif /* VCB200_H */
