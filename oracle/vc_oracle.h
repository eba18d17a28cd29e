/*
 * vc_oracle.h -- CPU oracle for the VoiceConversion.jl spectral-conversion hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a literal, single-threaded Float64 restatement of the
 * reference's Julia code (r9y9/VoiceConversion.jl).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (voiceconversion.jl_b200/) never links, imports or calls anything in this directory.
 *
 * Parity status:
 *   DTW (src/dtw.jl)              -- PINNED by the reference's own known-answer tests
 *                                    (test/dtw.jl:7-19, 21-31) -> tests/golden/dtw_*.json
 *   constructW                    -- PINNED by test/trajectory_gmmmap.jl:1-34 (exact structure)
 *   GMMMap ctor / accessors       -- PINNED (shape only) by test/gmmmap.jl:1-15 on the real model
 *   GV (gv.jl, TrajectoryGVGMMMap), diffgmm -- PARITY UNPINNED (the reference has no tests for them:
 *                                    "TODO: tests", src/diffgmm.jl:8); checked through their defining
 *                                    properties (tests/test_oracle_gv.py)
 *   fvconvert / vc numerics       -- PARITY UNPINNED by the reference: its tests only assert
 *                                    isfinite (test/vc.jl:26,50,72).  Julia 0.5 is not runnable in
 *                                    this image; the oracle is instead cross-checked against
 *                                    independent NumPy/SciPy formulations (tests/test_oracle_*.py).
 *
 * All matrices are column-major (Julia layout): Matrix{Float64}(D,T) == T contiguous D-vectors.
 */
#ifndef VC_ORACLE_H
#define VC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { VCO_OK = 0, VCO_EDIM = 1, VCO_ENOTPD = 2, VCO_ESINGULAR = 3, VCO_EARG = 4, VCO_ENOMEM = 5 };

typedef struct vco_gmmmap vco_gmmmap;
typedef struct vco_traj vco_traj;
typedef struct vco_dtw vco_dtw;

/* ---- src/gmmmap.jl:57-96, src/gmm.jl:8-20 ---- */
int vco_gmmmap_create(const double* weights, const double* mu /*(2D,M)*/,
                      const double* sigma /*(2D,2D,M)*/, int twoD, int M, int swap,
                      vco_gmmmap** out);
void vco_gmmmap_destroy(vco_gmmmap* g);
int vco_gmmmap_dim(const vco_gmmmap* g);
int vco_gmmmap_ncomponents(const vco_gmmmap* g);
int vco_gmmmap_length(const vco_gmmmap* g);
/* read-only views of the precomputed parameters, for tests: which = 0 mux(D,M), 1 muy(D,M),
 * 2 A(D,D,M) = SyxSxx^-1, 3 Sxx, 4 Sxy, 5 Syx, 6 Syy, 7 weights(M) */
const double* vco_gmmmap_param(const vco_gmmmap* g, int which);

/* src/gmm.jl:24-30 (vector) -- posterior has M entries */
int vco_predict_proba(const vco_gmmmap* g, const double* x, double* posterior);
/* src/gmm.jl:44-47 -- returns 1-based index of first max */
int vco_predict(const vco_gmmmap* g, const double* x, int* mhat);
/* src/gmmmap.jl:101-118 */
int vco_fvconvert(vco_gmmmap* g, const double* x, int xlen, double* y);
/* src/common.jl:7-26 : fm is (rows, T); out is (rows, T) */
int vco_vc_fbf(vco_gmmmap* g, const double* fm, int rows, int64_t T, double* out);
/* same loop run on nthreads OpenMP threads (for the "fair CPU" baseline only) */
int vco_vc_fbf_mt(const vco_gmmmap* g, const double* fm, int rows, int64_t T, double* out,
                  int nthreads);

/* ---- src/trajectory_gmmmap.jl:1-110 ---- */
int vco_traj_create(vco_gmmmap* g /*borrowed*/, int T, vco_traj** out);
void vco_traj_destroy(vco_traj* t);
int vco_traj_length(const vco_traj* t);
int vco_traj_dim(const vco_traj* t);
const double* vco_traj_Dy(const vco_traj* t); /* (2Ds,2Ds,M) */
/* constructW (src/trajectory_gmmmap.jl:39-61) as COO triplets, 0-based; returns nnz (or -1).
 * Call with rows==NULL to query nnz. */
int64_t vco_constructW(int D, int T, int64_t* rows, int64_t* cols, double* vals);
/* fvconvert(tgmm, X (2Ds,T)) -> Y (Ds,T); optional side outputs mhat (T, 1-based), Ey (2Ds,T) */
int vco_traj_fvconvert(vco_traj* t, const double* X, int xrows, int T, double* Y, int* mhat,
                       double* Ey);
/* src/common.jl:31-63 : fm (1+2Ds, T) -> out (1+Ds, T) */
int vco_vc_traj(vco_traj* t, const double* fm, int rows, int64_t T, double* out);
/* batch of utterances (ragged; offsets in frames, nseq+1 entries), independent fresh converter
 * state of chunk limit `limit` per utterance, OpenMP over utterances */
int vco_vc_traj_batch_mt(vco_gmmmap* g, int limit, const double* fm, int rows,
                         const int64_t* offsets, int64_t nseq, double* out, int nthreads);

/* ---- src/gv.jl:6-21 : fvpostf!(VarianceScaling(s2), src (D,T)) in place ---- */
void vco_variance_scaling(const double* s2, double* src, int D, int64_t T);

/* ---- src/trajectory_gmmmap.jl:112-189 : TrajectoryGVGMMMap ---- */
typedef struct vco_trajgv vco_trajgv;
int vco_trajgv_create(vco_traj* t /*borrowed*/, const double* muv /*(Ds)*/, const double* svv /*(Ds,Ds)*/,
                      vco_trajgv** out);
void vco_trajgv_destroy(vco_trajgv* v);
/* fvconvert(tgv, X (2Ds,T); epochs, alpha) -> Y (Ds,T) */
int vco_trajgv_fvconvert(vco_trajgv* v, const double* X, int xrows, int T, int epochs, double alpha,
                         double* Y);
/* vc(tgv, fm (1+2Ds,T)) -> out (1+Ds,T), chunked by length(tgv.tgmm) */
int vco_vc_trajgv(vco_trajgv* v, const double* fm, int rows, int64_t T, int epochs, double alpha,
                  double* out);

/* ---- src/diffgmm.jl:9-25 on the joint parameters: mu (2D,M), sigma (2D,2D,M) ---- */
void vco_diffgmm(const double* mu, const double* sigma, int twoD, int M, double* mu_out, double* sigma_out);

/* ---- src/dtw.jl ---- */
int vco_dtw_create(int fstep, int bstep, vco_dtw** out);
void vco_dtw_destroy(vco_dtw* d);
/* fit!(d, template (D,S), sequence (D,T)) -> path (T, 1-based)  src/dtw.jl:93-128 */
int vco_dtw_fit(vco_dtw* d, const double* tmpl, int D, int S, const double* seq, int T,
                int64_t* path);
/* set_template! / update! / backward  src/dtw.jl:53-90,133-145 */
int vco_dtw_set_template(vco_dtw* d, const double* tmpl, int D, int S);
int vco_dtw_update(vco_dtw* d, const double* v, int vlen);
int vco_dtw_backward(const vco_dtw* d, int64_t* path /* ncols-1 entries */);
/* table access for tests: costtable (S, ncols), backpointer (S, ncols) */
int vco_dtw_tables(const vco_dtw* d, int* S, int* ncols, const double** cost,
                   const int64_t** backptr);
/* ragged batch, OpenMP over pairs (offsets in frames, npairs+1 entries) */
int vco_dtw_fit_batch_mt(const double* tmpl, const int64_t* tmpl_off, const double* seq,
                         const int64_t* seq_off, int64_t npairs, int D, int fstep, int bstep,
                         int64_t* paths, double* final_cost, int nthreads);

/* ---- "next" rows ---- */
/* push_delta  src/datasets.jl:6-13 : src (D,T) -> out (2D,T) */
void vco_push_delta(const double* src, int D, int T, double* out);
/* align  src/align.jl:8-35 : newtgt (D,S) */
int vco_align(const double* src, int D, int S, const double* tgt, int T, double* newtgt,
              int64_t* path);

int vco_max_threads(void);
#ifdef __cplusplus
}
#endif
#endif
