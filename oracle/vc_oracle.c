/*
 * vc_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see vc_oracle.h for the rules).
 *
 * Literal Float64 restatement of r9y9/VoiceConversion.jl's conversion hot path.  Every function
 * cites the reference file:line it follows (paths relative to the reference checkout).
 * Third-party Julia arithmetic that is not vendored in the reference (Distributions.MvNormal /
 * PDMats Cholesky log-pdf, StatsFuns.logsumexp, Base dense `^-1` = LU inverse, Base sparse `\`)
 * is restated from its published definition; REQUIRE:1-8 only gives lower bounds
 * (Distributions >= 0.6.2, StatsFuns/StatsBase unbounded), there is no lock file.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC  (see oracle/Makefile).
 * -ffp-contract=off matters: DTW parity is bit-exact and the reference adds/multiplies without FMA.
 */
#include "vc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* small dense helpers (column-major)                                                          */
/* ------------------------------------------------------------------------------------------ */

/* Dense inverse through LU with partial pivoting: what Julia's `A^-1` / `inv(A)` does for a
 * general Matrix{Float64} (getrf + getri).  a is n x n column-major, overwritten with inv(a). */
static int lu_inverse(double* a, int n) {
    int* piv = (int*)malloc(sizeof(int) * (size_t)n);
    double* lu = (double*)malloc(sizeof(double) * (size_t)n * n);
    if (!piv || !lu) { free(piv); free(lu); return VCO_ENOMEM; }
    memcpy(lu, a, sizeof(double) * (size_t)n * n);
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(lu[k + (size_t)k * n]);
        for (int i = k + 1; i < n; ++i) {
            double v = fabs(lu[i + (size_t)k * n]);
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (best == 0.0) { free(piv); free(lu); return VCO_ESINGULAR; }
        if (p != k)
            for (int j = 0; j < n; ++j) {
                double t = lu[k + (size_t)j * n];
                lu[k + (size_t)j * n] = lu[p + (size_t)j * n];
                lu[p + (size_t)j * n] = t;
            }
        double d = lu[k + (size_t)k * n];
        for (int i = k + 1; i < n; ++i) lu[i + (size_t)k * n] /= d;
        for (int j = k + 1; j < n; ++j) {
            double t = lu[k + (size_t)j * n];
            if (t != 0.0)
                for (int i = k + 1; i < n; ++i) lu[i + (size_t)j * n] -= lu[i + (size_t)k * n] * t;
        }
    }
    /* solve A X = I column by column */
    for (int c = 0; c < n; ++c) {
        double* x = a + (size_t)c * n;
        for (int i = 0; i < n; ++i) x[i] = (i == c) ? 1.0 : 0.0;
        for (int k = 0; k < n; ++k) {
            int p = piv[k];
            if (p != k) { double t = x[k]; x[k] = x[p]; x[p] = t; }
        }
        for (int k = 0; k < n; ++k) {
            double t = x[k];
            if (t != 0.0)
                for (int i = k + 1; i < n; ++i) x[i] -= lu[i + (size_t)k * n] * t;
        }
        for (int k = n - 1; k >= 0; --k) {
            x[k] /= lu[k + (size_t)k * n];
            double t = x[k];
            for (int i = 0; i < k; ++i) x[i] -= lu[i + (size_t)k * n] * t;
        }
    }
    free(piv);
    free(lu);
    return VCO_OK;
}

/* Lower Cholesky factor of Hermitian(S, :U) -- src/gmm.jl:16 mirrors the UPPER triangle, then
 * PDMats takes the Cholesky factor.  l (n x n col-major) receives L with S = L L'. */
static int chol_lower_from_upper(const double* s, int n, double* l) {
    memset(l, 0, sizeof(double) * (size_t)n * n);
    for (int j = 0; j < n; ++j) {
        for (int i = j; i < n; ++i) {
            /* Hermitian(:U): element (i,j) with i >= j is read from (j,i) */
            double v = s[j + (size_t)i * n];
            for (int k = 0; k < j; ++k) v -= l[i + (size_t)k * n] * l[j + (size_t)k * n];
            if (i == j) {
                if (!(v > 0.0)) return VCO_ENOTPD;
                l[j + (size_t)j * n] = sqrt(v);
            } else {
                l[i + (size_t)j * n] = v / l[j + (size_t)j * n];
            }
        }
    }
    return VCO_OK;
}

/* c (m x n) = a (m x k) * b (k x n), column-major, plain triple loop */
static void matmul(const double* a, const double* b, double* c, int m, int k, int n) {
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) {
            double s = 0.0;
            for (int p = 0; p < k; ++p) s += a[i + (size_t)p * m] * b[p + (size_t)j * k];
            c[i + (size_t)j * m] = s;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* GMMMap -- src/gmmmap.jl, src/gmm.jl                                                         */
/* ------------------------------------------------------------------------------------------ */

struct vco_gmmmap {
    int D, M;
    double *w, *mux, *muy;          /* (M), (D,M), (D,M) */
    double *Sxx, *Sxy, *Syx, *Syy;  /* (D,D,M) each */
    double* A;                      /* SyxSxx^-1 (D,D,M)   src/gmmmap.jl:34-36 */
    double* Ey;                     /* scratch (D,M)       src/gmmmap.jl:84   */
    /* px = GaussianMixtureModel(mux, Sxx, w)              src/gmmmap.jl:87   */
    double* L;                      /* Cholesky factors (D,D,M) */
    double* logdet;                 /* (M) */
    double* lpr;                    /* scratch (M) */
};

void vco_gmmmap_destroy(vco_gmmmap* g) {
    if (!g) return;
    free(g->w); free(g->mux); free(g->muy);
    free(g->Sxx); free(g->Sxy); free(g->Syx); free(g->Syy);
    free(g->A); free(g->Ey); free(g->L); free(g->logdet); free(g->lpr);
    free(g);
}

int vco_gmmmap_create(const double* weights, const double* mu, const double* sigma, int twoD,
                      int M, int swap, vco_gmmmap** out) {
    if (!weights || !mu || !sigma || !out || twoD < 2 || M < 1) return VCO_EARG;
    /* src/gmmmap.jl:43  D = size(mu,1)>>1 ; an odd joint dimension makes GMMMapParam's
     * D x D assignment fail (:35) -> DimensionMismatch */
    if (twoD & 1) return VCO_EDIM;
    int D = twoD >> 1;
    size_t DD = (size_t)D * D;
    vco_gmmmap* g = (vco_gmmmap*)calloc(1, sizeof(*g));
    if (!g) return VCO_ENOMEM;
    g->D = D; g->M = M;
    g->w = (double*)malloc(sizeof(double) * M);
    g->mux = (double*)malloc(sizeof(double) * D * M);
    g->muy = (double*)malloc(sizeof(double) * D * M);
    g->Sxx = (double*)malloc(sizeof(double) * DD * M);
    g->Sxy = (double*)malloc(sizeof(double) * DD * M);
    g->Syx = (double*)malloc(sizeof(double) * DD * M);
    g->Syy = (double*)malloc(sizeof(double) * DD * M);
    g->A = (double*)malloc(sizeof(double) * DD * M);
    g->Ey = (double*)calloc((size_t)D * M, sizeof(double));
    g->L = (double*)malloc(sizeof(double) * DD * M);
    g->logdet = (double*)malloc(sizeof(double) * M);
    g->lpr = (double*)malloc(sizeof(double) * M);
    if (!g->w || !g->mux || !g->muy || !g->Sxx || !g->Sxy || !g->Syx || !g->Syy || !g->A ||
        !g->Ey || !g->L || !g->logdet || !g->lpr) { vco_gmmmap_destroy(g); return VCO_ENOMEM; }

    /* MixtureModel(normals, weights) -> Categorical(weights) requires a probability vector
     * (ext: Distributions.isprobvec: all >= 0 and sum ~= 1 with rtol sqrt(eps)) */
    double wsum = 0.0;
    for (int m = 0; m < M; ++m) {
        if (!(weights[m] >= 0.0)) { vco_gmmmap_destroy(g); return VCO_EARG; }
        wsum += weights[m];
    }
    if (!(fabs(wsum - 1.0) <= 1.4901161193847656e-08 * fmax(fabs(wsum), 1.0))) {
        vco_gmmmap_destroy(g); return VCO_EARG;
    }
    memcpy(g->w, weights, sizeof(double) * M);

    /* split_joint_gmm  src/gmmmap.jl:41-52, then optional swap :74-78 */
    for (int m = 0; m < M; ++m) {
        const double* mum = mu + (size_t)m * twoD;
        const double* sm = sigma + (size_t)m * twoD * twoD;
        double* mx = swap ? g->muy : g->mux;
        double* my = swap ? g->mux : g->muy;
        for (int i = 0; i < D; ++i) {
            mx[i + (size_t)m * D] = mum[i];
            my[i + (size_t)m * D] = mum[D + i];
        }
        double* xx = (swap ? g->Syy : g->Sxx) + m * DD;
        double* yy = (swap ? g->Sxx : g->Syy) + m * DD;
        double* xy = (swap ? g->Syx : g->Sxy) + m * DD;
        double* yx = (swap ? g->Sxy : g->Syx) + m * DD;
        for (int j = 0; j < D; ++j)
            for (int i = 0; i < D; ++i) {
                xx[i + (size_t)j * D] = sm[i + (size_t)j * twoD];
                xy[i + (size_t)j * D] = sm[i + (size_t)(D + j) * twoD];
                yx[i + (size_t)j * D] = sm[(D + i) + (size_t)j * twoD];
                yy[i + (size_t)j * D] = sm[(D + i) + (size_t)(D + j) * twoD];
            }
    }
    /* GMMMapParam  src/gmmmap.jl:34-36 : A[:,:,m] = Syx[:,:,m] * Sxx[:,:,m]^-1 */
    double* inv = (double*)malloc(sizeof(double) * DD);
    if (!inv) { vco_gmmmap_destroy(g); return VCO_ENOMEM; }
    for (int m = 0; m < M; ++m) {
        memcpy(inv, g->Sxx + m * DD, sizeof(double) * DD);
        int rc = lu_inverse(inv, D);
        if (rc != VCO_OK) { free(inv); vco_gmmmap_destroy(g); return rc; }
        matmul(g->Syx + m * DD, inv, g->A + m * DD, D, D, D);
    }
    free(inv);
    /* GaussianMixtureModel  src/gmm.jl:8-20 : MvNormal(mux[:,m], Array(Hermitian(Sxx[:,:,m]))) */
    for (int m = 0; m < M; ++m) {
        int rc = chol_lower_from_upper(g->Sxx + m * DD, D, g->L + m * DD);
        if (rc != VCO_OK) { vco_gmmmap_destroy(g); return rc; }
        double ld = 0.0;
        for (int k = 0; k < D; ++k) ld += log(g->L[m * DD + k + (size_t)k * D]);
        g->logdet[m] = 2.0 * ld;
    }
    *out = g;
    return VCO_OK;
}

int vco_gmmmap_dim(const vco_gmmmap* g) { return g->D; }          /* src/gmmmap.jl:94 */
int vco_gmmmap_ncomponents(const vco_gmmmap* g) { return g->M; }   /* src/gmmmap.jl:95 */
int vco_gmmmap_length(const vco_gmmmap* g) { (void)g; return 1; }  /* src/gmmmap.jl:93 */

const double* vco_gmmmap_param(const vco_gmmmap* g, int which) {
    switch (which) {
        case 0: return g->mux; case 1: return g->muy; case 2: return g->A;
        case 3: return g->Sxx; case 4: return g->Sxy; case 5: return g->Syx;
        case 6: return g->Syy; case 7: return g->w;
        default: return 0;
    }
}

#define LOG2PI 1.8378770664093454835606594728112

/* (ext) Distributions: logpdf(MvNormal) = -(D*log2pi + logdetcov)/2 - sqmahal/2, with
 * sqmahal = || L \ (x - mu) ||^2 through the stored Cholesky factor (PDMats invquad/whiten). */
static double mvn_logpdf(const vco_gmmmap* g, int m, const double* x, double* z) {
    int D = g->D;
    const double* L = g->L + (size_t)m * D * D;
    const double* mu = g->mux + (size_t)m * D;
    double q = 0.0;
    for (int i = 0; i < D; ++i) {
        double v = x[i] - mu[i];
        for (int k = 0; k < i; ++k) v -= L[i + (size_t)k * D] * z[k];
        z[i] = v / L[i + (size_t)i * D];
        q += z[i] * z[i];
    }
    return -(D * LOG2PI + g->logdet[m]) / 2.0 - q / 2.0;
}

/* src/gmm.jl:24-30.  Returns the number of entries written (components with p > 0), the
 * `find(p .> 0.)` filter is kept literally. */
static int predict_proba_core(const vco_gmmmap* g, const double* x, double* post, double* z) {
    int n = 0;
    for (int m = 0; m < g->M; ++m) {
        if (g->w[m] > 0.0) post[n++] = mvn_logpdf(g, m, x, z) + log(g->w[m]);
    }
    /* (ext) StatsFuns.logsumexp: u = maximum; log(sum(exp(x - u))) + u */
    double u = -INFINITY;
    for (int i = 0; i < n; ++i) if (post[i] > u) u = post[i];
    double logprob;
    if (n == 0) logprob = -INFINITY;
    else if (isinf(u)) logprob = u;
    else {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += exp(post[i] - u);
        logprob = log(s) + u;
    }
    for (int i = 0; i < n; ++i) post[i] = exp(post[i] - logprob);
    return n;
}

int vco_predict_proba(const vco_gmmmap* g, const double* x, double* posterior) {
    double* z = (double*)malloc(sizeof(double) * g->D);
    if (!z) return VCO_ENOMEM;
    int n = predict_proba_core(g, x, posterior, z);
    for (int i = n; i < g->M; ++i) posterior[i] = NAN; /* shorter vector in the reference */
    free(z);
    return VCO_OK;
}

/* src/gmm.jl:44-47 : indmax(posterior) -- first maximum, 1-based, index into the FILTERED list */
static int predict_core(const vco_gmmmap* g, const double* x, double* post, double* z) {
    int n = predict_proba_core(g, x, post, z);
    int best = 0;
    for (int i = 1; i < n; ++i) if (post[i] > post[best]) best = i;
    return best + 1;
}

int vco_predict(const vco_gmmmap* g, const double* x, int* mhat) {
    double* z = (double*)malloc(sizeof(double) * (g->D + g->M));
    if (!z) return VCO_ENOMEM;
    *mhat = predict_core(g, x, z + g->D, z);
    free(z);
    return VCO_OK;
}

/* src/gmmmap.jl:101-118 with caller-provided scratch (Ey (D,M), post (M), z (D)) */
static int fvconvert_core(const vco_gmmmap* g, const double* x, double* y, double* Ey,
                          double* post, double* z) {
    int D = g->D, M = g->M;
    /* Eq. (11)  :109-111 */
    for (int m = 0; m < M; ++m) {
        const double* A = g->A + (size_t)m * D * D;
        const double* mx = g->mux + (size_t)m * D;
        const double* my = g->muy + (size_t)m * D;
        double* e = Ey + (size_t)m * D;
        for (int i = 0; i < D; ++i) z[i] = x[i] - mx[i];
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += A[i + (size_t)k * D] * z[k];
            e[i] = my[i] + s;
        }
    }
    /* Eq. (9)  :114 */
    int n = predict_proba_core(g, x, post, z);
    /* Eq. (13) :117  Ey * posterior -- a (D,M) x (n) product: DimensionMismatch if a zero
     * weight shortened the posterior */
    if (n != M) return VCO_EDIM;
    for (int i = 0; i < D; ++i) {
        double s = 0.0;
        for (int m = 0; m < M; ++m) s += Ey[i + (size_t)m * D] * post[m];
        y[i] = s;
    }
    return VCO_OK;
}

int vco_fvconvert(vco_gmmmap* g, const double* x, int xlen, double* y) {
    if (xlen != g->D) return VCO_EDIM; /* src/gmmmap.jl:102 */
    double* z = (double*)malloc(sizeof(double) * g->D);
    if (!z) return VCO_ENOMEM;
    int rc = fvconvert_core(g, x, y, g->Ey, g->lpr, z);
    free(z);
    return rc;
}

/* src/common.jl:7-26 */
int vco_vc_fbf(vco_gmmmap* g, const double* fm, int rows, int64_t T, double* out) {
    if (rows - 1 != g->D) return VCO_EDIM;
    double* z = (double*)malloc(sizeof(double) * g->D);
    if (!z) return VCO_ENOMEM;
    int rc = VCO_OK;
    for (int64_t t = 0; t < T && rc == VCO_OK; ++t)                        /* :17-19 */
        rc = fvconvert_core(g, fm + t * rows + 1, out + t * rows + 1, g->Ey, g->lpr, z);
    for (int64_t t = 0; t < T; ++t) out[t * rows] = fm[t * rows];          /* :23 */
    free(z);
    return rc;
}

int vco_vc_fbf_mt(const vco_gmmmap* g, const double* fm, int rows, int64_t T, double* out,
                  int nthreads) {
    if (rows - 1 != g->D) return VCO_EDIM;
    int rc = VCO_OK;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        double* scratch = (double*)malloc(sizeof(double) * ((size_t)g->D * g->M + g->M + g->D));
        double* Ey = scratch; double* post = Ey + (size_t)g->D * g->M; double* z = post + g->M;
#pragma omp for schedule(static)
        for (int64_t t = 0; t < T; ++t) {
            int r = scratch ? fvconvert_core(g, fm + t * rows + 1, out + t * rows + 1, Ey, post, z)
                            : VCO_ENOMEM;
            if (r != VCO_OK) {
#pragma omp critical
                rc = r;
            }
            out[t * rows] = fm[t * rows];
        }
        free(scratch);
    }
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* TrajectoryGMMMap -- src/trajectory_gmmmap.jl:1-110                                          */
/* ------------------------------------------------------------------------------------------ */

struct vco_traj {
    vco_gmmmap* g;
    int T;      /* current number of column-blocks of W (length(t), :34) -- mutable state (Q3) */
    double* Dy; /* (2Ds,2Ds,M)  :24-28 */
};

int vco_traj_create(vco_gmmmap* g, int T, vco_traj** out) {
    if (!g || !out || T < 1) return VCO_EARG;
    int D2 = g->D, M = g->M; /* D2 = 2*Ds = dim(g) */
    size_t DD = (size_t)D2 * D2;
    vco_traj* t = (vco_traj*)calloc(1, sizeof(*t));
    if (!t) return VCO_ENOMEM;
    t->g = g; t->T = T;
    t->Dy = (double*)malloc(sizeof(double) * DD * M);
    double* tmp = (double*)malloc(sizeof(double) * DD);
    if (!t->Dy || !tmp) { free(tmp); vco_traj_destroy(t); return VCO_ENOMEM; }
    for (int m = 0; m < M; ++m) {
        /* :26  Dy = Syy - (SyxSxx^-1) * Sxy ; :27 Dy = Dy^-1 */
        matmul(g->A + m * DD, g->Sxy + m * DD, tmp, D2, D2, D2);
        double* d = t->Dy + m * DD;
        for (size_t i = 0; i < DD; ++i) d[i] = g->Syy[m * DD + i] - tmp[i];
        int rc = lu_inverse(d, D2);
        if (rc != VCO_OK) { free(tmp); vco_traj_destroy(t); return rc; }
    }
    free(tmp);
    *out = t;
    return VCO_OK;
}

void vco_traj_destroy(vco_traj* t) { if (t) { free(t->Dy); free(t); } }
int vco_traj_length(const vco_traj* t) { return t->T; }     /* :34 */
int vco_traj_dim(const vco_traj* t) { return t->g->D; }     /* :35 */
const double* vco_traj_Dy(const vco_traj* t) { return t->Dy; }

/* Non-zeros of row `i` (0..2D-1) of the t-th (0-based) row-block of W: compute_wt  :39-53 */
static int w_row(int t, int i, int D, int T, int64_t* col, double* val) {
    if (i < D) { col[0] = (int64_t)t * D + i; val[0] = 1.0; return 1; }   /* w0 = I at block t */
    int n = 0, k = i - D;
    if (t >= 1) { col[n] = (int64_t)(t - 1) * D + k; val[n] = -0.5; ++n; } /* :45-47 (t >= 2, 1-based) */
    if (t < T - 1) { col[n] = (int64_t)(t + 1) * D + k; val[n] = 0.5; ++n; } /* :48-50 */
    return n;
}

int64_t vco_constructW(int D, int T, int64_t* rows, int64_t* cols, double* vals) {
    if (D < 1 || T < 1) return -1;
    int64_t nnz = 0;
    for (int t = 0; t < T; ++t)
        for (int i = 0; i < 2 * D; ++i) {
            int64_t c[2]; double v[2];
            int n = w_row(t, i, D, T, c, v);
            for (int k = 0; k < n; ++k) {
                if (rows) { rows[nnz] = (int64_t)2 * D * t + i; cols[nnz] = c[k]; vals[nnz] = v[k]; }
                ++nnz;
            }
        }
    return nnz;
}

/* General band LU with partial pivoting + solve (what a sparse direct `\` on a non-Hermitian
 * banded matrix amounts to).  LAPACK-style band storage: (i,j) at ab[kl+ku+i-j + j*ldab],
 * ldab = 2*kl+ku+1, first kl rows are fill space (zero on entry). */
static int band_lu_solve(int n, int kl, int ku, double* ab, double* b) {
    int ldab = 2 * kl + ku + 1;
#define AB(i, j) ab[(size_t)(j) * ldab + (kl + ku + (i) - (j))]
    int* piv = (int*)malloc(sizeof(int) * (size_t)n);
    if (!piv) return VCO_ENOMEM;
    int ju = 0;
    for (int j = 0; j < n; ++j) {
        int km = (kl < n - 1 - j) ? kl : n - 1 - j;
        int jp = 0;
        double best = fabs(AB(j, j));
        for (int i = 1; i <= km; ++i) {
            double v = fabs(AB(j + i, j));
            if (v > best) { best = v; jp = i; }
        }
        piv[j] = j + jp;
        if (best == 0.0) { free(piv); return VCO_ESINGULAR; }
        int cand = j + ku + jp; if (cand > n - 1) cand = n - 1;
        if (cand > ju) ju = cand;
        if (jp != 0)
            for (int c = j; c <= ju; ++c) { double t = AB(j + jp, c); AB(j + jp, c) = AB(j, c); AB(j, c) = t; }
        double d = AB(j, j);
        for (int i = 1; i <= km; ++i) AB(j + i, j) /= d;
        for (int c = j + 1; c <= ju; ++c) {
            double t = AB(j, c);
            if (t != 0.0)
                for (int i = 1; i <= km; ++i) AB(j + i, c) -= AB(j + i, j) * t;
        }
    }
    for (int j = 0; j < n; ++j) {
        int km = (kl < n - 1 - j) ? kl : n - 1 - j;
        int p = piv[j];
        if (p != j) { double t = b[j]; b[j] = b[p]; b[p] = t; }
        double t = b[j];
        for (int i = 1; i <= km; ++i) b[j + i] -= AB(j + i, j) * t;
    }
    for (int j = n - 1; j >= 0; --j) {
        b[j] /= AB(j, j);
        double t = b[j];
        int lo = j - ku - kl; if (lo < 0) lo = 0;
        for (int i = lo; i < j; ++i) b[i] -= AB(i, j) * t;
    }
#undef AB
    free(piv);
    return VCO_OK;
}

/* src/trajectory_gmmmap.jl:65-110 */
/* keep_ab / keep_rhs (optional): malloc'ed copies of the band matrix R = W'Dy^-1 W (LAPACK band
 * storage, before factorisation) and of r = W'Dy^-1 Ey, for the GV gradient (:150-156). */
static int traj_core(vco_traj* tg, const double* X, int xrows, int T, double* Y, int* mhat_out,
                     double* Ey_out, double** keep_ab, double** keep_rhs, int* kl_out, int* ldab_out) {
    vco_gmmmap* g = tg->g;
    int D2 = g->D;
    if (xrows & 1) return VCO_EDIM;
    int D = xrows >> 1;                     /* :66 */
    if (2 * D != D2) return VCO_EDIM;       /* :67 */
    if (T < 1) return VCO_EARG;
    if (T != tg->T) tg->T = T;              /* :70-72 W rebuilt for the new length; state kept */

    int n = D * T, kl = 3 * D - 1, ku = kl, ldab = 2 * kl + ku + 1;
    if (kl > n - 1) { kl = ku = (n > 1 ? n - 1 : 0); ldab = 2 * kl + ku + 1; }
    int* mhat = (int*)malloc(sizeof(int) * (size_t)T);
    double* Ey = (double*)malloc(sizeof(double) * (size_t)D2 * T);
    double* ab = (double*)calloc((size_t)ldab * n, sizeof(double));
    double* rhs = (double*)calloc((size_t)n, sizeof(double));
    double* scratch = (double*)malloc(sizeof(double) * (size_t)(g->M + D2));
    if (!mhat || !Ey || !ab || !rhs || !scratch) {
        free(mhat); free(Ey); free(ab); free(rhs); free(scratch); return VCO_ENOMEM;
    }
    /* :82  mhat = predict(g.px, X) */
    for (int t = 0; t < T; ++t)
        mhat[t] = predict_core(g, X + (size_t)t * D2, scratch, scratch + g->M);
    /* :85-89  Ey[:,t] = muy[:,m] + A[:,:,m] * (X[:,t] - mux[:,m]) */
    for (int t = 0; t < T; ++t) {
        int m = mhat[t] - 1;
        const double* A = g->A + (size_t)m * D2 * D2;
        double* z = scratch;
        for (int i = 0; i < D2; ++i) z[i] = X[(size_t)t * D2 + i] - g->mux[(size_t)m * D2 + i];
        for (int i = 0; i < D2; ++i) {
            double s = 0.0;
            for (int k = 0; k < D2; ++k) s += A[i + (size_t)k * D2] * z[k];
            Ey[(size_t)t * D2 + i] = g->muy[(size_t)m * D2 + i] + s;
        }
    }
    /* :95  Dy^-1 = blkdiag(Dy[:,:,mhat[t]]...) ; :103-105  R = (W'Dy^-1) W, r = (W'Dy^-1) Ey.
     * W is never stored: its rows come from w_row() (the same rule vco_constructW emits). */
#define AB(i, j) ab[(size_t)(j) * ldab + (kl + ku + (i) - (j))]
    for (int t = 0; t < T; ++t) {
        const double* P = tg->Dy + (size_t)(mhat[t] - 1) * D2 * D2;
        for (int i = 0; i < D2; ++i) {
            int64_t ci[2]; double vi[2];
            int ni = w_row(t, i, D, T, ci, vi);
            for (int j = 0; j < D2; ++j) {
                double p = P[i + (size_t)j * D2];
                int64_t cj[2]; double vj[2];
                int nj = w_row(t, j, D, T, cj, vj);
                for (int a = 0; a < ni; ++a) {
                    rhs[ci[a]] += vi[a] * p * Ey[(size_t)t * D2 + j];
                    for (int b = 0; b < nj; ++b) AB(ci[a], cj[b]) += vi[a] * p * vj[b];
                }
            }
        }
    }
#undef AB
    if (keep_ab) {
        *keep_ab = (double*)malloc(sizeof(double) * (size_t)ldab * n);
        *keep_rhs = (double*)malloc(sizeof(double) * (size_t)n);
        if (!*keep_ab || !*keep_rhs) { free(mhat); free(Ey); free(ab); free(rhs); free(scratch); return VCO_ENOMEM; }
        memcpy(*keep_ab, ab, sizeof(double) * (size_t)ldab * n);
        memcpy(*keep_rhs, rhs, sizeof(double) * (size_t)n);
        *kl_out = kl; *ldab_out = ldab;
    }
    int rc = band_lu_solve(n, kl, ku, ab, rhs); /* :105  y = R \ r */
    if (rc == VCO_OK) memcpy(Y, rhs, sizeof(double) * (size_t)n); /* :109 reshape(y, D, T) */
    if (mhat_out) memcpy(mhat_out, mhat, sizeof(int) * (size_t)T);
    if (Ey_out) memcpy(Ey_out, Ey, sizeof(double) * (size_t)D2 * T);
    free(mhat); free(Ey); free(ab); free(rhs); free(scratch);
    return rc;
}

int vco_traj_fvconvert(vco_traj* tg, const double* X, int xrows, int T, double* Y, int* mhat_out,
                       double* Ey_out) {
    return traj_core(tg, X, xrows, T, Y, mhat_out, Ey_out, 0, 0, 0, 0);
}

/* ------------------------------------------------------------------------------------------ */
/* GV post-filter -- src/gv.jl:6-21                                                            */
/* ------------------------------------------------------------------------------------------ */

/* src[:,:] = sqrt(s2 ./ var(src, 2)) .* (src .- mean(src, 2)) .+ mean(src, 2)   (src/gv.jl:10-15)
 * mean = sum/T, var = sum((x - mean)^2)/(T - 1) (Julia's corrected two-pass varm), both summed in
 * frame order.  T == 1 gives NaN exactly as in Julia (0/0). */
void vco_variance_scaling(const double* s2, double* src, int D, int64_t T) {
    for (int i = 0; i < D; ++i) {
        double sum = 0.0;
        for (int64_t t = 0; t < T; ++t) sum += src[t * D + i];
        const double mu = sum / (double)T;
        double ss = 0.0;
        for (int64_t t = 0; t < T; ++t) { const double d = src[t * D + i] - mu; ss += d * d; }
        const double var = ss / (double)(T - 1);
        const double sc = sqrt(s2[i] / var);
        for (int64_t t = 0; t < T; ++t) src[t * D + i] = sc * (src[t * D + i] - mu) + mu;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* TrajectoryGVGMMMap -- src/trajectory_gmmmap.jl:112-189                                      */
/* ------------------------------------------------------------------------------------------ */

struct vco_trajgv {
    vco_traj* t;   /* borrowed */
    int Ds;
    double* muv;   /* (Ds)      :116 */
    double* pv;    /* (Ds,Ds)   :118  inv(Svv) */
};

void vco_trajgv_destroy(vco_trajgv* v) { if (v) { free(v->muv); free(v->pv); free(v); } }

int vco_trajgv_create(vco_traj* t, const double* muv, const double* svv, vco_trajgv** out) {
    if (!t || !muv || !svv || !out) return VCO_EARG;
    const int Ds = t->g->D / 2;
    for (int i = 0; i < Ds; ++i)
        if (muv[i] < 0.0) return VCO_EARG;          /* :124  @assert sum(muv .< 0) == 0 */
    vco_trajgv* v = (vco_trajgv*)calloc(1, sizeof(*v));
    if (!v) return VCO_ENOMEM;
    v->t = t; v->Ds = Ds;
    v->muv = (double*)malloc(sizeof(double) * Ds);
    v->pv = (double*)malloc(sizeof(double) * (size_t)Ds * Ds);
    if (!v->muv || !v->pv) { vco_trajgv_destroy(v); return VCO_ENOMEM; }
    memcpy(v->muv, muv, sizeof(double) * Ds);
    memcpy(v->pv, svv, sizeof(double) * (size_t)Ds * Ds);
    int rc = lu_inverse(v->pv, Ds);                  /* :125  inv(Svv) */
    if (rc != VCO_OK) { vco_trajgv_destroy(v); return rc; }
    *out = v;
    return VCO_OK;
}

/* fvconvert(tgv, X; epochs, alpha)  :140-172, gvgrad :175-189 */
int vco_trajgv_fvconvert(vco_trajgv* v, const double* X, int xrows, int T, int epochs, double alpha,
                         double* Y) {
    const int D = v->Ds;
    double *ab = 0, *r = 0;
    int kl = 0, ldab = 0;
    int rc = traj_core(v->t, X, xrows, T, Y, 0, 0, &ab, &r, &kl, &ldab);   /* :146 y0 */
    if (rc != VCO_OK) { free(ab); free(r); return rc; }
    const int n = D * T, ku = kl;
    vco_variance_scaling(v->muv, Y, D, T);                                   /* :150 eq. (58) */
    const double omega = 1.0 / (2.0 * (double)T);                            /* :152 */
    double* dy = (double*)malloc(sizeof(double) * (size_t)n);
    double* gv = (double*)malloc(sizeof(double) * 3 * (size_t)D);
    if (!dy || !gv) { free(ab); free(r); free(dy); free(gv); return VCO_ENOMEM; }
    double *muy = gv + D, *coef = gv + 2 * D;
    for (int epoch = 0; epoch < epochs; ++epoch) {
        /* gvgrad: gv = var(y, 2), muy = mean(y, 2); v[:,t] = -2/T * (pv' (gv - muv)) .* (y[:,t] - muy) */
        for (int i = 0; i < D; ++i) {
            double sum = 0.0;
            for (int t = 0; t < T; ++t) sum += Y[(size_t)t * D + i];
            muy[i] = sum / (double)T;
            double ss = 0.0;
            for (int t = 0; t < T; ++t) { const double d = Y[(size_t)t * D + i] - muy[i]; ss += d * d; }
            gv[i] = ss / (double)(T - 1);
        }
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += v->pv[k + (size_t)i * D] * (gv[k] - v->muv[k]);  /* pv' */
            coef[i] = -2.0 / (double)T * s;
        }
        /* :161  dy = omega * (-(W'Dy^-1 W) y + W'Dy^-1 Ey) + vec(gvgrad) */
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            const int j0 = i - kl > 0 ? i - kl : 0, j1 = i + ku < n - 1 ? i + ku : n - 1;
            for (int j = j0; j <= j1; ++j) s += ab[(size_t)j * ldab + (kl + ku + i - j)] * Y[j];
            dy[i] = omega * (-s + r[i]) + coef[i % D] * (Y[i] - muy[i % D]);
        }
        for (int i = 0; i < n; ++i) {
            if (dy[i] != dy[i]) { free(ab); free(r); free(dy); free(gv); return VCO_EARG; }  /* :165 @assert */
            Y[i] = Y[i] + alpha * dy[i];                                      /* :168 eq. (52) */
        }
    }
    free(ab); free(r); free(dy); free(gv);
    return VCO_OK;
}

/* vc(c::TrajectoryConverter, fm) with a GV converter: src/common.jl:31-63, default epochs / alpha */
int vco_vc_trajgv(vco_trajgv* v, const double* fm, int rows, int64_t T, int epochs, double alpha,
                  double* out) {
    int srows = rows - 1, Dout = (srows >> 1) + 1;
    if (T < 1) return VCO_EARG;
    int64_t limit = vco_traj_length(v->t);
    int rc = VCO_OK;
    int64_t count = 0;
    double* conv = (double*)malloc(sizeof(double) * (size_t)(Dout - 1) * (size_t)(limit < T ? limit : T));
    double* phrase = (double*)malloc(sizeof(double) * (size_t)srows * (size_t)(limit < T ? limit : T));
    if (!conv || !phrase) { free(conv); free(phrase); return VCO_ENOMEM; }
    for (;;) {
        int64_t b = count * limit, e = (count + 1) * limit < T ? (count + 1) * limit : T;
        int len = (int)(e - b);
        for (int t = 0; t < len; ++t) memcpy(phrase + (size_t)t * srows, fm + (b + t) * rows + 1, sizeof(double) * srows);
        rc = vco_trajgv_fvconvert(v, phrase, srows, len, epochs, alpha, conv);
        if (rc != VCO_OK) break;
        for (int t = 0; t < len; ++t)
            memcpy(out + (b + t) * Dout + 1, conv + (size_t)t * (Dout - 1), sizeof(double) * (Dout - 1));
        if (e == T) break;
        ++count;
    }
    for (int64_t t = 0; t < T; ++t) out[t * Dout] = fm[t * rows];
    free(conv); free(phrase);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* diffgmm -- src/diffgmm.jl:9-25 ([Kobayashi 2014] eqs. 6-8), restated on the joint parameters */
/* ------------------------------------------------------------------------------------------ */

/* GMMMapParam(w, mux, muy - mux, Sxx, Sxy - Sxx, (Sxy - Sxx)', Sxx + Syy - Sxy - Syx) written back as
 * a joint (2D, M) mean and (2D, 2D, M) covariance, so that GMMMap(w, mu', sigma') builds exactly
 * that parameter set. */
void vco_diffgmm(const double* mu, const double* sigma, int twoD, int M, double* mu_out, double* sigma_out) {
    const int D = twoD / 2;
    for (int m = 0; m < M; ++m) {
        const double* u = mu + (size_t)m * twoD;
        const double* S = sigma + (size_t)m * twoD * twoD;
        double* uo = mu_out + (size_t)m * twoD;
        double* So = sigma_out + (size_t)m * twoD * twoD;
#define SG(a, b) S[(a) + (size_t)(b) * twoD]
#define SO(a, b) So[(a) + (size_t)(b) * twoD]
        for (int i = 0; i < D; ++i) { uo[i] = u[i]; uo[D + i] = u[D + i] - u[i]; }
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) {
                SO(i, j) = SG(i, j);                                   /* Sxx */
                SO(i, D + j) = SG(i, D + j) - SG(i, j);                /* Sxy - Sxx */
                SO(D + j, i) = SG(i, D + j) - SG(i, j);                /* (Sxy - Sxx)' */
                SO(D + i, D + j) = SG(i, j) + SG(D + i, D + j) - SG(i, D + j) - SG(D + i, j);
            }
#undef SG
#undef SO
    }
}

/* src/common.jl:31-63 */
int vco_vc_traj(vco_traj* tg, const double* fm, int rows, int64_t T, double* out) {
    int srows = rows - 1;            /* :35 src = fm[2:end,:] */
    int Dout = (srows >> 1) + 1;     /* :38 */
    if (T < 1) return VCO_EARG;
    int64_t limit = vco_traj_length(tg); /* :43, read once */
    double* phrase = (double*)malloc(sizeof(double) * (size_t)srows * (size_t)(limit < T ? limit : T));
    double* conv = (double*)malloc(sizeof(double) * (size_t)(Dout - 1) * (size_t)(limit < T ? limit : T));
    if (!phrase || !conv) { free(phrase); free(conv); return VCO_ENOMEM; }
    int64_t count = 0;
    int rc = VCO_OK;
    for (;;) {
        int64_t b = count * limit;                                /* 0-based begin */
        int64_t e = (count + 1) * limit; if (e > T) e = T;        /* exclusive end */
        int len = (int)(e - b);
        for (int t = 0; t < len; ++t)
            memcpy(phrase + (size_t)t * srows, fm + (b + t) * rows + 1, sizeof(double) * srows);
        rc = vco_traj_fvconvert(tg, phrase, srows, len, conv, 0, 0);  /* :51 */
        if (rc != VCO_OK) break;
        for (int t = 0; t < len; ++t)
            memcpy(out + (b + t) * Dout + 1, conv + (size_t)t * (Dout - 1), sizeof(double) * (Dout - 1));
        if (e == T) break;
        ++count;
    }
    for (int64_t t = 0; t < T; ++t) out[t * Dout] = fm[t * rows];  /* :60 */
    free(phrase); free(conv);
    return rc;
}

int vco_vc_traj_batch_mt(vco_gmmmap* g, int limit, const double* fm, int rows,
                         const int64_t* offsets, int64_t nseq, double* out, int nthreads) {
    int Dout = ((rows - 1) >> 1) + 1;
    int rc = VCO_OK;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
    for (int64_t s = 0; s < nseq; ++s) {
        vco_traj* t = 0;
        int r = vco_traj_create(g, limit, &t);
        if (r == VCO_OK)
            r = vco_vc_traj(t, fm + offsets[s] * rows, rows, offsets[s + 1] - offsets[s],
                            out + offsets[s] * Dout);
        vco_traj_destroy(t);
        if (r != VCO_OK) {
#pragma omp critical
            rc = r;
        }
    }
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* DTW -- src/dtw.jl                                                                           */
/* ------------------------------------------------------------------------------------------ */

struct vco_dtw {
    int fstep, bstep;        /* :12-13 */
    int D, S, ncols;         /* template is (D,S); tables are (S, ncols) */
    double* tmpl;
    double* cost;
    int64_t* bp;
};

int vco_dtw_create(int fstep, int bstep, vco_dtw** out) {    /* :19-21 */
    if (!out) return VCO_EARG;
    vco_dtw* d = (vco_dtw*)calloc(1, sizeof(*d));
    if (!d) return VCO_ENOMEM;
    d->fstep = fstep; d->bstep = bstep;
    *out = d;
    return VCO_OK;
}

void vco_dtw_destroy(vco_dtw* d) { if (d) { free(d->tmpl); free(d->cost); free(d->bp); free(d); } }

/* :23-31  transition(d, from, to) */
static inline double dtw_transition(int from, int to) {
    if (to == from + 1) return 0.0;
    if (from == to) return 1.0;
    return 2.0;
}

/* :33-35  observation = sumabs2(v - template[:,i]); strict left-to-right, no FMA */
static inline double dtw_observation(const double* v, const double* tcol, int D) {
    double s = 0.0;
    for (int k = 0; k < D; ++k) {
        double df = v[k] - tcol[k];
        double sq = df * df;
        s = s + sq;
    }
    return s;
}

/* one column of the recurrence  :104-125 (fit!) == :68-86 (update!) ; 1-based state ids */
static void dtw_column(const vco_dtw* d, const double* v, const double* last, double* cur,
                       int64_t* curbp) {
    int S = d->S;
    for (int i = 1; i <= S; ++i) {
        int minindex = i;
        double ocost = dtw_observation(v, d->tmpl + (size_t)(i - 1) * d->D, d->D);
        double tcost = dtw_transition(minindex, i);
        double mincost = last[minindex - 1] + ocost + tcost;
        for (int j = i - d->bstep; j <= i + d->fstep; ++j) {
            if (j < 1 || j > S) continue;
            double c = last[j - 1] + ocost + dtw_transition(j, i);
            if (c < mincost) { mincost = c; minindex = j; }
        }
        cur[i - 1] = mincost;
        curbp[i - 1] = minindex;
    }
}

static int dtw_set_tmpl(vco_dtw* d, const double* tmpl, int D, int S) {
    double* t = (double*)malloc(sizeof(double) * (size_t)D * S);
    if (!t) return VCO_ENOMEM;
    memcpy(t, tmpl, sizeof(double) * (size_t)D * S);
    free(d->tmpl);
    d->tmpl = t; d->D = D; d->S = S;
    return VCO_OK;
}

static int dtw_alloc_tables(vco_dtw* d, int S, int ncols) {
    free(d->cost); free(d->bp);
    d->cost = (double*)calloc((size_t)S * ncols, sizeof(double));
    d->bp = (int64_t*)malloc(sizeof(int64_t) * (size_t)S * ncols);
    if (!d->cost || !d->bp) return VCO_ENOMEM;
    for (size_t i = 0; i < (size_t)S * ncols; ++i) d->bp[i] = 1;   /* :47 ones(Int,S,T+1) */
    for (int i = 0; i < S; ++i) { d->cost[i] = (double)(i + 1); d->bp[i] = i + 1; } /* :49-50 */
    d->ncols = ncols;
    return VCO_OK;
}

/* :133-145 */
int vco_dtw_backward(const vco_dtw* d, int64_t* path) {
    int T = d->ncols - 1, S = d->S;
    if (T < 1) return VCO_EARG;
    const double* lastcol = d->cost + (size_t)T * S;
    int best = 0;
    for (int i = 1; i < S; ++i) if (lastcol[i] < lastcol[best]) best = i;   /* indmin: first min */
    path[T - 1] = best + 1;
    for (int i = T; i >= 2; --i)                                            /* reverse(2:T) */
        path[i - 2] = d->bp[(size_t)i * S + (path[i - 1] - 1)];            /* backpointer[minpath[i], i+1] */
    return VCO_OK;
}

/* :93-128 */
int vco_dtw_fit(vco_dtw* d, const double* tmpl, int D, int S, const double* seq, int T,
                int64_t* path) {
    if (D < 1 || S < 1 || T < 1) return VCO_EARG;
    int rc = dtw_alloc_tables(d, S, T + 1);       /* :98 lazy_init!(d,S,T) */
    if (rc != VCO_OK) return rc;
    rc = dtw_set_tmpl(d, tmpl, D, S);             /* :102 */
    if (rc != VCO_OK) return rc;
    for (int t = 0; t < T; ++t)
        dtw_column(d, seq + (size_t)t * D, d->cost + (size_t)t * S, d->cost + (size_t)(t + 1) * S,
                   d->bp + (size_t)(t + 1) * S);
    return vco_dtw_backward(d, path);             /* :127 */
}

/* :53-56 + lazy_init!(d,S) :38-41 */
int vco_dtw_set_template(vco_dtw* d, const double* tmpl, int D, int S) {
    int rc = dtw_set_tmpl(d, tmpl, D, S);
    if (rc != VCO_OK) return rc;
    return dtw_alloc_tables(d, S, 1);
}

/* :61-90  (hcat growth) */
int vco_dtw_update(vco_dtw* d, const double* v, int vlen) {
    if (!d->tmpl || !d->cost) return VCO_EARG;
    if (vlen != d->D) return VCO_EDIM;
    int S = d->S, n = d->ncols;
    double* nc = (double*)realloc(d->cost, sizeof(double) * (size_t)S * (n + 1));
    if (!nc) return VCO_ENOMEM;
    d->cost = nc;
    int64_t* nb = (int64_t*)realloc(d->bp, sizeof(int64_t) * (size_t)S * (n + 1));
    if (!nb) return VCO_ENOMEM;
    d->bp = nb;
    dtw_column(d, v, d->cost + (size_t)(n - 1) * S, d->cost + (size_t)n * S, d->bp + (size_t)n * S);
    d->ncols = n + 1;
    return VCO_OK;
}

int vco_dtw_tables(const vco_dtw* d, int* S, int* ncols, const double** cost,
                   const int64_t** backptr) {
    if (S) *S = d->S;
    if (ncols) *ncols = d->ncols;
    if (cost) *cost = d->cost;
    if (backptr) *backptr = d->bp;
    return VCO_OK;
}

int vco_dtw_fit_batch_mt(const double* tmpl, const int64_t* tmpl_off, const double* seq,
                         const int64_t* seq_off, int64_t npairs, int D, int fstep, int bstep,
                         int64_t* paths, double* final_cost, int nthreads) {
    int rc = VCO_OK;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
    for (int64_t p = 0; p < npairs; ++p) {
        vco_dtw* d = 0;
        int r = vco_dtw_create(fstep, bstep, &d);
        int S = (int)(tmpl_off[p + 1] - tmpl_off[p]), T = (int)(seq_off[p + 1] - seq_off[p]);
        if (r == VCO_OK)
            r = vco_dtw_fit(d, tmpl + tmpl_off[p] * D, D, S, seq + seq_off[p] * D, T,
                            paths + seq_off[p]);
        if (r == VCO_OK && final_cost)
            final_cost[p] = d->cost[(size_t)T * S + (paths[seq_off[p] + T - 1] - 1)];
        vco_dtw_destroy(d);
        if (r != VCO_OK) {
#pragma omp critical
            rc = r;
        }
    }
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* "next" rows                                                                                 */
/* ------------------------------------------------------------------------------------------ */

/* src/datasets.jl:6-13 : repmat(src,2), then interior delta = -0.5 x[t-1] + 0.5 x[t+1];
 * frames 1 and T keep delta = copy of static (quirk Q2) */
void vco_push_delta(const double* src, int D, int T, double* out) {
    for (int t = 0; t < T; ++t)
        for (int i = 0; i < D; ++i) {
            out[(size_t)t * 2 * D + i] = src[(size_t)t * D + i];
            out[(size_t)t * 2 * D + D + i] = src[(size_t)t * D + i];
        }
    for (int t = 1; t < T - 1; ++t)
        for (int i = 0; i < D; ++i)
            out[(size_t)t * 2 * D + D + i] =
                -0.5 * src[(size_t)(t - 1) * D + i] + 0.5 * src[(size_t)(t + 1) * D + i];
}

/* src/align.jl:8-35 */
int vco_align(const double* src, int D, int S, const double* tgt, int T, double* newtgt,
              int64_t* path) {
    vco_dtw* d = 0;
    int rc = vco_dtw_create(0, 2, &d);                    /* :16 */
    if (rc != VCO_OK) return rc;
    rc = vco_dtw_fit(d, src, D, S, tgt, T, path);         /* :17 */
    vco_dtw_destroy(d);
    if (rc != VCO_OK) return rc;
    memset(newtgt, 0, sizeof(double) * (size_t)D * S);    /* :20 */
    for (int t = 0; t < T; ++t)                           /* :21 later duplicate wins */
        memcpy(newtgt + (size_t)(path[t] - 1) * D, tgt + (size_t)t * D, sizeof(double) * D);
    /* :25 hole = setdiff(path[1]:path[end], path), visited in increasing order */
    char* hit = (char*)calloc((size_t)S + 2, 1);
    if (!hit) return VCO_ENOMEM;
    for (int t = 0; t < T; ++t) hit[path[t]] = 1;
    for (int64_t i = path[0]; i <= path[T - 1]; ++i) {
        if (hit[i]) continue;
        if (i > 1 && i < S)                               /* :27 */
            for (int j = 0; j < D; ++j)
                newtgt[(size_t)(i - 1) * D + j] =
                    (newtgt[(size_t)(i - 2) * D + j] + newtgt[(size_t)i * D + j]) / 2.0;
    }
    free(hit);
    return VCO_OK;
}

int vco_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
