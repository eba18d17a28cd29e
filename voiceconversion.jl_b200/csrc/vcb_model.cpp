// vcb_model.cpp -- K0: one-time model preprocessing on the host in Float64.
//
// Replaces GMMMap(weights, mu, Sigma; swap) (reference src/gmmmap.jl:62-90): split_joint_gmm
// (:41-52), GMMMapParam's A_m = Syx_m * Sxx_m^-1 (:34-36, dense LU inverse), and
// GaussianMixtureModel (src/gmm.jl:8-20: MvNormal over Hermitian(Sxx, :U) => Cholesky, PD check),
// then derives the operands the kernels consume:
//   Linv_m = chol(Sxx_m)^-1, offsets Linv_m (mux_m - xbar), b_m = muy_m - A_m (mux_m - xbar),
//   c_m = log w_m - (D log 2pi + log|Sxx_m|) / 2,
// and packs them (fp32 rows for the CUDA-core kernel, tf32 hi/lo images in the UMMA canonical
// K-major layout for the tcgen05 kernel).  Also TrajectoryGMMMap's Dy_m = (Syy_m - A_m Sxy_m)^-1
// (src/trajectory_gmmmap.jl:24-28).
#include "vcb_model.h"
#include "vcb_kernels.h"

#include <cmath>
#include <cstring>
#include <limits>

namespace vcb {

// In-place inverse of a general n x n column-major matrix: Gauss-Jordan with partial (row)
// pivoting.  Returns false when a pivot is exactly zero (Julia: SingularException).
static bool invert_general(std::vector<double>& a, int n) {
    std::vector<double> aug((size_t)n * 2 * n);
    auto at = [&](int r, int c) -> double& { return aug[(size_t)r * 2 * n + c]; };  // row-major
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
            at(r, c) = a[r + (size_t)c * n];
            at(r, n + c) = (r == c) ? 1.0 : 0.0;
        }
    for (int col = 0; col < n; ++col) {
        int piv = col;
        double best = std::fabs(at(col, col));
        for (int r = col + 1; r < n; ++r)
            if (std::fabs(at(r, col)) > best) { best = std::fabs(at(r, col)); piv = r; }
        if (best == 0.0 || !std::isfinite(best)) return false;
        if (piv != col)
            for (int c = 0; c < 2 * n; ++c) std::swap(at(piv, c), at(col, c));
        double inv = 1.0 / at(col, col);
        for (int c = 0; c < 2 * n; ++c) at(col, c) *= inv;
        for (int r = 0; r < n; ++r) {
            if (r == col) continue;
            double f = at(r, col);
            if (f == 0.0) continue;
            for (int c = 0; c < 2 * n; ++c) at(r, c) -= f * at(col, c);
        }
    }
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) a[r + (size_t)c * n] = at(r, n + c);
    return true;
}

bool invert_matrix(std::vector<double>& a, int n) { return invert_general(a, n); }

// GMMMapParam(w, mux, muy - mux, Sxx, Sxy - Sxx, (Sxy - Sxx)', Sxx + Syy - Sxy - Syx)  ([Kobayashi 2014]
// eqs. 6-8) written back as joint parameters, so that vcb_gmmmap_create builds exactly that model.
void diffgmm_params(const double* mu, const double* sigma, int twoD, int M, double* mu_out, double* sigma_out) {
    const int D = twoD / 2;
    for (int m = 0; m < M; ++m) {
        const double* u = mu + (size_t)m * twoD;
        const double* S = sigma + (size_t)m * twoD * twoD;
        double* uo = mu_out + (size_t)m * twoD;
        double* So = sigma_out + (size_t)m * twoD * twoD;
        auto s = [&](int r, int c) { return S[r + (size_t)c * twoD]; };
        for (int i = 0; i < D; ++i) {
            uo[i] = u[i];
            uo[D + i] = u[D + i] - u[i];
        }
        for (int c = 0; c < D; ++c)
            for (int r = 0; r < D; ++r) {
                const double xy = s(r, D + c) - s(r, c);
                So[r + (size_t)c * twoD] = s(r, c);
                So[r + (size_t)(D + c) * twoD] = xy;
                So[(D + c) + (size_t)r * twoD] = xy;
                So[(D + r) + (size_t)(D + c) * twoD] = s(r, c) + s(D + r, D + c) - s(r, D + c) - s(D + r, c);
            }
    }
}

// Lower Cholesky factor of the matrix whose UPPER triangle is stored in s (Hermitian(s, :U),
// src/gmm.jl:16).  Row-major output l[i*n + j].  Returns false if not positive definite.
static bool cholesky_from_upper(const double* s, int n, std::vector<double>& l) {
    l.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j <= i; ++j) {
            double acc = s[j + (size_t)i * n];  // element (j,i), j <= i: upper triangle
            for (int k = 0; k < j; ++k) acc -= l[(size_t)i * n + k] * l[(size_t)j * n + k];
            if (i == j) {
                if (!(acc > 0.0) || !std::isfinite(acc)) return false;
                l[(size_t)i * n + i] = std::sqrt(acc);
            } else {
                l[(size_t)i * n + j] = acc / l[(size_t)j * n + j];
            }
        }
    }
    return true;
}

// Inverse of a lower-triangular row-major matrix (row-major output).
static void invert_lower(const std::vector<double>& l, int n, double* out) {
    std::memset(out, 0, sizeof(double) * (size_t)n * n);
    for (int c = 0; c < n; ++c) {
        out[(size_t)c * n + c] = 1.0 / l[(size_t)c * n + c];
        for (int r = c + 1; r < n; ++r) {
            double acc = 0.0;
            for (int k = c; k < r; ++k) acc += l[(size_t)r * n + k] * out[(size_t)k * n + c];
            out[(size_t)r * n + c] = -acc / l[(size_t)r * n + r];
        }
    }
}

static inline float tf32_rn(float v) {
    uint32_t u;
    std::memcpy(&u, &v, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return v;  // inf / nan
    u += 0x1000u;
    u &= 0xFFFFE000u;
    std::memcpy(&v, &u, 4);
    return v;
}

void tf32_split(double v, float& hi, float& lo) {
    hi = tf32_rn((float)v);
    lo = tf32_rn((float)(v - (double)hi));
}

// Index (in floats) of element (row n, reduction index k) inside a K-major, non-swizzled UMMA
// operand image with `rows` rows: 8x(16 B) core matrices, SBO = 128 B between 8-row groups,
// LBO = rows*16 B between 16-byte K slices.  (Same mapping is used by the kernel's A loader.)
static inline size_t umma_kmajor_index(int n, int k, int rows) {
    return (size_t)(k >> 2) * ((size_t)rows * 4) + (size_t)(n >> 3) * 32 + (size_t)(n & 7) * 4 + (k & 3);
}

int32_t build_gmmmap(const double* weights, const double* mu, const double* sigma, int twoD, int M,
                     int swap, vcb_gmmmap& g) {
    if (!weights || !mu || !sigma) return fail(VCB_EARG, "null model pointer");
    if (twoD < 2 || M < 1) return fail(VCB_EARG, "bad model size (twoD=%d, M=%d)", twoD, M);
    // src/gmmmap.jl:43: D = size(mu,1)>>1; an odd joint dimension fails in GMMMapParam (:35)
    if (twoD & 1) return fail(VCB_EDIM, "joint dimension %d is odd", twoD);
    const int D = twoD / 2;
    const size_t DD = (size_t)D * D;
    g.D = D;
    g.M = M;
    g.DP = round_up(D, 8);
    g.DS = simt_padded_dim(D);
    g.KS = round_up(g.DS + 1, 4);

    // MixtureModel(normals, weights) needs a probability vector (ext: Distributions.isprobvec).
    // Zero weights make the reference's fvconvert throw (src/gmm.jl:26-27 + src/gmmmap.jl:117,
    // SURVEY H8/Q1), so they are rejected here.
    double wsum = 0.0;
    for (int m = 0; m < M; ++m) {
        if (!(weights[m] > 0.0) || !std::isfinite(weights[m]))
            return fail(VCB_EARG, "weights[%d] = %g: weights must be positive (zero-weight components make the reference's fvconvert throw)", m, weights[m]);
        wsum += weights[m];
    }
    if (!(std::fabs(wsum - 1.0) <= 1.4901161193847656e-08 * std::fmax(std::fabs(wsum), 1.0)))
        return fail(VCB_EARG, "weights sum to %.17g, not a probability vector", wsum);

    g.src_w.assign(weights, weights + M);
    g.src_mu.assign(mu, mu + (size_t)twoD * M);
    g.src_sigma.assign(sigma, sigma + (size_t)twoD * twoD * M);
    g.src_swap = swap;
    g.w.assign(weights, weights + M);
    g.mux.resize((size_t)D * M);
    g.muy.resize((size_t)D * M);
    g.Sxx.resize(DD * M);
    g.Sxy.resize(DD * M);
    g.Syx.resize(DD * M);
    g.Syy.resize(DD * M);
    g.A.resize(DD * M);
    // split_joint_gmm (src/gmmmap.jl:41-52) + swap (:74-78)
    const int xo = swap ? D : 0, yo = swap ? 0 : D;
    for (int m = 0; m < M; ++m) {
        const double* mm = mu + (size_t)m * twoD;
        const double* sm = sigma + (size_t)m * twoD * twoD;
        for (int i = 0; i < D; ++i) {
            g.mux[(size_t)m * D + i] = mm[xo + i];
            g.muy[(size_t)m * D + i] = mm[yo + i];
        }
        for (int c = 0; c < D; ++c)
            for (int r = 0; r < D; ++r) {
                size_t o = m * DD + r + (size_t)c * D;
                g.Sxx[o] = sm[(xo + r) + (size_t)(xo + c) * twoD];
                g.Sxy[o] = sm[(xo + r) + (size_t)(yo + c) * twoD];
                g.Syx[o] = sm[(yo + r) + (size_t)(xo + c) * twoD];
                g.Syy[o] = sm[(yo + r) + (size_t)(yo + c) * twoD];
            }
    }
    // centring vector: weighted mean of the source means (subtracted from x in Float64 on device)
    g.xbar.assign(D, 0.0);
    for (int m = 0; m < M; ++m)
        for (int i = 0; i < D; ++i) g.xbar[i] += g.w[m] * g.mux[(size_t)m * D + i];

    std::vector<double> inv(DD), chol, linv_all(DD * M), cst(M);
    std::vector<double> offw((size_t)D * M), offa((size_t)D * M);
    for (int m = 0; m < M; ++m) {
        // A_m = Syx_m * Sxx_m^-1  (dense inverse, src/gmmmap.jl:35)
        std::copy(g.Sxx.begin() + m * DD, g.Sxx.begin() + (m + 1) * DD, inv.begin());
        if (!invert_general(inv, D)) return fail(VCB_ESINGULAR, "Sxx[:,:,%d] is singular", m + 1);
        for (int c = 0; c < D; ++c)
            for (int r = 0; r < D; ++r) {
                double acc = 0.0;
                for (int k = 0; k < D; ++k) acc += g.Syx[m * DD + r + (size_t)k * D] * inv[k + (size_t)c * D];
                g.A[m * DD + r + (size_t)c * D] = acc;
            }
        // MvNormal(mux_m, Hermitian(Sxx_m)) -> Cholesky (src/gmm.jl:16-17)
        if (!cholesky_from_upper(g.Sxx.data() + m * DD, D, chol))
            return fail(VCB_ENOTPD, "Sxx[:,:,%d] is not positive definite", m + 1);
        double logdet = 0.0;
        for (int i = 0; i < D; ++i) logdet += std::log(chol[(size_t)i * D + i]);
        logdet *= 2.0;
        cst[m] = std::log(g.w[m]) - 0.5 * (D * 1.8378770664093454835606594728112 + logdet);
        double* li = linv_all.data() + m * DD;
        invert_lower(chol, D, li);
        for (int r = 0; r < D; ++r) {
            double ow = 0.0, oa = 0.0;
            for (int k = 0; k < D; ++k) {
                double dm = g.mux[(size_t)m * D + k] - g.xbar[k];
                ow += li[(size_t)r * D + k] * dm;
                oa += g.A[m * DD + r + (size_t)k * D] * dm;
            }
            offw[(size_t)m * D + r] = -ow;                            // z = Linv xc + offw
            offa[(size_t)m * D + r] = g.muy[(size_t)m * D + r] - oa;  // E = A xc + offa
        }
    }

    // ---- device uploads: Float64 exact-path operands
    std::vector<double> linv_cm(DD * M);
    for (int mm = 0; mm < M; ++mm)
        for (int r = 0; r < D; ++r)
            for (int k = 0; k < D; ++k) linv_cm[mm * DD + (size_t)k * D + r] = linv_all[mm * DD + (size_t)r * D + k];
    if (g.d_linv.upload(linv_all) != cudaSuccess || g.d_linv_cm.upload(linv_cm) != cudaSuccess || g.d_mux.upload(g.mux) != cudaSuccess ||
        g.d_muy.upload(g.muy) != cudaSuccess || g.d_A.upload(g.A) != cudaSuccess ||
        g.d_c.upload(cst) != cudaSuccess || g.d_xbar.upload(g.xbar) != cudaSuccess)
        return fail(VCB_ECUDA, "model upload failed: %s", cudaGetErrorString(cudaGetLastError()));

    // ---- fp32 rows for the CUDA-core kernel: [M][2*DP][KS], row = [offset | coeffs (D) | 0-pad]
    if (g.DS) {
        const int DP = g.DS, KS = g.KS;
        std::vector<float> w32((size_t)M * 2 * DP * KS, 0.0f), c32(M);
        for (int m = 0; m < M; ++m) {
            c32[m] = (float)cst[m];
            const double* li = linv_all.data() + m * DD;
            for (int r = 0; r < D; ++r) {
                float* zw = &w32[((size_t)m * 2 * DP + r) * KS];
                float* ew = &w32[((size_t)m * 2 * DP + DP + r) * KS];
                zw[0] = (float)offw[(size_t)m * D + r];
                ew[0] = (float)offa[(size_t)m * D + r];
                for (int k = 0; k < D; ++k) {
                    zw[1 + k] = (float)li[(size_t)r * D + k];
                    ew[1 + k] = (float)g.A[m * DD + r + (size_t)k * D];
                }
            }
        }
        if (g.d_w32.upload(w32) != cudaSuccess || g.d_c32.upload(c32) != cudaSuccess)
            return fail(VCB_ECUDA, "model upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    }

    // ---- tf32 hi/lo operand images for the tcgen05 kernels
    {
        vcb_tc_pack& tc = g.tc;
        const int DP = g.DP;
        // K layout: [xc (D) | 0-pad to 8] with the offset carried by two "ones" columns (B holds
        // the offset's tf32 hi part in the first and its lo part in the second, so ONE hi*hi MMA
        // adds it at full precision).  The ones columns sit in the padding of the last data k-step
        // when it has two spare columns, otherwise in an extra k-step that needs a single MMA
        // instead of the three passes of a data k-step.
        {
            const int kdata = round_up(D, 8);
            tc.c1 = (kdata - D >= 2) ? D : kdata;
            tc.KP = (kdata - D >= 2) ? kdata : kdata + 8;
            tc.koff = (kdata - D >= 2) ? 0 : 1;
        }
        // the whitening-only (arg-max) kernel runs CTA-pair MMAs, where a CTA stores half of every B
        // stage: plan its chunk width for that mode (the single-CTA fallback then runs on fewer stages)
        const TcPlan pc = tc_plan(M, tc.KP, 2 * DP, DP + 2), pw = tc_plan(M, tc.KP, DP, 4, DP == 48);
        tc.GC = pc.G; tc.NC = pc.N; tc.NCHC = pc.G ? (M + pc.G - 1) / pc.G : 0;
        tc.GW = pw.G; tc.NW = pw.N; tc.NCHW = pw.G ? (M + pw.G - 1) / pw.G : 0;
        auto put = [&](std::vector<float>& img, size_t base, int rows, int n, int k, double v) {
            float hi, lo;
            tf32_split(v, hi, lo);
            img[base + umma_kmajor_index(n, k, rows)] = hi;
            img[base + (size_t)rows * tc.KP + umma_kmajor_index(n, k, rows)] = lo;
        };
        auto put_off = [&](std::vector<float>& img, size_t base, int rows, int n, double v) {
            float hi, lo;
            tf32_split(v, hi, lo);
            img[base + umma_kmajor_index(n, tc.c1, rows)] = hi;       // both parts in the hi image
            img[base + umma_kmajor_index(n, tc.c1 + 1, rows)] = lo;
        };
        std::vector<float> bc((size_t)tc.NCHC * 2 * tc.NC * tc.KP, 0.0f);
        std::vector<float> bw((size_t)tc.NCHW * 2 * tc.NW * tc.KP, 0.0f);
        const int mpad = std::max(std::max(tc.NCHC * tc.GC, tc.NCHW * tc.GW), M);
        std::vector<float> cpad(mpad, -std::numeric_limits<float>::infinity());
        for (int m = 0; m < M; ++m) cpad[m] = (float)cst[m];
        for (int m = 0; m < M; ++m) {
            const double* li = linv_all.data() + m * DD;
            if (tc.GC) {
                const int ch = m / tc.GC, gi = m % tc.GC;
                const size_t base = (size_t)ch * 2 * tc.NC * tc.KP;
                for (int r = 0; r < D; ++r) {
                    const int nz = gi * 2 * DP + r, ne = gi * 2 * DP + DP + r;
                    for (int k = 0; k < D; ++k) {
                        put(bc, base, tc.NC, nz, k, li[(size_t)r * D + k]);
                        put(bc, base, tc.NC, ne, k, g.A[m * DD + r + (size_t)k * D]);
                    }
                    put_off(bc, base, tc.NC, nz, offw[(size_t)m * D + r]);
                    put_off(bc, base, tc.NC, ne, offa[(size_t)m * D + r]);
                }
            }
            if (tc.GW) {
                const int ch = m / tc.GW, gi = m % tc.GW;
                const size_t base = (size_t)ch * 2 * tc.NW * tc.KP;
                for (int r = 0; r < D; ++r) {
                    const int nz = gi * DP + r;
                    for (int k = 0; k < D; ++k) put(bw, base, tc.NW, nz, k, li[(size_t)r * D + k]);
                    put_off(bw, base, tc.NW, nz, offw[(size_t)m * D + r]);
                }
            }
        }
        // row-half split of every chunk image for the CTA-pair kernels (each CTA of a pair holds N/2
        // rows of B; LBO of the half image is (N/2)*16 bytes)
        auto split_pair = [&](const std::vector<float>& img, int nch, int n) {
            std::vector<float> out(img.size(), 0.0f);
            if (n <= 0 || (n % 16)) return out;
            const int h = n / 2, kc = tc.KP / 4;
            for (int ch = 0; ch < nch; ++ch) {
                const size_t cb = (size_t)ch * 2 * n * tc.KP;
                for (int part = 0; part < 2; ++part)
                    for (int j = 0; j < kc; ++j)
                        for (int row = 0; row < n; ++row) {
                            const int rk = row / h, r2 = row % h;
                            const size_t src = cb + (size_t)part * n * tc.KP + (size_t)j * n * 4 + (size_t)(row >> 3) * 32 + (row & 7) * 4;
                            const size_t dst = cb + (size_t)rk * n * tc.KP + (size_t)part * h * tc.KP + (size_t)j * h * 4 +
                                               (size_t)(r2 >> 3) * 32 + (r2 & 7) * 4;
                            for (int e = 0; e < 4; ++e) out[dst + e] = img[src + e];
                        }
            }
            return out;
        };
        const std::vector<float> bc2 = split_pair(bc, tc.NCHC, tc.NC), bw2 = split_pair(bw, tc.NCHW, tc.NW);
        if (tc.Bc.upload(bc) != cudaSuccess || tc.Bw.upload(bw) != cudaSuccess ||
            tc.Bc2.upload(bc2) != cudaSuccess || tc.Bw2.upload(bw2) != cudaSuccess ||
            tc.cst.upload(cpad) != cudaSuccess)
            return fail(VCB_ECUDA, "model upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    return VCB_OK;
}

int32_t build_traj(const vcb_gmmmap& g, vcb_traj& t) {
    if (g.D & 1) return fail(VCB_EDIM, "TrajectoryGMMMap needs static+delta features: dim(g) = %d is odd", g.D);
    const int D2 = g.D, M = g.M;
    const size_t DD = (size_t)D2 * D2;
    t.g = &g;
    t.device = g.device;
    t.Ds = D2 / 2;
    t.Dy.resize(DD * M);
    std::vector<double> d(DD), psym(DD * M);
    for (int m = 0; m < M; ++m) {
        // Dy = Syy - (Syx Sxx^-1) * Sxy ; Dy = Dy^-1   (src/trajectory_gmmmap.jl:26-27)
        for (int c = 0; c < D2; ++c)
            for (int r = 0; r < D2; ++r) {
                double acc = 0.0;
                for (int k = 0; k < D2; ++k) acc += g.A[m * DD + r + (size_t)k * D2] * g.Sxy[m * DD + k + (size_t)c * D2];
                d[r + (size_t)c * D2] = g.Syy[m * DD + r + (size_t)c * D2] - acc;
            }
        if (!invert_general(d, D2)) return fail(VCB_ESINGULAR, "Dy[:,:,%d] is singular", m + 1);
        std::copy(d.begin(), d.end(), t.Dy.begin() + m * DD);
        // the band solver is a Cholesky: use the symmetric part (the LU inverse is symmetric to
        // ~1e-12 relative)
        for (int c = 0; c < D2; ++c)
            for (int r = 0; r < D2; ++r)
                psym[m * DD + r + (size_t)c * D2] = 0.5 * (d[r + (size_t)c * D2] + d[c + (size_t)r * D2]);
    }
    if (t.d_P.upload(psym) != cudaSuccess || t.d_err.upload(std::vector<int>(1, 0)) != cudaSuccess)
        return fail(VCB_ECUDA, "model upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return VCB_OK;
}

}  // namespace vcb
