"""Seeded synthetic workloads for the BASELINE.json configurations.

Everything is Float64, generated on the host with ``numpy.random.default_rng(seed)`` and shared
verbatim by the CUDA path, the parity tests and the CPU baseline.  Shapes follow the reference's
Julia conventions (column-major ``(D, T)`` feature matrices; ``(2D, M)`` means; ``(2D, 2D, M)``
covariances).  Nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

from typing import NamedTuple, Tuple

import numpy as np


class JointGMM(NamedTuple):
    weights: np.ndarray  # (M,)
    means: np.ndarray    # (2D, M)  column-major
    covars: np.ndarray   # (2D, 2D, M)


def random_joint_gmm(seed: int, M: int, joint_dim: int, lam_lo: float = 1e-4, lam_hi: float = 1.0,
                     mean_scale: float = 1.0) -> JointGMM:
    """SPD joint GMM: weights ~ normalised Gamma(2); means ~ N(0,1) * mean_scale / (1 + k/4) per
    coefficient k; covariance Q diag(lambda) Q' with Q random orthogonal and lambda log-uniform in
    [lam_lo, lam_hi] ("CMU-Arctic-shaped": cond <= lam_hi/lam_lo)."""
    rng = np.random.default_rng(seed)
    w = rng.gamma(2.0, size=M)
    w /= w.sum()
    half = joint_dim // 2
    k = np.concatenate([np.arange(half), np.arange(joint_dim - half)])
    means = rng.standard_normal((joint_dim, M)) * (mean_scale / (1.0 + k / 4.0))[:, None]
    covars = np.empty((joint_dim, joint_dim, M), order="F")
    for m in range(M):
        q, _ = np.linalg.qr(rng.standard_normal((joint_dim, joint_dim)))
        lam = np.exp(rng.uniform(np.log(lam_lo), np.log(lam_hi), size=joint_dim))
        s = (q * lam) @ q.T
        covars[:, :, m] = 0.5 * (s + s.T)
    return JointGMM(w, np.asfortranarray(means), covars)


def sample_source_frames(gmm: JointGMM, T: int, seed: int, chunk: int = 65536) -> np.ndarray:
    """T frames from the model's own source marginal p(x); returns (D, T) column-major."""
    rng = np.random.default_rng(seed)
    D = gmm.means.shape[0] // 2
    M = gmm.weights.shape[0]
    L = np.stack([np.linalg.cholesky(gmm.covars[:D, :D, m]) for m in range(M)])
    mux = gmm.means[:D, :].T  # (M, D)
    X = np.empty((D, T), order="F")
    for b in range(0, T, chunk):
        e = min(T, b + chunk)
        comp = rng.choice(M, size=e - b, p=gmm.weights)
        z = rng.standard_normal((e - b, D))
        X[:, b:e] = (mux[comp] + np.einsum("tij,tj->ti", L[comp], z)).T
    return X


def fbf_feature_matrix(gmm: JointGMM, T: int, seed: int) -> np.ndarray:
    """(1+D, T) feature matrix for ``vc(::GMMMap, fm)``: row 1 = N(0,1) 'power', rest = frames."""
    rng = np.random.default_rng(seed + 7)
    X = sample_source_frames(gmm, T, seed)
    fm = np.empty((X.shape[0] + 1, T), order="F")
    fm[0, :] = rng.standard_normal(T)
    fm[1:, :] = X
    return fm


def push_delta(src: np.ndarray) -> np.ndarray:
    """src/datasets.jl:6-13 (boundary frames keep delta = copy of static)."""
    D, T = src.shape
    out = np.empty((2 * D, T), order="F")
    out[:D] = src
    out[D:] = src
    if T > 2:
        out[D:, 1:T - 1] = -0.5 * src[:, 0:T - 2] + 0.5 * src[:, 2:T]
    return out


def trajectory_utterances(gmm: JointGMM, n_utt: int, frames: int | Tuple[int, int], seed: int,
                          rho: float = 0.9, seg: int = 20):
    """Utterances for ``vc(::TrajectoryGMMMap, fm)``.  gmm has joint dim 4*Ds laid out
    [x_static; x_delta; y_static; y_delta].  Static tracks are AR(1)-smoothed (rho) excursions
    around a piecewise-constant mixture-mean sequence, then ``push_delta``.

    Returns (fm, offsets): fm (1+2Ds, total_frames) column-major, offsets (n_utt+1,) int64.
    """
    rng = np.random.default_rng(seed)
    Ds = gmm.means.shape[0] // 4
    M = gmm.weights.shape[0]
    Ls = np.stack([np.linalg.cholesky(gmm.covars[:Ds, :Ds, m]) for m in range(M)])
    if isinstance(frames, int):
        lens = np.full(n_utt, frames, dtype=np.int64)
    else:
        lens = rng.integers(frames[0], frames[1] + 1, size=n_utt).astype(np.int64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    fm = np.empty((1 + 2 * Ds, int(offsets[-1])), order="F")
    for u in range(n_utt):
        T = int(lens[u])
        nseg = (T + seg - 1) // seg
        comp = np.repeat(rng.choice(M, size=nseg, p=gmm.weights), seg)[:T]
        eps = np.einsum("tij,tj->ti", Ls[comp], rng.standard_normal((T, Ds)))
        noise = np.empty_like(eps)
        noise[0] = eps[0]
        c = np.sqrt(1.0 - rho * rho)
        for t in range(1, T):
            noise[t] = rho * noise[t - 1] + c * eps[t]
        static = (gmm.means[:Ds, comp].T + noise).T  # (Ds, T)
        b, e = int(offsets[u]), int(offsets[u + 1])
        fm[0, b:e] = rng.standard_normal(T)
        fm[1:, b:e] = push_delta(np.asfortranarray(static))
    return fm, offsets


def c4_utterances(gmm: JointGMM, utt_ids, frames: int, seed: int = 1004, rho: float = 0.9, seg: int = 20):
    """Utterances of the C4 batch BY INDEX: utterance ``u`` depends only on ``(seed, u)``, so a rank
    generates exactly its own shard of the 8192-utterance batch (same statistics as
    ``trajectory_utterances``: piecewise-constant mixture sequence, AR(1)-smoothed excursions,
    ``push_delta``).  Returns (fm (1+2Ds, n*frames) column-major, offsets (n+1,))."""
    from scipy.signal import lfilter
    utt_ids = np.asarray(utt_ids, dtype=np.int64)
    Ds = gmm.means.shape[0] // 4
    M = gmm.weights.shape[0]
    Ls = np.stack([np.linalg.cholesky(gmm.covars[:Ds, :Ds, m]) for m in range(M)])
    T = int(frames)
    n = len(utt_ids)
    offsets = (np.arange(n + 1, dtype=np.int64) * T)
    fm = np.empty((1 + 2 * Ds, n * T), order="F")
    nseg = (T + seg - 1) // seg
    c = np.sqrt(1.0 - rho * rho)
    for k, u in enumerate(utt_ids):
        rng = np.random.default_rng([seed, int(u)])
        comp = np.repeat(rng.choice(M, size=nseg, p=gmm.weights), seg)[:T]
        eps = np.einsum("tij,tj->ti", Ls[comp], rng.standard_normal((T, Ds)))
        drive = c * eps
        drive[0] = eps[0]
        noise = lfilter([1.0], [1.0, -rho], drive, axis=0)
        static = (gmm.means[:Ds, comp].T + noise).T
        b = k * T
        fm[0, b:b + T] = rng.standard_normal(T)
        fm[1:, b:b + T] = push_delta(np.asfortranarray(static))
    return fm, offsets


def dtw_pairs(n_pairs: int, dim: int, len_range: Tuple[int, int], seed: int, noise: float = 0.05):
    """Parallel-utterance pairs for ``DTWs.fit!``: template = smoothed random walk (dim, S);
    sequence = template resampled along a random monotone warp (local rate in [0.5, 2]) + noise.

    Returns (tmpl, tmpl_off, seq, seq_off) -- ragged column-wise concatenations + offsets.
    """
    rng = np.random.default_rng(seed)
    S = rng.integers(len_range[0], len_range[1] + 1, size=n_pairs)
    T = rng.integers(len_range[0], len_range[1] + 1, size=n_pairs)
    toff = np.concatenate([[0], np.cumsum(S)]).astype(np.int64)
    soff = np.concatenate([[0], np.cumsum(T)]).astype(np.int64)
    tm = np.empty((dim, int(toff[-1])), order="F")
    sq = np.empty((dim, int(soff[-1])), order="F")
    kern = np.hanning(9)
    kern /= kern.sum()
    for p in range(n_pairs):
        s, t = int(S[p]), int(T[p])
        walk = np.cumsum(rng.standard_normal((dim, s + 8)) * 0.3, axis=1)
        tpl = np.stack([np.convolve(walk[k], kern, mode="valid") for k in range(dim)])  # (dim, s)
        rate = np.exp(rng.uniform(np.log(0.5), np.log(2.0), size=t))
        pos = np.cumsum(rate)
        pos = (pos - pos[0]) / max(pos[-1] - pos[0], 1e-9) * (s - 1)
        lo = np.floor(pos).astype(np.int64).clip(0, s - 2 if s > 1 else 0)
        fr = pos - lo
        hi = np.minimum(lo + 1, s - 1)
        sqp = tpl[:, lo] * (1 - fr) + tpl[:, hi] * fr + noise * rng.standard_normal((dim, t))
        tm[:, toff[p]:toff[p + 1]] = tpl
        sq[:, soff[p]:soff[p + 1]] = sqp
    return tm, toff, sq, soff


# --- the BASELINE.json configurations -----------------------------------------------------------

def config_c1(T: int = 1_000_000, stress: bool = False):
    """C1: 24-dim mcep, 64-mixture full-cov joint GMM, frame-by-frame conversion of T frames."""
    gmm = random_joint_gmm(1001, 64, 48, *((1e-7, 2.0) if stress else (1e-4, 1.0)))
    return gmm, fbf_feature_matrix(gmm, T, 1001)


def config_c2(n_utt: int = 1000, frames: int = 500, M: int = 64, seed: int = 1002):
    """C2: static+delta (48-dim source), 64 mixtures, n_utt utterances x frames."""
    gmm = random_joint_gmm(seed, M, 96)
    fm, off = trajectory_utterances(gmm, n_utt, frames, seed)
    return gmm, fm, off


def config_c3(n_pairs: int = 1000, seed: int = 1003):
    """C3: DTW of n_pairs parallel utterance pairs, ~600x600 frames, 24-dim."""
    return dtw_pairs(n_pairs, 24, (550, 650), seed)


def config_c4(n_utt: int = 8192, frames: int = 500):
    """C4: 128-mixture trajectory conversion, 8192 utterances (1024 per GPU on 8 GPUs)."""
    return config_c2(n_utt, frames, M=128, seed=1004)
