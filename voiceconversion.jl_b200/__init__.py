"""voiceconversion.jl_b200 -- host-side mirror of VoiceConversion.jl's conversion API over libvcb200.

Julia is not available in the build image, so this Python layer plays the role of the Julia shim
(``julia/VoiceConversionB200.jl``): same names, argument meaning and error behaviour as the
reference (``src/VoiceConversion.jl:12-38``), each method = argument checks + one call through the
C ABI (``include/vcb200.h``).  All arithmetic happens in the CUDA library; nothing here computes.

Array conventions
  * numpy arrays use the Julia shapes -- ``(D, T)`` feature matrices, ``(2D, M)`` means,
    ``(2D, 2D, M)`` covariances -- and are passed column-major (``order='F'``; a C-contiguous
    ``(T, D)`` array's ``.T`` is accepted without a copy).
  * torch CUDA tensors are "frame-major": a contiguous ``(T, rows)`` float64 tensor, which is the
    same memory as Julia's ``(rows, T)`` matrix.  They select the device-resident (``*_dev``) entry
    points on ``torch.cuda.current_stream()``.

The directory name contains a dot, so import it through the repo-root alias module ``vcb200``.
"""
from __future__ import annotations

import ctypes as C
from typing import List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import (ArgumentError, CudaError, DimensionMismatch, PosDefException, SingularException,  # noqa: F401
                   VCBError, device_count, init, launch_count, set_device, set_kernel_variant, stage_timing,
                   stage_times)
from . import dtws as DTWs  # noqa: N812  (Julia sub-module name, src/dtw.jl:1)
from . import jld, shard, synth  # noqa: F401

__all__ = [
    "AbstractConverter", "FrameByFrameConverter", "TrajectoryConverter", "GMMMapParam", "GMMMap",
    "TrajectoryGMMMap", "TrajectoryGVGMMMap", "VarianceScaling", "fvpostf", "fvpostf_", "diffgmm",
    "fvconvert", "fvconvert_gv", "vc", "vc_batch", "vc_static_batch", "ncomponents", "dim", "predict_proba",
    "predict", "constructW", "push_delta", "align", "align_batch", "DTWs", "DimensionMismatch",
    "PosDefException", "SingularException", "ArgumentError", "CudaError", "VCBError",
    "set_device", "device_count", "init", "set_kernel_variant", "launch_count", "traj_status", "pinned_empty",
]


def _is_torch(a) -> bool:
    return hasattr(a, "data_ptr") and hasattr(a, "is_cuda")


def _f64(a) -> np.ndarray:
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_dev_tensor(t, cols: Optional[int] = None):
    import torch
    if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.dim() == 2):
        raise ArgumentError(_lib.EARG, "device input must be a contiguous 2-D float64 CUDA tensor (T, rows)")
    if cols is not None and t.shape[1] != cols:
        raise DimensionMismatch(_lib.EDIM, "Inconsistent dimentions.")


# ---- type hierarchy (src/common.jl:2-4) -------------------------------------------------------
class AbstractConverter:
    pass


class FrameByFrameConverter(AbstractConverter):
    pass


class TrajectoryConverter(AbstractConverter):
    pass


class GMMMapParam(NamedTuple):
    """src/gmmmap.jl:10-21 (field names transliterated)."""
    weights: np.ndarray
    mux: np.ndarray
    muy: np.ndarray
    Sxx: np.ndarray
    Sxy: np.ndarray
    Syx: np.ndarray
    Syy: np.ndarray
    SyxSxxinv: np.ndarray


class GMMMap(FrameByFrameConverter):
    """``GMMMap(weights, mu, Sigma; swap=false)``  (src/gmmmap.jl:57-91)."""

    def __init__(self, weights, means, covars, swap: bool = False):
        w, mu, sg = _f64(weights), _f64(means), _f64(covars)
        if mu.ndim != 2 or sg.ndim != 3 or w.ndim != 1:
            raise ArgumentError(_lib.EARG, "expected weights (M,), means (2D, M), covars (2D, 2D, M)")
        twoD, M = mu.shape
        if sg.shape != (twoD, twoD, M) or w.shape != (M,):
            raise DimensionMismatch(_lib.EDIM, "Inconsistent dimentions.")
        self._h = C.c_void_p()
        _lib.check(_lib.lib().vcb_gmmmap_create(_lib.ptr(w), _lib.ptr(mu), _lib.ptr(sg), twoD, M, int(bool(swap)),
                                                C.byref(self._h)))
        d, m = C.c_int32(), C.c_int32()
        _lib.check(_lib.lib().vcb_gmmmap_dim(self._h, C.byref(d)))
        _lib.check(_lib.lib().vcb_gmmmap_ncomponents(self._h, C.byref(m)))
        self._dim, self._M = d.value, m.value
        self._params = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().vcb_gmmmap_destroy(h)
            except Exception:
                pass
            self._h = None

    def __len__(self) -> int:          # Base.length(g::GMMMap) = 1   src/gmmmap.jl:93
        return 1

    @property
    def dim(self) -> int:              # src/gmmmap.jl:94
        return self._dim

    @property
    def ncomponents(self) -> int:      # src/gmmmap.jl:95
        return self._M

    @property
    def size(self) -> Tuple[int, int]:  # src/gmmmap.jl:96
        return (self.dim, len(self))

    @property
    def params(self) -> GMMMapParam:
        if self._params is None:
            D, M = self._dim, self._M
            shapes = [(D, M), (D, M), (D, D, M), (D, D, M), (D, D, M), (D, D, M), (D, D, M), (M,)]
            got = []
            for which, shp in enumerate(shapes):
                buf = np.empty(shp, order="F")
                _lib.check(_lib.lib().vcb_gmmmap_get_param(self._h, which, _lib.ptr(buf)))
                got.append(buf)
            self._params = GMMMapParam(got[7], got[0], got[1], got[3], got[4], got[5], got[6], got[2])
        return self._params


class TrajectoryGMMMap(TrajectoryConverter):
    """``TrajectoryGMMMap(g::GMMMap, T::Int)``  (src/trajectory_gmmmap.jl:3-37).

    ``len(t)`` is the number of frames W is currently built for; like the reference it changes
    whenever ``fvconvert`` sees a different length (src/trajectory_gmmmap.jl:70-72).
    """

    def __init__(self, g: GMMMap, T: int):
        if not isinstance(g, GMMMap):
            raise ArgumentError(_lib.EARG, "TrajectoryGMMMap needs a GMMMap")
        if int(T) < 1:
            raise ArgumentError(_lib.EARG, "T must be positive")
        self.gmmmap = g
        self._h = C.c_void_p()
        _lib.check(_lib.lib().vcb_traj_create(g._h, C.byref(self._h)))
        self._T = int(T)
        self.Ey = np.zeros(0)          # tgmm.Ey, kept for the GV variant (:90-91)
        self._Dy = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().vcb_traj_destroy(h)
            except Exception:
                pass
            self._h = None

    def __len__(self) -> int:          # div(size(W,2), div(dim,2))   :34
        return self._T

    @property
    def dim(self) -> int:              # :35
        return self.gmmmap.dim

    @property
    def ncomponents(self) -> int:      # :36
        return self.gmmmap.ncomponents

    @property
    def size(self) -> Tuple[int, int]:  # :37
        return (self.dim, len(self))

    @property
    def Dy(self) -> np.ndarray:        # :24-28
        if self._Dy is None:
            d, M = self.dim, self.ncomponents
            buf = np.empty((d, d, M), order="F")
            _lib.check(_lib.lib().vcb_traj_get_Dy(self._h, _lib.ptr(buf)))
            self._Dy = buf
        return self._Dy


class TrajectoryGVGMMMap(TrajectoryConverter):
    """``TrajectoryGVGMMMap(tgmm, mu_v, S_vv)``  (src/trajectory_gmmmap.jl:114-133): trajectory
    conversion followed by gradient ascent on the likelihood with the global-variance term."""

    def __init__(self, tgmm: TrajectoryGMMMap, mu_v, sigma_vv):
        if not isinstance(tgmm, TrajectoryGMMMap):
            raise ArgumentError(_lib.EARG, "TrajectoryGVGMMMap needs a TrajectoryGMMMap")
        self.tgmm = tgmm
        self.mu_v = _f64(mu_v)
        self.sigma_vv = _f64(sigma_vv)
        Ds = tgmm.dim // 2
        if self.mu_v.shape != (Ds,) or self.sigma_vv.shape != (Ds, Ds):
            raise DimensionMismatch(_lib.EDIM, "GV statistics must be (Ds,) and (Ds, Ds)")
        self._h = C.c_void_p()
        _lib.check(_lib.lib().vcb_trajgv_create(tgmm._h, _lib.ptr(self.mu_v), _lib.ptr(self.sigma_vv), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().vcb_trajgv_destroy(h)
            except Exception:
                pass
            self._h = None

    def __len__(self) -> int:           # :129
        return len(self.tgmm)

    @property
    def dim(self) -> int:               # :130
        return self.tgmm.dim

    @property
    def ncomponents(self) -> int:       # :131
        return self.tgmm.ncomponents

    @property
    def size(self):                     # :132 refers to an undefined variable `t` and throws in the reference
        raise NameError("t not defined (Base.size(g::TrajectoryGVGMMMap), src/trajectory_gmmmap.jl:132)")


class VarianceScaling(NamedTuple):
    """``VarianceScaling(sigma2)``  (src/gv.jl:6-8)."""
    sigma2: np.ndarray


class _PinnedBuffer:
    """Owner of one vcb_host_alloc allocation (freed when the last array view dies)."""

    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        _lib.check(_lib.lib().vcb_host_alloc(C.byref(self.ptr), max(int(nbytes), 1)))
        self.nbytes = int(nbytes)

    def __del__(self):
        p = getattr(self, "ptr", None)
        if p:
            try:
                _lib.lib().vcb_host_free(p)
            except Exception:
                pass
            self.ptr = None


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """Column-major array in page-locked host memory (``vcb_host_alloc``): host<->device copies of such
    arrays run at PCIe speed and keep the library's copy/compute pipelines asynchronous; ordinary
    (pageable) arrays are staged by the driver at a fraction of that."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape))
    owner = _PinnedBuffer(n * dt.itemsize)
    buf = (C.c_char * (n * dt.itemsize)).from_address(owner.ptr.value)
    a = np.frombuffer(buf, dtype=dt, count=n).reshape(shape, order="F")
    a.flags.writeable = True
    _PINNED_OWNERS[id(buf)] = owner           # keep the allocation alive as long as the ctypes buffer is
    import weakref
    weakref.finalize(buf, _PINNED_OWNERS.pop, id(buf), None)
    return a


_PINNED_OWNERS: dict = {}


def traj_status(t: "TrajectoryGMMMap") -> None:
    """After device-resident (``*_dev``) conversions: synchronises the current stream and raises
    ``PosDefException`` if any of them met a non-positive pivot of ``W' D^-1 W`` (an indefinite
    ``Dy``) since the last query.  Host-array conversions raise by themselves."""
    st = None
    try:
        import torch
        if torch.cuda.is_available():
            st = _stream_ptr()
    except ImportError:
        pass
    _lib.check(_lib.lib().vcb_traj_status(t._h, st))


def _check_offsets(off: np.ndarray, total: int) -> np.ndarray:
    """offsets (n+1) must start at >= 0, be non-decreasing and stay inside the array."""
    if off.ndim != 1 or len(off) < 1:
        raise ArgumentError(_lib.EARG, "offsets must be a vector of n+1 frame offsets")
    if len(off) > 1 and (off[0] < 0 or np.any(np.diff(off) < 0) or off[-1] > total):
        raise ArgumentError(_lib.EARG, "offsets must be non-decreasing and within the %d frames of the array" % total)
    return off


def fvpostf(vs: VarianceScaling, src, offsets=None):
    """``fvpostf(vs, src)`` (src/gv.jl:17-21): per-dimension variance scaling of one (D, T) matrix,
    of a concatenated batch with ``offsets`` (one filter per utterance), or of a frame-major CUDA
    tensor.  Returns a filtered copy."""
    L = _lib.lib()
    s2 = _f64(vs.sigma2)
    if _is_torch(src):
        import torch
        _check_dev_tensor(src)
        T, D = src.shape
        if s2.shape != (D,):
            raise DimensionMismatch(_lib.EDIM, "VarianceScaling has %d entries, src has %d rows" % (s2.size, D))
        off = np.array([0, T], dtype=np.int64) if offsets is None else np.ascontiguousarray(offsets, dtype=np.int64)
        _check_offsets(off, T)
        out = torch.empty_like(src)
        d_s2 = torch.from_numpy(s2).to(src.device)
        _lib.check(L.vcb_variance_scaling_batch_dev(_lib.ptr(d_s2), D, _lib.ptr(src), D, _lib.ptr(off), len(off) - 1,
                                                    _lib.ptr(out), D, _stream_ptr()))
        return out
    src = _f64(src)
    D, T = src.shape
    if s2.shape != (D,):
        raise DimensionMismatch(_lib.EDIM, "VarianceScaling has %d entries, src has %d rows" % (s2.size, D))
    off = np.array([0, T], dtype=np.int64) if offsets is None else np.ascontiguousarray(offsets, dtype=np.int64)
    _check_offsets(off, T)
    out = np.empty_like(src, order="F")
    _lib.check(L.vcb_variance_scaling_batch(_lib.ptr(s2), D, _lib.ptr(src), D, _lib.ptr(off), len(off) - 1, _lib.ptr(out), D))
    return out


def fvpostf_(vs: VarianceScaling, src: np.ndarray, offsets=None) -> np.ndarray:
    """``fvpostf!(vs, src)`` (src/gv.jl:10-15): in place on a column-major float64 array."""
    src[...] = fvpostf(vs, src, offsets)
    return src


def diffgmm(params):
    """``diffgmm(params::GMMMapParam)`` (src/diffgmm.jl:9-25).  Accepts a GMMMapParam (returns the
    differential GMMMapParam, A recomputed as in the 7-argument constructor src/gmmmap.jl:23-38)
    or a ``(weights, means, covars)`` joint triple (returns the joint triple of the differential
    model, ready for ``GMMMap(...)``)."""
    L = _lib.lib()
    if isinstance(params, GMMMapParam):
        D, M = params.mux.shape
        mu = np.asfortranarray(np.concatenate([params.mux, params.muy], axis=0))
        sg = np.empty((2 * D, 2 * D, M), order="F")
        sg[:D, :D], sg[:D, D:], sg[D:, :D], sg[D:, D:] = params.Sxx, params.Sxy, params.Syx, params.Syy
        w, mo, so = diffgmm((params.weights, mu, sg))
        return GMMMap(w, mo, so).params
    w, mu, sg = params[0], _f64(params[1]), _f64(params[2])
    mo, so = np.empty_like(mu, order="F"), np.empty_like(sg, order="F")
    _lib.check(L.vcb_diffgmm(_lib.ptr(mu), _lib.ptr(sg), mu.shape[0], mu.shape[1], _lib.ptr(mo), _lib.ptr(so)))
    return type(params)(w, mo, so) if hasattr(params, "_fields") else (w, mo, so)


def dim(c) -> int:
    return c.dim


def ncomponents(c) -> int:
    return c.ncomponents


# ---- posterior helpers (src/gmm.jl:24-58) -------------------------------------------------------
def predict_proba(g: GMMMap, X) -> np.ndarray:
    """``predict_proba(g.px, X)``: X (D,) or (D, T) -> (M,) or (M, T)."""
    X = _f64(X)
    vec = X.ndim == 1
    if vec:
        X = X.reshape(-1, 1, order="F")
    post = np.empty((g.ncomponents, X.shape[1]), order="F")
    _lib.check(_lib.lib().vcb_gmmmap_predict_proba(g._h, _lib.ptr(X), X.shape[0], X.shape[1], X.shape[0], _lib.ptr(post)))
    return post[:, 0].copy() if vec else post


def predict(g: GMMMap, X):
    """``predict(g.px, X)``: 1-based index of the most likely mixture per frame."""
    X = _f64(X)
    vec = X.ndim == 1
    if vec:
        X = X.reshape(-1, 1, order="F")
    out = np.empty(X.shape[1], dtype=np.int64)
    _lib.check(_lib.lib().vcb_gmmmap_predict(g._h, _lib.ptr(X), X.shape[0], X.shape[1], X.shape[0], _lib.ptr(out)))
    return int(out[0]) if vec else out


# ---- fvconvert ----------------------------------------------------------------------------------
def fvconvert(c, X, return_aux: bool = False):
    """``fvconvert(g::GMMMap, x::Vector)`` (src/gmmmap.jl:101-118) -- also accepts a (D, T) matrix
    or a frame-major CUDA tensor -- and ``fvconvert(t::TrajectoryGMMMap, X::Matrix)``
    (src/trajectory_gmmmap.jl:65-110)."""
    L = _lib.lib()
    if isinstance(c, GMMMap):
        if _is_torch(X):
            import torch
            _check_dev_tensor(X)
            T, rows = X.shape
            Y = torch.empty((T, c.dim), dtype=torch.float64, device=X.device)
            _lib.check(L.vcb_gmmmap_convert_dev(c._h, _lib.ptr(X), rows, T, rows, _lib.ptr(Y), c.dim, _stream_ptr()))
            return Y
        X = _f64(X)
        vec = X.ndim == 1
        if vec:
            X = X.reshape(-1, 1, order="F")
        rows, T = X.shape
        Y = np.empty((c.dim, T), order="F")
        _lib.check(L.vcb_gmmmap_convert(c._h, _lib.ptr(X), rows, T, rows, _lib.ptr(Y), c.dim))
        return Y[:, 0].copy() if vec else Y
    if isinstance(c, TrajectoryGMMMap):
        Ds = c.dim // 2
        if _is_torch(X):
            import torch
            _check_dev_tensor(X)
            T, rows = X.shape
            off = np.array([0, T], dtype=np.int64)
            Y = torch.empty((T, Ds), dtype=torch.float64, device=X.device)
            _lib.check(L.vcb_traj_convert_batch_dev(c._h, _lib.ptr(X), rows, rows, _lib.ptr(off), 1, 0, _lib.ptr(Y), Ds,
                                                    None, None, _stream_ptr()))
            c._T = T
            return Y
        X = _f64(X)
        if X.ndim != 2:
            raise ArgumentError(_lib.EARG, "fvconvert(::TrajectoryGMMMap, X) needs a matrix")
        rows, T = X.shape
        off = np.array([0, T], dtype=np.int64)
        Y = np.empty((Ds, T), order="F")
        mhat = np.empty(T, dtype=np.int64)
        Ey = np.empty((rows, T), order="F")
        _lib.check(L.vcb_traj_convert_batch(c._h, _lib.ptr(X), rows, rows, _lib.ptr(off), 1, 0, _lib.ptr(Y), Ds,
                                            _lib.ptr(mhat), _lib.ptr(Ey)))
        c._T = T                                   # W rebuilt for the new length (:70-72)
        c.Ey = Ey.reshape(-1, order="F")           # tgmm.Ey = vec(Ey)   (:90-91)
        return (Y, mhat, Ey) if return_aux else Y
    if isinstance(c, TrajectoryGVGMMMap):
        return _fvconvert_gv(c, X)
    raise TypeError("fvconvert: unsupported converter type")


def _fvconvert_gv(c: "TrajectoryGVGMMMap", X, epochs: int = 100, alpha: float = 1.0e-5):
    L = _lib.lib()
    Ds = c.dim // 2
    if _is_torch(X):
        import torch
        _check_dev_tensor(X)
        T, rows = X.shape
        off = np.array([0, T], dtype=np.int64)
        Y = torch.empty((T, Ds), dtype=torch.float64, device=X.device)
        _lib.check(L.vcb_trajgv_convert_batch_dev(c._h, _lib.ptr(X), rows, rows, _lib.ptr(off), 1, 0, int(epochs),
                                                  float(alpha), _lib.ptr(Y), Ds, _stream_ptr()))
        c.tgmm._T = T
        return Y
    X = _f64(X)
    if X.ndim != 2:
        raise ArgumentError(_lib.EARG, "fvconvert(::TrajectoryGVGMMMap, X) needs a matrix")
    rows, T = X.shape
    off = np.array([0, T], dtype=np.int64)
    Y = np.empty((Ds, T), order="F")
    _lib.check(L.vcb_trajgv_convert_batch(c._h, _lib.ptr(X), rows, rows, _lib.ptr(off), 1, 0, int(epochs), float(alpha),
                                          _lib.ptr(Y), Ds))
    c.tgmm._T = T
    return Y


def fvconvert_gv(c: "TrajectoryGVGMMMap", X, epochs: int = 100, alpha: float = 1.0e-5):
    """``fvconvert(tgv, X; epochs=100, α=1.0e-5)`` (src/trajectory_gmmmap.jl:140-172) with its
    keyword arguments."""
    return _fvconvert_gv(c, X, epochs, alpha)


# ---- vc -----------------------------------------------------------------------------------------
def vc(c, fm, out=None):
    """``vc(c, fm)`` (src/common.jl:7-26 and :31-63).  Row 1 of ``fm`` (power) passes through.
    ``out`` (frame-by-frame host path only) is an optional preallocated column-major result
    buffer, e.g. page-locked memory."""
    L = _lib.lib()
    if isinstance(c, FrameByFrameConverter):
        if _is_torch(fm):
            import torch
            _check_dev_tensor(fm)
            T, rows = fm.shape
            out = torch.empty_like(fm)
            _lib.check(L.vcb_gmmmap_vc_dev(c._h, _lib.ptr(fm), rows, T, _lib.ptr(out), _stream_ptr()))
            return out
        fm = _f64(fm)
        if fm.ndim != 2:
            raise ArgumentError(_lib.EARG, "vc needs a feature matrix")
        rows, T = fm.shape
        if out is None:
            out = np.empty_like(fm, order="F")
        elif out.shape != fm.shape or out.dtype != np.float64 or not out.flags.f_contiguous:
            raise ArgumentError(_lib.EARG, "out must be a column-major float64 array shaped like fm")
        _lib.check(L.vcb_gmmmap_vc(c._h, _lib.ptr(fm), rows, T, _lib.ptr(out)))
        return out
    if isinstance(c, TrajectoryConverter):
        if _is_torch(fm):
            T = fm.shape[0]
            out, = vc_batch(c, fm, np.array([0, T], dtype=np.int64), _split=False)
            return out
        fm = _f64(fm)
        out = vc_batch(c, [fm])[0]
        return out
    raise TypeError("vc: unsupported converter type")


def vc_batch(c, fms, offsets=None, _split: bool = True, epochs: int = 100, alpha: float = 1.0e-5, out=None):
    """Batch extension of ``vc(c::TrajectoryConverter, fm)``: every utterance is converted as
    ``vc(c, fm_s)`` would with the chunk limit ``len(c)`` read once (src/common.jl:43), all in one
    library call.  ``fms`` is a list of (1+2Ds, T_s) matrices, or one concatenated matrix /
    frame-major CUDA tensor together with ``offsets`` (n+1, in frames).

    Like the reference, the converter remembers the length of the last chunk it solved
    (src/trajectory_gmmmap.jl:70-72), so ``len(c)`` may change.  ``out`` (host path with
    ``offsets``) is an optional preallocated column-major (1+Ds, total) result buffer, e.g.
    page-locked memory.
    """
    L = _lib.lib()
    limit = len(c)
    Ds = c.dim // 2
    gv = isinstance(c, TrajectoryGVGMMMap)      # vc passes no keywords: epochs = 100, alpha = 1e-5
    if gv:
        c, cgv = c.tgmm, c
    if _is_torch(fms):
        import torch
        _check_dev_tensor(fms)
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        total, rows = fms.shape
        _check_offsets(off, total)
        out = torch.empty((total, Ds + 1), dtype=torch.float64, device=fms.device)
        if gv:
            _lib.check(L.vcb_trajgv_vc_batch_dev(cgv._h, _lib.ptr(fms), rows, _lib.ptr(off), len(off) - 1, limit,
                                                 int(epochs), float(alpha), _lib.ptr(out), _stream_ptr()))
        else:
            _lib.check(L.vcb_traj_vc_batch_dev(c._h, _lib.ptr(fms), rows, _lib.ptr(off), len(off) - 1, limit,
                                               _lib.ptr(out), _stream_ptr()))
        _update_len(c, off, limit)
        return (out,) if not _split else [out[off[i]:off[i + 1]] for i in range(len(off) - 1)]
    if offsets is None:
        mats = [_f64(m) for m in fms]
        rows = mats[0].shape[0]
        for m in mats:
            if m.ndim != 2 or m.shape[0] != rows:
                raise DimensionMismatch(_lib.EDIM, "Inconsistent dimentions.")
        off = np.concatenate([[0], np.cumsum([m.shape[1] for m in mats])]).astype(np.int64)
        fm = np.asfortranarray(np.concatenate(mats, axis=1)) if len(mats) > 1 else mats[0]
    else:
        fm = _f64(fms)
        rows = fm.shape[0]
        off = _check_offsets(np.ascontiguousarray(offsets, dtype=np.int64), fm.shape[1])
    if out is None:
        out = np.empty((Ds + 1, fm.shape[1]), order="F")
    elif out.shape != (Ds + 1, fm.shape[1]) or out.dtype != np.float64 or not out.flags.f_contiguous:
        raise ArgumentError(_lib.EARG, "out must be a column-major float64 array of shape (1 + dim/2, total frames)")
    if gv:
        _lib.check(L.vcb_trajgv_vc_batch(cgv._h, _lib.ptr(fm), rows, _lib.ptr(off), len(off) - 1, limit, int(epochs),
                                         float(alpha), _lib.ptr(out)))
    else:
        _lib.check(L.vcb_traj_vc_batch(c._h, _lib.ptr(fm), rows, _lib.ptr(off), len(off) - 1, limit, _lib.ptr(out)))
    _update_len(c, off, limit)
    if offsets is not None and not _split:
        return (out,)
    return [np.asfortranarray(out[:, off[i]:off[i + 1]]) for i in range(len(off) - 1)]


def vc_static_batch(c: TrajectoryGMMMap, fm, offsets):
    """``vc(c, [fm[1,:]; push_delta(fm[2:end,:])])`` for a batch in one call (the pattern of
    bin/vc.jl:76-82): ``fm`` is (1+Ds, total) -- power row and STATIC features -- column-major, or a
    frame-major CUDA tensor (total, 1+Ds); the delta rows are appended on the device."""
    L = _lib.lib()
    limit = len(c)
    Ds = c.dim // 2
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    if _is_torch(fm):
        import torch
        _check_dev_tensor(fm)
        total, rows = fm.shape
        _check_offsets(off, total)
        out = torch.empty((total, Ds + 1), dtype=torch.float64, device=fm.device)
        _lib.check(L.vcb_traj_vc_static_batch_dev(c._h, _lib.ptr(fm), rows, _lib.ptr(off), len(off) - 1, limit,
                                                  _lib.ptr(out), _stream_ptr()))
    else:
        fm = _f64(fm)
        _check_offsets(off, fm.shape[1])
        out = np.empty((Ds + 1, fm.shape[1]), order="F")
        _lib.check(L.vcb_traj_vc_static_batch(c._h, _lib.ptr(fm), fm.shape[0], _lib.ptr(off), len(off) - 1, limit,
                                              _lib.ptr(out)))
    _update_len(c, off, limit)
    return out


def _update_len(c: TrajectoryGMMMap, off: np.ndarray, limit: int) -> None:
    # length of the last chunk of the last non-empty utterance (quirk Q3)
    for i in range(len(off) - 2, -1, -1):
        T = int(off[i + 1] - off[i])
        if T > 0:
            last = T % limit if limit > 0 else T
            c._T = last if last else min(limit, T)
            return


# ---- W (never used by the library; provided because the reference exposes and tests it) ---------
def constructW(D: int, T: int):
    """``constructW(D, T)`` (src/trajectory_gmmmap.jl:39-61) as a scipy.sparse CSC matrix."""
    import scipy.sparse as sp
    rows: List[int] = []
    cols: List[int] = []
    vals: List[float] = []
    eye = np.arange(D)
    for t in range(T):
        r0 = 2 * D * t
        rows.extend(r0 + eye); cols.extend(t * D + eye); vals.extend([1.0] * D)
        if t >= 1:
            rows.extend(r0 + D + eye); cols.extend((t - 1) * D + eye); vals.extend([-0.5] * D)
        if t < T - 1:
            rows.extend(r0 + D + eye); cols.extend((t + 1) * D + eye); vals.extend([0.5] * D)
    return sp.csc_matrix((vals, (rows, cols)), shape=(2 * D * T, D * T))


# ---- callers either side of the path ------------------------------------------------
def push_delta(src, offsets=None) -> np.ndarray:
    """``push_delta(src)`` (src/datasets.jl:6-13); with ``offsets`` a ragged batch in one call."""
    s = _f64(src)
    off = np.array([0, s.shape[1]], dtype=np.int64) if offsets is None else np.ascontiguousarray(offsets, dtype=np.int64)
    _check_offsets(off, s.shape[1])
    out = np.empty((2 * s.shape[0], s.shape[1]), order="F")
    _lib.check(_lib.lib().vcb_push_delta_batch(_lib.ptr(s), s.shape[0], _lib.ptr(off), len(off) - 1, _lib.ptr(out)))
    return out


def align_batch(src, src_off, tgt, tgt_off):
    """Batch of ``align(src, tgt)`` (src/align.jl:8-35).  Returns (newtgt, paths)."""
    s, t = _f64(src), _f64(tgt)
    if s.shape[0] != t.shape[0]:
        raise DimensionMismatch(_lib.EDIM, "order of feature vector must be equal")
    so = _check_offsets(np.ascontiguousarray(src_off, dtype=np.int64), s.shape[1])
    to = _check_offsets(np.ascontiguousarray(tgt_off, dtype=np.int64), t.shape[1])
    if len(so) != len(to):
        raise ArgumentError(_lib.EARG, "src_off and tgt_off must describe the same number of pairs")
    newtgt = np.empty_like(s, order="F")
    paths = np.empty(t.shape[1], dtype=np.int64)
    _lib.check(_lib.lib().vcb_align_batch(_lib.ptr(s), _lib.ptr(so), _lib.ptr(t), _lib.ptr(to), len(so) - 1, s.shape[0],
                                          _lib.ptr(newtgt), _lib.ptr(paths)))
    return newtgt, paths


def align(src, tgt):
    """``align(src, tgt)`` -> (src, newtgt)  (src/align.jl:8-35)."""
    s, t = _f64(src), _f64(tgt)
    newtgt, _ = align_batch(s, [0, s.shape[1]], t, [0, t.shape[1]])
    return s, newtgt
