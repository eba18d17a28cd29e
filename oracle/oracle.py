"""ctypes binding of the CPU oracle (oracle/libvc_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package.

Arrays use the reference's Julia shapes -- (D, T) feature matrices, (2D, M) means,
(2D, 2D, M) covariances -- stored column-major (``order='F'``), which is what the C code reads.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvc_oracle.so")

OK, EDIM, ENOTPD, ESINGULAR, EARG, ENOMEM = range(6)


class OracleError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"oracle {what} failed with code {code}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("vc_oracle.c", "vc_oracle.h"))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_m:
        subprocess.run(["make", "-C", _HERE, "libvc_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.vco_gmmmap_create.argtypes = [_dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.vco_gmmmap_destroy.argtypes = [C.c_void_p]
        L.vco_gmmmap_destroy.restype = None
        for n in ("vco_gmmmap_dim", "vco_gmmmap_ncomponents", "vco_gmmmap_length"):
            getattr(L, n).argtypes = [C.c_void_p]
        L.vco_gmmmap_param.argtypes = [C.c_void_p, C.c_int]
        L.vco_gmmmap_param.restype = _dp
        L.vco_predict_proba.argtypes = [C.c_void_p, _dp, _dp]
        L.vco_predict.argtypes = [C.c_void_p, _dp, _ip]
        L.vco_fvconvert.argtypes = [C.c_void_p, _dp, C.c_int, _dp]
        L.vco_vc_fbf.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int64, _dp]
        L.vco_vc_fbf_mt.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int64, _dp, C.c_int]
        L.vco_traj_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.vco_traj_destroy.argtypes = [C.c_void_p]
        L.vco_traj_destroy.restype = None
        L.vco_traj_length.argtypes = [C.c_void_p]
        L.vco_traj_dim.argtypes = [C.c_void_p]
        L.vco_traj_Dy.argtypes = [C.c_void_p]
        L.vco_traj_Dy.restype = _dp
        L.vco_constructW.argtypes = [C.c_int, C.c_int, _lp, _lp, _dp]
        L.vco_constructW.restype = C.c_int64
        L.vco_traj_fvconvert.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, _dp, _ip, _dp]
        L.vco_vc_traj.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int64, _dp]
        L.vco_vc_traj_batch_mt.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, _lp, C.c_int64, _dp, C.c_int]
        L.vco_variance_scaling.argtypes = [_dp, _dp, C.c_int, C.c_int64]
        L.vco_variance_scaling.restype = None
        L.vco_trajgv_create.argtypes = [C.c_void_p, _dp, _dp, C.POINTER(C.c_void_p)]
        L.vco_trajgv_destroy.argtypes = [C.c_void_p]
        L.vco_trajgv_destroy.restype = None
        L.vco_trajgv_fvconvert.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_int, C.c_double, _dp]
        L.vco_vc_trajgv.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int64, C.c_int, C.c_double, _dp]
        L.vco_diffgmm.argtypes = [_dp, _dp, C.c_int, C.c_int, _dp, _dp]
        L.vco_diffgmm.restype = None
        L.vco_dtw_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.vco_dtw_destroy.argtypes = [C.c_void_p]
        L.vco_dtw_destroy.restype = None
        L.vco_dtw_fit.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, _dp, C.c_int, _lp]
        L.vco_dtw_set_template.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int]
        L.vco_dtw_update.argtypes = [C.c_void_p, _dp, C.c_int]
        L.vco_dtw_backward.argtypes = [C.c_void_p, _lp]
        L.vco_dtw_tables.argtypes = [C.c_void_p, _ip, _ip, C.POINTER(_dp), C.POINTER(_lp)]
        L.vco_dtw_fit_batch_mt.argtypes = [_dp, _lp, _dp, _lp, C.c_int64, C.c_int, C.c_int, C.c_int, _lp, _dp, C.c_int]
        L.vco_push_delta.argtypes = [_dp, C.c_int, C.c_int, _dp]
        L.vco_push_delta.restype = None
        L.vco_align.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_int, _dp, _lp]
        L.vco_max_threads.argtypes = []
        _lib = L
    return _lib


def _f64(a) -> np.ndarray:
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def _check(rc: int, what: str):
    if rc != OK:
        raise OracleError(rc, what)


def max_threads() -> int:
    return int(lib().vco_max_threads())


class GMMMap:
    """src/gmmmap.jl:57-96"""

    def __init__(self, weights, means, covars, swap: bool = False):
        w, mu, sg = _f64(weights), _f64(means), _f64(covars)
        twoD, M = mu.shape
        assert sg.shape == (twoD, twoD, M) and w.shape == (M,)
        self._h = C.c_void_p()
        _check(lib().vco_gmmmap_create(_p(w), _p(mu), _p(sg), twoD, M, int(swap), C.byref(self._h)), "GMMMap")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vco_gmmmap_destroy(self._h)
            self._h = None

    def __len__(self):
        return lib().vco_gmmmap_length(self._h)

    @property
    def dim(self):
        return lib().vco_gmmmap_dim(self._h)

    @property
    def ncomponents(self):
        return lib().vco_gmmmap_ncomponents(self._h)

    @property
    def size(self):
        return (self.dim, len(self))

    def param(self, which: int, shape):
        n = int(np.prod(shape))
        buf = np.ctypeslib.as_array(lib().vco_gmmmap_param(self._h, which), shape=(n,))
        return buf.copy().reshape(shape, order="F")

    def predict_proba(self, x):
        x = _f64(x)
        post = np.empty(self.ncomponents)
        _check(lib().vco_predict_proba(self._h, _p(x), _p(post)), "predict_proba")
        return post

    def predict(self, X):
        X = _f64(X)
        if X.ndim == 1:
            X = X.reshape(-1, 1, order="F")
        out = np.empty(X.shape[1], dtype=np.int64)
        m = C.c_int()
        for t in range(X.shape[1]):
            col = np.ascontiguousarray(X[:, t])
            _check(lib().vco_predict(self._h, _p(col), C.byref(m)), "predict")
            out[t] = m.value
        return out

    def fvconvert(self, x):
        x = _f64(x)
        y = np.empty(self.dim)
        _check(lib().vco_fvconvert(self._h, _p(x), x.shape[0], _p(y)), "fvconvert")
        return y

    def vc(self, fm, nthreads: int = 0):
        fm = _f64(fm)
        out = np.empty_like(fm, order="F")
        if nthreads > 0:
            rc = lib().vco_vc_fbf_mt(self._h, _p(fm), fm.shape[0], fm.shape[1], _p(out), nthreads)
        else:
            rc = lib().vco_vc_fbf(self._h, _p(fm), fm.shape[0], fm.shape[1], _p(out))
        _check(rc, "vc(fbf)")
        return out


class TrajectoryGMMMap:
    """src/trajectory_gmmmap.jl:3-37"""

    def __init__(self, g: GMMMap, T: int):
        self.gmmmap = g
        self._h = C.c_void_p()
        _check(lib().vco_traj_create(g._h, T, C.byref(self._h)), "TrajectoryGMMMap")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vco_traj_destroy(self._h)
            self._h = None

    def __len__(self):
        return lib().vco_traj_length(self._h)

    @property
    def dim(self):
        return lib().vco_traj_dim(self._h)

    @property
    def ncomponents(self):
        return self.gmmmap.ncomponents

    @property
    def size(self):
        return (self.dim, len(self))

    @property
    def Dy(self):
        d, M = self.dim, self.ncomponents
        return np.ctypeslib.as_array(lib().vco_traj_Dy(self._h), shape=(d * d * M,)).copy().reshape((d, d, M), order="F")

    def fvconvert(self, X, return_aux: bool = False):
        X = _f64(X)
        rows, T = X.shape
        Y = np.empty((rows // 2, T), order="F")
        mhat = np.empty(T, dtype=np.int32)
        Ey = np.empty((rows, T), order="F")
        _check(lib().vco_traj_fvconvert(self._h, _p(X), rows, T, _p(Y), mhat.ctypes.data_as(_ip), _p(Ey)), "fvconvert(traj)")
        return (Y, mhat.astype(np.int64), Ey) if return_aux else Y

    def vc(self, fm):
        fm = _f64(fm)
        rows, T = fm.shape
        out = np.empty((((rows - 1) >> 1) + 1, T), order="F")
        _check(lib().vco_vc_traj(self._h, _p(fm), rows, T, _p(out)), "vc(traj)")
        return out


class TrajectoryGVGMMMap:
    """src/trajectory_gmmmap.jl:112-189"""

    def __init__(self, tgmm: TrajectoryGMMMap, mu_v, sigma_vv):
        self.tgmm = tgmm
        mu_v, sigma_vv = _f64(mu_v), _f64(sigma_vv)
        self._h = C.c_void_p()
        _check(lib().vco_trajgv_create(tgmm._h, _p(mu_v), _p(sigma_vv), C.byref(self._h)), "TrajectoryGVGMMMap")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vco_trajgv_destroy(self._h)
            self._h = None

    def __len__(self):
        return len(self.tgmm)

    def fvconvert(self, X, epochs: int = 100, alpha: float = 1.0e-5):
        X = _f64(X)
        rows, T = X.shape
        Y = np.empty((rows // 2, T), order="F")
        _check(lib().vco_trajgv_fvconvert(self._h, _p(X), rows, T, epochs, alpha, _p(Y)), "fvconvert(trajgv)")
        return Y

    def vc(self, fm, epochs: int = 100, alpha: float = 1.0e-5):
        fm = _f64(fm)
        rows, T = fm.shape
        out = np.empty((((rows - 1) >> 1) + 1, T), order="F")
        _check(lib().vco_vc_trajgv(self._h, _p(fm), rows, T, epochs, alpha, _p(out)), "vc(trajgv)")
        return out


def fvpostf(sigma2, src):
    """fvpostf(VarianceScaling(sigma2), src)  src/gv.jl:10-21 (returns a filtered copy)."""
    out = _f64(src).copy(order="F")
    s2 = _f64(sigma2)
    lib().vco_variance_scaling(_p(s2), _p(out), out.shape[0], out.shape[1])
    return out


def diffgmm(means, covars):
    """src/diffgmm.jl:9-25 applied to the joint parameters; returns (means', covars')."""
    mu, sg = _f64(means), _f64(covars)
    mo, so = np.empty_like(mu, order="F"), np.empty_like(sg, order="F")
    lib().vco_diffgmm(_p(mu), _p(sg), mu.shape[0], mu.shape[1], _p(mo), _p(so))
    return mo, so


def vc_traj_batch(g: GMMMap, limit: int, fm, offsets, nthreads: int = 1):
    """Independent `vc(TrajectoryGMMMap(g, limit), fm_s)` per utterance s; fm is the column-wise
    concatenation, offsets (nseq+1) in frames."""
    fm = _f64(fm)
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    rows = fm.shape[0]
    out = np.empty((((rows - 1) >> 1) + 1, fm.shape[1]), order="F")
    _check(lib().vco_vc_traj_batch_mt(g._h, limit, _p(fm), rows, off.ctypes.data_as(_lp), len(off) - 1, _p(out), nthreads), "vc_traj_batch")
    return out


def constructW(D: int, T: int):
    """src/trajectory_gmmmap.jl:55-61 as a dense (2DT, DT) array (small sizes only)."""
    nnz = lib().vco_constructW(D, T, None, None, None)
    r = np.empty(nnz, dtype=np.int64)
    c = np.empty(nnz, dtype=np.int64)
    v = np.empty(nnz)
    lib().vco_constructW(D, T, r.ctypes.data_as(_lp), c.ctypes.data_as(_lp), _p(v))
    return r, c, v


class DTW:
    """src/dtw.jl:11-21"""

    def __init__(self, fstep: int = 0, bstep: int = 1):
        self.fstep, self.bstep = fstep, bstep
        self._h = C.c_void_p()
        _check(lib().vco_dtw_create(fstep, bstep, C.byref(self._h)), "DTW")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vco_dtw_destroy(self._h)
            self._h = None

    def fit(self, template, sequence):
        tm, sq = _f64(template), _f64(sequence)
        path = np.empty(sq.shape[1], dtype=np.int64)
        _check(lib().vco_dtw_fit(self._h, _p(tm), tm.shape[0], tm.shape[1], _p(sq), sq.shape[1], path.ctypes.data_as(_lp)), "fit!")
        return path

    def set_template(self, template):
        tm = _f64(template)
        _check(lib().vco_dtw_set_template(self._h, _p(tm), tm.shape[0], tm.shape[1]), "set_template!")

    def update(self, v):
        v = _f64(v)
        _check(lib().vco_dtw_update(self._h, _p(v), v.shape[0]), "update!")

    def backward(self):
        S, n, _, _ = self._tables_raw()
        path = np.empty(n - 1, dtype=np.int64)
        _check(lib().vco_dtw_backward(self._h, path.ctypes.data_as(_lp)), "backward")
        return path

    def _tables_raw(self):
        S, n = C.c_int(), C.c_int()
        cp, bp = _dp(), _lp()
        lib().vco_dtw_tables(self._h, C.byref(S), C.byref(n), C.byref(cp), C.byref(bp))
        return S.value, n.value, cp, bp

    def tables(self):
        S, n, cp, bp = self._tables_raw()
        cost = np.ctypeslib.as_array(cp, shape=(S * n,)).copy().reshape((S, n), order="F")
        back = np.ctypeslib.as_array(bp, shape=(S * n,)).copy().reshape((S, n), order="F")
        return cost, back


def dtw_fit_batch(tmpl, tmpl_off, seq, seq_off, fstep=0, bstep=1, nthreads=1):
    tm, sq = _f64(tmpl), _f64(seq)
    to = np.ascontiguousarray(tmpl_off, dtype=np.int64)
    so = np.ascontiguousarray(seq_off, dtype=np.int64)
    n = len(to) - 1
    paths = np.empty(sq.shape[1], dtype=np.int64)
    fc = np.empty(n)
    _check(lib().vco_dtw_fit_batch_mt(_p(tm), to.ctypes.data_as(_lp), _p(sq), so.ctypes.data_as(_lp), n, tm.shape[0], fstep, bstep, paths.ctypes.data_as(_lp), _p(fc), nthreads), "dtw_fit_batch")
    return paths, fc


def push_delta(src):
    s = _f64(src)
    out = np.empty((2 * s.shape[0], s.shape[1]), order="F")
    lib().vco_push_delta(_p(s), s.shape[0], s.shape[1], _p(out))
    return out


def align(src, tgt):
    s, t = _f64(src), _f64(tgt)
    if s.shape[0] != t.shape[0]:
        raise OracleError(EDIM, "align")
    newtgt = np.empty_like(s, order="F")
    path = np.empty(t.shape[1], dtype=np.int64)
    _check(lib().vco_align(_p(s), s.shape[0], s.shape[1], _p(t), t.shape[1], _p(newtgt), path.ctypes.data_as(_lp)), "align")
    return s, newtgt, path
