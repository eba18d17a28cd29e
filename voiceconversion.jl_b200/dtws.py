"""``DTWs`` -- mirror of the reference's DTW sub-module (src/dtw.jl) over libvcb200.

Julia's ``fit!``, ``update!`` and ``set_template!`` are spelled ``fit``, ``update`` and
``set_template`` here.  State indices are 1-based like the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _f64(a) -> np.ndarray:
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


class DTW:
    """``DTW(; fstep=0, bstep=1)``  (src/dtw.jl:11-21)."""

    def __init__(self, fstep: int = 0, bstep: int = 1):
        self.fstep = int(fstep)
        self.bstep = int(bstep)
        self.template = np.zeros((1, 1), order="F")
        self.costtable = np.zeros((1, 1), order="F")
        self.backpointer = np.zeros((1, 1), dtype=np.int64, order="F")
        self.final_cost = None


def fit_batch(d: DTW, templates, tmpl_off, sequences, seq_off):
    """Batch extension of ``fit!``: pairs are stored back to back, ``*_off`` are frame offsets
    (npairs+1).  Accepts numpy arrays (D, total) or frame-major CUDA tensors (total, D).
    Returns (paths, final_cost)."""
    L = _lib.lib()
    to = np.ascontiguousarray(tmpl_off, dtype=np.int64)
    so = np.ascontiguousarray(seq_off, dtype=np.int64)
    n = len(to) - 1
    nt = templates.shape[0] if hasattr(templates, "is_cuda") else np.shape(templates)[1]
    ns = sequences.shape[0] if hasattr(sequences, "is_cuda") else np.shape(sequences)[1]
    for off, total in ((to, nt), (so, ns)):
        if off.ndim != 1 or len(off) != len(to) or n < 0 or (n > 0 and (off[0] < 0 or np.any(np.diff(off) < 0) or off[-1] > total)):
            raise _lib.ArgumentError(_lib.EARG, "offsets must be (npairs+1,), non-decreasing and within the arrays")
    if hasattr(templates, "is_cuda"):
        import torch
        if templates.shape[1] != sequences.shape[1]:
            raise _lib.DimensionMismatch(_lib.EDIM, "template and sequence dimensions differ")
        paths = torch.empty(sequences.shape[0], dtype=torch.int64, device=sequences.device)
        fc = torch.empty(n, dtype=torch.float64, device=sequences.device)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(L.vcb_dtw_fit_batch_dev(_lib.ptr(templates), _lib.ptr(to), _lib.ptr(sequences), _lib.ptr(so), n,
                                           templates.shape[1], d.fstep, d.bstep, _lib.ptr(paths), _lib.ptr(fc), st))
        return paths, fc
    tm, sq = _f64(templates), _f64(sequences)
    if tm.shape[0] != sq.shape[0]:
        raise _lib.DimensionMismatch(_lib.EDIM, "template and sequence dimensions differ")
    paths = np.empty(sq.shape[1], dtype=np.int64)
    fc = np.empty(n)
    _lib.check(L.vcb_dtw_fit_batch(_lib.ptr(tm), _lib.ptr(to), _lib.ptr(sq), _lib.ptr(so), n, tm.shape[0],
                                   d.fstep, d.bstep, _lib.ptr(paths), _lib.ptr(fc)))
    return paths, fc


def fit(d: DTW, template, sequence=None, tables: bool = False) -> np.ndarray:
    """``fit!(d, template, sequence)`` / ``fit!(d, sequence)``  (src/dtw.jl:93-130): the 1-based
    template index aligned with every sequence frame.

    The fused kernel keeps the cost table on the chip, so ``d.costtable`` / ``d.backpointer`` are NOT
    filled (the reference leaves its S x (T+1) tables behind, :122-123).  ``tables=True`` rebuilds them
    exactly through ``update`` (one library call per frame) for callers that read them."""
    if sequence is None:
        template, sequence = d.template, template
    tm, sq = _f64(template), _f64(sequence)
    if tables:
        if tm.shape[0] != sq.shape[0]:
            raise _lib.DimensionMismatch(_lib.EDIM, "template and sequence dimensions differ")
        set_template(d, tm)
        for t in range(sq.shape[1]):
            update(d, sq[:, t])
        d.final_cost = float(d.costtable[:, -1].min())
        return backward(d)
    d.template = tm
    paths, fc = fit_batch(d, tm, [0, tm.shape[1]], sq, [0, sq.shape[1]])
    d.final_cost = float(fc[0])
    return paths


def set_template(d: DTW, template) -> None:
    """``set_template!(d, template)`` + ``lazy_init!(d, S)``  (src/dtw.jl:38-41, 53-56)."""
    d.template = _f64(template)
    S = d.template.shape[1]
    d.costtable = np.arange(1, S + 1, dtype=np.float64).reshape(S, 1, order="F")
    d.backpointer = np.arange(1, S + 1, dtype=np.int64).reshape(S, 1, order="F")


def update(d: DTW, v) -> None:
    """``update!(d, v)``  (src/dtw.jl:61-90): append one column to the tables."""
    v = _f64(v)
    D, S = d.template.shape
    if v.shape != (D,):
        raise _lib.DimensionMismatch(_lib.EDIM, "Inconsistent dimentions.")
    last = np.ascontiguousarray(d.costtable[:, -1])
    newcost = np.empty(S)
    newbp = np.empty(S, dtype=np.int64)
    _lib.check(_lib.lib().vcb_dtw_update(_lib.ptr(d.template), D, S, _lib.ptr(last), _lib.ptr(v), d.fstep, d.bstep,
                                         _lib.ptr(newcost), _lib.ptr(newbp)))
    d.costtable = np.asfortranarray(np.hstack([d.costtable, newcost[:, None]]))
    d.backpointer = np.asfortranarray(np.hstack([d.backpointer, newbp[:, None]]))


def backward(d: DTW) -> np.ndarray:
    """``backward(d)``  (src/dtw.jl:133-145) on the tables built by ``update``: pure index
    chasing over T entries, done on the host."""
    T = d.costtable.shape[1] - 1
    path = np.zeros(T, dtype=np.int64)
    path[T - 1] = int(np.argmin(d.costtable[:, T])) + 1     # indmin: first minimum
    for i in range(T, 1, -1):
        path[i - 2] = d.backpointer[path[i - 1] - 1, i]
    return path
