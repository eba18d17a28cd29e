"""ctypes binding of libvcb200.so -- the same symbols a Julia ``ccall`` shim binds
(include/vcb200.h).

The shared library is built in-tree (``voiceconversion.jl_b200/libvcb200.so``) by
``__graft_entry__.build()`` / ``make -C voiceconversion.jl_b200/csrc``.  There is no fallback of
any kind: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvcb200.so")

OK, EDIM, ENOTPD, ESINGULAR, EARG, ENOMEM, ECUDA, EUNSUPPORTED = range(8)


class VCBError(RuntimeError):
    """Base class; ``code`` is the vcb_status value."""

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


class DimensionMismatch(VCBError):      # Julia DimensionMismatch (src/gmmmap.jl:102)
    pass


class PosDefException(VCBError):        # Julia PosDefException from MvNormal (src/gmm.jl:17)
    pass


class SingularException(VCBError):      # Julia SingularException from `^-1` (src/gmmmap.jl:35)
    pass


class ArgumentError(VCBError):
    pass


class CudaError(VCBError):
    pass


_EXC = {EDIM: DimensionMismatch, ENOTPD: PosDefException, ESINGULAR: SingularException,
        EARG: ArgumentError, ECUDA: CudaError}

_d = C.POINTER(C.c_double)
_l = C.POINTER(C.c_int64)
_i32 = C.c_int32
_i64 = C.c_int64
_vp = C.c_void_p

# symbol -> (restype, argtypes); must list every function include/vcb200.h declares
# (tests/test_abi.py cross-checks this table against the header and the built library).
SIGNATURES = {
    "vcb_version": (_i32, []),
    "vcb_last_error": (_i32, [C.c_char_p, C.c_size_t]),
    "vcb_device_count": (_i32, [C.POINTER(_i32)]),
    "vcb_set_device": (_i32, [_i32]),
    "vcb_init": (_i32, [_i32]),
    "vcb_num_devices": (_i32, [C.POINTER(_i32)]),
    "vcb_host_alloc": (_i32, [C.POINTER(_vp), C.c_size_t]),
    "vcb_host_free": (_i32, [_vp]),
    "vcb_host_register": (_i32, [_vp, C.c_size_t]),
    "vcb_host_unregister": (_i32, [_vp]),
    "vcb_set_kernel_variant": (_i32, [_i32]),
    "vcb_launch_count": (_i64, []),
    "vcb_stage_timing": (_i32, [_i32]),
    "vcb_stage_times": (_i32, [_vp, _i32, C.POINTER(_i32)]),
    "vcb_gmmmap_create": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, C.POINTER(_vp)]),
    "vcb_gmmmap_destroy": (_i32, [_vp]),
    "vcb_gmmmap_dim": (_i32, [_vp, C.POINTER(_i32)]),
    "vcb_gmmmap_ncomponents": (_i32, [_vp, C.POINTER(_i32)]),
    "vcb_gmmmap_get_param": (_i32, [_vp, _i32, _vp]),
    "vcb_gmmmap_convert": (_i32, [_vp, _vp, _i32, _i64, _i64, _vp, _i64]),
    "vcb_gmmmap_convert_dev": (_i32, [_vp, _vp, _i32, _i64, _i64, _vp, _i64, _vp]),
    "vcb_gmmmap_vc": (_i32, [_vp, _vp, _i32, _i64, _vp]),
    "vcb_gmmmap_vc_dev": (_i32, [_vp, _vp, _i32, _i64, _vp, _vp]),
    "vcb_gmmmap_predict_proba": (_i32, [_vp, _vp, _i32, _i64, _i64, _vp]),
    "vcb_gmmmap_predict": (_i32, [_vp, _vp, _i32, _i64, _i64, _vp]),
    "vcb_traj_create": (_i32, [_vp, C.POINTER(_vp)]),
    "vcb_traj_destroy": (_i32, [_vp]),
    "vcb_traj_status": (_i32, [_vp, _vp]),
    "vcb_traj_get_Dy": (_i32, [_vp, _vp]),
    "vcb_traj_convert_batch": (_i32, [_vp, _vp, _i32, _i64, _vp, _i64, _i32, _vp, _i64, _vp, _vp]),
    "vcb_traj_convert_batch_dev": (_i32, [_vp, _vp, _i32, _i64, _vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp]),
    "vcb_traj_vc_batch": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, _vp]),
    "vcb_traj_vc_batch_dev": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, _vp, _vp]),
    "vcb_dtw_fit_batch": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "vcb_dtw_fit_batch_dev": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp]),
    "vcb_dtw_update": (_i32, [_vp, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp]),
    "vcb_traj_vc_static_batch": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, _vp]),
    "vcb_traj_vc_static_batch_dev": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, _vp, _vp]),
    "vcb_trajgv_create": (_i32, [_vp, _vp, _vp, C.POINTER(_vp)]),
    "vcb_trajgv_destroy": (_i32, [_vp]),
    "vcb_trajgv_convert_batch": (_i32, [_vp, _vp, _i32, _i64, _vp, _i64, _i32, _i32, C.c_double, _vp, _i64]),
    "vcb_trajgv_convert_batch_dev": (_i32, [_vp, _vp, _i32, _i64, _vp, _i64, _i32, _i32, C.c_double, _vp, _i64, _vp]),
    "vcb_trajgv_vc_batch": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, _i32, C.c_double, _vp]),
    "vcb_trajgv_vc_batch_dev": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, _i32, C.c_double, _vp, _vp]),
    "vcb_variance_scaling_batch": (_i32, [_vp, _i32, _vp, _i64, _vp, _i64, _vp, _i64]),
    "vcb_variance_scaling_batch_dev": (_i32, [_vp, _i32, _vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    "vcb_diffgmm": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "vcb_push_delta_batch": (_i32, [_vp, _i32, _vp, _i64, _vp]),
    "vcb_align_batch": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Loads libvcb200.so (once).  Raises if it has not been built -- there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C voiceconversion.jl_b200/csrc`). There is no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    lib().vcb_last_error(buf, len(buf))
    return buf.value.decode(errors="replace")


def check(rc: int) -> None:
    if rc != OK:
        raise _EXC.get(rc, VCBError)(rc, last_error() or f"libvcb200 status {rc}")


def ptr(a) -> int:
    """Address of a numpy array's or torch tensor's first element."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


def set_device(device: int) -> None:
    check(lib().vcb_set_device(device))


def init(ndev: int = 0) -> int:
    """Multi-device mode: the host-array batch calls shard over devices 0..ndev-1 (0 = all visible,
    1 = single device again).  Returns the number of devices in use."""
    check(lib().vcb_init(int(ndev)))
    n = _i32(0)
    check(lib().vcb_num_devices(C.byref(n)))
    return n.value


def device_count() -> int:
    n = _i32(0)
    check(lib().vcb_device_count(C.byref(n)))
    return n.value


def set_kernel_variant(variant: int) -> None:
    """0 = auto, 1 = CUDA-core fp32 posterior kernel (and the per-column barrier DTW kernel instead of the
    persistent warp-pipeline one), 2 = tcgen05 3xTF32 kernel (tests / profiling)."""
    check(lib().vcb_set_kernel_variant(variant))


def launch_count() -> int:
    return int(lib().vcb_launch_count())


def stage_timing(enable: bool) -> None:
    """Profiling aid: record CUDA events between the stages of trajectory / DTW device calls."""
    check(lib().vcb_stage_timing(int(bool(enable))))


def stage_times():
    """Stage durations (ms) of the most recent trajectory / DTW device call (synchronises on it)."""
    buf = (C.c_double * 12)()
    n = _i32(0)
    check(lib().vcb_stage_times(buf, 12, C.byref(n)))
    return [buf[i] for i in range(n.value)]
