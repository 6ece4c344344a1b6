"""ctypes binding of include/stacker_cuda.h — the same C ABI the Rust crate binds in rust/src/ffi.rs.

There is no fallback: if the CUDA library has not been built (python __graft_entry__.py, or
`make -C libstacker.rs_b200/csrc`) importing this module raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstacker_cuda.so")

STK_OK, STK_ERR_BAD_ARG, STK_ERR_CUDA, STK_ERR_NOT_ENOUGH, STK_ERR_ECC_NOCONV, STK_ERR_ECC_NAN, \
    STK_ERR_CRITERIA, STK_ERR_STATE, STK_ERR_UNSUPPORTED, STK_ERR_NOMEM = range(10)
STK_TERM_COUNT, STK_TERM_EPS = 1, 2
STK_BORDER_CONSTANT = 0
STK_ABI_VERSION = 5


class EccConfig(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32),
        ("motion_type", C.c_int32), ("criteria_type", C.c_int32), ("max_count", C.c_int32),
        ("epsilon", C.c_double), ("gauss_filt_size", C.c_int32), ("device", C.c_int32),
        ("lanes", C.c_int32), ("seed_reference", C.c_int32), ("align", C.c_int32),
        ("ecc_width", C.c_int32), ("ecc_height", C.c_int32),
    ]


class FrameResult(C.Structure):
    _fields_ = [
        ("tag", C.c_int64), ("warp", C.c_float * 9), ("rho", C.c_double),
        ("iterations", C.c_int32), ("status", C.c_int32),
    ]


class PeerHandle(C.Structure):
    """stk_peer_handle: opaque bytes a rank publishes so that the other ranks can map its partial stack,
    its flag block and (rank 0) its output buffer."""
    _fields_ = [("bytes", C.c_ubyte * 256)]


# every symbol include/stacker_cuda.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "stk_abi_version": (C.c_int, []),
    "stk_last_error": (C.c_char_p, []),
    "stk_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "stk_scaled_size": (C.c_int, [C.c_int, C.c_int, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "stk_pinned_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "stk_pinned_free": (C.c_int, [_P]),
    "stk_ecc_create": (C.c_int, [C.POINTER(EccConfig), C.POINTER(_P)]),
    "stk_ecc_destroy": (C.c_int, [_P]),
    "stk_ecc_set_reference": (C.c_int, [_P, _P, C.c_size_t]),
    "stk_ecc_set_reference_device": (C.c_int, [_P, _P, C.c_size_t]),
    "stk_ecc_set_input_stream": (C.c_int, [_P, _P, C.c_int]),
    "stk_ecc_submit_frame": (C.c_int, [_P, _P, C.c_size_t, C.c_int64]),
    "stk_ecc_submit_frame_pinned": (C.c_int, [_P, _P, C.c_size_t, C.c_int64]),
    "stk_ecc_submit_frame_device": (C.c_int, [_P, _P, C.c_size_t, C.c_int64]),
    "stk_ecc_acquire_frame_buffer": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "stk_ecc_submit_acquired": (C.c_int, [_P, _P, C.c_int64]),
    "stk_ecc_release_frame_buffer": (C.c_int, [_P, _P]),
    "stk_ecc_submit_warp": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int64]),
    "stk_ecc_submit_warp_device": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int64]),
    "stk_ecc_submit_warp_affine": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int64]),
    "stk_ecc_submit_warp_affine_device": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int64]),
    "stk_ecc_sync": (C.c_int, [_P]),
    "stk_ecc_results": (C.c_int, [_P, C.POINTER(FrameResult), C.c_int, C.POINTER(C.c_int)]),
    "stk_ecc_finish": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "stk_ecc_partial": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "stk_ecc_finish_from": (C.c_int, [_P, _P, C.c_int, _P, C.c_size_t]),
    "stk_ecc_finish_device": (C.c_int, [_P, _P, C.c_int, _P]),
    "stk_ecc_peer_export": (C.c_int, [_P, C.POINTER(PeerHandle)]),
    "stk_ecc_peer_connect": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(PeerHandle)]),
    "stk_ecc_peer_connect_local": (C.c_int, [C.POINTER(_P), C.c_int]),
    "stk_ecc_peer_reduce": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    "stk_ecc_peer_reduce_scatter": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "stk_ecc_peer_slice_to_host": (C.c_int, [_P, _P]),
    "stk_ecc_peer_disconnect": (C.c_int, [_P]),
    "stk_ecc_reset": (C.c_int, [_P]),
    "stk_ecc_launch_count": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "stk_ecc_set_profiling": (C.c_int, [_P, C.c_int]),
    "stk_ecc_stage_times": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "stk_prep_grey_blur": (C.c_int, [_P, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_size_t]),
    "stk_grey_resize_area": (C.c_int, [_P, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_size_t]),
    "stk_ecc_debug_iteration": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_int,
                                          C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "stk_ecc_debug_timing": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_int)]),
    "stk_tenengrad": (C.c_int, [_P, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "stk_tenengrad_device": (C.c_int, [_P, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "stk_tenengrad_batch_device": (C.c_int, [_P, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "stk_sharpness_all": (C.c_int, [_P, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "stk_sharpness_all_batch_device": (C.c_int, [_P, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
}


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the sm_100a library first (python -c 'import __graft_entry__ as g; "
            "g.build()' or make -C libstacker.rs_b200/csrc).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.stk_abi_version() != STK_ABI_VERSION:
        raise ImportError("libstacker_cuda.so ABI version mismatch")
    return lib


lib = load()


def last_error() -> str:
    msg = lib.stk_last_error()
    return msg.decode("utf-8", "replace") if msg else ""
