"""Host-side mirror of the libstacker public API (/root/reference/src/lib.rs) over the C ABI.

The reference is a Rust crate; Rust is not available in this image, so this module (and the C++ mirror
in host/, and the uncompiled Rust crate in rust/) drive the SAME extern "C" entry points the Rust
wrappers bind.  What stays on the host is what the reference keeps on the host: file decode
(imgcodecs::imread) and, for keypoint_match, the ORB / BFMatcher / findHomography stages — both through
OpenCV (`cv2`), which the crate links as well.  Everything numeric on the ECC align-and-stack path runs
in the CUDA library; there is no CPU fallback.
"""
from __future__ import annotations

import atexit
import ctypes as C
import enum
import os
import threading
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import lib

BORDER_CONSTANT = 0          # opencv::core::BORDER_CONSTANT
BORDER_REPLICATE, BORDER_REFLECT, BORDER_WRAP, BORDER_REFLECT_101 = 1, 2, 3, 4      # the other cv::BorderTypes the warp path implements
RANSAC = 8                   # opencv::calib3d::RANSAC
LMEDS = 4


# ---- errors: /root/reference/src/lib.rs:27-45 -----------------------------------------------------
class StackerError(Exception):
    """Base of the crate's error enum."""


class OpenCvError(StackerError):
    """StackerError::OpenCvError — any failure the reference gets from OpenCV, including
    findTransformECC's StsNoConv (src/lib.rs:777)."""


class NotEnoughFiles(StackerError):
    def __init__(self):
        super().__init__("Not enough files")


class NotImplementedError_(StackerError):
    """StackerError::NotImplemented."""


class InvalidPathEncoding(StackerError):
    pass


class InvalidParams(StackerError):
    def __init__(self, msg):
        super().__init__(f"Invalid parameter(s) {msg}")


class ProcessingError(StackerError):
    def __init__(self, msg):
        super().__init__(f"Internal error {msg}")


def _raise(rc: int):
    msg = _ffi.last_error()
    if rc in (_ffi.STK_ERR_ECC_NOCONV, _ffi.STK_ERR_ECC_NAN, _ffi.STK_ERR_CRITERIA):
        raise OpenCvError(msg)
    if rc == _ffi.STK_ERR_BAD_ARG:
        raise InvalidParams(msg)
    if rc == _ffi.STK_ERR_NOT_ENOUGH:
        raise NotEnoughFiles()
    if rc == _ffi.STK_ERR_UNSUPPORTED:
        raise NotImplementedError_(msg)
    raise ProcessingError(msg)


def _check(rc: int):
    if rc != _ffi.STK_OK:
        _raise(rc)


# ---- parameter types: src/lib.rs:48-73, :603-623 -----------------------------------------------------
class MotionType(enum.IntEnum):
    """Discriminants are OpenCV's MOTION_* (src/lib.rs:603-609)."""
    Translation = 0
    Euclidean = 1
    Affine = 2
    Homography = 3


@dataclass(frozen=True)
class EccMatchParameters:
    motion_type: MotionType
    max_count: Optional[int]
    epsilon: Optional[float]
    gauss_filt_size: int


@dataclass(frozen=True)
class KeyPointMatchParameters:
    """Defaults are `impl Default` at src/utils.rs:250-261."""
    method: int = RANSAC
    ransac_reproj_threshold: float = 3.0
    match_keep_ratio: float = 0.75
    match_ratio: float = 0.8
    border_mode: int = BORDER_CONSTANT
    border_value: Sequence[float] = (0.0, 0.0, 0.0, 0.0)


def term_criteria(params: EccMatchParameters):
    """From<EccMatchParameters> for Result<TermCriteria> (src/utils.rs:159-170): (typ, max_count, epsilon);
    unset fields stay at TermCriteria::default() == 0."""
    typ, mc, eps = 0, 0, 0.0
    if params.max_count is not None:
        typ |= _ffi.STK_TERM_COUNT
        mc = int(params.max_count)
    if params.epsilon is not None:
        typ |= _ffi.STK_TERM_EPS
        eps = float(params.epsilon)
    return typ, mc, eps


# ---- low-level stack context ---------------------------------------------------------------------------
def _device_view(obj, expect_shape=None):
    """(ptr, pitch_bytes) of an array living on the GPU (anything with __cuda_array_interface__, e.g. a
    torch CUDA tensor of shape HxWxC uint8), or None for host arrays.  The array must be 8-bit with dense
    pixels (only the row pitch may be padded) and, when `expect_shape` is given, of exactly that shape: the
    library reads height * pitch bytes from the pointer, so anything else would be an out-of-bounds device read.
    Stream contract: the library's lane streams are non-blocking; EccStack orders every device frame behind the
    stream that produced it (`_producer_stream`: torch's current stream for torch tensors, the `stream` entry of the
    interface otherwise) through stk_ecc_set_input_stream — a device-side dependency, nothing blocks on the host."""
    iface = getattr(obj, "__cuda_array_interface__", None)
    if iface is None:
        return None
    shape = tuple(int(v) for v in iface["shape"])
    if iface.get("typestr") not in ("|u1", "<u1", ">u1"):
        raise OpenCvError(f"device frame must be 8-bit unsigned (got typestr {iface.get('typestr')!r})")
    if expect_shape is not None and shape != tuple(expect_shape):
        raise OpenCvError(f"device frame has shape {shape}, the stack expects {tuple(expect_shape)}")
    strides = iface.get("strides")
    dense = [1]
    for d in reversed(shape[1:]):
        dense.insert(0, dense[0] * d)
    if strides:
        if tuple(int(v) for v in strides[1:]) != tuple(dense[1:]):
            raise OpenCvError(f"device frame must have dense pixels (strides {tuple(strides)})")
        if int(strides[0]) < dense[0]:
            raise OpenCvError("device frame rows overlap")
    pitch = int(strides[0]) if strides else dense[0]
    return int(iface["data"][0]), pitch


def _producer_stream(obj):
    """The CUDA stream handle (int; 0 = the legacy default stream) whose queued work produces `obj`, or None if unknown:
    torch's current stream on the tensor's device, else the `stream` entry of __cuda_array_interface__ (v3: 1 = legacy
    default, 2 = per-thread default — not expressible as a handle here, treated as unknown)."""
    import sys
    torch = sys.modules.get("torch")
    if torch is not None and isinstance(obj, torch.Tensor):
        return int(torch.cuda.current_stream(obj.device).cuda_stream)
    st = getattr(obj, "__cuda_array_interface__", {}).get("stream")
    if st is None or st == 2:
        return None
    return 0 if st == 1 else int(st)


class EccStack:
    """One stack on one CUDA device: wraps stk_ecc_ctx.  `params=None` makes a warp-only context
    (keypoint_match tail)."""

    def __init__(self, width: int, height: int, channels: int = 3, params: Optional[EccMatchParameters] = None,
                 device: int = -1, lanes: int = 0, seed_reference: bool = True, ecc_size=None):
        """`ecc_size=(w, h)`: ecc_match_scaling_down — ECC runs on greys INTER_AREA-resized to that size
        (see scaled_size) and the matrix is rescaled to full resolution on the device."""
        cfg = _ffi.EccConfig()
        if ecc_size is not None:
            cfg.ecc_width, cfg.ecc_height = int(ecc_size[0]), int(ecc_size[1])
        cfg.width, cfg.height, cfg.channels = int(width), int(height), int(channels)
        cfg.device, cfg.lanes = int(device), int(lanes)
        cfg.seed_reference = 1 if seed_reference else 0
        if params is not None:
            typ, mc, eps = term_criteria(params)
            cfg.align = 1
            cfg.motion_type = int(params.motion_type)
            cfg.criteria_type, cfg.max_count, cfg.epsilon = typ, mc, eps
            cfg.gauss_filt_size = int(params.gauss_filt_size)
        else:
            cfg.align = 0
        self.width, self.height, self.channels = cfg.width, cfg.height, cfg.channels
        self.lanes = int(lanes) if lanes > 0 else 4        # the library's default
        self._ctx = C.c_void_p()
        self._keep = []          # host arrays that must outlive asynchronous copies
        self._submitted = 0      # frames handed to the library since the last reset (sizes the results() buffer)
        _check(lib.stk_ecc_create(C.byref(cfg), C.byref(self._ctx)))

    # -- lifetime
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            lib.stk_ecc_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- frames
    def _host_frame(self, frame: np.ndarray):
        if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != self.channels:
            raise OpenCvError(f"frame must be 8-bit HxWx{self.channels} (got {frame.dtype} {frame.shape})")
        if frame.shape[0] != self.height or frame.shape[1] != self.width:
            raise OpenCvError(f"frame size {frame.shape[1]}x{frame.shape[0]} differs from the stack's "
                              f"{self.width}x{self.height}")
        if frame.strides[2] != 1 or frame.strides[1] != self.channels:
            frame = np.ascontiguousarray(frame)
        return frame, frame.ctypes.data, frame.strides[0]

    def _order_after(self, frame):
        """Tell the library which stream produced the device frame about to be submitted (only when it changes)."""
        ps = _producer_stream(frame)
        if ps != getattr(self, "_input_stream", -1):
            _check(lib.stk_ecc_set_input_stream(self._ctx, C.c_void_p(ps or 0), 0 if ps is None else 1))
            self._input_stream = ps

    def set_reference(self, frame):
        dv = _device_view(frame, (self.height, self.width, self.channels))
        if dv is not None:
            self._order_after(frame)
            _check(lib.stk_ecc_set_reference_device(self._ctx, dv[0], dv[1]))
            self._keep.append(frame)
        else:
            f, ptr, pitch = self._host_frame(frame)
            self._keep.append(f)      # a page-locked reference is copied asynchronously: keep it until the next sync
            _check(lib.stk_ecc_set_reference(self._ctx, ptr, pitch))

    def submit(self, frame, tag: int = 0, pinned: bool = False):
        dv = _device_view(frame, (self.height, self.width, self.channels))
        if dv is not None:
            self._keep.append(frame)
            self._order_after(frame)
            _check(lib.stk_ecc_submit_frame_device(self._ctx, dv[0], dv[1], int(tag)))
            self._submitted += 1
            return
        f, ptr, pitch = self._host_frame(frame)
        if pinned:
            self._keep.append(f)
            _check(lib.stk_ecc_submit_frame_pinned(self._ctx, ptr, pitch, int(tag)))
            self._submitted += 1
        else:
            _check(lib.stk_ecc_submit_frame(self._ctx, ptr, pitch, int(tag)))
            self._submitted += 1

    # -- host feed: pinned ring buffers as decode targets
    def acquire_buffer(self) -> np.ndarray:
        """A free pinned frame buffer of the context's ring as an HxWxC uint8 array (blocks until one is free).
        Fill it from any thread, then submit_acquired() it (or release_buffer())."""
        ptr, pitch = C.c_void_p(), C.c_size_t()
        _check(lib.stk_ecc_acquire_frame_buffer(self._ctx, C.byref(ptr), C.byref(pitch)))
        n = self.height * self.width * self.channels
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,))
        return arr.reshape(self.height, self.width, self.channels)

    def submit_acquired(self, buf: np.ndarray, tag: int = 0):
        _check(lib.stk_ecc_submit_acquired(self._ctx, buf.ctypes.data, int(tag)))
        self._submitted += 1

    def release_buffer(self, buf: np.ndarray):
        _check(lib.stk_ecc_release_frame_buffer(self._ctx, buf.ctypes.data))

    def submit_warp(self, frame, h, border_mode: int = BORDER_CONSTANT, border_value=(0, 0, 0, 0), tag: int = 0):
        hm = (C.c_double * 9)(*np.asarray(h, np.float64).reshape(9))
        bv = (C.c_double * 4)(*[float(v) for v in border_value])
        dv = _device_view(frame, (self.height, self.width, self.channels))
        if dv is not None:
            self._keep.append(frame)
            self._order_after(frame)
            _check(lib.stk_ecc_submit_warp_device(self._ctx, dv[0], dv[1], hm, int(border_mode), bv, int(tag)))
            self._submitted += 1
            return
        f, ptr, pitch = self._host_frame(frame)
        _check(lib.stk_ecc_submit_warp(self._ctx, ptr, pitch, hm, int(border_mode), bv, int(tag)))
        self._submitted += 1

    def submit_warp_affine(self, frame, m, border_mode: int = BORDER_CONSTANT, border_value=(0, 0, 0, 0), tag: int = 0):
        """warp_affine(img_f32, M 2x3 f64, ..) + accumulate (src/lib.rs:782-790 with a caller-supplied matrix)."""
        mm = (C.c_double * 6)(*np.asarray(m, np.float64).reshape(-1)[:6])
        bv = (C.c_double * 4)(*[float(v) for v in border_value])
        dv = _device_view(frame, (self.height, self.width, self.channels))
        if dv is not None:
            self._keep.append(frame)
            self._order_after(frame)
            _check(lib.stk_ecc_submit_warp_affine_device(self._ctx, dv[0], dv[1], mm, int(border_mode), bv, int(tag)))
            self._submitted += 1
            return
        f, ptr, pitch = self._host_frame(frame)
        _check(lib.stk_ecc_submit_warp_affine(self._ctx, ptr, pitch, mm, int(border_mode), bv, int(tag)))
        self._submitted += 1

    # -- completion
    def sync(self):
        rc = lib.stk_ecc_sync(self._ctx)
        self._keep.clear()
        _check(rc)

    def results(self):
        n = C.c_int(0)
        cap = max(1, self._submitted)
        buf = (_ffi.FrameResult * cap)()
        _check(lib.stk_ecc_results(self._ctx, buf, cap, C.byref(n)))
        out = []
        for i in range(n.value):
            r = buf[i]
            out.append(dict(tag=r.tag, warp=np.array(r.warp[:], np.float32).reshape(3, 3), rho=r.rho,
                            iterations=r.iterations, status=r.status))
        return out

    def finish(self, divisor: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """`out`: an HxWxC float32 array to receive the stack (e.g. page-locked memory: the device-to-host copy of a
        4K stack is 2 ms into pinned memory, several times that into pageable memory)."""
        shape = (self.height, self.width, self.channels)
        if out is None:
            out = np.empty(shape, np.float32)
        elif out.dtype != np.float32 or out.shape != shape or out.strides[1:] != (4 * self.channels, 4):
            raise InvalidParams(f"out must be a float32 array of shape {shape} with dense rows")
        _check(lib.stk_ecc_finish(self._ctx, int(divisor), out.ctypes.data, out.strides[0]))
        self._keep.clear()
        return out

    def partial(self):
        """(device pointer, n_floats) of this device's partial stack, lanes already summed."""
        ptr, n = C.c_void_p(), C.c_size_t()
        _check(lib.stk_ecc_partial(self._ctx, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def finish_from(self, d_sum_ptr: Optional[int], divisor: int) -> np.ndarray:
        out = np.empty((self.height, self.width, self.channels), np.float32)
        _check(lib.stk_ecc_finish_from(self._ctx, d_sum_ptr, int(divisor), out.ctypes.data, out.strides[0]))
        self._keep.clear()
        return out

    def finish_device(self, d_sum_ptr: Optional[int], divisor: int, d_out_ptr: int):
        _check(lib.stk_ecc_finish_device(self._ctx, d_sum_ptr, int(divisor), d_out_ptr))

    # -- multi-GPU exchange over peer memory (stk_ecc_peer_*; csrc/peer_reduce.cuh)
    def peer_export(self) -> bytes:
        """This rank's stk_peer_handle as bytes: all-gather them and give the list to peer_connect."""
        h = _ffi.PeerHandle()
        _check(lib.stk_ecc_peer_export(self._ctx, C.byref(h)))
        return bytes(h.bytes)

    def peer_connect(self, rank: int, world: int, handles):
        """`handles[r]` = rank r's peer_export() bytes (one process per GPU; CUDA IPC mappings)."""
        if len(handles) != world:
            raise InvalidParams(f"{len(handles)} peer handles for world size {world}")
        arr = (_ffi.PeerHandle * world)()
        for r, b in enumerate(handles):
            if len(b) != C.sizeof(_ffi.PeerHandle):
                raise InvalidParams(f"peer handle {r} has {len(b)} bytes")
            C.memmove(C.byref(arr[r]), bytes(b), len(b))
        _check(lib.stk_ecc_peer_connect(self._ctx, int(rank), int(world), arr))

    @staticmethod
    def peer_connect_local(stacks):
        """All ranks' contexts live in THIS process (one per device, rank order): peer access, no IPC."""
        arr = (C.c_void_p * len(stacks))(*[s._ctx.value for s in stacks])
        _check(lib.stk_ecc_peer_connect_local(arr, len(stacks)))

    def peer_reduce(self, divisor: int) -> Optional[int]:
        """The exchange step + divide (collective, asynchronous).  Rank 0: device pointer of the finished
        stack (valid after sync()); other ranks: None."""
        out = C.c_void_p()
        _check(lib.stk_ecc_peer_reduce(self._ctx, int(divisor), C.byref(out)))
        return out.value

    def peer_reduce_scatter(self, divisor: int):
        """The exchange step + divide with every rank keeping the finished pixels of its own slice:
        (device pointer of the slice, begin, count) in floats of the flat H*W*C stack."""
        ptr, b, n = C.c_void_p(), C.c_size_t(), C.c_size_t()
        _check(lib.stk_ecc_peer_reduce_scatter(self._ctx, int(divisor), C.byref(ptr), C.byref(b), C.byref(n)))
        return ptr.value, b.value, n.value

    def peer_slice_to_host(self, host_ptr: int):
        """Queue the device-to-host copy of this rank's slice into the dense host stack at `host_ptr` (done
        after sync())."""
        _check(lib.stk_ecc_peer_slice_to_host(self._ctx, C.c_void_p(int(host_ptr))))

    def peer_disconnect(self):
        _check(lib.stk_ecc_peer_disconnect(self._ctx))

    def set_profiling(self, enabled: bool):
        _check(lib.stk_ecc_set_profiling(self._ctx, 1 if enabled else 0))

    def stage_times(self):
        """dict(prep_ms, loop_ms, warp_ms, frames, iterations) summed over the frames submitted with
        profiling on."""
        ms = (C.c_double * 3)()
        nf, it = C.c_int64(), C.c_int64()
        _check(lib.stk_ecc_stage_times(self._ctx, ms, C.byref(nf), C.byref(it)))
        return dict(prep_ms=ms[0], loop_ms=ms[1], warp_ms=ms[2], frames=nf.value, iterations=it.value)

    def debug_iteration(self, frame: np.ndarray, warp_in):
        """One ECC iteration from `warp_in` (3x3): (totals f64[NV], warp_out 3x3 f32, rho, status)."""
        f, ptr, pitch = self._host_frame(frame)
        win = (C.c_float * 9)(*np.asarray(warp_in, np.float32).reshape(9))
        tot = (C.c_double * 128)()
        nv, rho, status = C.c_int(), C.c_double(), C.c_int()
        wout = (C.c_float * 9)()
        _check(lib.stk_ecc_debug_iteration(self._ctx, ptr, pitch, win, tot, 128, C.byref(nv), wout, C.byref(rho),
                                           C.byref(status)))
        return (np.array(tot[:nv.value]), np.array(wout[:], np.float32).reshape(3, 3), rho.value, status.value)

    def debug_timing(self, frame: np.ndarray, warp_in, iters: int = 3):
        """%globaltimer stamps (ns) of the last of `iters` back-to-back iteration kernels:
        (tiles[n_tiles, 4], tail[4])."""
        f, ptr, pitch = self._host_frame(frame)
        win = (C.c_float * 9)(*np.asarray(warp_in, np.float32).reshape(9))
        cap = 4 * 65536 + 4
        buf = (C.c_uint64 * cap)()
        nt = C.c_int()
        _check(lib.stk_ecc_debug_timing(self._ctx, ptr, pitch, win, int(iters), buf, cap, C.byref(nt)))
        a = np.frombuffer(buf, dtype=np.uint64, count=nt.value * 4 + 4).copy()
        return a[:nt.value * 4].reshape(nt.value, 4), a[nt.value * 4:]

    def reset(self):
        _check(lib.stk_ecc_reset(self._ctx))
        self._keep.clear()
        self._submitted = 0

    def launch_count(self) -> int:
        n = C.c_int64()
        _check(lib.stk_ecc_launch_count(self._ctx, C.byref(n)))
        return n.value


def scaled_size(width: int, height: int, scale_down: float):
    """utils::scale_image's size rule (src/utils.rs:186-200) + ecc_match_scaling_down's validation
    (src/lib.rs:876-888), evaluated by the library: (small_width, small_height) or InvalidParams."""
    sw, sh = C.c_int(), C.c_int()
    _check(lib.stk_scaled_size(int(width), int(height), C.c_float(scale_down), C.byref(sw), C.byref(sh)))
    return sw.value, sh.value


def grey_resize_area(frame: np.ndarray, out_width: int, out_height: int, device: int = -1) -> np.ndarray:
    """cvtColor(BGR2GRAY) (skipped for 2-D input) + cv::resize(INTER_AREA): utils::scale_image on the grey
    frame (src/utils.rs:186-214)."""
    frame = np.ascontiguousarray(frame)
    h, w = frame.shape[:2]
    ch = 1 if frame.ndim == 2 else frame.shape[2]
    out = np.empty((out_height, out_width), np.uint8)
    _check(lib.stk_grey_resize_area(frame.ctypes.data, frame.strides[0], w, h, ch, int(out_width), int(out_height),
                                    device, out.ctypes.data, out.strides[0]))
    return out


def prep_grey_blur(frame: np.ndarray, ksize: int, device: int = -1) -> np.ndarray:
    """The f32 plane findTransformECC builds from an 8-bit BGR frame: grey -> f32 -> GaussianBlur(k)."""
    frame = np.ascontiguousarray(frame)
    h, w, ch = frame.shape
    out = np.empty((h, w), np.float32)
    _check(lib.stk_prep_grey_blur(frame.ctypes.data, frame.strides[0], w, h, ch, int(ksize), device,
                                  out.ctypes.data, out.strides[0]))
    return out


# ---- decode (host; stays OpenCV like the reference) -----------------------------------------------------
def imread(path, flags=None):
    """utils::imread (src/utils.rs:111-117)."""
    import cv2
    p = os.fspath(path)
    if not isinstance(p, str):
        raise InvalidPathEncoding(repr(path))
    img = cv2.imread(p, cv2.IMREAD_UNCHANGED if flags is None else flags)
    if img is None:
        # OpenCV returns an empty Mat; the first OpenCV call on it fails -> OpenCvError in the reference
        raise OpenCvError(f"imread failed for {p}")
    return img


def _load(item):
    return item if isinstance(item, np.ndarray) else imread(item)


def _check_colour_frame(img: np.ndarray):
    if img.dtype != np.uint8:
        raise OpenCvError("findTransformECC: images must have 8uC1 type after cvtColor (got a non-8-bit file)")
    if img.ndim != 3 or img.shape[2] not in (3, 4):
        raise OpenCvError("cvtColor(BGR2GRAY): input must have 3 or 4 channels")


# ---- context cache ---------------------------------------------------------------------------------------
# A context owns ~0.7 GB of device buffers, pinned staging and CUDA graphs for a 4K stack; creating and destroying
# it costs 50-300 ms (cudaMalloc / cudaHostAlloc / cudaFree), several times the 19 ms the stack itself takes.  The
# one-shot plugin calls therefore park their single-device context here and the next call with the same geometry and
# parameters resets and reuses it.  STK_CONTEXT_CACHE=<n> bounds the idle contexts kept per process (default 2,
# 0 disables); clear_context_cache() releases them.
_CTX_CACHE = {}            # key -> idle contexts of that geometry, most recently parked last
_CTX_CACHE_ORDER = []      # (key, stack) in parking order: the front is evicted first (LRU)
_CTX_CACHE_LOCK = threading.Lock()
_CTX_CACHE_MAX = max(0, int(os.environ.get("STK_CONTEXT_CACHE", "2")))


def _ctx_key(w, h, ch, params, device, ecc_size, seed_reference, lanes):
    pk = None if params is None else (int(params.motion_type), params.max_count, params.epsilon, int(params.gauss_filt_size))
    return (int(device), int(w), int(h), int(ch), pk, None if ecc_size is None else tuple(int(v) for v in ecc_size),
            bool(seed_reference), int(lanes))


def _acquire_stack(w, h, ch, params, device, ecc_size=None, seed_reference=True, lanes=0):
    key = _ctx_key(w, h, ch, params, device, ecc_size, seed_reference, lanes)
    with _CTX_CACHE_LOCK:
        idle = _CTX_CACHE.get(key)
        st = idle.pop() if idle else None
        if st is not None:
            _CTX_CACHE_ORDER.remove((key, st))
    if st is not None:
        try:
            st.reset()
            return st, key
        except StackerError:
            st.close()
    return EccStack(w, h, ch, params, device=device, lanes=lanes, seed_reference=seed_reference, ecc_size=ecc_size), key


def _release_stack(st, key, reusable: bool):
    """Park a context for the next call of the same geometry; beyond STK_CONTEXT_CACHE idle contexts the least recently
    parked one is destroyed (a service whose frame size changes does not keep the first sizes it ever saw)."""
    if not (reusable and _CTX_CACHE_MAX > 0):
        st.close()
        return
    evicted = []
    with _CTX_CACHE_LOCK:
        st._keep.clear()
        _CTX_CACHE.setdefault(key, []).append(st)
        _CTX_CACHE_ORDER.append((key, st))
        while len(_CTX_CACHE_ORDER) > _CTX_CACHE_MAX:
            k, old = _CTX_CACHE_ORDER.pop(0)
            _CTX_CACHE[k].remove(old)
            if not _CTX_CACHE[k]:
                del _CTX_CACHE[k]
            evicted.append(old)
    for old in evicted:
        old.close()


def clear_context_cache():
    """Destroy the idle contexts kept by ecc_match / keypoint_match (frees their device and pinned memory)."""
    with _CTX_CACHE_LOCK:
        stacks = [st for _, st in _CTX_CACHE_ORDER]
        _CTX_CACHE.clear()
        del _CTX_CACHE_ORDER[:]
    for st in stacks:
        st.close()


atexit.register(clear_context_cache)


# ---- ecc_match: src/lib.rs:702-847 ----------------------------------------------------------------------
def ecc_match(files: Iterable, params: EccMatchParameters, scale_down_width: Optional[float] = None, *,
              device: int = -1, devices=None, workers: Optional[int] = None, return_details: bool = False,
              pinned: bool = False, out: Optional[np.ndarray] = None):
    """Align every frame to the first with ECC and average them.

    `files`: paths (decoded on host threads with cv2.imread(IMREAD_UNCHANGED), as
    utils::read_grey_and_f32 does) or already-decoded HxWxC uint8 arrays.
    `devices`: several CUDA devices of the box, driven from this one process — one context per device, frames
    dealt round-robin, one fused exchange + divide over NVLink peer memory (what the Rust crate and
    `libstacker::ecc_match_on_devices` do); default: the single `device`.
    `pinned=True`: the caller's arrays live in page-locked memory (stk_pinned_alloc, torch pin_memory): they are
    uploaded from where they are, no staging copy; they must stay untouched until the call returns.  Pageable
    arrays and files go through the context's pinned ring, filled by `workers` host threads.
    `out`: optional HxWxC float32 array (ideally page-locked) that receives the stack on a single device.
    Returns the stacked image, float32 HxWxC in [0,1] (the reference's CV_32FC3 Mat).
    Errors: NotEnoughFiles (empty input); OpenCvError (no COUNT/EPS criteria, ECC non-convergence, bad
    image type); InvalidParams (scale_down_width out of range)."""
    items = list(files)
    if not items:
        raise NotEnoughFiles()
    typ, _, _ = term_criteria(params)
    first = _load(items[0])
    _check_colour_frame(first)
    h, w, ch = first.shape
    ecc_size = None
    if scale_down_width is not None:
        # ecc_match_scaling_down (src/lib.rs:849-1028): width validation (:876-888) comes before any ECC call
        ecc_size = scaled_size(w, h, float(np.float32(scale_down_width)))
    if not typ:
        raise OpenCvError("findTransformECC: criteria.type must have COUNT or EPS set")
    devs = [int(d) for d in devices] if devices else [int(device)]
    if len(set(devs)) != len(devs):
        raise InvalidParams("devices must be distinct")
    stacks = []
    cache_key, ok = None, False
    try:
        if len(devs) == 1:
            st0, cache_key = _acquire_stack(w, h, ch, params, devs[0], ecc_size, True)
            stacks.append(st0)
            st0.set_reference(first)
        else:
            for k, d in enumerate(devs):
                # only the first context seeds its accumulator with the unwarped frame 0 (src/lib.rs:752-754)
                stacks.append(EccStack(w, h, ch, params, device=d, ecc_size=ecc_size, seed_reference=(k == 0)))
                stacks[-1].set_reference(first)
        if len(stacks) > 1:
            EccStack.peer_connect_local(stacks)
        nd = len(stacks)
        n_workers = workers or min(8, os.cpu_count() or 1)
        rest = items[1:]
        if rest:
            if pinned and all(isinstance(i, np.ndarray) for i in rest):
                for k, fr in enumerate(rest):
                    _check_colour_frame(fr)
                    stacks[(k + 1) % nd].submit(fr, tag=k + 1, pinned=True)
            else:
                # Rayon: one task per frame (src/lib.rs:746-749).  Each task decodes and copies its frame into a
                # pinned ring buffer of the context it is dealt to (no library lock held); the main thread hands
                # the buffers over in file order, so the summation order — and the result — does not depend on
                # thread timing.  The window keeps at most one ring's worth of tasks alive per context, so a task
                # can always get its buffer.
                def load_into_ring(k, item):
                    fr = _load(item)
                    _check_colour_frame(fr)
                    if fr.shape != (h, w, ch):
                        raise OpenCvError(f"frame size {fr.shape[1]}x{fr.shape[0]} differs from the stack's {w}x{h}")
                    buf = stacks[(k + 1) % nd].acquire_buffer()
                    np.copyto(buf, fr)
                    return buf

                window = max(1, min(n_workers, 2 * stacks[0].lanes))
                with ThreadPoolExecutor(max_workers=window) as ex:
                    pending = []
                    it = iter(enumerate(rest))
                    done = False
                    while pending or not done:
                        while not done and len(pending) < window:
                            try:
                                k, item = next(it)
                            except StopIteration:
                                done = True
                                break
                            pending.append((k, ex.submit(load_into_ring, k, item)))
                        if pending:
                            k, fut = pending.pop(0)
                            stacks[(k + 1) % nd].submit_acquired(fut.result(), tag=k + 1)
        if nd == 1:
            out = stacks[0].finish(len(items), out)
        else:
            # Rayon's try_reduce + `/ n` as ONE exchange step: every exchange is queued before the first copy-out
            # (a copy into pageable memory blocks this thread until its device's exchange has finished)
            out = _finish_on_devices(stacks, len(items), (h, w, ch))   # its sync() also raises a frame's ECC failure (src/lib.rs:777)
        if return_details:
            res = sorted((r for st in stacks for r in st.results()), key=lambda r: r["tag"])
            ok = True
            return out, res
        ok = True
        return out
    finally:
        if cache_key is not None:
            _release_stack(stacks[0], cache_key, ok)      # a context that raised is destroyed, not reused
        else:
            for st in stacks:
                st.close()


# ---- keypoint_match: src/lib.rs:129-353 -----------------------------------------------------------------
def _orb(grey):
    import cv2
    return cv2.ORB_create().detectAndCompute(grey, None)      # utils::orb_detect_and_compute


def _scale_image(img, scale_down: float):
    """utils::scale_image (src/utils.rs:186-214) on the host, for the keypoint front end (which stays OpenCV)."""
    import cv2
    h, w = img.shape[:2]
    factor = float(scale_down) / float(w if w < h else h)
    return cv2.resize(img, (int(w * factor), int(h * factor)), interpolation=cv2.INTER_AREA)


def _keep_count(n_good: int, keep_ratio: float) -> int:
    """`(filtered_matches.len() as f32 * params.match_keep_ratio).round() as usize` (src/lib.rs:235, :472): f32 product, halves rounded
    AWAY from zero as Rust's f32::round does (6 matches at 0.75 keep 5; Python's round() would keep 4 and drop
    the frame below the 5-match minimum)."""
    x = np.float32(n_good) * np.float32(keep_ratio)
    return int(np.floor(x + np.float32(0.5))) if x >= 0 else 0


def _frame_homography(kp0, des0, img, params: KeyPointMatchParameters, scale_down: Optional[float] = None):
    """Host stages of src/lib.rs:200-287 (scale_down: :424-547, features on the INTER_AREA-downscaled grey and
    the homography taken back to full size with adjust_homography_for_scale_f64, src/utils.rs:218-248);
    None == the reference drops the frame."""
    import cv2
    grey = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    full_h, full_w = grey.shape
    if scale_down is not None:
        grey = _scale_image(grey, scale_down)
    kp, des = _orb(grey)
    if des is None or des0 is None:
        return None
    knn = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(des0, des, k=2)
    good = [m[0] for m in knn
            if len(m) == 2 and m[0].distance < np.float32(params.match_ratio) * m[1].distance]
    good.sort(key=lambda m: m.distance)
    good = good[:_keep_count(len(good), params.match_keep_ratio)]
    if len(good) < 5:
        return None
    src = np.float32([kp0[m.queryIdx].pt for m in good]).reshape(-1, 1, 2)
    dst = np.float32([kp[m.trainIdx].pt for m in good]).reshape(-1, 1, 2)
    try:
        hm, _ = cv2.findHomography(dst, src, params.method, params.ransac_reproj_threshold)
    except cv2.error:
        return None
    if hm is None or hm.shape != (3, 3) or abs(np.linalg.det(hm)) < 1e-6:
        return None
    if scale_down is not None:
        sx, sy = full_w / grey.shape[1], full_h / grey.shape[0]
        hm = hm.copy()
        hm[0, 2] *= sx
        hm[1, 2] *= sy
        hm[2, 0] /= sx
        hm[2, 1] /= sy
    return hm


def _finish_on_devices(stacks, divisor: int, shape) -> np.ndarray:
    """Rayon's try_reduce + `/ n` over several contexts of this process: ONE exchange step over NVLink peer
    memory, every device copying its slice of the result out.  Every exchange is queued before the first
    copy-out (a copy into pageable memory blocks this thread until its device's exchange has finished)."""
    out = np.empty(shape, np.float32)
    for st in stacks:
        st.peer_reduce_scatter(divisor)
    for st in stacks:
        st.peer_slice_to_host(out.ctypes.data)
    for st in stacks:
        st.sync()
    return out


def keypoint_match(files: Iterable, params: KeyPointMatchParameters = KeyPointMatchParameters(),
                   scale_down_width: Optional[float] = None, *, device: int = -1, devices=None,
                   workers: Optional[int] = None):
    """Returns (dropped, stacked f32 HxWxC).  ORB / BFMatcher / findHomography stay on the host (OpenCV),
    the final warp_perspective + accumulate + divide run on the GPU.  `devices`: several CUDA
    devices driven from this process (BASELINE configs[4]): the accepted frames are dealt round-robin to one
    warp-only context per device and the partial stacks meet in one exchange + divide.

    Deviation, documented: when frames are dropped the reference's result depends on how Rayon split the
    index range (src/lib.rs:307 seeds a worker's accumulator with a copy of frame 0); here dropped frames
    are simply left out and the sum is divided by n - dropped."""
    import cv2
    items = list(files)
    if not items:
        raise NotEnoughFiles()
    first = _load(items[0])
    _check_colour_frame(first)
    h, w, ch = first.shape
    sd = None
    if scale_down_width is not None:
        # keypoint_match_scale_down (src/lib.rs:355-600): only the upper bound is validated (:378-383)
        sd = float(np.float32(scale_down_width))
        if sd >= float(w):
            raise InvalidParams(f"scale_down_to was larger (or equal) to the full image width: full_size:{w}, "
                                f"scale_down_to:{sd:g}")
    grey0 = cv2.cvtColor(first, cv2.COLOR_BGR2GRAY)
    kp0, des0 = _orb(grey0 if sd is None else _scale_image(grey0, sd))
    dropped = 0

    def work(item):
        img = _load(item)
        _check_colour_frame(img)
        return img, _frame_homography(kp0, des0, img, params, sd)

    devs = [int(d) for d in devices] if devices else None
    if devs is not None and len(devs) > 1:
        if len(set(devs)) != len(devs):
            raise InvalidParams("devices must be distinct")
        stacks = []
        try:
            for k, d in enumerate(devs):
                stacks.append(EccStack(w, h, ch, None, device=d, seed_reference=(k == 0)))
                stacks[-1].set_reference(first)
            EccStack.peer_connect_local(stacks)
            n_workers = workers or min(8, os.cpu_count() or 1)
            accepted = 0
            with ThreadPoolExecutor(max_workers=n_workers) as ex:
                for k, (img, hm) in enumerate(ex.map(work, items[1:])):
                    if hm is None:
                        dropped += 1
                        continue
                    if img.shape[:2] != (h, w):
                        raise NotImplementedError_("frames of differing size")
                    accepted += 1
                    stacks[accepted % len(stacks)].submit_warp(img, hm, params.border_mode, params.border_value, tag=k + 1)
            if len(items) - dropped <= 0:
                raise InvalidParams("All images discarded: try modifying KeyPointMatchParameters::match_distance_threshold")
            return dropped, _finish_on_devices(stacks, len(items) - dropped, (h, w, ch))
        finally:
            for st in stacks:
                st.close()
    if devs:
        device = devs[0]
    st, cache_key = _acquire_stack(w, h, ch, None, device)
    ok = False
    try:
        st.set_reference(first)

        n_workers = workers or min(8, os.cpu_count() or 1)
        with ThreadPoolExecutor(max_workers=n_workers) as ex:
            for k, (img, hm) in enumerate(ex.map(work, items[1:])):
                if hm is None:
                    dropped += 1
                    continue
                if img.shape[:2] != (h, w):
                    raise NotImplementedError_("frames of differing size")
                st.submit_warp(img, hm, params.border_mode, params.border_value, tag=k + 1)
        if len(items) - dropped <= 0:
            raise InvalidParams("All images discarded: try modifying KeyPointMatchParameters::match_distance_threshold")
        out = st.finish(len(items) - dropped)
        ok = True
        return dropped, out
    finally:
        _release_stack(st, cache_key, ok)


# ---- sharpness_tenengrad: src/lib.rs:1101-1147 ----------------------------------------------------------
def sharpness_tenengrad(src_grey_mat, k_size: int, *, device: int = -1) -> float:
    if k_size not in (1, 3, 5, 7):
        raise InvalidParams("Kernel size must be 1, 3, 5, or 7")
    out = C.c_double()
    dv = _device_view(src_grey_mat)
    if dv is not None:
        shape = src_grey_mat.__cuda_array_interface__["shape"]
        ch = 1 if len(shape) == 2 else shape[2]
        _check(lib.stk_tenengrad_device(dv[0], dv[1], shape[1], shape[0], ch, k_size, device, C.byref(out)))
        return out.value
    a = np.asarray(src_grey_mat)
    if a.dtype != np.uint8:
        raise NotImplementedError_("sharpness_tenengrad: only 8-bit input is implemented on the GPU path")
    if a.ndim != 2:
        raise OpenCvError("sharpness_tenengrad expects a single-channel image")
    if a.strides[1] != 1:
        a = np.ascontiguousarray(a)
    _check(lib.stk_tenengrad(a.ctypes.data, a.strides[0], a.shape[1], a.shape[0], 1, k_size, device, C.byref(out)))
    return out.value


# ---- the other sharpness metrics: src/lib.rs:1032-1090, :1151-1166 ---------------------------------------
def sharpness_all(src_mat, *, device: int = -1):
    """(LAPM, LAPV, TENG(3), GLVN) of one 8-bit grey image (host array, or anything with
    __cuda_array_interface__) in ONE pass over the plane — what examples/main.rs:43-46 computes per file."""
    out = (C.c_double * 4)()
    dv = _device_view(src_mat)
    if dv is not None:
        shape = src_mat.__cuda_array_interface__["shape"]
        ch = 1 if len(shape) == 2 else shape[2]
        _check(lib.stk_sharpness_all_batch_device(dv[0], 0, dv[1], shape[1], shape[0], ch, 1, device, out))
        return tuple(out)
    a = np.asarray(src_mat)
    if a.dtype != np.uint8:
        raise NotImplementedError_("sharpness metrics: only 8-bit input is implemented on the GPU path")
    if a.ndim != 2:
        raise OpenCvError("the sharpness metrics expect a single-channel image")
    if a.strides[1] != 1:
        a = np.ascontiguousarray(a)
    _check(lib.stk_sharpness_all(a.ctypes.data, a.strides[0], a.shape[1], a.shape[0], 1, device, out))
    return tuple(out)


def sharpness_batch(frames, *, device: int = -1):
    """n same-sized frames resident on the device as one (n, H, W[, C]) uint8 array (e.g. a torch CUDA
    tensor; C = 3/4 fuses cvtColor(BGR2GRAY)): an (n, 4) array of (LAPM, LAPV, TENG(3), GLVN)."""
    iface = frames.__cuda_array_interface__
    shape = iface["shape"]
    if iface.get("strides"):
        raise InvalidParams("sharpness_batch needs a contiguous (n, H, W[, C]) array")
    n, h, w = shape[0], shape[1], shape[2]
    ch = 1 if len(shape) == 3 else shape[3]
    out = (C.c_double * (4 * n))()
    _check(lib.stk_sharpness_all_batch_device(int(iface["data"][0]), h * w * ch, w * ch, w, h, ch, n, device, out))
    return np.array(out[:], np.float64).reshape(n, 4)


def sharpness_modified_laplacian(src_mat, *, device: int = -1) -> float:
    """LAPM (Nayar89), src/lib.rs:1032-1068."""
    return sharpness_all(src_mat, device=device)[0]


def sharpness_variance_of_laplacian(src_mat, *, device: int = -1) -> float:
    """LAPV (Pech2000), src/lib.rs:1070-1090."""
    return sharpness_all(src_mat, device=device)[1]


def sharpness_normalized_gray_level_variance(src_mat, *, device: int = -1) -> float:
    """GLVN (Santos97), src/lib.rs:1151-1166."""
    return sharpness_all(src_mat, device=device)[3]
