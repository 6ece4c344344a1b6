"""Host-side logic and the C-ABI surface, no GPU needed."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_term_criteria_doctest(pkg):
    """The reference's only executed test on this path (/root/reference/src/utils.rs:148-158):
    max_count None, epsilon Some(0.1) -> epsilon == 0.1, typ == EPS."""
    p = pkg.EccMatchParameters(pkg.MotionType.Euclidean, None, 0.1, 3)
    typ, mc, eps = pkg.term_criteria(p)
    assert eps == 0.1 and typ == 2 and mc == 0
    assert pkg.term_criteria(pkg.EccMatchParameters(pkg.MotionType.Affine, 5000, 1e-5, 5)) == (3, 5000, 1e-5)
    assert pkg.term_criteria(pkg.EccMatchParameters(pkg.MotionType.Affine, 7, None, 5)) == (1, 7, 0.0)
    assert pkg.term_criteria(pkg.EccMatchParameters(pkg.MotionType.Affine, None, None, 5))[0] == 0


def test_motion_type_discriminants(pkg):
    # opencv::video::MOTION_* (/root/reference/src/lib.rs:603-609)
    assert [int(pkg.MotionType.Translation), int(pkg.MotionType.Euclidean), int(pkg.MotionType.Affine),
            int(pkg.MotionType.Homography)] == [0, 1, 2, 3]


def test_keypoint_defaults(pkg):
    # impl Default (/root/reference/src/utils.rs:250-261)
    d = pkg.KeyPointMatchParameters()
    assert (d.method, d.ransac_reproj_threshold, d.match_keep_ratio, d.match_ratio, d.border_mode) == (8, 3.0, 0.75, 0.8, 0)
    assert tuple(d.border_value) == (0.0, 0.0, 0.0, 0.0)


def test_abi_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "stacker_cuda.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(stk_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = ctypes.CDLL(pkg._ffi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/stacker_cuda.h but not exported"
    assert declared == set(pkg._ffi.SYMBOLS), declared ^ set(pkg._ffi.SYMBOLS)
    assert lib.stk_abi_version() == 5


def test_struct_layout_matches_header(pkg):
    # stk_ecc_config: 6 int32, double, 7 int32 -> 8-byte aligned; stk_frame_result: i64, 9 f32, f64, 2 i32
    assert ctypes.sizeof(pkg._ffi.EccConfig) == 64
    assert pkg._ffi.EccConfig.ecc_width.offset == 52
    assert pkg._ffi.EccConfig.epsilon.offset == 24
    assert ctypes.sizeof(pkg._ffi.FrameResult) == 64
    assert pkg._ffi.FrameResult.rho.offset == 48


def test_errors_before_any_gpu_work(pkg):
    with pytest.raises(pkg.NotEnoughFiles):
        pkg.ecc_match([], pkg.EccMatchParameters(pkg.MotionType.Homography, 10, 1e-3, 5))
    with pytest.raises(pkg.NotEnoughFiles):
        pkg.keypoint_match([])
    with pytest.raises(pkg.InvalidParams):
        pkg.sharpness_tenengrad(np.zeros((8, 8), np.uint8), 2)
    with pytest.raises(pkg.StackerError):
        pkg.ecc_match(["/nonexistent/a.jpg"], pkg.EccMatchParameters(pkg.MotionType.Homography, 10, 1e-3, 5))


def test_scaled_size_rule_and_validation(pkg):
    """utils::scale_image's size rule (src/utils.rs:186-200) and ecc_match_scaling_down's validation
    (src/lib.rs:876-888) live in the library (no GPU needed); the oracle restates the same rule."""
    from oracle import restate as R
    for w, h, sd in [(1024, 768, 300.0), (3840, 2160, 540.0), (150, 200, 70.0), (301, 201, 67.0), (333, 222, 221.5),
                     (6000, 4000, 1234.5)]:
        assert pkg.scaled_size(w, h, sd) == R.scaled_size(w, h, sd)
    with pytest.raises(pkg.InvalidParams, match="larger"):
        pkg.scaled_size(1024, 768, 1024.0)
    with pytest.raises(pkg.InvalidParams, match="too small"):
        pkg.scaled_size(1024, 768, 10.0)
    # landscape frame, height < scale_down < width: the rule ENLARGES (cv::resize INTER_AREA then runs its bilinear
    # "area mode"); legal in the reference, so legal here
    assert pkg.scaled_size(1000, 500, 800.0) == R.scaled_size(1000, 500, 800.0) == (1600, 800)
    # the validation runs before any device work
    frames = [np.zeros((96, 128, 3), np.uint8)] * 2
    with pytest.raises(pkg.InvalidParams):
        pkg.ecc_match(frames, pkg.EccMatchParameters(pkg.MotionType.Affine, 5, 1e-3, 3), 500.0)


def test_no_cpu_fallback_without_gpu(pkg):
    """On a box without a CUDA device the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.ProcessingError):
        pkg.sharpness_tenengrad(np.zeros((16, 16), np.uint8), 3)
    with pytest.raises(pkg.ProcessingError):
        pkg.sharpness_all(np.zeros((16, 16), np.uint8))
    with pytest.raises(pkg.ProcessingError):
        pkg.grey_resize_area(np.zeros((16, 16), np.uint8), 8, 8)
    frames = [np.zeros((32, 32, 3), np.uint8)] * 2
    with pytest.raises(pkg.ProcessingError):
        pkg.ecc_match(frames, pkg.EccMatchParameters(pkg.MotionType.Affine, 5, 1e-3, 3))


def test_product_package_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "libstacker.rs_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".rs")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("oracle/restate.py", "").replace("oracle/", "oracle_") or f.endswith((".cu", ".cuh")), f


def test_rank_by_sharpness_order():
    from oracle import restate as R
    # examples/main.rs:53-64: ascending sort, drop the worst, reverse -> sharpest first
    assert R.rank_by_sharpness([5.0, 1.0, 9.0, 3.0]) == [2, 0, 3]


def test_keep_count_rounds_half_away_from_zero(pkg):
    """src/lib.rs:235: `(len as f32 * match_keep_ratio).round() as usize` — Rust's f32::round rounds halves away
    from zero.  With the default ratio 0.75 the product lands on .5 whenever len % 4 == 2: 6 -> 5 (Python's
    round() would give 4, below the 5-match minimum, and drop a frame the reference keeps), 14 -> 11, 22 -> 17."""
    from oracle import cvref
    for fn in (pkg.api._keep_count, cvref.keep_count):
        assert [fn(n, 0.75) for n in (6, 14, 22, 4, 5, 7, 0)] == [5, 11, 17, 3, 4, 5, 0]
        assert fn(10, 0.8) == 8 and fn(3, 0.5) == 2 and fn(1, 0.5) == 1


def test_device_frame_validation(pkg):
    """Frames handed over through __cuda_array_interface__ are checked before the library sees the pointer."""
    class Fake:
        def __init__(self, shape, typestr="|u1", strides=None):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (4096, False), "version": 3,
                                             "strides": strides}
    dv = pkg.api._device_view
    assert dv(Fake((4, 8, 3)), (4, 8, 3)) == (4096, 24)
    assert dv(Fake((4, 8, 3), strides=(32, 3, 1)), (4, 8, 3)) == (4096, 32)
    with pytest.raises(pkg.OpenCvError):
        dv(Fake((4, 8, 3), typestr="<f4"), (4, 8, 3))
    with pytest.raises(pkg.OpenCvError):
        dv(Fake((4, 9, 3)), (4, 8, 3))
    with pytest.raises(pkg.OpenCvError):
        dv(Fake((4, 8, 3), strides=(48, 6, 2)), (4, 8, 3))


def test_context_cache_key_and_producer_stream(pkg):
    """Host logic of the context cache and of the stream hand-over (no GPU needed): the key separates everything a
    context is built from; the producer stream comes from __cuda_array_interface__ when the object is not a torch tensor."""
    api = pkg.api
    p1 = pkg.EccMatchParameters(pkg.MotionType.Homography, 5000, 1e-5, 5)
    p2 = pkg.EccMatchParameters(pkg.MotionType.Homography, None, 1e-5, 5)         # EPS only: a different TermCriteria
    k = api._ctx_key(640, 480, 3, p1, 0, None, True, 0)
    assert k == api._ctx_key(640, 480, 3, pkg.EccMatchParameters(pkg.MotionType.Homography, 5000, 1e-5, 5), 0, None, True, 0)
    others = [api._ctx_key(640, 480, 3, p2, 0, None, True, 0), api._ctx_key(640, 480, 4, p1, 0, None, True, 0),
              api._ctx_key(640, 480, 3, p1, 1, None, True, 0), api._ctx_key(640, 480, 3, p1, 0, (320, 240), True, 0),
              api._ctx_key(640, 480, 3, None, 0, None, True, 0), api._ctx_key(640, 480, 3, p1, 0, None, False, 0),
              api._ctx_key(640, 480, 3, p1, 0, None, True, 2)]
    assert len({k, *others}) == len(others) + 1

    class Fake:
        def __init__(self, stream):
            self.__cuda_array_interface__ = {"shape": (2, 2, 3), "typestr": "|u1", "data": (4096, False), "version": 3}
            if stream != "absent":
                self.__cuda_array_interface__["stream"] = stream
    assert api._producer_stream(Fake("absent")) is None        # unknown producer: no ordering is requested
    assert api._producer_stream(Fake(None)) is None
    assert api._producer_stream(Fake(1)) == 0                  # 1 = the legacy default stream
    assert api._producer_stream(Fake(2)) is None               # per-thread default stream: not expressible as a handle
    assert api._producer_stream(Fake(0x7f00dead)) == 0x7f00dead


def test_context_cache_evicts_least_recently_parked(pkg, monkeypatch):
    """Bookkeeping of the context cache with stand-in contexts (no GPU): at most STK_CONTEXT_CACHE idle contexts, the
    least recently parked one is destroyed first, a context that raised is destroyed at once."""
    api = pkg.api

    class Stand:
        def __init__(self, name):
            self.name, self._keep, self.closed = name, [1], False

        def close(self):
            self.closed = True

        def reset(self):
            pass

    api.clear_context_cache()
    monkeypatch.setattr(api, "_CTX_CACHE_MAX", 2)
    a, b, c, d = Stand("a"), Stand("b"), Stand("c"), Stand("d")
    api._release_stack(a, "k1", True)
    api._release_stack(b, "k2", True)
    api._release_stack(c, "k1", True)                # third idle context: the oldest (a) goes
    assert (a.closed, b.closed, c.closed) == (True, False, False) and c._keep == []
    api._release_stack(d, "k3", False)               # a context whose call raised is never parked
    assert d.closed and [s.name for _, s in api._CTX_CACHE_ORDER] == ["b", "c"]
    api.clear_context_cache()
    assert b.closed and c.closed and not api._CTX_CACHE and not api._CTX_CACHE_ORDER
