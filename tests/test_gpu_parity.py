"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle (oracle/restate.py,
itself pinned to cv2) and against cv2 directly where it is installed.

Bars (BASELINE.json north_star): warp matrices within 0.05 px corner displacement, 8-bit stack
max-abs-diff <= 1 (PSNR >= 50 dB), identical sharpness ordering; byte/integer/sampling work bit-exact."""
import math

import numpy as np
import pytest

from oracle import restate as R
from oracle import synth
from parity_util import assert_stack_parity, psnr8

pytestmark = pytest.mark.gpu

MOTIONS = [0, 1, 2, 3]


# ---- K1: grey + blur, bit-exact ------------------------------------------------------------------------
# widths that are multiples of 4 take the streaming kernel for k = 3 / 5 (bands of 120 columns, strips of rows):
# several bands, a last band that ends mid-warp, a single partial band, the smallest legal plane, 3-row planes
@pytest.mark.parametrize("size", [(320, 240), (257, 131), (64, 40), (1000, 37), (3840, 40), (244, 20), (124, 33),
                                  (120, 9), (8, 5), (480, 3)])
@pytest.mark.parametrize("k", [1, 3, 5, 7, 9])
def test_prep_bit_exact(pkg, size, k):
    w, h = size
    rng = np.random.default_rng(w * 31 + k)
    frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    got = pkg.prep_grey_blur(frame, k, device=0)
    want = R.gaussian_blur_f32(R.bgr2gray_u8(frame).astype(np.float32), k)
    assert np.array_equal(got, want)


def test_prep_large_kernel_close(pkg):
    rng = np.random.default_rng(5)
    frame = rng.integers(0, 256, (96, 160, 3), dtype=np.uint8)
    for k in (11, 15, 31):
        got = pkg.prep_grey_blur(frame, k, device=0)
        want = R.gaussian_blur_f32(R.bgr2gray_u8(frame).astype(np.float32), k)
        assert np.abs(got - want).max() <= 1e-4


@pytest.mark.parametrize("size", [(70, 50), (72, 50), (368, 31)])
@pytest.mark.parametrize("k", [3, 5])
def test_prep_bgra(pkg, size, k):
    w, h = size
    rng = np.random.default_rng(6 + w)
    frame = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    got = pkg.prep_grey_blur(frame, k, device=0)
    want = R.gaussian_blur_f32(R.bgr2gray_u8(frame[..., :3]).astype(np.float32), k)
    assert np.array_equal(got, want)


def test_prep_extreme_values(pkg):
    """all-255 and checkerboard planes: the packed 16-bit halves of the streaming kernel reach their maximum
    (256 * 255) without carrying into each other"""
    for k in (3, 5):
        for frame in (np.full((24, 256, 3), 255, np.uint8),
                      np.tile(np.array([[0, 255], [255, 0]], np.uint8), (12, 128))[..., None].repeat(3, axis=2)):
            got = pkg.prep_grey_blur(np.ascontiguousarray(frame), k, device=0)
            want = R.gaussian_blur_f32(R.bgr2gray_u8(frame).astype(np.float32), k)
            assert np.array_equal(got, want)


# ---- K4: final warp + accumulate, bit-exact ------------------------------------------------------------
def _rand_h(rng, w, h, big=False):
    g = synth.random_warp(rng, 3, w, h)
    if big:
        g[:2, 2] += rng.uniform(-0.4, 0.4, 2) * (w, h)
        g[:2, :2] += rng.uniform(-0.1, 0.1, (2, 2))
    return g


@pytest.mark.parametrize("size", [(320, 240), (333, 97), (64, 64), (31, 200)])
@pytest.mark.parametrize("big", [False, True])
def test_warp_only_matches_oracle_bit_exact(pkg, size, big):
    """keypoint_match tail: frame 0 unwarped + warpPerspective(frame_i, H_i f64) summed, / n."""
    w, h = size
    rng = np.random.default_rng(w + h + big)
    frames = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(4)]
    hs = [_rand_h(rng, w, h, big) for _ in range(3)]
    with pkg.EccStack(w, h, 3, None, device=0, lanes=1) as st:     # one lane: fixed summation order
        st.set_reference(frames[0])
        for f, hm in zip(frames[1:], hs):
            st.submit_warp(f, hm)
        got = st.finish(4)
    acc = R.to_f32_unit(frames[0])
    for f, hm in zip(frames[1:], hs):
        acc = acc + R.warp_linear(R.to_f32_unit(f), hm, w, h, perspective=True, inverse_map=False)
    want = acc * np.float32(1.0 / 4)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("chan", [3, 4])
def test_warp_only_unaligned_device_frames_and_strong_perspective(pkg, have_cv2, chan):
    """The rarely-taken paths of the second-generation K4: device frames whose base address and pitch are NOT 4-byte
    aligned (byte-granular word addressing), a homography whose w runs from ~0.3 to ~3 over the frame (columns outside
    the guarded f32 range evaluate every pixel in f64) and one that throws most pixels out of the source — all
    `array_equal` with cv2.warpPerspective on the CV_32F frame."""
    import cv2
    import torch
    w, h = 401, 302                                   # pitch 401 * chan: odd for chan = 3, and a partial tile on both edges
    rng = np.random.default_rng(77 + chan)
    frames = [rng.integers(0, 256, (h, w, chan), dtype=np.uint8) for _ in range(4)]
    hs = [_rand_h(rng, w, h, True),
          # inverse map with w falling from 1 to 0.16 across the frame: columns below 1/4 leave the guarded f32 range
          np.linalg.inv(np.array([[1.1, 0.0, 5.0], [0.0, 1.1, -4.0], [-2.1e-3, 0.0, 1.0]])),
          # inverse map whose w changes sign inside the frame (cv2: W ? 32 / W : 0, saturating casts): general path
          np.array([[1.0, 0.02, -3.0], [0.01, 1.0, 2.0], [4.0e-3, -1.5e-3, 1.0]])]
    k255 = np.float32(1.0 / 255.0)
    want = frames[0].astype(np.float32) * k255
    for f, hm in zip(frames[1:], hs):
        want = want + cv2.warpPerspective(f.astype(np.float32) * k255, hm, (w, h), flags=cv2.INTER_LINEAR)
    want = want * np.float32(1.0 / 4)
    # device copies at an odd byte offset inside a larger buffer (base % 4 == 1 or 3)
    devs = []
    for k, f in enumerate(frames):
        raw = torch.zeros(f.size + 8, dtype=torch.uint8, device="cuda:0")
        off = 1 + 2 * (k % 2)
        view = raw[off:off + f.size].view(h, w, chan)
        view.copy_(torch.from_numpy(f))
        assert view.data_ptr() % 4 != 0
        devs.append(view)
    with pkg.EccStack(w, h, chan, None, device=0, lanes=1) as st:
        st.set_reference(devs[0])
        for d, hm in zip(devs[1:], hs):
            st.submit_warp(d, hm)
        got = st.finish(4)
    assert np.array_equal(got, want)
    # the same through host frames (aligned staging: the 32-bit word-index form)
    with pkg.EccStack(w, h, chan, None, device=0, lanes=1) as st:
        st.set_reference(frames[0])
        for f, hm in zip(frames[1:], hs):
            st.submit_warp(f, hm)
        got = st.finish(4)
    assert np.array_equal(got, want)


def test_warp_only_border_value_and_bgra(pkg):
    w, h = 120, 90
    rng = np.random.default_rng(3)
    frames = [rng.integers(0, 256, (h, w, 4), dtype=np.uint8) for _ in range(2)]
    hm = _rand_h(rng, w, h, True)
    bv = (0.25, 0.5, 0.75, 1.0)
    with pkg.EccStack(w, h, 4, None, device=0, lanes=1) as st:
        st.set_reference(frames[0])
        st.submit_warp(frames[1], hm, pkg.BORDER_CONSTANT, bv)
        got = st.finish(2)
    want = (R.to_f32_unit(frames[0]) +
            R.warp_linear(R.to_f32_unit(frames[1]), hm, w, h, True, False, border_value=np.array(bv))) * np.float32(0.5)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("chan", [3, 4])
def test_warp_only_border_modes_bit_exact(pkg, mode, chan):
    """keypoint_match's border_mode (src/lib.rs:297): REPLICATE / REFLECT / WRAP / REFLECT_101 against the oracle
    (itself pinned bit-exact to cv2.warpPerspective in tests/test_oracle_vs_cv2.py)."""
    w, h = 200, 130
    rng = np.random.default_rng(70 + mode)
    frames = [rng.integers(0, 256, (h, w, chan), dtype=np.uint8) for _ in range(4)]
    hs = [_rand_h(rng, w, h, True) for _ in range(3)]
    hs[2][:2, 2] += np.array([2.5 * w, -1.7 * h])            # far outside: the reflect loop runs more than once
    with pkg.EccStack(w, h, chan, None, device=0, lanes=1) as st:
        st.set_reference(frames[0])
        for f, hm in zip(frames[1:], hs):
            st.submit_warp(f, hm, mode)
        got = st.finish(4)
    acc = R.to_f32_unit(frames[0])
    for f, hm in zip(frames[1:], hs):
        acc = acc + R.warp_linear(R.to_f32_unit(f), hm, w, h, True, False, border_mode=mode)
    assert np.array_equal(got, acc * np.float32(0.25))


@pytest.mark.parametrize("size", [(320, 240), (333, 97), (1921, 1083)])
@pytest.mark.parametrize("chan", [3, 4])
def test_warp_affine_bit_exact_vs_oracle_and_cv2(pkg, have_cv2, size, chan):
    """The 2x3 final warp of ecc_match (warp_affine, src/lib.rs:782-790; OpenCV's 10-bit fixed-point coordinates)
    through its own entry point: `array_equal` against the oracle and against cv2.warpAffine on the CV_32F frame."""
    w, h = size
    rng = np.random.default_rng(w * 7 + chan)
    n = 5
    frames = [rng.integers(0, 256, (h, w, chan), dtype=np.uint8) for _ in range(n)]
    ms = []
    for i in range(n - 1):
        g = synth.random_warp(rng, 2, w, h)[:2].copy()
        if i % 2:                                           # large motions: rims, taps outside, whole rows of border
            g[:, 2] += rng.uniform(-0.3, 0.3, 2) * (w, h)
            g[:, :2] += rng.uniform(-0.1, 0.1, (2, 2))
        ms.append(g)
    bv = (0.0, 0.0, 0.0, 0.0) if chan == 3 else (0.1, 0.2, 0.3, 0.4)
    with pkg.EccStack(w, h, chan, None, device=0, lanes=1) as st:     # one lane: fixed summation order
        st.set_reference(frames[0])
        for f, m in zip(frames[1:], ms):
            st.submit_warp_affine(f, m, pkg.BORDER_CONSTANT, bv)
        got = st.finish(n)
    acc = R.to_f32_unit(frames[0])
    for f, m in zip(frames[1:], ms):
        acc = acc + R.warp_linear(R.to_f32_unit(f), m, w, h, perspective=False, inverse_map=False, border_value=np.array(bv[:chan]))
    assert np.array_equal(got, acc * np.float32(1.0 / n))
    if have_cv2:
        import cv2
        k255 = np.float32(1 / 255.0)
        acc = frames[0].astype(np.float32) * k255
        for f, m in zip(frames[1:], ms):
            acc = acc + cv2.warpAffine(f.astype(np.float32) * k255, m, (w, h), flags=cv2.INTER_LINEAR,
                                       borderMode=cv2.BORDER_CONSTANT, borderValue=bv)
        assert np.array_equal(got, acc * np.float32(1.0 / n))


def test_warp_only_transparent_border_unsupported(pkg):
    frame = np.zeros((16, 16, 3), np.uint8)
    with pkg.EccStack(16, 16, 3, None, device=0, lanes=1) as st:
        st.set_reference(frame)
        with pytest.raises(pkg.StackerError):
            st.submit_warp(frame, np.eye(3), 5)


def test_warp_only_vs_cv2(pkg, have_cv2):
    if not have_cv2:
        pytest.skip("cv2 not installed")
    import cv2
    w, h = 400, 300
    rng = np.random.default_rng(9)
    frames = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(3)]
    hs = [_rand_h(rng, w, h, i == 1) for i in range(2)]
    with pkg.EccStack(w, h, 3, None, device=0, lanes=1) as st:
        st.set_reference(frames[0])
        for f, hm in zip(frames[1:], hs):
            st.submit_warp(f, hm)
        got = st.finish(3)
    acc = frames[0].astype(np.float32) * np.float32(1 / 255.0)
    for f, hm in zip(frames[1:], hs):
        acc = acc + cv2.warpPerspective(f.astype(np.float32) * np.float32(1 / 255.0), hm, (w, h), flags=cv2.INTER_LINEAR)
    assert np.array_equal(got, acc * np.float32(1.0 / 3))


# ---- K2: one iteration's reductions against the oracle ---------------------------------------------------
def _expected_totals(motion, tmpl, img, m32):
    """The NV sums in the kernel's layout (csrc/ecc_iter.cuh Layout<>), from the oracle's planes in f64."""
    hs, ws = tmpl.shape
    persp = motion == 3
    gxp, gyp = R.central_gradients(img)
    w_ = R.warp_linear(img, m32, ws, hs, persp, True).astype(np.float64)
    gx = R.warp_linear(gxp, m32, ws, hs, persp, True).astype(np.float64)
    gy = R.warp_linear(gyp, m32, ws, hs, persp, True).astype(np.float64)
    mk = R.warp_mask_nearest(m32, ws, hs, img.shape[1], img.shape[0], persp).astype(np.float64)
    t_ = tmpl.astype(np.float64)
    X = np.broadcast_to(np.arange(ws, dtype=np.float64)[None, :], (hs, ws))
    Y = np.broadcast_to(np.arange(hs, dtype=np.float64)[:, None], (hs, ws))
    m = np.asarray(m32, np.float64)
    if motion == 0:
        g, kron = [gx, gy], False
    elif motion == 1:
        c, s = m[0, 0], m[1, 0]
        g, kron = [gx * (-X * s - Y * c) + gy * (X * c - Y * s), gx, gy], False
    elif motion == 2:
        g, kron = [gx, gy], True
    else:
        den = X * m[2, 0] + Y * m[2, 1] + 1.0
        hx = -(X * m[0, 0] + Y * m[0, 1] + m[0, 2]) / den
        hy = -(X * m[1, 0] + Y * m[1, 1] + m[1, 2]) / den
        a, b = gx / den, gy / den
        g, kron = [a, b, hx * a + hy * b], True
    out = [mk.sum(), (mk * w_).sum(), (mk * w_ * w_).sum(), (mk * t_).sum(), (mk * t_ * t_).sum(), (mk * w_ * t_).sum()]
    qm = [np.ones_like(X), X, Y, X * X, X * Y, Y * Y] if kron else [np.ones_like(X)]
    zm = [np.ones_like(X), X, Y] if kron else [np.ones_like(X)]
    for i in range(len(g)):
        for j in range(i, len(g)):
            out += [(g[i] * g[j] * q).sum() for q in qm]
    for z in (w_, mk, mk * t_):
        for gi in g:
            out += [(gi * z * q).sum() for q in zm]
    return np.array(out)


@pytest.mark.parametrize("motion", MOTIONS)
@pytest.mark.parametrize("size", [(320, 240), (389, 211)])
def test_iteration_sums_match_oracle(pkg, motion, size):
    w, h = size
    st_ = synth.Stack(w, h, 2, motion, seed=20 + motion)
    f0, f1 = st_.frames()
    # a matrix near (not at) the truth, so the warp leaves the frame on one side
    rng = np.random.default_rng(motion)
    g = st_.truth[1].copy()
    g[:2, 2] += rng.uniform(-1.5, 1.5, 2)
    m32 = g.astype(np.float32)
    if motion != 3:
        m32[2] = (0, 0, 1)
    if motion == 1:   # keep it a rotation
        th = math.asin(float(m32[1, 0]))
        m32[0, 0] = m32[1, 1] = np.float32(math.cos(th)); m32[0, 1] = -m32[1, 0]
    params = pkg.EccMatchParameters(pkg.MotionType(motion), 50, 1e-5, 5)
    with pkg.EccStack(w, h, 3, params, device=0, lanes=1) as st:
        st.set_reference(f0)
        tot, m_out, rho, status = st.debug_iteration(f1, m32)
    tmpl = R.gaussian_blur_f32(R.bgr2gray_u8(f1).astype(np.float32), 5)
    img = R.gaussian_blur_f32(R.bgr2gray_u8(f0).astype(np.float32), 5)
    mm = m32 if motion == 3 else m32[:2]
    want = _expected_totals(motion, tmpl, img, mm)
    assert tot.shape == want.shape
    scale = np.maximum(np.abs(want), 1e-3 * np.abs(want).max() if motion < 2 else 0) + 1e-12
    rel = np.abs(tot - want) / np.maximum(np.abs(want), 1e-30)
    assert tot[0] == want[0], (tot[0], want[0])       # the masked pixel count is exact
    big = np.abs(want) > 1e-6 * np.abs(want).max()
    assert rel[big].max() < 2e-4, (np.argmax(rel * big), rel[big].max())
    # and the update the kernel derives from its own sums equals the oracle's epilogue on the oracle's sums
    sums = R.ecc_sums(motion, tmpl, img, *R.central_gradients(img), mm)
    rho_want, m_want = R.ecc_epilogue(motion, sums, mm)
    assert status == 0
    assert abs(rho - rho_want) < 1e-6
    assert synth.corner_displacement(m_out if motion == 3 else m_out[:2], m_want, w, h) < 2e-3


# ---- the whole path: ecc_match vs oracle / cv2 -----------------------------------------------------------
@pytest.mark.parametrize("motion", MOTIONS)
def test_ecc_match_small_vs_oracle(pkg, motion):
    w, h = 320, 240
    stack = synth.Stack(w, h, 6, motion, seed=47 + motion)
    frames = stack.frames()
    params = pkg.EccMatchParameters(pkg.MotionType(motion), 5000, 1e-5, 5)
    got, res = pkg.ecc_match(frames, params, None, device=0, return_details=True)
    want, warps, iters = R.ecc_match(frames, motion, 5000, 1e-5, 5)
    assert [r["status"] for r in res] == [0] * 5
    for r, wm, it in zip(res, warps[1:], iters[1:]):
        mine = r["warp"] if motion == 3 else r["warp"][:2]
        assert synth.corner_displacement(mine, wm, w, h) <= 0.05
        # the eps test on rho may trip an iteration or three apart; the matrix and stack bars above decide
        assert it > 40 or abs(r["iterations"] - it) <= 4
    assert_stack_parity(got, want, warps, motion, 6)


def test_ecc_match_config1_vs_cv2(pkg, have_cv2):
    """BASELINE config 1: Homography, 5 x 1024x768, max_count 5000, eps 1e-5, gauss 5 (examples/main.rs:105-114)."""
    if not have_cv2:
        pytest.skip("cv2 not installed")
    from oracle import cvref
    stack = synth.config_stack(1)
    frames = stack.frames()
    params = pkg.EccMatchParameters(pkg.MotionType.Homography, 5000, 1e-5, 5)
    got, res = pkg.ecc_match(frames, params, None, device=0, return_details=True)
    want, warps, _ = cvref.ecc_match(frames, 3, 5000, 1e-5, 5)
    for r, wm in zip(res, warps[1:]):
        assert synth.corner_displacement(r["warp"], wm, 1024, 768) <= 0.05
    g8, w8 = np.rint(got * 255.0), np.rint(want * 255.0)
    assert np.abs(g8 - w8).max() <= 1
    assert psnr8(g8, w8) >= 50.0


def test_ecc_match_config2_shape_vs_cv2(pkg, have_cv2):
    """BASELINE config 2 (Euclidean 1920x1080) on 5 of its 16 frames (all 16: tests/test_gpu_fullsize.py)."""
    if not have_cv2:
        pytest.skip("cv2 not installed")
    from oracle import cvref
    frames = synth.config_stack(2, n_frames=5).frames()
    params = pkg.EccMatchParameters(pkg.MotionType.Euclidean, 5000, 1e-5, 5)
    got, res = pkg.ecc_match(frames, params, None, device=0, return_details=True)
    want, warps, _ = cvref.ecc_match(frames, 1, 5000, 1e-5, 5)
    for r, wm in zip(res, warps[1:]):
        assert synth.corner_displacement(r["warp"][:2], wm, 1920, 1080) <= 0.05
    assert_stack_parity(got, want, warps, 1, 5)


def test_fixed_iteration_count_and_eps_only(pkg):
    """TermCriteria flag semantics (src/utils.rs:159-170): COUNT only -> exactly max_count iterations;
    EPS only -> OpenCV's 200-iteration cap with the eps test."""
    w, h = 256, 192
    frames = synth.Stack(w, h, 2, 2, seed=77).frames()
    _, res = pkg.ecc_match(frames, pkg.EccMatchParameters(pkg.MotionType.Affine, 7, None, 3), None, device=0, return_details=True)
    assert res[0]["iterations"] == 7
    _, res = pkg.ecc_match(frames, pkg.EccMatchParameters(pkg.MotionType.Affine, None, 1e-3, 3), None, device=0, return_details=True)
    _, _, it = R.find_transform_ecc(R.bgr2gray_u8(frames[1]), R.bgr2gray_u8(frames[0]), 2, R.term_criteria(None, 1e-3), 3)
    assert abs(res[0]["iterations"] - it) <= 1 and res[0]["iterations"] < 200


def test_ecc_errors(pkg):
    w, h = 128, 96
    frames = synth.Stack(w, h, 2, 0, seed=5).frames()
    with pytest.raises(pkg.NotEnoughFiles):
        pkg.ecc_match([], pkg.EccMatchParameters(pkg.MotionType.Affine, 10, 1e-4, 5))
    with pytest.raises(pkg.OpenCvError):      # neither COUNT nor EPS: CV_Assert inside findTransformECC
        pkg.ecc_match(frames, pkg.EccMatchParameters(pkg.MotionType.Affine, None, None, 5), device=0)
    # uncorrelated images: OpenCV raises StsNoConv -> the whole call fails (src/lib.rs:777)
    rng = np.random.default_rng(0)
    flat = np.full((h, w, 3), 7, np.uint8)
    with pytest.raises(pkg.OpenCvError):
        pkg.ecc_match([frames[0], flat], pkg.EccMatchParameters(pkg.MotionType.Translation, 20, 1e-4, 5), device=0)
    # a single frame is legal: the stack is frame 0 itself
    out = pkg.ecc_match(frames[:1], pkg.EccMatchParameters(pkg.MotionType.Translation, 20, 1e-4, 5), device=0)
    assert np.array_equal(out, R.to_f32_unit(frames[0]))


def test_context_cache_reuse_and_failure(pkg):
    """The one-shot plugin calls park their context and reuse it for the next call of the same geometry: identical
    results, a separate context per parameter set, a context whose call raised is destroyed rather than reused."""
    api = pkg.api
    w, h = 256, 192
    frames = synth.Stack(w, h, 5, 3, seed=12).frames()
    p1 = pkg.EccMatchParameters(pkg.MotionType.Homography, 5000, 1e-5, 5)
    p2 = pkg.EccMatchParameters(pkg.MotionType.Affine, 50, 1e-4, 3)
    pkg.clear_context_cache()
    a, ra = pkg.ecc_match(frames, p1, None, device=0, return_details=True)
    assert sum(len(v) for v in api._CTX_CACHE.values()) == 1
    b, rb = pkg.ecc_match(frames, p1, None, device=0, return_details=True)
    assert np.array_equal(a, b) and all(np.array_equal(x["warp"], y["warp"]) for x, y in zip(ra, rb))
    assert sum(len(v) for v in api._CTX_CACHE.values()) == 1          # the same context went back
    c = pkg.ecc_match(frames, p2, None, device=0)
    assert len(api._CTX_CACHE) == 2 and not np.array_equal(a, c)
    flat = np.full((h, w, 3), 7, np.uint8)
    with pytest.raises(pkg.OpenCvError):
        pkg.ecc_match([frames[0], flat], p1, None, device=0)
    assert sum(len(v) for v in api._CTX_CACHE.values()) == 1          # p1's context was destroyed, p2's is still parked
    d = pkg.ecc_match(frames, p1, None, device=0)
    assert np.array_equal(a, d)
    pkg.clear_context_cache()
    assert not api._CTX_CACHE


def test_device_frames_are_ordered_behind_their_producer_stream(pkg):
    """ABI v5 stream contract: a device frame whose producer has only QUEUED its writes (a side stream that first sleeps,
    then copies) may be submitted at once — the lane waits for the producer's stream on the device."""
    import torch
    w, h = 320, 240
    frames = synth.Stack(w, h, 5, 3, seed=21).frames()
    params = pkg.EccMatchParameters(pkg.MotionType.Homography, 5000, 1e-5, 5)
    want, rw = pkg.ecc_match(frames, params, None, device=0, return_details=True)
    host = [torch.from_numpy(f).pin_memory() for f in frames]
    bufs = [torch.zeros(h, w, 3, dtype=torch.uint8, device="cuda:0") for _ in frames]
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=0)
    with pkg.EccStack(w, h, 3, params, device=0) as st:
        with torch.cuda.stream(side):
            for k, (b, hf) in enumerate(zip(bufs, host)):
                torch.cuda._sleep(10_000_000)          # several ms: the copy below is queued, not done, when submit returns
                b.copy_(hf, non_blocking=True)
                if k == 0:
                    st.set_reference(b)
                else:
                    st.submit(b, tag=k)
        got = st.finish(len(frames))
        res = st.results()
    for a, b in zip(sorted(res, key=lambda r: r["tag"]), rw):
        assert np.array_equal(a["warp"], b["warp"])
    assert np.abs(got - want).max() <= 1e-6


def test_lanes_and_device_resident_input(pkg):
    """Same stack through 1 lane / 4 lanes, host and device-resident frames: identical warps, stack equal
    up to f32 summation order."""
    import torch
    w, h = 320, 240
    frames = synth.Stack(w, h, 6, 3, seed=91).frames()
    params = pkg.EccMatchParameters(pkg.MotionType.Homography, 5000, 1e-5, 5)
    outs, warps = [], []
    for lanes, dev in ((1, False), (4, False), (3, True)):
        with pkg.EccStack(w, h, 3, params, device=0, lanes=lanes) as st:
            src = [torch.from_numpy(f).cuda() for f in frames] if dev else frames
            st.set_reference(src[0])
            for k, f in enumerate(src[1:]):
                st.submit(f, tag=k + 1)
            outs.append(st.finish(len(frames)))
            warps.append([r["warp"] for r in st.results()])
    for ws in warps[1:]:
        for a, b in zip(warps[0], ws):
            assert np.array_equal(a, b)
    for o in outs[1:]:
        assert np.abs(o - outs[0]).max() <= 1e-6


def test_kernel_variants_agree(pkg, monkeypatch):
    """The homography iteration kernel exists in several builds: FastPersp coordinates (default), exact f64
    coordinates (STK_ECC_EXACT_COORDS=1), the compiled block geometries of the second-generation kernel
    (STK_ECC_CFG, csrc/ecc_iter_v2.cuh), the first-generation kernel (STK_ECC_GEN=1) and the host-driven loop
    (STK_LOOP_MODE=host).  Same stack, same bars; the geometries must also agree with each other to rounding."""
    w, h = 400, 300
    frames = synth.Stack(w, h, 5, 3, seed=50).frames()
    want, warps, _ = R.ecc_match(frames, 3, 5000, 1e-5, 5)
    params = pkg.EccMatchParameters(pkg.MotionType.Homography, 5000, 1e-5, 5)
    keys = ("STK_ECC_EXACT_COORDS", "STK_ECC_GEN", "STK_ECC_CFG", "STK_LOOP_MODE")
    envs = [{}, {"STK_ECC_EXACT_COORDS": "1"}, {"STK_ECC_GEN": "1"}, {"STK_LOOP_MODE": "host"}]
    envs += [{"STK_ECC_CFG": str(k)} for k in range(1, 8)]
    first = None
    for env in envs:
        for k in keys:
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        got, res = pkg.ecc_match(frames, params, None, device=0, return_details=True)
        for r, wm in zip(res, warps[1:]):
            assert synth.corner_displacement(r["warp"], wm, w, h) <= 0.05, env
        assert_stack_parity(got, want, warps, 3, 5)
        if "STK_ECC_EXACT_COORDS" not in env:
            if first is None:
                first = res
            for a, b in zip(first, res):
                assert synth.corner_displacement(a["warp"], b["warp"], w, h) <= 2e-3, env


# ---- K6: Tenengrad, bit-identical ------------------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 3, 5, 7])
@pytest.mark.parametrize("size", [(320, 240), (65, 33), (1000, 701)])
def test_tenengrad_exact(pkg, k, size):
    w, h = size
    rng = np.random.default_rng(w + k)
    grey = rng.integers(0, 256, (h, w), dtype=np.uint8)
    assert pkg.sharpness_tenengrad(grey, k, device=0) == R.sharpness_tenengrad(grey, k)


@pytest.mark.parametrize("size", [(16, 2), (32, 3), (512, 49), (528, 97), (2064, 50), (4112, 5), (3840, 2160)])
def test_tenengrad_stream_kernel_exact(pkg, size):
    """Tenengrad(3) on 16-byte aligned grey planes runs the dedicated streaming kernel (16 columns per thread, halo
    by shuffle, 48-row bands): plane edges, warp and block boundaries, band tails of 1 and 2 rows, a batch."""
    import ctypes as C
    import torch
    w, h = size
    rng = np.random.default_rng(w * 3 + h)
    n = 3
    greys = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    greys[1] = 255                                           # saturated plane: zero gradient everywhere
    greys[2, :, ::2] = 0
    greys[2, :, 1::2] = 255                                  # maximum horizontal gradient
    dev = torch.from_numpy(greys).cuda()
    out = (C.c_double * n)()
    assert pkg._ffi.lib.stk_tenengrad_batch_device(dev.data_ptr(), h * w, w, w, h, 1, 3, n, 0, out) == 0
    assert list(out) == [R.sharpness_tenengrad(g, 3) for g in greys]
    assert pkg.sharpness_tenengrad(greys[0], 3, device=0) == out[0]


def test_tenengrad_vs_cv2_and_ordering(pkg, have_cv2):
    """config 3's ranking step: identical sharpness ORDER (ties included) as the reference."""
    if not have_cv2:
        pytest.skip("cv2 not installed")
    from oracle import cvref
    stack = synth.config_stack(3, n_frames=8, width=640, height=360)
    greys = [R.bgr2gray_u8(f) for f in stack.frames()]
    mine = [pkg.sharpness_tenengrad(g, 3, device=0) for g in greys]
    ref = [cvref.sharpness_tenengrad(g, 3) for g in greys]
    assert mine == ref
    assert R.rank_by_sharpness(mine) == R.rank_by_sharpness(ref)


def test_tenengrad_invalid_k(pkg):
    with pytest.raises(pkg.InvalidParams):
        pkg.sharpness_tenengrad(np.zeros((8, 8), np.uint8), 4)


# ---- N1: ecc_match_scaling_down (src/lib.rs:849-1028) ------------------------------------------------------
@pytest.mark.parametrize("case", [(200, 150, 64.0), (256, 192, 96.0), (300, 240, 80.0), (640, 480, 120.0), (301, 201, 67.0),
                                  (150, 200, 70.0), (1024, 768, 300.0), (333, 222, 221.5), (64, 48, 11.0)])
def test_grey_resize_area_bit_exact(pkg, case):
    """cvtColor(BGR2GRAY) + resize(INTER_AREA): generic table path, 2x2 / 3x3 / 4x4 fast paths, portrait."""
    w, h, sd = case
    rng = np.random.default_rng(w * 7 + h)
    sw, sh = R.scaled_size(w, h, sd)
    assert pkg.scaled_size(w, h, sd) == (sw, sh)
    bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    want = R.resize_area_u8(R.bgr2gray_u8(bgr), sw, sh)
    assert np.array_equal(pkg.grey_resize_area(bgr, sw, sh, device=0), want)
    grey = R.bgr2gray_u8(bgr)
    assert np.array_equal(pkg.grey_resize_area(grey, sw, sh, device=0), want)


@pytest.mark.parametrize("case", [(320, 240, 250.0), (512, 384, 400.0), (200, 100, 199.0), (301, 201, 260.0), (333, 250, 250.9),
                                  (1024, 768, 800.0), (64, 48, 50.0)])
def test_grey_resize_area_enlarging_bit_exact(pkg, case):
    """landscape frame, height < scale_down < width: utils::scale_image ENLARGES and cv::resize(INTER_AREA) runs its
    8-bit bilinear kernels in "area mode" (K0's `up` branch) — against the restatement and the cv2 golden vectors."""
    import os
    w, h, sd = case
    rng = np.random.default_rng(w * 7 + h)
    sw, sh = R.scaled_size(w, h, sd)
    assert pkg.scaled_size(w, h, sd) == (sw, sh) and (sw > w or sh > h)
    grey = rng.integers(0, 256, (h, w), dtype=np.uint8)
    want = R.resize_area_u8(grey, sw, sh)
    assert np.array_equal(pkg.grey_resize_area(grey, sw, sh, device=0), want)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scale_up.npz"))
    key = f"resize_{w}x{h}_{int(sd)}"
    if key in g:
        assert np.array_equal(want, g[key])
    bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(pkg.grey_resize_area(bgr, sw, sh, device=0), R.resize_area_u8(R.bgr2gray_u8(bgr), sw, sh))


@pytest.mark.parametrize("case", [(2, 480, 360, 400.0, 66), (3, 400, 300, 330.0, 67)])
def test_ecc_match_scaling_down_enlarging_vs_oracle_and_golden(pkg, case):
    import os
    motion, w, h, sd, seed = case
    frames = synth.Stack(w, h, 5, motion, seed=seed).frames()
    params = pkg.EccMatchParameters(pkg.MotionType(motion), 60, 1e-5, 5)
    got, res = pkg.ecc_match(frames, params, sd, device=0, return_details=True)
    want, warps, _ = R.ecc_match_scaling_down(frames, motion, 60, 1e-5, 5, sd)
    assert [r["status"] for r in res] == [0] * 4
    for r, wm in zip(res, warps[1:]):
        mine = r["warp"] if motion == 3 else r["warp"][:2]
        assert synth.corner_displacement(mine, wm, w, h) <= 0.05
    assert_stack_parity(got, want, warps, motion, 5)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scale_up.npz"))
    for r, ref in zip(res, g[f"sd_m{motion}_warps"]):
        mine = r["warp"] if motion == 3 else r["warp"][:2]
        assert synth.corner_displacement(mine, ref if motion == 3 else ref[:2], w, h) <= 0.05
    assert_stack_parity(got, g[f"sd_m{motion}_stack8"].astype(np.float32) / np.float32(255.0), warps, motion, 5)


def test_grey_resize_area_golden(pkg):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scale_down.npz"))
    for w, h, sd in [(200, 150, 64.0), (256, 192, 96.0), (300, 240, 80.0), (640, 480, 120.0), (301, 201, 67.0), (150, 200, 70.0)]:
        rng = np.random.default_rng(w * 7 + h)
        grey = rng.integers(0, 256, (h, w), dtype=np.uint8)
        want = g[f"resize_{w}x{h}_{int(sd)}"]
        assert np.array_equal(pkg.grey_resize_area(grey, want.shape[1], want.shape[0], device=0), want)


SD_CASES = [(0, 480, 360, 240.0, 61), (1, 480, 360, 240.0, 62), (2, 640, 480, 320.0, 63), (3, 800, 600, 400.0, 65)]


@pytest.mark.parametrize("case", SD_CASES)
def test_ecc_match_scaling_down_vs_oracle_and_golden(pkg, case):
    """ECC on the INTER_AREA-downscaled greys, matrix rescaled on the device with the reference's two f32
    rules, full-size warp: against the restatement and against the committed cv2 vectors."""
    import os
    motion, w, h, sd, seed = case
    frames = synth.Stack(w, h, 5, motion, seed=seed).frames()
    params = pkg.EccMatchParameters(pkg.MotionType(motion), 60, 1e-5, 5)
    got, res = pkg.ecc_match(frames, params, sd, device=0, return_details=True)
    want, warps, iters = R.ecc_match_scaling_down(frames, motion, 60, 1e-5, 5, sd)
    assert [r["status"] for r in res] == [0] * 4
    for r, wm in zip(res, warps[1:]):
        mine = r["warp"] if motion == 3 else r["warp"][:2]
        assert synth.corner_displacement(mine, wm, w, h) <= 0.05
    assert_stack_parity(got, want, warps, motion, 5)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scale_down.npz"))
    for r, ref in zip(res, g[f"sd_m{motion}_warps"]):
        mine = r["warp"] if motion == 3 else r["warp"][:2]
        assert synth.corner_displacement(mine, ref if motion == 3 else ref[:2], w, h) <= 0.05
    assert_stack_parity(got, g[f"sd_m{motion}_stack8"].astype(np.float32) / np.float32(255.0), warps, motion, 5)


def test_ecc_match_scaling_down_config1_vs_cv2(pkg, have_cv2):
    """examples/main.rs:119-128 calls ecc_match with Some(width) on the config-1 stack."""
    if not have_cv2:
        pytest.skip("cv2 not installed")
    from oracle import cvref
    frames = synth.config_stack(1).frames()
    params = pkg.EccMatchParameters(pkg.MotionType.Homography, 5000, 1e-5, 5)
    got, res = pkg.ecc_match(frames, params, 400.0, device=0, return_details=True)
    want, warps, _ = cvref.ecc_match_scaling_down(frames, 3, 5000, 1e-5, 5, 400.0)
    for r, wm in zip(res, warps[1:]):
        assert synth.corner_displacement(r["warp"], wm, 1024, 768) <= 0.05
    g8, w8 = np.rint(got * 255.0), np.rint(want * 255.0)
    assert np.abs(g8 - w8).max() <= 1
    assert psnr8(g8, w8) >= 50.0


def test_ecc_match_scaling_down_errors(pkg):
    frames = synth.Stack(128, 96, 2, 0, seed=5).frames()
    p = pkg.EccMatchParameters(pkg.MotionType.Translation, 20, 1e-4, 5)
    with pytest.raises(pkg.InvalidParams):
        pkg.ecc_match(frames, p, 128.0, device=0)          # >= full width (src/lib.rs:876-881)
    with pytest.raises(pkg.InvalidParams):
        pkg.ecc_match(frames, p, 10.0, device=0)           # too small (src/lib.rs:883-888)
    with pytest.raises(pkg.NotEnoughFiles):
        pkg.ecc_match([], p, 64.0, device=0)


# ---- N3: the other sharpness metrics, bit-identical --------------------------------------------------------
@pytest.mark.parametrize("size", [(320, 240), (65, 33), (1000, 701), (7, 5), (1, 9), (12, 1)])
def test_sharpness_all_exact(pkg, size):
    w, h = size
    rng = np.random.default_rng(w * 3 + h)
    grey = rng.integers(0, 256, (h, w), dtype=np.uint8)
    got = pkg.sharpness_all(grey, device=0)
    want = (R.sharpness_modified_laplacian(grey), R.sharpness_variance_of_laplacian(grey),
            R.sharpness_tenengrad(grey, 3), R.sharpness_normalized_gray_level_variance(grey))
    assert got == want
    assert pkg.sharpness_modified_laplacian(grey, device=0) == want[0]
    assert pkg.sharpness_variance_of_laplacian(grey, device=0) == want[1]
    assert pkg.sharpness_normalized_gray_level_variance(grey, device=0) == want[3]


def test_sharpness_golden_and_batch(pkg):
    import os
    import torch
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sharpness.npz"))
    rng = np.random.default_rng(21)
    greys = {"rand": rng.integers(0, 256, (131, 257), dtype=np.uint8),
             "scene": R.bgr2gray_u8(synth.Stack(320, 240, 1, 0, seed=22).frames()[0])}
    for name, grey in greys.items():
        assert list(pkg.sharpness_all(grey, device=0)) == list(g[name])
    # batch entry point on device-resident BGR frames (grey conversion fused)
    frames = synth.config_stack(3, n_frames=5, width=320, height=180).frames()
    batch = torch.from_numpy(np.stack(frames)).cuda()
    got = pkg.sharpness_batch(batch, device=0)
    for row, f in zip(got, frames):
        grey = R.bgr2gray_u8(f)
        assert tuple(row) == (R.sharpness_modified_laplacian(grey), R.sharpness_variance_of_laplacian(grey),
                              R.sharpness_tenengrad(grey, 3), R.sharpness_normalized_gray_level_variance(grey))
    flat = np.full((30, 40), 9, np.uint8)
    assert pkg.sharpness_all(flat, device=0) == (0.0, 0.0, 0.0, 0.0)


# ---- keypoint_match: host front end (OpenCV) + GPU tail, with and without scale_down_width ----------------
@pytest.mark.parametrize("scale_down", [None, 300.0])
def test_keypoint_match_vs_cv2(pkg, have_cv2, scale_down):
    """src/lib.rs:129-353 and :355-600 against the same OpenCV calls made by the oracle: identical drops,
    stack within the parity bar."""
    if not have_cv2:
        pytest.skip("cv2 not installed")
    from oracle import cvref
    w, h = 800, 600
    frames = synth.Stack(w, h, 5, 3, seed=31).frames()
    params = pkg.KeyPointMatchParameters(method=pkg.RANSAC, ransac_reproj_threshold=5.0, match_keep_ratio=0.8, match_ratio=0.9)
    dropped, got = pkg.keypoint_match(frames, params, scale_down, device=0)
    want_dropped, want, hs = cvref.keypoint_match(frames, reproj=5.0, match_ratio=0.9, keep_ratio=0.8, scale_down=scale_down)
    assert dropped == want_dropped == 0
    g8, w8 = np.rint(got * 255.0), np.rint(want * 255.0)
    assert np.abs(g8 - w8).max() <= 1
    assert psnr8(g8, w8) >= 50.0
    if scale_down is not None:
        with pytest.raises(pkg.InvalidParams):
            pkg.keypoint_match(frames, params, float(w), device=0)


# ---- N2: host feed through the pinned frame ring ------------------------------------------------------------
def test_frame_ring_threads_and_files(pkg, tmp_path):
    """Decode threads fill ring buffers concurrently (more tasks than buffers, so acquire must block and
    recycle), files go through ecc_match's windowed feed: same warps, same stack as in-memory submission."""
    import threading
    import cv2
    w, h, n = 320, 240, 12
    frames = synth.Stack(w, h, n, 2, seed=93).frames()
    params = pkg.EccMatchParameters(pkg.MotionType.Affine, 200, 1e-5, 5)
    want, res_want = pkg.ecc_match(frames, params, None, device=0, return_details=True)
    with pkg.EccStack(w, h, 3, params, device=0, lanes=2) as st:         # ring of 4 buffers, 11 frames, 6 threads
        st.set_reference(frames[0])
        errs = []

        def worker(ids):
            try:
                for i in ids:
                    buf = st.acquire_buffer()
                    np.copyto(buf, frames[i])
                    st.submit_acquired(buf, tag=i)
            except Exception as e:          # pragma: no cover
                errs.append(e)
        ths = [threading.Thread(target=worker, args=(list(range(1 + t, n, 6)),)) for t in range(6)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        assert not errs
        b = st.acquire_buffer()             # an unused buffer can be handed back
        st.release_buffer(b)
        got = st.finish(n)
        res = sorted(st.results(), key=lambda r: r["tag"])
    for a, b in zip(res, sorted(res_want, key=lambda r: r["tag"])):
        assert np.array_equal(a["warp"], b["warp"])
    assert np.abs(got - want).max() <= 1e-6                      # f32 summation order only
    paths = []
    for i, f in enumerate(frames):
        p = str(tmp_path / f"f{i:02d}.png")
        cv2.imwrite(p, f)
        paths.append(p)
    from_files = pkg.ecc_match(paths, params, None, device=0, workers=5)
    assert np.array_equal(from_files, want)                      # same order of submission -> bit-identical
