"""The reference's own numeric engine, driven call-for-call as /root/reference/src/lib.rs drives it.

TEST INFRASTRUCTURE ONLY (checker and CPU baseline; never on the product path).

libstacker.rs does no arithmetic of its own on this path: it calls OpenCV (C++, 4.12.0 pinned,
/root/reference/.github/workflows/rust.yml:53) through the `opencv` crate.  The same OpenCV entry
points are importable here as Python `cv2` (4.13.0 in this image), so this module IS the reference
behaviour for parity purposes: each function makes exactly the calls, in the order and with the
arguments the Rust code makes, on frames shared in memory (decode stays outside).

  read_grey_and_f32   /root/reference/src/utils.rs:128-144
  ecc_match           /root/reference/src/lib.rs:719-847   (ecc_match_no_scaling)
  keypoint_match      /root/reference/src/lib.rs:146-353   (keypoint_match_no_scale)
  sharpness_tenengrad /root/reference/src/lib.rs:1101-1147
  ecc_match_scaling_down  /root/reference/src/lib.rs:849-1028, scale_image /root/reference/src/utils.rs:186-214
  keypoint_match(scale_down=...)  /root/reference/src/lib.rs:355-600 (adjust_homography_for_scale_f64, utils.rs:218-248)
  sharpness_modified_laplacian / _variance_of_laplacian / _normalized_gray_level_variance
                      /root/reference/src/lib.rs:1032-1090, :1151-1166
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import cv2

MOTION_TRANSLATION, MOTION_EUCLIDEAN, MOTION_AFFINE, MOTION_HOMOGRAPHY = 0, 1, 2, 3


def read_grey_and_f32(img_u8: np.ndarray):
    """src/utils.rs:128-144 minus the imread: (grey u8, colour f32 = img * 1/255)."""
    img_f32 = cv2.multiply(img_u8, np.array([1.0]), scale=1.0 / 255.0, dtype=cv2.CV_32F) \
        if False else (img_u8.astype(np.float32) * np.float32(1.0 / 255.0))
    grey = cv2.cvtColor(img_u8, cv2.COLOR_BGR2GRAY)
    return grey, img_f32


def term_criteria(max_count, epsilon):
    """src/utils.rs:159-170."""
    typ, mc, eps = 0, 0, 0.0
    if max_count is not None:
        typ |= cv2.TERM_CRITERIA_COUNT
        mc = int(max_count)
    if epsilon is not None:
        typ |= cv2.TERM_CRITERIA_EPS
        eps = float(epsilon)
    return (typ, mc, eps)


def align_frame(grey_i, grey0, motion, criteria, gauss):
    """src/lib.rs:763-777: identity init, template = frame i, input = frame 0."""
    warp = np.eye(3 if motion == MOTION_HOMOGRAPHY else 2, 3, dtype=np.float32)
    rho, warp = cv2.findTransformECC(grey_i, grey0, warp, motion, criteria, None, gauss)
    return rho, warp


def warp_frame(img_f32, warp, motion, border_mode=cv2.BORDER_CONSTANT, border_value=0):
    """src/lib.rs:780-803."""
    h, w = img_f32.shape[:2]
    if motion == MOTION_HOMOGRAPHY:
        return cv2.warpPerspective(img_f32, warp, (w, h), flags=cv2.INTER_LINEAR,
                                   borderMode=border_mode, borderValue=border_value)
    return cv2.warpAffine(img_f32, warp, (w, h), flags=cv2.INTER_LINEAR,
                          borderMode=border_mode, borderValue=border_value)


def ecc_match(frames_u8, motion, max_count, epsilon, gauss_filt_size, workers=None):
    """ecc_match_no_scaling over in-memory frames.  One task per frame like Rayon's
    into_par_iter().with_min_len(1) (src/lib.rs:746-749); partial sums are combined in index order
    (the reference's order is non-deterministic).
    Returns (stack f32, warps, rhos)."""
    if len(frames_u8) == 0:
        raise ValueError("NotEnoughFiles")
    criteria = term_criteria(max_count, epsilon)
    grey0, f32_0 = read_grey_and_f32(frames_u8[0])

    def task(i):
        if i == 0:
            return f32_0.copy(), None, None
        grey, f32 = read_grey_and_f32(frames_u8[i])
        rho, warp = align_frame(grey, grey0, motion, criteria, gauss_filt_size)
        return warp_frame(f32, warp, motion), warp, rho

    workers = workers or os.cpu_count() or 1
    if workers == 1:
        results = [task(i) for i in range(len(frames_u8))]
    else:
        with ThreadPoolExecutor(max_workers=workers) as ex:
            results = list(ex.map(task, range(len(frames_u8))))
    acc = None
    for warped, _, _ in results:
        acc = warped if acc is None else acc + warped
    stack = acc * np.float32(1.0 / len(frames_u8))       # MatExpr acc / n -> convertTo(scale = 1/n)
    return stack, [r[1] for r in results], [r[2] for r in results]


def scale_image(img, scale_down):
    """src/utils.rs:186-214: smaller dimension -> scale_down, sizes truncated, INTER_AREA."""
    h, w = img.shape[:2]
    factor = float(scale_down) / float(w if w < h else h)
    return cv2.resize(img, (int(w * factor), int(h * factor)), interpolation=cv2.INTER_AREA)


def ecc_match_scaling_down(frames_u8, motion, max_count, epsilon, gauss_filt_size, scale_down, workers=1):
    """ecc_match_scaling_down (src/lib.rs:849-1028) over in-memory frames.
    Returns (stack f32, full-resolution warps, rhos)."""
    if len(frames_u8) == 0:
        raise ValueError("NotEnoughFiles")
    criteria = term_criteria(max_count, epsilon)
    grey0, f32_0 = read_grey_and_f32(frames_u8[0])
    h, w = grey0.shape
    if scale_down >= w:
        raise ValueError("InvalidParams: scale_down_to was larger (or equal) to the full image width")
    if scale_down <= 10.0:
        raise ValueError("InvalidParams: scale_down_to was too small")
    grey0_small = scale_image(grey0, scale_down)
    sh, sw = grey0_small.shape

    def task(i):
        if i == 0:
            return f32_0.copy(), None, None
        grey, f32 = read_grey_and_f32(frames_u8[i])
        grey_small = scale_image(grey, scale_down)
        rho, m = align_frame(grey_small, grey0_small, motion, criteria, gauss_filt_size)
        if motion != MOTION_HOMOGRAPHY:
            m = m.copy()
            m[0, 2] *= np.float32(w) / np.float32(sw)          # src/lib.rs:946-949
            m[1, 2] *= np.float32(h) / np.float32(sh)
        else:
            sx, sy = np.float32(w / sw), np.float32(h / sh)    # src/utils.rs:228-241 (f32 variant)
            m = m.copy()
            m[0, 2] *= sx
            m[1, 2] *= sy
            m[2, 0] /= sx
            m[2, 1] /= sy
        return warp_frame(f32, m, motion), m, rho

    if workers == 1:
        results = [task(i) for i in range(len(frames_u8))]
    else:
        with ThreadPoolExecutor(max_workers=workers) as ex:
            results = list(ex.map(task, range(len(frames_u8))))
    acc = None
    for warped, _, _ in results:
        acc = warped if acc is None else acc + warped
    return acc * np.float32(1.0 / len(frames_u8)), [r[1] for r in results], [r[2] for r in results]


def sharpness_tenengrad(grey_u8, ksize):
    """src/lib.rs:1101-1147."""
    if ksize not in (1, 3, 5, 7):
        raise ValueError("Kernel size must be 1, 3, 5, or 7")
    gx = cv2.Sobel(grey_u8, cv2.CV_64F, 1, 0, ksize=ksize, scale=1.0, delta=0.0, borderType=cv2.BORDER_DEFAULT)
    gy = cv2.Sobel(grey_u8, cv2.CV_64F, 0, 1, ksize=ksize, scale=1.0, delta=0.0, borderType=cv2.BORDER_DEFAULT)
    s = cv2.add(cv2.multiply(gx, gx), cv2.multiply(gy, gy))
    return cv2.mean(s)[0]


def sharpness_modified_laplacian(grey_u8):
    """src/lib.rs:1032-1068."""
    m = np.array([[-1.0, 2.0, -1.0]], np.float64)
    g = cv2.getGaussianKernel(3, -1.0, cv2.CV_64F)
    lx = cv2.sepFilter2D(grey_u8, cv2.CV_64F, m, g, anchor=(-1, -1), delta=0.0, borderType=cv2.BORDER_DEFAULT)
    ly = cv2.sepFilter2D(grey_u8, cv2.CV_64F, g, m, anchor=(-1, -1), delta=0.0, borderType=cv2.BORDER_DEFAULT)
    return cv2.mean(np.abs(lx) + np.abs(ly))[0]


def sharpness_variance_of_laplacian(grey_u8):
    """src/lib.rs:1070-1090."""
    lap = cv2.Laplacian(grey_u8, cv2.CV_64F, ksize=3, scale=1.0, delta=0.0, borderType=cv2.BORDER_REPLICATE)
    _, sigma = cv2.meanStdDev(lap)
    return float(sigma[0, 0]) * float(sigma[0, 0])


def sharpness_normalized_gray_level_variance(grey_u8):
    """src/lib.rs:1151-1166."""
    mu, sigma = cv2.meanStdDev(grey_u8.astype(np.float64))
    return float(sigma[0, 0]) ** 2 / max(float(mu[0, 0]), float(np.finfo(np.float64).eps))


def keep_count(n_good: int, keep_ratio: float) -> int:
    """src/lib.rs:235: `(filtered_matches.len() as f32 * params.match_keep_ratio).round() as usize` — Rust's f32::round rounds
    halves AWAY from zero (Python's round() goes to even: 6 * 0.75 = 4.5 must keep 5, not 4)."""
    x = np.float32(n_good) * np.float32(keep_ratio)
    return int(np.floor(x + np.float32(0.5))) if x >= 0 else 0


def keypoint_homography(grey0_kp_des, grey_i, method, reproj, match_ratio, keep_ratio):
    """src/lib.rs:200-287: ORB -> BF kNN(2) -> Lowe ratio -> sort -> keep -> findHomography(dst->src).
    Returns a 3x3 f64 matrix or None when the reference would drop the frame."""
    kp0, des0 = grey0_kp_des
    orb = cv2.ORB_create()
    kp, des = orb.detectAndCompute(grey_i, None)
    if des is None or len(kp) < 2:
        return None
    matcher = cv2.BFMatcher(cv2.NORM_HAMMING, False)
    knn = matcher.knnMatch(des0, des, k=2)
    good = [m[0] for m in knn if len(m) == 2 and m[0].distance < np.float32(match_ratio) * m[1].distance]
    good.sort(key=lambda m: m.distance)
    good = good[:keep_count(len(good), keep_ratio)]
    if len(good) < 5:
        return None
    src = np.float32([kp0[m.queryIdx].pt for m in good]).reshape(-1, 1, 2)
    dst = np.float32([kp[m.trainIdx].pt for m in good]).reshape(-1, 1, 2)
    try:
        h, _ = cv2.findHomography(dst, src, method, reproj)
    except cv2.error:
        return None
    if h is None or h.shape != (3, 3) or abs(np.linalg.det(h)) < 1e-6:
        return None
    return h


def keypoint_match(frames_u8, method=cv2.RANSAC, reproj=3.0, match_ratio=0.8, keep_ratio=0.75,
                   border_mode=cv2.BORDER_CONSTANT, border_value=(0, 0, 0, 0), scale_down=None):
    """keypoint_match_no_scale, sequential fold (the reference's dropped-frame seeding quirk at
    src/lib.rs:307 depends on the Rayon split; with zero drops every order gives the same sum).
    Returns (dropped, stack f32, homographies)."""
    if len(frames_u8) == 0:
        raise ValueError("NotEnoughFiles")
    grey0, f32_0 = read_grey_and_f32(frames_u8[0])
    h0, w0 = grey0.shape
    if scale_down is not None:
        # keypoint_match_scale_down (src/lib.rs:355-600): features on the downscaled greys, homography
        # adjusted with adjust_homography_for_scale_f64 (src/utils.rs:218-248)
        if scale_down >= w0:
            raise ValueError("InvalidParams: scale_down_to was larger (or equal) to the full image width")
        grey0 = scale_image(grey0, scale_down)
    orb = cv2.ORB_create()
    kp0, des0 = orb.detectAndCompute(grey0, None)
    acc, dropped, hs = f32_0.copy(), 0, [None]
    for fr in frames_u8[1:]:
        grey, f32 = read_grey_and_f32(fr)
        if scale_down is not None:
            grey = scale_image(grey, scale_down)
        h = keypoint_homography((kp0, des0), grey, method, reproj, match_ratio, keep_ratio)
        if h is not None and scale_down is not None:
            sx, sy = w0 / grey.shape[1], h0 / grey.shape[0]
            h = h.copy()
            h[0, 2] *= sx
            h[1, 2] *= sy
            h[2, 0] /= sx
            h[2, 1] /= sy
        hs.append(h)
        if h is None:
            dropped += 1
            continue
        acc = acc + cv2.warpPerspective(f32, h, (w0, h0), flags=cv2.INTER_LINEAR,
                                        borderMode=border_mode, borderValue=border_value)
    return dropped, acc * np.float32(1.0 / (len(frames_u8) - dropped)), hs
