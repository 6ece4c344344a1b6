"""Generates tests/golden/*.npz with the REAL OpenCV (cv2) on seeded synthetic inputs.

Run in a container that has cv2 (this image: 4.13.0; the reference pins OpenCV 4.12.0):
    python tests/golden/make_golden.py
The vectors pin the oracle (oracle/restate.py) where cv2 is unavailable, and are used by the GPU suite
as a second, committed reference.  Inputs are regenerated from the seeds by oracle/synth.py, so only the
cv2 OUTPUTS are stored."""
import os
import sys

import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import synth, cvref  # noqa: E402


# every golden stack has 5 frames: the north-star bar (8-bit max-abs-diff <= 1) is asserted literally on stacks of
# >= 5 frames (a zero-blended rim pixel of ONE frame moves by ~3 levels per 1/32-px coordinate quantum; / n)
N_STACK = 5


def ecc_cases():
    out = {}
    for motion in (0, 1, 2, 3):
        st = synth.Stack(256, 192, N_STACK, motion, seed=100 + motion)
        frames = st.frames()
        stack, warps, rhos = cvref.ecc_match(frames, motion, 5000, 1e-5, 5, workers=1)
        out[f"m{motion}_warps"] = np.stack([np.vstack([w, [0, 0, 1]]) if w.shape[0] == 2 else w for w in warps[1:]]).astype(np.float32)
        out[f"m{motion}_rhos"] = np.array(rhos[1:], np.float64)
        out[f"m{motion}_stack8"] = np.rint(stack * 255.0).astype(np.uint8)
    return out


def primitive_cases():
    out = {}
    st = synth.Stack(200, 150, 2, 3, seed=55)
    f0, f1 = st.frames()
    out["grey"] = cv2.cvtColor(f0, cv2.COLOR_BGR2GRAY)
    for k in (3, 5, 7, 9):
        out[f"blur{k}"] = cv2.GaussianBlur(out["grey"].astype(np.float32), (k, k), 0)
    rng = np.random.default_rng(8)
    hm = synth.random_warp(rng, 3, 200, 150)
    hm[:2, 2] += (13.3, -7.7)
    am = synth.random_warp(rng, 2, 200, 150)[:2]
    f32 = f1.astype(np.float32) * np.float32(1 / 255.0)
    out["H"] = hm
    out["A"] = am
    out["warp_persp"] = cv2.warpPerspective(f32, hm, (200, 150), flags=cv2.INTER_LINEAR)
    out["warp_affine"] = cv2.warpAffine(f32, am, (200, 150), flags=cv2.INTER_LINEAR)
    for k in (1, 3, 5, 7):
        out[f"teng{k}"] = np.float64(cvref.sharpness_tenengrad(out["grey"], k))
    return out


RESIZE_CASES = [  # (width, height, scale_down): generic table path, 2x2, 3x3 and 4x4 integer fast paths, portrait
    (200, 150, 64.0), (256, 192, 96.0), (300, 240, 80.0), (640, 480, 120.0), (301, 201, 67.0), (150, 200, 70.0),
]
SCALE_DOWN_CASES = [  # (motion, width, height, scale_down, seed)
    (0, 480, 360, 240.0, 61), (1, 480, 360, 240.0, 62), (2, 640, 480, 320.0, 63), (3, 800, 600, 400.0, 65),
]


def scale_down_cases():
    """ecc_match_scaling_down (src/lib.rs:849-1028): INTER_AREA resize of the greys and the whole scaled path."""
    out = {}
    for w, h, sd in RESIZE_CASES:
        rng = np.random.default_rng(w * 7 + h)
        grey = rng.integers(0, 256, (h, w), dtype=np.uint8)
        out[f"resize_{w}x{h}_{int(sd)}"] = cvref.scale_image(grey, sd)
    for motion, w, h, sd, seed in SCALE_DOWN_CASES:
        frames = synth.Stack(w, h, N_STACK, motion, seed=seed).frames()
        stack, warps, _ = cvref.ecc_match_scaling_down(frames, motion, 60, 1e-5, 5, sd)
        out[f"sd_m{motion}_warps"] = np.stack([np.vstack([m, [0, 0, 1]]) if m.shape[0] == 2 else m for m in warps[1:]]).astype(np.float32)
        out[f"sd_m{motion}_stack8"] = np.rint(stack * 255.0).astype(np.uint8)
    return out


RESIZE_UP_CASES = [  # landscape frames with height < scale_down < width: utils::scale_image ENLARGES (cv::resize INTER_AREA
    # falls back to its 8-bit bilinear kernels in "area mode"); the last one enlarges x only (new height == height)
    (320, 240, 250.0), (512, 384, 400.0), (200, 100, 199.0), (301, 201, 260.0), (333, 250, 250.9),
]
SCALE_UP_CASES = [(2, 480, 360, 400.0, 66), (3, 400, 300, 330.0, 67)]   # (motion, width, height, scale_down, seed)


def scale_up_cases():
    """ecc_match_scaling_down when utils::scale_image enlarges the greys (src/utils.rs:186-214)."""
    out = {}
    for w, h, sd in RESIZE_UP_CASES:
        rng = np.random.default_rng(w * 7 + h)
        grey = rng.integers(0, 256, (h, w), dtype=np.uint8)
        out[f"resize_{w}x{h}_{int(sd)}"] = cvref.scale_image(grey, sd)
    for motion, w, h, sd, seed in SCALE_UP_CASES:
        frames = synth.Stack(w, h, N_STACK, motion, seed=seed).frames()
        stack, warps, _ = cvref.ecc_match_scaling_down(frames, motion, 60, 1e-5, 5, sd)
        out[f"sd_m{motion}_warps"] = np.stack([np.vstack([m, [0, 0, 1]]) if m.shape[0] == 2 else m for m in warps[1:]]).astype(np.float32)
        out[f"sd_m{motion}_stack8"] = np.rint(stack * 255.0).astype(np.uint8)
    return out


def sharpness_cases():
    """LAPM / LAPV / TENG(3) / GLVN (src/lib.rs:1032-1166) on a random and on a synthetic-scene grey plane."""
    out = {}
    rng = np.random.default_rng(21)
    greys = {"rand": rng.integers(0, 256, (131, 257), dtype=np.uint8),
             "scene": cv2.cvtColor(synth.Stack(320, 240, 1, 0, seed=22).frames()[0], cv2.COLOR_BGR2GRAY)}
    for name, g in greys.items():
        out[name] = np.array([cvref.sharpness_modified_laplacian(g), cvref.sharpness_variance_of_laplacian(g),
                              cvref.sharpness_tenengrad(g, 3), cvref.sharpness_normalized_gray_level_variance(g)], np.float64)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "scale_down.npz"), **scale_down_cases())
    np.savez_compressed(os.path.join(HERE, "scale_up.npz"), **scale_up_cases())
    np.savez_compressed(os.path.join(HERE, "sharpness.npz"), **sharpness_cases())
    np.savez_compressed(os.path.join(HERE, "ecc_256x192.npz"), **ecc_cases())
    np.savez_compressed(os.path.join(HERE, "primitives_200x150.npz"), **primitive_cases())
    print("cv2", cv2.__version__, "golden vectors written")
