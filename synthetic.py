"""Seeded synthetic stacks: the INPUTS of the parity tests and of the benchmark.

An input generator, not part of the oracle and not part of the product: it lives at the repo root so that
bench.py's GPU arm can make its frames without importing anything under oracle/ (`oracle/synth.py` re-exports it
for the tests).  Nothing in libstacker.rs_b200/ imports this file.

The reference ships no images (its example reads a third-party set that is not in the tree,
/root/reference/README.md:16-18, examples/main.rs:35), so stacks are synthesised:

  scene   : float canvas (W+2*MARGIN) x (H+2*MARGIN) x 3 = low-frequency colour field
            (bicubic-upsampled uniform noise at 1/128 and 1/32 scale, amplitudes 50/30, offset 50)
            + W*H/2500 random rotated rectangles (side 6-60 px, colour offset U(-70,70), anti-aliased)
            + Gaussian blur sigma 0.8
  frame i : bicubic warp of the canvas with a known ground-truth matrix G_i (frame 0: identity)
            + N(0, 3^2) noise, rounded and clipped to u8 BGR.

G_i maps frame-i pixel coordinates to frame-0 pixel coordinates, i.e. it is the matrix ECC is
expected to recover when called as the reference calls it (template = frame i, input = frame 0,
/root/reference/src/lib.rs:769-772).
"""
from __future__ import annotations

import numpy as np
import cv2

MARGIN = 48

MOTION_TRANSLATION, MOTION_EUCLIDEAN, MOTION_AFFINE, MOTION_HOMOGRAPHY = 0, 1, 2, 3


def make_scene(width: int, height: int, seed: int) -> np.ndarray:
    """Float32 BGR canvas of size (height+2*MARGIN, width+2*MARGIN, 3)."""
    rng = np.random.default_rng(seed)
    cw, ch = width + 2 * MARGIN, height + 2 * MARGIN
    canvas = np.full((ch, cw, 3), 50.0, np.float32)
    for scale, amp in ((128, 50.0), (32, 30.0)):
        lw, lh = max(2, cw // scale + 2), max(2, ch // scale + 2)
        low = rng.uniform(0.0, 1.0, (lh, lw, 3)).astype(np.float32)
        canvas += amp * cv2.resize(low, (cw, ch), interpolation=cv2.INTER_CUBIC)
    n_rect = max(8, (width * height) // 2500)
    layer = np.zeros((ch, cw, 3), np.float32)
    cx = rng.uniform(0, cw, n_rect)
    cy = rng.uniform(0, ch, n_rect)
    sw = rng.uniform(6, 60, n_rect)
    sh = rng.uniform(6, 60, n_rect)
    ang = rng.uniform(0, 180, n_rect)
    col = rng.uniform(-70, 70, (n_rect, 3))
    for i in range(n_rect):
        box = cv2.boxPoints(((float(cx[i]), float(cy[i])), (float(sw[i]), float(sh[i])), float(ang[i])))
        pts = np.round(box * 16).astype(np.int32)  # 4 fractional bits
        cv2.fillConvexPoly(layer, pts, tuple(float(c) for c in col[i]), lineType=cv2.LINE_AA, shift=4)
    canvas += layer
    canvas = cv2.GaussianBlur(canvas, (0, 0), 0.8)
    return canvas


def random_warp(rng: np.random.Generator, motion: int, width: int, height: int) -> np.ndarray:
    """Ground-truth 3x3 (float64) matrix mapping frame-i coordinates to frame-0 coordinates."""
    tx, ty = rng.uniform(-4, 4, 2)
    theta = rng.uniform(-0.01, 0.01) if motion != MOTION_TRANSLATION else 0.0
    c, s = np.cos(theta), np.sin(theta)
    a = np.array([[c, -s], [s, c]], np.float64)
    if motion in (MOTION_AFFINE, MOTION_HOMOGRAPHY):
        a = a + rng.uniform(-0.005, 0.005, (2, 2))
    g = np.eye(3)
    g[:2, :2] = a
    # rotate / shear about the image centre so the corner displacement stays inside MARGIN
    ctr = np.array([(width - 1) / 2.0, (height - 1) / 2.0])
    g[:2, 2] = ctr - a @ ctr + np.array([tx, ty])
    if motion == MOTION_HOMOGRAPHY:
        p = rng.uniform(-2e-6, 2e-6, 2)
        g[2, 0], g[2, 1] = p
        # keep the centre where the affine part put it
        g[2, 2] = 1.0 - p @ ctr
        g /= g[2, 2]
    return g


def render_frame(canvas: np.ndarray, g: np.ndarray, width: int, height: int,
                 rng: np.random.Generator | None, noise_sigma: float = 3.0,
                 blur_sigma: float = 0.0) -> np.ndarray:
    """u8 BGR frame: frame(x) = canvas(G x + MARGIN)."""
    off = np.array([[1, 0, MARGIN], [0, 1, MARGIN], [0, 0, 1]], np.float64)
    m = off @ g
    img = cv2.warpPerspective(canvas, m, (width, height),
                              flags=cv2.INTER_CUBIC | cv2.WARP_INVERSE_MAP,
                              borderMode=cv2.BORDER_REFLECT)
    if blur_sigma > 0:
        img = cv2.GaussianBlur(img, (0, 0), blur_sigma)
    if rng is not None and noise_sigma > 0:
        img = img + rng.standard_normal(img.shape, dtype=np.float32) * np.float32(noise_sigma)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


class Stack:
    """Lazy synthetic stack: frame(i) is reproducible for any i without rendering the others."""

    def __init__(self, width: int, height: int, n_frames: int, motion: int, seed: int,
                 noise_sigma: float = 3.0, blur_sigmas=None):
        self.width, self.height, self.n, self.motion, self.seed = width, height, n_frames, motion, seed
        self.noise_sigma = noise_sigma
        self.blur_sigmas = blur_sigmas
        self._canvas = None
        wrng = np.random.default_rng(seed + 1000)
        self.truth = [np.eye(3)] + [random_warp(wrng, motion, width, height) for _ in range(n_frames - 1)]

    @property
    def canvas(self):
        if self._canvas is None:
            self._canvas = make_scene(self.width, self.height, self.seed)
        return self._canvas

    def frame(self, i: int) -> np.ndarray:
        nrng = np.random.default_rng([self.seed + 2000, i])
        bs = 0.0 if self.blur_sigmas is None else float(self.blur_sigmas[i])
        return render_frame(self.canvas, self.truth[i], self.width, self.height, nrng,
                            self.noise_sigma, bs)

    def frames(self):
        return [self.frame(i) for i in range(self.n)]


# The five BASELINE.json configs (index = config number - 1).  seeds: scene seed = config number.
CONFIGS = {
    1: dict(width=1024, height=768, n_frames=5, motion=MOTION_HOMOGRAPHY, seed=1),
    2: dict(width=1920, height=1080, n_frames=16, motion=MOTION_EUCLIDEAN, seed=2),
    3: dict(width=3840, height=2160, n_frames=32, motion=MOTION_AFFINE, seed=3),
    4: dict(width=3840, height=2160, n_frames=64, motion=MOTION_HOMOGRAPHY, seed=4),
    5: dict(width=6000, height=4000, n_frames=256, motion=MOTION_HOMOGRAPHY, seed=5),
}


def config_stack(cfg: int, n_frames: int | None = None, width: int | None = None,
                 height: int | None = None) -> Stack:
    kw = dict(CONFIGS[cfg])
    if n_frames is not None:
        kw["n_frames"] = n_frames
    if width is not None:
        kw["width"] = width
    if height is not None:
        kw["height"] = height
    if cfg == 3:
        # distinct blur per frame so the Tenengrad values are well separated
        r = np.random.default_rng(kw["seed"] + 3000)
        kw["blur_sigmas"] = r.permutation(kw["n_frames"]) * 0.05 + 0.3
    return Stack(**kw)


def corner_displacement(m_a: np.ndarray, m_b: np.ndarray, width: int, height: int) -> float:
    """Max distance (px) between the images of the four frame corners under two matrices."""
    def full(m):
        m = np.asarray(m, np.float64)
        return np.vstack([m, [0, 0, 1]]) if m.shape[0] == 2 else m
    a, b = full(m_a), full(m_b)
    pts = np.array([[0, 0, 1], [width - 1, 0, 1], [0, height - 1, 1], [width - 1, height - 1, 1]], np.float64).T
    pa, pb = a @ pts, b @ pts
    pa, pb = pa[:2] / pa[2], pb[:2] / pb[2]
    return float(np.max(np.hypot(*(pa - pb))))
