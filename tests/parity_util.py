"""Shared parity checks (north_star bars): warp matrices within 0.05 px corner displacement; 8-bit stack
max-abs-diff <= 1 and PSNR >= 50 dB."""
import math

import numpy as np

from oracle import restate as R


def psnr8(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)


def coverage(warps, motion, width, height):
    """Pixels every warped frame covers with all four bilinear taps inside its source (the zero-blended
    rim of a warped frame is a 1-px step of ~100 grey levels, where a 1/32-px coordinate quantum moves the
    8-bit value by several levels)."""
    ones = np.full((height, width), 255, np.uint8)
    cov = np.ones((height, width), bool)
    for m in warps:
        if m is None:
            continue
        cov &= R.final_warp(ones, m, motion) == np.float32(1.0)
    return cov


def assert_stack_parity(got_f32, want_f32, warps, motion, n_frames):
    """The north-star bars, literally: 8-bit max-abs-diff <= 1 and PSNR >= 50 dB.  Every parity stack has at least
    5 frames: one frame's zero-blended rim pixel moves by ~3 grey levels per 1/32-px coordinate quantum, and the
    bars are stated for stacks, where that is divided by n."""
    assert n_frames >= 5, "parity stacks have at least 5 frames"
    g8, w8 = np.rint(got_f32 * 255.0), np.rint(want_f32 * 255.0)
    d = np.abs(g8 - w8)
    assert psnr8(g8, w8) >= 50.0, psnr8(g8, w8)
    assert d.max() <= 1, d.max()
