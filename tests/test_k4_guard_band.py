"""Soundness of the guarded f32 coordinates of the second-generation final-warp kernel (csrc/warp_acc.cuh, FastInv),
checked on the CPU by emulating its f32 arithmetic in NumPy — no GPU needed.

The kernel evaluates the displacement t = 32 (u - x) of cv2.warpPerspective's inverse map in f32, rounds it through the
1.5 * 2^23 magic constant and ACCEPTS the result only when the residual |t - rint(t)| is at most 0.5 - band, band being the
per-column error bound FastInv::init computes; every other pixel is recomputed with OpenCV's own f64 sequence.  Bit-exactness
of the warp (the `array_equal` GPU tests against cv2) therefore rests on: an accepted pixel's rounded coordinate equals the
exact one.  This test draws random stack-like and strongly perspective inverse maps, runs the emulation with the worst-case
error of rcp.approx (+-2^-23) and asserts that no accepted coordinate differs from rint of the exact value."""
import numpy as np

f32 = np.float32


def _fma32(a, b, c):
    # fused multiply-add on f32 inputs: the product of two f32 is exact in f64; one rounding to f32 (up to double rounding)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def _check(m, w, h, tile_y, rng, n=4096):
    xs = rng.integers(0, w, n).astype(np.float64)
    ys = (tile_y + rng.integers(0, 32, n)).astype(np.float64)
    y_lo, y_hi = f32(tile_y), f32(tile_y + 31)
    wcd = m[2, 0] * xs + m[2, 2]
    a32 = (32.0 * (xs * (m[0, 0] - wcd) + m[0, 2])).astype(f32)
    b32 = (32.0 * (m[0, 1] - m[2, 1] * xs)).astype(f32)
    g32 = (32.0 * (m[1, 0] * xs + m[1, 2])).astype(f32)
    d32 = (32.0 * (m[1, 1] - wcd)).astype(f32)
    m7_32, wc, m7 = f32(32.0 * m[2, 1]), wcd.astype(f32), f32(m[2, 1])
    su = np.abs(a32) + np.abs(b32) * y_hi
    sv = np.abs(g32) + (np.abs(d32) + np.abs(m7_32) * y_hi) * y_hi
    w_lo = _fma32(np.full_like(wc, m7), np.full_like(wc, y_lo), wc)
    w_hi = _fma32(np.full_like(wc, m7), np.full_like(wc, y_hi), wc)
    wmin = np.minimum(w_lo, w_hi) * f32(0.99)
    ok = (wmin > 0.25) & (np.maximum(w_lo, w_hi) < 4) & (np.maximum(su, sv) < 1e6)
    with np.errstate(divide="ignore", invalid="ignore"):
        band = (np.maximum(su, sv) / wmin * f32(12.0 / 16777216.0) + f32(1.0 / 262144.0)).astype(f32)
    thr = np.where(ok, f32(0.5) - band, f32(-1))
    yf = ys.astype(f32)
    wf = _fma32(np.full_like(yf, m7), yf, wc)
    nu = _fma32(b32, yf, a32)
    nv = _fma32(_fma32(np.full_like(yf, -m7_32), yf, d32), yf, g32)
    wd = m[2, 0] * xs + m[2, 1] * ys + m[2, 2]
    exact = (32.0 * ((m[0, 0] * xs + m[0, 1] * ys + m[0, 2]) / wd - xs), 32.0 * ((m[1, 0] * xs + m[1, 1] * ys + m[1, 2]) / wd - ys))
    accepted = wrong = 0
    for eps in (-2.0 ** -23, 0.0, 2.0 ** -23):            # rcp.approx.ftz.f32: relative error at most 2^-23
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            r = ((1.0 / wf.astype(np.float64)) * (1 + eps)).astype(f32)
            for num, t_ex in zip((nu, nv), exact):
                t = num.astype(np.float64) * r.astype(np.float64)             # the product inside fma(n, r, magic)
                q = np.rint((t + 12582912.0).astype(f32).astype(np.float64) - 12582912.0)
                res = (t - q).astype(f32)
                acc = np.abs(res) <= thr                                       # NaN compares false: sent to the exact path
                accepted += int(acc.sum())
                wrong += int((acc & (q != np.rint(t_ex))).sum())
    return accepted, wrong, 6 * n


def test_accepted_coordinates_equal_the_exact_ones():
    rng = np.random.default_rng(1)
    w, h = 3840, 2160
    tot_acc = tot = 0
    for trial in range(90):
        ang, sc = rng.uniform(-0.05, 0.05), rng.uniform(0.95, 1.05)
        m = np.array([[sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(-40, 40)],
                      [sc * np.sin(ang), sc * np.cos(ang), rng.uniform(-40, 40)],
                      [rng.uniform(-3e-5, 3e-5), rng.uniform(-3e-5, 3e-5), 1.0]])
        if trial % 3 == 0:
            m[2, :2] *= 10                                 # w between ~0 and 2.5 over the frame: columns leave the f32 range
        acc, wrong, n = _check(m, w, h, int(rng.integers(0, h // 32)) * 32, rng)
        assert wrong == 0
        tot_acc += acc
        tot += n
    assert tot_acc > 0.8 * tot                              # the fast evaluation is the common case, not the exception


def test_stack_like_motion_rarely_needs_the_exact_path():
    """ECC-sized motions (a few pixels): fewer than 1 % of the coordinates fall inside the guard band."""
    rng = np.random.default_rng(2)
    w, h = 3840, 2160
    tot_acc = tot = 0
    for _ in range(30):
        m = np.eye(3) + rng.uniform(-1, 1, (3, 3)) * np.array([[2e-3, 2e-3, 6.0], [2e-3, 2e-3, 6.0], [2e-7, 2e-7, 0.0]])
        acc, wrong, n = _check(m, w, h, int(rng.integers(0, h // 32)) * 32, rng)
        assert wrong == 0
        tot_acc += acc
        tot += n
    assert tot_acc > 0.99 * tot


def test_byte_splice_conversion_is_exact():
    """K4 turns a tap byte b into fl(b * fl(1/255)) without a conversion instruction: PRMT splices b under the exponent of
    2^23 (the float 2^23 + b) and ONE fused multiply-add computes fma(2^23 + b, k, -(2^23 k)).  2^23 k is exact (a power-of-two
    multiple of k), so the FMA's single rounding is that of b * k: the value OpenCV's convertTo(CV_32F, 1/255) produces."""
    k = f32(1.0 / 255.0)
    c = f32(8388608.0) * k
    assert float(c) == 8388608.0 * float(k)                                   # exact: no rounding in 2^23 * k
    b = np.arange(256)
    spliced = (np.uint32(0x4B000000) | b.astype(np.uint32)).view(f32)
    assert np.array_equal(spliced, (8388608.0 + b).astype(f32))
    fused = (spliced.astype(np.float64) * float(k) - float(c)).astype(f32)    # exact in f64, rounded once
    assert np.array_equal(fused, b.astype(f32) * k)
    # the fraction splice: (1.5 * 2^23 + q) / 32 - 1.5 * 2^23 / 32 == q / 32 for the 5-bit fractions
    q = np.arange(32)
    sp = (np.uint32(0x4B400000) | q.astype(np.uint32)).view(f32)
    assert np.array_equal((sp.astype(np.float64) * (1.0 / 32) - 12582912.0 / 32).astype(f32), (q / 32.0).astype(f32))
