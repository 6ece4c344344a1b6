"""NumPy restatement of the arithmetic on libstacker's ECC align-and-stack path.

TEST INFRASTRUCTURE ONLY — this file is the CPU *oracle* the CUDA kernels are checked against.
Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may import it; the product
path (libstacker.rs_b200/) never does and fails loudly when its CUDA library is missing.

Where the algorithm lives.  The reference (/root/reference/src/lib.rs) contains no arithmetic of
its own on this path: every numeric step is a call into OpenCV 4.12.0 C++ (pinned in
/root/reference/.github/workflows/rust.yml:53 and README.md:24) through the `opencv` crate 0.97.2
(/root/reference/Cargo.toml:19).  OpenCV is a third-party dependency that is NOT vendored under
/root/reference, so each function below restates the published OpenCV algorithm (upstream file
named for orientation) and cites the reference call site it stands for.

How it is pinned.  The same OpenCV functions are importable in this image as Python `cv2` 4.13.0.
tests/test_oracle_vs_cv2.py checks every function here against the real cv2 call (bit-exact for the
integer/byte/sampling work, <= 1e-3 px corner displacement for the ECC solver) and
tests/golden/ holds cv2-generated vectors (made by tests/golden/make_golden.py) so the check also
runs where cv2 is absent.  The reference's own tests pin only the TermCriteria flag semantics
(/root/reference/src/utils.rs:148-158); that doctest is restated in tests/test_host_api.py.
"""
from __future__ import annotations

import math
import numpy as np

MOTION_TRANSLATION, MOTION_EUCLIDEAN, MOTION_AFFINE, MOTION_HOMOGRAPHY = 0, 1, 2, 3
TERM_COUNT, TERM_EPS = 1, 2

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS          # 32
AB_BITS = 10
AB_SCALE = 1 << AB_BITS                   # 1024


class EccNoConvergence(Exception):
    """Mirrors cv::Error::StsNoConv raised by findTransformECC (-> StackerError::OpenCvError)."""


# --------------------------------------------------------------------------------------------
# read_grey_and_f32                                   /root/reference/src/utils.rs:128-144
# --------------------------------------------------------------------------------------------
def bgr2gray_u8(bgr: np.ndarray) -> np.ndarray:
    """cvtColor(BGR2GRAY) on 8-bit input [OpenCV color_rgb.simd.hpp, RGB2Gray<uchar>]:
    15-bit fixed point, coefficients B 3735, G 19235, R 9798, rounding constant 1<<14."""
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)


def to_f32_unit(img_u8: np.ndarray) -> np.ndarray:
    """Mat::convert_to(CV_32F, 1/255) (/root/reference/src/utils.rs:133, :20):
    float(src) * float(1/255), one rounding."""
    return img_u8.astype(np.float32) * np.float32(1.0 / 255.0)


# --------------------------------------------------------------------------------------------
# TermCriteria                                        /root/reference/src/utils.rs:159-170
# --------------------------------------------------------------------------------------------
def term_criteria(max_count, epsilon):
    """(typ, max_count, epsilon) exactly as From<EccMatchParameters> builds it: unset fields stay
    at TermCriteria::default() = 0."""
    typ, mc, eps = 0, 0, 0.0
    if max_count is not None:
        typ |= TERM_COUNT
        mc = int(max_count)
    if epsilon is not None:
        typ |= TERM_EPS
        eps = float(epsilon)
    return typ, mc, eps


# --------------------------------------------------------------------------------------------
# GaussianBlur / gradients inside findTransformECC    [OpenCV video/src/ecc.cpp, imgproc/smooth]
# --------------------------------------------------------------------------------------------
def gaussian_taps(ksize: int) -> np.ndarray:
    """getGaussianKernel(ksize, sigma<=0, CV_32F): fixed tables for ksize <= 9, else sampled
    Gaussian with sigma = 0.3*((ksize-1)*0.5 - 1) + 0.8, normalised."""
    small = {
        1: [1.0],
        3: [0.25, 0.5, 0.25],
        5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
        7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125],
        # cv2 4.13 also tabulates k = 9 ([4 13 30 51 60 51 30 13 4] / 256); checked against
        # cv2.getGaussianKernel in tests/test_oracle_vs_cv2.py
        9: [0.015625, 0.05078125, 0.1171875, 0.19921875, 0.234375, 0.19921875, 0.1171875, 0.05078125, 0.015625],
    }
    if ksize in small:
        return np.array(small[ksize], np.float32)
    sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    scale2x = -0.5 / (sigma * sigma)
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(scale2x * x * x)
    k /= k.sum()
    return k.astype(np.float32)


def _reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    idx = np.abs(idx) % period
    return np.where(idx >= n, period - idx, idx)


def gaussian_blur_f32(img: np.ndarray, ksize: int) -> np.ndarray:
    """GaussianBlur(CV_32F, (k,k), sigma 0) == separable row then column filter, f32 accumulation,
    BORDER_REFLECT_101.  For u8-valued input and k <= 7 every intermediate is dyadic and exact."""
    assert ksize >= 1 and ksize % 2 == 1
    img = img.astype(np.float32)
    if ksize == 1:
        return img.copy()
    taps = gaussian_taps(ksize)
    r = ksize // 2
    h, w = img.shape
    cols = _reflect101(np.arange(-r, w + r), w)
    rows = _reflect101(np.arange(-r, h + r), h)
    tmp = np.zeros((h, w), np.float32)
    pad = img[:, cols]
    for k in range(ksize):
        tmp += taps[k] * pad[:, k:k + w]
    out = np.zeros((h, w), np.float32)
    pad = tmp[rows, :]
    for k in range(ksize):
        out += taps[k] * pad[k:k + h, :]
    return out


def central_gradients(img: np.ndarray):
    """filter2D(img, [-0.5 0 0.5]) and its transpose, BORDER_REFLECT_101 => exactly 0 on the
    first/last column (row)."""
    gx = np.zeros_like(img)
    gy = np.zeros_like(img)
    gx[:, 1:-1] = (img[:, 2:] - img[:, :-2]) * np.float32(0.5)
    gy[1:-1, :] = (img[2:, :] - img[:-2, :]) * np.float32(0.5)
    return gx, gy


# --------------------------------------------------------------------------------------------
# warpAffine / warpPerspective coordinate rules        [OpenCV imgproc/src/imgwarp.cpp]
# --------------------------------------------------------------------------------------------
def invert_affine(m: np.ndarray) -> np.ndarray:
    """The closed form warpAffine uses when WARP_INVERSE_MAP is not set."""
    m = np.asarray(m, np.float64).reshape(2, 3).copy()
    d = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[1, 1] * d, m[0, 0] * d
    m[0, 0] = a11
    m[0, 1] *= -d
    m[1, 0] *= -d
    m[1, 1] = a22
    b1 = -m[0, 0] * m[0, 2] - m[0, 1] * m[1, 2]
    b2 = -m[1, 0] * m[0, 2] - m[1, 1] * m[1, 2]
    m[0, 2], m[1, 2] = b1, b2
    return m


def invert_3x3(m: np.ndarray) -> np.ndarray:
    """cv::invert(DECOMP_LU) special case for 3x3 double (adjugate / determinant)."""
    s = np.asarray(m, np.float64).reshape(3, 3)
    d = (s[0, 0] * (s[1, 1] * s[2, 2] - s[1, 2] * s[2, 1])
         - s[0, 1] * (s[1, 0] * s[2, 2] - s[1, 2] * s[2, 0])
         + s[0, 2] * (s[1, 0] * s[2, 1] - s[1, 1] * s[2, 0]))
    if d == 0:
        return np.zeros((3, 3))
    d = 1.0 / d
    t = np.empty((3, 3))
    t[0, 0] = (s[1, 1] * s[2, 2] - s[1, 2] * s[2, 1]) * d
    t[0, 1] = (s[0, 2] * s[2, 1] - s[0, 1] * s[2, 2]) * d
    t[0, 2] = (s[0, 1] * s[1, 2] - s[0, 2] * s[1, 1]) * d
    t[1, 0] = (s[1, 2] * s[2, 0] - s[1, 0] * s[2, 2]) * d
    t[1, 1] = (s[0, 0] * s[2, 2] - s[0, 2] * s[2, 0]) * d
    t[1, 2] = (s[0, 2] * s[1, 0] - s[0, 0] * s[1, 2]) * d
    t[2, 0] = (s[1, 0] * s[2, 1] - s[1, 1] * s[2, 0]) * d
    t[2, 1] = (s[0, 1] * s[2, 0] - s[0, 0] * s[2, 1]) * d
    t[2, 2] = (s[0, 0] * s[1, 1] - s[0, 1] * s[1, 0]) * d
    return t


def _sat_int(v: np.ndarray) -> np.ndarray:
    """saturate_cast<int>(double): round half to even, clamp to int32."""
    return np.clip(np.rint(v), -2147483648.0, 2147483647.0).astype(np.int64)


def perspective_fixed_coords(im: np.ndarray, width: int, height: int, nearest: bool = False):
    """Per destination pixel: (X, Y) in 1/32 px (linear) or whole px (nearest), from the INVERSE map
    `im` (3x3 f64) as WarpPerspectiveInvoker computes them: 64-px column blocks, X0 evaluated at the
    block start, W = TAB/W, clamp, round half-even."""
    im = np.asarray(im, np.float64).reshape(3, 3)
    bw = min(64, width)
    x = np.arange(width)
    xb = (x // bw) * bw
    x1 = (x - xb).astype(np.float64)
    xb = xb.astype(np.float64)
    y = np.arange(height, dtype=np.float64)[:, None]
    x0 = im[0, 0] * xb[None, :] + im[0, 1] * y + im[0, 2]
    y0 = im[1, 0] * xb[None, :] + im[1, 1] * y + im[1, 2]
    w0 = im[2, 0] * xb[None, :] + im[2, 1] * y + im[2, 2]
    w = w0 + im[2, 0] * x1[None, :]
    num = 1.0 if nearest else float(INTER_TAB_SIZE)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(w != 0, num / w, 0.0)
    fx = np.maximum(-2147483648.0, np.minimum(2147483647.0, (x0 + im[0, 0] * x1[None, :]) * w))
    fy = np.maximum(-2147483648.0, np.minimum(2147483647.0, (y0 + im[1, 0] * x1[None, :]) * w))
    return _sat_int(fx), _sat_int(fy)


def affine_fixed_coords(im: np.ndarray, width: int, height: int, nearest: bool = False):
    """WarpAffineInvoker: 10-bit fixed point; linear -> 1/32 px units, nearest -> whole px."""
    im = np.asarray(im, np.float64).reshape(2, 3)
    x = np.arange(width, dtype=np.float64)
    y = np.arange(height, dtype=np.float64)
    adelta = _sat_int(im[0, 0] * x * AB_SCALE)
    bdelta = _sat_int(im[1, 0] * x * AB_SCALE)
    round_delta = AB_SCALE // 2 if nearest else AB_SCALE // INTER_TAB_SIZE // 2
    x0 = _sat_int((im[0, 1] * y + im[0, 2]) * AB_SCALE) + round_delta
    y0 = _sat_int((im[1, 1] * y + im[1, 2]) * AB_SCALE) + round_delta
    shift = AB_BITS if nearest else AB_BITS - INTER_BITS
    # the sums wrap as C int (they never do at sane sizes); the shift is arithmetic (floor)
    xx = (x0[:, None] + adelta[None, :]) >> shift
    yy = (y0[:, None] + bdelta[None, :]) >> shift
    return xx, yy


BORDER_CONSTANT, BORDER_REPLICATE, BORDER_REFLECT, BORDER_WRAP, BORDER_REFLECT_101 = 0, 1, 2, 3, 4


def border_interpolate(p: np.ndarray, length: int, border_mode: int) -> np.ndarray:
    """cv::borderInterpolate for the modes that always yield a valid index (REPLICATE, REFLECT, WRAP,
    REFLECT_101) — what KeyPointMatchParameters::border_mode (/root/reference/src/lib.rs:66-68, used at
    :297) may select besides BORDER_CONSTANT."""
    p = np.asarray(p, np.int64).copy()
    if border_mode == BORDER_REPLICATE:
        return np.clip(p, 0, length - 1)
    if border_mode in (BORDER_REFLECT, BORDER_REFLECT_101):
        if length == 1:
            return np.zeros_like(p)
        delta = 1 if border_mode == BORDER_REFLECT_101 else 0
        while True:
            neg, big = p < 0, p >= length
            if not (neg.any() or big.any()):
                return p
            p = np.where(neg, -p - 1 + delta, p)
            big = p >= length
            p = np.where(big, length - 1 - (p - length) - delta, p)
    if border_mode == BORDER_WRAP:
        return np.mod(p, length)
    raise ValueError(f"unsupported border mode {border_mode}")


def sample_bilinear_fixed(src: np.ndarray, xq: np.ndarray, yq: np.ndarray, border_value=0.0,
                          border_mode: int = BORDER_CONSTANT):
    """remapBilinear<float>: integer part = q >> 5 (saturated to short), weights from the 5-bit fraction,
    value = s00*w00 + s01*w01 + s10*w10 + s11*w11 summed left to right in f32.  BORDER_CONSTANT: taps outside
    the source are the border value; REPLICATE / REFLECT / WRAP / REFLECT_101: every tap's coordinates go
    through borderInterpolate on their own."""
    h, w = src.shape[:2]
    chan = 1 if src.ndim == 2 else src.shape[2]
    s = src.reshape(h, w, chan).astype(np.float32)
    sx = np.clip(xq >> INTER_BITS, -32768, 32767)
    sy = np.clip(yq >> INTER_BITS, -32768, 32767)
    ax = (xq & (INTER_TAB_SIZE - 1)).astype(np.float32) / np.float32(INTER_TAB_SIZE)
    ay = (yq & (INTER_TAB_SIZE - 1)).astype(np.float32) / np.float32(INTER_TAB_SIZE)
    one = np.float32(1.0)
    w00 = ((one - ay) * (one - ax))[..., None]
    w01 = ((one - ay) * ax)[..., None]
    w10 = (ay * (one - ax))[..., None]
    w11 = (ay * ax)[..., None]
    bv = np.broadcast_to(np.asarray(border_value, np.float32).reshape(-1)[:chan] if np.ndim(border_value)
                         else np.full(chan, border_value, np.float32), (chan,))

    if border_mode != BORDER_CONSTANT:
        x0, x1 = border_interpolate(sx, w, border_mode), border_interpolate(sx + 1, w, border_mode)
        y0, y1 = border_interpolate(sy, h, border_mode), border_interpolate(sy + 1, h, border_mode)
        out = s[y0, x0] * w00
        out = out + s[y0, x1] * w01
        out = out + s[y1, x0] * w10
        out = out + s[y1, x1] * w11
        return out.reshape(xq.shape + ((chan,) if src.ndim == 3 else ()))

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        v = s[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
        return np.where(ok[..., None], v, bv)

    out = tap(sy, sx) * w00
    out = out + tap(sy, sx + 1) * w01
    out = out + tap(sy + 1, sx) * w10
    out = out + tap(sy + 1, sx + 1) * w11
    # OpenCV writes the border value untouched when all four taps are outside
    allout = (sx >= w) | (sx + 1 < 0) | (sy >= h) | (sy + 1 < 0)
    out = np.where(allout[..., None], bv, out)
    return out.reshape(xq.shape + ((chan,) if src.ndim == 3 else ()))


def warp_linear(src: np.ndarray, m: np.ndarray, width: int, height: int, perspective: bool,
                inverse_map: bool, border_value=0.0, border_mode: int = BORDER_CONSTANT) -> np.ndarray:
    """warpPerspective / warpAffine, INTER_LINEAR, border mode CONSTANT (default) / REPLICATE / REFLECT / WRAP /
    REFLECT_101
    (/root/reference/src/lib.rs:780-803 forward map; ECC-internal warps use WARP_INVERSE_MAP)."""
    if perspective:
        im = np.asarray(m, np.float64).reshape(3, 3)
        if not inverse_map:
            im = invert_3x3(im)
        xq, yq = perspective_fixed_coords(im, width, height)
    else:
        im = np.asarray(m, np.float64).reshape(2, 3)
        if not inverse_map:
            im = invert_affine(im)
        xq, yq = affine_fixed_coords(im, width, height)
    return sample_bilinear_fixed(src, xq, yq, border_value, border_mode)


def warp_mask_nearest(m: np.ndarray, width: int, height: int, src_w: int, src_h: int,
                      perspective: bool) -> np.ndarray:
    """INTER_NEAREST + WARP_INVERSE_MAP warp of an all-ones mask: 1 where the rounded source
    coordinate is inside the source image."""
    if perspective:
        xn, yn = perspective_fixed_coords(np.asarray(m, np.float64).reshape(3, 3), width, height, nearest=True)
    else:
        xn, yn = affine_fixed_coords(np.asarray(m, np.float64).reshape(2, 3), width, height, nearest=True)
    return (xn >= 0) & (xn < src_w) & (yn >= 0) & (yn < src_h)


# --------------------------------------------------------------------------------------------
# findTransformECC                                   /root/reference/src/lib.rs:769-777
#                                                    [OpenCV video/src/ecc.cpp]
# --------------------------------------------------------------------------------------------
N_PARAMS = {MOTION_TRANSLATION: 2, MOTION_EUCLIDEAN: 3, MOTION_AFFINE: 6, MOTION_HOMOGRAPHY: 8}


def jacobian_rows(motion: int, gx, gy, m32, xg, yg):
    """List of P planes: the per-pixel Jacobian row (image_jacobian_*_ECC)."""
    if motion == MOTION_TRANSLATION:
        return [gx, gy]
    if motion == MOTION_EUCLIDEAN:
        c, s = m32[0, 0], m32[1, 0]
        hat_x = -(xg * s) - (yg * c)
        hat_y = (xg * c) - (yg * s)
        return [gx * hat_x + gy * hat_y, gx, gy]
    if motion == MOTION_AFFINE:
        return [gx * xg, gy * xg, gx * yg, gy * yg, gx, gy]
    h0, h1, h2 = m32[0, 0], m32[1, 0], m32[2, 0]
    h3, h4, h5 = m32[0, 1], m32[1, 1], m32[2, 1]
    h6, h7 = m32[0, 2], m32[1, 2]
    den = xg * h2 + yg * h5 + np.float32(1.0)
    hat_x = (-xg * h0 - yg * h3 - h6) / den
    hat_y = (-xg * h1 - yg * h4 - h7) / den
    a = gx / den
    b = gy / den
    t = hat_x * a + hat_y * b
    return [a * xg, b * xg, t * xg, a * yg, b * yg, t * yg, a, b]


def ecc_sums(motion: int, tmpl: np.ndarray, img: np.ndarray, gxp: np.ndarray, gyp: np.ndarray,
             m32: np.ndarray, acc=np.float64):
    """One iteration's reductions in the single-pass form the CUDA kernel implements.
    Returns dict(n, Sw, Sww, St, Stt, Swt, H[P,P], A[P], Am[P], B[P])."""
    hs, ws = tmpl.shape
    hd, wd = img.shape
    persp = motion == MOTION_HOMOGRAPHY
    w_ = warp_linear(img, m32, ws, hs, persp, inverse_map=True)
    gx = warp_linear(gxp, m32, ws, hs, persp, inverse_map=True)
    gy = warp_linear(gyp, m32, ws, hs, persp, inverse_map=True)
    mask = warp_mask_nearest(m32, ws, hs, wd, hd, persp)
    xg = np.broadcast_to(np.arange(ws, dtype=np.float32)[None, :], (hs, ws))
    yg = np.broadcast_to(np.arange(hs, dtype=np.float32)[:, None], (hs, ws))
    jac = jacobian_rows(motion, gx, gy, np.asarray(m32, np.float32), xg, yg)
    p = len(jac)
    mk = mask.astype(acc)
    wd_, td_ = w_.astype(acc), tmpl.astype(acc)
    out = dict(
        n=mk.sum(), Sw=(mk * wd_).sum(), Sww=(mk * wd_ * wd_).sum(),
        St=(mk * td_).sum(), Stt=(mk * td_ * td_).sum(), Swt=(mk * wd_ * td_).sum(),
        H=np.zeros((p, p)), A=np.zeros(p), Am=np.zeros(p), B=np.zeros(p))
    jd = [j.astype(acc) for j in jac]
    for i in range(p):
        out["A"][i] = (jd[i] * wd_).sum()
        out["Am"][i] = (jd[i] * mk).sum()
        out["B"][i] = (jd[i] * mk * td_).sum()
        for k in range(i, p):
            out["H"][i, k] = out["H"][k, i] = (jd[i] * jd[k]).sum()
    return out


def ecc_epilogue(motion: int, sums: dict, m32: np.ndarray):
    """rho, lambda, delta-p and the matrix update in f64, matrix stored back as f32.
    Raises EccNoConvergence like ecc.cpp (NaN rho; lambda denominator <= 0)."""
    n = sums["n"]
    w_mean, t_mean = sums["Sw"] / n, sums["St"] / n
    img_n2 = sums["Sww"] - sums["Sw"] * sums["Sw"] / n
    tmp_n2 = sums["Stt"] - sums["St"] * sums["St"] / n
    corr = sums["Swt"] - sums["St"] * sums["Sw"] / n
    with np.errstate(all="ignore"):
        rho = corr / math.sqrt(img_n2 * tmp_n2) if img_n2 * tmp_n2 > 0 else float("nan")
    if math.isnan(rho):
        raise EccNoConvergence("NaN encountered.")
    ip = sums["A"] - w_mean * sums["Am"]
    tp = sums["B"] - t_mean * sums["Am"]
    hinv = np.linalg.inv(sums["H"])
    iph = hinv @ ip
    lam_n = img_n2 - ip @ iph
    lam_d = corr - tp @ iph
    if lam_d <= 0.0:
        raise EccNoConvergence("The algorithm stopped before its convergence.")
    lam = lam_n / lam_d
    dp = hinv @ (lam * tp - ip)
    m = np.array(m32, np.float32, copy=True)
    f = np.float32
    if motion == MOTION_TRANSLATION:
        m[0, 2] += f(dp[0]); m[1, 2] += f(dp[1])
    elif motion == MOTION_AFFINE:
        m[0, 0] += f(dp[0]); m[1, 0] += f(dp[1]); m[0, 1] += f(dp[2])
        m[1, 1] += f(dp[3]); m[0, 2] += f(dp[4]); m[1, 2] += f(dp[5])
    elif motion == MOTION_HOMOGRAPHY:
        m[0, 0] += f(dp[0]); m[1, 0] += f(dp[1]); m[2, 0] += f(dp[2]); m[0, 1] += f(dp[3])
        m[1, 1] += f(dp[4]); m[2, 1] += f(dp[5]); m[0, 2] += f(dp[6]); m[1, 2] += f(dp[7])
    else:
        new_theta = f(dp[0]) + f(math.asin(float(m[1, 0])))
        m[0, 2] += f(dp[1]); m[1, 2] += f(dp[2])
        m[0, 0] = m[1, 1] = f(math.cos(float(new_theta)))
        m[1, 0] = f(math.sin(float(new_theta)))
        m[0, 1] = -m[1, 0]
    return rho, m


def find_transform_ecc(tmpl_u8: np.ndarray, img_u8: np.ndarray, motion: int, criteria,
                       gauss_filt_size: int = 5, warp_init=None, acc=np.float64):
    """findTransformECC(templateImage, inputImage, warp, motion, criteria, noArray(), gauss).
    Returns (rho, warp f32 [2x3 | 3x3], iterations)."""
    typ, max_count, epsilon = criteria
    if not (typ & (TERM_COUNT | TERM_EPS)):
        raise ValueError("criteria.type must have COUNT or EPS set")   # CV_Assert in ecc.cpp
    n_iter = max_count if typ & TERM_COUNT else 200
    eps = epsilon if typ & TERM_EPS else -1.0
    rows = 3 if motion == MOTION_HOMOGRAPHY else 2
    m = np.eye(rows, 3, dtype=np.float32) if warp_init is None else np.array(warp_init, np.float32)
    tmpl = gaussian_blur_f32(tmpl_u8.astype(np.float32), gauss_filt_size)
    img = gaussian_blur_f32(img_u8.astype(np.float32), gauss_filt_size)
    gxp, gyp = central_gradients(img)
    rho, last_rho, it = -1.0, -eps, 0
    i = 1
    while i <= n_iter and abs(rho - last_rho) >= eps:
        sums = ecc_sums(motion, tmpl, img, gxp, gyp, m, acc)
        last_rho = rho
        rho, m = ecc_epilogue(motion, sums, m)
        it = i
        i += 1
    return rho, m, it


# --------------------------------------------------------------------------------------------
# ecc_match_no_scaling                               /root/reference/src/lib.rs:719-847
# --------------------------------------------------------------------------------------------
def final_warp(frame_u8: np.ndarray, m, motion: int, border_value=0.0, border_mode: int = BORDER_CONSTANT) -> np.ndarray:
    """convert_to(CV_32F, 1/255) then warp_affine | warp_perspective forward map
    (/root/reference/src/lib.rs:780-803)."""
    h, w = frame_u8.shape[:2]
    return warp_linear(to_f32_unit(frame_u8), m, w, h, perspective=(motion == MOTION_HOMOGRAPHY),
                       inverse_map=False, border_value=border_value, border_mode=border_mode)


def ecc_match(frames_u8, motion: int, max_count, epsilon, gauss_filt_size: int):
    """Returns (stack f32 HxWxC in [0,1], [warp per frame], [iterations per frame])."""
    if len(frames_u8) == 0:
        raise ValueError("NotEnoughFiles")
    crit = term_criteria(max_count, epsilon)
    grey0 = bgr2gray_u8(frames_u8[0])
    acc = to_f32_unit(frames_u8[0])
    warps, iters = [None], [0]
    for fr in frames_u8[1:]:
        _, m, it = find_transform_ecc(bgr2gray_u8(fr), grey0, motion, crit, gauss_filt_size)
        warps.append(m)
        iters.append(it)
        acc = acc + final_warp(fr, m, motion)
    return acc * np.float32(1.0 / len(frames_u8)), warps, iters


# --------------------------------------------------------------------------------------------
# ecc_match_scaling_down                             /root/reference/src/lib.rs:849-1028
#                                                    /root/reference/src/utils.rs:186-248
# --------------------------------------------------------------------------------------------
def scaled_size(width: int, height: int, scale_down: float):
    """utils::scale_image: the SMALLER dimension becomes `scale_down`; both new sizes are truncated."""
    factor = float(scale_down) / float(width if width < height else height)
    return int(width * factor), int(height * factor)


def _area_tab(ssize: int, dsize: int, scale: float):
    """computeResizeAreaTab [OpenCV imgproc/resize.cpp]: per destination index the list of
    (source index, f32 weight), built in f64."""
    tab = [[] for _ in range(dsize)]
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = int(math.ceil(fsx1)), int(math.floor(fsx2))
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab[dx].append((sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab[dx].append((sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab[dx].append((sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def _linear_area_tab(ssize: int, dsize: int):
    """cv::resize's coefficient table for INTER_AREA when it is NOT a pure down-scale ("area mode" of the bilinear
    path, imgproc/resize.cpp): per destination index the first source index and the two 11-bit fixed-point
    weights saturate_cast<short>(w * 2048)."""
    inv = dsize / ssize
    scale = 1.0 / inv
    ofs = np.zeros(dsize, np.int64)
    coef = np.zeros((dsize, 2), np.int64)
    for d in range(dsize):
        s = int(np.floor(d * scale))
        f = np.float32((d + 1) - (s + 1) * inv)
        f = np.float32(0.0) if f <= 0 else np.float32(f - np.floor(f))
        if s < 0:
            f, s = np.float32(0.0), 0
        if s >= ssize - 1:
            f, s = np.float32(0.0), ssize - 1
        ofs[d] = s
        coef[d, 0] = int(np.clip(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048.0))), -32768, 32767))
        coef[d, 1] = int(np.clip(np.rint(np.float32(f * np.float32(2048.0))), -32768, 32767))
    return ofs, coef


def resize_area_up_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=INTER_AREA) when at least one axis ENLARGES (utils::scale_image on
    a landscape frame with height < scale_down_width < width, /root/reference/src/utils.rs:186-214): OpenCV
    emulates "area" with its 8-bit bilinear kernels — horizontal pass in 11-bit fixed point, vertical pass
    ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2."""
    sh, sw = src.shape
    xo, xa = _linear_area_tab(sw, dw)
    yo, ya = _linear_area_tab(sh, dh)
    s = src.astype(np.int64)
    rows = s[:, xo] * xa[None, :, 0] + s[:, np.minimum(xo + 1, sw - 1)] * xa[None, :, 1]
    s0, s1 = rows[yo, :], rows[np.minimum(yo + 1, sh - 1), :]
    b0, b1 = ya[:, 0][:, None], ya[:, 1][:, None]
    out = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def resize_area_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=INTER_AREA) for single-channel 8-bit planes.
    Down-scaling: integer scale factors take OpenCV's fast path (2x2: (sum + 2) >> 2, otherwise
    rint(sum * f32(1/area))); everything else the generic f32 path: per source row buf += S*alpha in table order,
    then sum = beta*buf for the first row of a destination row and sum += beta*buf after, saturate_cast<uchar>.
    An enlarging axis sends the whole call through resize_area_up_u8."""
    sh, sw = src.shape
    if dw > sw or dh > sh:
        return resize_area_up_u8(src, dw, dh)
    scale_x, scale_y = 1.0 / (dw / sw), 1.0 / (dh / sh)     # cv::resize: inv_scale = dsize/ssize; scale = 1/inv_scale
    ix, iy = int(round(scale_x)), int(round(scale_y))
    if abs(scale_x - ix) < np.finfo(np.float64).eps and abs(scale_y - iy) < np.finfo(np.float64).eps:
        s = src[:dh * iy, :dw * ix].reshape(dh, iy, dw, ix).astype(np.int64).sum(axis=(1, 3))
        if ix == 2 and iy == 2:
            return ((s + 2) >> 2).astype(np.uint8)
        return np.clip(np.rint(s.astype(np.float32) * np.float32(1.0 / (ix * iy))), 0, 255).astype(np.uint8)
    xt, yt = _area_tab(sw, dw, scale_x), _area_tab(sh, dh, scale_y)

    def dense(tab):
        k = max(len(t) for t in tab)
        idx = np.zeros((len(tab), k), np.int64)
        wgt = np.zeros((len(tab), k), np.float32)
        for d, t in enumerate(tab):
            for j, (si, a) in enumerate(t):
                idx[d, j], wgt[d, j] = si, a
        return idx, wgt

    xi, xw = dense(xt)
    yi, yw = dense(yt)
    s32 = src.astype(np.float32)
    # horizontal pass for every source row: buf[sy][dx] (adding S*0 for padded entries changes nothing)
    buf = np.zeros((sh, dw), np.float32)
    for j in range(xi.shape[1]):
        buf = buf + s32[:, xi[:, j]] * xw[None, :, j]
    out = np.zeros((dh, dw), np.float32)
    for j in range(yi.shape[1]):
        out = out + yw[:, j, None] * buf[yi[:, j], :]
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def rescale_warp(m_small: np.ndarray, motion: int, full_size, small_size) -> np.ndarray:
    """Full-resolution matrix from the one estimated on the downscaled greys, with the reference's f32
    arithmetic: non-homography scales ONLY the translation column (src/lib.rs:941-951); homography goes
    through adjust_homography_for_scale_f32 (src/utils.rs:218-248)."""
    fw, fh = full_size
    sw, sh = small_size
    m = np.array(m_small, np.float32, copy=True)
    if motion != MOTION_HOMOGRAPHY:
        m[0, 2] *= np.float32(fw) / np.float32(sw)
        m[1, 2] *= np.float32(fh) / np.float32(sh)
        return m
    sx, sy = np.float32(fw / sw), np.float32(fh / sh)
    m[0, 2] *= sx
    m[1, 2] *= sy
    m[2, 0] /= sx
    m[2, 1] /= sy
    return m


def ecc_match_scaling_down(frames_u8, motion: int, max_count, epsilon, gauss_filt_size: int, scale_down: float):
    """Returns (stack f32, [full-resolution warp per frame], [iterations per frame])."""
    if len(frames_u8) == 0:
        raise ValueError("NotEnoughFiles")
    h, w = frames_u8[0].shape[:2]
    if scale_down >= w:
        raise ValueError("InvalidParams: scale_down_to was larger (or equal) to the full image width")
    if scale_down <= 10.0:
        raise ValueError("InvalidParams: scale_down_to was too small")
    crit = term_criteria(max_count, epsilon)
    sw, sh = scaled_size(w, h, scale_down)
    grey0 = resize_area_u8(bgr2gray_u8(frames_u8[0]), sw, sh)
    acc = to_f32_unit(frames_u8[0])
    warps, iters = [None], [0]
    for fr in frames_u8[1:]:
        small = resize_area_u8(bgr2gray_u8(fr), sw, sh)
        _, m, it = find_transform_ecc(small, grey0, motion, crit, gauss_filt_size)
        mf = rescale_warp(m, motion, (w, h), (sw, sh))
        warps.append(mf)
        iters.append(it)
        acc = acc + final_warp(fr, mf, motion)
    return acc * np.float32(1.0 / len(frames_u8)), warps, iters


# --------------------------------------------------------------------------------------------
# sharpness_tenengrad                                /root/reference/src/lib.rs:1101-1147
# --------------------------------------------------------------------------------------------
SOBEL_KERNELS = {            # getDerivKernels(dx=1, dy=0, ksize): (derivative taps, smoothing taps)
    1: ([-1, 0, 1], [1]),
    3: ([-1, 0, 1], [1, 2, 1]),
    5: ([-1, -2, 0, 2, 1], [1, 4, 6, 4, 1]),
    7: ([-1, -4, -5, 0, 5, 4, 1], [1, 6, 15, 20, 15, 6, 1]),
}


def _sep_corr_int(img: np.ndarray, kx, ky) -> np.ndarray:
    h, w = img.shape
    rx, ry = len(kx) // 2, len(ky) // 2
    cols = _reflect101(np.arange(-rx, w + rx), w)
    rows = _reflect101(np.arange(-ry, h + ry), h)
    pad = img[:, cols]
    tmp = np.zeros((h, w), np.int64)
    for k, c in enumerate(kx):
        tmp += c * pad[:, k:k + w]
    pad = tmp[rows, :]
    out = np.zeros((h, w), np.int64)
    for k, c in enumerate(ky):
        out += c * pad[k:k + h, :]
    return out


def sharpness_tenengrad(grey_u8: np.ndarray, ksize: int) -> float:
    """Sobel dx, dy in CV_64F, mean(gx^2 + gy^2).  All values are integers < 2^53 so the f64 result is
    exact; cv::mean multiplies it by the rounded reciprocal: sum * (1.0 / N)."""
    if ksize not in SOBEL_KERNELS:
        raise ValueError("Kernel size must be 1, 3, 5, or 7")      # StackerError::InvalidParams
    d, s = SOBEL_KERNELS[ksize]
    g = grey_u8.astype(np.int64)
    gx = _sep_corr_int(g, d, s)
    gy = _sep_corr_int(g, s, d)
    total = int((gx * gx + gy * gy).sum())
    return float(total) * (1.0 / float(grey_u8.size))


# --------------------------------------------------------------------------------------------
# the other three sharpness metrics                  /root/reference/src/lib.rs:1032-1090, :1151-1166
# (SURVEY §8(f) N3).  On 8-bit input every intermediate is an integer (or a multiple of 1/4), so the
# sums are exact in f64 and only the last few scalar operations round — restated in cv2's order.
# --------------------------------------------------------------------------------------------
def _replicate(idx: np.ndarray, n: int) -> np.ndarray:
    return np.clip(idx, 0, n - 1)


def _mean_std_from_sums(s: int, sq: int, n: int):
    """cv::meanStdDev's scalar tail: scale = 1/N; mean = s*scale; sigma = sqrt(max(sq*scale - mean^2, 0))."""
    scale = 1.0 / float(n)
    mean = float(s) * scale
    var = max(float(sq) * scale - mean * mean, 0.0)
    return mean, math.sqrt(var)


def sharpness_modified_laplacian(grey_u8: np.ndarray) -> float:
    """LAPM (src/lib.rs:1032-1068): sepFilter2D with [-1 2 -1] along one axis and getGaussianKernel(3,-1) =
    [1 2 1]/4 along the other, BORDER_REFLECT_101, CV_64F; mean(|lx| + |ly|).  4*lx, 4*ly are integers."""
    g = grey_u8.astype(np.int64)
    lx4 = _sep_corr_int(g, [-1, 2, -1], [1, 2, 1])
    ly4 = _sep_corr_int(g, [1, 2, 1], [-1, 2, -1])
    total4 = int((np.abs(lx4) + np.abs(ly4)).sum())
    return (float(total4) / 4.0) * (1.0 / float(grey_u8.size))


def laplacian3_int(grey_u8: np.ndarray) -> np.ndarray:
    """cv::Laplacian(ksize=3, scale 1, BORDER_REPLICATE): kernel [[2 0 2], [0 -8 0], [2 0 2]]."""
    h, w = grey_u8.shape
    g = grey_u8.astype(np.int64)
    pad = g[_replicate(np.arange(-1, h + 1), h)][:, _replicate(np.arange(-1, w + 1), w)]
    return 2 * (pad[:-2, :-2] + pad[:-2, 2:] + pad[2:, :-2] + pad[2:, 2:]) - 8 * pad[1:-1, 1:-1]


def sharpness_variance_of_laplacian(grey_u8: np.ndarray) -> float:
    """LAPV (src/lib.rs:1070-1090): meanStdDev of the CV_64F Laplacian, sigma^2."""
    lap = laplacian3_int(grey_u8)
    _, sigma = _mean_std_from_sums(int(lap.sum()), int((lap * lap).sum()), lap.size)
    return sigma * sigma


def sharpness_normalized_gray_level_variance(grey_u8: np.ndarray) -> float:
    """GLVN (src/lib.rs:1151-1166): sigma^2 / max(mu, f64::EPSILON) of the image itself."""
    g = grey_u8.astype(np.int64)
    mu, sigma = _mean_std_from_sums(int(g.sum()), int((g * g).sum()), g.size)
    return (sigma ** 2) / max(mu, float(np.finfo(np.float64).eps))


def rank_by_sharpness(values):
    """examples/main.rs:53-64: sort ascending by Tenengrad, drop the worst, reverse (sharpest first).
    Returns the frame indices in stacking order."""
    order = sorted(range(len(values)), key=lambda i: values[i])
    return list(reversed(order[1:]))
