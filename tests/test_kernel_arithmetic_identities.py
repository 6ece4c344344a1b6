"""The integer / exponent tricks the streaming kernels rely on, restated in NumPy and checked exhaustively on the CPU
(no GPU needed).  Each one is a claim DESIGN.md makes about why a kernel is bit-identical to OpenCV."""
import numpy as np

f32 = np.float32


def test_grey_weights_split_into_byte_dot_products():
    """K1 (csrc/prep.cuh, grey_pairs): cvtColor's 15-bit weights are split into bytes so that one pixel costs two 4-way
    byte dot products: 3735 = 14*256 + 151, 19235 = 75*256 + 35, 9798 = 38*256 + 70.  For every (B, G, R):
    (dp4a_hi << 8) + dp4a_lo == 3735 B + 19235 G + 9798 R, and the rounded shift is OpenCV's RGB2Gray<uchar>."""
    assert (14 * 256 + 151, 75 * 256 + 35, 38 * 256 + 70) == (3735, 19235, 9798)
    rng = np.random.default_rng(0)
    bgr = np.concatenate([rng.integers(0, 256, (200000, 3)), np.array([[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255]])])
    b, g, r = bgr[:, 0].astype(np.int64), bgr[:, 1].astype(np.int64), bgr[:, 2].astype(np.int64)
    hi = 14 * b + 75 * g + 38 * r
    lo = 151 * b + 35 * g + 70 * r
    full = 3735 * b + 19235 * g + 9798 * r
    assert np.array_equal((hi << 8) + lo, full)
    assert hi.max() < 2 ** 16 and lo.max() < 2 ** 16                        # each dot product fits 16 bits
    grey = (full + 16384) >> 15
    assert grey.max() == 255 and grey.min() == 0                             # 3735 + 19235 + 9798 == 2^15: no overflow
    assert 3735 + 19235 + 9798 == 1 << 15


def test_blurred_value_splice_is_exact():
    """K1: with 8-bit input and the dyadic taps [1 4 6 4 1]/16 (k = 5) resp. [1 2 1]/4 (k = 3) the blurred pixel is v / 256
    resp. v / 16 with an integer v below 2^16 resp. 2^12; the f32 value comes from splicing v under an exponent
    (0x47000000 | v is 32768 + v/256; 0x49000000 | v is 524288 + v/16) and one subtraction — exact for every v."""
    v = np.arange(1 << 16, dtype=np.uint32)
    assert (16 * 255) * 16 < (1 << 16)                                       # k = 5: sum of taps 16 * 16, max value
    sp = (np.uint32(0x47000000) | v).view(f32)
    assert np.array_equal(sp - f32(32768.0), (v / 256.0).astype(f32))
    assert np.array_equal((sp - f32(32768.0)).astype(np.float64), v / 256.0)  # and v/256 is exactly representable
    v3 = np.arange(4 * 255 * 4 + 1, dtype=np.uint32)                          # k = 3: taps sum 4 * 4
    sp3 = (np.uint32(0x49000000) | v3).view(f32)
    assert np.array_equal((sp3 - f32(524288.0)).astype(np.float64), v3 / 16.0)


def test_packed_16bit_blur_never_carries():
    """K1 filters two pixels per 32-bit register: the horizontal and the vertical [1 4 6 4 1] pass on 8-bit input stay
    below 2^16 per half (16 * 255 after one pass, 256 * 255 after both), so the halves never carry into each other."""
    assert 16 * 255 < 1 << 16 and 256 * 255 < 1 << 16
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (5, 1000, 2)).astype(np.uint64)
    taps = np.array([1, 4, 6, 4, 1], dtype=np.uint64)
    packed = a[..., 0] | (a[..., 1] << np.uint64(16))
    acc = (packed * taps[:, None]).sum(axis=0)
    lo, hi = acc & np.uint64(0xFFFF), acc >> np.uint64(16)
    assert np.array_equal(lo, (a[..., 0] * taps[:, None]).sum(axis=0))
    assert np.array_equal(hi, (a[..., 1] * taps[:, None]).sum(axis=0))


def test_tenengrad_band_sum_fits_32_bits():
    """K6 (csrc/tenengrad.cuh): a thread accumulates gx^2 + gy^2 of 48 rows x 16 columns in 32 bits before widening.
    |gx|, |gy| <= 4 * 255 for the 3x3 Sobel on 8-bit input."""
    assert 48 * 16 * 2 * (4 * 255) ** 2 < 1 << 32


def test_fastpersp_coordinates_move_few_quanta():
    """K2 (csrc/ecc_iter.cuh, FastPersp): Homography sample positions are evaluated as x + (alpha + beta y) / w in f32 from
    per-column constants prepared in f64, then quantised to 1/32 px.  DESIGN states that against OpenCV's f64 evaluation
    this moves the quantum for ~1e-4 of the pixels, by one step: emulate the f32 arithmetic and count."""
    def fma32(a, b, c):
        return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)

    rng = np.random.default_rng(4)
    w, h, n = 3840, 2160, 20000
    moved = total = 0
    worst = 0
    for _ in range(40):
        m = (np.eye(3) + rng.uniform(-1, 1, (3, 3)) * np.array([[2e-3, 2e-3, 6.0], [2e-3, 2e-3, 6.0], [2e-7, 2e-7, 0.0]])).astype(f32)
        md = m.astype(np.float64)                     # the kernel reads the f32 matrix and prepares constants in f64
        xs = rng.integers(0, w, n).astype(np.float64)
        ys = rng.integers(0, h, n).astype(np.float64)
        alpha = (xs * (md[0, 0] - md[2, 2] - md[2, 0] * xs) + md[0, 2]).astype(f32)
        beta = (md[0, 1] - md[2, 1] * xs).astype(f32)
        wc = (md[2, 0] * xs + md[2, 2]).astype(f32)
        yf = ys.astype(f32)
        wf = fma32(np.full(n, m[2, 1]), yf, wc)
        rw = (1.0 / wf.astype(np.float64)).astype(f32)
        du = (fma32(beta, yf, alpha).astype(np.float64) * rw.astype(np.float64)).astype(f32)
        q_fast = np.rint(du.astype(np.float64) * 32.0) + 32.0 * xs
        u = (md[0, 0] * xs + md[0, 1] * ys + md[0, 2]) / (md[2, 0] * xs + md[2, 1] * ys + md[2, 2])
        q_exact = np.rint(32.0 * u)
        d = np.abs(q_fast - q_exact)
        moved += int((d != 0).sum())
        worst = max(worst, int(d.max()))
        total += n
    assert worst <= 1                                  # never more than one 1/32-px step
    assert moved / total < 1e-3                        # measured ~1e-4
