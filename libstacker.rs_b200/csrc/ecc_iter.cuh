// K2 — one ECC iteration, fused: inverse-map warp of the blurred reference image I and of its two
// central-difference gradient planes (derived on the fly from I, never stored), nearest-neighbour mask,
// per-pixel Jacobian for the four motion models, and every reduction the update needs, accumulated in
// registers -> warp shuffles -> block -> per-tile partials (f64) -> the last block to finish sums the
// partials in a fixed order, solves the PxP normal equations in f64, updates the f32 warp matrix, runs the
// convergence test and sets the CUDA-graph WHILE condition.  No per-pixel plane is written and no
// iteration round-trips to the host.
//
// Replaces the body of OpenCV's findTransformECC loop (modules/video/src/ecc.cpp), which the reference
// reaches through opencv::video::find_transform_ecc at /root/reference/src/lib.rs:769-777
// (template = frame i, input = frame 0, identity init, no mask).  Restated on the CPU in
// oracle/restate.py (ecc_sums / ecc_epilogue), which is pinned against cv2.findTransformECC.
//
// Work decomposition ("column owner"): a tile is 128 columns x R rows; a 256-thread block owns one tile,
// warps 0-3 take the upper half of the rows, warps 4-7 the lower half, and every thread keeps ONE column
// x for its whole row range.  Because X is constant per thread, the Kronecker structure of the affine /
// homography Jacobians ( J = g (x) [X, Y, 1] ) lets a pixel accumulate only g_i*g_j*{1,Y,Y^2} and
// g_i*z*{1,Y}; the X factors are folded in once per tile.  Algorithmic traffic: 4N (T) + 4N (I) bytes
// per iteration.
#pragma once
#include "common.cuh"

namespace stk {

constexpr int kEccThreads = 256;
constexpr int kEccStripW = 128;     // columns per tile
constexpr int kEccRowParts = 2;     // row halves per tile (kEccThreads / kEccStripW)

// status values written by the device loop (== stacker_cuda.h STK_* codes)
constexpr int kStatusOk = 0, kStatusNoConv = 4, kStatusNaN = 5;

struct EccState {
  float m[9];               // current warp, row-major 3x3 (rows 0-1 used by the 2x3 models)
  double inv[9];            // inverse map for the final forward warp (3x3, or 2x3 in [0..5])
  double rho, last_rho;
  double eps;
  int max_iter;
  int iters;
  int status;
  int cont;                 // 1 while the loop should run another iteration
  unsigned int tile_counter;
  int pad;
};

struct EccIterParams {
  const float* img;         // I  : blurred reference (frame 0), f32
  const float* tmpl;        // T  : blurred current frame, f32
  int pitch;                // floats, both planes
  int width, height;        // template size == image size on this path
  int rows_per_tile;        // R (even)
  int n_strips, n_bands;
  double* partials;         // [n_tiles][NV]
  EccState* st;
  cudaGraphConditionalHandle handle;
  int use_handle;
  double* totals_out;       // optional: the NV reduced sums of this iteration (test hook), else null
};

// ---- per-model constants -------------------------------------------------------------------------
template <int MOTION> struct Model;
template <> struct Model<kTranslation> { static constexpr int P = 2, G = 2; static constexpr bool kron = false, persp = false; };
template <> struct Model<kEuclidean>   { static constexpr int P = 3, G = 3; static constexpr bool kron = false, persp = false; };
template <> struct Model<kAffine>      { static constexpr int P = 6, G = 2; static constexpr bool kron = true,  persp = false; };
template <> struct Model<kHomography>  { static constexpr int P = 8, G = 3; static constexpr bool kron = true,  persp = true;  };

template <int MOTION> struct Layout {
  using M = Model<MOTION>;
  static constexpr int G = M::G;
  static constexpr int NP = G * (G + 1) / 2;                 // products g_i g_j
  static constexpr int QM = M::kron ? 6 : 1;                 // moments {1,X,Y,XX,XY,YY} | {1}
  static constexpr int ZM = M::kron ? 3 : 1;                 // moments {1,X,Y} | {1}
  static constexpr int kScal = 6;                            // n Sw Sww St Stt Swt
  static constexpr int kH = kScal;                           // H block  [NP][QM]
  static constexpr int kZ = kH + NP * QM;                    // proj block [3 z][G][ZM]
  static constexpr int NV = kZ + 3 * G * ZM;
};

// ---- sampling ------------------------------------------------------------------------------------
struct Sample { float w, gx2, gy2; };   // bilinear I, 2*bilinear(GX), 2*bilinear(GY)

struct Taps {     // the 12 values of I a bilinear sample of (I, GX, GY) touches
  float m0, m1;             // row sy-1 : cols sx, sx+1
  float a_1, a0, a1, a2;    // row sy   : cols sx-1 .. sx+2
  float b_1, b0, b1, b2;    // row sy+1
  float c0, c1;             // row sy+2 : cols sx, sx+1
};

__device__ __forceinline__ Taps load_taps(const float* __restrict__ img, int pitch, int sx, int sy) {
  const float* p = img + (ptrdiff_t)sy * pitch + sx;
  Taps t;
  t.m0 = __ldg(p - pitch);     t.m1 = __ldg(p - pitch + 1);
  t.a_1 = __ldg(p - 1);        t.a0 = __ldg(p);             t.a1 = __ldg(p + 1);         t.a2 = __ldg(p + 2);
  t.b_1 = __ldg(p + pitch - 1); t.b0 = __ldg(p + pitch);    t.b1 = __ldg(p + pitch + 1); t.b2 = __ldg(p + pitch + 2);
  t.c0 = __ldg(p + 2 * pitch); t.c1 = __ldg(p + 2 * pitch + 1);
  return t;
}

__device__ __forceinline__ float lerp(float a, float b, float t) { return fmaf(t, b - a, a); }

__device__ __forceinline__ Sample interp(const Taps& t, float ax, float ay) {
  Sample s;
  s.w = lerp(lerp(t.a0, t.a1, ax), lerp(t.b0, t.b1, ax), ay);
  // GX taps (2*GX = I[c+1] - I[c-1]) at (sy,sx) (sy,sx+1) (sy+1,sx) (sy+1,sx+1)
  s.gx2 = lerp(lerp(t.a1 - t.a_1, t.a2 - t.a0, ax), lerp(t.b1 - t.b_1, t.b2 - t.b0, ax), ay);
  // GY taps (2*GY = I[r+1] - I[r-1])
  s.gy2 = lerp(lerp(t.b0 - t.m0, t.b1 - t.m1, ax), lerp(t.c0 - t.a0, t.c1 - t.a1, ax), ay);
  return s;
}

// general path: every tap obeys the plane's own rule — I, GX, GY are 0 outside the image (constant
// border), GX is 0 on the first/last column, GY on the first/last row (filter2D + BORDER_REFLECT_101).
__device__ __noinline__ Sample sample_general(const float* __restrict__ img, int pitch, int w, int h,
                                              int sx, int sy, float ax, float ay) {
  float vi[4], vx[4], vy[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = sx + (k & 1), r = sy + (k >> 1);
    const bool in = (unsigned)c < (unsigned)w && (unsigned)r < (unsigned)h;
    float i0 = 0.f, dx = 0.f, dy = 0.f;
    if (in) {
      const float* q = img + (ptrdiff_t)r * pitch + c;
      i0 = __ldg(q);
      if (c >= 1 && c <= w - 2) dx = __ldg(q + 1) - __ldg(q - 1);
      if (r >= 1 && r <= h - 2) dy = __ldg(q + pitch) - __ldg(q - pitch);
    }
    vi[k] = i0; vx[k] = dx; vy[k] = dy;
  }
  Sample s;
  s.w = lerp(lerp(vi[0], vi[1], ax), lerp(vi[2], vi[3], ax), ay);
  s.gx2 = lerp(lerp(vx[0], vx[1], ax), lerp(vx[2], vx[3], ax), ay);
  s.gy2 = lerp(lerp(vy[0], vy[1], ax), lerp(vy[2], vy[3], ax), ay);
  return s;
}

// ---- coordinates ----------------------------------------------------------------------------------
// Per-thread (fixed column x) constants and per-row evaluation of the quantised source position
// (Xq, Yq in 1/32 px) exactly as OpenCV's WARP_INVERSE_MAP warps compute it: f64 projective divide
// + round-half-even for warpPerspective, 10-bit fixed point for warpAffine.
template <bool PERSP> struct Coord;

template <> struct Coord<true> {
  double cx, cy, cw, m01, m11, m21;
  __device__ __forceinline__ void init(const float* m, int x) {
    const double xd = (double)x;
    cx = fma((double)m[0], xd, (double)m[2]);
    cy = fma((double)m[3], xd, (double)m[5]);
    cw = fma((double)m[6], xd, (double)m[8]);
    m01 = (double)m[1]; m11 = (double)m[4]; m21 = (double)m[7];
  }
  // returns false when the position is not representable (treated as outside the image)
  __device__ __forceinline__ bool at(int y, int& xq, int& yq) const {
    const double yd = (double)y;
    const double w = fma(m21, yd, cw);
    const double rw = (w != 0.0) ? __drcp_rn(w) : 0.0;
    const double fx = fma(m01, yd, cx) * rw;
    const double fy = fma(m11, yd, cy) * rw;
    xq = rint_magic_scaled(fx, 32.0);
    yq = rint_magic_scaled(fy, 32.0);
    return coord_in_range(fx) && coord_in_range(fy);
  }
  // same, plus the separately rounded INTER_NEAREST coordinate (round half even of u, v) for the mask
  __device__ __forceinline__ bool at_with_nearest(int y, int& xq, int& yq, int& xn, int& yn) const {
    const double yd = (double)y;
    const double w = fma(m21, yd, cw);
    const double rw = (w != 0.0) ? __drcp_rn(w) : 0.0;
    const double fx = fma(m01, yd, cx) * rw;
    const double fy = fma(m11, yd, cy) * rw;
    xq = rint_magic_scaled(fx, 32.0);
    yq = rint_magic_scaled(fy, 32.0);
    xn = rint_magic(fx);
    yn = rint_magic(fy);
    return coord_in_range(fx) && coord_in_range(fy);
  }
};

template <> struct Coord<false> {
  double m01, m02, m11, m12;
  int adelta, bdelta;
  __device__ __forceinline__ void init(const float* m, int x) {
    const double xd = (double)x;
    adelta = rint_magic(__dmul_rn(__dmul_rn((double)m[0], xd), kAbScale));
    bdelta = rint_magic(__dmul_rn(__dmul_rn((double)m[3], xd), kAbScale));
    m01 = (double)m[1]; m02 = (double)m[2]; m11 = (double)m[4]; m12 = (double)m[5];
  }
  __device__ __forceinline__ bool at(int y, int& xq, int& yq) const {
    const double yd = (double)y;
    const double fx = __dadd_rn(__dmul_rn(m01, yd), m02);
    const double fy = __dadd_rn(__dmul_rn(m11, yd), m12);
    // X0 = rint(fx * 1024) + 16 ; X = (X0 + adelta) >> 5    (round_delta = AB_SCALE/INTER_TAB_SIZE/2)
    xq = (rint_magic_scaled(fx, kAbScale) + 16 + adelta) >> (kAbBits - kInterBits);
    yq = (rint_magic_scaled(fy, kAbScale) + 16 + bdelta) >> (kAbBits - kInterBits);
    return (fabs(fx) < 1.0e6) && (fabs(fy) < 1.0e6);
  }
  // INTER_NEAREST: round_delta = AB_SCALE/2, shift by AB_BITS
  __device__ __forceinline__ bool at_with_nearest(int y, int& xq, int& yq, int& xn, int& yn) const {
    const double yd = (double)y;
    const double fx = __dadd_rn(__dmul_rn(m01, yd), m02);
    const double fy = __dadd_rn(__dmul_rn(m11, yd), m12);
    const int rx = rint_magic_scaled(fx, kAbScale), ry = rint_magic_scaled(fy, kAbScale);
    xq = (rx + 16 + adelta) >> (kAbBits - kInterBits);
    yq = (ry + 16 + bdelta) >> (kAbBits - kInterBits);
    xn = (rx + 512 + adelta) >> kAbBits;
    yn = (ry + 512 + bdelta) >> kAbBits;
    return (fabs(fx) < 1.0e6) && (fabs(fy) < 1.0e6);
  }
};

// ---- interior test ---------------------------------------------------------------------------------
// A tile is "interior" when all of its pixels sample 2 px inside the image, so no tap needs a border
// rule and the mask is all ones.  The image of a rectangle under an affine map, or under a projective
// map with w > 0 on its four corners, is the convex hull of the corner images.
template <bool PERSP>
__device__ bool tile_interior(const float* m, int x0, int x1, int y0, int y1, int w, int h) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double x = (k & 1) ? (double)x1 : (double)x0;
    const double y = (k & 2) ? (double)y1 : (double)y0;
    double u = (double)m[0] * x + (double)m[1] * y + (double)m[2];
    double v = (double)m[3] * x + (double)m[4] * y + (double)m[5];
    if (PERSP) {
      const double ww = (double)m[6] * x + (double)m[7] * y + (double)m[8];
      if (!(ww > 1e-9)) { ok = false; continue; }
      u /= ww; v /= ww;
    }
    ok = ok && (u >= 2.0) && (u <= (double)(w - 3)) && (v >= 2.0) && (v <= (double)(h - 3));
  }
  return ok;
}

// ---- accumulation ----------------------------------------------------------------------------------
template <int MOTION> struct Accum {
  using L = Layout<MOTION>;
  static constexpr int G = L::G, NP = L::NP;
  static constexpr bool kron = Model<MOTION>::kron;
  static constexpr int PY = kron ? 3 : 1;     // {1, Y, Y^2}
  static constexpr int ZY = kron ? 2 : 1;     // {1, Y}
  float n, sw, sww, st, stt, swt;
  float p[NP][PY];
  float z[3][G][ZY];

  __device__ __forceinline__ void clear() {
    n = sw = sww = st = stt = swt = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i)
#pragma unroll
      for (int k = 0; k < PY; ++k) p[i][k] = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int i = 0; i < G; ++i)
#pragma unroll
        for (int k = 0; k < ZY; ++k) z[a][i][k] = 0.f;
  }

  // g: Jacobian generators, w_: warped image, t_: template, mk: mask (0/1), yf: row as float
  template <bool INTERIOR>
  __device__ __forceinline__ void add(const float (&g)[G], float w_, float t_, float mk, float yf) {
    const float yy = yf * yf;
    int idx = 0;
#pragma unroll
    for (int i = 0; i < G; ++i)
#pragma unroll
      for (int j = i; j < G; ++j) {
        const float pr = g[i] * g[j];
        p[idx][0] += pr;
        if (kron) { p[idx][1] = fmaf(pr, yf, p[idx][1]); p[idx][2] = fmaf(pr, yy, p[idx][2]); }
        ++idx;
      }
    const float wm = INTERIOR ? w_ : w_ * mk;
    const float tm = INTERIOR ? t_ : t_ * mk;
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const float gw = g[i] * w_;                      // A  : unmasked
      const float gm = INTERIOR ? g[i] : g[i] * mk;    // Am : masked
      const float gt = g[i] * tm;                      // B  : masked
      z[0][i][0] += gw; z[1][i][0] += gm; z[2][i][0] += gt;
      if (kron) {
        z[0][i][1] = fmaf(gw, yf, z[0][i][1]);
        z[1][i][1] = fmaf(gm, yf, z[1][i][1]);
        z[2][i][1] = fmaf(gt, yf, z[2][i][1]);
      }
    }
    n += INTERIOR ? 1.f : mk;
    sw += wm; sww = fmaf(wm, w_, sww);
    st += tm; stt = fmaf(tm, t_, stt);
    swt = fmaf(wm, t_, swt);
  }

  // fold the per-thread column coordinate X into the moments and emit the NV values of this thread
  __device__ __forceinline__ void emit(float xf, float (&v)[L::NV]) const {
    v[0] = n; v[1] = sw; v[2] = sww; v[3] = st; v[4] = stt; v[5] = swt;
    const float xx = xf * xf;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      if (kron) {
        float* o = &v[L::kH + i * 6];
        o[0] = p[i][0]; o[1] = xf * p[i][0]; o[2] = p[i][1];
        o[3] = xx * p[i][0]; o[4] = xf * p[i][1]; o[5] = p[i][2];
      } else {
        v[L::kH + i] = p[i][0];
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int i = 0; i < G; ++i) {
        if (kron) {
          float* o = &v[L::kZ + (a * G + i) * 3];
          o[0] = z[a][i][0]; o[1] = xf * z[a][i][0]; o[2] = z[a][i][1];
        } else {
          v[L::kZ + a * G + i] = z[a][i][0];
        }
      }
  }
};

// ---- epilogue: totals -> normal equations -> update -------------------------------------------------
// Assemble H (PxP), A, Am, B from the reduced totals `t` (f64, layout of Layout<MOTION>).
template <int MOTION>
__device__ void assemble(const double* t, double (*hm)[8], double* a, double* am, double* b) {
  using L = Layout<MOTION>;
  constexpr int P = Model<MOTION>::P, G = L::G;
  auto pidx = [](int i, int j) {  // index of product g_i g_j (i <= j) in the upper-triangular enumeration
    if (i > j) { int s = i; i = j; j = s; }
    return i * G - i * (i - 1) / 2 + (j - i);
  };
  if (!Model<MOTION>::kron) {
    for (int i = 0; i < P; ++i) {
      for (int j = 0; j < P; ++j) hm[i][j] = t[L::kH + pidx(i, j)];
      a[i] = t[L::kZ + 0 * G + i]; am[i] = t[L::kZ + 1 * G + i]; b[i] = t[L::kZ + 2 * G + i];
    }
  } else {
    // J_k = g[k % G] * q[k / G],  q = (X, Y, 1)
    // pair moment index in {1, X, Y, XX, XY, YY}: (X,X)=3 (X,Y)=4 (X,1)=1 (Y,Y)=5 (Y,1)=2 (1,1)=0
    const int pm[3][3] = {{3, 4, 1}, {4, 5, 2}, {1, 2, 0}};
    const int zm[3] = {1, 2, 0};
    for (int k = 0; k < P; ++k) {
      const int gk = k % G, qk = k / G;
      for (int l = 0; l < P; ++l) {
        const int gl = l % G, ql = l / G;
        hm[k][l] = t[L::kH + pidx(gk, gl) * 6 + pm[qk][ql]];
      }
      a[k] = t[L::kZ + (0 * G + gk) * 3 + zm[qk]];
      am[k] = t[L::kZ + (1 * G + gk) * 3 + zm[qk]];
      b[k] = t[L::kZ + (2 * G + gk) * 3 + zm[qk]];
    }
  }
}

// inverse map used by the final forward warp: exactly OpenCV's arithmetic (no FMA contraction).
__device__ inline void compute_inverse(EccState* st, bool persp) {
  double s[9];
  for (int i = 0; i < 9; ++i) s[i] = (double)st->m[i];
  double* o = st->inv;
  if (!persp) {
    double d = __dsub_rn(__dmul_rn(s[0], s[4]), __dmul_rn(s[1], s[3]));
    d = (d != 0.0) ? __ddiv_rn(1.0, d) : 0.0;
    const double a11 = __dmul_rn(s[4], d), a22 = __dmul_rn(s[0], d);
    const double i00 = a11, i01 = __dmul_rn(s[1], -d), i10 = __dmul_rn(s[3], -d), i11 = a22;
    o[0] = i00; o[1] = i01; o[3] = i10; o[4] = i11;
    o[2] = __dsub_rn(__dmul_rn(-i00, s[2]), __dmul_rn(i01, s[5]));
    o[5] = __dsub_rn(__dmul_rn(-i10, s[2]), __dmul_rn(i11, s[5]));
    o[6] = 0.0; o[7] = 0.0; o[8] = 1.0;
  } else {
    auto det2 = [](double a, double b, double c, double d) { return __dsub_rn(__dmul_rn(a, d), __dmul_rn(b, c)); };
    // det = s00*(s11 s22 - s12 s21) - s01*(s10 s22 - s12 s20) + s02*(s10 s21 - s11 s20)
    double d = __dadd_rn(__dsub_rn(__dmul_rn(s[0], det2(s[4], s[5], s[7], s[8])),
                                   __dmul_rn(s[1], det2(s[3], s[5], s[6], s[8]))),
                         __dmul_rn(s[2], det2(s[3], s[4], s[6], s[7])));
    if (d == 0.0) { for (int i = 0; i < 9; ++i) o[i] = 0.0; return; }
    d = __ddiv_rn(1.0, d);
    o[0] = __dmul_rn(det2(s[4], s[5], s[7], s[8]), d);
    o[1] = __dmul_rn(det2(s[2], s[1], s[8], s[7]), d);
    o[2] = __dmul_rn(det2(s[1], s[2], s[4], s[5]), d);
    o[3] = __dmul_rn(det2(s[5], s[3], s[8], s[6]), d);
    o[4] = __dmul_rn(det2(s[0], s[2], s[6], s[8]), d);
    o[5] = __dmul_rn(det2(s[2], s[0], s[5], s[3]), d);
    o[6] = __dmul_rn(det2(s[3], s[4], s[6], s[7]), d);
    o[7] = __dmul_rn(det2(s[1], s[0], s[7], s[6]), d);
    o[8] = __dmul_rn(det2(s[0], s[1], s[3], s[4]), d);
  }
}

// One thread: f64 solve of the P x P SPD system with two right-hand sides (Gaussian elimination with
// partial pivoting on a shared-memory scratch), lambda, delta-p, matrix update, convergence test.
template <int MOTION>
__device__ __noinline__ void ecc_solve_and_update(const double* tot, EccState* st, double (*hm)[8],
                                                  double* ip, double* tp) {
  constexpr int P = Model<MOTION>::P;
  double* am = ip + 16;     // scratch laid out by the caller: ip[8] tp[8] am[8] (ip+16)
  assemble<MOTION>(tot, hm, ip, am, tp);   // ip <- A, tp <- B for now
  const double n = tot[0], sw = tot[1], sww = tot[2], s_t = tot[3], stt = tot[4], swt = tot[5];
  const double wbar = sw / n, tbar = s_t / n;
  const double in2 = sww - sw * sw / n;
  const double tn2 = stt - s_t * s_t / n;
  const double corr = swt - s_t * sw / n;
  const double rho = corr / sqrt(in2 * tn2);
  st->last_rho = st->rho;
  st->rho = rho;
  st->iters += 1;
  if (!(rho == rho)) {              // NaN -> cv::Error::StsNoConv "NaN encountered."
    st->status = kStatusNaN; st->cont = 0; return;
  }
  for (int k = 0; k < P; ++k) { ip[k] = ip[k] - wbar * am[k]; tp[k] = tp[k] - tbar * am[k]; }
  // solve H [y1 y2] = [ip tp]
  double y1[8], y2[8];
  bool singular = false;
  {
    double r1[8], r2[8];
    for (int k = 0; k < P; ++k) { r1[k] = ip[k]; r2[k] = tp[k]; }
    for (int k = 0; k < P; ++k) {
      int piv = k; double best = fabs(hm[k][k]);
      for (int i = k + 1; i < P; ++i) { const double v = fabs(hm[i][k]); if (v > best) { best = v; piv = i; } }
      if (!(best > 0.0)) { singular = true; break; }
      if (piv != k) {
        for (int j = 0; j < P; ++j) { const double s = hm[k][j]; hm[k][j] = hm[piv][j]; hm[piv][j] = s; }
        double s = r1[k]; r1[k] = r1[piv]; r1[piv] = s;
        s = r2[k]; r2[k] = r2[piv]; r2[piv] = s;
      }
      const double rp = 1.0 / hm[k][k];
      for (int i = k + 1; i < P; ++i) {
        const double f = hm[i][k] * rp;
        if (f != 0.0) {
          for (int j = k + 1; j < P; ++j) hm[i][j] -= f * hm[k][j];
          r1[i] -= f * r1[k]; r2[i] -= f * r2[k];
        }
      }
    }
    if (!singular) {
      for (int k = P - 1; k >= 0; --k) {
        double s1 = r1[k], s2 = r2[k];
        for (int j = k + 1; j < P; ++j) { s1 -= hm[k][j] * y1[j]; s2 -= hm[k][j] * y2[j]; }
        y1[k] = s1 / hm[k][k]; y2[k] = s2 / hm[k][k];
      }
    } else {
      for (int k = 0; k < P; ++k) { y1[k] = 0.0; y2[k] = 0.0; }   // Mat::inv() of a singular matrix is all zeros
    }
  }
  double lam_n = in2, lam_d = corr;
  for (int k = 0; k < P; ++k) { lam_n -= ip[k] * y1[k]; lam_d -= tp[k] * y1[k]; }
  if (lam_d <= 0.0) {               // "The algorithm stopped before its convergence."
    st->rho = -1.0; st->status = kStatusNoConv; st->cont = 0; return;
  }
  const double lam = lam_n / lam_d;
  float dp[8];
  for (int k = 0; k < P; ++k) dp[k] = (float)(lam * y2[k] - y1[k]);   // deltaP is CV_32F
  float* m = st->m;
  if (MOTION == kTranslation) {
    m[2] += dp[0]; m[5] += dp[1];
  } else if (MOTION == kAffine) {
    m[0] += dp[0]; m[3] += dp[1]; m[1] += dp[2]; m[4] += dp[3]; m[2] += dp[4]; m[5] += dp[5];
  } else if (MOTION == kHomography) {
    m[0] += dp[0]; m[3] += dp[1]; m[6] += dp[2]; m[1] += dp[3]; m[4] += dp[4]; m[7] += dp[5];
    m[2] += dp[6]; m[5] += dp[7];
  } else {
    const double th = (double)dp[0] + asin((double)m[3]);
    m[2] += dp[1]; m[5] += dp[2];
    m[0] = m[4] = (float)cos(th);
    m[3] = (float)sin(th);
    m[1] = -m[3];
  }
  // for (i = 1; i <= maxIter && fabs(rho - last_rho) >= eps; ++i)
  st->cont = (st->iters < st->max_iter) && (fabs(st->rho - st->last_rho) >= st->eps) ? 1 : 0;
}

template <bool B> struct BoolTag { static constexpr bool value = B; };

// ---- the iteration kernel ----------------------------------------------------------------------------
template <int MOTION>
__global__ void __launch_bounds__(kEccThreads, 2) ecc_iter_kernel(const EccIterParams p) {
  using L = Layout<MOTION>;
  using Md = Model<MOTION>;
  constexpr int NV = L::NV, G = L::G;
  __shared__ float s_m[9];
  __shared__ int s_interior;
  __shared__ float s_red[kEccThreads / 32][NV];
  __shared__ int s_last;
  __shared__ double s_tot[NV];
  __shared__ double s_h[8][8];
  __shared__ double s_vec[24];

  EccState* st = p.st;
  // a frame whose loop already stopped (only reachable in the host-driven fallback loop)
  if (st->cont == 0) return;

  const int tid = threadIdx.x;
  const int strip = blockIdx.x % p.n_strips, band = blockIdx.x / p.n_strips;
  const int x0 = strip * kEccStripW;
  const int y0 = band * p.rows_per_tile;
  const int y1 = min(y0 + p.rows_per_tile, p.height);
  if (tid < 9) s_m[tid] = st->m[tid];
  __syncthreads();
  if (tid == 0)
    s_interior = tile_interior<Md::persp>(s_m, x0, min(x0 + kEccStripW, p.width) - 1, y0, y1 - 1, p.width, p.height) ? 1 : 0;
  __syncthreads();
  const bool interior = s_interior != 0;

  const int x = x0 + (tid & (kEccStripW - 1));
  const int part = tid / kEccStripW;
  const int half = (y1 - y0 + kEccRowParts - 1) / kEccRowParts;
  const int ya = y0 + part * half;
  const int yb = min(ya + half, y1);
  const float xf = (float)x;

  Accum<MOTION> acc;
  acc.clear();

  if (x < p.width && ya < yb) {
    Coord<Md::persp> co;
    co.init(s_m, x);
    // f32 Jacobian constants (OpenCV evaluates the Jacobian on f32 grids with the f32 matrix)
    float jc0 = 0.f, jc1 = 0.f, jc2 = 0.f, h3 = 0.f, h4 = 0.f, h5 = 0.f, ec = 0.f, es = 0.f;
    if (MOTION == kHomography) {
      jc0 = fmaf(xf, s_m[0], s_m[2]);     // X h0 + h6
      jc1 = fmaf(xf, s_m[3], s_m[5]);     // X h1 + h7
      jc2 = fmaf(xf, s_m[6], 1.0f);       // X h2 + 1
      h3 = s_m[1]; h4 = s_m[4]; h5 = s_m[7];
    } else if (MOTION == kEuclidean) {
      ec = s_m[0]; es = s_m[3];
    }
    const float* trow = p.tmpl + (size_t)ya * p.pitch + x;

    auto body = [&](int y, const Sample& s, float t_, float mk, auto interior_tag) {
      constexpr bool kInterior = decltype(interior_tag)::value;
      const float yf = (float)y;
      float g[G];
      if (MOTION == kTranslation) {
        g[0] = 0.5f * s.gx2; g[1] = 0.5f * s.gy2;
      } else if (MOTION == kEuclidean) {
        const float gx = 0.5f * s.gx2, gy = 0.5f * s.gy2;
        const float hx = -(xf * es) - yf * ec;
        const float hy = xf * ec - yf * es;
        g[0] = fmaf(gx, hx, gy * hy); g[1] = gx; g[2] = gy;
      } else if (MOTION == kAffine) {
        g[0] = 0.5f * s.gx2; g[1] = 0.5f * s.gy2;
      } else {
        const float den = fmaf(yf, h5, jc2);
        const float rden = __frcp_rn(den);
        const float hx = -fmaf(yf, h3, jc0) * rden;
        const float hy = -fmaf(yf, h4, jc1) * rden;
        const float hr = 0.5f * rden;
        g[0] = s.gx2 * hr; g[1] = s.gy2 * hr;
        g[2] = fmaf(hx, g[0], hy * g[1]);
      }
      acc.template add<kInterior>(g, s.w, t_, mk, yf);
    };

    if (interior) {
      // software pipeline: fetch the 12 taps + template value of row y+1 while row y is reduced
      int xq, yq;
      co.at(ya, xq, yq);
      Taps tp = load_taps(p.img, p.pitch, xq >> kInterBits, yq >> kInterBits);
      float tv = __ldg(trow);
      float ax = (float)(xq & (kInterTab - 1)) * (1.f / kInterTab);
      float ay = (float)(yq & (kInterTab - 1)) * (1.f / kInterTab);
      for (int y = ya; y < yb; ++y) {
        Taps tn = tp; float tvn = tv, axn = ax, ayn = ay;
        if (y + 1 < yb) {
          int xq2, yq2;
          co.at(y + 1, xq2, yq2);
          tn = load_taps(p.img, p.pitch, xq2 >> kInterBits, yq2 >> kInterBits);
          tvn = __ldg(trow + (size_t)(y + 1 - ya) * p.pitch);
          axn = (float)(xq2 & (kInterTab - 1)) * (1.f / kInterTab);
          ayn = (float)(yq2 & (kInterTab - 1)) * (1.f / kInterTab);
        }
        const Sample s = interp(tp, ax, ay);
        body(y, s, tv, 1.f, BoolTag<true>{});
        tp = tn; tv = tvn; ax = axn; ay = ayn;
      }
    } else {
      for (int y = ya; y < yb; ++y) {
        int xq, yq, xn, yn;
        const bool ok = co.at_with_nearest(y, xq, yq, xn, yn);
        Sample s; s.w = 0.f; s.gx2 = 0.f; s.gy2 = 0.f;
        float mk = 0.f;
        if (ok) {
          const int sx = xq >> kInterBits, sy = yq >> kInterBits;
          if (sx >= -1 && sx < p.width && sy >= -1 && sy < p.height) {
            const float ax = (float)(xq & (kInterTab - 1)) * (1.f / kInterTab);
            const float ay = (float)(yq & (kInterTab - 1)) * (1.f / kInterTab);
            s = sample_general(p.img, p.pitch, p.width, p.height, sx, sy, ax, ay);
          }
          // OpenCV warps an all-ones mask with INTER_NEAREST: 1 where the rounded position is inside
          mk = ((unsigned)xn < (unsigned)p.width && (unsigned)yn < (unsigned)p.height) ? 1.f : 0.f;
        }
        const float tv = __ldg(trow + (size_t)(y - ya) * p.pitch);
        body(y, s, tv, mk, BoolTag<false>{});
      }
    }
  }

  // ---- block reduction: registers -> warp shuffle -> smem -> f64 partial of this tile -------------
  float v[NV];
  acc.emit(xf, v);
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float s = warp_sum(v[i]);
    if (lane == 0) s_red[wid][i] = s;
  }
  __syncthreads();
  double* part_out = p.partials + (size_t)blockIdx.x * NV;
  if (tid < NV) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kEccThreads / 32; ++w) s += (double)s_red[w][tid];
    part_out[tid] = s;
  }

  // ---- last block: deterministic cross-tile sum + solve --------------------------------------------
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int done = atomicAdd(&st->tile_counter, 1u);
    s_last = (done == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int n_tiles = gridDim.x;
  for (int i = wid; i < NV; i += kEccThreads / 32) {
    double s = 0.0;
    for (int t = lane; t < n_tiles; t += 32) s += __ldcg(p.partials + (size_t)t * NV + i);
    s = warp_sum(s);
    if (lane == 0) s_tot[i] = s;
  }
  __syncthreads();
  if (p.totals_out && tid < NV) p.totals_out[tid] = s_tot[tid];
  if (tid == 0) {
    st->tile_counter = 0;
    ecc_solve_and_update<MOTION>(s_tot, st, s_h, s_vec, s_vec + 8);
    if (st->cont == 0) compute_inverse(st, Md::persp);
    if (p.use_handle) cudaGraphSetConditional(p.handle, (unsigned)st->cont);
  }
}

// State initialisation at the head of each frame's loop (identity warp, rho = -1, last_rho = -eps).
__global__ void ecc_init_kernel(EccState* st, int persp, int max_iter, double eps,
                                cudaGraphConditionalHandle handle, int use_handle) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int i = 0; i < 9; ++i) st->m[i] = (i == 0 || i == 4 || i == 8) ? 1.f : 0.f;
  st->rho = -1.0;
  st->last_rho = -eps;
  st->eps = eps;
  st->max_iter = max_iter;
  st->iters = 0;
  st->status = kStatusOk;
  st->tile_counter = 0;
  // for (i = 1; i <= maxIter && fabs(rho - last_rho) >= eps; ...) evaluated before the first iteration
  st->cont = (max_iter >= 1 && fabs(-1.0 - (-eps)) >= eps) ? 1 : 0;
  compute_inverse(st, persp != 0);
  if (use_handle) cudaGraphSetConditional(handle, (unsigned)st->cont);
}

}  // namespace stk
