// K2 — one ECC iteration, fused: inverse-map warp of the blurred reference image I and of its two
// central-difference gradient planes (derived on the fly from I, never stored), nearest-neighbour mask,
// per-pixel Jacobian for the four motion models, and every reduction the update needs, accumulated in
// registers -> warp shuffles -> block -> per-tile partials (f64) -> the last block to finish sums the
// partials in a fixed order, one warp solves the PxP normal equations in f64 (Gauss-Jordan, one lane per
// row), updates the f32 warp matrix, runs the convergence test and sets the CUDA-graph WHILE condition.
// No per-pixel plane is written and no iteration round-trips to the host.
//
// Replaces the body of OpenCV's findTransformECC loop (modules/video/src/ecc.cpp), which the reference
// reaches through opencv::video::find_transform_ecc at /root/reference/src/lib.rs:769-777
// (template = frame i, input = frame 0, identity init, no mask).  Restated on the CPU in
// oracle/restate.py (ecc_sums / ecc_epilogue), which is pinned against cv2.findTransformECC.
//
// Work decomposition ("column owner"): the frame is cut into 128-column strips of 16-row chunks and the
// chunk list is dealt evenly to one persistent 256-thread block per resident slot (2 per SM); within a run
// of chunks of one strip every thread keeps ONE column x.  Because X is constant per thread, the Kronecker
// structure of the affine / homography Jacobians ( J = g (x) [X, Y, 1] ) lets a pixel accumulate only
// g_i*g_j*{1,Y,Y^2} and g_i*z*{1,Y}; the X factors are folded in once per run.
//
// Data movement: the tile is walked in chunks of 16 rows.  The bounding boxes of all chunks' sample
// positions are computed up front (one thread per chunk); an elected lane then keeps four stages of TMA
// tiled copies in flight (cp.async.bulk.tensor, completion on "full" mbarriers, stage release on "empty"
// mbarriers with one arrival per warp — no block-wide barrier in the loop): a 144x32 box of I around the
// warped chunk and the 128x16 block of T.  HBM/L2 latency is off the critical path and the 12 taps per
// pixel are LDS at constant offsets.  Pixels whose taps touch the image border (a thin rim), and chunks
// whose samples would not fit a box (extreme warps), take the general path: direct global loads with
// every border rule applied per tap.
// Algorithmic traffic: 4N (T) + 4N (I) bytes per iteration.
#pragma once
#include <cuda.h>

#include <climits>

#include "common.cuh"

namespace stk {

constexpr int kEccThreads = 256;
constexpr int kEccStripW = 128;     // columns per tile
constexpr int kEccRowParts = kEccThreads / kEccStripW;   // 2: rows of a chunk are split over two thread halves
constexpr int kChunkH = 16;         // rows per chunk
constexpr int kChunkRowsPerThread = kChunkH / kEccRowParts;
constexpr int kBoxW = 144;          // I box: 128 columns + drift/halo margin
constexpr int kBoxH = 32;           // I box: 16 rows + drift/halo margin
constexpr int kEccStages = 4;
constexpr int kMaxChunks = 128;      // chunks per run (box table size); longer runs are split
constexpr int kImgStageBytes = kBoxW * kBoxH * 4;          // 18432
constexpr int kTmplStageBytes = kEccStripW * kChunkH * 4;  // 8192
constexpr int kStageBytes = kImgStageBytes + kTmplStageBytes;
constexpr int kEccDynSmem = kEccStages * kStageBytes;

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
  // ecc_match_scaling_down: factors that take the matrix estimated on the downscaled planes to full
  // resolution once the loop has ended (0 = no rescale).  Set once per context, never by the init kernel.
  float rescale_x, rescale_y;
};

struct alignas(64) EccIterParams {
  CUtensorMap tm_img;       // I : f32 [H][W], box kBoxW x (chunk height + 16) of the kernel configuration, zero fill outside
  CUtensorMap tm_tmpl;      // T : f32 [H][W], box kEccStripW x chunk height
  const float* img;         // same planes, for the general path
  const float* tmpl;
  int pitch;                // floats, both planes
  int width, height;        // template size == image size on this path
  int n_strips;             // 128-column strips
  int chunks_per_strip;     // ceil(height / chunk height of the kernel configuration)
  double* partials;         // [NV][tiles_pad]: value-major so the cross-block sum reads coalesced
  int tiles_pad;            // gridDim.x rounded up to 32
  EccState* st;
  cudaGraphConditionalHandle handle;
  int use_handle;
  unsigned frac_magic;      // 0x4B400000 (1.5 * 2^23): kept in a register by the lean pixel body (ecc_iter_v2.cuh::lean_pixel)
  int rim_weight;           // cost of a chunk of the first/last strip in 1/8 of an interior chunk (8 = no bias)
  double* totals_out;       // optional: the NV reduced sums of this iteration (test hook), else null
  unsigned long long* timing_out;   // optional: %globaltimer stamps per tile [n_tiles][4] + tail [4] (profiling hook)
};

// ---- per-model constants -------------------------------------------------------------------------
template <int MOTION> struct Model;
template <> struct Model<kTranslation> { static constexpr int P = 2, G = 2; static constexpr bool kron = false, persp = false; };
template <> struct Model<kEuclidean>   { static constexpr int P = 3, G = 3; static constexpr bool kron = false, persp = false; };
template <> struct Model<kAffine>      { static constexpr int P = 6, G = 2; static constexpr bool kron = true,  persp = false; };
template <> struct Model<kHomography>  { static constexpr int P = 8, G = 3; static constexpr bool kron = true,  persp = true;  };

// Layout of the NV reduced sums ("totals"):
//   [0..5]  n Sw Sww St Stt Swt                      (masked scalar sums)
//   [kH..]  for each product g_i g_j (i <= j): its moments {1, X, Y, XX, XY, YY}  (kron)  |  {1}
//   [kZ..]  for z in {w, m, m*t}: for each g_i: moments {1, X, Y}  (kron)  |  {1}
// where g = (gx, gy) [Translation, Affine], (gx*hatX + gy*hatY, gx, gy) [Euclidean], (a, b, t) [Homography].
template <int MOTION> struct Layout {
  using M = Model<MOTION>;
  static constexpr int G = M::G;
  static constexpr int NP = G * (G + 1) / 2;
  static constexpr int QM = M::kron ? 6 : 1;
  static constexpr int ZM = M::kron ? 3 : 1;
  static constexpr int kScal = 6;
  static constexpr int kH = kScal;
  static constexpr int kZ = kH + NP * QM;
  static constexpr int NV = kZ + 3 * G * ZM;
};

// ---- TMA / mbarrier primitives (sm_90+ PTX) ----------------------------------------------------------
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded spin: a copy that never lands (bad descriptor, lost arrival) traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  for (unsigned spins = 0; !mbar_try_wait(b, parity); ++spins) {
    if (spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// ---- sampling ------------------------------------------------------------------------------------
struct Sample { float w, gx2, gy2; };   // bilinear I, 2*bilinear(GX), 2*bilinear(GY)

__device__ __forceinline__ float lerp(float a, float b, float t) { return fmaf(t, b - a, a); }
// 1 MUFU; |rel err| <= 2^-23 on the normal range (the denominators here are ~1)
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// packed FP32 pairs (Blackwell FFMA2 / FADD2 / FMUL2, PTX *.f32x2): one issue slot for two lanes of work.  The
// SASS forms take a scalar register as a broadcast operand (R.F32) and negate packed sources, so
// f2(s) operands and a - b cost nothing extra.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }   // a - b
__device__ __forceinline__ float2 lerp2(float2 a, float2 b, float2 t) { return fma2(t, sub2(b, a), a); }

// fast path: 12 taps from the shared-memory box (row stride kBoxW), p = &box[sy][sx]
__device__ __forceinline__ Sample sample_box(const float* p, float ax, float ay) {
  const float m0 = p[-kBoxW], m1 = p[-kBoxW + 1];
  const float a_1 = p[-1], a0 = p[0], a1 = p[1], a2 = p[2];
  const float b_1 = p[kBoxW - 1], b0 = p[kBoxW], b1 = p[kBoxW + 1], b2 = p[kBoxW + 2];
  const float c0 = p[2 * kBoxW], c1 = p[2 * kBoxW + 1];
  Sample s;
  s.w = lerp(lerp(a0, a1, ax), lerp(b0, b1, ax), ay);
  // GX taps (2*GX = I[c+1] - I[c-1]) at (sy,sx) (sy,sx+1) (sy+1,sx) (sy+1,sx+1)
  s.gx2 = lerp(lerp(a1 - a_1, a2 - a0, ax), lerp(b1 - b_1, b2 - b0, ax), ay);
  // GY taps (2*GY = I[r+1] - I[r-1])
  s.gy2 = lerp(lerp(b0 - m0, b1 - m1, ax), lerp(c0 - a0, c1 - a1, ax), ay);
  return s;
}

// Same values as sample_box with the arithmetic paired up: the two gradient planes share every step
// (first/second horizontal tap, top/bottom row), so (gx2, gy2) comes out as one packed value in 10 packed
// instructions instead of 20 scalar ones.  Operation order per lane is the one of sample_box.
__device__ __forceinline__ void sample_box_packed(const float* p, float ax, float ay, float& w, float2& gxy2) {
  const float m0 = p[-kBoxW], m1 = p[-kBoxW + 1];
  const float a_1 = p[-1], a0 = p[0], a1 = p[1], a2 = p[2];
  const float b_1 = p[kBoxW - 1], b0 = p[kBoxW], b1 = p[kBoxW + 1], b2 = p[kBoxW + 2];
  const float c0 = p[2 * kBoxW], c1 = p[2 * kBoxW + 1];
  w = lerp(lerp(a0, a1, ax), lerp(b0, b1, ax), ay);
  // top row of the 2x2 gradient taps: (GX, GY) at (sy, sx) and (sy, sx+1); bottom row: at sy+1
  const float2 t0 = sub2(f2(a1, b0), f2(a_1, m0)), t1 = sub2(f2(a2, b1), f2(a0, m1));
  const float2 u0 = sub2(f2(b1, c0), f2(b_1, a0)), u1 = sub2(f2(b2, c1), f2(b0, a1));
  const float2 ax2 = f2(ax);
  gxy2 = lerp2(lerp2(t0, t1, ax2), lerp2(u0, u1, ax2), f2(ay));
}

// careful path: same 12 taps from the box, with the border rules of the gradient planes applied as 0/1
// factors.  TMA zero-filled the part of the box outside the image, which already gives I = 0 there and
// GX = GY = 0 for taps whose row (GX) or column (GY) is outside; what is left is GX = 0 on columns <= 0 and
// >= W-1 and GY = 0 on rows <= 0 and >= H-1 (filter2D's BORDER_REFLECT_101 + the constant border).
__device__ __forceinline__ Sample sample_box_rules(const float* p, float ax, float ay, int sx, int sy, int w, int h) {
  const float zx0 = ((unsigned)(sx - 1) <= (unsigned)(w - 3)) ? 1.f : 0.f;       // column sx     in [1, W-2]
  const float zx1 = ((unsigned)sx <= (unsigned)(w - 3)) ? 1.f : 0.f;             // column sx + 1 in [1, W-2]
  const float zy0 = ((unsigned)(sy - 1) <= (unsigned)(h - 3)) ? 1.f : 0.f;
  const float zy1 = ((unsigned)sy <= (unsigned)(h - 3)) ? 1.f : 0.f;
  const float m0 = p[-kBoxW], m1 = p[-kBoxW + 1];
  const float a_1 = p[-1], a0 = p[0], a1 = p[1], a2 = p[2];
  const float b_1 = p[kBoxW - 1], b0 = p[kBoxW], b1 = p[kBoxW + 1], b2 = p[kBoxW + 2];
  const float c0 = p[2 * kBoxW], c1 = p[2 * kBoxW + 1];
  Sample s;
  s.w = lerp(lerp(a0, a1, ax), lerp(b0, b1, ax), ay);
  s.gx2 = lerp(lerp(zx0 * (a1 - a_1), zx1 * (a2 - a0), ax), lerp(zx0 * (b1 - b_1), zx1 * (b2 - b0), ax), ay);
  s.gy2 = lerp(lerp(zy0 * (b0 - m0), zy0 * (b1 - m1), ax), lerp(zy1 * (c0 - a0), zy1 * (c1 - a1), ax), ay);
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
// Exact evaluation (per-thread fixed column x): the quantised source position (Xq, Yq in 1/32 px) as
// OpenCV's WARP_INVERSE_MAP warps compute it — f64 projective divide + round-half-even for
// warpPerspective, 10-bit fixed point for warpAffine — and the separately rounded INTER_NEAREST position
// the mask uses.
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

// Fast projective coordinates (interior chunks only): the DISPLACEMENT (u - x, v - y) is evaluated in
// f32 from per-thread constants prepared in f64,
//     u - x = (alpha_x + beta_x y) / w ,  v - y = (gamma_x + delta_x y - m21 y^2) / w ,  w = wc_x + m21 y
// and quantised with round-half-even: 32 u = 32 x + rint(32 (u - x)) because x is an integer.  For the
// near-identity warps ECC works on, the displacement is a few pixels, so its f32 error (~1e-6 px) moves
// the 1/32-px quantisation for ~1e-4 of the pixels by one step — far inside the 0.05 px parity bar — at
// a third of the cost of the f64 divide.  EccIterParams::exact_coords switches it off for validation.
struct FastPersp {
  float alpha, beta, gamma, delta, wc, m21;
  __device__ __forceinline__ void init(const float* m, int x) {
    const double xd = (double)x;
    const double m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21d = m[7], m22 = m[8];
    alpha = (float)(xd * (m00 - m22 - m20 * xd) + m02);
    beta = (float)(m01 - m21d * xd);
    gamma = (float)(m10 * xd + m12);
    delta = (float)(m11 - m22 - m20 * xd);
    wc = (float)(m20 * xd + m22);
    m21 = (float)m21d;
  }
  // qx, qy: rint(32 du), rint(32 dv); du, dv, rw also returned for the Jacobian
  __device__ __forceinline__ void at(float yf, int& qx, int& qy, float& du, float& dv, float& rw) const {
    const float w = fmaf(m21, yf, wc);
    rw = rcp_approx(w);
    du = fmaf(beta, yf, alpha) * rw;
    dv = fmaf(fmaf(-m21, yf, delta), yf, gamma) * rw;
    // 1.5 * 2^23 magic: the low mantissa bits of (32 d + magic) are rint(32 d), |32 d| < 2^22
    qx = __float_as_int(fmaf(du, 32.0f, 12582912.0f)) - 0x4B400000;
    qy = __float_as_int(fmaf(dv, 32.0f, 12582912.0f)) - 0x4B400000;
  }
  // rint(du), rint(dv): the INTER_NEAREST position relative to (x, y), for the mask
  static __device__ __forceinline__ void nearest(float du, float dv, int& nx, int& ny) {
    nx = __float_as_int(du + 12582912.0f) - 0x4B400000;
    ny = __float_as_int(dv + 12582912.0f) - 0x4B400000;
  }
};

// ---- chunk geometry -------------------------------------------------------------------------------
// Bounding box of the sample positions of the pixel rectangle [x0,x1] x [y0,y1] (inclusive).  The image
// of a rectangle under an affine map, or under a projective map with w > 0 on its four corners, is the
// convex hull of the corner images.  ok == false: no usable bound (w <= 0 somewhere / NaN).
template <bool PERSP>
__device__ bool chunk_bounds(const float* m, int x0, int x1, int y0, int y1, double& umin, double& umax,
                             double& vmin, double& vmax) {
  bool ok = true;
  umin = vmin = 1e300; umax = vmax = -1e300;
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
    if (!(u == u) || !(v == v)) ok = false;
    umin = fmin(umin, u); umax = fmax(umax, u); vmin = fmin(vmin, v); vmax = fmax(vmax, v);
  }
  return ok;
}

// The same bounding box in f32 (second-generation kernel): the table is on every block's start-up path, and the f64
// divides made it ~1 us of the ~3 us a block needs before its first chunk lands.  |error| <~ 1e-3 px at 4K-class
// coordinates — absorbed by the 2-px box margin and the 1/16-px slack of the interior test (callers add it).
template <bool PERSP>
__device__ bool chunk_bounds_f32(const float* m, int x0, int x1, int y0, int y1, float& umin, float& umax,
                                 float& vmin, float& vmax) {
  bool ok = true;
  umin = vmin = 3.0e38f; umax = vmax = -3.0e38f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float x = (k & 1) ? (float)x1 : (float)x0;
    const float y = (k & 2) ? (float)y1 : (float)y0;
    float u = fmaf(m[0], x, fmaf(m[1], y, m[2]));
    float v = fmaf(m[3], x, fmaf(m[4], y, m[5]));
    if (PERSP) {
      const float ww = fmaf(m[6], x, fmaf(m[7], y, m[8]));
      if (!(ww > 1e-6f)) { ok = false; continue; }
      const float rw = 1.0f / ww;
      u *= rw; v *= rw;
    }
    if (!(u == u) || !(v == v)) ok = false;
    umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
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

  // g: Jacobian generators, w_: warped image, t_: template, mk: mask (0/1), yf: row as float.
  // INTERIOR: mask is known to be 1 (and the pixel count is added by the caller).
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
    if (!INTERIOR) n += mk;
    sw += wm; sww = fmaf(wm, w_, sww);
    st += tm; stt = fmaf(tm, t_, stt);
    swt = fmaf(wm, t_, swt);
  }

  // fold the per-thread column coordinate X into the moments and emit the NV values of this thread.
  // TWICE_NEG (homography with FastPersp coordinates): the generators were accumulated as (2a, 2b, -2t) —
  // the pixel loop skips the 0.5 and the negation — so products carry a factor 4 (and the sign of the t
  // factors), the projections a factor 2; both are exact powers of two, undone here once per run.
  template <bool TWICE_NEG = false>
  __device__ __forceinline__ void emit(float xf, float (&v)[L::NV]) const {
    v[0] = n; v[1] = sw; v[2] = sww; v[3] = st; v[4] = stt; v[5] = swt;
    const float xx = xf * xf;
    int idx = 0;
#pragma unroll
    for (int gi = 0; gi < G; ++gi)
#pragma unroll
      for (int gj = gi; gj < G; ++gj) {
        const float f = !TWICE_NEG ? 1.f : (((gi == G - 1) != (gj == G - 1)) ? -0.25f : 0.25f);
        const float p0 = TWICE_NEG ? p[idx][0] * f : p[idx][0];
        if (kron) {
          const float p1 = TWICE_NEG ? p[idx][1] * f : p[idx][1];
          const float p2 = TWICE_NEG ? p[idx][2] * f : p[idx][2];
          float* o = &v[L::kH + idx * 6];
          o[0] = p0; o[1] = xf * p0; o[2] = p1;
          o[3] = xx * p0; o[4] = xf * p1; o[5] = p2;
        } else {
          v[L::kH + idx] = p0;
        }
        ++idx;
      }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const float f = !TWICE_NEG ? 1.f : (i == G - 1 ? -0.5f : 0.5f);
        const float z0 = TWICE_NEG ? z[a][i][0] * f : z[a][i][0];
        if (kron) {
          const float z1 = TWICE_NEG ? z[a][i][1] * f : z[a][i][1];
          float* o = &v[L::kZ + (a * G + i) * 3];
          o[0] = z0; o[1] = xf * z0; o[2] = z1;
        } else {
          v[L::kZ + a * G + i] = z0;
        }
      }
  }
};

// Homography accumulator with the sums stored as register PAIRS so the interior path updates two of them
// per instruction.  Same sums, same per-sum operation order as Accum<kHomography>; generators are
// (g0, g1, g2) = (2a, 2b, -2t) (see Accum::emit<TWICE_NEG>).
//   h[k]  = (p00, p01), (p02, p12), (p11, p22) for k = 0..2, each with moments {1, Y, Y^2}: the first two pairs are
//           (g0, g1) times a broadcast scalar, the third two scalar squares — no register moves to build them
//   za/zm/zt[m] = (g0 z, g1 z) for z = w / 1 / t, moments {1, Y} ; zc[m] = (g2 w, g2 t) ; zm2[m] = g2
//   s1 = (Sw, St), s2 = (Sww, Stt)
struct AccumH2 {
  using L = Layout<kHomography>;
  static constexpr int G = 3;
  float n, swt;
  float2 s1, s2;
  float2 h[3][3];
  float2 za[2], zm[2], zt[2], zc[2];
  float zm2[2];

  __device__ __forceinline__ void clear() {
    const float2 o = f2(0.f);
    n = swt = 0.f; s1 = s2 = o;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int m = 0; m < 3; ++m) h[k][m] = o;
#pragma unroll
    for (int m = 0; m < 2; ++m) { za[m] = zm[m] = zt[m] = zc[m] = o; zm2[m] = 0.f; }
  }

  // interior pixel (mask 1; the caller counts it): g01 = (g0, g1)
  __device__ __forceinline__ void add_packed(float2 g01, float g2, float w_, float t_, float yf) {
    const float2 y2 = f2(yf), yy2 = f2(yf * yf);
    const float2 q0 = mul2(g01, f2(g01.x));                       // g0 g0, g0 g1
    const float2 q1 = mul2(g01, f2(g2));                          // g0 g2, g1 g2
    const float2 q2 = f2(g01.y * g01.y, g2 * g2);                 // g1 g1, g2 g2
    h[0][0] = add2(h[0][0], q0); h[0][1] = fma2(q0, y2, h[0][1]); h[0][2] = fma2(q0, yy2, h[0][2]);
    h[1][0] = add2(h[1][0], q1); h[1][1] = fma2(q1, y2, h[1][1]); h[1][2] = fma2(q1, yy2, h[1][2]);
    h[2][0] = add2(h[2][0], q2); h[2][1] = fma2(q2, y2, h[2][1]); h[2][2] = fma2(q2, yy2, h[2][2]);
    const float2 wt = f2(w_, t_);
    const float2 gw = mul2(g01, f2(w_)), gt = mul2(g01, f2(t_)), gc = mul2(f2(g2), wt);
    za[0] = add2(za[0], gw);  za[1] = fma2(gw, y2, za[1]);
    zm[0] = add2(zm[0], g01); zm[1] = fma2(g01, y2, zm[1]);
    zt[0] = add2(zt[0], gt);  zt[1] = fma2(gt, y2, zt[1]);
    zc[0] = add2(zc[0], gc);  zc[1] = fma2(gc, y2, zc[1]);
    zm2[0] += g2; zm2[1] = fmaf(g2, yf, zm2[1]);
    s1 = add2(s1, wt);
    s2 = fma2(wt, wt, s2);
    swt = fmaf(w_, t_, swt);
  }

  // general pixel (rim / careful / slow paths): scalar arithmetic on the same registers
  template <bool INTERIOR>
  __device__ __forceinline__ void add(const float (&g)[3], float w_, float t_, float mk, float yf) {
    const float yy = yf * yf;
    auto up = [&](float& a0, float& a1, float& a2, float pr) { a0 += pr; a1 = fmaf(pr, yf, a1); a2 = fmaf(pr, yy, a2); };
    up(h[0][0].x, h[0][1].x, h[0][2].x, g[0] * g[0]);
    up(h[0][0].y, h[0][1].y, h[0][2].y, g[0] * g[1]);
    up(h[1][0].x, h[1][1].x, h[1][2].x, g[0] * g[2]);
    up(h[1][0].y, h[1][1].y, h[1][2].y, g[1] * g[2]);
    up(h[2][0].x, h[2][1].x, h[2][2].x, g[1] * g[1]);
    up(h[2][0].y, h[2][1].y, h[2][2].y, g[2] * g[2]);
    const float wm = INTERIOR ? w_ : w_ * mk;
    const float tm = INTERIOR ? t_ : t_ * mk;
    auto uz = [&](float& a0, float& a1, float v) { a0 += v; a1 = fmaf(v, yf, a1); };
    uz(za[0].x, za[1].x, g[0] * w_); uz(za[0].y, za[1].y, g[1] * w_); uz(zc[0].x, zc[1].x, g[2] * w_);
    uz(zm[0].x, zm[1].x, INTERIOR ? g[0] : g[0] * mk); uz(zm[0].y, zm[1].y, INTERIOR ? g[1] : g[1] * mk);
    uz(zm2[0], zm2[1], INTERIOR ? g[2] : g[2] * mk);
    uz(zt[0].x, zt[1].x, g[0] * tm); uz(zt[0].y, zt[1].y, g[1] * tm); uz(zc[0].y, zc[1].y, g[2] * tm);
    if (!INTERIOR) n += mk;
    s1.x += wm; s2.x = fmaf(wm, w_, s2.x);
    s1.y += tm; s2.y = fmaf(tm, t_, s2.y);
    swt = fmaf(wm, t_, swt);
  }

  template <bool TWICE_NEG>
  __device__ __forceinline__ void emit(float xf, float (&v)[L::NV]) const {
    static_assert(TWICE_NEG, "AccumH2 holds (2a, 2b, -2t) sums");
    v[0] = n; v[1] = s1.x; v[2] = s2.x; v[3] = s1.y; v[4] = s2.y; v[5] = swt;
    const float xx = xf * xf;
    // product order of Layout<>: (0,0) (0,1) (0,2) (1,1) (1,2) (2,2); the sign flips where exactly one factor is g2
    const float pf[6] = {0.25f, 0.25f, -0.25f, 0.25f, -0.25f, 0.25f};
    // where each product lives: pair index, .y?
    const int pk[6] = {0, 0, 1, 2, 1, 2};
    const bool py[6] = {false, true, false, false, true, true};
#pragma unroll
    for (int idx = 0; idx < 6; ++idx) {
      const float2 m0 = h[pk[idx]][0], m1 = h[pk[idx]][1], m2 = h[pk[idx]][2];
      const float p0 = (py[idx] ? m0.y : m0.x) * pf[idx];
      const float p1 = (py[idx] ? m1.y : m1.x) * pf[idx];
      const float p2 = (py[idx] ? m2.y : m2.x) * pf[idx];
      float* o = &v[L::kH + idx * 6];
      o[0] = p0; o[1] = xf * p0; o[2] = p1; o[3] = xx * p0; o[4] = xf * p1; o[5] = p2;
    }
    // projections: z in {w, m, m t} x g_i, moments {1, X, Y}
    const float z0[3][3] = {{za[0].x, za[0].y, zc[0].x}, {zm[0].x, zm[0].y, zm2[0]}, {zt[0].x, zt[0].y, zc[0].y}};
    const float z1[3][3] = {{za[1].x, za[1].y, zc[1].x}, {zm[1].x, zm[1].y, zm2[1]}, {zt[1].x, zt[1].y, zc[1].y}};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float f = i == 2 ? -0.5f : 0.5f;
        const float v0 = z0[a][i] * f, v1 = z1[a][i] * f;
        float* o = &v[L::kZ + (a * 3 + i) * 3];
        o[0] = v0; o[1] = xf * v0; o[2] = v1;
      }
  }
};

// Third form of the Homography accumulator ("premultiplied"): the same 41 sums with the row coordinate folded into the
// GENERATORS instead of into the products.  With gy_i = g_i * Y (3 multiplies per pixel),
//     sum g_i g_j        = fma(g_i,  g_j,  .)      sum g_i z     = fma(g_i,  z, .)
//     sum g_i g_j Y      = fma(gy_i, g_j,  .)      sum g_i z Y   = fma(gy_i, z, .)
//     sum g_i g_j Y^2    = fma(gy_i, gy_j, .)      sum g_i, sum g_i Y : plain adds
// every sum is ONE instruction and no product g_i g_j / g_i z is ever formed on its own: 44 FP32-pipe cycles per pixel
// instead of 54 (AccumH2: 15 products + 39 accumulations), 27 issue slots instead of 30.  Each update rounds once
// (fused) where AccumH2 rounded the product first — at least as accurate.
//   hA[m] = (p00, p11), hB[m] = (p02, p12): packed; hC[m] = (p01, p22): two scalar chains; m = moment {1, Y, Y^2}
//   za/zt/zm[m] = (g0 z, g1 z) for z = w / t / 1; zc[m] = (g2 w, g2 t); zm2[m] = g2; s1 = (Sw, St), s2 = (Sww, Stt)
struct AccumH3 {
  using L = Layout<kHomography>;
  static constexpr int G = 3;
  float n, swt;
  float2 s1, s2;
  float2 hA[3], hB[3], hC[3];
  float2 za[2], zm[2], zt[2], zc[2];
  float zm2[2];

  __device__ __forceinline__ void clear() {
    const float2 o = f2(0.f);
    n = swt = 0.f; s1 = s2 = o;
#pragma unroll
    for (int m = 0; m < 3; ++m) hA[m] = hB[m] = hC[m] = o;
#pragma unroll
    for (int m = 0; m < 2; ++m) { za[m] = zm[m] = zt[m] = zc[m] = o; zm2[m] = 0.f; }
  }

  // interior pixel (mask 1; the caller counts it): g01 = (g0, g1)
  __device__ __forceinline__ void add_packed(float2 g01, float g2, float w_, float t_, float yf) {
    const float2 gy01 = mul2(g01, f2(yf));
    const float gy2 = g2 * yf;
    hA[0] = fma2(g01, g01, hA[0]); hA[1] = fma2(gy01, g01, hA[1]); hA[2] = fma2(gy01, gy01, hA[2]);
    hB[0] = fma2(g01, f2(g2), hB[0]); hB[1] = fma2(gy01, f2(g2), hB[1]); hB[2] = fma2(gy01, f2(gy2), hB[2]);
    hC[0].x = fmaf(g01.x, g01.y, hC[0].x); hC[1].x = fmaf(gy01.x, g01.y, hC[1].x); hC[2].x = fmaf(gy01.x, gy01.y, hC[2].x);
    hC[0].y = fmaf(g2, g2, hC[0].y); hC[1].y = fmaf(gy2, g2, hC[1].y); hC[2].y = fmaf(gy2, gy2, hC[2].y);
    const float2 wt = f2(w_, t_);
    za[0] = fma2(g01, f2(w_), za[0]); za[1] = fma2(gy01, f2(w_), za[1]);
    zt[0] = fma2(g01, f2(t_), zt[0]); zt[1] = fma2(gy01, f2(t_), zt[1]);
    zc[0] = fma2(f2(g2), wt, zc[0]);  zc[1] = fma2(f2(gy2), wt, zc[1]);
    zm[0] = add2(zm[0], g01);         zm[1] = add2(zm[1], gy01);
    zm2[0] += g2; zm2[1] += gy2;
    s1 = add2(s1, wt);
    s2 = fma2(wt, wt, s2);
    swt = fmaf(w_, t_, swt);
  }

  // general pixel (rim / careful / slow paths): scalar arithmetic on the same registers
  template <bool INTERIOR>
  __device__ __forceinline__ void add(const float (&g)[3], float w_, float t_, float mk, float yf) {
    const float y0 = g[0] * yf, y1 = g[1] * yf, y2 = g[2] * yf;
    auto up = [&](float& a0, float& a1, float& a2, float gi, float gj, float yi, float yj) {
      a0 = fmaf(gi, gj, a0); a1 = fmaf(yi, gj, a1); a2 = fmaf(yi, yj, a2);
    };
    up(hA[0].x, hA[1].x, hA[2].x, g[0], g[0], y0, y0);
    up(hA[0].y, hA[1].y, hA[2].y, g[1], g[1], y1, y1);
    up(hB[0].x, hB[1].x, hB[2].x, g[0], g[2], y0, y2);
    up(hB[0].y, hB[1].y, hB[2].y, g[1], g[2], y1, y2);
    up(hC[0].x, hC[1].x, hC[2].x, g[0], g[1], y0, y1);
    up(hC[0].y, hC[1].y, hC[2].y, g[2], g[2], y2, y2);
    const float wm = INTERIOR ? w_ : w_ * mk;
    const float tm = INTERIOR ? t_ : t_ * mk;
    auto uz = [&](float& a0, float& a1, float gi, float yi, float z) { a0 = fmaf(gi, z, a0); a1 = fmaf(yi, z, a1); };
    uz(za[0].x, za[1].x, g[0], y0, w_); uz(za[0].y, za[1].y, g[1], y1, w_); uz(zc[0].x, zc[1].x, g[2], y2, w_);
    uz(zt[0].x, zt[1].x, g[0], y0, tm); uz(zt[0].y, zt[1].y, g[1], y1, tm); uz(zc[0].y, zc[1].y, g[2], y2, tm);
    if (INTERIOR) {
      zm[0].x += g[0]; zm[1].x += y0; zm[0].y += g[1]; zm[1].y += y1; zm2[0] += g[2]; zm2[1] += y2;
    } else {
      uz(zm[0].x, zm[1].x, g[0], y0, mk); uz(zm[0].y, zm[1].y, g[1], y1, mk); uz(zm2[0], zm2[1], g[2], y2, mk);
      n += mk;
    }
    s1.x += wm; s2.x = fmaf(wm, w_, s2.x);
    s1.y += tm; s2.y = fmaf(tm, t_, s2.y);
    swt = fmaf(wm, t_, swt);
  }

  template <bool TWICE_NEG>
  __device__ __forceinline__ void emit(float xf, float (&v)[L::NV]) const {
    static_assert(TWICE_NEG, "AccumH3 holds (2a, 2b, -2t) sums");
    v[0] = n; v[1] = s1.x; v[2] = s2.x; v[3] = s1.y; v[4] = s2.y; v[5] = swt;
    const float xx = xf * xf;
    // product order of Layout<>: (0,0) (0,1) (0,2) (1,1) (1,2) (2,2); the sign flips where exactly one factor is g2
    const float pf[6] = {0.25f, 0.25f, -0.25f, 0.25f, -0.25f, 0.25f};
#pragma unroll
    for (int idx = 0; idx < 6; ++idx) {
      float m[3];
#pragma unroll
      for (int k = 0; k < 3; ++k)
        m[k] = idx == 0 ? hA[k].x : idx == 1 ? hC[k].x : idx == 2 ? hB[k].x : idx == 3 ? hA[k].y : idx == 4 ? hB[k].y : hC[k].y;
      const float p0 = m[0] * pf[idx], p1 = m[1] * pf[idx], p2 = m[2] * pf[idx];
      float* o = &v[L::kH + idx * 6];
      o[0] = p0; o[1] = xf * p0; o[2] = p1; o[3] = xx * p0; o[4] = xf * p1; o[5] = p2;
    }
    const float z0[3][3] = {{za[0].x, za[0].y, zc[0].x}, {zm[0].x, zm[0].y, zm2[0]}, {zt[0].x, zt[0].y, zc[0].y}};
    const float z1[3][3] = {{za[1].x, za[1].y, zc[1].x}, {zm[1].x, zm[1].y, zm2[1]}, {zt[1].x, zt[1].y, zc[1].y}};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float f = i == 2 ? -0.5f : 0.5f;
        const float v0 = z0[a][i] * f, v1 = z1[a][i] * f;
        float* o = &v[L::kZ + (a * 3 + i) * 3];
        o[0] = v0; o[1] = xf * v0; o[2] = v1;
      }
  }
};

// Affine accumulator in the premultiplied packed form of AccumH3 (round 2, last session: the 2x3 models still ran the
// scalar body — 113 warp-instructions per pixel against Homography's 79, ncu on BASELINE configs[2]'s motion model).
// Generators are accumulated as (g0, g1) = (2 gx, 2 gy) — what the sampler returns — and the exact factors 1/4 (products)
// and 1/2 (projections) are undone once per run in emit().  26 sums, 16 instructions per interior pixel:
//   hA[m] = (p00, p11), hC[m] = p01 for the moments m = {1, Y, Y^2}; za/zm/zt[m] = (g0 z, g1 z) for z = w / 1 / t, m = {1, Y}
struct AccumA3 {
  using L = Layout<kAffine>;
  static constexpr int G = 2;
  static constexpr bool kTwice = true;
  float n, swt;
  float2 s1, s2;
  float2 hA[3];
  float hC[3];
  float2 za[2], zm[2], zt[2];

  __device__ __forceinline__ void clear() {
    const float2 o = f2(0.f);
    n = swt = 0.f; s1 = s2 = o;
#pragma unroll
    for (int m = 0; m < 3; ++m) { hA[m] = o; hC[m] = 0.f; }
#pragma unroll
    for (int m = 0; m < 2; ++m) za[m] = zm[m] = zt[m] = o;
  }

  // interior pixel (mask 1; the caller counts it): g01 = (2 gx, 2 gy)
  __device__ __forceinline__ void add_packed(float2 g01, float w_, float t_, float yf) {
    const float2 gy01 = mul2(g01, f2(yf));
    hA[0] = fma2(g01, g01, hA[0]); hA[1] = fma2(gy01, g01, hA[1]); hA[2] = fma2(gy01, gy01, hA[2]);
    hC[0] = fmaf(g01.x, g01.y, hC[0]); hC[1] = fmaf(gy01.x, g01.y, hC[1]); hC[2] = fmaf(gy01.x, gy01.y, hC[2]);
    const float2 wt = f2(w_, t_);
    za[0] = fma2(g01, f2(w_), za[0]); za[1] = fma2(gy01, f2(w_), za[1]);
    zt[0] = fma2(g01, f2(t_), zt[0]); zt[1] = fma2(gy01, f2(t_), zt[1]);
    zm[0] = add2(zm[0], g01);         zm[1] = add2(zm[1], gy01);
    s1 = add2(s1, wt);
    s2 = fma2(wt, wt, s2);
    swt = fmaf(w_, t_, swt);
  }

  // general pixel (rim / careful / slow paths), scalar arithmetic on the same registers; g = (2 gx, 2 gy)
  template <bool INTERIOR>
  __device__ __forceinline__ void add(const float (&g)[2], float w_, float t_, float mk, float yf) {
    const float y0 = g[0] * yf, y1 = g[1] * yf;
    hA[0].x = fmaf(g[0], g[0], hA[0].x); hA[1].x = fmaf(y0, g[0], hA[1].x); hA[2].x = fmaf(y0, y0, hA[2].x);
    hA[0].y = fmaf(g[1], g[1], hA[0].y); hA[1].y = fmaf(y1, g[1], hA[1].y); hA[2].y = fmaf(y1, y1, hA[2].y);
    hC[0] = fmaf(g[0], g[1], hC[0]); hC[1] = fmaf(y0, g[1], hC[1]); hC[2] = fmaf(y0, y1, hC[2]);
    const float wm = INTERIOR ? w_ : w_ * mk;
    const float tm = INTERIOR ? t_ : t_ * mk;
    za[0].x = fmaf(g[0], w_, za[0].x); za[1].x = fmaf(y0, w_, za[1].x);
    za[0].y = fmaf(g[1], w_, za[0].y); za[1].y = fmaf(y1, w_, za[1].y);
    zt[0].x = fmaf(g[0], tm, zt[0].x); zt[1].x = fmaf(y0, tm, zt[1].x);
    zt[0].y = fmaf(g[1], tm, zt[0].y); zt[1].y = fmaf(y1, tm, zt[1].y);
    if (INTERIOR) {
      zm[0].x += g[0]; zm[1].x += y0; zm[0].y += g[1]; zm[1].y += y1;
    } else {
      zm[0].x = fmaf(g[0], mk, zm[0].x); zm[1].x = fmaf(y0, mk, zm[1].x);
      zm[0].y = fmaf(g[1], mk, zm[0].y); zm[1].y = fmaf(y1, mk, zm[1].y);
      n += mk;
    }
    s1.x += wm; s2.x = fmaf(wm, w_, s2.x);
    s1.y += tm; s2.y = fmaf(tm, t_, s2.y);
    swt = fmaf(wm, t_, swt);
  }

  template <bool UNUSED>
  __device__ __forceinline__ void emit(float xf, float (&v)[L::NV]) const {
    v[0] = n; v[1] = s1.x; v[2] = s2.x; v[3] = s1.y; v[4] = s2.y; v[5] = swt;
    const float xx = xf * xf;
    // product order of Layout<>: (0,0) (0,1) (1,1)
#pragma unroll
    for (int idx = 0; idx < 3; ++idx) {
      float m[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) m[k] = idx == 0 ? hA[k].x : idx == 1 ? hC[k] : hA[k].y;
      const float p0 = m[0] * 0.25f, p1 = m[1] * 0.25f, p2 = m[2] * 0.25f;
      float* o = &v[L::kH + idx * 6];
      o[0] = p0; o[1] = xf * p0; o[2] = p1; o[3] = xx * p0; o[4] = xf * p1; o[5] = p2;
    }
    const float z0[3][2] = {{za[0].x, za[0].y}, {zm[0].x, zm[0].y}, {zt[0].x, zt[0].y}};
    const float z1[3][2] = {{za[1].x, za[1].y}, {zm[1].x, zm[1].y}, {zt[1].x, zt[1].y}};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float v0 = z0[a][i] * 0.5f, v1 = z1[a][i] * 0.5f;
        float* o = &v[L::kZ + (a * 2 + i) * 3];
        o[0] = v0; o[1] = xf * v0; o[2] = v1;
      }
  }
};

// ---- Jacobian generators (OpenCV evaluates them on f32 grids with the f32 matrix) --------------------
template <int MOTION> struct Jac {
  static constexpr int G = Model<MOTION>::G;
  float jc0, jc1, jc2, h3, h4, h5, ec, es, xf;
  __device__ __forceinline__ void init(const float* m, float xf_) {
    xf = xf_;
    jc0 = jc1 = jc2 = h3 = h4 = h5 = ec = es = 0.f;
    if (MOTION == kHomography) {
      jc0 = fmaf(xf, m[0], m[2]);     // X h0 + h6
      jc1 = fmaf(xf, m[3], m[5]);     // X h1 + h7
      jc2 = fmaf(xf, m[6], 1.0f);     // X h2 + 1
      h3 = m[1]; h4 = m[4]; h5 = m[7];
    } else if (MOTION == kEuclidean) {
      ec = m[0]; es = m[3];
    }
  }
  __device__ __forceinline__ void eval(const Sample& s, float yf, float (&g)[G]) const {
    if (MOTION == kTranslation || MOTION == kAffine) {
      g[0] = 0.5f * s.gx2; g[1] = 0.5f * s.gy2;
    } else if (MOTION == kEuclidean) {
      const float gx = 0.5f * s.gx2, gy = 0.5f * s.gy2;
      const float hx = -(xf * es) - yf * ec;
      const float hy = xf * ec - yf * es;
      g[0] = fmaf(gx, hx, gy * hy); g[1] = gx; g[G - 1] = gy;
    } else {
      const float den = fmaf(yf, h5, jc2);
      const float rden = rcp_approx(den);
      const float hx = -fmaf(yf, h3, jc0) * rden;
      const float hy = -fmaf(yf, h4, jc1) * rden;
      const float hr = 0.5f * rden;
      g[0] = s.gx2 * hr; g[1] = s.gy2 * hr;
      g[G - 1] = fmaf(hx, g[0], hy * g[1]);
    }
  }
};

// ---- epilogue ---------------------------------------------------------------------------------------
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

// Full-resolution matrix from the one estimated on the downscaled greys, in the reference's f32 arithmetic:
// the 2x3 models scale only the translation column (/root/reference/src/lib.rs:941-951), the homography
// goes through adjust_homography_for_scale_f32 (/root/reference/src/utils.rs:218-248).
__device__ inline void rescale_to_full(EccState* st, bool persp) {
  const float sx = st->rescale_x, sy = st->rescale_y;
  if (sx == 0.f) return;
  st->m[2] = __fmul_rn(st->m[2], sx);
  st->m[5] = __fmul_rn(st->m[5], sy);
  if (persp) {
    st->m[6] = __fdiv_rn(st->m[6], sx);
    st->m[7] = __fdiv_rn(st->m[7], sy);
  }
}

// One warp: lane r < P owns row r of the augmented system [H | ip | tp] in registers; Gauss-Jordan in
// f64 (H is symmetric positive definite: no pivoting), then lambda, delta-p, the f32 matrix update and the
// convergence test.  `tot` is the f64 totals vector in shared memory.
template <int MOTION>
__device__ __noinline__ void ecc_epilogue_warp(const double* tot, EccState* st, int lane) {
  using L = Layout<MOTION>;
  constexpr int P = Model<MOTION>::P, G = L::G;
  constexpr unsigned kFull = 0xffffffffu;
  const int r = lane < P ? lane : 0;
  const double n = tot[0], sw = tot[1], sww = tot[2], s_t = tot[3], stt = tot[4], swt = tot[5];
  const double wbar = sw / n, tbar = s_t / n;
  const double in2 = sww - sw * sw / n;
  const double tn2 = stt - s_t * s_t / n;
  const double corr = swt - s_t * sw / n;
  const double rho = corr / sqrt(in2 * tn2);

  // row r of H and of the three projections
  double row[P + 2];
  double a_r, am_r, b_r;
  {
    const int gk = Model<MOTION>::kron ? r % G : r;
    const int qk = Model<MOTION>::kron ? r / G : 0;
#pragma unroll
    for (int l = 0; l < P; ++l) {
      const int gl = Model<MOTION>::kron ? l % G : l;
      const int ql = Model<MOTION>::kron ? l / G : 0;
      const int i = gk < gl ? gk : gl, j = gk < gl ? gl : gk;
      const int pi = i * G - i * (i - 1) / 2 + (j - i);      // index of g_i g_j in the upper-triangular order
      if (Model<MOTION>::kron) {
        // moment of (q_k, q_l) in {1, X, Y, XX, XY, YY}; q = (X, Y, 1)
        const int lo = qk < ql ? qk : ql, hi = qk < ql ? ql : qk;
        const int mom = (lo == 0) ? (hi == 0 ? 3 : hi == 1 ? 4 : 1) : (lo == 1) ? (hi == 1 ? 5 : 2) : 0;
        row[l] = tot[L::kH + pi * 6 + mom];
      } else {
        row[l] = tot[L::kH + pi];
      }
    }
    if (Model<MOTION>::kron) {
      const int zm = qk == 0 ? 1 : qk == 1 ? 2 : 0;
      a_r = tot[L::kZ + (0 * G + gk) * 3 + zm];
      am_r = tot[L::kZ + (1 * G + gk) * 3 + zm];
      b_r = tot[L::kZ + (2 * G + gk) * 3 + zm];
    } else {
      a_r = tot[L::kZ + 0 * G + gk]; am_r = tot[L::kZ + 1 * G + gk]; b_r = tot[L::kZ + 2 * G + gk];
    }
  }
  const double ip_r = a_r - wbar * am_r;
  const double tp_r = b_r - tbar * am_r;
  row[P] = ip_r;
  row[P + 1] = tp_r;

  bool singular = false;
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const double pivot = __shfl_sync(kFull, row[k], k);
    if (!(fabs(pivot) > 0.0)) singular = true;
    const double rp = 1.0 / pivot;
    const double f = row[k] * rp;
#pragma unroll
    for (int j = k + 1; j < P + 2; ++j) {
      const double pk = __shfl_sync(kFull, row[j], k);
      row[j] = (lane == k) ? pk * rp : fma(-f, pk, row[j]);
    }
    row[k] = (lane == k) ? 1.0 : 0.0;
  }
  double y1 = row[P], y2 = row[P + 1];
  if (singular) { y1 = 0.0; y2 = 0.0; }          // Mat::inv() of a singular matrix is all zeros
  const bool act = lane < P;
  const double lam_n = in2 - warp_sum(act ? ip_r * y1 : 0.0);
  const double lam_d = corr - warp_sum(act ? tp_r * y1 : 0.0);

  int status = kStatusOk;
  double rho_out = rho;
  if (!(rho == rho)) status = kStatusNaN;                        // "NaN encountered."
  else if (lam_d <= 0.0) { status = kStatusNoConv; rho_out = -1.0; }   // "The algorithm stopped before its convergence."

  if (status == kStatusOk) {
    const double lam = lam_n / lam_d;
    const float dp = (float)(lam * y2 - y1);                     // deltaP is CV_32F
    float* m = st->m;
    if (MOTION == kEuclidean) {
      const float d0 = __shfl_sync(kFull, dp, 0), d1 = __shfl_sync(kFull, dp, 1), d2 = __shfl_sync(kFull, dp, 2);
      if (lane == 0) {
        const double th = (double)d0 + asin((double)m[3]);
        m[2] += d1; m[5] += d2;
        m[0] = m[4] = (float)cos(th);
        m[3] = (float)sin(th);
        m[1] = -m[3];
      }
    } else if (act) {
      int idx;
      if (MOTION == kTranslation) idx = lane == 0 ? 2 : 5;
      else if (MOTION == kAffine) idx = (lane & 1) * 3 + (lane >> 1);            // 0 3 1 4 2 5
      else idx = (lane % 3) * 3 + lane / 3;                                      // 0 3 6 1 4 7 2 5
      m[idx] += dp;
    }
  }
  __syncwarp();
  if (lane == 0) {
    const double prev = st->rho;
    st->last_rho = prev;
    st->rho = rho_out;
    const int iters = st->iters + 1;
    st->iters = iters;
    st->status = status;
    // for (i = 1; i <= maxIter && fabs(rho - last_rho) >= eps; ++i)
    st->cont = (status == kStatusOk && iters < st->max_iter && fabs(rho_out - prev) >= st->eps) ? 1 : 0;
  }
  __syncwarp();
}

// ---- warp reduction of a register vector -----------------------------------------------------------
// Butterfly "transpose-reduce": 64 values per lane are summed across the 32 lanes in 62 shuffles (each
// stage halves the values a lane still carries) instead of 64 x 5; lane l ends with the totals of values
// 2l and 2l+1.  Values beyond 64 (and vectors shorter than 32) take the plain 5-step butterfly.
template <int NV>
__device__ __forceinline__ void warp_reduce_vector(float (&v)[NV], int lane, float* out /* [NV] in smem */) {
  if constexpr (NV >= 32) {
    float r[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] = i < NV ? v[i] : 0.f;
#pragma unroll
    for (int half = 32, o = 16; half >= 2; half >>= 1, o >>= 1) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < half; ++i) {
        const float lo = r[i], hi = r[i + half];
        const float send = up ? lo : hi;
        const float keep = up ? hi : lo;
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    if (2 * lane < NV) out[2 * lane] = r[0];
    if (2 * lane + 1 < NV) out[2 * lane + 1] = r[1];
#pragma unroll
    for (int i = 64; i < NV; ++i) {
      const float t = warp_sum(v[i]);
      if (lane == 0) out[i] = t;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float t = warp_sum(v[i]);
      if (lane == 0) out[i] = t;
    }
  }
}

// ---- end of a block's work: publish its partial; the last block sums all partials and solves --------
template <int MOTION, int THREADS>
__device__ __forceinline__ void finish_iteration(const EccIterParams& p, EccState* st, const double* s_accum,
                                                 double* s_tot, int* s_last) {
  using L = Layout<MOTION>;
  constexpr int NV = L::NV;
  constexpr int kWarps = THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (p.timing_out && tid == 0) p.timing_out[(size_t)blockIdx.x * 4 + 1] = global_ns();
  for (int i = tid; i < NV; i += THREADS) p.partials[(size_t)i * p.tiles_pad + blockIdx.x] = s_accum[i];

  // ---- last block: deterministic cross-block sum + solve --------------------------------------------
  __threadfence();
  __syncthreads();
  if (p.timing_out && tid == 0) p.timing_out[(size_t)blockIdx.x * 4 + 2] = global_ns();
  if (tid == 0) {
    const unsigned int done = atomicAdd(&st->tile_counter, 1u);
    *s_last = (done == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  // value-major partials: lane l of a warp reads blocks l, l+32, ... of one value — 16 independent coalesced
  // loads in flight per lane and two values per round, so the whole sum costs a few L2 round trips
  // instead of one per block (measured: 26 us -> ~5 us at 270 partials).  Fixed order => deterministic.
  const int n_tiles = gridDim.x;
  for (int i = wid; i < NV; i += 2 * kWarps) {
    const int i2 = i + kWarps;
    const double* src0 = p.partials + (size_t)i * p.tiles_pad;
    const double* src1 = p.partials + (size_t)(i2 < NV ? i2 : i) * p.tiles_pad;
    double s0 = 0.0, s1 = 0.0;
    for (int base = 0; base < n_tiles; base += 512) {
      double v0[16], v1[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int t = base + lane + 32 * k;
        v0[k] = t < n_tiles ? __ldcg(src0 + t) : 0.0;
        v1[k] = t < n_tiles ? __ldcg(src1 + t) : 0.0;
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) { s0 += v0[k]; s1 += v1[k]; }
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (lane == 0) { s_tot[i] = s0; if (i2 < NV) s_tot[i2] = s1; }
  }
  __syncthreads();
  if (p.totals_out) for (int i = tid; i < NV; i += THREADS) p.totals_out[i] = s_tot[i];
  unsigned long long* tail = p.timing_out ? p.timing_out + (size_t)gridDim.x * 4 : nullptr;
  if (tail && tid == 0) tail[0] = global_ns();        // cross-block sum done
  if (wid == 0) {
    ecc_epilogue_warp<MOTION>(s_tot, st, lane);
    if (tail && lane == 0) tail[1] = global_ns();     // solve + update done
    if (lane == 0) {
      st->tile_counter = 0;
      if (st->cont == 0) { rescale_to_full(st, Model<MOTION>::persp); compute_inverse(st, Model<MOTION>::persp); }
      if (p.use_handle) cudaGraphSetConditional(p.handle, (unsigned)st->cont);
      if (tail) { tail[2] = global_ns(); tail[3] = (unsigned long long)blockIdx.x; }
    }
  }
}

// Work split.  Chunks are dealt strip-major to the persistent blocks in contiguous runs of (nearly) equal
// COST, where a chunk of the first or last strip counts rim_weight/8: those are the chunks whose samples can
// leave the image, which sends the warp that owns the rim columns down the careful path (measured: blocks
// working there took 45-47 us against 38 us elsewhere).  chunk_at() inverts the cumulative cost.
__device__ __forceinline__ int chunk_at(long long t, int n_strips, int cps, int wr) {
  const long long first = (long long)cps * wr;
  if (n_strips == 1 || t < first) return (int)(t / wr);
  const long long mid = (long long)(n_strips - 2) * cps * 8;
  if (t < first + mid) return cps + (int)((t - first) / 8);
  return cps * (n_strips - 1) + (int)((t - first - mid) / wr);
}
__device__ __forceinline__ void block_chunk_range(int n_strips, int cps, int wr, int& g0, int& g1) {
  const long long total_w = (long long)cps * wr * (n_strips > 1 ? 2 : 1) + (long long)max(n_strips - 2, 0) * cps * 8;
  g0 = chunk_at((long long)blockIdx.x * total_w / gridDim.x, n_strips, cps, wr);
  g1 = blockIdx.x + 1 == gridDim.x ? n_strips * cps
                                   : chunk_at((long long)(blockIdx.x + 1) * total_w / gridDim.x, n_strips, cps, wr);
}

template <int MOTION, bool FAST> struct AccumFor { using type = Accum<MOTION>; };
template <> struct AccumFor<kHomography, true> { using type = AccumH2; };

// ---- the iteration kernel ----------------------------------------------------------------------------
// EXACT = false (homography only): FastPersp coordinates for the boxed pixels; true: f64 everywhere.
//
// Persistent, balanced: the frame is cut into 128-column strips of 16-row chunks; the chunk list (strip
// major) is split evenly over the resident blocks, so every block owns a contiguous run of chunks — one
// or two "segments" (a run that crosses into the next strip).  Per segment the thread's column is fixed;
// at a segment end the block folds its sums into an f64 shared accumulator.
template <int MOTION, bool EXACT>
__global__ void __launch_bounds__(kEccThreads, 512 / kEccThreads) ecc_iter_kernel(const __grid_constant__ EccIterParams p) {
  using L = Layout<MOTION>;
  using Md = Model<MOTION>;
  constexpr int NV = L::NV, G = L::G;
  constexpr int kWarps = kEccThreads / 32;
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ float s_m[9];
  __shared__ int s_box[kMaxChunks][3];          // xlo (multiple of 4, or INT_MIN = no box), ylo, interior flag
  __shared__ alignas(8) uint64_t s_full[kEccStages];
  __shared__ alignas(8) uint64_t s_empty[kEccStages];
  __shared__ float s_red[kWarps][NV];
  __shared__ double s_accum[NV];
  __shared__ int s_last;
  __shared__ double s_tot[NV];

  EccState* st = p.st;
  // PDL: let the next kernel of the chain be launched now, and wait until the previous one (whose last block
  // writes the state read below) has completed and flushed.  Both are no-ops without a programmatic edge.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // a frame whose loop already stopped (a later kernel of the unrolled chain, or the host-driven loop)
  if (st->cont == 0) return;

  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  if (p.timing_out && tid == 0) p.timing_out[(size_t)blockIdx.x * 4 + 0] = global_ns();
  if (tid < 9) s_m[tid] = st->m[tid];
  if (tid < NV) s_accum[tid] = 0.0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kEccStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const int cps = p.chunks_per_strip;
  int g0, g1;
  block_chunk_range(p.n_strips, cps, p.rim_weight, g0, g1);
  int cc = 0;                       // chunks this block has consumed so far: drives stage and phase

  const int col = tid & (kEccStripW - 1);
  const int part = tid / kEccStripW;
  constexpr bool fast_coords = Md::persp && !EXACT;

  while (g0 < g1) {
    // ---- segment: chunks [c_first, c_first + nseg) of one strip ---------------------------------
    const int strip = g0 / cps, c_first = g0 - strip * cps;
    const int nseg = min(min(cps - c_first, g1 - g0), kMaxChunks);
    const int x0 = strip * kEccStripW;
    const int x1 = min(x0 + kEccStripW, p.width) - 1;       // inclusive
    const int y0 = c_first * kChunkH;
    const int y1 = min(y0 + nseg * kChunkH, p.height);      // exclusive
    __syncthreads();    // s_m / barriers ready (first pass); previous segment fully consumed (later passes)

    // box table: for each chunk, where its 144x32 window of I starts — or "no box" when the sample
    // positions of the chunk do not fit one.  TMA zero-fills whatever part of a box lies outside the
    // image; pixels whose taps touch the border are handled by the careful variant (see `safe` below).
    if (tid < nseg) {
      const int cy0 = y0 + tid * kChunkH;
      const int cy1 = min(cy0 + kChunkH, y1) - 1;
      double umin, umax, vmin, vmax;
      bool ok = chunk_bounds<Md::persp>(s_m, x0, x1, cy0, cy1, umin, umax, vmin, vmax);
      int xlo = 0, ylo = 0;
      if (ok) ok = fabs(umin) < 1e8 && fabs(umax) < 1e8 && fabs(vmin) < 1e8 && fabs(vmax) < 1e8;
      if (ok) {
        // the innermost TMA coordinate must land on a 16-byte boundary (4 floats): an unaligned start
        // raises "illegal instruction" on sm_100 (measured, scripts/tma_probe3.cu)
        xlo = ((int)floor(umin) - 2) & ~3;
        ylo = (int)floor(vmin) - 2;
        ok = ((int)floor(umax) + 3 - xlo < kBoxW) && ((int)floor(vmax) + 3 - ylo < kBoxH);
      }
      s_box[tid][0] = ok ? xlo : INT_MIN;
      s_box[tid][1] = ylo;
      // interior chunk: every sample of the chunk (quantised position within 1/64 px + f32 slack of the real
      // one) has its four taps and their gradient stencils off the border rows/columns, so no border rule
      // applies, the mask is 1 everywhere and the pixel loop needs no per-pixel test
      s_box[tid][2] = (ok && floor(umin - 0.0625) >= 1.0 && floor(umax + 0.0625) <= (double)(p.width - 3) &&
                       floor(vmin - 0.0625) >= 1.0 && floor(vmax + 0.0625) <= (double)(p.height - 3)) ? 1 : 0;
    }
    __syncthreads();

    // producer (thread 0): TMA issue for local chunk c (global sequence number cc + c)
    auto issue = [&](int c) {
      const int gc = cc + c;
      const int s = gc % kEccStages;
      // the stage last held sequence number gc - kEccStages: wait until every warp has released it
      if (gc >= kEccStages) mbar_wait(&s_empty[s], (unsigned)(gc / kEccStages - 1) & 1u);
      const int xlo = s_box[c][0], ylo = s_box[c][1];
      const bool boxed = xlo != INT_MIN;
      unsigned char* stage = dyn + s * kStageBytes;
      mbar_expect_tx(&s_full[s], (boxed ? kImgStageBytes : 0) + kTmplStageBytes);
      if (boxed) tma_load_2d(stage, &p.tm_img, xlo, ylo, &s_full[s]);
      tma_load_2d(stage + kImgStageBytes, &p.tm_tmpl, x0, y0 + c * kChunkH, &s_full[s]);
    };
    if (tid == 0) {
      for (int c = 0; c < kEccStages - 1 && c < nseg; ++c) issue(c);
    }

    const int x = x0 + col;
    const float xf = (float)x;
    const bool col_ok = x < p.width;
    typename AccumFor<MOTION, fast_coords>::type acc;
    acc.clear();
    int n_safe = 0;
    FastPersp fp;
    if (fast_coords) fp.init(s_m, x);

    // general path for one pixel: exact coordinates, every border rule, nearest-neighbour mask
    auto slow_pixel = [&](int y, float t_) {
      Coord<Md::persp> co;
      Jac<MOTION> jac;
      co.init(s_m, x);
      jac.init(s_m, xf);
      int xq, yq, xn, yn;
      const bool ok = co.at_with_nearest(y, xq, yq, xn, yn);
      Sample smp; smp.w = 0.f; smp.gx2 = 0.f; smp.gy2 = 0.f;
      float mk = 0.f;
      if (ok) {
        const int sx = xq >> kInterBits, sy = yq >> kInterBits;
        if (sx >= -1 && sx < p.width && sy >= -1 && sy < p.height) {
          const float ax = (float)(xq & (kInterTab - 1)) * (1.f / kInterTab);
          const float ay = (float)(yq & (kInterTab - 1)) * (1.f / kInterTab);
          smp = sample_general(p.img, p.pitch, p.width, p.height, sx, sy, ax, ay);
        }
        // OpenCV warps an all-ones mask with INTER_NEAREST: 1 where the rounded position is inside
        mk = ((unsigned)xn < (unsigned)p.width && (unsigned)yn < (unsigned)p.height) ? 1.f : 0.f;
      }
      float g[G];
      const float yf = (float)y;
      jac.eval(smp, yf, g);
      if (fast_coords) { g[0] *= 2.f; g[1] *= 2.f; g[G - 1] *= -2.f; }     // the run accumulates (2a, 2b, -2t)
      acc.template add<false>(g, smp.w, t_, mk, yf);
    };

    for (int c = 0; c < nseg; ++c) {
      const int gc = cc + c;
      const int s = gc % kEccStages;
      if (tid == 0 && c + kEccStages - 1 < nseg) issue(c + kEccStages - 1);
      __syncwarp();
      mbar_wait(&s_full[s], (unsigned)(gc / kEccStages) & 1u);
      const float* box = reinterpret_cast<const float*>(dyn + s * kStageBytes);
      const float* tbox = reinterpret_cast<const float*>(dyn + s * kStageBytes + kImgStageBytes);
      const int xlo = s_box[c][0], ylo = s_box[c][1];
      const bool boxed = xlo != INT_MIN;
      const int cy0 = y0 + c * kChunkH;
      const int ya = cy0 + part * kChunkRowsPerThread;
      const int yb = min(ya + kChunkRowsPerThread, y1);
      // lanes beyond the image width sit the chunk out; warp votes below use the mask of the lanes that work
      const unsigned wmask = __ballot_sync(0xffffffffu, col_ok && ya < yb);
      if (col_ok && ya < yb) {
        const float* trow = tbox + (ya - cy0) * kEccStripW + col;
        if (boxed && s_box[c][2] != 0 && yb - ya == kChunkRowsPerThread) {
          // lean path (interior chunk, full height): straight-line code for the thread's 8 rows — no border
          // rule, no mask, no vote, no branch — so the rows interleave freely in the schedule
          const float* bp0 = box + (ya - ylo) * kBoxW + (x - xlo);
          const float yf0 = (float)ya;
          if constexpr (fast_coords) {
#pragma unroll
            for (int r = 0; r < kChunkRowsPerThread; ++r) {
              const float yf = yf0 + (float)r;
              const float t_ = trow[r * kEccStripW];
              // FastPersp::at with the two coordinates carried as a pair
              const float rw = rcp_approx(fmaf(fp.m21, yf, fp.wc));
              const float2 d = mul2(f2(fmaf(fp.beta, yf, fp.alpha), fmaf(fmaf(-fp.m21, yf, fp.delta), yf, fp.gamma)), f2(rw));
              const float2 qf = fma2(d, f2(32.0f), f2(12582912.0f));
              const int qx = __float_as_int(qf.x) - 0x4B400000, qy = __float_as_int(qf.y) - 0x4B400000;
              const float* bp = bp0 + ((qy >> kInterBits) + r) * kBoxW + (qx >> kInterBits);
              const float ax = (float)(qx & (kInterTab - 1)) * (1.f / kInterTab);
              const float ay = (float)(qy & (kInterTab - 1)) * (1.f / kInterTab);
              float w_;
              float2 gxy2;
              sample_box_packed(bp, ax, ay, w_, gxy2);
              const float2 g01 = mul2(gxy2, f2(rw));                          // 2a, 2b
              const float2 uv = add2(f2(xf, yf), d);                          // sample position (u, v)
              const float g2 = fmaf(uv.x, g01.x, uv.y * g01.y);               // -2t  (t = hatX a + hatY b, hat = -(u, v))
              acc.add_packed(g01, g2, w_, t_, yf);
            }
          } else {
            // the other instantiations (2x3 models with OpenCV's exact 10-bit fixed point, homography with exact
            // f64 coordinates): same straight-line shape, exact coordinates, scalar sums
            Coord<Md::persp> co;
            Jac<MOTION> jac;
            co.init(s_m, x);
            jac.init(s_m, xf);
#pragma unroll
            for (int r = 0; r < kChunkRowsPerThread; ++r) {
              const float yf = yf0 + (float)r;
              const float t_ = trow[r * kEccStripW];
              int xq, yq;
              co.at(ya + r, xq, yq);
              const float* bp = box + ((yq >> kInterBits) - ylo) * kBoxW + ((xq >> kInterBits) - xlo);
              const float ax = (float)(xq & (kInterTab - 1)) * (1.f / kInterTab);
              const float ay = (float)(yq & (kInterTab - 1)) * (1.f / kInterTab);
              const Sample smp = sample_box(bp, ax, ay);
              float g[G];
              jac.eval(smp, yf, g);
              acc.template add<true>(g, smp.w, t_, 1.f, yf);
            }
          }
          n_safe += kChunkRowsPerThread;
        } else if (boxed) {
          float yf = (float)ya;
          Coord<Md::persp> co;
          Jac<MOTION> jac;
          if (!fast_coords) { co.init(s_m, x); jac.init(s_m, xf); }
#pragma unroll 2
          for (int y = ya; y < yb; ++y, yf += 1.0f) {
            float g[G];
            Sample smp;
            const float t_ = trow[(y - ya) * kEccStripW];
            int sx, sy, xn = 0, yn = 0;   // integer sample position / nearest position in the image
            float ax, ay, du = 0.f, dv = 0.f, rw = 0.f;
            if (fast_coords) {
              int qx, qy;
              fp.at(yf, qx, qy, du, dv, rw);
              sx = x + (qx >> kInterBits);
              sy = y + (qy >> kInterBits);
              ax = (float)(qx & (kInterTab - 1)) * (1.f / kInterTab);
              ay = (float)(qy & (kInterTab - 1)) * (1.f / kInterTab);
            } else {
              int xq, yq;
              co.at_with_nearest(y, xq, yq, xn, yn);
              sx = xq >> kInterBits;
              sy = yq >> kInterBits;
              ax = (float)(xq & (kInterTab - 1)) * (1.f / kInterTab);
              ay = (float)(yq & (kInterTab - 1)) * (1.f / kInterTab);
            }
            // safe: the four bilinear taps and their gradient stencils stay off the border rows/columns,
            // so no border rule applies and the mask is 1.  The branch is taken warp-wide: a warp on the
            // rim runs the careful variant for all its lanes (no divergence), every other warp the plain one.
            const bool safe = (unsigned)(sx - 1) <= (unsigned)(p.width - 4) && (unsigned)(sy - 1) <= (unsigned)(p.height - 4);
            const float* bp = box + (sy - ylo) * kBoxW + (sx - xlo);
            const bool all_safe = __all_sync(wmask, safe);
            float mk = 1.f;
            if (all_safe) {
              smp = sample_box(bp, ax, ay);
            } else {
              smp = sample_box_rules(bp, ax, ay, sx, sy, p.width, p.height);
              if (fast_coords) { FastPersp::nearest(du, dv, xn, yn); xn += x; yn += y; }
              // OpenCV warps an all-ones mask with INTER_NEAREST: 1 where the rounded position is inside
              mk = ((unsigned)xn < (unsigned)p.width && (unsigned)yn < (unsigned)p.height) ? 1.f : 0.f;
            }
            if (fast_coords) {
              // a = gx / den, b = gy / den, t = hatX a + hatY b with hatX = -u, hatY = -v, den = w;
              // accumulated as (2a, 2b, -2t), see Accum::emit
              g[0] = smp.gx2 * rw; g[1] = smp.gy2 * rw;
              g[G - 1] = fmaf(xf + du, g[0], (yf + dv) * g[1]);
            } else {
              jac.eval(smp, yf, g);
            }
            if (all_safe) { acc.template add<true>(g, smp.w, t_, 1.f, yf); ++n_safe; }
            else acc.template add<false>(g, smp.w, t_, mk, yf);
          }
        } else {
          for (int y = ya; y < yb; ++y) slow_pixel(y, trow[(y - ya) * kEccStripW]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[s]);      // this warp is done with stage s
    }
    acc.n += (float)n_safe;

    // ---- segment end: registers -> warp transpose-reduce -> smem -> f64 block accumulator ----------
    {
      float v[NV];
      acc.template emit<fast_coords>(xf, v);
      warp_reduce_vector<NV>(v, lane, s_red[wid]);
    }
    __syncthreads();
    if (tid < NV) {
      double sres = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) sres += (double)s_red[w][tid];
      s_accum[tid] += sres;
    }
    cc += nseg;
    g0 += nseg;
  }
  finish_iteration<MOTION, kEccThreads>(p, st, s_accum, s_tot, &s_last);
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
