// K4 — final warp + accumulate, and K5 — lane sum + scale.
//
// K4 replaces, per frame, Mat::convert_to(CV_32F, 1/255) (/root/reference/src/utils.rs:133),
// imgproc::warp_affine | warp_perspective(INTER_LINEAR, BORDER_CONSTANT) and the MatExpr `acc + warped`
// (/root/reference/src/lib.rs:780-814; keypoint_match tail :289-316).  The f32 colour frame and the
// warped frame are never materialised: the kernel gathers the 8-bit source, converts the four taps,
// interpolates and adds into the f32 accumulator.  Algorithmic traffic 3N (u8 gather) + 12N + 12N.
//
// Bit-exactness vs OpenCV (oracle/restate.py::warp_linear, pinned against cv2): destination -> source
// coordinates in f64 with OpenCV's operation order (no FMA contraction), quantised to 1/32 px
// (perspective: round-half-even of 32*u, 64-px column blocks; affine: 10-bit fixed point), weights
// from the 5-bit fractions, value = s00*w00 + s01*w01 + s10*w10 + s11*w11 summed left to right in f32,
// taps outside the source replaced by the border value.
//
// K5 replaces Rayon's try_reduce of the per-thread partial sums and the final MatExpr `/ n`
// (/root/reference/src/lib.rs:819-839, :339-346): out = (sum over lanes) * float(1/n).
#pragma once
#include "common.cuh"

namespace stk {

constexpr int kWarpBatch = 4;             // frames gathered per launch: the accumulator is read and written once per batch

struct WarpFrame {
  const uint8_t* src;     // u8 interleaved
  size_t src_pitch;
  const double* inv_ptr;  // device pointer to the inverse map (ECC path), or null -> inv[]
  const int* status_ptr;  // device pointer to the frame's ECC status (frame skipped when != 0), or null
  double inv[9];
  float border[4];
  int border_mode;        // cv::BorderTypes: 0 CONSTANT, 1 REPLICATE, 2 REFLECT, 3 WRAP, 4 REFLECT_101
  int pad;
};

struct WarpAccParams {
  WarpFrame f[kWarpBatch];
  int n;                  // frames in this launch, 1..kWarpBatch, added in index order
  float* acc;             // f32 interleaved, width*channels floats per row, dense
  int width, height;      // destination size (== accumulator size)
  int src_width, src_height;
  int store;              // 1: acc = sum of the batch (first batch on this lane), 0: acc += ...
  unsigned frac_magic;    // 0x4B400000 handed over in a REGISTER: (q & 31) | magic is then one LOP3 (two with an immediate)
};

// Correctly rounded 1/w for w in the normal range (|w| in [2^-500, 2^500]): the instruction sequence of
// __drcp_rn without its special-case branch.  scripts/rcp_check.cu compares it with IEEE division over 2^28
// values on the GPU (0 mismatches).
__device__ __forceinline__ double rcp_rn_normal(double w) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(w));
  double t = __fma_rn(-w, r, 1.0);
  t = __fma_rn(t, t, t);
  r = __fma_rn(r, t, r);
  const double e = __fma_rn(-w, r, 1.0);
  return __fma_rn(r, e, r);
}

// cv::borderInterpolate for the modes that always yield a source index (REPLICATE 1, REFLECT 2, WRAP 3,
// REFLECT_101 4), restated in oracle/restate.py::border_interpolate and pinned against cv2 there
__device__ __forceinline__ int border_index(int p, int len, int mode) {
  if ((unsigned)p < (unsigned)len) return p;
  if (mode == 1) return p < 0 ? 0 : len - 1;
  if (mode == 3) { const int r = p % len; return r < 0 ? r + len : r; }
  if (len == 1) return 0;
  const int delta = mode == 4 ? 1 : 0;
  do {
    if (p < 0) p = -p - 1 + delta;
    else p = len - 1 - (p - len) - delta;
  } while ((unsigned)p >= (unsigned)len);
  return p;
}

constexpr int kWarpBX = 32, kWarpBY = 8;   // thread block
constexpr int kWarpRows = 4;               // destination rows per thread: the block tile is 32 x 32 pixels
constexpr int kWarpTH = kWarpBY * kWarpRows;

// Packed f32x2 multiply, written as PTX.  Only the MULTIPLIES of the blend are packed: ptxas fuses a packed multiply
// feeding a packed add into FFMA2 even for the `.rn` forms (seen in the SASS of the first batched build, with or
// without -fmad=false, and as a bit mismatch against cv2 at 6000x4000), so the adds stay scalar __fadd_rn, which is
// never fused.
__device__ __forceinline__ float2 mul2_rn_exact(float2 a, float2 b) {
  float2 d;
  asm("{\n .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mul.rn.f32x2 rd, ra, rb;\n mov.b64 {%0, %1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// value = s00*w00 + s01*w01 + s10*w10 + s11*w11, left to right in f32 (OpenCV's remapBilinear order), for the C channels
// of one pixel; taps are raw bytes, converted as fl(byte * fl(1/255)).  The multiplies of a channel pair ride the packed
// f32x2 form (each half is an IEEE round-to-nearest multiply, so the values are those of the scalar sequence).
template <int C>
__device__ __forceinline__ void blend_taps(const unsigned (&t00)[C], const unsigned (&t01)[C], const unsigned (&t10)[C],
                                           const unsigned (&t11)[C], float w00, float w01, float w10, float w11, float* out) {
  const float k255 = (float)(1.0 / 255.0);
  constexpr int P = C / 2;
#pragma unroll
  for (int q = 0; q < P; ++q) {
    const float2 k2 = make_float2(k255, k255);
    const float2 s00 = mul2_rn_exact(make_float2((float)t00[2 * q], (float)t00[2 * q + 1]), k2);
    const float2 s01 = mul2_rn_exact(make_float2((float)t01[2 * q], (float)t01[2 * q + 1]), k2);
    const float2 s10 = mul2_rn_exact(make_float2((float)t10[2 * q], (float)t10[2 * q + 1]), k2);
    const float2 s11 = mul2_rn_exact(make_float2((float)t11[2 * q], (float)t11[2 * q + 1]), k2);
    const float2 p00 = mul2_rn_exact(s00, make_float2(w00, w00)), p01 = mul2_rn_exact(s01, make_float2(w01, w01));
    const float2 p10 = mul2_rn_exact(s10, make_float2(w10, w10)), p11 = mul2_rn_exact(s11, make_float2(w11, w11));
    float2 v;
    v.x = __fadd_rn(__fadd_rn(__fadd_rn(p00.x, p01.x), p10.x), p11.x);
    v.y = __fadd_rn(__fadd_rn(__fadd_rn(p00.y, p01.y), p10.y), p11.y);
    out[2 * q] = v.x;
    out[2 * q + 1] = v.y;
  }
  if (C & 1) {
    constexpr int c = C - 1;
    const float s00 = __fmul_rn((float)t00[c], k255), s01 = __fmul_rn((float)t01[c], k255);
    const float s10 = __fmul_rn((float)t10[c], k255), s11 = __fmul_rn((float)t11[c], k255);
    out[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s00, w00), __fmul_rn(s01, w01)), __fmul_rn(s10, w10)), __fmul_rn(s11, w11));
  }
}

// acc[c] = fl(acc[c] + v[c]) on the thread's OWN pixel slot of the block's accumulator tile in shared memory (lane
// stride C floats: conflict-free for C = 3; one 128-bit access for C = 4)
template <int C>
__device__ __forceinline__ void tile_add(float* slot, const float (&v)[C]) {
  if constexpr (C == 4) {
    float4 a = *reinterpret_cast<float4*>(slot);
    a.x = __fadd_rn(a.x, v[0]); a.y = __fadd_rn(a.y, v[1]); a.z = __fadd_rn(a.z, v[2]); a.w = __fadd_rn(a.w, v[3]);
    *reinterpret_cast<float4*>(slot) = a;
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) slot[c] = __fadd_rn(slot[c], v[c]);
  }
}

// One destination pixel by the general rules (any map, any border mode): coordinates in f64 with OpenCV's operation
// order and saturation, every tap either inside the source or replaced per the border mode.  (cx0 .. cw1, adelta,
// bdelta: the per-column terms of the caller.)
template <int C, bool PERSP>
__device__ __forceinline__ void general_pixel(const WarpFrame& f, const double* mtx, double cx0, double cx1, double cy0, double cy1,
                                              double cw0, double cw1, int adelta, int bdelta, int y, int sw, int sh, float (&v)[C]) {
  const double yd = (double)y;
  int xq, yq;
  if (PERSP) {
    const double X0 = __dadd_rn(__dadd_rn(cx0, __dmul_rn(mtx[1], yd)), mtx[2]);
    const double Y0 = __dadd_rn(__dadd_rn(cy0, __dmul_rn(mtx[4], yd)), mtx[5]);
    const double W0 = __dadd_rn(__dadd_rn(cw0, __dmul_rn(mtx[7], yd)), mtx[8]);
    const double W = __dadd_rn(W0, cw1);
    const double W32 = (W != 0.0) ? __dmul_rn(__drcp_rn(W), (double)kInterTab) : 0.0;      // == INTER_TAB_SIZE / W
    // cvRound with saturation == clamp to [INT_MIN, INT_MAX] then round half to even
    xq = __double2int_rn(__dmul_rn(__dadd_rn(X0, cx1), W32));
    yq = __double2int_rn(__dmul_rn(__dadd_rn(Y0, cy1), W32));
  } else {
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(mtx[1], yd), mtx[2]), kAbScale)) + 16;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(mtx[4], yd), mtx[5]), kAbScale)) + 16;
    xq = (int)((unsigned)X0 + (unsigned)adelta) >> (kAbBits - kInterBits);
    yq = (int)((unsigned)Y0 + (unsigned)bdelta) >> (kAbBits - kInterBits);
  }
  // integer part is stored as short in OpenCV's map (saturate_cast<short>)
  const int sx = max(-32768, min(32767, xq >> kInterBits));
  const int sy = max(-32768, min(32767, yq >> kInterBits));
  const float ax = (float)(xq & (kInterTab - 1)) * (1.f / kInterTab);
  const float ay = (float)(yq & (kInterTab - 1)) * (1.f / kInterTab);
  const float w00 = __fmul_rn(1.f - ay, 1.f - ax), w01 = __fmul_rn(1.f - ay, ax);
  const float w10 = __fmul_rn(ay, 1.f - ax), w11 = __fmul_rn(ay, ax);
  const float k255 = (float)(1.0 / 255.0);
  const uint8_t* r0 = f.src + (ptrdiff_t)sy * (ptrdiff_t)f.src_pitch + (ptrdiff_t)sx * C;
  const uint8_t* r1 = r0 + f.src_pitch;
  if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
    // all four taps inside the source: the common case, no border logic
    unsigned t00[C], t01[C], t10[C], t11[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { t00[c] = __ldg(r0 + c); t01[c] = __ldg(r0 + C + c); t10[c] = __ldg(r1 + c); t11[c] = __ldg(r1 + C + c); }
    blend_taps<C>(t00, t01, t10, t11, w00, w01, w10, w11, v);
  } else if (f.border_mode != 0) {
    // REPLICATE / REFLECT / WRAP / REFLECT_101: each tap's coordinates go through borderInterpolate on their own
    const int x0 = border_index(sx, sw, f.border_mode), x1 = border_index(sx + 1, sw, f.border_mode);
    const int y0 = border_index(sy, sh, f.border_mode), y1 = border_index(sy + 1, sh, f.border_mode);
    const uint8_t* q0 = f.src + (size_t)y0 * f.src_pitch;
    const uint8_t* q1 = f.src + (size_t)y1 * f.src_pitch;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float s00 = __fmul_rn((float)__ldg(q0 + x0 * C + c), k255), s01 = __fmul_rn((float)__ldg(q0 + x1 * C + c), k255);
      const float s10 = __fmul_rn((float)__ldg(q1 + x0 * C + c), k255), s11 = __fmul_rn((float)__ldg(q1 + x1 * C + c), k255);
      v[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s00, w00), __fmul_rn(s01, w01)), __fmul_rn(s10, w10)),
                       __fmul_rn(s11, w11));
    }
  } else if (sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0) {
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = f.border[c];
  } else {
    const bool x0in = (unsigned)sx < (unsigned)sw, x1in = (unsigned)(sx + 1) < (unsigned)sw;
    const bool y0in = (unsigned)sy < (unsigned)sh, y1in = (unsigned)(sy + 1) < (unsigned)sh;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float s00 = (y0in && x0in) ? __fmul_rn((float)__ldg(r0 + c), k255) : f.border[c];
      const float s01 = (y0in && x1in) ? __fmul_rn((float)__ldg(r0 + C + c), k255) : f.border[c];
      const float s10 = (y1in && x0in) ? __fmul_rn((float)__ldg(r1 + c), k255) : f.border[c];
      const float s11 = (y1in && x1in) ? __fmul_rn((float)__ldg(r1 + C + c), k255) : f.border[c];
      v[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s00, w00), __fmul_rn(s01, w01)), __fmul_rn(s10, w10)),
                       __fmul_rn(s11, w11));
    }
  }
}

// One frame's contribution to the block's 32 x 32 tile: the C interpolated values of each of the thread's four pixels
// are ADDED to the accumulator tile s_acc (tile row r at s_acc + r * 32 * C).  Every thread owns its pixels' slots for
// the whole launch, so consecutive frames need no barrier.  mtx = the frame's inverse map (9 doubles, shared memory).
template <int C, bool PERSP>
__device__ __forceinline__ void warp_tile(const WarpFrame& f, const double* mtx, float* s_acc, int width, int height,
                                          int sw, int sh) {
  const int x = blockIdx.x * kWarpBX + threadIdx.x;
  const int y_base = blockIdx.y * kWarpTH + threadIdx.y;
  constexpr int kRow = kWarpBX * C;

  // per-thread (column) terms
  double cx0 = 0, cx1 = 0, cy0 = 0, cy1 = 0, cw0 = 0, cw1 = 0;
  int adelta = 0, bdelta = 0;
  if (PERSP) {
    // WarpPerspectiveInvoker: X0/Y0/W0 at the start of the 64-px column block, then + M*x1
    const int xb = width >= 64 ? (x & ~63) : 0;
    const double xbd = (double)xb, x1 = (double)(x - xb);
    cx0 = __dmul_rn(mtx[0], xbd); cx1 = __dmul_rn(mtx[0], x1);
    cy0 = __dmul_rn(mtx[3], xbd); cy1 = __dmul_rn(mtx[3], x1);
    cw0 = __dmul_rn(mtx[6], xbd); cw1 = __dmul_rn(mtx[6], x1);
  } else {
    const double xd = (double)x;
    adelta = __double2int_rn(__dmul_rn(__dmul_rn(mtx[0], xd), kAbScale));
    bdelta = __double2int_rn(__dmul_rn(__dmul_rn(mtx[3], xd), kAbScale));
  }

  // Interior tiles (the bulk of a frame): when the four corners of the tile land inside
  // [1/16, sw-1-1/16] x [1/16, sh-1-1/16] with w in (1e-9, 1e9), every tap of every pixel of the tile is inside
  // the source (a projective map with w > 0 sends the rectangle into the hull of its corner images; the slack
  // covers the 1/32-px quantisation), so the pixel loop needs no bounds test, clamp or border select, the
  // reciprocal no special cases and cvRound no saturation.  Decided per warp (lanes 0-3 take one corner each,
  // inequalities multiplied through by w: no division), no extra barrier.
  bool lean = blockIdx.x * kWarpBX + kWarpBX <= width && blockIdx.y * kWarpTH + kWarpTH <= height;
  {
    const int k = threadIdx.x & 3;
    const double cxk = (double)(blockIdx.x * kWarpBX + ((k & 1) ? kWarpBX - 1 : 0));
    const double cyk = (double)(blockIdx.y * kWarpTH + ((k & 2) ? kWarpTH - 1 : 0));
    const double nu = mtx[0] * cxk + mtx[1] * cyk + mtx[2], nv = mtx[3] * cxk + mtx[4] * cyk + mtx[5];
    const double ww = PERSP ? mtx[6] * cxk + mtx[7] * cyk + mtx[8] : 1.0;
    const double lo = 0.0625 * ww, uhi = ((double)(sw - 1) - 0.0625) * ww, vhi = ((double)(sh - 1) - 0.0625) * ww;
    lean = __all_sync(0xffffffffu, lean && ww > 1e-9 && ww < 1e9 && nu >= lo && nu < uhi && nv >= lo && nv < vhi);
  }

  if (lean) {
#pragma unroll
    for (int rr = 0; rr < kWarpRows; ++rr) {
      const int ry = threadIdx.y + rr * kWarpBY;
      const double yd = (double)(y_base + rr * kWarpBY);
      int xq, yq;
      if (PERSP) {
        const double X0 = __dadd_rn(__dadd_rn(cx0, __dmul_rn(mtx[1], yd)), mtx[2]);
        const double Y0 = __dadd_rn(__dadd_rn(cy0, __dmul_rn(mtx[4], yd)), mtx[5]);
        const double W0 = __dadd_rn(__dadd_rn(cw0, __dmul_rn(mtx[7], yd)), mtx[8]);
        const double W32 = __dmul_rn(rcp_rn_normal(__dadd_rn(W0, cw1)), (double)kInterTab);
        xq = rint_magic(__dmul_rn(__dadd_rn(X0, cx1), W32));
        yq = rint_magic(__dmul_rn(__dadd_rn(Y0, cy1), W32));
      } else {
        const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(mtx[1], yd), mtx[2]), kAbScale)) + 16;
        const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(mtx[4], yd), mtx[5]), kAbScale)) + 16;
        xq = (int)((unsigned)X0 + (unsigned)adelta) >> (kAbBits - kInterBits);
        yq = (int)((unsigned)Y0 + (unsigned)bdelta) >> (kAbBits - kInterBits);
      }
      const float ax = (float)(xq & (kInterTab - 1)) * (1.f / kInterTab);
      const float ay = (float)(yq & (kInterTab - 1)) * (1.f / kInterTab);
      const float w00 = __fmul_rn(1.f - ay, 1.f - ax), w01 = __fmul_rn(1.f - ay, ax);
      const float w10 = __fmul_rn(ay, 1.f - ax), w11 = __fmul_rn(ay, ax);
      const uint8_t* r0 = f.src + (ptrdiff_t)(yq >> kInterBits) * (ptrdiff_t)f.src_pitch + (xq >> kInterBits) * C;
      const uint8_t* r1 = r0 + f.src_pitch;
      unsigned t00[C], t01[C], t10[C], t11[C];
#pragma unroll
      for (int c = 0; c < C; ++c) { t00[c] = __ldg(r0 + c); t01[c] = __ldg(r0 + C + c); t10[c] = __ldg(r1 + c); t11[c] = __ldg(r1 + C + c); }
      float v[C];
      blend_taps<C>(t00, t01, t10, t11, w00, w01, w10, w11, v);
      tile_add<C>(s_acc + ry * kRow + threadIdx.x * C, v);
    }
    return;
  }
#pragma unroll
  for (int rr = 0; rr < kWarpRows; ++rr) {
    const int y = y_base + rr * kWarpBY;
    if (x < width && y < height) {
      float v[C];
      general_pixel<C, PERSP>(f, mtx, cx0, cx1, cy0, cy1, cw0, cw1, adelta, bdelta, y, sw, sh, v);
      tile_add<C>(s_acc + (threadIdx.y + rr * kWarpBY) * kRow + threadIdx.x * C, v);
    }
  }
}

// 32 x 32 destination tile per 256-thread block, four rows per thread (so the per-thread column terms of the
// coordinate transform, the parameter loads and the index math are paid once per four pixels).
// BATCHED: up to kWarpBatch frames per launch.  The block loads its tile of the accumulator into shared memory once
// (coalesced 128-bit accesses: a tile row is 32*C contiguous floats), every thread then adds frame after frame, in
// index order, into the slots of its own four pixels — the same f32 sequence as one launch per frame, so results are
// bit-identical to the unbatched form — and the tile is written back once: 3N + 24N/k bytes per frame instead of 27N.
// Between frames there is NO barrier (a thread touches only its own slots), so the warps of a block drift apart and
// the byte gathers of one overlap the arithmetic of the others; the accumulator costs no registers, which keeps the
// kernel at 8 blocks per SM (ncu r2: the register-resident form ran at 5 blocks per SM with `long_scoreboard` on the
// gathers as the top stall).
// ncu (profiles/) showed the first versions to be issue-bound (390, then 234 instructions per pixel), so the
// per-pixel path is kept lean: the inverse maps are staged once per block, the 64-px column block start is a mask,
// the reciprocal replaces the f64 divide (32/W == 32 * (1/W) exactly: a power-of-two scaling commutes with
// rounding), __double2int_rn supplies the saturation, pixels whose four taps are inside the source skip all border
// selects, and the multiplies of a channel pair are packed f32x2 instructions.
// A frame whose ECC status is non-zero contributes nothing (the reference aborts the whole stack, src/lib.rs:777);
// in store mode the accumulator is still written (zeros + the other frames), never left uninitialised.
template <int C, bool PERSP>
__global__ void __launch_bounds__(kWarpBX * kWarpBY, 8) warp_accumulate_kernel(const __grid_constant__ WarpAccParams p) {
  constexpr int kRow = kWarpBX * C;                 // floats per tile row
  constexpr int kTile = kWarpTH * kRow;
  constexpr int kThreads = kWarpBX * kWarpBY;
  __shared__ __align__(16) float s_acc[kTile];
  __shared__ double s_m[kWarpBatch][9];
  __shared__ int s_skip[kWarpBatch];
  const int tid = threadIdx.y * kWarpBX + threadIdx.x;
  if (tid < 9 * p.n) {
    const int j = tid / 9, i = tid - 9 * j;
    s_m[j][i] = p.f[j].inv_ptr ? p.f[j].inv_ptr[i] : p.f[j].inv[i];
  }
  if (tid < p.n) s_skip[tid] = (p.f[tid].status_ptr && *p.f[tid].status_ptr != 0) ? 1 : 0;

  const int tx0 = blockIdx.x * kWarpBX;
  const int ty0 = blockIdx.y * kWarpTH;
  const int row_elems = min(kWarpBX, p.width - tx0) * C;        // valid floats in this tile row
  const size_t row_base = (size_t)tx0 * C;                       // float offset of the tile inside a row
  const bool vec_ok = (((size_t)p.width * C) % 4 == 0);          // then every float4 of the tile is all in or all out

  // accumulator tile: global -> shared (zeros in store mode and outside the image)
  if (vec_ok) {
    for (int i = tid; i < kTile / 4; i += kThreads) {
      const int r = i / (kRow / 4), q = i - r * (kRow / 4);
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!p.store && ty0 + r < p.height && q * 4 < row_elems)
        a = *reinterpret_cast<const float4*>(p.acc + (size_t)(ty0 + r) * p.width * C + row_base + q * 4);
      *reinterpret_cast<float4*>(&s_acc[i * 4]) = a;
    }
  } else {
    for (int i = tid; i < kTile; i += kThreads) {
      const int r = i / kRow, q = i - r * kRow;
      s_acc[i] = (!p.store && ty0 + r < p.height && q < row_elems) ? p.acc[(size_t)(ty0 + r) * p.width * C + row_base + q] : 0.f;
    }
  }
  __syncthreads();

  for (int j = 0; j < p.n; ++j) {
    if (s_skip[j]) continue;                                   // block-uniform
    warp_tile<C, PERSP>(p.f[j], s_m[j], s_acc, p.width, p.height, p.src_width, p.src_height);
  }
  __syncthreads();

  if (vec_ok) {
    for (int i = tid; i < kTile / 4; i += kThreads) {
      const int r = i / (kRow / 4), q = i - r * (kRow / 4);
      if (ty0 + r < p.height && q * 4 < row_elems)
        *reinterpret_cast<float4*>(p.acc + (size_t)(ty0 + r) * p.width * C + row_base + q * 4) = *reinterpret_cast<const float4*>(&s_acc[i * 4]);
    }
  } else {
    for (int i = tid; i < kTile; i += kThreads) {
      const int r = i / kRow, q = i - r * kRow;
      if (ty0 + r < p.height && q < row_elems) p.acc[(size_t)(ty0 + r) * p.width * C + row_base + q] = s_acc[i];
    }
  }
}

// ---- K4, second generation ------------------------------------------------------------------------------
// Same results, bit for bit (same tests), rebuilt around what the round-2 ncu capture of the kernel above showed
// (profiles/r2_ncu_raw_warp_accumulate.csv): 171 warp-instructions per pixel and the L1 data pipe at 85 % — 12 byte
// gathers + 12 conversions per pixel, the nine matrix entries re-read from shared memory for every pixel and 12
// local-memory (spill) accesses per pixel because of the 32-register cap that the occupancy needed, plus 23 FP64
// operations per pixel for the coordinates.
//   * taps as aligned 32-bit words: the two horizontal taps of a source row are 2*C contiguous bytes; three aligned
//     words cover them for any alignment, PRMT shifts them into place and splices each byte under the exponent of
//     2^23, and ONE fused multiply-add turns that into fl(byte * fl(1/255)) exactly (2^23 * k is exact, so
//     fma(2^23 + b, k, -2^23 k) rounds b*k once): 6 loads per pixel instead of 12, no conversion instruction;
//   * perspective coordinates by guarded f32: the displacement (u - x, v - y) of a projective map is a ratio of
//     per-column polynomials in y (as in the ECC kernel's FastPersp), evaluated in f32 from constants prepared in
//     f64.  Its error is bounded per column (see FastInv::band); a pixel whose 32*displacement lands closer than that
//     to a rounding boundary — about one in a thousand — recomputes in f64 with OpenCV's operation order, so the
//     quantised coordinates are those of cv2.warpPerspective for every pixel;
//   * the per-column constants and the interior test of every frame of the batch are computed ONCE per block (warp j
//     does frame j, lane = column) into shared memory, before the block's only barrier — the first build had every
//     thread redo them per frame: 135 of its 149 instructions per pixel-frame went there and into 64-bit addressing;
//   * the accumulator slice lives in registers for the whole launch, 64 registers per thread: nothing spills, and the
//     24 independent word loads of a thread's four pixels hide the L2 latency that the old kernel needed 64 warps per
//     SM for.
constexpr int kWarp2Rows = 4;
constexpr int kWarp2TH = kWarpBY * kWarp2Rows;
constexpr int kFastInvN = 8;               // floats per column and frame, see FastInv

__device__ __forceinline__ float2 fma2_rn_exact(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%6, %7};\n"
      " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0, %1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 add2_rn_exact(float2 a, float2 b) {
  float2 d;
  asm("{\n .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n add.rn.f32x2 rd, ra, rb;\n mov.b64 {%0, %1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
// PRMT in its default mode; selector nibbles stay below 8 here, so no masking is needed (__byte_perm adds one)
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
  unsigned d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// Per-column (destination x) constants of one frame's inverse map for the guarded f32 coordinates:
//   u - x = (alpha + beta y) / W(y),   v - y = (gamma + (delta - m7 y) y) / W(y),   W(y) = wc + m7 y,
// held pre-multiplied by 32 (exact), so that t = n * rcp(W) is the displacement in 1/32-px quanta.
struct FastInv {
  float a32, b32, g32, d32, m7_32, wc, m7, thr;     // thr = 0.5 - band
  // y_lo, y_hi: the rows of the tile.  band bounds |t_f32 - t_exact| for both coordinates:
  //   coefficient roundings + the fmas: <= 3 * 2^-24 * S with S = the sum of the absolute terms of the numerator;
  //   W: two coefficient roundings + one fma, rcp.approx.ftz: 2^-23, the product: 2^-24 — < 9 * 2^-24 relative to |t| <= S / W;
  // 12 * 2^-24 * S / Wmin covers both with margin; the floor 2^-18 covers the f64 side's own rounding (< 2^-30).
  // W outside [1/4, 4] (far from the affine-like maps this path is for) or displacements beyond the range of the
  // magic-constant rounding: thr < 0 sends every pixel of the column through the exact evaluation.
  __device__ __forceinline__ void init(const double* m, int x, float y_lo, float y_hi) {
    const double xd = (double)x;
    const double wcd = m[6] * xd + m[8];
    a32 = (float)(32.0 * (xd * (m[0] - wcd) + m[2]));
    b32 = (float)(32.0 * (m[1] - m[7] * xd));
    g32 = (float)(32.0 * (m[3] * xd + m[5]));
    d32 = (float)(32.0 * (m[4] - wcd));
    m7_32 = (float)(32.0 * m[7]);
    wc = (float)wcd;
    m7 = (float)m[7];
    const float su = fabsf(a32) + fabsf(b32) * y_hi;
    const float sv = fabsf(g32) + (fabsf(d32) + fabsf(m7_32) * y_hi) * y_hi;
    const float w_lo = fmaf(m7, y_lo, wc), w_hi = fmaf(m7, y_hi, wc);
    const float wmin = fminf(w_lo, w_hi) * 0.99f;
    const bool ok = wmin > 0.25f && fmaxf(w_lo, w_hi) < 4.0f && fmaxf(su, sv) < 1.0e6f;
    thr = ok ? 0.5f - fmaf(fmaxf(su, sv) / wmin, 12.0f / 16777216.0f, 1.0f / 262144.0f) : -1.0f;
  }
};

// The two adjacent taps of a source row are 2*C contiguous bytes that start `o` bytes into the aligned word at wp: three
// words cover them for any o in 0..3 (C == 4 with o == 0 needs two).  Loading and converting are separate steps so that
// a thread can have the words of all its pixels in flight before it touches the first.
struct TapWords { unsigned w0, w1, w2; };
template <int C>
__device__ __forceinline__ TapWords load_tap_words(const unsigned* wp, unsigned o) {
  TapWords t;
  t.w0 = __ldg(wp); t.w1 = __ldg(wp + 1);
  t.w2 = 0;
  if (C == 3 || o != 0) t.w2 = __ldg(wp + 2);
  return t;
}
// s0[c], s1[c] = fl(byte * fl(1/255)) of the two taps: PRMT shifts the bytes into place and splices each under the
// exponent of 2^23; fma(2^23 + b, k, -2^23 k) rounds b*k once because 2^23 k is exact
template <int C>
__device__ __forceinline__ void convert_tap_pair(const TapWords& t, unsigned o, float (&s0)[C], float (&s1)[C]) {
  const unsigned sel = 0x3210u + 0x1111u * o;
  const unsigned A = prmt(t.w0, t.w1, sel), B = prmt(t.w1, t.w2, sel);      // bytes 0..3 and 4..7 of the pair
  constexpr unsigned kMagic = 0x4B000000u;                                   // 2^23
  const float k255 = (float)(1.0 / 255.0);
  const float kc = -8388608.0f * k255;                                       // exact
  unsigned b[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { b[i] = prmt(A, kMagic, 0x7540u + i); b[4 + i] = prmt(B, kMagic, 0x7540u + i); }
  float v[8];
#pragma unroll
  for (int i = 0; i + 1 < 2 * C; i += 2) {
    const float2 r = fma2_rn_exact(make_float2(__uint_as_float(b[i]), __uint_as_float(b[i + 1])), make_float2(k255, k255),
                                   make_float2(kc, kc));
    v[i] = r.x; v[i + 1] = r.y;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) { s0[c] = v[c]; s1[c] = v[C + c]; }
}

// value = s00*w00 + s01*w01 + s10*w10 + s11*w11, left to right in f32, on converted taps (see blend_taps).  wa = (w00, w01),
// wb = (w10, w11).  Channel pairs multiply by a broadcast weight; the odd channel's four products ride two packed
// multiplies across the taps.  Every half of a packed multiply is an IEEE round-to-nearest product, every add a scalar
// __fadd_rn: the values are those of the scalar sequence.
template <int C>
__device__ __forceinline__ void blend_vals(const float (&s00)[C], const float (&s01)[C], const float (&s10)[C],
                                           const float (&s11)[C], float2 wa, float2 wb, float (&out)[C]) {
  constexpr int P = C / 2;
#pragma unroll
  for (int q = 0; q < P; ++q) {
    const float2 p00 = mul2_rn_exact(make_float2(s00[2 * q], s00[2 * q + 1]), make_float2(wa.x, wa.x));
    const float2 p01 = mul2_rn_exact(make_float2(s01[2 * q], s01[2 * q + 1]), make_float2(wa.y, wa.y));
    const float2 p10 = mul2_rn_exact(make_float2(s10[2 * q], s10[2 * q + 1]), make_float2(wb.x, wb.x));
    const float2 p11 = mul2_rn_exact(make_float2(s11[2 * q], s11[2 * q + 1]), make_float2(wb.y, wb.y));
    out[2 * q] = __fadd_rn(__fadd_rn(__fadd_rn(p00.x, p01.x), p10.x), p11.x);
    out[2 * q + 1] = __fadd_rn(__fadd_rn(__fadd_rn(p00.y, p01.y), p10.y), p11.y);
  }
  if (C & 1) {
    constexpr int c = C - 1;
    const float2 pa = mul2_rn_exact(make_float2(s00[c], s01[c]), wa);       // (s00 w00, s01 w01)
    const float2 pb = mul2_rn_exact(make_float2(s10[c], s11[c]), wb);       // (s10 w10, s11 w11)
    out[c] = __fadd_rn(__fadd_rn(__fadd_rn(pa.x, pa.y), pb.x), pb.y);
  }
}

// Interior test of one frame for the block's tile, evaluated by one warp (lanes 0-3 take one corner each): the four
// corners land inside the source with w > 0 (then every pixel of the tile does: a projective map with w > 0 sends the
// rectangle into the hull of its corner images; the 1/16 px slack covers the 1/32-px quantisation), with the margins
// the word loads need — taps from column 1 on (an aligned word may start 3 bytes before the tap) and three columns short
// of the last one (it may end 6 bytes after the second tap).
template <bool PERSP>
__device__ __forceinline__ bool tile_is_interior(const double* mtx, int lane, int width, int height, int sw, int sh) {
  // a tile that sticks out of the destination on the right / at the bottom is judged on its part inside: the threads
  // outside compute on the clamped column / row and never store
  const int k = lane & 3;
  const double cxk = (double)min((int)(blockIdx.x * kWarpBX + ((k & 1) ? kWarpBX - 1 : 0)), width - 1);
  const double cyk = (double)min((int)(blockIdx.y * kWarp2TH + ((k & 2) ? kWarp2TH - 1 : 0)), height - 1);
  const double nu = mtx[0] * cxk + mtx[1] * cyk + mtx[2], nv = mtx[3] * cxk + mtx[4] * cyk + mtx[5];
  const double ww = PERSP ? mtx[6] * cxk + mtx[7] * cyk + mtx[8] : 1.0;
  const double ulo = 1.0625 * ww, vlo = 0.0625 * ww;
  const double uhi = ((double)(sw - 3) - 0.0625) * ww, vhi = ((double)(sh - 1) - 0.0625) * ww;
  return __all_sync(0xffffffffu, ww > 1e-9 && ww < 1e9 && nu >= ulo && nu < uhi && nv >= vlo && nv < vhi);
}

// One frame's contribution to the thread's kWarp2Rows pixels (column x, rows y_base + rr * kWarpBY), added to acc.
//   fic: this column's FastInv constants in shared memory (stride 32 floats), lean: the tile is interior for this frame
//   ALIGNED: the frame's base address and pitch are multiples of 4 (word index arithmetic in 32 bits)
template <int C, bool PERSP, bool ALIGNED>
__device__ __forceinline__ void warp_tile_v2(const WarpFrame& f, const double* __restrict__ mtx, const float* fic, bool lean,
                                             float (&acc)[kWarp2Rows][C], int width, int height, int sw, int sh,
                                             unsigned frac_magic) {
  const int x = blockIdx.x * kWarpBX + threadIdx.x;
  const int y_base = blockIdx.y * kWarp2TH + threadIdx.y;

  if (lean) {
    // threads beyond the right / bottom edge of the destination work on the clamped column / row (their pixels are never
    // stored): a partial tile runs this path like a full one
    const int xc = min(x, width - 1);
    float a32 = 0.f, b32 = 0.f, g32 = 0.f, d32 = 0.f, m7_32 = 0.f, wc = 0.f, m7 = 0.f, thr = 0.f;
    int adelta = 0, bdelta = 0;
    if (PERSP) {
      a32 = fic[0]; b32 = fic[32]; g32 = fic[64]; d32 = fic[96]; m7_32 = fic[128]; wc = fic[160]; m7 = fic[192]; thr = fic[224];
    } else {
      adelta = __float_as_int(fic[0]);
      bdelta = __float_as_int(fic[32]);
    }
    const unsigned pitch_w = (unsigned)(f.src_pitch >> 2);
    const float2 magic2 = make_float2(12582912.0f, 12582912.0f), neg_magic2 = make_float2(-12582912.0f, -12582912.0f);
    // phase 1: quantised source coordinates of the thread's pixels (the only phase with a data-dependent branch)
    int xq[kWarp2Rows], yq[kWarp2Rows];
#pragma unroll
    for (int rr = 0; rr < kWarp2Rows; ++rr) {
      const int y = min(y_base + rr * kWarpBY, height - 1);
      if (PERSP) {
        const float yf = (float)y;
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(m7, yf, wc)));
        // (tu, tv) = numerators * r: the displacement in 1/32-px quanta; q = rint(t) through the 1.5 * 2^23 magic constant,
        // res = t - q (both coordinates per packed instruction)
        const float2 n2 = make_float2(fmaf(b32, yf, a32), fmaf(fmaf(-m7_32, yf, d32), yf, g32));
        const float2 r2 = make_float2(r, r);
        const float2 q2 = fma2_rn_exact(n2, r2, magic2);
        const float2 i2 = add2_rn_exact(q2, neg_magic2);
        const float2 res = fma2_rn_exact(n2, r2, make_float2(-i2.x, -i2.y));
        xq[rr] = (xc << kInterBits) + (__float_as_int(q2.x) - 0x4B400000);
        yq[rr] = (y << kInterBits) + (__float_as_int(q2.y) - 0x4B400000);
        if (!(fabsf(res.x) <= thr && fabsf(res.y) <= thr)) {
          // too close to a rounding boundary for the f32 evaluation, outside its range (thr < 0) or not a number (a
          // comparison with NaN is false, so the negated form sends it here too): OpenCV's own f64 sequence
          const int xb = width >= 64 ? (xc & ~63) : 0;
          const double xbd = (double)xb, x1 = (double)(xc - xb), yd = (double)y;
          const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(mtx[0], xbd), __dmul_rn(mtx[1], yd)), mtx[2]);
          const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(mtx[3], xbd), __dmul_rn(mtx[4], yd)), mtx[5]);
          const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(mtx[6], xbd), __dmul_rn(mtx[7], yd)), mtx[8]);
          const double W32 = __dmul_rn(rcp_rn_normal(__dadd_rn(W0, __dmul_rn(mtx[6], x1))), (double)kInterTab);
          xq[rr] = rint_magic(__dmul_rn(__dadd_rn(X0, __dmul_rn(mtx[0], x1)), W32));
          yq[rr] = rint_magic(__dmul_rn(__dadd_rn(Y0, __dmul_rn(mtx[3], x1)), W32));
        }
      } else {
        const double yd = (double)y;
        const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(mtx[1], yd), mtx[2]), kAbScale)) + 16;
        const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(mtx[4], yd), mtx[5]), kAbScale)) + 16;
        xq[rr] = (int)((unsigned)X0 + (unsigned)adelta) >> (kAbBits - kInterBits);
        yq[rr] = (int)((unsigned)Y0 + (unsigned)bdelta) >> (kAbBits - kInterBits);
      }
    }
    // phase 2: all tap words of all pixels in flight (6 loads per pixel)
    TapWords t0[kWarp2Rows], t1[kWarp2Rows];
    unsigned o0[kWarp2Rows], o1[kWarp2Rows];
#pragma unroll
    for (int rr = 0; rr < kWarp2Rows; ++rr) {
      if (ALIGNED) {
        // interior: both integer positions are non-negative and the frame is < 4 GiB, so 32-bit word indices do
        const unsigned bx = (unsigned)(xq[rr] >> kInterBits) * C;
        const unsigned wi = (unsigned)(yq[rr] >> kInterBits) * pitch_w + (bx >> 2);
        const unsigned* base = reinterpret_cast<const unsigned*>(f.src);
        o0[rr] = o1[rr] = bx & 3u;
        t0[rr] = load_tap_words<C>(base + wi, o0[rr]);
        t1[rr] = load_tap_words<C>(base + (wi + pitch_w), o1[rr]);
      } else {
        const uint8_t* r0 = f.src + (ptrdiff_t)(yq[rr] >> kInterBits) * (ptrdiff_t)f.src_pitch + (xq[rr] >> kInterBits) * C;
        const uintptr_t u0 = reinterpret_cast<uintptr_t>(r0), u1 = u0 + f.src_pitch;
        o0[rr] = (unsigned)u0 & 3u; o1[rr] = (unsigned)u1 & 3u;
        t0[rr] = load_tap_words<C>(reinterpret_cast<const unsigned*>(u0 & ~(uintptr_t)3), o0[rr]);
        t1[rr] = load_tap_words<C>(reinterpret_cast<const unsigned*>(u1 & ~(uintptr_t)3), o1[rr]);
      }
    }
    // phase 3: weights, conversion, blend, accumulate
#pragma unroll
    for (int rr = 0; rr < kWarp2Rows; ++rr) {
      // fractions k/32 spliced under the 1.5 * 2^23 exponent: (2^23 * 1.5 + k) / 32 - 2^23 * 1.5 / 32 is exact.
      // axy = (ax, ay); weights as OpenCV forms them: w00 = (1-ay)(1-ax), w01 = (1-ay) ax, w10 = ay (1-ax), w11 = ay ax
      const float2 axy = fma2_rn_exact(make_float2(__uint_as_float(((unsigned)xq[rr] & (kInterTab - 1)) | frac_magic),
                                                   __uint_as_float(((unsigned)yq[rr] & (kInterTab - 1)) | frac_magic)),
                                       make_float2(1.f / kInterTab, 1.f / kInterTab),
                                       make_float2(-12582912.0f / kInterTab, -12582912.0f / kInterTab));
      const float2 one_m = add2_rn_exact(make_float2(1.f, 1.f), make_float2(-axy.x, -axy.y));      // (1 - ax, 1 - ay)
      const float2 xs = make_float2(one_m.x, axy.x);
      const float2 wa = mul2_rn_exact(make_float2(one_m.y, one_m.y), xs);                            // (w00, w01)
      const float2 wb = mul2_rn_exact(make_float2(axy.y, axy.y), xs);                                // (w10, w11)
      float s00[C], s01[C], s10[C], s11[C], v[C];
      convert_tap_pair<C>(t0[rr], o0[rr], s00, s01);
      convert_tap_pair<C>(t1[rr], o1[rr], s10, s11);
      blend_vals<C>(s00, s01, s10, s11, wa, wb, v);
#pragma unroll
      for (int c = 0; c < C; ++c) acc[rr][c] = __fadd_rn(acc[rr][c], v[c]);
    }
    return;
  }

  // rim tiles and maps that leave the source: the general per-pixel path (every tap by its own rule)
  double cx0 = 0, cx1 = 0, cy0 = 0, cy1 = 0, cw0 = 0, cw1 = 0;
  int adelta = 0, bdelta = 0;
  if (PERSP) {
    const int xb = width >= 64 ? (x & ~63) : 0;
    const double xbd = (double)xb, x1 = (double)(x - xb);
    cx0 = __dmul_rn(mtx[0], xbd); cx1 = __dmul_rn(mtx[0], x1);
    cy0 = __dmul_rn(mtx[3], xbd); cy1 = __dmul_rn(mtx[3], x1);
    cw0 = __dmul_rn(mtx[6], xbd); cw1 = __dmul_rn(mtx[6], x1);
  } else {
    const double xd = (double)x;
    adelta = __double2int_rn(__dmul_rn(__dmul_rn(mtx[0], xd), kAbScale));
    bdelta = __double2int_rn(__dmul_rn(__dmul_rn(mtx[3], xd), kAbScale));
  }
#pragma unroll 1
  for (int rr = 0; rr < kWarp2Rows; ++rr) {
    const int y = y_base + rr * kWarpBY;
    if (x < width && y < height) {
      float v[C];
      general_pixel<C, PERSP>(f, mtx, cx0, cx1, cy0, cy1, cw0, cw1, adelta, bdelta, y, sw, sh, v);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        // dynamic rr: select the row's slot without indexing the register array
#pragma unroll
        for (int q = 0; q < kWarp2Rows; ++q) if (q == rr) acc[q][c] = __fadd_rn(acc[q][c], v[c]);
      }
    }
  }
}

template <int C, bool PERSP, bool ALIGNED>
__global__ void __launch_bounds__(kWarpBX * kWarpBY, 4) warp_accumulate_v2_kernel(const __grid_constant__ WarpAccParams p) {
  __shared__ double s_m[kWarpBatch][9];
  __shared__ float s_fi[kWarpBatch][kFastInvN][kWarpBX];
  __shared__ int s_flag[kWarpBatch];              // bit 0: skip the frame (failed ECC), bit 1: interior tile
  const int tid = threadIdx.y * kWarpBX + threadIdx.x;
  if (tid < 9 * p.n) {
    const int j = tid / 9, i = tid - 9 * j;
    s_m[j][i] = p.f[j].inv_ptr ? p.f[j].inv_ptr[i] : p.f[j].inv[i];
  }
  if ((int)threadIdx.y < p.n) {
    // warp j prepares frame j: interior test of the tile, per-column constants (lane = column)
    const int j = threadIdx.y, lane = threadIdx.x;
    double m[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = p.f[j].inv_ptr ? p.f[j].inv_ptr[i] : p.f[j].inv[i];
    const bool lean = tile_is_interior<PERSP>(m, lane, p.width, p.height, p.src_width, p.src_height);
    const int x = min((int)(blockIdx.x * kWarpBX + lane), p.width - 1);      // columns beyond the edge: see warp_tile_v2
    if (PERSP) {
      FastInv fi;
      fi.init(m, x, (float)(blockIdx.y * kWarp2TH), (float)min((int)(blockIdx.y * kWarp2TH + kWarp2TH - 1), p.height - 1));
      s_fi[j][0][lane] = fi.a32; s_fi[j][1][lane] = fi.b32; s_fi[j][2][lane] = fi.g32; s_fi[j][3][lane] = fi.d32;
      s_fi[j][4][lane] = fi.m7_32; s_fi[j][5][lane] = fi.wc; s_fi[j][6][lane] = fi.m7; s_fi[j][7][lane] = fi.thr;
    } else {
      const double xd = (double)x;
      s_fi[j][0][lane] = __int_as_float(__double2int_rn(__dmul_rn(__dmul_rn(m[0], xd), kAbScale)));
      s_fi[j][1][lane] = __int_as_float(__double2int_rn(__dmul_rn(__dmul_rn(m[3], xd), kAbScale)));
    }
    if (lane == 0) s_flag[j] = ((p.f[j].status_ptr && *p.f[j].status_ptr != 0) ? 1 : 0) | (lean ? 2 : 0);
  }

  const int x = blockIdx.x * kWarpBX + threadIdx.x;
  const int y_base = blockIdx.y * kWarp2TH + threadIdx.y;
  const size_t row_floats = (size_t)p.width * C;
  const bool vec4 = C == 4 && ((reinterpret_cast<uintptr_t>(p.acc) & 15) == 0);
  float acc[kWarp2Rows][C];
#pragma unroll
  for (int rr = 0; rr < kWarp2Rows; ++rr) {
    const int y = y_base + rr * kWarpBY;
    const bool in = x < p.width && y < p.height && !p.store;
    const float* src = p.acc + (size_t)y * row_floats + (size_t)x * C;
    if (C == 4 && vec4) {
      const float4 a = in ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
      acc[rr][0] = a.x; acc[rr][1] = a.y; acc[rr][2] = a.z; acc[rr][C - 1] = a.w;
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) acc[rr][c] = in ? src[c] : 0.f;
    }
  }
  __syncthreads();

  for (int j = 0; j < p.n; ++j) {
    const int flag = s_flag[j];                                // block-uniform
    if (flag & 1) continue;
    warp_tile_v2<C, PERSP, ALIGNED>(p.f[j], s_m[j], &s_fi[j][0][threadIdx.x], (flag & 2) != 0, acc, p.width, p.height,
                                    p.src_width, p.src_height, p.frac_magic);
  }

#pragma unroll
  for (int rr = 0; rr < kWarp2Rows; ++rr) {
    const int y = y_base + rr * kWarpBY;
    if (x < p.width && y < p.height) {
      float* dst = p.acc + (size_t)y * row_floats + (size_t)x * C;
      if (C == 4 && vec4) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[rr][0], acc[rr][1], acc[rr][2], acc[rr][C - 1]);
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) dst[c] = acc[rr][c];
      }
    }
  }
}

// frame 0 enters the stack unwarped: acc = u8 * (1/255)        (/root/reference/src/lib.rs:752-754)
// four bytes in, one 128-bit store out per thread when the row layout allows it
__global__ void seed_accumulator_kernel(const uint8_t* src, size_t src_pitch, float* acc, int row_elems,
                                        int height, int vec_ok) {
  const int y = blockIdx.y;
  if (y >= height) return;
  const float k255 = (float)(1.0 / 255.0);
  const uint8_t* row = src + (size_t)y * src_pitch;
  float* out = acc + (size_t)y * row_elems;
  if (vec_ok) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= row_elems) return;
    const uchar4 b = *reinterpret_cast<const uchar4*>(row + i);
    float4 o;
    o.x = __fmul_rn((float)b.x, k255); o.y = __fmul_rn((float)b.y, k255);
    o.z = __fmul_rn((float)b.z, k255); o.w = __fmul_rn((float)b.w, k255);
    *reinterpret_cast<float4*>(out + i) = o;
  } else {
    for (int k = 0; k < 4; ++k) {
      const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4 + k;
      if (i < row_elems) out[i] = __fmul_rn((float)__ldg(row + i), k255);
    }
  }
}

struct LaneSumParams {
  const float* lanes[16];
  int n_lanes;
  float* out;
  size_t n;         // floats
  float scale;      // float(1/divisor), or 1
  int apply_scale;
};

// out[i] = ((l0[i] + l1[i]) + l2[i] ...) [* scale]; n_lanes may be 0 (zeros).  `out` may alias lanes[0].
__global__ void lane_sum_scale_kernel(const LaneSumParams p) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = p.n / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.n_lanes > 0) s = reinterpret_cast<const float4*>(p.lanes[0])[i];
    for (int l = 1; l < p.n_lanes; ++l) {
      const float4 t = reinterpret_cast<const float4*>(p.lanes[l])[i];
      s.x = __fadd_rn(s.x, t.x); s.y = __fadd_rn(s.y, t.y); s.z = __fadd_rn(s.z, t.z); s.w = __fadd_rn(s.w, t.w);
    }
    if (p.apply_scale) { s.x = __fmul_rn(s.x, p.scale); s.y = __fmul_rn(s.y, p.scale); s.z = __fmul_rn(s.z, p.scale); s.w = __fmul_rn(s.w, p.scale); }
    reinterpret_cast<float4*>(p.out)[i] = s;
  }
  // tail
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
    float s = p.n_lanes > 0 ? p.lanes[0][i] : 0.f;
    for (int l = 1; l < p.n_lanes; ++l) s = __fadd_rn(s, p.lanes[l][i]);
    if (p.apply_scale) s = __fmul_rn(s, p.scale);
    p.out[i] = s;
  }
}

}  // namespace stk
