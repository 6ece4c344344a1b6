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
    float v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = 0.f;
    if (x < width && y < height) {
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
