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

struct WarpAccParams {
  const uint8_t* src;     // u8 interleaved
  size_t src_pitch;
  float* acc;             // f32 interleaved, width*channels floats per row, dense
  int width, height;      // destination size (== accumulator size)
  int src_width, src_height;
  const double* inv_ptr;  // device pointer to the inverse map (ECC path), or null -> inv[]
  const int* status_ptr;  // device pointer to the frame's ECC status (skip when != 0), or null
  double inv[9];
  float border[4];
  int border_mode;        // cv::BorderTypes: 0 CONSTANT, 1 REPLICATE, 2 REFLECT, 3 WRAP, 4 REFLECT_101
  int store;              // 1: acc = v (first frame on this lane), 0: acc += v
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

// 32 x 32 destination tile per 256-thread block, four rows per thread (so the per-thread column terms of the
// coordinate transform, the parameter loads and the index math are paid once per four pixels).  The C
// interpolated values of the tile are staged in shared memory and the accumulator read-modify-write is done
// as coalesced 128-bit accesses (a tile row is 32*C contiguous floats).  ncu (profiles/) showed the first
// versions to be issue-bound (390, then 234 instructions per pixel), so the per-pixel path is kept lean: the
// inverse map is staged once per block, the 64-px column block start is a mask, the reciprocal replaces
// the f64 divide (32/W == 32 * (1/W) exactly: a power-of-two scaling commutes with rounding),
// __double2int_rn supplies the saturation, and pixels whose four taps are inside the source skip all
// border selects.
template <int C, bool PERSP>
__global__ void __launch_bounds__(kWarpBX * kWarpBY, 8) warp_accumulate_kernel(const WarpAccParams p) {
  __shared__ __align__(16) float s_val[kWarpTH][kWarpBX * C];
  __shared__ double s_m[9];
  if (p.status_ptr && *p.status_ptr != 0) return;
  const int tid = threadIdx.y * kWarpBX + threadIdx.x;
  if (tid < 9) {
    double mv = p.inv[0];
#pragma unroll
    for (int i = 1; i < 9; ++i) if (tid == i) mv = p.inv[i];
    if (p.inv_ptr) mv = p.inv_ptr[tid];
    s_m[tid] = mv;
  }
  __syncthreads();
  const int x = blockIdx.x * kWarpBX + threadIdx.x;
  const int y_base = blockIdx.y * kWarpTH + threadIdx.y;
  const float k255 = (float)(1.0 / 255.0);
  const int sw = p.src_width, sh = p.src_height;

  // per-thread (column) terms
  double cx0 = 0, cx1 = 0, cy0 = 0, cy1 = 0, cw0 = 0, cw1 = 0;
  int adelta = 0, bdelta = 0;
  if (PERSP) {
    // WarpPerspectiveInvoker: X0/Y0/W0 at the start of the 64-px column block, then + M*x1
    const int xb = p.width >= 64 ? (x & ~63) : 0;
    const double xbd = (double)xb, x1 = (double)(x - xb);
    cx0 = __dmul_rn(s_m[0], xbd); cx1 = __dmul_rn(s_m[0], x1);
    cy0 = __dmul_rn(s_m[3], xbd); cy1 = __dmul_rn(s_m[3], x1);
    cw0 = __dmul_rn(s_m[6], xbd); cw1 = __dmul_rn(s_m[6], x1);
  } else {
    const double xd = (double)x;
    adelta = __double2int_rn(__dmul_rn(__dmul_rn(s_m[0], xd), kAbScale));
    bdelta = __double2int_rn(__dmul_rn(__dmul_rn(s_m[3], xd), kAbScale));
  }

  // Interior tiles (the bulk of a frame): when the four corners of the tile land inside
  // [1/16, sw-1-1/16] x [1/16, sh-1-1/16] with w in (1e-9, 1e9), every tap of every pixel of the tile is inside
  // the source (a projective map with w > 0 sends the rectangle into the hull of its corner images; the slack
  // covers the 1/32-px quantisation), so the pixel loop needs no bounds test, clamp or border select, the
  // reciprocal no special cases and cvRound no saturation.  Decided per warp (lanes 0-3 take one corner each,
  // inequalities multiplied through by w: no division), no extra barrier.
  bool lean = blockIdx.x * kWarpBX + kWarpBX <= p.width && blockIdx.y * kWarpTH + kWarpTH <= p.height;
  {
    const int k = threadIdx.x & 3;
    const double cxk = (double)(blockIdx.x * kWarpBX + ((k & 1) ? kWarpBX - 1 : 0));
    const double cyk = (double)(blockIdx.y * kWarpTH + ((k & 2) ? kWarpTH - 1 : 0));
    const double nu = s_m[0] * cxk + s_m[1] * cyk + s_m[2], nv = s_m[3] * cxk + s_m[4] * cyk + s_m[5];
    const double ww = PERSP ? s_m[6] * cxk + s_m[7] * cyk + s_m[8] : 1.0;
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
        const double X0 = __dadd_rn(__dadd_rn(cx0, __dmul_rn(s_m[1], yd)), s_m[2]);
        const double Y0 = __dadd_rn(__dadd_rn(cy0, __dmul_rn(s_m[4], yd)), s_m[5]);
        const double W0 = __dadd_rn(__dadd_rn(cw0, __dmul_rn(s_m[7], yd)), s_m[8]);
        const double W32 = __dmul_rn(rcp_rn_normal(__dadd_rn(W0, cw1)), (double)kInterTab);
        xq = rint_magic(__dmul_rn(__dadd_rn(X0, cx1), W32));
        yq = rint_magic(__dmul_rn(__dadd_rn(Y0, cy1), W32));
      } else {
        const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(s_m[1], yd), s_m[2]), kAbScale)) + 16;
        const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(s_m[4], yd), s_m[5]), kAbScale)) + 16;
        xq = (int)((unsigned)X0 + (unsigned)adelta) >> (kAbBits - kInterBits);
        yq = (int)((unsigned)Y0 + (unsigned)bdelta) >> (kAbBits - kInterBits);
      }
      const float ax = (float)(xq & (kInterTab - 1)) * (1.f / kInterTab);
      const float ay = (float)(yq & (kInterTab - 1)) * (1.f / kInterTab);
      const float w00 = __fmul_rn(1.f - ay, 1.f - ax), w01 = __fmul_rn(1.f - ay, ax);
      const float w10 = __fmul_rn(ay, 1.f - ax), w11 = __fmul_rn(ay, ax);
      const uint8_t* r0 = p.src + (ptrdiff_t)(yq >> kInterBits) * (ptrdiff_t)p.src_pitch + (xq >> kInterBits) * C;
      const uint8_t* r1 = r0 + p.src_pitch;
      unsigned t00[C], t01[C], t10[C], t11[C];
#pragma unroll
      for (int c = 0; c < C; ++c) { t00[c] = __ldg(r0 + c); t01[c] = __ldg(r0 + C + c); t10[c] = __ldg(r1 + c); t11[c] = __ldg(r1 + C + c); }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float s00 = __fmul_rn((float)t00[c], k255), s01 = __fmul_rn((float)t01[c], k255);
        const float s10 = __fmul_rn((float)t10[c], k255), s11 = __fmul_rn((float)t11[c], k255);
        s_val[ry][threadIdx.x * C + c] =
            __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s00, w00), __fmul_rn(s01, w01)), __fmul_rn(s10, w10)), __fmul_rn(s11, w11));
      }
    }
  } else {
#pragma unroll
  for (int rr = 0; rr < kWarpRows; ++rr) {
    const int y = y_base + rr * kWarpBY;
    float v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = 0.f;
    if (x < p.width && y < p.height) {
      const double yd = (double)y;
      int xq, yq;
      if (PERSP) {
        const double X0 = __dadd_rn(__dadd_rn(cx0, __dmul_rn(s_m[1], yd)), s_m[2]);
        const double Y0 = __dadd_rn(__dadd_rn(cy0, __dmul_rn(s_m[4], yd)), s_m[5]);
        const double W0 = __dadd_rn(__dadd_rn(cw0, __dmul_rn(s_m[7], yd)), s_m[8]);
        const double W = __dadd_rn(W0, cw1);
        const double W32 = (W != 0.0) ? __dmul_rn(__drcp_rn(W), (double)kInterTab) : 0.0;      // == INTER_TAB_SIZE / W
        // cvRound with saturation == clamp to [INT_MIN, INT_MAX] then round half to even
        xq = __double2int_rn(__dmul_rn(__dadd_rn(X0, cx1), W32));
        yq = __double2int_rn(__dmul_rn(__dadd_rn(Y0, cy1), W32));
      } else {
        const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(s_m[1], yd), s_m[2]), kAbScale)) + 16;
        const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(s_m[4], yd), s_m[5]), kAbScale)) + 16;
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
      const uint8_t* r0 = p.src + (ptrdiff_t)sy * (ptrdiff_t)p.src_pitch + (ptrdiff_t)sx * C;
      const uint8_t* r1 = r0 + p.src_pitch;
      if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
        // all four taps inside the source: the common case, no border logic
        unsigned t00[C], t01[C], t10[C], t11[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { t00[c] = __ldg(r0 + c); t01[c] = __ldg(r0 + C + c); t10[c] = __ldg(r1 + c); t11[c] = __ldg(r1 + C + c); }
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float s00 = __fmul_rn((float)t00[c], k255), s01 = __fmul_rn((float)t01[c], k255);
          const float s10 = __fmul_rn((float)t10[c], k255), s11 = __fmul_rn((float)t11[c], k255);
          v[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s00, w00), __fmul_rn(s01, w01)), __fmul_rn(s10, w10)),
                           __fmul_rn(s11, w11));
        }
      } else if (p.border_mode != 0) {
        // REPLICATE / REFLECT / WRAP / REFLECT_101: each tap's coordinates go through borderInterpolate on their own
        const int x0 = border_index(sx, sw, p.border_mode), x1 = border_index(sx + 1, sw, p.border_mode);
        const int y0 = border_index(sy, sh, p.border_mode), y1 = border_index(sy + 1, sh, p.border_mode);
        const uint8_t* q0 = p.src + (size_t)y0 * p.src_pitch;
        const uint8_t* q1 = p.src + (size_t)y1 * p.src_pitch;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float s00 = __fmul_rn((float)__ldg(q0 + x0 * C + c), k255), s01 = __fmul_rn((float)__ldg(q0 + x1 * C + c), k255);
          const float s10 = __fmul_rn((float)__ldg(q1 + x0 * C + c), k255), s11 = __fmul_rn((float)__ldg(q1 + x1 * C + c), k255);
          v[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s00, w00), __fmul_rn(s01, w01)), __fmul_rn(s10, w10)),
                           __fmul_rn(s11, w11));
        }
      } else if (sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = p.border[c];
      } else {
        const bool x0in = (unsigned)sx < (unsigned)sw, x1in = (unsigned)(sx + 1) < (unsigned)sw;
        const bool y0in = (unsigned)sy < (unsigned)sh, y1in = (unsigned)(sy + 1) < (unsigned)sh;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float s00 = (y0in && x0in) ? __fmul_rn((float)__ldg(r0 + c), k255) : p.border[c];
          const float s01 = (y0in && x1in) ? __fmul_rn((float)__ldg(r0 + C + c), k255) : p.border[c];
          const float s10 = (y1in && x0in) ? __fmul_rn((float)__ldg(r1 + c), k255) : p.border[c];
          const float s11 = (y1in && x1in) ? __fmul_rn((float)__ldg(r1 + C + c), k255) : p.border[c];
          v[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s00, w00), __fmul_rn(s01, w01)), __fmul_rn(s10, w10)),
                           __fmul_rn(s11, w11));
        }
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) s_val[threadIdx.y + rr * kWarpBY][threadIdx.x * C + c] = v[c];
  }
  }
  __syncthreads();

  // coalesced accumulate: tile row r holds floats [x_tile0*C, x_tile0*C + 32*C) of accumulator row y0 + r
  const int tx0 = blockIdx.x * kWarpBX;
  const int ty0 = blockIdx.y * kWarpTH;
  const int row_elems = min(kWarpBX, p.width - tx0) * C;        // valid floats in this tile row
  const size_t row_base = (size_t)tx0 * C;                       // float offset of the tile inside a row
  const bool vec_ok = (((size_t)p.width * C) % 4 == 0) && (row_base % 4 == 0);
  if (vec_ok) {
    constexpr int kVecPerRow = kWarpBX * C / 4;
    constexpr int kVecs = kWarpTH * kVecPerRow;
#pragma unroll
    for (int i0 = 0; i0 < kVecs; i0 += kWarpBX * kWarpBY) {
      const int i = i0 + tid;
      if (i >= kVecs) break;
      const int r = i / kVecPerRow, q = i - r * kVecPerRow;
      const int yy = ty0 + r;
      if (yy >= p.height || q * 4 >= row_elems) continue;
      float* a = p.acc + (size_t)yy * p.width * C + row_base + q * 4;
      const float4 nv = *reinterpret_cast<const float4*>(&s_val[r][q * 4]);
      if (q * 4 + 4 <= row_elems) {
        float4 o = nv;
        if (!p.store) {
          const float4 cur = *reinterpret_cast<const float4*>(a);
          o.x = __fadd_rn(cur.x, nv.x); o.y = __fadd_rn(cur.y, nv.y); o.z = __fadd_rn(cur.z, nv.z); o.w = __fadd_rn(cur.w, nv.w);
        }
        *reinterpret_cast<float4*>(a) = o;
      } else {
        const float e[4] = {nv.x, nv.y, nv.z, nv.w};
        for (int k = 0; k < 4 && q * 4 + k < row_elems; ++k) a[k] = p.store ? e[k] : __fadd_rn(a[k], e[k]);
      }
    }
  } else {
    for (int i = tid; i < kWarpTH * kWarpBX * C; i += kWarpBX * kWarpBY) {
      const int r = i / (kWarpBX * C), q = i - r * (kWarpBX * C);
      const int yy = ty0 + r;
      if (yy >= p.height || q >= row_elems) continue;
      float* a = p.acc + (size_t)yy * p.width * C + row_base + q;
      *a = p.store ? s_val[r][q] : __fadd_rn(*a, s_val[r][q]);
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
