// K0 — grey + INTER_AREA downscale for ecc_match_scaling_down: 8-bit BGR(A) frame -> cvtColor(BGR2GRAY)
// -> cv::resize(INTER_AREA) -> 8-bit grey plane at the ECC working size.
//
// Replaces utils::scale_image on the grey frame (/root/reference/src/utils.rs:186-214, called at
// /root/reference/src/lib.rs:891-892 and :921-922) together with the cvt_color of read_grey_and_f32
// (/root/reference/src/utils.rs:136-142).  The full-resolution grey plane is never written: algorithmic
// traffic is 3N bytes in + n bytes out (n = small size).  The colour resize the reference also computes
// (lib.rs:919-920) is only ever asked for its size, so it is not computed at all.
//
// An enlarging axis (landscape frame with height < scale_down_width < width) sends the call through OpenCV's 8-bit
// bilinear kernels in "area mode" (oracle/restate.py::resize_area_up_u8): integer arithmetic, bit-exact.
//
// Bit-exact restatement of OpenCV's two INTER_AREA down-scaling paths for 8-bit single-channel input
// (oracle/restate.py::resize_area_u8, pinned against cv2.resize):
//   * integer scale in both directions (resizeAreaFast_Invoker): integer block sum; 2x2 -> (s + 2) >> 2,
//     otherwise saturate(rint(float(s) * (1.f / area)));
//   * otherwise (ResizeArea_Invoker): f32, per source row  buf = sum_k S * alpha_k  in table order, then
//     sum = beta_0 * buf_0 ; sum += beta_j * buf_j, no FMA contraction, saturate(rint(sum)).  The
//     (index, weight) tables are computeResizeAreaTab's, built on the host in f64 and stored as f32.
#pragma once
#include "common.cuh"

namespace stk {

struct ResizeAreaParams {
  const uint8_t* src;     // interleaved u8, `channels` per pixel (1 = already grey)
  size_t src_pitch;
  uint8_t* dst;           // u8 grey, dst_pitch bytes per row
  int dst_pitch;
  int sw, sh, dw, dh, channels;
  int ix, iy;             // > 0: integer-scale fast path with this block size
  // generic path tables (device): per destination column / row the first source index, the tap count and
  // kx / ky weights (dense, zero padded)
  const int* xfirst; const int* xcount; const float* xw; int kx;
  const int* yfirst; const int* ycount; const float* yw; int ky;
  // an axis enlarges (landscape frame, height < scale_down_width < width): OpenCV emulates INTER_AREA with its
  // 8-bit bilinear kernels ("area mode"): xfirst / yfirst = first source index, xcoef / ycoef = the two 11-bit weights
  int up; const int* xcoef; const int* ycoef;
};

__device__ __forceinline__ int grey_at(const uint8_t* row, int x, int ch) {
  const uint8_t* px = row + (size_t)x * ch;
  return ch == 1 ? (int)__ldg(px) : bgr2gray(__ldg(px), __ldg(px + 1), __ldg(px + 2));
}

constexpr int kResizeBX = 32, kResizeBY = 8;

__global__ void __launch_bounds__(kResizeBX * kResizeBY) resize_area_grey_kernel(const ResizeAreaParams p) {
  const int dx = blockIdx.x * kResizeBX + threadIdx.x;
  const int dy = blockIdx.y * kResizeBY + threadIdx.y;
  if (dx >= p.dw || dy >= p.dh) return;
  const int ch = p.channels;
  int out;
  if (p.up) {
    // HResizeLinear<uchar,int,short> then VResizeLinear<uchar,int,short>:
    //   S = g[x0]*a0 + g[x1]*a1 ;  out = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
    const int x0 = p.xfirst[dx], x1 = min(x0 + 1, p.sw - 1), a0 = p.xcoef[2 * dx], a1 = p.xcoef[2 * dx + 1];
    const int y0 = p.yfirst[dy], y1 = min(y0 + 1, p.sh - 1), b0 = p.ycoef[2 * dy], b1 = p.ycoef[2 * dy + 1];
    const uint8_t* r0 = p.src + (size_t)y0 * p.src_pitch;
    const uint8_t* r1 = p.src + (size_t)y1 * p.src_pitch;
    const int s0 = grey_at(r0, x0, ch) * a0 + grey_at(r0, x1, ch) * a1;
    const int s1 = grey_at(r1, x0, ch) * a0 + grey_at(r1, x1, ch) * a1;
    out = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
  } else if (p.ix > 0) {
    int s = 0;
    for (int ky = 0; ky < p.iy; ++ky) {
      const uint8_t* row = p.src + (size_t)(dy * p.iy + ky) * p.src_pitch;
      for (int kx = 0; kx < p.ix; ++kx) s += grey_at(row, dx * p.ix + kx, ch);
    }
    if (p.ix == 2 && p.iy == 2) {
      out = (s + 2) >> 2;
    } else {
      const float scale = __fdiv_rn(1.f, (float)(p.ix * p.iy));
      out = __float2int_rn(__fmul_rn((float)s, scale));
    }
  } else {
    const int x0 = p.xfirst[dx], nx = p.xcount[dx];
    const int y0 = p.yfirst[dy], ny = p.ycount[dy];
    const float* xw = p.xw + (size_t)dx * p.kx;
    const float* yw = p.yw + (size_t)dy * p.ky;
    float sum = 0.f;
    for (int ky = 0; ky < ny; ++ky) {
      const uint8_t* row = p.src + (size_t)(y0 + ky) * p.src_pitch;
      float buf = 0.f;
      for (int kx = 0; kx < nx; ++kx) buf = __fadd_rn(buf, __fmul_rn((float)grey_at(row, x0 + kx, ch), __ldg(xw + kx)));
      const float t = __fmul_rn(__ldg(yw + ky), buf);
      sum = ky == 0 ? t : __fadd_rn(sum, t);
    }
    out = __float2int_rn(sum);
  }
  p.dst[(size_t)dy * p.dst_pitch + dx] = (uint8_t)min(max(out, 0), 255);
}

}  // namespace stk
