// K1 — frame preparation: 8-bit BGR(A) -> grey (cvtColor BGR2GRAY, exact 15-bit fixed point) -> f32
// -> separable Gaussian blur (BORDER_REFLECT_101) -> f32 plane.
//
// Replaces, per frame, utils::read_grey_and_f32's cvt_color (/root/reference/src/utils.rs:136-142) and
// the `src.convertTo(f32); GaussianBlur(k x k, sigma 0)` at the top of OpenCV's findTransformECC that
// the reference reaches through /root/reference/src/lib.rs:769-777.  The grey and the un-blurred f32
// planes are never written to HBM: algorithmic traffic is 3N bytes in (u8 BGR) + 4N bytes out.
//
// Exactness: for k <= 9 OpenCV's taps are dyadic, so with 8-bit input every intermediate is exactly
// representable in f32 and the result is bit-identical to cv2.GaussianBlur regardless of summation
// order.  k >= 11 uses taps sampled in f64 on the host (same formula as getGaussianKernel).
#pragma once
#include "common.cuh"

namespace stk {

constexpr int kPrepTW = 64;       // output tile width
constexpr int kPrepTH = 32;       // output tile height
constexpr int kPrepThreads = 256;
constexpr int kMaxGaussRadius = 15;

struct PrepParams {
  const uint8_t* src;   // interleaved u8, `channels` per pixel
  size_t src_pitch;     // bytes
  float* dst;           // f32 plane
  int dst_pitch;        // floats
  int width, height, channels;
  int radius;           // k / 2
  float taps[2 * kMaxGaussRadius + 1];
};

// dynamic smem: grey[(TH+2r)][(TW+2r)] floats + tmp[(TH+2r)][TW] floats
__global__ void __launch_bounds__(kPrepThreads) prep_grey_blur_kernel(const PrepParams p) {
  extern __shared__ float smem[];
  const int r = p.radius;
  const int gw = kPrepTW + 2 * r, gh = kPrepTH + 2 * r;
  float* grey = smem;
  float* tmp = smem + gw * gh;
  const int x0 = blockIdx.x * kPrepTW, y0 = blockIdx.y * kPrepTH;
  const int tid = threadIdx.x;

  // 1. grey tile with reflected halo
  for (int i = tid; i < gw * gh; i += kPrepThreads) {
    const int ty = i / gw, tx = i - ty * gw;
    const int sx = reflect101(x0 + tx - r, p.width);
    const int sy = reflect101(y0 + ty - r, p.height);
    const uint8_t* px = p.src + (size_t)sy * p.src_pitch + (size_t)sx * p.channels;
    grey[i] = (float)bgr2gray(px[0], px[1], px[2]);
  }
  __syncthreads();

  // 2. horizontal pass (symmetric taps: centre + pairs)
  for (int i = tid; i < gh * kPrepTW; i += kPrepThreads) {
    const int ty = i / kPrepTW, tx = i - ty * kPrepTW;
    const float* row = grey + ty * gw + tx + r;
    float s = p.taps[r] * row[0];
    for (int k = 1; k <= r; ++k) s += p.taps[r + k] * (row[-k] + row[k]);
    tmp[i] = s;
  }
  __syncthreads();

  // 3. vertical pass
  for (int i = tid; i < kPrepTH * kPrepTW; i += kPrepThreads) {
    const int ty = i / kPrepTW, tx = i - ty * kPrepTW;
    const int x = x0 + tx, y = y0 + ty;
    if (x >= p.width || y >= p.height) continue;
    const float* col = tmp + (ty + r) * kPrepTW + tx;
    float s = p.taps[r] * col[0];
    for (int k = 1; k <= r; ++k) s += p.taps[r + k] * (col[-k * kPrepTW] + col[k * kPrepTW]);
    p.dst[(size_t)y * p.dst_pitch + x] = s;
  }
}

}  // namespace stk
