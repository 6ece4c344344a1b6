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
//
// One block = one 128 x 32 output tile: grey tile with reflected halo -> shared memory, horizontal pass
// -> shared memory, vertical pass -> 128-bit coalesced stores.  Thread indexing is 2-D (lane = column,
// warp = row) so the inner loops carry no integer division.
#pragma once
#include "common.cuh"

namespace stk {

constexpr int kPrepTW = 128;      // output tile width
constexpr int kPrepTH = 32;       // output tile height
constexpr int kPrepThreads = 256;
constexpr int kMaxGaussRadius = 15;

struct PrepParams {
  const uint8_t* src;   // interleaved u8, `channels` per pixel
  size_t src_pitch;     // bytes
  float* dst;           // f32 plane
  int dst_pitch;        // floats (multiple of 4)
  int width, height, channels;
  int radius;           // k / 2
  float taps[2 * kMaxGaussRadius + 1];
};

__host__ __device__ inline int prep_grey_pitch(int r) { return kPrepTW + 2 * r + 1; }   // odd: no bank conflicts on column walks
inline size_t prep_smem_bytes(int r) {
  return (size_t)((kPrepTH + 2 * r) * prep_grey_pitch(r) + (kPrepTH + 2 * r) * kPrepTW) * sizeof(float);
}

// dynamic smem: grey[(TH+2r)][gp] floats + tmp[(TH+2r)][TW] floats
__global__ void __launch_bounds__(kPrepThreads) prep_grey_blur_kernel(const PrepParams p) {
  extern __shared__ __align__(16) float prep_smem[];
  const int r = p.radius;
  const int gw = kPrepTW + 2 * r, gh = kPrepTH + 2 * r, gp = prep_grey_pitch(r);
  float* tmp = prep_smem;                       // [gh][TW], 16-byte aligned rows
  float* grey = prep_smem + gh * kPrepTW;       // [gh][gp]
  const int x0 = blockIdx.x * kPrepTW, y0 = blockIdx.y * kPrepTH;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int ch = p.channels;

  // 1. grey tile with reflected halo
  for (int ty = wrp; ty < gh; ty += kPrepThreads / 32) {
    const int sy = reflect101(y0 + ty - r, p.height);
    const uint8_t* row = p.src + (size_t)sy * p.src_pitch;
    for (int tx = lane; tx < gw; tx += 32) {
      const int sx = reflect101(x0 + tx - r, p.width);
      const uint8_t* px = row + (size_t)sx * ch;
      grey[ty * gp + tx] = (float)bgr2gray(__ldg(px), __ldg(px + 1), __ldg(px + 2));
    }
  }
  __syncthreads();

  // 2. horizontal pass (symmetric taps: centre + pairs)
  for (int ty = wrp; ty < gh; ty += kPrepThreads / 32) {
#pragma unroll
    for (int q = 0; q < kPrepTW / 32; ++q) {
      const int tx = lane + 32 * q;
      const float* c = grey + ty * gp + tx + r;
      float s = p.taps[r] * c[0];
      for (int k = 1; k <= r; ++k) s += p.taps[r + k] * (c[-k] + c[k]);
      tmp[ty * kPrepTW + tx] = s;
    }
  }
  __syncthreads();

  // 3. vertical pass: each thread 4 consecutive columns, 128-bit smem loads and global stores
  const int cx = lane * 4;                       // 32 lanes x 4 = 128 columns
  const bool full4 = x0 + cx + 3 < p.width;
  for (int ty = wrp; ty < kPrepTH; ty += kPrepThreads / 32) {
    const int y = y0 + ty;
    if (y >= p.height || x0 + cx >= p.width) continue;
    const float* c = tmp + (ty + r) * kPrepTW + cx;
    const float4 v0 = *reinterpret_cast<const float4*>(c);
    const float t0 = p.taps[r];
    float4 s = make_float4(t0 * v0.x, t0 * v0.y, t0 * v0.z, t0 * v0.w);
    for (int k = 1; k <= r; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(c - k * kPrepTW);
      const float4 b = *reinterpret_cast<const float4*>(c + k * kPrepTW);
      const float t = p.taps[r + k];
      s.x += t * (a.x + b.x); s.y += t * (a.y + b.y); s.z += t * (a.z + b.z); s.w += t * (a.w + b.w);
    }
    float* o = p.dst + (size_t)y * p.dst_pitch + x0 + cx;
    if (full4) {
      *reinterpret_cast<float4*>(o) = s;
    } else {
      const float e[4] = {s.x, s.y, s.z, s.w};
      for (int k = 0; k < 4 && x0 + cx + k < p.width; ++k) o[k] = e[k];
    }
  }
}

}  // namespace stk
