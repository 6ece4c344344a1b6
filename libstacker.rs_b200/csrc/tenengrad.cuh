// K6 — Tenengrad sharpness: Sobel dx and dy (ksize 1/3/5/7, BORDER_REFLECT_101), gx^2 + gy^2, sum.
//
// Replaces the five OpenCV calls of sharpness_tenengrad (/root/reference/src/lib.rs:1111-1146):
// sobel x2 (CV_64F), multiply x2, add, mean.  On 8-bit input every intermediate is an integer, so the
// kernel works in int32 / uint64 and the host multiplies the exact uint64 sum by 1.0/N in f64 (as cv::mean does) — bit-identical
// to the CV_64F pipeline as long as the sum stays below 2^53 (always for ksize <= 5; for ksize 7 up to
// ~1.7e5 mean-square gradient per pixel at 24 MPx, far above natural images).  Algorithmic traffic: N
// bytes per frame (3N when the grey conversion is fused in).
#pragma once
#include "common.cuh"

namespace stk {

constexpr int kTenTW = 64, kTenTH = 32, kTenThreads = 256, kTenMaxR = 3;

struct TenengradParams {
  const uint8_t* src;     // frame 0 of the batch
  size_t frame_stride;    // bytes between frames
  size_t pitch;
  int width, height, channels;
  int radius;             // 1 for ksize 1 and 3, 2 for 5, 3 for 7
  int dtap[2 * kTenMaxR + 1];   // derivative taps
  int stap[2 * kTenMaxR + 1];   // smoothing taps
  unsigned long long* sums;     // one per frame
};

__global__ void __launch_bounds__(kTenThreads) tenengrad_kernel(const TenengradParams p) {
  constexpr int GW = kTenTW + 2 * kTenMaxR, GH = kTenTH + 2 * kTenMaxR;
  __shared__ short s_g[GH * GW];
  __shared__ int s_d[GH * kTenTW];     // horizontal derivative
  __shared__ int s_s[GH * kTenTW];     // horizontal smoothing
  __shared__ unsigned long long s_part[kTenThreads / 32];
  const int r = p.radius;
  const int gw = kTenTW + 2 * r, gh = kTenTH + 2 * r;
  const int x0 = blockIdx.x * kTenTW, y0 = blockIdx.y * kTenTH;
  const uint8_t* src = p.src + (size_t)blockIdx.z * p.frame_stride;
  const int tid = threadIdx.x;

  for (int i = tid; i < gw * gh; i += kTenThreads) {
    const int ty = i / gw, tx = i - ty * gw;
    const int sx = reflect101(x0 + tx - r, p.width);
    const int sy = reflect101(y0 + ty - r, p.height);
    const uint8_t* px = src + (size_t)sy * p.pitch + (size_t)sx * p.channels;
    s_g[i] = (short)(p.channels == 1 ? (int)px[0] : bgr2gray(px[0], px[1], px[2]));
  }
  __syncthreads();
  for (int i = tid; i < gh * kTenTW; i += kTenThreads) {
    const int ty = i / kTenTW, tx = i - ty * kTenTW;
    const short* row = s_g + ty * gw + tx;
    int d = 0, s = 0;
    for (int k = 0; k <= 2 * r; ++k) { const int v = row[k]; d += p.dtap[k] * v; s += p.stap[k] * v; }
    s_d[i] = d; s_s[i] = s;
  }
  __syncthreads();
  unsigned long long local = 0;
  for (int i = tid; i < kTenTH * kTenTW; i += kTenThreads) {
    const int ty = i / kTenTW, tx = i - ty * kTenTW;
    if (x0 + tx >= p.width || y0 + ty >= p.height) continue;
    int gx = 0, gy = 0;
    for (int k = 0; k <= 2 * r; ++k) {
      gx += p.stap[k] * s_d[(ty + k) * kTenTW + tx];
      gy += p.dtap[k] * s_s[(ty + k) * kTenTW + tx];
    }
    local += (unsigned long long)((long long)gx * gx) + (unsigned long long)((long long)gy * gy);
  }
  local = warp_sum(local);
  if ((tid & 31) == 0) s_part[tid >> 5] = local;
  __syncthreads();
  if (tid == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < kTenThreads / 32; ++w) t += s_part[w];
    atomicAdd(p.sums + blockIdx.z, t);       // integer: order-independent, deterministic
  }
}

}  // namespace stk
