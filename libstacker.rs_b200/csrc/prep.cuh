// K1 — frame preparation: 8-bit BGR(A) (or an already-grey plane, channels == 1: the downscaled grey of
// ecc_match_scaling_down) -> grey (cvtColor BGR2GRAY, exact 15-bit fixed point) -> f32
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

// multiple of 4 floats: the horizontal pass reads its window as aligned 128-bit loads
__host__ __device__ inline int prep_grey_pitch(int r) { return (kPrepTW + 2 * r + 3) / 4 * 4 + 4; }
inline size_t prep_smem_bytes(int r) {
  return (size_t)((kPrepTH + 2 * r) * prep_grey_pitch(r) + (kPrepTH + 2 * r) * kPrepTW) * sizeof(float);
}

// dynamic smem: tmp[(TH+2r)][TW] floats + grey[(TH+2r)][gp] floats
// R > 0: compile-time radius (loops unrolled, taps in registers); R == 0: runtime radius p.radius.
template <int R>
__global__ void __launch_bounds__(kPrepThreads) prep_grey_blur_kernel(const PrepParams p) {
  extern __shared__ __align__(16) float prep_smem[];
  const int r = R > 0 ? R : p.radius;
  const int gw = kPrepTW + 2 * r, gh = kPrepTH + 2 * r, gp = prep_grey_pitch(r);
  float* tmp = prep_smem;                       // [gh][TW], 16-byte aligned rows
  float* grey = prep_smem + gh * kPrepTW;       // [gh][gp]
  const int x0 = blockIdx.x * kPrepTW, y0 = blockIdx.y * kPrepTH;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int ch = p.channels;
  constexpr int kWarps = kPrepThreads / 32;

  // 1. grey tile with halo; tiles whose halo stays inside the image skip the BORDER_REFLECT_101 index math
  const bool interior = x0 - r >= 0 && x0 + kPrepTW + r <= p.width && y0 - r >= 0 && y0 + kPrepTH + r <= p.height;
  // raw-byte staging: 16-byte aligned rows of the source tile, aliased onto `tmp` (dead until pass 2)
  const int raw_pitch = (gw * ch + 30) / 16 * 16;
  if (interior && raw_pitch <= kPrepTW * (int)sizeof(float)) {
    // The tile's source bytes come in as aligned 128-bit loads (an aligned chunk that holds one valid byte
    // lies in a mapped page, so the slop before/after a row is harmless) and the grey conversion reads them
    // from shared memory: ~1 global load per lane and row instead of 3 byte gathers per pixel.
    unsigned char* raw = reinterpret_cast<unsigned char*>(tmp);
    const uint8_t* base = p.src + (size_t)(y0 - r) * p.src_pitch + (size_t)(x0 - r) * ch;
    const int row_bytes = gw * ch;
    // every load of the warp's rows is in flight before the first store (one DRAM round trip, not one per row)
    constexpr int kRowsPerWarp = (kPrepTH + 2 * (R > 0 ? R : kMaxGaussRadius) + kWarps - 1) / kWarps;
    uint4 v[kRowsPerWarp];
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; ++i) {
      const int ty = wrp + i * kWarps;
      const uint8_t* row = base + (size_t)ty * p.src_pitch;
      const int off = (int)((uintptr_t)row & 15);
      if (ty < gh && lane * 16 < off + row_bytes) v[i] = __ldg(reinterpret_cast<const uint4*>(row - off) + lane);
    }
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; ++i) {
      const int ty = wrp + i * kWarps;
      const int off = (int)((uintptr_t)(base + (size_t)ty * p.src_pitch) & 15);
      if (ty < gh && lane * 16 < off + row_bytes) *reinterpret_cast<uint4*>(raw + ty * raw_pitch + lane * 16) = v[i];
    }
    __syncthreads();
    constexpr int kCols = (kPrepTW + 2 * (R > 0 ? R : kMaxGaussRadius) + 31) / 32;
    for (int ty = wrp; ty < gh; ty += kWarps) {
      const int off = (int)((uintptr_t)(base + (size_t)ty * p.src_pitch) & 15);
      const unsigned char* rrow = raw + ty * raw_pitch + off;
#pragma unroll
      for (int q = 0; q < kCols; ++q) {
        const int tx = lane + 32 * q;
        if (tx < gw) {
          const unsigned char* px = rrow + tx * ch;
          grey[ty * gp + tx] = ch == 1 ? (float)px[0] : (float)bgr2gray(px[0], px[1], px[2]);
        }
      }
    }
  } else if (interior) {
    const uint8_t* base = p.src + (size_t)(y0 - r) * p.src_pitch + (size_t)(x0 - r) * ch;
    // all byte loads of a row are issued before the first conversion (memory-level parallelism)
    constexpr int kCols = (kPrepTW + 2 * (R > 0 ? R : kMaxGaussRadius) + 31) / 32;
    for (int ty = wrp; ty < gh; ty += kWarps) {
      const uint8_t* row = base + (size_t)ty * p.src_pitch;
      unsigned b[kCols], g[kCols], rch[kCols];
#pragma unroll
      for (int q = 0; q < kCols; ++q) {
        const int tx = lane + 32 * q;
        if (tx < gw) {
          const uint8_t* px = row + tx * ch;
          b[q] = __ldg(px);
          if (ch != 1) { g[q] = __ldg(px + 1); rch[q] = __ldg(px + 2); }
        }
      }
#pragma unroll
      for (int q = 0; q < kCols; ++q) {
        const int tx = lane + 32 * q;
        if (tx < gw) grey[ty * gp + tx] = ch == 1 ? (float)b[q] : (float)bgr2gray(b[q], g[q], rch[q]);
      }
    }
  } else {
    for (int ty = wrp; ty < gh; ty += kWarps) {
      const int sy = reflect101(y0 + ty - r, p.height);
      const uint8_t* row = p.src + (size_t)sy * p.src_pitch;
      for (int tx = lane; tx < gw; tx += 32) {
        const int sx = reflect101(x0 + tx - r, p.width);
        const uint8_t* px = row + (size_t)sx * ch;
        grey[ty * gp + tx] = ch == 1 ? (float)__ldg(px) : (float)bgr2gray(__ldg(px), __ldg(px + 1), __ldg(px + 2));
      }
    }
  }
  __syncthreads();

  // taps (symmetric: centre + pairs) — registers when R is a template constant
  float tap[(R > 0 ? R : kMaxGaussRadius) + 1];
#pragma unroll
  for (int k = 0; k <= (R > 0 ? R : kMaxGaussRadius); ++k) tap[k] = (k <= r) ? p.taps[r + k] : 0.f;

  // 2. horizontal pass.  k = 5 (the reference's default): a thread produces 4 adjacent outputs from the 8 grey
  // values they span — two aligned 128-bit loads and one 128-bit store instead of 20 + 4 scalar accesses (the
  // kernel was bound by shared-memory wavefronts, profiles/r1_summary.md `pw_r1c`).
  if constexpr (R == 2) {
    for (int ty = wrp; ty < gh; ty += kWarps) {
      const float4 lo = *reinterpret_cast<const float4*>(grey + ty * gp + lane * 4);
      const float4 hi = *reinterpret_cast<const float4*>(grey + ty * gp + lane * 4 + 4);
      const float g[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float s = tap[0] * g[j + 2];
        s += tap[1] * (g[j + 1] + g[j + 3]);
        s += tap[2] * (g[j] + g[j + 4]);
        o[j] = s;
      }
      *reinterpret_cast<float4*>(tmp + ty * kPrepTW + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
  } else
  for (int ty = wrp; ty < gh; ty += kWarps) {
#pragma unroll
    for (int q = 0; q < kPrepTW / 32; ++q) {
      const int tx = lane + 32 * q;
      const float* c = grey + ty * gp + tx + r;
      float s = tap[0] * c[0];
      if (R > 0) {
#pragma unroll
        for (int k = 1; k <= R; ++k) s += tap[k] * (c[-k] + c[k]);
      } else {
        for (int k = 1; k <= r; ++k) s += p.taps[r + k] * (c[-k] + c[k]);
      }
      tmp[ty * kPrepTW + tx] = s;
    }
  }
  __syncthreads();

  // 3. vertical pass: each thread 4 consecutive columns, 128-bit smem loads and global stores
  const int cx = lane * 4;                       // 32 lanes x 4 = 128 columns
  const bool full4 = x0 + cx + 3 < p.width;
  if constexpr (R == 1 || R == 2) {
    // a warp owns 4 consecutive output rows: the 4 + 2R rows of `tmp` they need are loaded once (2 loads per
    // output row instead of 2R + 1)
    constexpr int kRows = kPrepTH / kWarps;                 // 4
    const int tyb = wrp * kRows;
    float4 w[kRows + 2 * R];
#pragma unroll
    for (int i = 0; i < kRows + 2 * R; ++i) w[i] = *reinterpret_cast<const float4*>(tmp + (tyb + i) * kPrepTW + cx);
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
      const int y = y0 + tyb + j;
      if (y >= p.height || x0 + cx >= p.width) continue;
      const float4 c0 = w[j + R];
      float4 s = make_float4(tap[0] * c0.x, tap[0] * c0.y, tap[0] * c0.z, tap[0] * c0.w);
#pragma unroll
      for (int k = 1; k <= R; ++k) {
        const float4 a = w[j + R - k], b = w[j + R + k];
        s.x += tap[k] * (a.x + b.x); s.y += tap[k] * (a.y + b.y); s.z += tap[k] * (a.z + b.z); s.w += tap[k] * (a.w + b.w);
      }
      float* o = p.dst + (size_t)y * p.dst_pitch + x0 + cx;
      if (full4) {
        *reinterpret_cast<float4*>(o) = s;
      } else {
        const float e[4] = {s.x, s.y, s.z, s.w};
        for (int k = 0; k < 4 && x0 + cx + k < p.width; ++k) o[k] = e[k];
      }
    }
  } else
  for (int ty = wrp; ty < kPrepTH; ty += kWarps) {
    const int y = y0 + ty;
    if (y >= p.height || x0 + cx >= p.width) continue;
    const float* c = tmp + (ty + r) * kPrepTW + cx;
    const float4 v0 = *reinterpret_cast<const float4*>(c);
    float4 s = make_float4(tap[0] * v0.x, tap[0] * v0.y, tap[0] * v0.z, tap[0] * v0.w);
    auto acc_pair = [&](int k, float t) {
      const float4 a = *reinterpret_cast<const float4*>(c - k * kPrepTW);
      const float4 b = *reinterpret_cast<const float4*>(c + k * kPrepTW);
      s.x += t * (a.x + b.x); s.y += t * (a.y + b.y); s.z += t * (a.z + b.z); s.w += t * (a.w + b.w);
    };
    if (R > 0) {
#pragma unroll
      for (int k = 1; k <= R; ++k) acc_pair(k, tap[k]);
    } else {
      for (int k = 1; k <= r; ++k) acc_pair(k, p.taps[r + k]);
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
