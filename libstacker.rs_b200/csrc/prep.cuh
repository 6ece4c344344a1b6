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

// ---------------------------------------------------------------------------------------------------------------
// Streaming variant for gauss_filt_size 3 and 5 (the reference's default, examples/main.rs:122): no shared memory.
//
// With 8-bit input and the dyadic taps [1 2 1]/4, [1 4 6 4 1]/16 the blurred value is v / 16 resp. v / 256 with v
// an INTEGER below 2^16 (A3.1 of SURVEY §8), so the whole filter runs in exact integer arithmetic — two pixels per
// 32-bit register (16-bit halves never carry into each other: 256 * 255 < 2^16) — and the result is bit-identical to
// cv2.GaussianBlur on the f32 plane (every f32 intermediate of OpenCV's pipeline is exactly representable, so its
// summation order cannot matter).
//
// A warp owns a band of 120 output columns (lanes 1..30 store 4 columns each, lanes 0 and 31 only supply the
// horizontal halo) and streams down a strip of rows: per row a lane loads its 4 pixels (12 contiguous bytes for
// BGR: the warp reads 384 contiguous bytes), converts them to grey, exchanges two packed registers with its
// neighbours (the +-2 column halo), does the horizontal pass, pushes the row into a 2R+1-row register window and
// emits one finished row: 128-bit coalesced stores.  Loads of the next four rows are issued before the current four
// are processed.  The u16 -> f32 conversion is an exponent splice (0x47000000 | v is 32768 + v/256) and one FADD,
// not a conversion-unit I2F.
constexpr int kPrepBandCols = 120;
constexpr int kPrepStreamThreads = 128;

struct PrepStreamParams {
  const uint8_t* src;
  size_t src_pitch;
  float* dst;
  int dst_pitch;        // floats
  int width, height;
  int strip_rows;       // output rows per warp
  int n_bands, n_strips;
};

template <bool V>
struct PrepTag { static constexpr bool value = V; };

template <int C>
struct RawRow { uint32_t w[C == 3 ? 3 : (C == 4 ? 4 : 1)]; };

template <int C>
__device__ __forceinline__ void load_raw(RawRow<C>& r, const uint8_t* px) {
  if constexpr (C == 3) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(px);
    r.w[0] = __ldg(q); r.w[1] = __ldg(q + 1); r.w[2] = __ldg(q + 2);
  } else if constexpr (C == 4) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(px));
    r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
  } else {
    r.w[0] = __ldg(reinterpret_cast<const uint32_t*>(px));
  }
}

// four greys of a lane as two packed pairs: lo = g0 | g1 << 16, hi = g2 | g3 << 16.
// cvtColor's 15-bit weights (3735, 19235, 9798) are split into bytes so that one pixel costs two 4-way byte dot
// products on the raw (B, G, R, x) word instead of three extractions and three multiply-adds:
//   3735 = 14*256 + 151,  19235 = 75*256 + 35,  9798 = 38*256 + 70;  the 4th weight is 0 (next pixel's B / alpha).
template <int C>
__device__ __forceinline__ void grey_pairs(const RawRow<C>& r, uint32_t& lo, uint32_t& hi) {
  if constexpr (C == 1) {
    lo = __byte_perm(r.w[0], 0, 0x4140);     // bytes (b0, 0, b1, 0)
    hi = __byte_perm(r.w[0], 0, 0x4342);
  } else {
    uint32_t px[4];
    if constexpr (C == 3) {
      // 12 bytes: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
      px[0] = r.w[0];
      px[1] = __funnelshift_r(r.w[0], r.w[1], 24);
      px[2] = __funnelshift_r(r.w[1], r.w[2], 16);
      px[3] = r.w[2] >> 8;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) px[i] = r.w[i];
    }
    constexpr uint32_t kWLo = 151u | (35u << 8) | (70u << 16);
    constexpr uint32_t kWHi = 14u | (75u << 8) | (38u << 16);
    uint32_t y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = (__dp4a(px[i], kWLo, 16384u) + (__dp4a(px[i], kWHi, 0u) << 8)) >> 15;
    lo = y[0] | (y[1] << 16);
    hi = y[2] | (y[3] << 16);
  }
}

template <int C, int R>   // C channels in {1, 3, 4}; R = 1 (k = 3) or 2 (k = 5)
__global__ void __launch_bounds__(kPrepStreamThreads) prep_stream_kernel(const PrepStreamParams p) {
  const int warp = (blockIdx.x * kPrepStreamThreads + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= p.n_bands * p.n_strips) return;
  // consecutive warps walk along a strip of bands: neighbouring warps read neighbouring bytes of the same rows
  const int band = warp % p.n_bands, strip = warp / p.n_bands;
  const int x = band * kPrepBandCols + 4 * (lane - 1);       // first of this lane's 4 columns
  const int y0 = strip * p.strip_rows;
  const int y1 = min(y0 + p.strip_rows, p.height);
  const bool stores = x >= 0 && x < p.width && lane >= 1 && lane <= 30;   // width % 4 == 0: a lane is all in or all out
  const bool left_edge = x == 0, right_edge = x + 4 == p.width;
  const int n_in = (y1 - y0) + 2 * R;                        // input rows y0-R .. y1+R-1 (reflected at the border)
  // halo lanes outside the image read a clamped (valid) address; what they produce is never used: the lane next
  // to them applies the mirror rule instead
  const int xc = min(max(x, 0), p.width - 4);
  const uint8_t* src_lane = p.src + (size_t)xc * C;
  const uint32_t pitch = (uint32_t)p.src_pitch;              // host guarantees pitch * height < 2^32
  const int h = p.height;

  // magic exponent: LSB of the spliced mantissa weighs 1/256 (k = 5) resp. 1/16 (k = 3)
  constexpr uint32_t kSplice = R == 2 ? 0x47000000u : 0x49000000u;
  constexpr float kBias = R == 2 ? 32768.f : 524288.f;

  uint32_t wa[2 * R + 1], wb[2 * R + 1];     // window of horizontally filtered rows: pairs (h0,h1) and (h2,h3)
#pragma unroll
  for (int i = 0; i < 2 * R + 1; ++i) wa[i] = wb[i] = 0;
  const uint32_t m_left = left_edge ? 0xffffffffu : 0u, m_right = right_edge ? 0xffffffffu : 0u;
  float* out_lane = p.dst + x;               // only dereferenced when `stores`

  constexpr int U = R == 2 ? 5 : 6;          // rows per group: a multiple of the window length (no register moves)
  auto row_ptr = [&](int i) {                // input row i of the strip = image row y0 - R + i, BORDER_REFLECT_101
    int sy = y0 - R + i;                     // one reflection is enough (height > 2R)
    sy = sy < 0 ? -sy : sy;
    sy = sy >= h ? 2 * h - 2 - sy : sy;
    return src_lane + (uint32_t)sy * pitch;
  };
  // one input row: grey, halo exchange, horizontal pass, push into the window; kStore: emit output row `yo`
  auto row = [&](const RawRow<C>& raw, auto store_tag, int yo) {
    uint32_t lo, hi;
    grey_pairs<C>(raw, lo, hi);
    uint32_t lm = __shfl_up_sync(0xffffffffu, hi, 1);      // left neighbour's (g2, g3) = (g[-2], g[-1])
    uint32_t rp = __shfl_down_sync(0xffffffffu, lo, 1);    // right neighbour's (g0, g1) = (g[4], g[5])
    // BORDER_REFLECT_101: (g[-2], g[-1]) = (g2, g1) and (g[W], g[W+1]) = (g[W-2], g[W-3]) = (g2, g1) too
    const uint32_t mirror = __byte_perm(lo, hi, 0x3254);
    lm = (mirror & m_left) | (lm & ~m_left);
    rp = (mirror & m_right) | (rp & ~m_right);
    const uint32_t s1 = __byte_perm(lm, lo, 0x5432);       // (g[-1], g0)
    const uint32_t s2 = __byte_perm(lo, hi, 0x5432);       // (g1, g2)
    const uint32_t s3 = __byte_perm(hi, rp, 0x5432);       // (g3, g4)
    uint32_t ha, hb;
    if constexpr (R == 2) {
      ha = (lm + hi) + 4u * (s1 + s2) + 6u * lo;           // (h0, h1)
      hb = (lo + rp) + 4u * (s2 + s3) + 6u * hi;           // (h2, h3)
    } else {
      ha = s1 + 2u * lo + s2;                              // h0 = g[-1] + 2 g0 + g1 ; h1 = g0 + 2 g1 + g2
      hb = s2 + 2u * hi + s3;
    }
#pragma unroll
    for (int k = 0; k < 2 * R; ++k) { wa[k] = wa[k + 1]; wb[k] = wb[k + 1]; }
    wa[2 * R] = ha; wb[2 * R] = hb;
    if constexpr (decltype(store_tag)::value) {
      uint32_t va, vb;
      if constexpr (R == 2) {
        va = (wa[0] + wa[4]) + 4u * (wa[1] + wa[3]) + 6u * wa[2];
        vb = (wb[0] + wb[4]) + 4u * (wb[1] + wb[3]) + 6u * wb[2];
      } else {
        va = wa[0] + 2u * wa[1] + wa[2];
        vb = wb[0] + 2u * wb[1] + wb[2];
      }
      float4 o;
      o.x = __uint_as_float(__byte_perm(va, kSplice, 0x7410)) - kBias;
      o.y = __uint_as_float(__byte_perm(va, kSplice, 0x7432)) - kBias;
      o.z = __uint_as_float(__byte_perm(vb, kSplice, 0x7410)) - kBias;
      o.w = __uint_as_float(__byte_perm(vb, kSplice, 0x7432)) - kBias;
      if (stores) *reinterpret_cast<float4*>(out_lane + (size_t)yo * p.dst_pitch) = o;
    }
  };
  using Yes = PrepTag<true>;
  using No = PrepTag<false>;
  // a group of U input rows starting at input row `base`; whole groups carry no per-row tests
  auto fetch = [&](RawRow<C>* dst, int base) {
    if (base + U <= n_in) {
#pragma unroll
      for (int j = 0; j < U; ++j) load_raw<C>(dst[j], row_ptr(base + j));
    } else {
#pragma unroll
      for (int j = 0; j < U; ++j)
        if (base + j < n_in) load_raw<C>(dst[j], row_ptr(base + j));
    }
  };
  auto process = [&](const RawRow<C>* rows, int base) {
    if (base + U <= n_in) {
#pragma unroll
      for (int j = 0; j < U; ++j) row(rows[j], Yes(), y0 + base + j - 2 * R);
    } else {
#pragma unroll
      for (int j = 0; j < U; ++j)
        if (base + j < n_in) row(rows[j], Yes(), y0 + base + j - 2 * R);
    }
  };

  // the first 2R input rows only fill the window; then two groups of rows ping-pong: the loads of one are in
  // flight while the other is filtered
  RawRow<C> pro[2 * R], ga[U], gb[U];
#pragma unroll
  for (int j = 0; j < 2 * R; ++j) load_raw<C>(pro[j], row_ptr(j));
  fetch(ga, 2 * R);
#pragma unroll
  for (int j = 0; j < 2 * R; ++j) row(pro[j], No(), 0);
  for (int base = 2 * R; base < n_in; base += 2 * U) {
    fetch(gb, base + U);
    process(ga, base);
    fetch(ga, base + 2 * U);
    process(gb, base + U);
  }
}

// the streaming kernel's preconditions (everything else takes the tiled kernel)
inline bool prep_stream_ok(const void* src, size_t pitch, int width, int height, int channels, int radius) {
  if (radius != 1 && radius != 2) return false;
  if (channels != 1 && channels != 3 && channels != 4) return false;
  if (width % 4 != 0 || width < 8 || height < 2 * radius + 1) return false;
  if ((unsigned long long)pitch * (unsigned long long)height >= (1ull << 32)) return false;   // 32-bit row offsets
  const size_t align = channels == 4 ? 16 : 4;
  return ((uintptr_t)src % align) == 0 && pitch % align == 0;
}

}  // namespace stk
