// Shared device helpers for the sm_100a align-and-stack kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace stk {

constexpr int kInterBits = 5;                 // OpenCV INTER_BITS: sub-pixel position quantised to 1/32 px
constexpr int kInterTab = 1 << kInterBits;    // INTER_TAB_SIZE
constexpr int kAbBits = 10;                   // warpAffine fixed point (AB_BITS)
constexpr double kAbScale = 1024.0;           // AB_SCALE

enum Motion : int { kTranslation = 0, kEuclidean = 1, kAffine = 2, kHomography = 3 };

// round-half-even double -> int32 through the 1.5*2^52 magic constant: one DADD on the FP64 pipe instead
// of a conversion-unit F2I.  Valid for |v| < 2^31 (callers range-check first).
__device__ __forceinline__ int rint_magic(double v) {
  return __double2loint(__dadd_rn(v, 6755399441055744.0));
}
// fused variant: rint(v * s)
__device__ __forceinline__ int rint_magic_scaled(double v, double s) {
  return __double2loint(__fma_rn(v, s, 6755399441055744.0));
}
// |v| < 2^26  (so that |32 v| < 2^31); false for NaN/Inf
__device__ __forceinline__ bool coord_in_range(double v) {
  return (unsigned)(__double2hiint(v) & 0x7fffffff) < 0x41900000u;
}

__host__ __device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
  }
  return i;
}

// cvtColor(BGR2GRAY) for 8-bit input: 15-bit fixed point (OpenCV RGB2Gray<uchar>), exact.
__device__ __forceinline__ int bgr2gray(int b, int g, int r) {
  return (b * 3735 + g * 19235 + r * 9798 + 16384) >> 15;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace stk
