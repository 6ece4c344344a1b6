// Microbenchmark: FP32 FMA issue rate with scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100.
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float2 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }
      else x[i] = __ffma2_rn(x[i], aa, bb);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f); else k<1><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = 148.0 * 8 * 256 * 16.0 * iters;
      printf("mode %d (%s): %.3f ms  %.1f TFLOP/s (2 flop per fma)\n", mode, mode ? "FFMA2" : "FFMA", ms, 2 * fma / ms / 1e9);
    }
  }
  return 0;
}
