// What do the pieces of K2's pixel body really cost on the FP32 pipe of sm_100?  (round 2: the ncu capture shows
// "math pipe throttle" as the top stall while the fma pipe reads 55 % busy — the packed f32x2 instructions
// must hold the pipe longer than the 2 cycles measured with loop-invariant operands in ffma2_probe.cu.)
// Each variant runs the same geometry as the real kernel (2 x 256-thread blocks per SM, 4 warps per scheduler)
// and reports cycles per call per scheduler = elapsed SM cycles / (calls per warp x 4 warps).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I libstacker.rs_b200/csrc -o scripts/pipe_probe scripts/pipe_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <type_traits>
#include "ecc_iter_v2.cuh"

using namespace stk;

template <int MODE>
__global__ void __launch_bounds__(256, 2) probe(float* out, int iters, float a, float b, unsigned magic) {
  __shared__ float box[kBoxW * 32];
  for (int i = threadIdx.x; i < kBoxW * 32; i += 256) box[i] = (float)(i % 97) * 0.01f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float r = 0.f;
  if (MODE == 0) {                       // the accumulate of one interior pixel (AccumH2::add_packed)
    AccumH2 acc; acc.clear();
    float2 g01 = f2(a + lane, b - lane); float g2 = a * b, w_ = a + 1.f, t_ = b + 2.f, yf = 3.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc.add_packed(g01, g2, w_, t_, yf);
        g01.x += a; g01.y += b; g2 += a; w_ += b; t_ += a; yf += 1.f;      // 6 scalar adds so nothing is loop-invariant
      }
    }
    float v[AccumH2::L::NV]; acc.emit<true>(1.f, v);
    for (int i = 0; i < AccumH2::L::NV; ++i) r += v[i];
  } else if (MODE == 1) {                // the sampling of one pixel: 12 LDS + packed interpolation
    float sw = 0.f; float2 sg = f2(0.f);
    float ax = a * 0.1f, ay = b * 0.1f;
    int off = kBoxW + 1 + lane;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float w_; float2 g;
        sample_box_packed(box + off + k * kBoxW, ax, ay, w_, g);
        sw += w_; sg = add2(sg, g);
        ax += 0.001f; ay += 0.002f;
      }
      off = (off + 3) & 1023;
    }
    r = sw + sg.x + sg.y;
  } else if (MODE == 2) {                // 9 packed FMAs per call: acc = q * y + acc with three distinct register pairs each
    float2 h[9]; for (int i = 0; i < 9; ++i) h[i] = f2(0.f);
    float2 q0 = f2(a, b), q1 = f2(b, a), q2 = f2(a + b, a - b); float yf = 1.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float2 y2 = f2(yf), yy2 = f2(yf * yf);
        h[0] = add2(h[0], q0); h[1] = fma2(q0, y2, h[1]); h[2] = fma2(q0, yy2, h[2]);
        h[3] = add2(h[3], q1); h[4] = fma2(q1, y2, h[4]); h[5] = fma2(q1, yy2, h[5]);
        h[6] = add2(h[6], q2); h[7] = fma2(q2, y2, h[7]); h[8] = fma2(q2, yy2, h[8]);
        q0.x += a; q1.y += b; q2.x += a; yf += 1.f;
      }
    }
    for (int i = 0; i < 9; ++i) r += h[i].x + h[i].y;
  } else if (MODE == 3) {                // the same 18 sums with scalar instructions
    float h[18]; for (int i = 0; i < 18; ++i) h[i] = 0.f;
    float q[6] = {a, b, b, a, a + b, a - b}; float yf = 1.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float yy = yf * yf;
#pragma unroll
        for (int j = 0; j < 6; ++j) { h[3 * j] += q[j]; h[3 * j + 1] = fmaf(q[j], yf, h[3 * j + 1]); h[3 * j + 2] = fmaf(q[j], yy, h[3 * j + 2]); }
        q[0] += a; q[3] += b; q[4] += a; yf += 1.f;
      }
    }
    for (int i = 0; i < 18; ++i) r += h[i];
  } else if (MODE == 5 || MODE == 6) {   // the premultiplied accumulator: packed interior form (5), scalar form (6)
    AccumH3 acc; acc.clear();
    float2 g01 = f2(a + lane, b - lane); float g2 = a * b, w_ = a + 1.f, t_ = b + 2.f, yf = 3.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (MODE == 5) acc.add_packed(g01, g2, w_, t_, yf);
        else { const float g[3] = {g01.x, g01.y, g2}; acc.add<true>(g, w_, t_, 1.f, yf); }
        g01.x += a; g01.y += b; g2 += a; w_ += b; t_ += a; yf += 1.f;
      }
    }
    float v[AccumH3::L::NV]; acc.emit<true>(1.f, v);
    for (int i = 0; i < AccumH3::L::NV; ++i) r += v[i];
  } else if (MODE >= 7 && MODE <= 10) {  // the whole interior pixel (ecc_iter_v2.cuh::lean_pixel): 7 = AccumH2, 8 = AccumH3, 9 = AccumH3 + leaner body, 10 = 9 in 4-row groups
    typename std::conditional<MODE == 7, AccumH2, AccumH3>::type acc; acc.clear();
    FastPersp fp; fp.alpha = a * 0.3f; fp.beta = 1e-4f * b; fp.gamma = 0.2f * b; fp.delta = 1e-4f * a; fp.wc = 1.f + 1e-6f * lane; fp.m21 = 1e-7f;
    const float xf = (float)lane;
    float yf0 = 0.f;
    constexpr unsigned kMagicHi = 0x4B400000u >> kInterBits;
    const unsigned bi0 = (unsigned)(2 * kBoxW + 2 + lane) - kMagicHi * (unsigned)(kBoxW + 1);
    constexpr int U = MODE == 10 ? 4 : 8;
    for (int it = 0; it < iters * (8 / U); ++it) {
#pragma unroll
      for (int k = 0; k < U; ++k)
        lean_pixel<(MODE >= 9) ? 1 : 0>(fp, xf, yf0 + (float)k, box[(k + 20) * kBoxW + lane], box, bi0 + (unsigned)(k * kBoxW), magic, acc);
      yf0 = (yf0 < 8.f) ? yf0 + (float)U : 0.f;           // stay inside the 32-row box
    }
    float v[AccumH2::L::NV]; acc.template emit<true>(1.f, v);
    for (int i = 0; i < AccumH2::L::NV; ++i) r += v[i];
  } else if (MODE == 4) {                // packed FMA, accumulator pair only varying (the ffma2_probe case)
    float2 x[9]; for (int i = 0; i < 9; ++i) x[i] = f2(a + i, b + i);
    const float2 aa = f2(a), bb = f2(b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int i = 0; i < 9; ++i) x[i] = fma2(x[i], aa, bb);
    }
    for (int i = 0; i < 9; ++i) r += x[i].x + x[i].y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, float* out, double pipe_cycles_expected) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  float best = 1e9f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    probe<MODE><<<148 * 2, 256>>>(out, iters, 1.0001f, 0.5f, 0x4B400000u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double cycles = best * 1e-3 * khz * 1e3;
  const double calls_per_sched = (double)iters * 8 * 4;      // 4 warps per scheduler
  printf("%-44s %8.3f ms  %6.1f cycles per call per scheduler (model: %.0f)  [clock %.0f MHz nominal]\n", name, best,
         cycles / calls_per_sched, pipe_cycles_expected, khz / 1e3);
}

int main() {
  float* out; cudaMalloc(&out, 148 * 2 * 256 * 4);
  run<0>("accumulate (AccumH2::add_packed) + 6 FADD", out, 54 + 6);
  run<1>("sample (12 LDS + packed bilinear) + 4", out, 26 + 6);
  run<2>("9 packed sums, 3 distinct pairs (+yy, 4 FADD)", out, 18 + 5);
  run<3>("18 scalar sums (+yy, 4 FADD)", out, 18 + 5);
  run<4>("9 FFMA2, invariant multiplicand/addend", out, 18);
  run<5>("accumulate premultiplied packed (AccumH3) + 6 FADD", out, 44 + 6);
  run<6>("accumulate premultiplied scalar (AccumH3) + 6 FADD", out, 44 + 6);
  run<7>("whole interior pixel, AccumH2", out, 97);
  run<8>("whole interior pixel, AccumH3", out, 87);
  run<9>("whole interior pixel, AccumH3, leaner body", out, 87);
  run<10>("whole interior pixel, AccumH3, leaner body, 4-row groups", out, 87);
  return 0;
}
