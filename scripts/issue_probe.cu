// What bounds a mixed FP32 instruction stream on sm_100 — the issue port, the FMA pipe or register-file reads?
// Each test runs 2 x 256-thread blocks per SM (4 warps per scheduler, the geometry of the ECC iteration kernel) and
// reports elapsed SM cycles per loop trip per scheduler / 4 warps, next to the number of instructions in the trip.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/issue_probe scripts/issue_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }

constexpr int N = 16;

template <int T>
__global__ void __launch_bounds__(256, 2) probe(float* out, int iters, float a, float b, int ia) {
  __shared__ float sm[2048];
  for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = (float)(i % 97) * 0.01f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float r = 0.f;
  float2 acc[N], x[N];
  float s[2 * N], y[2 * N];
  int k[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { acc[i] = f2(a + i, b - i); x[i] = f2(a * i + lane, b * i - lane); k[i] = ia + i * lane; }
#pragma unroll
  for (int i = 0; i < 2 * N; ++i) { s[i] = a * i + lane; y[i] = b + i; }
  const float2 ca = f2(a, a * 1.5f), cb = f2(b, b * 0.5f);
  int off = lane;
  for (int it = 0; it < iters; ++it) {
    if (T == 1 || T == 2 || T == 8 || T == 9) {          // packed, accumulator pair + two invariant pairs
#pragma unroll
      for (int i = 0; i < N; ++i) acc[i] = fma2(acc[i], ca, cb);
    }
    if (T == 2 || T == 5) {                              // + N integer ALU instructions (independent)
#pragma unroll
      for (int i = 0; i < N; ++i) k[i] = (k[i] ^ ia) + i;
    }
    if (T == 3) {                                        // packed: distinct pair * broadcast scalar + accumulator pair
#pragma unroll
      for (int i = 0; i < N; ++i) acc[i] = fma2(x[i], f2(y[i & 3]), acc[i]);
    }
    if (T == 4 || T == 5 || T == 10) {                   // scalar: 2N fma, distinct s[i], shared y, own accumulator
#pragma unroll
      for (int i = 0; i < N; ++i) { acc[i].x = fmaf(s[2 * i], y[i & 3], acc[i].x); acc[i].y = fmaf(s[2 * i + 1], y[(i + 1) & 3], acc[i].y); }
    }
    if (T == 6) {                                        // packed square: x*x + acc
#pragma unroll
      for (int i = 0; i < N; ++i) acc[i] = fma2(x[i], x[i], acc[i]);
    }
    if (T == 7) {                                        // packed, three distinct pairs
#pragma unroll
      for (int i = 0; i < N; ++i) acc[i] = fma2(x[i], x[(i + 5) % N], acc[i]);
    }
    if (T == 8 || T == 10) {                             // + N shared-memory loads (conflict-free, independent)
#pragma unroll
      for (int i = 0; i < N; ++i) s[i] += sm[(off + 36 * i) & 2047];
      off = (off + 7) & 2047;
    }
    if (T == 9) {                                        // + N scalar fma with an immediate-like invariant operand
#pragma unroll
      for (int i = 0; i < N; ++i) s[i] = fmaf(s[i], a, b);
    }
    if (T == 11) {                                       // scalar accumulate with immediate multiplier: acc += s * const
#pragma unroll
      for (int i = 0; i < N; ++i) { acc[i].x = fmaf(s[2 * i], 3.0f, acc[i].x); acc[i].y = fmaf(s[2 * i + 1], 5.0f, acc[i].y); }
    }
    if (T == 12) {                                       // scalar add: acc += s
#pragma unroll
      for (int i = 0; i < N; ++i) { acc[i].x += s[2 * i]; acc[i].y += s[2 * i + 1]; }
    }
    if (T == 13) {                                       // packed add: acc += x
#pragma unroll
      for (int i = 0; i < N; ++i) acc[i] = __fadd2_rn(acc[i], x[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) r += acc[i].x + acc[i].y + x[i].x + (float)k[i];
#pragma unroll
  for (int i = 0; i < 2 * N; ++i) r += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int T>
void run(const char* name, float* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  float best = 1e9f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    probe<T><<<148 * 2, 256>>>(out, iters, 1.0001f, 0.5f, 12345);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double cycles = best * 1e-3 * khz * 1e3;
  printf("%-64s %8.3f ms  %6.2f cycles per trip per warp-slot\n", name, best, cycles / ((double)iters * 4));
}

int main() {
  float* out; cudaMalloc(&out, 148 * 2 * 256 * 4);
  run<1>("T1  16 FFMA2 acc*c+d (invariant pairs)", out);
  run<2>("T2  T1 + 16 integer ALU pairs (LOP3+IADD)", out);
  run<3>("T3  16 FFMA2 pair*bcast+acc", out);
  run<4>("T4  32 FFMA s*y+acc", out);
  run<5>("T5  T4 + 16 integer ALU pairs", out);
  run<6>("T6  16 FFMA2 x*x+acc", out);
  run<7>("T7  16 FFMA2 x*x'+acc (three distinct pairs)", out);
  run<8>("T8  T1 + 16 LDS+FADD", out);
  run<9>("T9  T1 + 16 FFMA s*a+b (invariant operands)", out);
  run<10>("T10 T4 + 16 LDS+FADD", out);
  run<11>("T11 32 FFMA s*imm+acc", out);
  run<12>("T12 32 FADD acc+=s", out);
  run<13>("T13 16 FADD2 acc+=x", out);
  return 0;
}
