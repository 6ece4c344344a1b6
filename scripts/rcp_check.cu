// GPU check: rcp_rn_normal (csrc/warp_acc.cuh) == IEEE 1.0 / w, bit for bit, over 2^28 doubles drawn from
// the ranges the interior warp path can see (w ~ 1, and wide-exponent values).   nvcc -arch=sm_100a, run on a B200.
#include <cstdio>
#include <cstdint>
#include "../libstacker.rs_b200/csrc/warp_acc.cuh"

__device__ unsigned long long splitmix(unsigned long long& s) {
  unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void check(unsigned long long* bad, int per_thread) {
  unsigned long long s = 0x1234567ull + (unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x) * 7919ull;
  unsigned long long n = 0;
  for (int i = 0; i < per_thread; ++i) {
    const unsigned long long r = splitmix(s);
    double w;
    if (i & 1) {   // near 1: mantissa random, exponent in [-2, 2]
      w = __longlong_as_double((r & 0x000FFFFFFFFFFFFFull) | ((unsigned long long)(1021 + (r >> 60) % 5) << 52));
    } else {       // wide: exponent in [-30, 30]
      w = __longlong_as_double((r & 0x000FFFFFFFFFFFFFull) | ((unsigned long long)(993 + (r >> 56) % 61) << 52));
    }
    if (r & (1ull << 55)) w = -w;
    const double a = stk::rcp_rn_normal(w), b = __ddiv_rn(1.0, w);
    if (__double_as_longlong(a) != __double_as_longlong(b)) ++n;
  }
  if (n) atomicAdd(bad, n);
}

int main() {
  unsigned long long* d; unsigned long long h = 0;
  cudaMalloc(&d, 8); cudaMemset(d, 0, 8);
  check<<<1024, 256>>>(d, 1024);
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("rcp_rn_normal vs 1/w: %llu mismatches in %llu values (%s)\n", h, 1024ull * 256 * 1024, cudaGetErrorString(cudaGetLastError()));
  return h != 0;
}
