// Sweep descriptor parameters to find what this GPU/driver accepts for a plain f32 2D tiled load.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

__global__ void probe2d(const __grid_constant__ CUtensorMap tensor_map, float* out, int x, int y, int n) {
  extern __shared__ __align__(1024) unsigned char dyn[];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(dyn, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, n * 4);
  } else token = bar.arrive();
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float*>(dyn)[i];
}
__global__ void probe3d(const __grid_constant__ CUtensorMap tensor_map, float* out, int x, int y, int n) {
  extern __shared__ __align__(1024) unsigned char dyn[];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_3d_global_to_shared(dyn, &tensor_map, x, y, 0, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, n * 4);
  } else token = bar.arrive();
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float*>(dyn)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int bw = atoi(argv[1]), bh = atoi(argv[2]), swz = atoi(argv[3]), rank = atoi(argv[4]), x0 = atoi(argv[5]);
  printf("bw %d bh %d swizzle %d rank %d x0 %d: ", bw, bh, swz, rank, x0);
  const int W = 512, H = 240, pitch = 512;
  std::vector<float> h(pitch * H);
  for (int y = 0; y < H; ++y) for (int x = 0; x < pitch; ++x) h[y * pitch + x] = y * 1000 + x;
  float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)ptr;
  alignas(64) CUtensorMap tm;
  cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, 1}; cuuint64_t gs[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * 4 * H};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}; cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, gdim, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  (CUtensorMapSwizzle)swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 0; }
  float* out; cudaMalloc(&out, bw * bh * 4);
  cudaFuncSetAttribute(probe2d, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  cudaFuncSetAttribute(probe3d, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  if (rank == 2) probe2d<<<1, 128, bw * bh * 4>>>(tm, out, x0, 21, bw * bh);
  else probe3d<<<1, 128, bw * bh * 4>>>(tm, out, x0, 21, bw * bh);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> o(bw * bh); cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) if (o[y * bw + x] != (21 + y) * 1000 + x0 + x) ++bad;
  printf("OK, mismatches vs linear layout: %d (first %g %g row1 %g)\n", bad, o[0], o[1], o[bw]);
  return 0;
}
