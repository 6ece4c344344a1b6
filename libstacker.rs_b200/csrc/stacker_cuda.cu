// C ABI of the sm_100a align-and-stack library (see include/stacker_cuda.h).
//
// Host-side orchestration only: contexts, lanes (one CUDA stream + one CUDA graph with a device-driven
// WHILE loop per lane), pinned staging and result records.  All arithmetic is in the kernels included
// below; there is no CPU fallback.
#include "../../include/stacker_cuda.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "ecc_iter.cuh"
#include "ecc_iter_v2.cuh"
#include "prep.cuh"
#include "resize_area.cuh"
#include "tenengrad.cuh"
#include "warp_acc.cuh"
#include "peer_reduce.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(STK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// getGaussianKernel(k, sigma <= 0, CV_32F): tabulated dyadic taps for k <= 9 (as cv2 4.13 returns them),
// sampled Gaussian otherwise.
void gaussian_taps(int k, float* taps) {
  static const float t1[] = {1.f};
  static const float t3[] = {0.25f, 0.5f, 0.25f};
  static const float t5[] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
  static const float t7[] = {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f};
  static const float t9[] = {0.015625f, 0.05078125f, 0.1171875f, 0.19921875f, 0.234375f,
                             0.19921875f, 0.1171875f, 0.05078125f, 0.015625f};
  const float* tab = k == 1 ? t1 : k == 3 ? t3 : k == 5 ? t5 : k == 7 ? t7 : k == 9 ? t9 : nullptr;
  if (tab) { for (int i = 0; i < k; ++i) taps[i] = tab[i]; return; }
  const double sigma = 0.3 * ((k - 1) * 0.5 - 1) + 0.8;
  const double scale2x = -0.5 / (sigma * sigma);
  std::vector<double> v(k);
  double sum = 0;
  for (int i = 0; i < k; ++i) { const double x = i - (k - 1) * 0.5; v[i] = std::exp(scale2x * x * x); sum += v[i]; }
  for (int i = 0; i < k; ++i) taps[i] = (float)(v[i] / sum);
}

// OpenCV's inverse of the forward map, f64, same operation order as imgwarp.cpp / cv::invert 3x3.
// Compiled for the host without FMA contraction (x86-64 baseline), like the OpenCV it mirrors.
void invert_perspective_host(const double* s, double* o) {
  auto det2 = [](double a, double b, double c, double d) { return a * d - b * c; };
  double d = s[0] * det2(s[4], s[5], s[7], s[8]) - s[1] * det2(s[3], s[5], s[6], s[8]) +
             s[2] * det2(s[3], s[4], s[6], s[7]);
  if (d == 0.0) { for (int i = 0; i < 9; ++i) o[i] = 0.0; return; }
  d = 1.0 / d;
  o[0] = det2(s[4], s[5], s[7], s[8]) * d;
  o[1] = det2(s[2], s[1], s[8], s[7]) * d;
  o[2] = det2(s[1], s[2], s[4], s[5]) * d;
  o[3] = det2(s[5], s[3], s[8], s[6]) * d;
  o[4] = det2(s[0], s[2], s[6], s[8]) * d;
  o[5] = det2(s[2], s[0], s[5], s[3]) * d;
  o[6] = det2(s[3], s[4], s[6], s[7]) * d;
  o[7] = det2(s[1], s[0], s[7], s[6]) * d;
  o[8] = det2(s[0], s[1], s[3], s[4]) * d;
}

// cv::warpAffine's inverse of the 2x3 forward map (imgwarp.cpp), f64, same operation order; `o` = [iM00 iM01 iM02;
// iM10 iM11 iM12] followed by the projective row (0 0 1) the kernel's parameter block carries.
void invert_affine_host(const double* m, double* o) {
  double d = m[0] * m[4] - m[1] * m[3];
  d = d != 0.0 ? 1.0 / d : 0.0;
  const double a11 = m[4] * d, a22 = m[0] * d;
  o[0] = a11; o[1] = m[1] * (-d); o[3] = m[3] * (-d); o[4] = a22;
  o[2] = -o[0] * m[2] - o[1] * m[5];
  o[5] = -o[3] * m[2] - o[4] * m[5];
  o[6] = 0.0; o[7] = 0.0; o[8] = 1.0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point table (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_plane_tensor_map(CUtensorMap* tm, const float* base, int width, int height, int pitch_floats, int box_w,
                          int box_h) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr)
      return fail(STK_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    fn = (EncodeTiledFn)ptr;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)width, (cuuint64_t)height};
  const cuuint64_t gstride[1] = {(cuuint64_t)pitch_floats * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(STK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return STK_OK;
}

// computeResizeAreaTab (OpenCV imgproc/resize.cpp), as restated in oracle/restate.py::_area_tab: per
// destination index the contiguous run of source indices and their f32 weights, built in f64.
struct AreaTab {
  std::vector<int> first, count;
  std::vector<float> w;     // [dsize][k], zero padded
  int k = 0;
};

void build_area_tab(int ssize, int dsize, double scale, AreaTab& t) {
  std::vector<std::vector<float>> wts(dsize);
  t.first.assign(dsize, 0);
  t.count.assign(dsize, 0);
  t.k = 1;
  for (int dx = 0; dx < dsize; ++dx) {
    const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    const double cell = std::min(scale, ssize - fsx1);
    int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
    sx2 = std::min(sx2, ssize - 1);
    sx1 = std::min(sx1, sx2);
    int first = sx1;
    if (sx1 - fsx1 > 1e-3) { first = sx1 - 1; wts[dx].push_back((float)((sx1 - fsx1) / cell)); }
    for (int sx = sx1; sx < sx2; ++sx) wts[dx].push_back((float)(1.0 / cell));
    if (fsx2 - sx2 > 1e-3) wts[dx].push_back((float)(std::min(std::min(fsx2 - sx2, 1.0), cell) / cell));
    t.first[dx] = std::max(first, 0);
    t.count[dx] = (int)wts[dx].size();
    t.k = std::max(t.k, t.count[dx]);
  }
  t.w.assign((size_t)dsize * t.k, 0.f);
  for (int dx = 0; dx < dsize; ++dx)
    for (size_t j = 0; j < wts[dx].size(); ++j) t.w[(size_t)dx * t.k + j] = wts[dx][j];
}

// device-side description of one INTER_AREA downscale (sw x sh -> dw x dh)
struct AreaPlan {
  int sw = 0, sh = 0, dw = 0, dh = 0;
  int ix = 0, iy = 0;              // integer-scale fast path when > 0
  int kx = 0, ky = 0;
  int *xfirst = nullptr, *xcount = nullptr, *yfirst = nullptr, *ycount = nullptr;
  float *xw = nullptr, *yw = nullptr;
  bool up = false;                 // an axis enlarges: OpenCV's 8-bit bilinear "area mode" (xfirst/yfirst + coefficients)
  int *xcoef = nullptr, *ycoef = nullptr;   // [2 * dsize] 11-bit fixed-point weights
};

void free_area_plan(AreaPlan& a) {
  cudaFree(a.xfirst); cudaFree(a.xcount); cudaFree(a.yfirst); cudaFree(a.ycount); cudaFree(a.xw); cudaFree(a.yw);
  cudaFree(a.xcoef); cudaFree(a.ycoef);
  a = AreaPlan();
}

// cv::resize's coefficient table when INTER_AREA is not a pure down-scale ("area mode" of the bilinear path,
// imgproc/resize.cpp), as restated in oracle/restate.py::_linear_area_tab: first source index and the two weights
// saturate_cast<short>(w * INTER_RESIZE_COEF_SCALE) per destination index.
void build_linear_area_tab(int ssize, int dsize, std::vector<int>& ofs, std::vector<int>& coef) {
  const double inv = (double)dsize / ssize, scale = 1.0 / inv;
  ofs.assign(dsize, 0);
  coef.assign((size_t)dsize * 2, 0);
  for (int d = 0; d < dsize; ++d) {
    int s = (int)std::floor(d * scale);
    float f = (float)((d + 1) - (s + 1) * inv);
    f = f <= 0 ? 0.f : f - (float)(int)std::floor(f);
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    ofs[d] = s;
    auto to_short = [](float v) { const long r = std::lrintf(v); return (int)std::max(-32768l, std::min(32767l, r)); };
    coef[2 * d] = to_short((1.f - f) * 2048.f);
    coef[2 * d + 1] = to_short(f * 2048.f);
  }
}

int make_area_plan(int sw, int sh, int dw, int dh, AreaPlan& a) {
  a.sw = sw; a.sh = sh; a.dw = dw; a.dh = dh;
  if (dw > sw || dh > sh) {
    std::vector<int> xo, xc, yo, yc;
    build_linear_area_tab(sw, dw, xo, xc);
    build_linear_area_tab(sh, dh, yo, yc);
    auto up = [](const void* h, size_t bytes, void** d) -> bool {
      return cudaMalloc(d, bytes) == cudaSuccess && cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    const bool ok = up(xo.data(), sizeof(int) * dw, (void**)&a.xfirst) && up(xc.data(), sizeof(int) * 2 * dw, (void**)&a.xcoef) &&
                    up(yo.data(), sizeof(int) * dh, (void**)&a.yfirst) && up(yc.data(), sizeof(int) * 2 * dh, (void**)&a.ycoef);
    if (!ok) { free_area_plan(a); return fail(STK_ERR_NOMEM, "cannot upload the resize tables"); }
    a.up = true;
    return STK_OK;
  }
  // cv::resize: inv_scale = dsize / ssize ; scale = 1 / inv_scale ; is_area_fast when both are integers
  const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
  const int ix = (int)std::nearbyint(scale_x), iy = (int)std::nearbyint(scale_y);
  if (std::fabs(scale_x - ix) < 2.220446049250313e-16 && std::fabs(scale_y - iy) < 2.220446049250313e-16) {
    a.ix = ix; a.iy = iy;
    return STK_OK;
  }
  AreaTab tx, ty;
  build_area_tab(sw, dw, scale_x, tx);
  build_area_tab(sh, dh, scale_y, ty);
  a.kx = tx.k; a.ky = ty.k;
  auto up = [](const void* h, size_t bytes, void** d) -> bool {
    return cudaMalloc(d, bytes) == cudaSuccess && cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
  };
  const bool ok = up(tx.first.data(), sizeof(int) * dw, (void**)&a.xfirst) && up(tx.count.data(), sizeof(int) * dw, (void**)&a.xcount) &&
                  up(tx.w.data(), sizeof(float) * tx.w.size(), (void**)&a.xw) && up(ty.first.data(), sizeof(int) * dh, (void**)&a.yfirst) &&
                  up(ty.count.data(), sizeof(int) * dh, (void**)&a.ycount) && up(ty.w.data(), sizeof(float) * ty.w.size(), (void**)&a.yw);
  if (!ok) { free_area_plan(a); return fail(STK_ERR_NOMEM, "cannot upload the INTER_AREA tables"); }
  return STK_OK;
}

int launch_resize(const AreaPlan& a, const uint8_t* d_src, size_t pitch, int channels, uint8_t* d_dst, int dst_pitch,
                  cudaStream_t s) {
  stk::ResizeAreaParams p = {};
  p.src = d_src; p.src_pitch = pitch; p.dst = d_dst; p.dst_pitch = dst_pitch;
  p.sw = a.sw; p.sh = a.sh; p.dw = a.dw; p.dh = a.dh; p.channels = channels;
  p.ix = a.ix; p.iy = a.iy;
  p.up = a.up ? 1 : 0; p.xcoef = a.xcoef; p.ycoef = a.ycoef;
  p.xfirst = a.xfirst; p.xcount = a.xcount; p.xw = a.xw; p.kx = a.kx;
  p.yfirst = a.yfirst; p.ycount = a.ycount; p.yw = a.yw; p.ky = a.ky;
  dim3 block(stk::kResizeBX, stk::kResizeBY);
  dim3 grid((a.dw + stk::kResizeBX - 1) / stk::kResizeBX, (a.dh + stk::kResizeBY - 1) / stk::kResizeBY);
  stk::resize_area_grey_kernel<<<grid, block, 0, s>>>(p);
  CU(cudaGetLastError());
  return STK_OK;
}

struct Lane {
  cudaStream_t stream = nullptr;
  float* tmpl = nullptr;            // T plane
  uint8_t* d_frames[stk::kWarpBatch] = {};  // device staging for host-submitted frames: one per slot of the warp batch
                                            // (a frame's bytes must outlive its deferred final warp)
  stk::EccState* pend_st = nullptr;         // [kWarpBatch] copies of the finished ECC state (inverse map, status) per pending frame
  stk::WarpFrame pend[stk::kWarpBatch];     // frames whose final warp + accumulate is queued for the next batched launch
  int n_pend = 0;
  unsigned staging_used = 0;                // bit k: d_frames[k] holds a pending frame
  bool pend_persp = true;
  uint8_t* small = nullptr;         // downscaled grey (ecc_match_scaling_down), ew x eh, small_pitch bytes per row
  uint8_t* h_stage = nullptr;       // pinned staging for pageable host buffers
  int* h_cont = nullptr;            // pinned "loop continues" word of the host-driven loop (STK_LOOP_MODE=host)
  cudaEvent_t stage_free = nullptr; // H2D out of h_stage finished
  cudaEvent_t drained = nullptr;    // lane's queued work finished (set_reference / the peer exchange join the lanes on the device)
  bool wait_x = false;              // the next accumulator write of this lane must wait for the exchange in flight (x_done)
  stk::EccState* st = nullptr;
  double* partials = nullptr;
  float* acc = nullptr;
  bool acc_used = false;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraphConditionalHandle handle = 0;
  CUtensorMap tm_tmpl;
};

// Host feed: a ring of pinned frame buffers owned by the context.  Decode tasks fill a
// buffer (straight from the decoder, or one memcpy from a pageable Mat) WITHOUT holding the context lock and
// hand it over; the upload is asynchronous and the buffer returns to the ring guarded by its `uploaded` event.
struct RingBuf {
  uint8_t* host = nullptr;
  cudaEvent_t uploaded = nullptr;
};

// Multi-GPU exchange (csrc/peer_reduce.cuh): this context's view of every rank's partial stack / flag block and of
// the root's output buffer, mapped through CUDA IPC (one process per GPU) or peer access (one process).
struct PeerLink {
  bool connected = false;
  int rank = 0, world = 0;
  uint32_t step = 0;
  uint32_t* flags = nullptr;                  // local flag block (cudaMalloc, exported)
  const float* partial[stk::kMaxPeers] = {};
  uint32_t* pflags[stk::kMaxPeers] = {};
  float* root_out = nullptr;
  size_t slice_begin = 0, slice_end = 0;      // this rank's slice of the last exchange
  bool scattered = false;                     // last exchange was a reduce-scatter (result slice in d_out)
  std::vector<void*> opened;                  // IPC mappings to close on disconnect
};

struct ResultSlot {
  int64_t tag;
  stk::EccState* host;      // pinned copy of the frame's final state (null for warp-only frames)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // profiling: prep | loop | warp boundaries
};

}  // namespace

struct stk_ecc_ctx {
  stk_ecc_config cfg;
  int device = 0;
  int sm_count = 0;
  int n_lanes = 0;
  std::vector<Lane> lanes;
  float* img = nullptr;             // I plane (blurred reference grey)
  uint8_t* d_ref = nullptr;         // device copy of the reference frame when it came from the host
  float* d_out = nullptr;           // scaled result before D2H
  int ew = 0, eh = 0;               // size of the planes ECC runs on (== frame size unless scaling down)
  bool scaled = false;
  AreaPlan area;
  int small_pitch = 0;
  int pitch_f = 0;                  // floats per row of I / T
  size_t frame_bytes = 0;           // width*channels*height (dense staging)
  size_t acc_floats = 0;
  // ECC tiling
  int n_strips = 0, chunks_per_strip = 0, n_tiles = 0, nv = 0;   // n_tiles = persistent blocks of the ECC kernel
  int max_iter = 0;
  double eps = 0;
  CUtensorMap tm_img;
  bool exact_coords = false;
  bool pdl = false;            // programmatic dependent launch between the chained iteration kernels (opt-in STK_ECC_PDL=1: measured no gain on one lane and -5 % with four, the waiting blocks hold SM slots)
  int loop_unroll = 4;         // iteration kernels per WHILE-body pass (STK_ECC_UNROLL)
  int warp_batch = stk::kWarpBatch;   // frames per final-warp launch, 1..kWarpBatch (STK_WARP_BATCH)
  int warp_gen = 2;                   // K4 generation: 2 = word loads + guarded f32 coordinates, 1 = the round-1 kernel (STK_WARP_GEN)
  int rim_weight = 10;         // cost of a rim-strip chunk in 1/8 of an interior one (STK_ECC_RIM_WEIGHT)
  void* iter_fn = nullptr;     // the iteration kernel of this context (generation + geometry, see iter_variant)
  int iter_threads = 0, iter_smem = 0, iter_chunk_h = 0, iter_box_h = 0, iter_min_blocks = 0;
  int iter_gen = 2, iter_cfg = -1;          // -1: the default geometry of iter_variant
  bool host_loop = false;
  bool have_ref = false;
  // Device-side ordering between stacks (no host synchronisation in reset / set_reference / the peer exchange):
  //   ref_ready  lane 0 finished the reference plane: every other lane waits for it before its first frame;
  //   x_stream   the multi-GPU exchange (lane sum, announce/wait, reduce-scatter + divide, close) runs here, behind every
  //              lane's `drained` event, so that the NEXT stack's prep and ECC iterations overlap it; x_done closes it and
  //              is what the next stack's first accumulator writes (seed, final warps) wait for.
  cudaEvent_t ref_ready = nullptr, x_done = nullptr;
  cudaStream_t x_stream = nullptr;
  // producer stream of device-resident inputs (stk_ecc_set_input_stream): each *_device submission is ordered behind it
  bool order_inputs = false;
  cudaStream_t input_stream = nullptr;
  cudaEvent_t input_ev = nullptr;
  stk::PrepParams prep_proto;
  std::mutex mu;
  std::vector<RingBuf> ring;        // lazily allocated: 2 buffers per lane
  std::vector<int> ring_free;
  std::mutex ring_mu;
  std::condition_variable ring_cv;
  int next_lane = 0;
  std::vector<ResultSlot> results;
  std::vector<stk::EccState*> state_chunks;   // pinned, kChunk states each
  size_t states_used = 0;
  std::atomic<int64_t> launches{0};
  int64_t iter_launches_counted = 0;
  bool profiling = false;
  PeerLink peer;
  static constexpr size_t kChunk = 256;
};

namespace {

// The iteration kernel of a context: generation 2 (csrc/ecc_iter_v2.cuh) in one of its compiled geometries, or
// the first-generation kernel (STK_ECC_GEN=1, kept as the measured baseline).  Homography has two coordinate
// flavours: FastPersp (default) and exact f64 (STK_ECC_EXACT_COORDS=1).  Geometries other than the default are
// compiled for the default Homography flavour only (STK_ECC_CFG picks one; scripts/k2_variants.py measures them).
struct IterVariant { void* fn; int threads, smem, chunk_h, box_h, min_blocks; };

template <int MOTION, bool EXACT, class CFG>
IterVariant v2_variant() {
  return {(void*)stk::ecc_iter_v2_kernel<MOTION, EXACT, CFG>, CFG::kThreads, CFG::kDynSmem, CFG::kChunkH, CFG::kBoxH, CFG::kMinBlocks};
}

// default geometry: 128x32 chunks (16 rows per thread), 2 TMA stages, 2 blocks per SM, premultiplied accumulator and
// the leaner pixel body — the fastest of the measured variants on one lane and on four (scripts/k2_variants.py,
// profiles/r2_summary.md: 3 412 -> 3 616 frames/s on the 13-frame 4K stack against the packed AccumH2 body, cfg 2)
using DefaultEccCfg = stk::EccCfg16;

IterVariant iter_variant(int motion, bool exact, int gen, int cfg) {
  if (gen == 1) {
    void* fn;
    switch (motion) {
      case STK_MOTION_TRANSLATION: fn = (void*)stk::ecc_iter_kernel<stk::kTranslation, true>; break;
      case STK_MOTION_EUCLIDEAN: fn = (void*)stk::ecc_iter_kernel<stk::kEuclidean, true>; break;
      case STK_MOTION_AFFINE: fn = (void*)stk::ecc_iter_kernel<stk::kAffine, true>; break;
      default: fn = exact ? (void*)stk::ecc_iter_kernel<stk::kHomography, true> : (void*)stk::ecc_iter_kernel<stk::kHomography, false>;
    }
    return {fn, stk::kEccThreads, stk::kEccDynSmem, stk::kChunkH, stk::kBoxH, 2};
  }
  switch (motion) {
    case STK_MOTION_TRANSLATION: return v2_variant<stk::kTranslation, true, DefaultEccCfg>();
    case STK_MOTION_EUCLIDEAN: return v2_variant<stk::kEuclidean, true, DefaultEccCfg>();
    case STK_MOTION_AFFINE: return v2_variant<stk::kAffine, true, DefaultEccCfg>();
    default: break;
  }
  if (exact) return v2_variant<stk::kHomography, true, DefaultEccCfg>();
  switch (cfg) {
    case 1: return v2_variant<stk::kHomography, false, stk::EccCfg1>();
    case 2: return v2_variant<stk::kHomography, false, stk::EccCfg2>();
    case 3: return v2_variant<stk::kHomography, false, stk::EccCfg3>();
    case 4: return v2_variant<stk::kHomography, false, stk::EccCfg4>();
    case 5: return v2_variant<stk::kHomography, false, stk::EccCfg5>();
    case 6: return v2_variant<stk::kHomography, false, stk::EccCfg6>();
    case 7: return v2_variant<stk::kHomography, false, stk::EccCfg7>();
    case 8: return v2_variant<stk::kHomography, false, stk::EccCfg8>();
    case 9: return v2_variant<stk::kHomography, false, stk::EccCfg9>();
    case 10: return v2_variant<stk::kHomography, false, stk::EccCfg10>();
    case 11: return v2_variant<stk::kHomography, false, stk::EccCfg11>();
    case 12: return v2_variant<stk::kHomography, false, stk::EccCfg12>();
    case 13: return v2_variant<stk::kHomography, false, stk::EccCfg13>();
    case 14: return v2_variant<stk::kHomography, false, stk::EccCfg14>();
    case 15: return v2_variant<stk::kHomography, false, stk::EccCfg15>();
    case 16: return v2_variant<stk::kHomography, false, stk::EccCfg16>();
    case 17: return v2_variant<stk::kHomography, false, stk::EccCfg17>();
    case 18: return v2_variant<stk::kHomography, false, stk::EccCfg18>();
    case 19: return v2_variant<stk::kHomography, false, stk::EccCfg19>();
    case 0: return v2_variant<stk::kHomography, false, stk::EccCfg0>();
    default: return v2_variant<stk::kHomography, false, DefaultEccCfg>();
  }
}

int model_nv(int motion) {
  switch (motion) {
    case STK_MOTION_TRANSLATION: return stk::Layout<stk::kTranslation>::NV;
    case STK_MOTION_EUCLIDEAN: return stk::Layout<stk::kEuclidean>::NV;
    case STK_MOTION_AFFINE: return stk::Layout<stk::kAffine>::NV;
    default: return stk::Layout<stk::kHomography>::NV;
  }
}

stk::EccIterParams iter_params(stk_ecc_ctx* c, Lane& ln, bool use_handle) {
  stk::EccIterParams p;
  memset(&p, 0, sizeof p);
  p.tm_img = c->tm_img;
  p.tm_tmpl = ln.tm_tmpl;
  p.img = c->img;
  p.tmpl = ln.tmpl;
  p.pitch = c->pitch_f;
  p.width = c->ew;
  p.height = c->eh;
  p.n_strips = c->n_strips;
  p.chunks_per_strip = c->chunks_per_strip;
  p.partials = ln.partials;
  p.tiles_pad = (c->n_tiles + 31) / 32 * 32;
  p.st = ln.st;
  p.handle = ln.handle;
  p.use_handle = use_handle ? 1 : 0;
  p.rim_weight = c->rim_weight;
  p.frac_magic = 0x4B400000u;
  p.totals_out = nullptr;
  p.timing_out = nullptr;
  return p;
}

// graph per lane:  init -> WHILE(handle) { ecc_iter_kernel }
int build_lane_graph(stk_ecc_ctx* c, Lane& ln) {
  const bool persp = c->cfg.motion_type == STK_MOTION_HOMOGRAPHY;
  CU(cudaGraphCreate(&ln.graph, 0));
  CU(cudaGraphConditionalHandleCreate(&ln.handle, ln.graph, 1, cudaGraphCondAssignDefault));

  cudaGraphNode_t init_node, cond_node, iter_node;
  {
    stk::EccState* st = ln.st;
    int persp_i = persp ? 1 : 0, max_iter = c->max_iter, use = 1;
    double eps = c->eps;
    cudaGraphConditionalHandle h = ln.handle;
    void* args[] = {&st, &persp_i, &max_iter, &eps, &h, &use};
    cudaKernelNodeParams kp = {};
    kp.func = (void*)stk::ecc_init_kernel;
    kp.gridDim = dim3(1);
    kp.blockDim = dim3(32);
    kp.kernelParams = args;
    CU(cudaGraphAddKernelNode(&init_node, ln.graph, nullptr, 0, &kp));
  }
  {
    cudaGraphNodeParams np = {};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = ln.handle;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    CU(cudaGraphAddNode(&cond_node, ln.graph, &init_node, 1, &np));
    cudaGraph_t body = np.conditional.phGraph_out[0];
    stk::EccIterParams ip = iter_params(c, ln, true);
    void* args[] = {&ip};
    cudaKernelNodeParams kp = {};
    kp.func = c->iter_fn;
    kp.gridDim = dim3(c->n_tiles);
    kp.blockDim = dim3(c->iter_threads);
    kp.sharedMemBytes = c->iter_smem;
    kp.kernelParams = args;
    // The WHILE body holds `loop_unroll` copies of the iteration kernel in a chain: a kernel that finds the
    // loop already finished (cont == 0) returns at once, so at most unroll-1 empty launches are spent per
    // frame, while the condition round trip of the WHILE node (measured ~10 us against ~3 us for a
    // kernel-to-kernel edge) is paid once per `loop_unroll` iterations.
    // Inside the chain the edges are PROGRAMMATIC (PDL): kernel i+1 may be launched and have its blocks
    // scheduled while kernel i is still in its serial tail; it waits at griddepcontrol.wait before reading
    // the state kernel i writes.  This takes the launch latency off the iteration's critical path.
    cudaGraphNode_t prev = nullptr;
    for (int u = 0; u < c->loop_unroll; ++u) {
      if (prev && c->pdl) {
        CU(cudaGraphAddKernelNode(&iter_node, body, nullptr, 0, &kp));
        cudaGraphEdgeData ed = {};
        ed.from_port = cudaGraphKernelNodePortProgrammatic;
        ed.to_port = cudaGraphKernelNodePortDefault;
        ed.type = cudaGraphDependencyTypeProgrammatic;
        CU(cudaGraphAddDependencies_v2(body, &prev, &iter_node, &ed, 1));
      } else {
        CU(cudaGraphAddKernelNode(&iter_node, body, prev ? &prev : nullptr, prev ? 1 : 0, &kp));
      }
      prev = iter_node;
    }
  }
  CU(cudaGraphInstantiate(&ln.exec, ln.graph, 0));
  return STK_OK;
}

// frame (full size, BGR(A)) -> blurred f32 plane at the ECC working size; `small` = the lane's downscaled
// grey buffer when the context scales down
int launch_prep(stk_ecc_ctx* c, const uint8_t* d_src, size_t pitch, uint8_t* small, float* dst, cudaStream_t s) {
  stk::PrepParams p = c->prep_proto;
  p.src = d_src;
  p.src_pitch = pitch;
  p.dst = dst;
  if (c->scaled) {
    int rc = launch_resize(c->area, d_src, pitch, c->cfg.channels, small, c->small_pitch, s);
    if (rc) return rc;
    c->launches++;
    p.src = small;
    p.src_pitch = (size_t)c->small_pitch;
    p.channels = 1;
  }
  // gauss_filt_size 3 / 5 on 4-byte aligned rows: the shared-memory-free streaming kernel (exact integers)
  static const bool stream_on = [] { const char* e = getenv("STK_PREP_STREAM"); return !(e && e[0] == '0'); }();
  static const int strip_env = [] { const char* e = getenv("STK_PREP_STRIP"); return e ? atoi(e) : 0; }();
  if (stream_on && stk::prep_stream_ok(p.src, p.src_pitch, c->ew, c->eh, p.channels, p.radius)) {
    stk::PrepStreamParams q;
    q.src = p.src; q.src_pitch = p.src_pitch; q.dst = p.dst; q.dst_pitch = p.dst_pitch;
    q.width = c->ew; q.height = c->eh;
    q.n_bands = (c->ew + stk::kPrepBandCols - 1) / stk::kPrepBandCols;
    const int key = p.channels * 10 + p.radius;
    void (*kern)(const stk::PrepStreamParams) =
        key == 11 ? stk::prep_stream_kernel<1, 1> : key == 12 ? stk::prep_stream_kernel<1, 2>
      : key == 31 ? stk::prep_stream_kernel<3, 1> : key == 32 ? stk::prep_stream_kernel<3, 2>
      : key == 41 ? stk::prep_stream_kernel<4, 1> : stk::prep_stream_kernel<4, 2>;
    // ONE resident wave: strips sized from the kernel's real occupancy (2R halo rows are re-read per strip, so
    // strips are as tall as one wave allows, at least 8 rows)
    static std::atomic<int> occ_cache[64];
    int per_sm = occ_cache[key].load();
    if (per_sm <= 0) {
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, stk::kPrepStreamThreads, 0));
      occ_cache[key].store(per_sm);
    }
    const int warp_slots = std::max(1, c->sm_count * per_sm * (stk::kPrepStreamThreads / 32));
    const int strips_wanted = std::max(1, warp_slots / q.n_bands);
    q.strip_rows = strip_env > 0 ? strip_env : std::max(8, (c->eh + strips_wanted - 1) / strips_wanted);
    q.n_strips = (c->eh + q.strip_rows - 1) / q.strip_rows;
    const int warps = q.n_bands * q.n_strips;
    const int blocks = (warps * 32 + stk::kPrepStreamThreads - 1) / stk::kPrepStreamThreads;
    kern<<<blocks, stk::kPrepStreamThreads, 0, s>>>(q);
    c->launches++;
    CU(cudaGetLastError());
    return STK_OK;
  }
  const size_t smem = stk::prep_smem_bytes(p.radius);
  dim3 grid((c->ew + stk::kPrepTW - 1) / stk::kPrepTW, (c->eh + stk::kPrepTH - 1) / stk::kPrepTH);
  switch (p.radius) {   // radii 1..4 (gauss_filt_size 3..9) get unrolled instantiations
    case 1: stk::prep_grey_blur_kernel<1><<<grid, stk::kPrepThreads, smem, s>>>(p); break;
    case 2: stk::prep_grey_blur_kernel<2><<<grid, stk::kPrepThreads, smem, s>>>(p); break;
    case 3: stk::prep_grey_blur_kernel<3><<<grid, stk::kPrepThreads, smem, s>>>(p); break;
    case 4: stk::prep_grey_blur_kernel<4><<<grid, stk::kPrepThreads, smem, s>>>(p); break;
    default: stk::prep_grey_blur_kernel<0><<<grid, stk::kPrepThreads, smem, s>>>(p); break;
  }
  c->launches++;
  CU(cudaGetLastError());
  return STK_OK;
}

// K4 for the frames queued on a lane: ONE launch gathers up to kWarpBatch frames and touches the accumulator once
int flush_warps(stk_ecc_ctx* c, Lane& ln) {
  if (ln.n_pend == 0) return STK_OK;
  if (ln.wait_x) {       // the previous stack's exchange may still read this lane's accumulator
    CU(cudaStreamWaitEvent(ln.stream, c->x_done, 0));
    ln.wait_x = false;
  }
  stk::WarpAccParams p = {};
  for (int j = 0; j < ln.n_pend; ++j) p.f[j] = ln.pend[j];
  p.n = ln.n_pend;
  p.acc = ln.acc;
  p.width = c->cfg.width;
  p.height = c->cfg.height;
  p.src_width = c->cfg.width;
  p.src_height = c->cfg.height;
  p.store = ln.acc_used ? 0 : 1;
  p.frac_magic = 0x4B400000u;
  dim3 block(stk::kWarpBX, stk::kWarpBY);
  const int ch = c->cfg.channels;
  if (c->warp_gen == 1) {
    // first-generation kernel (accumulator tile in shared memory, byte gathers): kept for A/B measurements, STK_WARP_GEN=1
    dim3 grid((p.width + stk::kWarpBX - 1) / stk::kWarpBX, (p.height + stk::kWarpTH - 1) / stk::kWarpTH);
    if (ch == 3) {
      if (ln.pend_persp) stk::warp_accumulate_kernel<3, true><<<grid, block, 0, ln.stream>>>(p);
      else stk::warp_accumulate_kernel<3, false><<<grid, block, 0, ln.stream>>>(p);
    } else {
      if (ln.pend_persp) stk::warp_accumulate_kernel<4, true><<<grid, block, 0, ln.stream>>>(p);
      else stk::warp_accumulate_kernel<4, false><<<grid, block, 0, ln.stream>>>(p);
    }
  } else {
    dim3 grid((p.width + stk::kWarpBX - 1) / stk::kWarpBX, (p.height + stk::kWarp2TH - 1) / stk::kWarp2TH);
    // word-index addressing in 32 bits when every frame of the batch starts and strides on a 4-byte boundary
    bool aligned = true;
    for (int j = 0; j < p.n; ++j)
      aligned = aligned && (reinterpret_cast<uintptr_t>(p.f[j].src) % 4 == 0) && (p.f[j].src_pitch % 4 == 0) &&
                (p.f[j].src_pitch * (size_t)p.src_height < ((size_t)1 << 32));
#define STK_LAUNCH_WARP2(C, P)                                                                         \
    do {                                                                                               \
      if (aligned) stk::warp_accumulate_v2_kernel<C, P, true><<<grid, block, 0, ln.stream>>>(p);       \
      else stk::warp_accumulate_v2_kernel<C, P, false><<<grid, block, 0, ln.stream>>>(p);              \
    } while (0)
    if (ch == 3) {
      if (ln.pend_persp) STK_LAUNCH_WARP2(3, true); else STK_LAUNCH_WARP2(3, false);
    } else {
      if (ln.pend_persp) STK_LAUNCH_WARP2(4, true); else STK_LAUNCH_WARP2(4, false);
    }
#undef STK_LAUNCH_WARP2
  }
  c->launches++;
  ln.n_pend = 0;
  ln.staging_used = 0;
  CU(cudaGetLastError());
  ln.acc_used = true;
  return STK_OK;
}

// Queue one frame's final warp + accumulate on its lane (launched when kWarpBatch frames are waiting, or at the next
// sync / finish / exchange).  from_state: the matrix and status are the lane's ECC state as of this point of the
// stream (a device-side copy is taken, the lane's state is reused by the next frame).  The frame's bytes (d_src) must
// stay valid until the batch is launched and has run: the lane's staging slots / the caller's device buffer.
int launch_warp(stk_ecc_ctx* c, Lane& ln, const uint8_t* d_src, size_t pitch, bool persp,
                const double* inv_host, const float* border, bool from_state) {
  if (ln.n_pend > 0 && ln.pend_persp != persp) { int rc = flush_warps(c, ln); if (rc) return rc; }
  const int max_batch = c->warp_batch;
  const int j = ln.n_pend;
  stk::WarpFrame& f = ln.pend[j];
  memset(&f, 0, sizeof f);
  f.src = d_src;
  f.src_pitch = pitch;
  if (from_state) {
    CU(cudaMemcpyAsync(ln.pend_st + j, ln.st, sizeof(stk::EccState), cudaMemcpyDeviceToDevice, ln.stream));
    f.inv_ptr = ln.pend_st[j].inv;            // device address arithmetic only
    f.status_ptr = &ln.pend_st[j].status;
  }
  if (inv_host) for (int i = 0; i < 9; ++i) f.inv[i] = inv_host[i];
  for (int i = 0; i < 4; ++i) f.border[i] = border ? border[i] : 0.f;
  f.border_mode = border ? (int)border[4] : STK_BORDER_CONSTANT;      // border[4]: the cv::BorderTypes value
  for (int k = 0; k < stk::kWarpBatch; ++k) if (d_src == ln.d_frames[k] && d_src) ln.staging_used |= 1u << k;
  ln.pend_persp = persp;
  ln.n_pend = j + 1;
  if (ln.n_pend >= max_batch) return flush_warps(c, ln);
  return STK_OK;
}

int alloc_result_state(stk_ecc_ctx* c, stk::EccState** out) {
  const size_t chunk = c->states_used / stk_ecc_ctx::kChunk, off = c->states_used % stk_ecc_ctx::kChunk;
  if (chunk >= c->state_chunks.size()) {
    stk::EccState* p = nullptr;
    CU(cudaHostAlloc((void**)&p, sizeof(stk::EccState) * stk_ecc_ctx::kChunk, cudaHostAllocDefault));
    memset(p, 0, sizeof(stk::EccState) * stk_ecc_ctx::kChunk);
    c->state_chunks.push_back(p);
  }
  *out = c->state_chunks[chunk] + off;
  c->states_used++;
  return STK_OK;
}

// Stage a host frame into the lane's device buffer (dense rows).  `pinned` = the caller's buffer is
// page-locked and stays valid until sync, so it is copied from directly.
int upload_frame(stk_ecc_ctx* c, Lane& ln, uint8_t* d_frame, const uint8_t* host, size_t pitch, bool pinned) {
  const size_t row = (size_t)c->cfg.width * c->cfg.channels;
  if (pinned) {
    CU(cudaMemcpy2DAsync(d_frame, row, host, pitch, row, c->cfg.height, cudaMemcpyHostToDevice, ln.stream));
    return STK_OK;
  }
  CU(cudaEventSynchronize(ln.stage_free));
  if (pitch == row) {
    memcpy(ln.h_stage, host, row * c->cfg.height);
  } else {
    for (int y = 0; y < c->cfg.height; ++y) memcpy(ln.h_stage + (size_t)y * row, host + (size_t)y * pitch, row);
  }
  CU(cudaMemcpyAsync(d_frame, ln.h_stage, row * c->cfg.height, cudaMemcpyHostToDevice, ln.stream));
  CU(cudaEventRecord(ln.stage_free, ln.stream));
  return STK_OK;
}

Lane& pick_lane(stk_ecc_ctx* c) {
  Lane& ln = c->lanes[c->next_lane];
  c->next_lane = (c->next_lane + 1) % c->n_lanes;
  return ln;
}

int ring_acquire(stk_ecc_ctx* c, int* idx) {
  std::unique_lock<std::mutex> lk(c->ring_mu);
  if (c->ring.empty()) {
    const int n = 2 * c->n_lanes;
    std::vector<RingBuf> ring(n);
    for (auto& b : ring) {
      if (cudaHostAlloc((void**)&b.host, c->frame_bytes, cudaHostAllocDefault) != cudaSuccess ||
          cudaEventCreateWithFlags(&b.uploaded, cudaEventDisableTiming) != cudaSuccess) {
        for (auto& q : ring) { if (q.host) cudaFreeHost(q.host); if (q.uploaded) cudaEventDestroy(q.uploaded); }
        return fail(STK_ERR_NOMEM, "cannot allocate the pinned frame ring (%d x %zu bytes)", n, c->frame_bytes);
      }
    }
    c->ring.swap(ring);
    for (int i = 0; i < n; ++i) c->ring_free.push_back(i);
  }
  c->ring_cv.wait(lk, [&] { return !c->ring_free.empty(); });
  const int i = c->ring_free.back();
  c->ring_free.pop_back();
  lk.unlock();
  // the upload that last used this buffer must have landed before the caller overwrites it
  if (cudaEventSynchronize(c->ring[i].uploaded) != cudaSuccess) {
    { std::lock_guard<std::mutex> g(c->ring_mu); c->ring_free.push_back(i); }
    c->ring_cv.notify_one();
    return fail(STK_ERR_CUDA, "cudaEventSynchronize(ring buffer) failed");
  }
  *idx = i;
  return STK_OK;
}

void ring_release(stk_ecc_ctx* c, int idx) {
  { std::lock_guard<std::mutex> g(c->ring_mu); c->ring_free.push_back(idx); }
  c->ring_cv.notify_one();
}

int ring_index_of(stk_ecc_ctx* c, const uint8_t* buf) {
  std::lock_guard<std::mutex> g(c->ring_mu);
  for (size_t i = 0; i < c->ring.size(); ++i) if (c->ring[i].host == buf) return (int)i;
  return -1;
}

void copy_rows(uint8_t* dst, const uint8_t* src, size_t pitch, size_t row, int height) {
  if (pitch == row) { memcpy(dst, src, row * height); return; }
  for (int y = 0; y < height; ++y) memcpy(dst + (size_t)y * row, src + (size_t)y * pitch, row);
}

// the ECC part of one frame on its lane: prep -> device loop -> warp+accumulate -> result record
int enqueue_align(stk_ecc_ctx* c, Lane& ln, const uint8_t* d_src, size_t pitch, int64_t tag) {
  ResultSlot slot;
  slot.tag = tag;
  slot.host = nullptr;
  auto mark = [&](int i) -> int {
    if (!c->profiling) return STK_OK;
    CU(cudaEventCreate(&slot.ev[i]));
    CU(cudaEventRecord(slot.ev[i], ln.stream));
    return STK_OK;
  };
  int rc = mark(0);
  if (rc) return rc;
  rc = launch_prep(c, d_src, pitch, ln.small, ln.tmpl, ln.stream);
  if (rc) return rc;
  if ((rc = mark(1))) return rc;
  const bool persp = c->cfg.motion_type == STK_MOTION_HOMOGRAPHY;
  if (!c->host_loop) {
    CU(cudaGraphLaunch(ln.exec, ln.stream));
    c->launches++;   // init kernel; the iteration launches are added from the result records
  } else {
    // host-driven fallback (STK_LOOP_MODE=host): same kernels, convergence polled every few iterations
    stk::ecc_init_kernel<<<1, 32, 0, ln.stream>>>(ln.st, persp ? 1 : 0, c->max_iter, c->eps, 0, 0);
    c->launches++;
    CU(cudaGetLastError());
    stk::EccIterParams ip = iter_params(c, ln, false);
    void* args[] = {&ip};
    int done_iters = 0;
    if (!ln.h_cont) CU(cudaHostAlloc((void**)&ln.h_cont, sizeof(int), cudaHostAllocDefault));   // once per lane
    int* h_cont = ln.h_cont;
    *h_cont = 1;
    while (*h_cont && done_iters < c->max_iter) {
      const int chunk = std::min(4, c->max_iter - done_iters);
      for (int i = 0; i < chunk; ++i)
        CU(cudaLaunchKernel(c->iter_fn, dim3(c->n_tiles), dim3(c->iter_threads), args, c->iter_smem, ln.stream));
      done_iters += chunk;
      CU(cudaMemcpyAsync(h_cont, &ln.st->cont, sizeof(int), cudaMemcpyDeviceToHost, ln.stream));
      CU(cudaStreamSynchronize(ln.stream));
    }
  }
  if ((rc = mark(2))) return rc;
  rc = launch_warp(c, ln, d_src, pitch, persp, nullptr, nullptr, true);
  if (rc) return rc;
  if ((rc = mark(3))) return rc;
  rc = alloc_result_state(c, &slot.host);
  if (rc) return rc;
  CU(cudaMemcpyAsync(slot.host, ln.st, sizeof(stk::EccState), cudaMemcpyDeviceToHost, ln.stream));
  c->results.push_back(slot);
  return STK_OK;
}

int flush_all(stk_ecc_ctx* c) {
  for (auto& ln : c->lanes) { int rc = flush_warps(c, ln); if (rc) return rc; }
  return STK_OK;
}

int sync_all(stk_ecc_ctx* c) {
  { int rc = flush_all(c); if (rc) return rc; }
  for (auto& ln : c->lanes) CU(cudaStreamSynchronize(ln.stream));
  CU(cudaStreamSynchronize(c->x_stream));
  // account for the device-launched iteration kernels and surface per-frame failures
  int first_err = STK_OK;
  int64_t iters = 0;
  for (auto& r : c->results) {
    if (!r.host) continue;
    iters += r.host->iters;
    if (r.host->status != 0 && first_err == STK_OK) first_err = r.host->status;
  }
  c->launches += iters - c->iter_launches_counted;
  c->iter_launches_counted = iters;
  if (c->peer.connected) {
    uint32_t perr = 0;
    CU(cudaMemcpy(&perr, c->peer.flags + stk::kPeerError, sizeof perr, cudaMemcpyDeviceToHost));
    if (perr) {
      cudaMemset(c->peer.flags + stk::kPeerError, 0, sizeof perr);
      return fail(STK_ERR_CUDA, "peer exchange step %u timed out waiting for another rank", perr);
    }
  }
  if (first_err == STK_ERR_ECC_NOCONV)
    return fail(first_err, "findTransformECC: the algorithm stopped before its convergence (lambda denominator <= 0)");
  if (first_err == STK_ERR_ECC_NAN) return fail(first_err, "findTransformECC: NaN encountered");
  return first_err;
}

int lane_sum(stk_ecc_ctx* c, float* out, const float* const* extra, int n_extra, bool scale, int divisor,
             cudaStream_t s) {
  stk::LaneSumParams p = {};
  int n = 0;
  for (int i = 0; i < n_extra; ++i) p.lanes[n++] = extra[i];
  p.n_lanes = n;
  p.out = out;
  p.n = c->acc_floats;
  p.apply_scale = scale ? 1 : 0;
  p.scale = scale ? (float)(1.0 / (double)divisor) : 1.f;
  const int blocks = c->sm_count * 8;
  stk::lane_sum_scale_kernel<<<blocks, 256, 0, s>>>(p);
  c->launches++;
  CU(cudaGetLastError());
  return STK_OK;
}

int check_ctx(stk_ecc_ctx* c) {
  if (!c) return fail(STK_ERR_BAD_ARG, "null context");
  CU(cudaSetDevice(c->device));
  return STK_OK;
}

}  // namespace

extern "C" {

int stk_abi_version(void) { return STK_ABI_VERSION; }
const char* stk_last_error(void) { return g_err.c_str(); }

int stk_device_count(int* count) {
  if (!count) return fail(STK_ERR_BAD_ARG, "null count");
  CU(cudaGetDeviceCount(count));
  return STK_OK;
}

int stk_scaled_size(int width, int height, float scale_down, int* sw, int* sh) {
  if (!sw || !sh) return fail(STK_ERR_BAD_ARG, "null argument");
  if (width <= 0 || height <= 0) return fail(STK_ERR_BAD_ARG, "bad frame size %dx%d", width, height);
  if (scale_down >= (float)width)
    return fail(STK_ERR_BAD_ARG, "scale_down_to was larger (or equal) to the full image width: full_size:%d, scale_down_to:%g",
                width, (double)scale_down);
  if (scale_down <= 10.0f) return fail(STK_ERR_BAD_ARG, "scale_down_to was too small scale_down_to:%g", (double)scale_down);
  const double factor = (double)scale_down / (double)(width < height ? width : height);
  const int nw = (int)((double)width * factor), nh = (int)((double)height * factor);
  if (nw < 1 || nh < 1) return fail(STK_ERR_BAD_ARG, "scaled size %dx%d is empty", nw, nh);
  *sw = nw; *sh = nh;
  return STK_OK;
}

int stk_grey_resize_area(const uint8_t* img, size_t pitch, int width, int height, int channels, int out_width,
                         int out_height, int device, uint8_t* out, size_t out_pitch) {
  if (!img || !out) return fail(STK_ERR_BAD_ARG, "null argument");
  if (width <= 0 || height <= 0 || out_width <= 0 || out_height <= 0) return fail(STK_ERR_BAD_ARG, "bad size");
  if (channels != 1 && channels != 3 && channels != 4) return fail(STK_ERR_UNSUPPORTED, "channels must be 1, 3 or 4");
  const size_t row = (size_t)width * channels;
  if (pitch < row || out_pitch < (size_t)out_width) return fail(STK_ERR_BAD_ARG, "pitch too small");
  if (device >= 0) CU(cudaSetDevice(device));
  AreaPlan plan;
  int rc = make_area_plan(width, height, out_width, out_height, plan);
  if (rc) return rc;
  uint8_t *d_src = nullptr, *d_dst = nullptr;
  cudaError_t e = cudaMalloc((void**)&d_src, row * height);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_dst, (size_t)out_width * out_height);
  if (e == cudaSuccess) e = cudaMemcpy2D(d_src, row, img, pitch, row, height, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = launch_resize(plan, d_src, row, channels, d_dst, out_width, 0);
    if (rc == STK_OK) e = cudaMemcpy2D(out, out_pitch, d_dst, out_width, out_width, out_height, cudaMemcpyDeviceToHost);
  }
  cudaFree(d_src); cudaFree(d_dst);
  free_area_plan(plan);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(STK_ERR_CUDA, "grey_resize_area: %s", cudaGetErrorString(e));
  return STK_OK;
}

int stk_pinned_alloc(void** ptr, size_t bytes) {
  if (!ptr) return fail(STK_ERR_BAD_ARG, "null ptr");
  CU(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
  return STK_OK;
}
int stk_pinned_free(void* ptr) {
  CU(cudaFreeHost(ptr));
  return STK_OK;
}

int stk_ecc_create(const stk_ecc_config* cfg, stk_ecc_ctx** out) {
  if (!cfg || !out) return fail(STK_ERR_BAD_ARG, "null argument");
  *out = nullptr;
  if (cfg->width <= 0 || cfg->height <= 0) return fail(STK_ERR_BAD_ARG, "bad frame size %dx%d", cfg->width, cfg->height);
  if (cfg->width > 32766 || cfg->height > 32766) return fail(STK_ERR_BAD_ARG, "frame larger than OpenCV's remap limit (32767)");
  if (cfg->channels != 3 && cfg->channels != 4) return fail(STK_ERR_UNSUPPORTED, "channels must be 3 or 4 (got %d)", cfg->channels);
  if (cfg->align) {
    if (cfg->motion_type < 0 || cfg->motion_type > 3) return fail(STK_ERR_BAD_ARG, "bad motion type %d", cfg->motion_type);
    if (!(cfg->criteria_type & (STK_TERM_COUNT | STK_TERM_EPS)))
      return fail(STK_ERR_CRITERIA, "TermCriteria needs COUNT or EPS (findTransformECC asserts)");
    if (cfg->gauss_filt_size < 1 || cfg->gauss_filt_size % 2 == 0 || cfg->gauss_filt_size / 2 > stk::kMaxGaussRadius)
      return fail(STK_ERR_BAD_ARG, "gauss_filt_size must be odd, in [1, %d]", 2 * stk::kMaxGaussRadius + 1);
    if ((cfg->ecc_width != 0) != (cfg->ecc_height != 0) || cfg->ecc_width < 0 || cfg->ecc_height < 0)
      return fail(STK_ERR_BAD_ARG, "ecc_width/ecc_height must both be 0 or both be positive");
    // (an ECC size larger than the frame is legal: utils::scale_image enlarges a landscape frame when
    //  height < scale_down_width < width — the INTER_AREA resize then runs OpenCV's bilinear "area mode")
  }
  int dev = cfg->device;
  if (dev < 0) CU(cudaGetDevice(&dev));
  CU(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(STK_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);

  stk_ecc_ctx* c = new (std::nothrow) stk_ecc_ctx();
  if (!c) return fail(STK_ERR_NOMEM, "out of host memory");
  c->cfg = *cfg;
  c->device = dev;
  c->sm_count = prop.multiProcessorCount;
  c->n_lanes = cfg->lanes > 0 ? std::min(cfg->lanes, 16) : 4;
  c->scaled = cfg->align && cfg->ecc_width > 0;
  c->ew = c->scaled ? cfg->ecc_width : cfg->width;
  c->eh = c->scaled ? cfg->ecc_height : cfg->height;
  c->small_pitch = (c->ew + 15) / 16 * 16;
  c->pitch_f = (c->ew + 31) / 32 * 32;
  c->frame_bytes = (size_t)cfg->width * cfg->channels * cfg->height;
  c->acc_floats = (size_t)cfg->width * cfg->channels * cfg->height;
  c->max_iter = (cfg->criteria_type & STK_TERM_COUNT) ? cfg->max_count : 200;
  c->eps = (cfg->criteria_type & STK_TERM_EPS) ? cfg->epsilon : -1.0;
  const char* ec = getenv("STK_ECC_EXACT_COORDS");
  c->exact_coords = ec && strcmp(ec, "1") == 0;
  if (const char* pd = getenv("STK_ECC_PDL")) c->pdl = strcmp(pd, "0") != 0;
  if (const char* un = getenv("STK_ECC_UNROLL")) c->loop_unroll = std::max(1, std::min(16, atoi(un)));
  if (const char* rw = getenv("STK_ECC_RIM_WEIGHT")) c->rim_weight = std::max(8, std::min(32, atoi(rw)));
  if (const char* wb = getenv("STK_WARP_BATCH")) c->warp_batch = std::max(1, std::min(stk::kWarpBatch, atoi(wb)));
  if (const char* wg = getenv("STK_WARP_GEN")) c->warp_gen = atoi(wg) == 1 ? 1 : 2;
  const char* lm = getenv("STK_LOOP_MODE");
  c->host_loop = lm && strcmp(lm, "host") == 0;

  int rc = STK_OK;
  auto cleanup = [&](int code) { stk_ecc_destroy(c); return code; };

  if (cfg->align) {
    // work split: 128-column strips x row chunks, dealt evenly to one persistent block per resident slot
    if (const char* g = getenv("STK_ECC_GEN")) c->iter_gen = atoi(g) == 1 ? 1 : 2;
    if (const char* g = getenv("STK_ECC_CFG")) c->iter_cfg = std::max(-1, std::min(stk::kEccCfgCount - 1, atoi(g)));
    const IterVariant iv = iter_variant(cfg->motion_type, c->exact_coords, c->iter_gen, c->iter_cfg);
    c->iter_fn = iv.fn;
    c->iter_threads = iv.threads;
    c->iter_smem = iv.smem;
    c->iter_chunk_h = iv.chunk_h;
    c->iter_box_h = iv.box_h;
    c->iter_min_blocks = iv.min_blocks;
    if (cudaFuncSetAttribute(c->iter_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, c->iter_smem) != cudaSuccess)
      return cleanup(fail(STK_ERR_CUDA, "cannot reserve %d bytes of dynamic shared memory for the ECC kernel", c->iter_smem));
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, c->iter_fn, c->iter_threads, c->iter_smem) != cudaSuccess || occ < 1) occ = 1;
    // persistent blocks per SM and launch: the kernel's occupancy, unless STK_ECC_BLOCKS_PER_SM asks for fewer (with
    // several lanes the other slot is taken by another frame's kernel, and a block that owns twice the chunks pays its
    // start-up and its fold once for twice the work)
    // Measured (4K Homography, 13-frame stacks): four lanes 3 181 -> 3 430 frames/s with ONE block per SM and launch,
    // one lane 55.7 -> 64.1 us per iteration — so the grid follows the lane count of the context.
    if (const char* bps = getenv("STK_ECC_BLOCKS_PER_SM")) occ = std::max(1, std::min(occ, atoi(bps)));
    else if (c->n_lanes >= 3) occ = 1;
    const int slots = c->sm_count * occ;
    c->n_strips = (c->ew + stk::kEccStripW - 1) / stk::kEccStripW;
    c->chunks_per_strip = (c->eh + c->iter_chunk_h - 1) / c->iter_chunk_h;
    const long long total_chunks = (long long)c->n_strips * c->chunks_per_strip;
    // at least two chunks per block so the per-run fold/reduction stays amortised on small frames
    c->n_tiles = (int)std::max(1LL, std::min<long long>(slots, total_chunks / 2));
    c->nv = model_nv(cfg->motion_type);
    stk::PrepParams& pp = c->prep_proto;
    memset(&pp, 0, sizeof pp);
    pp.dst_pitch = c->pitch_f;
    pp.width = c->ew;
    pp.height = c->eh;
    pp.channels = cfg->channels;
    pp.radius = cfg->gauss_filt_size / 2;
    gaussian_taps(cfg->gauss_filt_size, pp.taps);
    const size_t smem = stk::prep_smem_bytes(pp.radius);
    if (smem > 48 * 1024) {
      if (cudaFuncSetAttribute(stk::prep_grey_blur_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return cleanup(fail(STK_ERR_CUDA, "cannot reserve %zu bytes of shared memory for the blur", smem));
    }
    if (cudaMalloc((void**)&c->img, (size_t)c->pitch_f * c->eh * sizeof(float)) != cudaSuccess)
      return cleanup(fail(STK_ERR_NOMEM, "cudaMalloc(I plane) failed"));
    rc = make_plane_tensor_map(&c->tm_img, c->img, c->ew, c->eh, c->pitch_f, stk::kBoxW, c->iter_box_h);
    if (rc) return cleanup(rc);
    if (c->scaled) {
      rc = make_area_plan(cfg->width, cfg->height, c->ew, c->eh, c->area);
      if (rc) return cleanup(rc);
    }

  }
  c->lanes.resize(c->n_lanes);
  if (cudaEventCreateWithFlags(&c->ref_ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->x_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->x_stream, cudaStreamNonBlocking) != cudaSuccess)
    return cleanup(fail(STK_ERR_CUDA, "cudaEventCreate / cudaStreamCreate failed"));
  for (auto& ln : c->lanes) {
    if (cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking) != cudaSuccess) return cleanup(fail(STK_ERR_CUDA, "cudaStreamCreate failed"));
    if (cudaEventCreateWithFlags(&ln.stage_free, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ln.drained, cudaEventDisableTiming) != cudaSuccess) return cleanup(fail(STK_ERR_CUDA, "cudaEventCreate failed"));
    if (cudaMalloc((void**)&ln.acc, c->acc_floats * sizeof(float)) != cudaSuccess) return cleanup(fail(STK_ERR_NOMEM, "cudaMalloc(accumulator) failed"));
    if (cfg->align) {
      if (cudaMalloc((void**)&ln.tmpl, (size_t)c->pitch_f * c->eh * sizeof(float)) != cudaSuccess) return cleanup(fail(STK_ERR_NOMEM, "cudaMalloc(T plane) failed"));
      rc = make_plane_tensor_map(&ln.tm_tmpl, ln.tmpl, c->ew, c->eh, c->pitch_f, stk::kEccStripW, c->iter_chunk_h);
      if (rc) return cleanup(rc);
      if (c->scaled && cudaMalloc((void**)&ln.small, (size_t)c->small_pitch * c->eh) != cudaSuccess) return cleanup(fail(STK_ERR_NOMEM, "cudaMalloc(downscaled grey) failed"));
      if (cudaMalloc((void**)&ln.st, sizeof(stk::EccState)) != cudaSuccess) return cleanup(fail(STK_ERR_NOMEM, "cudaMalloc(state) failed"));
      if (cudaMalloc((void**)&ln.pend_st, sizeof(stk::EccState) * stk::kWarpBatch) != cudaSuccess) return cleanup(fail(STK_ERR_NOMEM, "cudaMalloc(pending states) failed"));
      {
        // the rescale factors are the only fields the init kernel leaves alone
        stk::EccState zero;
        memset(&zero, 0, sizeof zero);
        if (c->scaled) {
          if (cfg->motion_type == STK_MOTION_HOMOGRAPHY) {     // src/utils.rs:228-241: f64 ratio, then `as f32`
            zero.rescale_x = (float)((double)cfg->width / (double)c->ew);
            zero.rescale_y = (float)((double)cfg->height / (double)c->eh);
          } else {                                             // src/lib.rs:946-949: f32 / f32
            zero.rescale_x = (float)cfg->width / (float)c->ew;
            zero.rescale_y = (float)cfg->height / (float)c->eh;
          }
        }
        if (cudaMemcpyAsync(ln.st, &zero, sizeof zero, cudaMemcpyHostToDevice, ln.stream) != cudaSuccess ||
            cudaStreamSynchronize(ln.stream) != cudaSuccess) return cleanup(fail(STK_ERR_CUDA, "state upload failed"));
      }
      if (cudaMalloc((void**)&ln.partials, (size_t)((c->n_tiles + 31) / 32 * 32) * c->nv * sizeof(double)) != cudaSuccess) return cleanup(fail(STK_ERR_NOMEM, "cudaMalloc(partials) failed"));
      if (!c->host_loop) {
        rc = build_lane_graph(c, ln);
        if (rc) return cleanup(rc);
      }
    }
  }
  *out = c;
  return STK_OK;
}

int stk_ecc_destroy(stk_ecc_ctx* c) {
  if (!c) return STK_OK;
  cudaSetDevice(c->device);
  for (auto& ln : c->lanes) {
    if (ln.stream) cudaStreamSynchronize(ln.stream);
    if (ln.exec) cudaGraphExecDestroy(ln.exec);
    if (ln.graph) cudaGraphDestroy(ln.graph);
    cudaFree(ln.tmpl); cudaFree(ln.small); cudaFree(ln.st); cudaFree(ln.pend_st); cudaFree(ln.partials); cudaFree(ln.acc);
    for (auto* d : ln.d_frames) cudaFree(d);
    if (ln.h_stage) cudaFreeHost(ln.h_stage);
    if (ln.h_cont) cudaFreeHost(ln.h_cont);
    if (ln.stage_free) cudaEventDestroy(ln.stage_free);
    if (ln.drained) cudaEventDestroy(ln.drained);
    if (ln.stream) cudaStreamDestroy(ln.stream);
  }
  if (c->x_stream) { cudaStreamSynchronize(c->x_stream); cudaStreamDestroy(c->x_stream); }
  if (c->ref_ready) cudaEventDestroy(c->ref_ready);
  if (c->input_ev) cudaEventDestroy(c->input_ev);
  if (c->x_done) cudaEventDestroy(c->x_done);
  for (auto& r : c->results) for (auto& e : r.ev) if (e) cudaEventDestroy(e);
  for (auto& b : c->ring) { if (b.host) cudaFreeHost(b.host); if (b.uploaded) cudaEventDestroy(b.uploaded); }
  for (void* m : c->peer.opened) cudaIpcCloseMemHandle(m);
  cudaFree(c->peer.flags);
  cudaFree(c->img); cudaFree(c->d_ref); cudaFree(c->d_out);
  free_area_plan(c->area);
  for (auto* p : c->state_chunks) cudaFreeHost(p);
  delete c;
  return STK_OK;
}

// a staging buffer for the NEXT frame submitted to this lane: one that holds no frame whose warp is still queued
// (at most kWarpBatch - 1 are, so one is always free)
static int ensure_host_staging(stk_ecc_ctx* c, Lane& ln, bool need_pinned_stage, uint8_t** d_frame) {
  int k = 0;
  while (k < stk::kWarpBatch - 1 && (ln.staging_used >> k) & 1u) ++k;
  uint8_t*& slot = ln.d_frames[k];
  if (!slot) CU(cudaMalloc((void**)&slot, c->frame_bytes));
  if (need_pinned_stage && !ln.h_stage) CU(cudaHostAlloc((void**)&ln.h_stage, c->frame_bytes, cudaHostAllocDefault));
  *d_frame = slot;
  return STK_OK;
}

// device-resident input: the lane that takes it waits for what its producer's stream has queued so far
static int order_after_input(stk_ecc_ctx* c, cudaStream_t lane_stream) {
  if (!c->order_inputs) return STK_OK;
  if (!c->input_ev) CU(cudaEventCreateWithFlags(&c->input_ev, cudaEventDisableTiming));
  CU(cudaEventRecord(c->input_ev, c->input_stream));
  CU(cudaStreamWaitEvent(lane_stream, c->input_ev, 0));
  return STK_OK;
}

// lane 0's stream waits until every lane has finished what is queued on it so far
static int join_lanes_on_lane0(stk_ecc_ctx* c) {
  Lane& l0 = c->lanes[0];
  for (auto& ln : c->lanes) {
    if (&ln == &l0) continue;
    CU(cudaEventRecord(ln.drained, ln.stream));
    CU(cudaStreamWaitEvent(l0.stream, ln.drained, 0));
  }
  return STK_OK;
}

// The reference plane is built on lane 0.  No host synchronisation: lane 0 first waits for whatever the other lanes still
// have queued (frames of an earlier stack read the plane it is about to overwrite), and every other lane then waits for
// `ref_ready` before its next frame.
static int set_reference_impl(stk_ecc_ctx* c, const uint8_t* d_bgr, size_t pitch) {
  Lane& l0 = c->lanes[0];
  int rc = join_lanes_on_lane0(c);
  if (rc) return rc;
  if (c->cfg.align) {
    rc = launch_prep(c, d_bgr, pitch, l0.small, c->img, l0.stream);
    if (rc) return rc;
  }
  CU(cudaEventRecord(c->ref_ready, l0.stream));
  for (auto& ln : c->lanes) if (&ln != &l0) CU(cudaStreamWaitEvent(ln.stream, c->ref_ready, 0));
  if (c->cfg.seed_reference) {
    if (l0.wait_x) {       // the previous stack's exchange may still read lane 0's accumulator
      CU(cudaStreamWaitEvent(l0.stream, c->x_done, 0));
      l0.wait_x = false;
    }
    const int row = c->cfg.width * c->cfg.channels;
    dim3 grid((row + 1023) / 1024, c->cfg.height);
    // 4-byte loads / 16-byte stores need the rows of both buffers to start on those boundaries
    const int vec_ok = (row % 4 == 0) && (pitch % 4 == 0) && (((uintptr_t)d_bgr) % 4 == 0);
    stk::seed_accumulator_kernel<<<grid, 256, 0, l0.stream>>>(d_bgr, pitch, l0.acc, row, c->cfg.height, vec_ok);
    c->launches++;
    CU(cudaGetLastError());
    l0.acc_used = true;
  }
  c->have_ref = true;
  return STK_OK;
}

int stk_ecc_set_reference(stk_ecc_ctx* c, const uint8_t* bgr, size_t pitch) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!bgr) return fail(STK_ERR_BAD_ARG, "null frame");
  const size_t row = (size_t)c->cfg.width * c->cfg.channels;
  if (pitch < row) return fail(STK_ERR_BAD_ARG, "pitch %zu < row bytes %zu", pitch, row);
  std::lock_guard<std::mutex> g(c->mu);
  if (!c->d_ref) CU(cudaMalloc((void**)&c->d_ref, c->frame_bytes));
  // stream-ordered on lane 0: a blocking cudaMemcpy from pageable memory may return before the DMA has
  // landed and is ordered only against the legacy stream, which the (non-blocking) lane streams ignore
  CU(cudaMemcpy2DAsync(c->d_ref, row, bgr, pitch, row, c->cfg.height, cudaMemcpyHostToDevice, c->lanes[0].stream));
  return set_reference_impl(c, c->d_ref, row);
}

int stk_ecc_set_reference_device(stk_ecc_ctx* c, const uint8_t* d_bgr, size_t pitch) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!d_bgr) return fail(STK_ERR_BAD_ARG, "null frame");
  if (pitch < (size_t)c->cfg.width * c->cfg.channels) return fail(STK_ERR_BAD_ARG, "pitch too small");
  std::lock_guard<std::mutex> g(c->mu);
  rc = order_after_input(c, c->lanes[0].stream);
  if (rc) return rc;
  return set_reference_impl(c, d_bgr, pitch);
}

int stk_ecc_set_input_stream(stk_ecc_ctx* c, void* cuda_stream, int enabled) {
  int rc = check_ctx(c);
  if (rc) return rc;
  std::lock_guard<std::mutex> g(c->mu);
  c->order_inputs = enabled != 0;
  c->input_stream = enabled ? (cudaStream_t)cuda_stream : nullptr;
  return STK_OK;
}

// a filled ring buffer: asynchronous upload on the next lane, then align (inv == null) or warp-only; the
// buffer goes back to the ring at once, guarded by its event
static int submit_ring(stk_ecc_ctx* c, int idx, int64_t tag, const double* inv, const float* border, bool persp = true) {
  int rc = STK_OK;
  {
    std::lock_guard<std::mutex> g(c->mu);
    Lane& ln = pick_lane(c);
    RingBuf& rb = c->ring[idx];
    const size_t row = (size_t)c->cfg.width * c->cfg.channels;
    uint8_t* d_frame = nullptr;
    rc = ensure_host_staging(c, ln, false, &d_frame);
    cudaError_t e = cudaSuccess;
    if (rc == STK_OK) e = cudaMemcpyAsync(d_frame, rb.host, c->frame_bytes, cudaMemcpyHostToDevice, ln.stream);
    if (rc == STK_OK && e == cudaSuccess) e = cudaEventRecord(rb.uploaded, ln.stream);
    if (rc == STK_OK && e != cudaSuccess) rc = fail(STK_ERR_CUDA, "frame upload: %s", cudaGetErrorString(e));
    if (rc == STK_OK) {
      if (inv) {
        rc = launch_warp(c, ln, d_frame, row, persp, inv, border, false);
        if (rc == STK_OK) { ResultSlot slot; slot.tag = tag; slot.host = nullptr; c->results.push_back(slot); }
      } else {
        rc = enqueue_align(c, ln, d_frame, row, tag);
      }
    }
  }
  ring_release(c, idx);
  return rc;
}

static int submit_align(stk_ecc_ctx* c, const uint8_t* buf, size_t pitch, int64_t tag, int kind) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!buf) return fail(STK_ERR_BAD_ARG, "null frame");
  if (!c->cfg.align) return fail(STK_ERR_STATE, "context was created with align = 0");
  const size_t row = (size_t)c->cfg.width * c->cfg.channels;
  if (pitch < row) return fail(STK_ERR_BAD_ARG, "pitch %zu < row bytes %zu", pitch, row);
  if (kind == 0) {
    // pageable buffer: one copy into a ring buffer, made without the context lock (decode threads copy in
    // parallel), then the zero-copy path
    { std::lock_guard<std::mutex> g(c->mu); if (!c->have_ref) return fail(STK_ERR_STATE, "stk_ecc_set_reference must come first"); }
    int idx = -1;
    rc = ring_acquire(c, &idx);
    if (rc) return rc;
    copy_rows(c->ring[idx].host, buf, pitch, row, c->cfg.height);
    return submit_ring(c, idx, tag, nullptr, nullptr);
  }
  std::lock_guard<std::mutex> g(c->mu);
  if (!c->have_ref) return fail(STK_ERR_STATE, "stk_ecc_set_reference must come first");
  Lane& ln = pick_lane(c);
  if (kind == 2) {
    rc = order_after_input(c, ln.stream);
    if (rc) return rc;
    return enqueue_align(c, ln, buf, pitch, tag);
  }
  uint8_t* d_frame = nullptr;
  rc = ensure_host_staging(c, ln, false, &d_frame);
  if (rc) return rc;
  rc = upload_frame(c, ln, d_frame, buf, pitch, true);
  if (rc) return rc;
  return enqueue_align(c, ln, d_frame, row, tag);
}

int stk_ecc_submit_frame(stk_ecc_ctx* c, const uint8_t* bgr, size_t pitch, int64_t tag) { return submit_align(c, bgr, pitch, tag, 0); }
int stk_ecc_submit_frame_pinned(stk_ecc_ctx* c, const uint8_t* bgr, size_t pitch, int64_t tag) { return submit_align(c, bgr, pitch, tag, 1); }
int stk_ecc_submit_frame_device(stk_ecc_ctx* c, const uint8_t* d_bgr, size_t pitch, int64_t tag) { return submit_align(c, d_bgr, pitch, tag, 2); }

int stk_ecc_acquire_frame_buffer(stk_ecc_ctx* c, uint8_t** buf, size_t* pitch) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!buf) return fail(STK_ERR_BAD_ARG, "null argument");
  int idx = -1;
  rc = ring_acquire(c, &idx);
  if (rc) return rc;
  *buf = c->ring[idx].host;
  if (pitch) *pitch = (size_t)c->cfg.width * c->cfg.channels;
  return STK_OK;
}

int stk_ecc_release_frame_buffer(stk_ecc_ctx* c, uint8_t* buf) {
  if (!c || !buf) return fail(STK_ERR_BAD_ARG, "null argument");
  const int idx = ring_index_of(c, buf);
  if (idx < 0) return fail(STK_ERR_BAD_ARG, "not a buffer of this context's frame ring");
  ring_release(c, idx);
  return STK_OK;
}

int stk_ecc_submit_acquired(stk_ecc_ctx* c, uint8_t* buf, int64_t tag) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!buf) return fail(STK_ERR_BAD_ARG, "null frame");
  if (!c->cfg.align) return fail(STK_ERR_STATE, "context was created with align = 0");
  const int idx = ring_index_of(c, buf);
  if (idx < 0) return fail(STK_ERR_BAD_ARG, "not a buffer of this context's frame ring");
  { std::lock_guard<std::mutex> g(c->mu); if (!c->have_ref) { ring_release(c, idx); return fail(STK_ERR_STATE, "stk_ecc_set_reference must come first"); } }
  return submit_ring(c, idx, tag, nullptr, nullptr);
}

static int submit_warp(stk_ecc_ctx* c, const uint8_t* buf, size_t pitch, const double* h, int border_mode,
                       const double* border_value, int64_t tag, bool device, bool affine = false) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!buf || !h) return fail(STK_ERR_BAD_ARG, "null argument");
  if (border_mode != STK_BORDER_CONSTANT && border_mode != STK_BORDER_REPLICATE && border_mode != STK_BORDER_REFLECT &&
      border_mode != STK_BORDER_WRAP && border_mode != STK_BORDER_REFLECT_101)
    return fail(STK_ERR_UNSUPPORTED, "border mode %d: CONSTANT, REPLICATE, REFLECT, WRAP and REFLECT_101 are implemented "
                "(BORDER_TRANSPARENT would add uninitialised memory to the stack in the reference)", border_mode);
  const size_t row = (size_t)c->cfg.width * c->cfg.channels;
  if (pitch < row) return fail(STK_ERR_BAD_ARG, "pitch %zu < row bytes %zu", pitch, row);
  double inv[9];
  if (affine) invert_affine_host(h, inv);
  else invert_perspective_host(h, inv);
  float border[5] = {0, 0, 0, 0, (float)border_mode};
  if (border_value) for (int i = 0; i < 4; ++i) border[i] = (float)border_value[i];
  if (!device) {
    int idx = -1;
    rc = ring_acquire(c, &idx);
    if (rc) return rc;
    copy_rows(c->ring[idx].host, buf, pitch, row, c->cfg.height);
    return submit_ring(c, idx, tag, inv, border, !affine);
  }
  std::lock_guard<std::mutex> g(c->mu);
  Lane& ln = pick_lane(c);
  const uint8_t* d_src = buf;
  size_t d_pitch = pitch;
  rc = order_after_input(c, ln.stream);
  if (rc) return rc;
  rc = launch_warp(c, ln, d_src, d_pitch, !affine, inv, border, false);
  if (rc) return rc;
  { ResultSlot slot; slot.tag = tag; slot.host = nullptr; c->results.push_back(slot); }
  return STK_OK;
}

int stk_ecc_submit_warp(stk_ecc_ctx* c, const uint8_t* bgr, size_t pitch, const double h[9], int border_mode,
                        const double border_value[4], int64_t tag) {
  return submit_warp(c, bgr, pitch, h, border_mode, border_value, tag, false);
}
int stk_ecc_submit_warp_device(stk_ecc_ctx* c, const uint8_t* d_bgr, size_t pitch, const double h[9], int border_mode,
                               const double border_value[4], int64_t tag) {
  return submit_warp(c, d_bgr, pitch, h, border_mode, border_value, tag, true);
}

int stk_ecc_submit_warp_affine(stk_ecc_ctx* c, const uint8_t* bgr, size_t pitch, const double m[6], int border_mode,
                               const double border_value[4], int64_t tag) {
  return submit_warp(c, bgr, pitch, m, border_mode, border_value, tag, false, true);
}
int stk_ecc_submit_warp_affine_device(stk_ecc_ctx* c, const uint8_t* d_bgr, size_t pitch, const double m[6], int border_mode,
                                      const double border_value[4], int64_t tag) {
  return submit_warp(c, d_bgr, pitch, m, border_mode, border_value, tag, true, true);
}

int stk_ecc_sync(stk_ecc_ctx* c) {
  int rc = check_ctx(c);
  if (rc) return rc;
  std::lock_guard<std::mutex> g(c->mu);
  return sync_all(c);
}

int stk_ecc_results(stk_ecc_ctx* c, stk_frame_result* out, int capacity, int* count) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!count || (capacity > 0 && !out)) return fail(STK_ERR_BAD_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  if ((rc = flush_all(c))) return rc;
  for (auto& ln : c->lanes) CU(cudaStreamSynchronize(ln.stream));
  int n = 0;
  for (auto& r : c->results) {
    if (n >= capacity) break;
    stk_frame_result& o = out[n++];
    memset(&o, 0, sizeof o);
    o.tag = r.tag;
    if (r.host) {
      for (int i = 0; i < 9; ++i) o.warp[i] = r.host->m[i];
      o.rho = r.host->rho;
      o.iterations = r.host->iters;
      o.status = r.host->status;
    } else {
      o.warp[0] = o.warp[4] = o.warp[8] = 1.f;
    }
  }
  *count = n;
  return STK_OK;
}

int stk_ecc_partial(stk_ecc_ctx* c, float** d_partial, size_t* n_floats) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!d_partial) return fail(STK_ERR_BAD_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  rc = sync_all(c);
  if (rc) return rc;
  const float* used[16];
  int n = 0;
  for (auto& ln : c->lanes) if (ln.acc_used) used[n++] = ln.acc;
  Lane& l0 = c->lanes[0];
  if (!(n == 1 && used[0] == l0.acc)) {
    // lane 0 first so that `out` aliasing lanes[0] is read before it is written by the same thread
    const float* ordered[16];
    int m = 0;
    if (l0.acc_used) ordered[m++] = l0.acc;
    for (auto& ln : c->lanes) if (ln.acc_used && ln.acc != l0.acc) ordered[m++] = ln.acc;
    rc = lane_sum(c, l0.acc, ordered, m, false, 1, l0.stream);
    if (rc) return rc;
    CU(cudaStreamSynchronize(l0.stream));
  }
  for (auto& ln : c->lanes) ln.acc_used = false;
  l0.acc_used = true;
  *d_partial = l0.acc;
  if (n_floats) *n_floats = c->acc_floats;
  return STK_OK;
}

int stk_ecc_finish_device(stk_ecc_ctx* c, const float* d_sum, int divisor, float* d_out) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (divisor <= 0) return fail(STK_ERR_BAD_ARG, "divisor must be positive (got %d)", divisor);
  if (!d_out) return fail(STK_ERR_BAD_ARG, "null output");
  std::lock_guard<std::mutex> g(c->mu);
  const float* src = d_sum ? d_sum : c->lanes[0].acc;
  rc = lane_sum(c, d_out, &src, 1, true, divisor, c->lanes[0].stream);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->lanes[0].stream));
  return STK_OK;
}

int stk_ecc_finish_from(stk_ecc_ctx* c, const float* d_sum, int divisor, float* out, size_t out_pitch) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!out) return fail(STK_ERR_BAD_ARG, "null output");
  const size_t row = (size_t)c->cfg.width * c->cfg.channels * sizeof(float);
  if (out_pitch < row) return fail(STK_ERR_BAD_ARG, "out_pitch %zu < row bytes %zu", out_pitch, row);
  {
    std::lock_guard<std::mutex> g(c->mu);
    if (!c->d_out) CU(cudaMalloc((void**)&c->d_out, c->acc_floats * sizeof(float)));
  }
  rc = stk_ecc_finish_device(c, d_sum, divisor, c->d_out);
  if (rc) return rc;
  CU(cudaMemcpy2D(out, out_pitch, c->d_out, row, row, c->cfg.height, cudaMemcpyDeviceToHost));
  return STK_OK;
}

int stk_ecc_finish(stk_ecc_ctx* c, int divisor, float* out, size_t out_pitch) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (divisor <= 0) return fail(STK_ERR_BAD_ARG, "divisor must be positive (got %d)", divisor);
  if (!out) return fail(STK_ERR_BAD_ARG, "null output");
  const size_t row = (size_t)c->cfg.width * c->cfg.channels * sizeof(float);
  if (out_pitch < row) return fail(STK_ERR_BAD_ARG, "out_pitch %zu < row bytes %zu", out_pitch, row);
  {
    std::lock_guard<std::mutex> g(c->mu);
    rc = sync_all(c);
    if (rc) return rc;
    if (!c->d_out) CU(cudaMalloc((void**)&c->d_out, c->acc_floats * sizeof(float)));
    const float* used[16];
    int n = 0;
    for (auto& ln : c->lanes) if (ln.acc_used) used[n++] = ln.acc;
    rc = lane_sum(c, c->d_out, used, n, true, divisor, c->lanes[0].stream);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->lanes[0].stream));
  }
  CU(cudaMemcpy2D(out, out_pitch, c->d_out, row, row, c->cfg.height, cudaMemcpyDeviceToHost));
  return STK_OK;
}


/* ---- multi-GPU exchange over peer memory (csrc/peer_reduce.cuh) ------------------------------------------- */
}  // extern "C"
namespace {

int peer_prepare(stk_ecc_ctx* c) {
  if (!c->peer.flags) {
    CU(cudaMalloc((void**)&c->peer.flags, stk::kPeerFlagWords * sizeof(uint32_t)));
    CU(cudaMemset(c->peer.flags, 0, stk::kPeerFlagWords * sizeof(uint32_t)));
  }
  if (!c->d_out) CU(cudaMalloc((void**)&c->d_out, c->acc_floats * sizeof(float)));
  return STK_OK;
}

void peer_clear(stk_ecc_ctx* c) {
  for (void* m : c->peer.opened) cudaIpcCloseMemHandle(m);
  c->peer.opened.clear();
  c->peer.connected = false;
  c->peer.world = 0;
}

struct PeerHandleWire {          // what stk_peer_handle carries
  cudaIpcMemHandle_t partial, out, flags;
  uint64_t n_floats;
  int32_t device;
  int32_t abi;
};
static_assert(sizeof(PeerHandleWire) <= sizeof(stk_peer_handle), "stk_peer_handle too small");

template <int W>
void launch_peer_reduce(const stk::PeerReduceParams& p, int blocks, cudaStream_t s) {
  stk::peer_reduce_scale_kernel<W><<<blocks, 256, 0, s>>>(p);
}

}  // namespace
extern "C" {

int stk_ecc_peer_export(stk_ecc_ctx* c, stk_peer_handle* out) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!out) return fail(STK_ERR_BAD_ARG, "null handle");
  std::lock_guard<std::mutex> g(c->mu);
  if ((rc = peer_prepare(c))) return rc;
  PeerHandleWire w;
  memset(&w, 0, sizeof w);
  CU(cudaIpcGetMemHandle(&w.partial, c->lanes[0].acc));
  CU(cudaIpcGetMemHandle(&w.out, c->d_out));
  CU(cudaIpcGetMemHandle(&w.flags, c->peer.flags));
  w.n_floats = c->acc_floats;
  w.device = c->device;
  w.abi = STK_ABI_VERSION;
  memset(out, 0, sizeof *out);
  memcpy(out, &w, sizeof w);
  return STK_OK;
}

int stk_ecc_peer_connect(stk_ecc_ctx* c, int rank, int world, const stk_peer_handle* handles) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!handles) return fail(STK_ERR_BAD_ARG, "null handles");
  if (world < 1 || world > stk::kMaxPeers || rank < 0 || rank >= world)
    return fail(STK_ERR_BAD_ARG, "rank %d / world %d out of range (at most %d ranks)", rank, world, stk::kMaxPeers);
  std::lock_guard<std::mutex> g(c->mu);
  if ((rc = peer_prepare(c))) return rc;
  peer_clear(c);
  PeerLink& pl = c->peer;
  for (int r = 0; r < world; ++r) {
    PeerHandleWire w;
    memcpy(&w, &handles[r], sizeof w);
    if (w.abi != STK_ABI_VERSION || w.n_floats != c->acc_floats) {
      peer_clear(c);
      return fail(STK_ERR_BAD_ARG, "peer %d exported a different stack geometry or ABI", r);
    }
    if (r == rank) {
      pl.partial[r] = c->lanes[0].acc;
      pl.pflags[r] = pl.flags;
      if (r == 0) pl.root_out = c->d_out;
      continue;
    }
    void *mp = nullptr, *mf = nullptr, *mo = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&mp, w.partial, cudaIpcMemLazyEnablePeerAccess);
    if (e == cudaSuccess) { pl.opened.push_back(mp); e = cudaIpcOpenMemHandle(&mf, w.flags, cudaIpcMemLazyEnablePeerAccess); }
    if (e == cudaSuccess) { pl.opened.push_back(mf); if (r == 0) { e = cudaIpcOpenMemHandle(&mo, w.out, cudaIpcMemLazyEnablePeerAccess); if (e == cudaSuccess) pl.opened.push_back(mo); } }
    if (e != cudaSuccess) {
      cudaGetLastError();
      peer_clear(c);
      return fail(STK_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d, device %d) failed: %s", r, w.device, cudaGetErrorString(e));
    }
    pl.partial[r] = (const float*)mp;
    pl.pflags[r] = (uint32_t*)mf;
    if (r == 0) pl.root_out = (float*)mo;
  }
  pl.rank = rank;
  pl.world = world;
  pl.connected = true;
  // a (re)connect starts the step count afresh on EVERY rank (connect is collective): a context that already
  // exchanged in another group would otherwise disagree with a fresh one and spin into the timeout
  pl.step = 0;
  CU(cudaMemset(pl.flags, 0, stk::kPeerFlagWords * sizeof(uint32_t)));
  return STK_OK;
}

int stk_ecc_peer_connect_local(stk_ecc_ctx* const* ctxs, int world) {
  if (!ctxs) return fail(STK_ERR_BAD_ARG, "null contexts");
  if (world < 1 || world > stk::kMaxPeers) return fail(STK_ERR_BAD_ARG, "world %d out of range (at most %d ranks)", world, stk::kMaxPeers);
  for (int r = 0; r < world; ++r) {
    if (!ctxs[r]) return fail(STK_ERR_BAD_ARG, "null context %d", r);
    if (ctxs[r]->acc_floats != ctxs[0]->acc_floats) return fail(STK_ERR_BAD_ARG, "context %d has a different stack geometry", r);
    for (int q = 0; q < r; ++q)
      if (ctxs[q]->device == ctxs[r]->device) return fail(STK_ERR_BAD_ARG, "contexts %d and %d share device %d", q, r, ctxs[r]->device);
  }
  for (int r = 0; r < world; ++r) {
    stk_ecc_ctx* c = ctxs[r];
    CU(cudaSetDevice(c->device));
    std::lock_guard<std::mutex> g(c->mu);
    int rc = peer_prepare(c);
    if (rc) return rc;
    peer_clear(c);
    for (int q = 0; q < world; ++q) {
      if (q == r) continue;
      int can = 0;
      CU(cudaDeviceCanAccessPeer(&can, c->device, ctxs[q]->device));
      if (!can) return fail(STK_ERR_UNSUPPORTED, "device %d cannot access device %d's memory", c->device, ctxs[q]->device);
      cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) return fail(STK_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", c->device, ctxs[q]->device, cudaGetErrorString(e));
    }
  }
  for (int r = 0; r < world; ++r) {
    PeerLink& pl = ctxs[r]->peer;
    for (int q = 0; q < world; ++q) { pl.partial[q] = ctxs[q]->lanes[0].acc; pl.pflags[q] = ctxs[q]->peer.flags; }
    pl.root_out = ctxs[0]->d_out;
    pl.rank = r;
    pl.world = world;
    pl.connected = true;
    pl.step = 0;
    CU(cudaSetDevice(ctxs[r]->device));
    CU(cudaMemset(pl.flags, 0, stk::kPeerFlagWords * sizeof(uint32_t)));
  }
  return STK_OK;
}

int stk_ecc_peer_disconnect(stk_ecc_ctx* c) {
  int rc = check_ctx(c);
  if (rc) return rc;
  std::lock_guard<std::mutex> g(c->mu);
  for (auto& ln : c->lanes) CU(cudaStreamSynchronize(ln.stream));
  CU(cudaStreamSynchronize(c->x_stream));
  peer_clear(c);
  return STK_OK;
}

}  // extern "C"
namespace {
// scatter == false: finished pixels go to the root's output buffer (reduce);  true: every rank keeps the finished
// pixels of its own slice in its own output buffer (reduce-scatter), to be copied out by stk_ecc_peer_slice_to_host
int peer_exchange(stk_ecc_ctx* c, int divisor, bool scatter) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (divisor <= 0) return fail(STK_ERR_BAD_ARG, "divisor must be positive (got %d)", divisor);
  std::lock_guard<std::mutex> g(c->mu);
  PeerLink& pl = c->peer;
  if (!pl.connected) return fail(STK_ERR_STATE, "stk_ecc_peer_reduce before stk_ecc_peer_connect");
  if ((rc = flush_all(c))) return rc;
  Lane& l0 = c->lanes[0];
  cudaStream_t xs = c->x_stream;
  // join the lanes on the exchange stream (no host synchronisation), then sum them into lane 0's accumulator.  The lanes
  // themselves stay free: the next stack's prep and ECC iterations run while this exchange is in flight; only its first
  // accumulator writes wait for x_done (Lane::wait_x).
  const float* ordered[16];
  int m = 0;
  if (l0.acc_used) ordered[m++] = l0.acc;
  for (auto& ln : c->lanes) {
    CU(cudaEventRecord(ln.drained, ln.stream));
    CU(cudaStreamWaitEvent(xs, ln.drained, 0));
    if (&ln != &l0 && ln.acc_used) ordered[m++] = ln.acc;
  }
  if (pl.world == 1) {
    // a world of one: the "exchange" is the lane sum fused with the divide (what stk_ecc_finish_device does), but on the
    // exchange stream and without a host synchronisation — consecutive stacks of ONE device queue back to back as well
    rc = lane_sum(c, scatter ? c->d_out : pl.root_out, ordered, m, true, divisor, xs);
    if (rc) return rc;
    CU(cudaEventRecord(c->x_done, xs));
    for (auto& ln : c->lanes) { ln.acc_used = false; ln.wait_x = true; }
    ++pl.step;
    pl.slice_begin = 0;
    pl.slice_end = c->acc_floats;
    pl.scattered = scatter;
    return STK_OK;
  }
  if (!(m == 1 && ordered[0] == l0.acc)) {
    rc = lane_sum(c, l0.acc, ordered, m, false, 1, xs);
    if (rc) return rc;
  }
  for (auto& ln : c->lanes) { ln.acc_used = false; ln.wait_x = true; }
  l0.acc_used = true;

  stk::PeerReduceParams p;
  memset(&p, 0, sizeof p);
  for (int r = 0; r < pl.world; ++r) { p.partial[r] = pl.partial[r]; p.flags[r] = pl.pflags[r]; }
  p.out = scatter ? c->d_out : pl.root_out;
  // Slices in units of 4 floats, the last worker also takes the remainder.  With more than two ranks the ROOT
  // takes no slice: every finished pixel has to enter the root over its inbound links anyway (S bytes), and a
  // root slice would add (world-1) remote reads per pixel on those same links (measured at world 4: 260 us with
  // equal slices, the root's inbound side carrying 1.5 S).  Without it every rank's inbound traffic is S.
  {
    // (reduce-scatter: nothing converges on the root, every rank takes an equal slice)
    const bool rootless = pl.world > 2 && !scatter;
    const int workers = rootless ? pl.world - 1 : pl.world;
    const int w = rootless ? pl.rank - 1 : pl.rank;              // -1: the root of a world > 2
    const size_t n4 = c->acc_floats / 4, per = n4 / workers;
    if (w < 0) { p.begin = p.end = 0; }
    else {
      p.begin = (size_t)w * per * 4;
      p.end = w == workers - 1 ? c->acc_floats : (size_t)(w + 1) * per * 4;
    }
  }
  p.rank = pl.rank;
  p.world = pl.world;
  p.step = ++pl.step;
  p.scale = (float)(1.0 / (double)divisor);
  static const unsigned long long timeout_ms = [] { const char* e = getenv("STK_PEER_TIMEOUT_MS"); return e ? strtoull(e, nullptr, 10) : 30000ull; }();
  p.timeout_ns = timeout_ms * 1000000ull;
  // Grid of the reduce kernel (NVLink-bound).  Measured at world 2 on the 4K payload (scripts/peer_check.py,
  // profiles/r2_peer_blocks.log): 592 blocks 191 us, 148 blocks 191 us, 74 blocks 217 us, 37 blocks 384 us, 16 blocks 832 us —
  // one block per SM is as fast as four and leaves the other block slot of every SM to the next stack's ECC kernels, which
  // run beside the exchange.  STK_PEER_BLOCKS overrides.
  static const int blocks_env = [] { const char* e = getenv("STK_PEER_BLOCKS"); return e ? atoi(e) : 0; }();
  const int blocks = blocks_env > 0 ? blocks_env : c->sm_count;
  // announce + wait in a one-warp kernel of its own: a rank that is ahead of the others spins there without holding SM
  // slots, and the reduce kernel only starts when every partial is complete
  p.pre_waited = 1;
  stk::peer_announce_wait_kernel<<<1, 32, 0, xs>>>(p);
  CU(cudaGetLastError());
  switch (pl.world) {
    case 2: launch_peer_reduce<2>(p, blocks, xs); break;
    case 4: launch_peer_reduce<4>(p, blocks, xs); break;
    case 8: launch_peer_reduce<8>(p, blocks, xs); break;
    default: launch_peer_reduce<0>(p, blocks, xs); break;
  }
  CU(cudaGetLastError());
  stk::peer_wait_done_kernel<<<1, 32, 0, xs>>>(pl.flags, pl.world, p.step, p.timeout_ns);
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->x_done, xs));
  c->launches += 1;
  c->launches += 2;
  pl.slice_begin = p.begin;
  pl.slice_end = p.end;
  pl.scattered = scatter;
  return STK_OK;
}
}  // namespace
extern "C" {

int stk_ecc_peer_reduce(stk_ecc_ctx* c, int divisor, const float** d_out) {
  int rc = peer_exchange(c, divisor, false);
  if (rc) return rc;
  if (d_out) *d_out = c->peer.rank == 0 ? c->d_out : nullptr;
  return STK_OK;
}

int stk_ecc_peer_reduce_scatter(stk_ecc_ctx* c, int divisor, const float** d_slice, size_t* begin, size_t* count) {
  int rc = peer_exchange(c, divisor, true);
  if (rc) return rc;
  if (d_slice) *d_slice = c->d_out + c->peer.slice_begin;
  if (begin) *begin = c->peer.slice_begin;
  if (count) *count = c->peer.slice_end - c->peer.slice_begin;
  return STK_OK;
}

int stk_ecc_peer_slice_to_host(stk_ecc_ctx* c, float* out) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!out) return fail(STK_ERR_BAD_ARG, "null output");
  std::lock_guard<std::mutex> g(c->mu);
  PeerLink& pl = c->peer;
  if (!pl.connected || !pl.scattered) return fail(STK_ERR_STATE, "stk_ecc_peer_slice_to_host needs a preceding stk_ecc_peer_reduce_scatter");
  const size_t n = pl.slice_end - pl.slice_begin;
  if (n) CU(cudaMemcpyAsync(out + pl.slice_begin, c->d_out + pl.slice_begin, n * sizeof(float), cudaMemcpyDeviceToHost, c->x_stream));
  return STK_OK;
}

int stk_ecc_reset(stk_ecc_ctx* c) {
  int rc = check_ctx(c);
  if (rc) return rc;
  std::lock_guard<std::mutex> g(c->mu);
  // no host synchronisation: the next stk_ecc_set_reference joins the lanes on the device (the pinned result records and
  // staging buffers are only ever touched in stream order behind that join)
  for (auto& ln : c->lanes) { ln.acc_used = false; ln.n_pend = 0; ln.staging_used = 0; }
  for (auto& r : c->results) for (auto& e : r.ev) if (e) cudaEventDestroy(e);
  c->results.clear();
  c->states_used = 0;
  c->iter_launches_counted = 0;
  c->launches = 0;
  c->next_lane = 0;
  c->have_ref = false;
  return STK_OK;
}

int stk_ecc_launch_count(stk_ecc_ctx* c, int64_t* launches) {
  if (!c || !launches) return fail(STK_ERR_BAD_ARG, "null argument");
  *launches = c->launches.load();
  return STK_OK;
}

int stk_ecc_set_profiling(stk_ecc_ctx* c, int enabled) {
  if (!c) return fail(STK_ERR_BAD_ARG, "null context");
  std::lock_guard<std::mutex> g(c->mu);
  c->profiling = enabled != 0;
  return STK_OK;
}

int stk_ecc_stage_times(stk_ecc_ctx* c, double ms[3], int64_t* frames, int64_t* iterations) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!ms) return fail(STK_ERR_BAD_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  if ((rc = flush_all(c))) return rc;
  for (auto& ln : c->lanes) CU(cudaStreamSynchronize(ln.stream));
  ms[0] = ms[1] = ms[2] = 0.0;
  int64_t nf = 0, it = 0;
  for (auto& r : c->results) {
    if (!r.ev[0] || !r.ev[3]) continue;
    for (int k = 0; k < 3; ++k) {
      float t = 0.f;
      CU(cudaEventElapsedTime(&t, r.ev[k], r.ev[k + 1]));
      ms[k] += t;
    }
    ++nf;
    if (r.host) it += r.host->iters;
  }
  if (frames) *frames = nf;
  if (iterations) *iterations = it;
  return STK_OK;
}

/* ---- single stages ----------------------------------------------------------------------------------- */
int stk_prep_grey_blur(const uint8_t* bgr, size_t pitch, int width, int height, int channels, int ksize,
                       int device, float* out, size_t out_pitch) {
  if (!bgr || !out) return fail(STK_ERR_BAD_ARG, "null argument");
  stk_ecc_config cfg = {};
  cfg.width = width; cfg.height = height; cfg.channels = channels;
  cfg.motion_type = STK_MOTION_TRANSLATION; cfg.criteria_type = STK_TERM_COUNT; cfg.max_count = 1;
  cfg.gauss_filt_size = ksize; cfg.device = device; cfg.lanes = 1; cfg.seed_reference = 0; cfg.align = 1;
  if (out_pitch < (size_t)width * sizeof(float)) return fail(STK_ERR_BAD_ARG, "out_pitch too small");
  stk_ecc_ctx* c = nullptr;
  int rc = stk_ecc_create(&cfg, &c);
  if (rc) return rc;
  rc = stk_ecc_set_reference(c, bgr, pitch);
  if (rc == STK_OK && cudaStreamSynchronize(c->lanes[0].stream) != cudaSuccess)     // set_reference is asynchronous
    rc = fail(STK_ERR_CUDA, "prep failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc == STK_OK) {
    cudaError_t e = cudaMemcpy2D(out, out_pitch, c->img, (size_t)c->pitch_f * sizeof(float),
                                 (size_t)width * sizeof(float), height, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(STK_ERR_CUDA, "prep readback: %s", cudaGetErrorString(e));
  }
  stk_ecc_destroy(c);
  return rc;
}

int stk_ecc_debug_iteration(stk_ecc_ctx* c, const uint8_t* bgr, size_t pitch, const float warp_in[9],
                            double* totals, int cap, int* nv, float warp_out[9], double* rho, int* status) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!bgr || !warp_in || !totals || !nv) return fail(STK_ERR_BAD_ARG, "null argument");
  if (!c->cfg.align) return fail(STK_ERR_STATE, "context was created with align = 0");
  if (cap < c->nv) return fail(STK_ERR_BAD_ARG, "totals capacity %d < %d", cap, c->nv);
  std::lock_guard<std::mutex> g(c->mu);
  if (!c->have_ref) return fail(STK_ERR_STATE, "stk_ecc_set_reference must come first");
  Lane& ln = c->lanes[0];
  uint8_t* d_frame = nullptr;
  rc = ensure_host_staging(c, ln, true, &d_frame);
  if (rc) return rc;
  rc = upload_frame(c, ln, d_frame, bgr, pitch, false);
  if (rc) return rc;
  const size_t row = (size_t)c->cfg.width * c->cfg.channels;
  rc = launch_prep(c, d_frame, row, ln.small, ln.tmpl, ln.stream);
  if (rc) return rc;
  const bool persp = c->cfg.motion_type == STK_MOTION_HOMOGRAPHY;
  stk::ecc_init_kernel<<<1, 32, 0, ln.stream>>>(ln.st, persp ? 1 : 0, 1 << 30, -1.0, 0, 0);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ln.st->m, warp_in, 9 * sizeof(float), cudaMemcpyHostToDevice, ln.stream));
  double* d_tot = nullptr;
  CU(cudaMalloc((void**)&d_tot, sizeof(double) * c->nv));
  stk::EccIterParams ip = iter_params(c, ln, false);
  ip.totals_out = d_tot;
  void* args[] = {&ip};
  cudaError_t e = cudaLaunchKernel(c->iter_fn, dim3(c->n_tiles), dim3(c->iter_threads), args, c->iter_smem, ln.stream);
  stk::EccState hs;
  if (e == cudaSuccess) e = cudaMemcpyAsync(totals, d_tot, sizeof(double) * c->nv, cudaMemcpyDeviceToHost, ln.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&hs, ln.st, sizeof hs, cudaMemcpyDeviceToHost, ln.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ln.stream);
  cudaFree(d_tot);
  if (e != cudaSuccess) return fail(STK_ERR_CUDA, "debug iteration: %s", cudaGetErrorString(e));
  *nv = c->nv;
  if (warp_out) for (int i = 0; i < 9; ++i) warp_out[i] = hs.m[i];
  if (rho) *rho = hs.rho;
  if (status) *status = hs.status;
  return STK_OK;
}

int stk_ecc_debug_timing(stk_ecc_ctx* c, const uint8_t* bgr, size_t pitch, const float warp_in[9], int iters,
                         uint64_t* stamps, int cap, int* n_tiles) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!bgr || !warp_in || !stamps || !n_tiles) return fail(STK_ERR_BAD_ARG, "null argument");
  if (!c->cfg.align) return fail(STK_ERR_STATE, "context was created with align = 0");
  const int need = c->n_tiles * 4 + 4;
  if (cap < need) return fail(STK_ERR_BAD_ARG, "stamps capacity %d < %d", cap, need);
  std::lock_guard<std::mutex> g(c->mu);
  if (!c->have_ref) return fail(STK_ERR_STATE, "stk_ecc_set_reference must come first");
  Lane& ln = c->lanes[0];
  uint8_t* d_frame = nullptr;
  rc = ensure_host_staging(c, ln, true, &d_frame);
  if (rc) return rc;
  rc = upload_frame(c, ln, d_frame, bgr, pitch, false);
  if (rc) return rc;
  const size_t row = (size_t)c->cfg.width * c->cfg.channels;
  rc = launch_prep(c, d_frame, row, ln.small, ln.tmpl, ln.stream);
  if (rc) return rc;
  const bool persp = c->cfg.motion_type == STK_MOTION_HOMOGRAPHY;
  stk::ecc_init_kernel<<<1, 32, 0, ln.stream>>>(ln.st, persp ? 1 : 0, 1 << 30, -1.0, 0, 0);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ln.st->m, warp_in, 9 * sizeof(float), cudaMemcpyHostToDevice, ln.stream));
  unsigned long long* d_t = nullptr;
  CU(cudaMalloc((void**)&d_t, sizeof(unsigned long long) * need));
  stk::EccIterParams ip = iter_params(c, ln, false);
  ip.timing_out = d_t;
  void* args[] = {&ip};
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < std::max(1, iters) && e == cudaSuccess; ++i)
    e = cudaLaunchKernel(c->iter_fn, dim3(c->n_tiles), dim3(c->iter_threads), args, c->iter_smem, ln.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(stamps, d_t, sizeof(unsigned long long) * need, cudaMemcpyDeviceToHost, ln.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ln.stream);
  cudaFree(d_t);
  if (e != cudaSuccess) return fail(STK_ERR_CUDA, "debug timing: %s", cudaGetErrorString(e));
  *n_tiles = c->n_tiles;
  return STK_OK;
}

/* ---- Tenengrad ------------------------------------------------------------------------------------ */
}  // extern "C"  (reopened below)

namespace {
// Per-device scratch for the block sums of the sharpness kernels: cudaMalloc/cudaFree per call cost
// milliseconds (measured 4-9 ms per call with a few hundred MB of frames resident), far more than the kernels.
// Grow-only, one per device, calls serialised by the mutex for as long as they use it.
struct SumScratch {
  std::mutex mu;
  unsigned long long* d[64] = {};
  size_t cap[64] = {};
};
SumScratch g_scratch;

int scratch_for(int dev, size_t bytes, unsigned long long** out) {
  if (dev < 0 || dev >= 64) return fail(STK_ERR_BAD_ARG, "device ordinal %d out of range", dev);
  if (g_scratch.cap[dev] < bytes) {
    if (g_scratch.d[dev]) cudaFree(g_scratch.d[dev]);
    g_scratch.d[dev] = nullptr; g_scratch.cap[dev] = 0;
    const size_t want = std::max(bytes, (size_t)1 << 20);
    if (cudaMalloc((void**)&g_scratch.d[dev], want) != cudaSuccess) return fail(STK_ERR_NOMEM, "cudaMalloc(sum scratch) failed");
    g_scratch.cap[dev] = want;
  }
  *out = g_scratch.d[dev];
  return STK_OK;
}
}  // namespace

extern "C" {
static int tenengrad_taps(int ksize, stk::TenengradParams& p) {
  memset(p.dtap, 0, sizeof p.dtap);
  memset(p.stap, 0, sizeof p.stap);
  switch (ksize) {   // getDerivKernels(dx = 1, dy = 0, ksize)
    case 1: { const int d[] = {-1, 0, 1}, s[] = {0, 1, 0}; p.radius = 1; memcpy(p.dtap, d, sizeof d); memcpy(p.stap, s, sizeof s); break; }
    case 3: { const int d[] = {-1, 0, 1}, s[] = {1, 2, 1}; p.radius = 1; memcpy(p.dtap, d, sizeof d); memcpy(p.stap, s, sizeof s); break; }
    case 5: { const int d[] = {-1, -2, 0, 2, 1}, s[] = {1, 4, 6, 4, 1}; p.radius = 2; memcpy(p.dtap, d, sizeof d); memcpy(p.stap, s, sizeof s); break; }
    case 7: { const int d[] = {-1, -4, -5, 0, 5, 4, 1}, s[] = {1, 6, 15, 20, 15, 6, 1}; p.radius = 3; memcpy(p.dtap, d, sizeof d); memcpy(p.stap, s, sizeof s); break; }
    default: return fail(STK_ERR_BAD_ARG, "Kernel size must be 1, 3, 5, or 7");
  }
  return STK_OK;
}

int stk_tenengrad_batch_device(const uint8_t* d_imgs, size_t frame_stride, size_t pitch, int width, int height,
                               int channels, int ksize, int n, int device, double* out) {
  stk::TenengradParams p = {};
  int rc = tenengrad_taps(ksize, p);
  if (rc) return rc;
  if (!d_imgs || !out || n <= 0) return fail(STK_ERR_BAD_ARG, "null/empty argument");
  if (width <= 0 || height <= 0) return fail(STK_ERR_BAD_ARG, "bad size");
  if (channels != 1 && channels != 3 && channels != 4) return fail(STK_ERR_UNSUPPORTED, "channels must be 1, 3 or 4");
  if (pitch < (size_t)width * channels) return fail(STK_ERR_BAD_ARG, "pitch too small");
  if (device >= 0) CU(cudaSetDevice(device));
  if (ksize == 3 && channels == 1 && width % stk::kTsCols == 0 && height >= 2 && pitch % 16 == 0 && frame_stride % 16 == 0 &&
      ((uintptr_t)d_imgs) % 16 == 0 && getenv("STK_TENENGRAD_STREAM") == nullptr) {
    // k = 3 on 16-byte aligned grey planes: the dedicated streaming kernel, the whole batch in one launch
    unsigned long long* d_sums = nullptr;
    const size_t sum_bytes = sizeof(unsigned long long) * stk::kSumSlots * (size_t)n;
    int cur_dev = 0;
    CU(cudaGetDevice(&cur_dev));
    std::lock_guard<std::mutex> scratch_lock(g_scratch.mu);
    rc = scratch_for(cur_dev, sum_bytes, &d_sums);
    if (rc) return rc;
    cudaError_t e = cudaMemsetAsync(d_sums, 0, sum_bytes, 0);
    stk::TenStreamParams q = {};
    q.src = d_imgs; q.frame_stride = frame_stride; q.pitch = pitch; q.width = width; q.height = height;
    // bands of up to kTsBand rows, shortened on small batches so that the grid still fills the device
    q.bands = (height + stk::kTsBand - 1) / stk::kTsBand;
    static const int cols = [] { const char* e = getenv("STK_TENENGRAD_COLS"); return (e && atoi(e) == 8) ? 8 : 16; }();
    q.col_blocks = (width + stk::kTsThreads * cols - 1) / (stk::kTsThreads * cols);
    q.sums = d_sums;
    std::vector<unsigned long long> h((size_t)n * stk::kSumSlots);
    if (e == cudaSuccess) {
      const long long blocks = (long long)n * q.bands * q.col_blocks;
      if (cols == 8) stk::tenengrad_stream_kernel<8><<<(unsigned)blocks, stk::kTsThreads>>>(q);
      else stk::tenengrad_stream_kernel<16><<<(unsigned)blocks, stk::kTsThreads>>>(q);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(h.data(), d_sums, sum_bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(STK_ERR_CUDA, "tenengrad: %s", cudaGetErrorString(e));
    const double scale = 1.0 / ((double)width * (double)height);
    for (int i = 0; i < n; ++i) {
      unsigned long long t = 0;
      for (int k = 0; k < stk::kSumSlots; ++k) t += h[(size_t)i * stk::kSumSlots + k];
      out[i] = (double)t * scale;   // cv::mean: exact integer sum * (1.0 / N)
    }
    return STK_OK;
  }
  if (ksize == 3 && channels == 1 && width % 4 == 0 && width >= 8 && height >= 2 && pitch % 4 == 0 && frame_stride % 4 == 0 &&
      ((uintptr_t)d_imgs) % 4 == 0) {
    // k = 3 on grey planes: the streaming kernel (same exact integer sum, ~4x faster than the tiled one)
    std::vector<double> all((size_t)n * 4);
    rc = stk_sharpness_all_batch_device(d_imgs, frame_stride, pitch, width, height, channels, n, -1, all.data());
    if (rc) return rc;
    for (int i = 0; i < n; ++i) out[i] = all[(size_t)i * 4 + 2];
    return STK_OK;
  }
  unsigned long long* d_sums = nullptr;
  const size_t sum_bytes = sizeof(unsigned long long) * stk::kSumSlots * (size_t)n;
  int cur_dev = 0;
  CU(cudaGetDevice(&cur_dev));
  std::lock_guard<std::mutex> scratch_lock(g_scratch.mu);
  rc = scratch_for(cur_dev, sum_bytes, &d_sums);
  if (rc) return rc;
  cudaError_t e = cudaMemsetAsync(d_sums, 0, sum_bytes, 0);
  std::vector<unsigned long long> h((size_t)n * stk::kSumSlots);
  if (e == cudaSuccess) {
    p.src = d_imgs; p.frame_stride = frame_stride; p.pitch = pitch;
    p.width = width; p.height = height; p.channels = channels; p.sums = d_sums;
    for (int z0 = 0; z0 < n && e == cudaSuccess; z0 += 32768) {
      stk::TenengradParams q = p;
      q.src = d_imgs + (size_t)z0 * frame_stride;
      q.sums = d_sums + (size_t)z0 * stk::kSumSlots;
      dim3 grid((width + stk::kTenTW - 1) / stk::kTenTW, (height + stk::kTenTH - 1) / stk::kTenTH, std::min(32768, n - z0));
      stk::tenengrad_kernel<<<grid, stk::kTenThreads>>>(q);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(h.data(), d_sums, sum_bytes, cudaMemcpyDeviceToHost);
  }
  if (e != cudaSuccess) return fail(STK_ERR_CUDA, "tenengrad: %s", cudaGetErrorString(e));
  const double scale = 1.0 / ((double)width * (double)height);
  for (int i = 0; i < n; ++i) {
    unsigned long long t = 0;
    for (int k = 0; k < stk::kSumSlots; ++k) t += h[(size_t)i * stk::kSumSlots + k];
    out[i] = (double)t * scale;   // cv::mean: exact integer sum * (1.0 / N)
  }
  return STK_OK;
}

int stk_tenengrad_device(const uint8_t* d_img, size_t pitch, int width, int height, int channels, int ksize,
                         int device, double* out) {
  return stk_tenengrad_batch_device(d_img, 0, pitch, width, height, channels, ksize, 1, device, out);
}

int stk_tenengrad(const uint8_t* img, size_t pitch, int width, int height, int channels, int ksize, int device,
                  double* out) {
  stk::TenengradParams p = {};
  int rc = tenengrad_taps(ksize, p);
  if (rc) return rc;
  if (!img || !out) return fail(STK_ERR_BAD_ARG, "null argument");
  if (width <= 0 || height <= 0) return fail(STK_ERR_BAD_ARG, "bad size");
  if (channels != 1 && channels != 3 && channels != 4) return fail(STK_ERR_UNSUPPORTED, "channels must be 1, 3 or 4");
  const size_t row = (size_t)width * channels;
  if (pitch < row) return fail(STK_ERR_BAD_ARG, "pitch too small");
  if (device >= 0) CU(cudaSetDevice(device));
  uint8_t* d = nullptr;
  CU(cudaMalloc((void**)&d, row * height));
  cudaError_t e = cudaMemcpy2D(d, row, img, pitch, row, height, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d); return fail(STK_ERR_CUDA, "tenengrad upload: %s", cudaGetErrorString(e)); }
  rc = stk_tenengrad_batch_device(d, 0, row, width, height, channels, ksize, 1, -1, out);
  cudaFree(d);
  return rc;
}

/* ---- LAPM / LAPV / TENG(3) / GLVN in one pass ----------------------------------------------------------- */
// scalar tails in cv::mean / cv::meanStdDev's order (oracle/restate.py::_mean_std_from_sums); host code is
// compiled for baseline x86-64, so nothing here is contracted into an FMA
static void sharpness_from_sums(const unsigned long long* v, double n_px, double* out) {
  const double scale = 1.0 / n_px;
  auto mean_sigma = [&](double s, double sq, double& mean, double& sigma) {
    mean = s * scale;
    const double var = sq * scale - mean * mean;
    sigma = std::sqrt(var > 0.0 ? var : 0.0);
  };
  out[0] = ((double)v[1] / 4.0) * scale;                                  // LAPM: mean(|lx| + |ly|)
  double mean, sigma;
  mean_sigma((double)(long long)v[2], (double)v[3], mean, sigma);
  out[1] = sigma * sigma;                                                 // LAPV
  out[2] = (double)v[0] * scale;                                          // TENG
  mean_sigma((double)v[4], (double)v[5], mean, sigma);
  const double eps = 2.220446049250313e-16;
  out[3] = (sigma * sigma) / (mean > eps ? mean : eps);                   // GLVN
}

int stk_sharpness_all_batch_device(const uint8_t* d_imgs, size_t frame_stride, size_t pitch, int width, int height,
                                   int channels, int n, int device, double* out) {
  if (!d_imgs || !out || n <= 0) return fail(STK_ERR_BAD_ARG, "null/empty argument");
  if (width <= 0 || height <= 0) return fail(STK_ERR_BAD_ARG, "bad size");
  if (channels != 1 && channels != 3 && channels != 4) return fail(STK_ERR_UNSUPPORTED, "channels must be 1, 3 or 4");
  if (pitch < (size_t)width * channels) return fail(STK_ERR_BAD_ARG, "pitch too small");
  if (device >= 0) CU(cudaSetDevice(device));
  const size_t bytes = sizeof(unsigned long long) * stk::kSharpSums * stk::kSumSlots * (size_t)n;
  unsigned long long* d_sums = nullptr;
  int cur_dev = 0;
  CU(cudaGetDevice(&cur_dev));
  std::lock_guard<std::mutex> scratch_lock(g_scratch.mu);
  int rc = scratch_for(cur_dev, bytes, &d_sums);
  if (rc) return rc;
  cudaError_t e = cudaMemsetAsync(d_sums, 0, bytes, 0);
  std::vector<unsigned long long> h((size_t)n * stk::kSharpSums * stk::kSumSlots);
  for (int z0 = 0; z0 < n && e == cudaSuccess; z0 += 32768) {
    stk::SharpnessParams p = {};
    p.src = d_imgs + (size_t)z0 * frame_stride;
    p.frame_stride = frame_stride; p.pitch = pitch;
    p.width = width; p.height = height; p.channels = channels;
    p.sums = d_sums + (size_t)z0 * stk::kSharpSums * stk::kSumSlots;
    const bool stream_ok = channels == 1 && width % 4 == 0 && width >= 8 && height >= 2 && pitch % 4 == 0 &&
                           frame_stride % 4 == 0 && ((uintptr_t)d_imgs) % 4 == 0;
    if (stream_ok) {
      const int cols_per_block = stk::kStreamThreads * stk::kStreamCols;
      dim3 grid((width + cols_per_block - 1) / cols_per_block, (height + stk::kStreamBand - 1) / stk::kStreamBand, std::min(32768, n - z0));
      stk::sharpness_stream_kernel<<<grid, stk::kStreamThreads>>>(p);
    } else {
      dim3 grid((width + stk::kTenTW - 1) / stk::kTenTW, (height + stk::kTenTH - 1) / stk::kTenTH, std::min(32768, n - z0));
      stk::sharpness_all_kernel<<<grid, stk::kTenThreads>>>(p);
    }
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(h.data(), d_sums, bytes, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail(STK_ERR_CUDA, "sharpness: %s", cudaGetErrorString(e));
  for (int i = 0; i < n; ++i) {
    unsigned long long v[stk::kSharpSums] = {};
    for (int k = 0; k < stk::kSumSlots; ++k)
      for (int j = 0; j < stk::kSharpSums; ++j) v[j] += h[((size_t)i * stk::kSumSlots + k) * stk::kSharpSums + j];   // two's complement for [2]
    sharpness_from_sums(v, (double)width * (double)height, out + 4 * (size_t)i);
  }
  return STK_OK;
}

int stk_sharpness_all(const uint8_t* img, size_t pitch, int width, int height, int channels, int device, double out[4]) {
  if (!img || !out) return fail(STK_ERR_BAD_ARG, "null argument");
  if (width <= 0 || height <= 0) return fail(STK_ERR_BAD_ARG, "bad size");
  if (channels != 1 && channels != 3 && channels != 4) return fail(STK_ERR_UNSUPPORTED, "channels must be 1, 3 or 4");
  const size_t row = (size_t)width * channels;
  if (pitch < row) return fail(STK_ERR_BAD_ARG, "pitch too small");
  if (device >= 0) CU(cudaSetDevice(device));
  uint8_t* d = nullptr;
  CU(cudaMalloc((void**)&d, row * height));
  cudaError_t e = cudaMemcpy2D(d, row, img, pitch, row, height, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d); return fail(STK_ERR_CUDA, "sharpness upload: %s", cudaGetErrorString(e)); }
  const int rc = stk_sharpness_all_batch_device(d, 0, row, width, height, channels, 1, -1, out);
  cudaFree(d);
  return rc;
}

}  // extern "C"
