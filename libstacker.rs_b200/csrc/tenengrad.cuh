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
// Block sums land in one of kSumSlots per frame (slot = block index mod kSumSlots) and the host adds the slots:
// thousands of same-address atomics per frame serialise in L2 (measured: 2.5 ms per 4K frame with one
// address, see scripts/satellite_bench.py); integer sums are exact in any order.
constexpr int kSumSlots = 64;

struct TenengradParams {
  const uint8_t* src;     // frame 0 of the batch
  size_t frame_stride;    // bytes between frames
  size_t pitch;
  int width, height, channels;
  int radius;             // 1 for ksize 1 and 3, 2 for 5, 3 for 7
  int dtap[2 * kTenMaxR + 1];   // derivative taps
  int stap[2 * kTenMaxR + 1];   // smoothing taps
  unsigned long long* sums;     // [n_frames][kSumSlots]
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
    const int slot = (blockIdx.y * gridDim.x + blockIdx.x) % kSumSlots;
    atomicAdd(p.sums + (size_t)blockIdx.z * kSumSlots + slot, t);       // integer: order-independent, deterministic
  }
}

// ---- all four sharpness metrics of the crate in one pass -----------------------------------------------
// LAPM  sharpness_modified_laplacian            /root/reference/src/lib.rs:1032-1068
// LAPV  sharpness_variance_of_laplacian         /root/reference/src/lib.rs:1070-1090
// TENG  sharpness_tenengrad(k = 3)              /root/reference/src/lib.rs:1101-1147
// GLVN  sharpness_normalized_gray_level_variance /root/reference/src/lib.rs:1151-1166
// which examples/main.rs:37-50 evaluates per file with ~14 OpenCV passes over CV_64F planes.  All four are
// 3x3 stencils (or none) over the 8-bit grey plane, so one read of the plane (N bytes; 3N with the grey
// conversion fused in) yields six exact integer sums per frame:
//   [0] sum(gx^2 + gy^2)            Sobel 3x3, BORDER_REFLECT_101
//   [1] sum(|4 lx| + |4 ly|)        [-1 2 -1] x [1 2 1] and transposed, BORDER_REFLECT_101 (lx, ly in 1/4 units)
//   [2] sum(lap)  [3] sum(lap^2)    [[2 0 2],[0 -8 0],[2 0 2]], BORDER_REPLICATE  ([2] two's complement)
//   [4] sum(g)    [5] sum(g^2)
// The host finishes each metric with the same few f64 operations cv::mean / cv::meanStdDev use
// (oracle/restate.py, pinned bit-exact against cv2).
constexpr int kSharpSums = 6;

struct SharpnessParams {
  const uint8_t* src;
  size_t frame_stride, pitch;
  int width, height, channels;
  unsigned long long* sums;     // [n_frames][kSumSlots][kSharpSums]
};

__global__ void __launch_bounds__(kTenThreads) sharpness_all_kernel(const SharpnessParams p) {
  constexpr int GW = kTenTW + 2, GH = kTenTH + 2;
  __shared__ short s_g[GH][GW];
  __shared__ unsigned long long s_part[kTenThreads / 32][kSharpSums];
  const int x0 = blockIdx.x * kTenTW, y0 = blockIdx.y * kTenTH;
  const uint8_t* src = p.src + (size_t)blockIdx.z * p.frame_stride;
  const int tid = threadIdx.x;
  const int w = p.width, h = p.height;

  for (int i = tid; i < GW * GH; i += kTenThreads) {
    const int ty = i / GW, tx = i - ty * GW;
    const int sx = reflect101(x0 + tx - 1, w);
    const int sy = reflect101(y0 + ty - 1, h);
    const uint8_t* px = src + (size_t)sy * p.pitch + (size_t)sx * p.channels;
    s_g[ty][tx] = (short)(p.channels == 1 ? (int)px[0] : bgr2gray(px[0], px[1], px[2]));
  }
  __syncthreads();

  unsigned long long teng = 0, lapm = 0, lapsq = 0;
  long long lapsum = 0;
  unsigned int gsum = 0, gsq = 0;
  for (int i = tid; i < kTenTH * kTenTW; i += kTenThreads) {
    const int ty = i / kTenTW, tx = i - ty * kTenTW;
    const int x = x0 + tx, y = y0 + ty;
    if (x >= w || y >= h) continue;
    const int cy = ty + 1, cx = tx + 1;
    // REFLECT_101 neighbourhood
    const int a00 = s_g[cy - 1][cx - 1], a01 = s_g[cy - 1][cx], a02 = s_g[cy - 1][cx + 1];
    const int a10 = s_g[cy][cx - 1], a11 = s_g[cy][cx], a12 = s_g[cy][cx + 1];
    const int a20 = s_g[cy + 1][cx - 1], a21 = s_g[cy + 1][cx], a22 = s_g[cy + 1][cx + 1];
    const int gx = (a02 - a00) + 2 * (a12 - a10) + (a22 - a20);
    const int gy = (a20 - a00) + 2 * (a21 - a01) + (a22 - a02);
    teng += (unsigned long long)(gx * gx) + (unsigned long long)(gy * gy);
    // 4*lx: [-1 2 -1] along x, [1 2 1] along y ; 4*ly transposed
    const int r0 = 2 * a01 - a00 - a02, r1 = 2 * a11 - a10 - a12, r2 = 2 * a21 - a20 - a22;
    const int c0 = 2 * a10 - a00 - a20, c1 = 2 * a11 - a01 - a21, c2 = 2 * a12 - a02 - a22;
    lapm += (unsigned)(abs(r0 + 2 * r1 + r2) + abs(c0 + 2 * c1 + c2));
    // BORDER_REPLICATE corners (the clamped positions are always inside the tile)
    const int xl = cx - (x > 0 ? 1 : 0), xr = cx + (x < w - 1 ? 1 : 0);
    const int yu = cy - (y > 0 ? 1 : 0), yd = cy + (y < h - 1 ? 1 : 0);
    const int lap = 2 * (s_g[yu][xl] + s_g[yu][xr] + s_g[yd][xl] + s_g[yd][xr]) - 8 * a11;
    lapsum += lap;
    lapsq += (unsigned long long)(lap * lap);
    gsum += (unsigned)a11;
    gsq += (unsigned)(a11 * a11);
  }
  unsigned long long v[kSharpSums] = {teng, lapm, (unsigned long long)lapsum, lapsq, gsum, gsq};
#pragma unroll
  for (int k = 0; k < kSharpSums; ++k) {
    v[k] = warp_sum(v[k]);
    if ((tid & 31) == 0) s_part[tid >> 5][k] = v[k];
  }
  __syncthreads();
  if (tid < kSharpSums) {
    unsigned long long t = 0;
    for (int wp = 0; wp < kTenThreads / 32; ++wp) t += s_part[wp][tid];
    const int slot = (blockIdx.y * gridDim.x + blockIdx.x) % kSumSlots;
    atomicAdd(p.sums + ((size_t)blockIdx.z * kSumSlots + slot) * kSharpSums + tid, t);     // integer: order-independent
  }
}

// ---- streaming variant for 8-bit grey planes (what the crate's callers pass: IMREAD_GRAYSCALE) -----------------
// Same six sums as sharpness_all_kernel, restructured after measuring it at 48 us per 4K frame (3 % of HBM peak):
// a thread owns 4 adjacent columns (one aligned 32-bit load per row + the two neighbour bytes) and streams down a
// band of rows with a 3-row register window.  The vertical parts of all four stencils are shared per COLUMN
// (t = up + down, u = 2*mid, s = t + u, e = u - t, d = down - up, q = up' + down' with the REPLICATE rows), so a
// pixel costs ~18 integer operations; band sums fit 32 bits and are widened once per thread.
// Requirements (else the tiled kernel above runs): channels == 1, width % 4 == 0, pitch % 4 == 0, 4-byte aligned base.
constexpr int kStreamThreads = 256, kStreamBand = 32, kStreamCols = 4;

__global__ void __launch_bounds__(kStreamThreads) sharpness_stream_kernel(const SharpnessParams p) {
  __shared__ unsigned long long s_part[kStreamThreads / 32][kSharpSums];
  const int w = p.width, h = p.height;
  const int x0 = (blockIdx.x * kStreamThreads + threadIdx.x) * kStreamCols;
  const int y0 = blockIdx.y * kStreamBand;
  const uint8_t* src = p.src + (size_t)blockIdx.z * p.frame_stride;
  const int tid = threadIdx.x;
  unsigned int teng = 0, lapm = 0, lapsq = 0, gsum = 0, gsq = 0;
  int lapsum = 0;
  if (x0 < w) {
    const int xl = reflect101(x0 - 1, w), xr = reflect101(x0 + kStreamCols, w);
    const bool edge_l = x0 == 0, edge_r = x0 + kStreamCols == w;
    auto load_row = [&](int y, int (&v)[6]) {
      const uint8_t* row = src + (size_t)reflect101(y, h) * p.pitch;
      const unsigned int word = __ldg(reinterpret_cast<const unsigned int*>(row + x0));
      v[0] = __ldg(row + xl);
      v[1] = word & 0xff; v[2] = (word >> 8) & 0xff; v[3] = (word >> 16) & 0xff; v[4] = word >> 24;
      v[5] = __ldg(row + xr);
    };
    int up[6], mid[6], dn[6];
    load_row(y0 - 1, up);
    load_row(y0, mid);
    const int y_end = min(y0 + kStreamBand, h);
    for (int y = y0; y < y_end; ++y) {
      load_row(y + 1, dn);
      // BORDER_REPLICATE rows for the Laplacian: above row 0 is row 0, below row h-1 is row h-1
      const bool top = y == 0, bot = y == h - 1;
      int s[6], e[6], d[6], q[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int t = up[j] + dn[j], u = 2 * mid[j];
        s[j] = t + u; e[j] = u - t; d[j] = dn[j] - up[j];
        q[j] = (top ? mid[j] : up[j]) + (bot ? mid[j] : dn[j]);
      }
      // REPLICATE columns: left of column 0 is column 0, right of column w-1 is column w-1
      const int q_l = edge_l ? q[1] : q[0], q_r = edge_r ? q[4] : q[5];
#pragma unroll
      for (int i = 0; i < kStreamCols; ++i) {
        const int gx = s[i + 2] - s[i];
        const int gy = d[i] + 2 * d[i + 1] + d[i + 2];
        teng += (unsigned)(gx * gx) + (unsigned)(gy * gy);
        lapm += (unsigned)(abs(2 * s[i + 1] - s[i] - s[i + 2]) + abs(e[i] + 2 * e[i + 1] + e[i + 2]));
        const int ql = i == 0 ? q_l : q[i], qr = i == kStreamCols - 1 ? q_r : q[i + 2];
        const int c = mid[i + 1];
        const int lap = 2 * (ql + qr) - 8 * c;
        lapsum += lap;
        lapsq += (unsigned)(lap * lap);
        gsum += (unsigned)c;
        gsq += (unsigned)(c * c);
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) { up[j] = mid[j]; mid[j] = dn[j]; }
    }
  }
  unsigned long long v[kSharpSums] = {teng, lapm, (unsigned long long)(long long)lapsum, lapsq, gsum, gsq};
#pragma unroll
  for (int k = 0; k < kSharpSums; ++k) {
    v[k] = warp_sum(v[k]);
    if ((tid & 31) == 0) s_part[tid >> 5][k] = v[k];
  }
  __syncthreads();
  if (tid < kSharpSums) {
    unsigned long long t = 0;
    for (int wp = 0; wp < kStreamThreads / 32; ++wp) t += s_part[wp][tid];
    const int slot = (blockIdx.y * gridDim.x + blockIdx.x) % kSumSlots;
    atomicAdd(p.sums + ((size_t)blockIdx.z * kSumSlots + slot) * kSharpSums + tid, t);
  }
}

// ---- Tenengrad(3) alone, streaming, for batches of 8-bit grey planes ---------------------------------------------
// What sharpness_tenengrad (ksize 3) needs is ONE sum per frame, and ranking a stack asks for it on every frame
// (examples/main.rs:37-64).  Round-1 ncu put the fused four-metric kernel at 16-26 us per 4K plane — 6-8 % of HBM,
// ~26 integer operations per pixel behind three 1-4-byte loads per row — so the ranking metric gets its own kernel:
//   * a thread owns 16 adjacent columns: ONE 128-bit load per row; the two halo columns come from the neighbouring
//     lanes by shuffle (BORDER_REFLECT_101 at the plane's edges is a select of the lane's own second/second-last byte);
//   * rows stream through a 3-row register window that is ROTATED BY NAME (the loop is unrolled three times), so no
//     register is ever copied; the next two rows' loads are always in flight;
//   * per column and row the vertical terms s = up + 2 mid + down and d = down - up are formed once and shared by the
//     three pixels that use them: ~8 integer operations per pixel (gx = s[c+1] - s[c-1], gy = d[c-1] + 2 d[c] + d[c+1],
//     two multiply-adds into a 32-bit band sum, widened to 64 bits once per band);
//   * the whole batch is one launch; a block adds ONE 64-bit value to its frame's slot.
// Same exact integer sum as tenengrad_kernel, so the f64 result is bit-identical to the OpenCV CV_64F pipeline.
// Requirements (else the tiled kernels run): channels == 1, width % 16 == 0, pitch % 16 == 0, 16-byte aligned planes.
constexpr int kTsThreads = 128, kTsCols = 16, kTsBand = 48;     // 48 rows x 16 cols x 2 x 1020^2 < 2^32

struct TenStreamParams {
  const uint8_t* src;
  size_t frame_stride, pitch;
  int width, height;
  int bands;                    // ceil(height / kTsBand)
  int col_blocks;               // ceil(width / (kTsThreads * COLS))
  unsigned long long* sums;     // [n_frames][kSumSlots]
};

// COLS = 16 (one 128-bit load per row and thread) or 8 (64-bit loads, half the registers, twice the warps)
template <int COLS> struct TsRow;
template <> struct TsRow<16> { using type = uint4; };
template <> struct TsRow<8> { using type = uint2; };

template <int COLS>
__global__ void __launch_bounds__(kTsThreads) tenengrad_stream_kernel(const TenStreamParams p) {
  using Row = typename TsRow<COLS>::type;
  constexpr int NW = COLS / 4;
  __shared__ unsigned long long s_part[kTsThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31;
  // blockIdx.x = (frame, band, column block) flattened, column block fastest
  int b = blockIdx.x;
  const int cb = b % p.col_blocks; b /= p.col_blocks;
  const int band = b % p.bands;
  const int frame = b / p.bands;
  const int w = p.width, h = p.height;
  const int x0 = (cb * kTsThreads + tid) * COLS;
  const int y0 = band * kTsBand;
  const int y_end = min(y0 + kTsBand, h);
  const uint8_t* src = p.src + (size_t)frame * p.frame_stride;
  const bool active = x0 < w;
  const bool edge_l = x0 == 0, edge_r = x0 + COLS == w;
  const int xa = active ? x0 : 0;                 // idle threads load column block 0 again (they only feed shuffles)
  // halo bytes across warp boundaries (lane 0's left, lane 31's right) cannot come from a shuffle: one byte load each
  const bool need_l = lane == 0 && !edge_l && active, need_r = lane == 31 && !edge_r && active;
  const int xl = need_l ? x0 - 1 : xa, xr = need_r ? x0 + COLS : xa;

  // BORDER_REFLECT_101 row (|overshoot| is at most 2 here and height >= 2): no loop, no branch
  auto row_of = [&](int y) -> const uint8_t* {
    const int yy = y < 0 ? -y : (y >= h ? 2 * h - 2 - y : y);
    return src + (size_t)yy * p.pitch;
  };
  struct Raw { Row q; unsigned l, r; };
  auto load = [&](int y) -> Raw {
    const uint8_t* row = row_of(y);
    Raw t;
    t.q = __ldg(reinterpret_cast<const Row*>(row + xa));
    t.l = (lane == 0) ? (unsigned)__ldg(row + xl) : 0u;          // predicated single-byte loads on the warp's end lanes
    t.r = (lane == 31) ? (unsigned)__ldg(row + xr) : 0u;
    return t;
  };
  // the row as COLS + 2 ints: [halo left, own columns, halo right]
  auto expand = [&](const Raw& t, int (&v)[COLS + 2]) {
    unsigned wds[NW];
    if constexpr (COLS == 16) { wds[0] = t.q.x; wds[1] = t.q.y; wds[2] = t.q.z; wds[3] = t.q.w; }
    else { wds[0] = t.q.x; wds[1] = t.q.y; }
    // neighbours' edge bytes: lane-1's last byte, lane+1's first byte (a warp's lanes are adjacent column groups)
    const unsigned from_l = __shfl_up_sync(0xffffffffu, wds[NW - 1] >> 24, 1);
    const unsigned from_r = __shfl_down_sync(0xffffffffu, wds[0] & 0xffu, 1);
    // plane edges: REFLECT_101 = column 1 / column w-2, i.e. this thread's own second / second-last byte
    const unsigned left = edge_l ? ((wds[0] >> 8) & 0xffu) : (need_l ? t.l : from_l);
    const unsigned right = edge_r ? ((wds[NW - 1] >> 16) & 0xffu) : (need_r ? t.r : from_r);
    v[0] = (int)left;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      v[1 + 4 * k] = (int)(wds[k] & 0xffu);
      v[2 + 4 * k] = (int)((wds[k] >> 8) & 0xffu);
      v[3 + 4 * k] = (int)((wds[k] >> 16) & 0xffu);
      v[4 + 4 * k] = (int)(wds[k] >> 24);
    }
    v[COLS + 1] = (int)right;
  };

  unsigned int acc = 0;
  auto row_terms = [&](const int (&up)[COLS + 2], const int (&mid)[COLS + 2], const int (&dn)[COLS + 2]) {
    int sv[COLS + 2], dv[COLS + 2];
#pragma unroll
    for (int j = 0; j < COLS + 2; ++j) { sv[j] = up[j] + 2 * mid[j] + dn[j]; dv[j] = dn[j] - up[j]; }
#pragma unroll
    for (int i = 0; i < COLS; ++i) {
      const int gx = sv[i + 2] - sv[i];
      const int gy = dv[i] + 2 * dv[i + 1] + dv[i + 2];
      acc += (unsigned)(gx * gx) + (unsigned)(gy * gy);
    }
  };

  int ra[COLS + 2], rb[COLS + 2], rc[COLS + 2];
  expand(load(y0 - 1), ra);
  expand(load(y0), rb);
  int y = y0;
  // three rows per trip, the window rotating by name: (ra, rb, rc) -> (rb, rc, ra) -> (rc, ra, rb); the trip's three
  // loads are issued before the first row is touched.  (Measured in round 2: prefetching the NEXT trip's rows as well —
  // 111 registers, 4 blocks per SM instead of 5 — is slower, 7.6 against 6.8 us per 4K plane: occupancy hides the load
  // latency better than a deeper per-thread pipeline does here.  Forcing 6 or 7 blocks per SM through launch bounds — 80 / 72
  // registers — changes nothing either: 6.9 / 6.9 / 6.6 us per plane for 5 / 6 / 7 blocks; of the 6.8 us per plane 5.4 are the
  // kernel, the rest the call's own memset, copy-back and host sum.)
  for (; y + 3 <= y_end; y += 3) {
    const Raw q1 = load(y + 1), q2 = load(y + 2), q3 = load(y + 3);
    expand(q1, rc); if (active) row_terms(ra, rb, rc);
    expand(q2, ra); if (active) row_terms(rb, rc, ra);
    expand(q3, rb); if (active) row_terms(rc, ra, rb);
  }
  if (y < y_end) { expand(load(y + 1), rc); if (active) row_terms(ra, rb, rc); ++y; }
  if (y < y_end) { expand(load(y + 1), ra); if (active) row_terms(rb, rc, ra); ++y; }

  unsigned long long t = warp_sum((unsigned long long)acc);
  if (lane == 0) s_part[tid >> 5] = t;
  __syncthreads();
  if (tid == 0) {
    unsigned long long tot = 0;
#pragma unroll
    for (int k = 0; k < kTsThreads / 32; ++k) tot += s_part[k];
    const int slot = (band * p.col_blocks + cb) % kSumSlots;
    atomicAdd(p.sums + (size_t)frame * kSumSlots + slot, tot);       // integer: order-independent, deterministic
  }
}

}  // namespace stk
