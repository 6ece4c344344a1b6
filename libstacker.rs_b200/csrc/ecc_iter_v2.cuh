// K2, second generation — the same fused ECC iteration as ecc_iter.cuh (same sums, same epilogue, same
// accumulators; see the header of that file for the algorithm and the reference call it replaces,
// /root/reference/src/lib.rs:769-777), with the block-level plumbing rebuilt around what the round-1 ncu
// source view showed (profiles/r1_summary.md): 34 % of the stall samples were outside the pixel body —
// the producer thread waiting on "empty" barriers (which kept all 8 warps of a block in lock-step), the
// per-segment block barrier of the strip fold, and a pipeline that drained at every strip change.
//
//   * geometry is a compile-time configuration (EccCfg): threads per block (128 columns x THREADS/128 row
//     parts), rows per thread and chunk (RPT), rows per unrolled group, TMA stages, blocks per SM —
//     so occupancy / chunk height / pipeline depth are chosen by measurement, not baked in;
//   * the box table covers the block's WHOLE chunk range (all strips), so the TMA pipeline never drains
//     at a strip change; a thread folds its column sums when its strip changes, inside its warp only;
//   * no block barrier in the chunk loop and none at a fold: every warp owns an f64 accumulator row in
//     shared memory (lane l owns sums 2l, 2l+1 after the transpose-reduce), added across warps once, in
//     fixed order, when the block is done (deterministic);
//   * stage recycling without a producer: the LAST warp to finish a stage (shared-memory counter) issues
//     the TMA refill for it, so no warp ever waits for the others except through the data itself;
//   * the pixel body sheds its integer<->float conversions (fraction bits spliced under the 1.5*2^23
//     magic exponent, one packed FFMA) and the magic-constant subtractions (folded into the box pointer).
#pragma once
#include <type_traits>

#include "ecc_iter.cuh"

namespace stk {

template <int THREADS, int RPT, int UNROLL, int STAGES, int MINB, int PACK = 1, int BODY = 0>
struct EccCfg {
  static_assert(THREADS % kEccStripW == 0 && RPT % UNROLL == 0, "bad ECC kernel geometry");
  static constexpr int kThreads = THREADS;
  static constexpr int kRowParts = THREADS / kEccStripW;
  static constexpr int kRpt = RPT;              // rows per thread per chunk
  static constexpr int kUnroll = UNROLL;        // rows per straight-line group of the lean body
  static constexpr int kStages = STAGES;
  static constexpr int kMinBlocks = MINB;
  static constexpr int kPack = PACK;            // lean Homography body: 1 = packed f32x2 arithmetic, 0 = scalar, 2 = packed sums only,
                                                // 3 = packed with the premultiplied accumulator (AccumH3)
  static constexpr int kBody = BODY;            // lean_pixel variant (packed bodies only)
  static constexpr int kWarps = THREADS / 32;
  static constexpr int kChunkH = RPT * kRowParts;
  static constexpr int kBoxH = kChunkH + 16;    // same 16-row drift/halo margin as the first-generation kernel
  static constexpr int kImgBytes = kBoxW * kBoxH * 4;
  static constexpr int kTmplBytes = kEccStripW * kChunkH * 4;
  static constexpr int kStageBytes = kImgBytes + kTmplBytes;
  static constexpr int kDynSmem = STAGES * kStageBytes;
  static_assert(kBoxH <= 256 && kImgBytes % 128 == 0 && kTmplBytes % 128 == 0, "TMA box limits");
};

// Butterfly transpose-reduce of a register vector across the warp (see warp_reduce_vector) with the totals
// ADDED to this warp's own f64 row: lane l owns entries 2l and 2l+1 (and lane 0 the ones beyond 64), so
// there is no race and no barrier.
template <int NV>
__device__ __forceinline__ void warp_reduce_accumulate(float (&v)[NV], int lane, double* acc) {
  if constexpr (NV >= 32) {
    float r[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] = i < NV ? v[i] : 0.f;
#pragma unroll
    for (int half = 32, o = 16; half >= 2; half >>= 1, o >>= 1) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < half; ++i) {
        const float lo = r[i], hi = r[i + half];
        const float send = up ? lo : hi;
        const float keep = up ? hi : lo;
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    if (2 * lane < NV) acc[2 * lane] += (double)r[0];
    if (2 * lane + 1 < NV) acc[2 * lane + 1] += (double)r[1];
#pragma unroll
    for (int i = 64; i < NV; ++i) {
      const float t = warp_sum(v[i]);
      if (lane == 0) acc[i] += (double)t;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float t = warp_sum(v[i]);
      if (lane == 0) acc[i] += (double)t;
    }
  }
}

// One interior pixel of the Homography lean body (FastPersp coordinates, 12 taps from the shared-memory box, Jacobian
// generators (2a, 2b, -2t), sums): everything between the template value t_ and the accumulator update.
//   bi      index of the thread's box element for row r of the group with the magic offsets folded in:
//           ((q - M) >> 5) == (q >> 5) - (M >> 5) for M = 0x4B400000 (its low 5 bits are zero), unsigned wrap-around
//   magic   0x4B400000 held in a REGISTER (the caller gets it from the kernel parameters): (q & 31) | magic is then ONE
//           LOP3 per coordinate; as an immediate it takes two (LOP3 has a single immediate slot).
// V = 0: the round-2 first-session body; V = 1: fused LOP3, scalar (u, v) — a packed instruction holds the issue
// port for two cycles (scripts/pipe_probe.cu), so a packed op that needs a pair-forming move costs more than two scalars.
template <int V, class ACC>
__device__ __forceinline__ void lean_pixel(const FastPersp& fp, float xf, float yf, float t_, const float* box, unsigned bi,
                                           unsigned magic, ACC& acc) {
  const float rw = rcp_approx(fmaf(fp.m21, yf, fp.wc));
  const float2 d = mul2(f2(fmaf(fp.beta, yf, fp.alpha), fmaf(fmaf(-fp.m21, yf, fp.delta), yf, fp.gamma)), f2(rw));
  const float2 qf = fma2(d, f2(32.0f), f2(12582912.0f));
  const int qxb = __float_as_int(qf.x), qyb = __float_as_int(qf.y);    // M + rint(32 du), M + rint(32 dv)
  const float* bp = box + (bi + (unsigned)(qyb >> kInterBits) * (unsigned)kBoxW + (unsigned)(qxb >> kInterBits));
  // fractions k/32: splice the 5 low bits under the magic exponent (a float equal to 2^23*1.5 + k), one packed FFMA
  float2 axy;
  if constexpr (V == 0) {
    axy = fma2(f2(__int_as_float((qxb & (kInterTab - 1)) | 0x4B400000), __int_as_float((qyb & (kInterTab - 1)) | 0x4B400000)),
               f2(1.f / kInterTab), f2(-12582912.0f / kInterTab));
  } else {
    axy = fma2(f2(__int_as_float((qxb & (kInterTab - 1)) | magic), __int_as_float((qyb & (kInterTab - 1)) | magic)),
               f2(1.f / kInterTab), f2(-12582912.0f / kInterTab));
  }
  float w_;
  float2 gxy2;
  sample_box_packed(bp, axy.x, axy.y, w_, gxy2);
  const float2 g01 = mul2(gxy2, f2(rw));                          // 2a, 2b
  float g2;                                                       // -2t  (t = hatX a + hatY b, hat = -(u, v))
  if constexpr (V == 0) {
    const float2 uv = add2(f2(xf, yf), d);                        // sample position (u, v)
    g2 = fmaf(uv.x, g01.x, uv.y * g01.y);
  } else {
    g2 = fmaf(xf + d.x, g01.x, (yf + d.y) * g01.y);
  }
  acc.add_packed(g01, g2, w_, t_, yf);
}

template <class ACC, int NV, bool TWICE_NEG>
__device__ __noinline__ void fold_columns(ACC acc, float xf, int lane, double* row) {
  float v[NV];
  acc.template emit<TWICE_NEG>(xf, v);
  warp_reduce_accumulate<NV>(v, lane, row);
}

template <int MOTION, bool EXACT, class CFG>
__global__ void __launch_bounds__(CFG::kThreads, CFG::kMinBlocks) ecc_iter_v2_kernel(const __grid_constant__ EccIterParams p) {
  using L = Layout<MOTION>;
  using Md = Model<MOTION>;
  constexpr int NV = L::NV, G = L::G;
  constexpr int kThreads = CFG::kThreads, kWarps = CFG::kWarps, kStages = CFG::kStages;
  constexpr int kCH = CFG::kChunkH, kRpt = CFG::kRpt, kUnr = CFG::kUnroll, kBH = CFG::kBoxH;
  constexpr int kImgBytes = CFG::kImgBytes, kStageBytes = CFG::kStageBytes;
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ float s_m[9];
  __shared__ int s_box[kMaxChunks][4];          // xlo (multiple of 4, or INT_MIN = no box), ylo, interior flag, strip << 16 | chunk row
  __shared__ alignas(8) uint64_t s_full[kStages];
  __shared__ int s_cnt[kStages];                // warps that have finished the chunk held by the stage
  __shared__ double s_acc[kWarps][NV];
  __shared__ double s_accum[NV];
  __shared__ int s_last;
  __shared__ double s_tot[NV];

  EccState* st = p.st;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (st->cont == 0) return;      // the frame's loop already stopped (a later kernel of the unrolled chain)

  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  if (p.timing_out && tid == 0) p.timing_out[(size_t)blockIdx.x * 4 + 0] = global_ns();
  if (tid < 9) s_m[tid] = st->m[tid];
  for (int i = tid; i < kWarps * NV; i += kThreads) (&s_acc[0][0])[i] = 0.0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) { mbar_init(&s_full[s], 1); s_cnt[s] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const int cps = p.chunks_per_strip;           // chunks of kCH rows per strip
  int g0, g1;
  block_chunk_range(p.n_strips, cps, p.rim_weight, g0, g1);
  int cc = 0;                                   // chunks consumed so far: drives stage and phase

  const int col = tid & (kEccStripW - 1);
  const int part = tid / kEccStripW;
  constexpr bool fast_coords = Md::persp && !EXACT;

  // packed premultiplied accumulators: AccumH3 (Homography, FastPersp coordinates), AccumA3 (Affine, exact coordinates)
  constexpr bool affine_packed = MOTION == kAffine && CFG::kPack == 3;
  using AccT = typename std::conditional<fast_coords && CFG::kPack == 3, AccumH3,
               typename std::conditional<affine_packed, AccumA3, typename AccumFor<MOTION, fast_coords>::type>::type>::type;
  AccT acc;
  acc.clear();
  int n_safe = 0;
  int cur_strip = -1, x = 0;
  float xf = 0.f;
  bool col_ok = false;
  FastPersp fp;
  const unsigned frac_magic = p.frac_magic;     // 0x4B400000 in a register: see lean_pixel

  // column sums of the strip this thread just left -> its warp's f64 row (warp-collective, no block barrier).
  // The fold itself is an out-of-line call (fold_columns): it is rare (once per strip change) and its 130
  // temporaries must not take part in the register allocation of the pixel loop.
  auto fold = [&]() {
    acc.n += (float)n_safe;
    fold_columns<AccT, NV, fast_coords>(acc, xf, lane, s_acc[wid]);
    acc.clear();
    n_safe = 0;
  };

  while (g0 < g1) {
    const int nb = min(g1 - g0, kMaxChunks);    // one table batch (the whole range unless the frame is tiny-grid huge)
    __syncthreads();    // s_m / barriers ready (first pass); previous batch fully consumed (later passes)

    // box table: for each chunk, where its window of I starts — or "no box" when the sample positions of the
    // chunk do not fit one — and whether every sample of the chunk keeps its taps off the border
    if (tid < nb) {
      const int gi = g0 + tid;
      const int strip = gi / cps, crow = gi - strip * cps;
      const int x0 = strip * kEccStripW, x1 = min(x0 + kEccStripW, p.width) - 1;
      const int cy0 = crow * kCH, cy1 = min(cy0 + kCH, p.height) - 1;
      float umin, umax, vmin, vmax;
      bool ok = chunk_bounds_f32<Md::persp>(s_m, x0, x1, cy0, cy1, umin, umax, vmin, vmax);
      int xlo = 0, ylo = 0;
      if (ok) ok = fabsf(umin) < 1e6f && fabsf(umax) < 1e6f && fabsf(vmin) < 1e6f && fabsf(vmax) < 1e6f;
      if (ok) {
        // f32 bounds: widen by 1/32 px before flooring; the box keeps its 2-px margin on the low side and one more
        // column/row than the taps need on the high side
        xlo = ((int)floorf(umin - 0.03125f) - 2) & ~3;      // TMA: innermost coordinate on a 16-byte boundary
        ylo = (int)floorf(vmin - 0.03125f) - 2;
        ok = ((int)floorf(umax + 0.03125f) + 3 - xlo < kBoxW) && ((int)floorf(vmax + 0.03125f) + 3 - ylo < kBH);
      }
      s_box[tid][0] = ok ? xlo : INT_MIN;
      s_box[tid][1] = ylo;
      s_box[tid][2] = (ok && floorf(umin - 0.0625f) >= 1.0f && floorf(umax + 0.0625f) <= (float)(p.width - 3) &&
                       floorf(vmin - 0.0625f) >= 1.0f && floorf(vmax + 0.0625f) <= (float)(p.height - 3)) ? 1 : 0;
      s_box[tid][3] = (strip << 16) | crow;
    }
    __syncthreads();

    // TMA issue for local chunk c (sequence number cc + c) into its stage; the caller knows the stage is free
    auto issue = [&](int c) {
      const int s = (cc + c) % kStages;
      const int xlo = s_box[c][0], ylo = s_box[c][1], sc = s_box[c][3];
      const bool boxed = xlo != INT_MIN;
      unsigned char* stage = dyn + s * kStageBytes;
      mbar_expect_tx(&s_full[s], (boxed ? kImgBytes : 0) + CFG::kTmplBytes);
      if (boxed) tma_load_2d(stage, &p.tm_img, xlo, ylo, &s_full[s]);
      tma_load_2d(stage + kImgBytes, &p.tm_tmpl, (sc >> 16) * kEccStripW, (sc & 0xffff) * kCH, &s_full[s]);
    };
    if (tid == 0) {
      for (int c = 0; c < kStages && c < nb; ++c) issue(c);
    }

    // general path for one pixel: exact coordinates, every border rule, nearest-neighbour mask
    auto slow_pixel = [&](int y, float t_) {
      Coord<Md::persp> co;
      Jac<MOTION> jac;
      co.init(s_m, x);
      jac.init(s_m, xf);
      int xq, yq, xn, yn;
      const bool ok = co.at_with_nearest(y, xq, yq, xn, yn);
      Sample smp; smp.w = 0.f; smp.gx2 = 0.f; smp.gy2 = 0.f;
      float mk = 0.f;
      if (ok) {
        const int sx = xq >> kInterBits, sy = yq >> kInterBits;
        if (sx >= -1 && sx < p.width && sy >= -1 && sy < p.height) {
          const float ax = (float)(xq & (kInterTab - 1)) * (1.f / kInterTab);
          const float ay = (float)(yq & (kInterTab - 1)) * (1.f / kInterTab);
          smp = sample_general(p.img, p.pitch, p.width, p.height, sx, sy, ax, ay);
        }
        mk = ((unsigned)xn < (unsigned)p.width && (unsigned)yn < (unsigned)p.height) ? 1.f : 0.f;
      }
      float g[G];
      const float yf = (float)y;
      jac.eval(smp, yf, g);
      if (fast_coords) { g[0] *= 2.f; g[1] *= 2.f; g[G - 1] *= -2.f; }     // the run accumulates (2a, 2b, -2t)
      if (affine_packed) { g[0] *= 2.f; g[1] *= 2.f; }                       // the run accumulates (2 gx, 2 gy)
      acc.template add<false>(g, smp.w, t_, mk, yf);
    };

    for (int c = 0; c < nb; ++c) {
      const int gc = cc + c;
      const int s = gc % kStages;
      const int xlo = s_box[c][0], ylo = s_box[c][1], sc = s_box[c][3];
      const int strip = sc >> 16;
      if (strip != cur_strip) {               // warp-uniform: a chunk belongs to one strip
        if (cur_strip >= 0) fold();
        cur_strip = strip;
        x = strip * kEccStripW + col;
        xf = (float)x;
        col_ok = x < p.width;
        if (fast_coords) fp.init(s_m, x);
      }
      mbar_wait(&s_full[s], (unsigned)(gc / kStages) & 1u);
      if (p.timing_out && tid == 0 && gc == 0) p.timing_out[(size_t)blockIdx.x * 4 + 3] = global_ns();   // first chunk landed
      const float* box = reinterpret_cast<const float*>(dyn + s * kStageBytes);
      const float* tbox = reinterpret_cast<const float*>(dyn + s * kStageBytes + kImgBytes);
      const bool boxed = xlo != INT_MIN;
      const int cy0 = (sc & 0xffff) * kCH;
      const int ya = cy0 + part * kRpt;
      const int yb = min(ya + kRpt, p.height);
      // lanes beyond the image width sit the chunk out; warp votes below use the mask of the lanes that work
      const unsigned wmask = __ballot_sync(0xffffffffu, col_ok && ya < yb);
      if (col_ok && ya < yb) {
        const float* trow = tbox + (ya - cy0) * kEccStripW + col;
        if (boxed && s_box[c][2] != 0 && yb - ya == kRpt) {
          // lean path (interior chunk, full height): straight-line groups of kUnr rows — no border rule, no mask,
          // no vote, no branch — so the rows of a group interleave freely in the schedule
          if constexpr (fast_coords && (CFG::kPack == 0 || CFG::kPack == 2)) {
            // scalar arithmetic (kPack 0) / scalar sampling with packed sums (kPack 2): a packed f32x2 instruction with
            // three distinct register-pair operands holds the FP32 pipe for ~4 cycles, not 2 (scripts/pipe_probe.cu),
            // so packing only pays where the loop is issue-bound
            const float* bp0 = box + (ya - ylo) * kBoxW + (x - xlo);
            float yf0 = (float)ya;
#pragma unroll 1
            for (int rg = 0; rg < kRpt; rg += kUnr) {
#pragma unroll
              for (int r = 0; r < kUnr; ++r) {
                const float yf = yf0 + (float)r;
                const float t_ = trow[r * kEccStripW];
                int qx, qy;
                float du, dv, rw;
                fp.at(yf, qx, qy, du, dv, rw);
                const float* bp = bp0 + ((qy >> kInterBits) + r) * kBoxW + (qx >> kInterBits);
                const float ax = (float)(qx & (kInterTab - 1)) * (1.f / kInterTab);
                const float ay = (float)(qy & (kInterTab - 1)) * (1.f / kInterTab);
                const Sample smp = sample_box(bp, ax, ay);
                float g[G];
                g[0] = smp.gx2 * rw; g[1] = smp.gy2 * rw;                       // 2a, 2b
                g[G - 1] = fmaf(xf + du, g[0], (yf + dv) * g[1]);               // -2t
                if constexpr (CFG::kPack == 2) acc.add_packed(f2(g[0], g[1]), g[G - 1], smp.w, t_, yf);
                else acc.template add<true>(g, smp.w, t_, 1.f, yf);
              }
              trow += kUnr * kEccStripW;
              bp0 += kUnr * kBoxW;
              yf0 += (float)kUnr;
            }
          } else if constexpr (fast_coords) {
            // the magic offsets of the float->int splice are folded into the box index (see lean_pixel)
            constexpr unsigned kMagicHi = 0x4B400000u >> kInterBits;
            unsigned bi0 = (unsigned)((ya - ylo) * kBoxW + (x - xlo)) - kMagicHi * (unsigned)(kBoxW + 1);
            float yf0 = (float)ya;
#pragma unroll 1
            for (int rg = 0; rg < kRpt; rg += kUnr) {
#pragma unroll
              for (int r = 0; r < kUnr; ++r)
                lean_pixel<CFG::kBody>(fp, xf, yf0 + (float)r, trow[r * kEccStripW], box, bi0 + (unsigned)(r * kBoxW), frac_magic, acc);
              trow += kUnr * kEccStripW;
              bi0 += kUnr * kBoxW;
              yf0 += (float)kUnr;
            }
          } else if constexpr (affine_packed) {
            // Affine: OpenCV's exact 10-bit fixed-point coordinates, the packed sampler and the packed sums
            Coord<false> co;
            co.init(s_m, x);
#pragma unroll 1
            for (int rg = 0; rg < kRpt; rg += kUnr) {
#pragma unroll
              for (int r = 0; r < kUnr; ++r) {
                const int y = ya + rg + r;
                const float t_ = trow[(rg + r) * kEccStripW];
                int xq, yq;
                co.at(y, xq, yq);
                const float* bp = box + ((yq >> kInterBits) - ylo) * kBoxW + ((xq >> kInterBits) - xlo);
                // fractions k/32: the 5 low bits spliced under the 1.5 * 2^23 exponent, one packed FFMA
                const float2 axy = fma2(f2(__uint_as_float(((unsigned)xq & (kInterTab - 1)) | frac_magic),
                                           __uint_as_float(((unsigned)yq & (kInterTab - 1)) | frac_magic)),
                                        f2(1.f / kInterTab), f2(-12582912.0f / kInterTab));
                float w_;
                float2 gxy2;
                sample_box_packed(bp, axy.x, axy.y, w_, gxy2);
                acc.add_packed(gxy2, w_, t_, (float)y);
              }
            }
          } else {
            // the other instantiations (2x3 models with OpenCV's exact 10-bit fixed point, homography with exact
            // f64 coordinates): same shape, exact coordinates, scalar sums
            Coord<Md::persp> co;
            Jac<MOTION> jac;
            co.init(s_m, x);
            jac.init(s_m, xf);
#pragma unroll 1
            for (int rg = 0; rg < kRpt; rg += kUnr) {
#pragma unroll
              for (int r = 0; r < kUnr; ++r) {
                const int y = ya + rg + r;
                const float yf = (float)y;
                const float t_ = trow[(rg + r) * kEccStripW];
                int xq, yq;
                co.at(y, xq, yq);
                const float* bp = box + ((yq >> kInterBits) - ylo) * kBoxW + ((xq >> kInterBits) - xlo);
                const float ax = (float)(xq & (kInterTab - 1)) * (1.f / kInterTab);
                const float ay = (float)(yq & (kInterTab - 1)) * (1.f / kInterTab);
                const Sample smp = sample_box(bp, ax, ay);
                float g[G];
                jac.eval(smp, yf, g);
                acc.template add<true>(g, smp.w, t_, 1.f, yf);
              }
            }
          }
          n_safe += kRpt;
        } else if (boxed) {
          float yf = (float)ya;
          Coord<Md::persp> co;
          Jac<MOTION> jac;
          if (!fast_coords) { co.init(s_m, x); jac.init(s_m, xf); }
#pragma unroll 2
          for (int y = ya; y < yb; ++y, yf += 1.0f) {
            float g[G];
            Sample smp;
            const float t_ = trow[(y - ya) * kEccStripW];
            int sx, sy, xn = 0, yn = 0;   // integer sample position / nearest position in the image
            float ax, ay, du = 0.f, dv = 0.f, rw = 0.f;
            if (fast_coords) {
              int qx, qy;
              fp.at(yf, qx, qy, du, dv, rw);
              sx = x + (qx >> kInterBits);
              sy = y + (qy >> kInterBits);
              ax = (float)(qx & (kInterTab - 1)) * (1.f / kInterTab);
              ay = (float)(qy & (kInterTab - 1)) * (1.f / kInterTab);
            } else {
              int xq, yq;
              co.at_with_nearest(y, xq, yq, xn, yn);
              sx = xq >> kInterBits;
              sy = yq >> kInterBits;
              ax = (float)(xq & (kInterTab - 1)) * (1.f / kInterTab);
              ay = (float)(yq & (kInterTab - 1)) * (1.f / kInterTab);
            }
            // safe: the four bilinear taps and their gradient stencils stay off the border rows/columns, so no
            // border rule applies and the mask is 1.  The branch is taken warp-wide: a warp on the rim runs the
            // careful variant for all its lanes (no divergence), every other warp the plain one.
            const bool safe = (unsigned)(sx - 1) <= (unsigned)(p.width - 4) && (unsigned)(sy - 1) <= (unsigned)(p.height - 4);
            const float* bp = box + (sy - ylo) * kBoxW + (sx - xlo);
            const bool all_safe = __all_sync(wmask, safe);
            float mk = 1.f;
            if (all_safe) {
              smp = sample_box(bp, ax, ay);
            } else {
              smp = sample_box_rules(bp, ax, ay, sx, sy, p.width, p.height);
              if (fast_coords) { FastPersp::nearest(du, dv, xn, yn); xn += x; yn += y; }
              // OpenCV warps an all-ones mask with INTER_NEAREST: 1 where the rounded position is inside
              mk = ((unsigned)xn < (unsigned)p.width && (unsigned)yn < (unsigned)p.height) ? 1.f : 0.f;
            }
            if (fast_coords) {
              // a = gx / den, b = gy / den, t = hatX a + hatY b with hatX = -u, hatY = -v, den = w;
              // accumulated as (2a, 2b, -2t), see Accum::emit
              g[0] = smp.gx2 * rw; g[1] = smp.gy2 * rw;
              g[G - 1] = fmaf(xf + du, g[0], (yf + dv) * g[1]);
            } else {
              jac.eval(smp, yf, g);
              if (affine_packed) { g[0] *= 2.f; g[1] *= 2.f; }
            }
            if (all_safe) { acc.template add<true>(g, smp.w, t_, 1.f, yf); ++n_safe; }
            else acc.template add<false>(g, smp.w, t_, mk, yf);
          }
        } else {
          for (int y = ya; y < yb; ++y) slow_pixel(y, trow[(y - ya) * kEccStripW]);
        }
      }
      // stage release: the last warp to get here recycles the stage for chunk c + kStages of this batch
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();                       // this warp's reads of the stage are done before it is counted
        const int old = atomicAdd(&s_cnt[s], 1);
        if (old == kWarps - 1) {
          atomicExch(&s_cnt[s], 0);
          if (c + kStages < nb) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(c + kStages);
          }
        }
      }
    }
    cc += nb;
    g0 += nb;
  }
  if (cur_strip >= 0) fold();

  __syncthreads();
  if (tid < NV) {
    double sres = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) sres += s_acc[w][tid];
    s_accum[tid] = sres;
  }
  finish_iteration<MOTION, kThreads>(p, st, s_accum, s_tot, &s_last);
}

// ---- configurations ------------------------------------------------------------------------------------
// id 0 is the default; the others exist so that geometry is decided by measurement (STK_ECC_CFG=id,
// scripts/k2_variants.py) — they are instantiated for the Homography kernel only.
using EccCfg0 = EccCfg<256, 8, 8, 4, 2>;     // the first-generation geometry: 128x16 chunks, 2 blocks/SM, 8-row groups
using EccCfg1 = EccCfg<256, 8, 4, 2, 3>;     // 3 blocks/SM (<= 85 registers), 4-row groups, 2 stages (shared memory)
using EccCfg2 = EccCfg<256, 16, 8, 2, 2>;    // 128x32 chunks: half the per-chunk overhead
using EccCfg3 = EccCfg<768, 8, 4, 3, 1>;     // one 24-warp block per SM, 128x48 chunks
using EccCfg4 = EccCfg<512, 8, 8, 4, 1>;     // one 16-warp block per SM, 128x32 chunks
using EccCfg5 = EccCfg<384, 8, 4, 2, 2>;     // 2 x 12 warps per SM, 128x24 chunks
using EccCfg6 = EccCfg<256, 8, 4, 4, 2>;     // geometry 0 with 4-row groups (isolates the unroll depth)
using EccCfg7 = EccCfg<128, 8, 8, 2, 5>;     // five independent 4-warp blocks per SM, 128x8 chunks
using EccCfg8 = EccCfg<256, 16, 8, 2, 2, 0>; // geometry 2, SCALAR lean body
using EccCfg9 = EccCfg<256, 16, 8, 2, 2, 2>; // geometry 2, scalar sampling + packed sums
using EccCfg10 = EccCfg<256, 16, 4, 2, 3, 0>;// scalar body at 3 blocks/SM (<= 85 registers; chunk 128x32 needs 2 stages of 43.6 KB: 2 blocks by smem)
using EccCfg11 = EccCfg<256, 8, 4, 2, 3, 0>; // scalar body, 128x16 chunks, 3 blocks/SM
using EccCfg12 = EccCfg<256, 12, 4, 3, 2>;   // 128x24 chunks, THREE stages at 2 blocks/SM (106 KB): one more chunk of slack against warp skew
using EccCfg13 = EccCfg<256, 12, 6, 3, 2>;   // the same with 6-row groups
using EccCfg14 = EccCfg<256, 16, 8, 2, 2, 3>; // geometry 2 with the premultiplied accumulator (44 instead of 54 FP32-pipe cycles of sums per pixel)
using EccCfg15 = EccCfg<256, 8, 8, 4, 2, 3>;  // geometry 0 with the premultiplied accumulator
using EccCfg16 = EccCfg<256, 16, 8, 2, 2, 3, 1>; // cfg 14 with the leaner pixel body (fused LOP3, scalar (u, v))
using EccCfg17 = EccCfg<256, 16, 4, 2, 2, 3, 1>; // the same in 4-row groups
using EccCfg18 = EccCfg<512, 8, 8, 4, 1, 3, 1>;  // cfg 16's body in ONE 16-warp block per SM (all warps the same age: no old-block-first
                                                 // scheduling skew between two co-resident blocks), 128x32 chunks, 4 stages
using EccCfg19 = EccCfg<512, 16, 8, 2, 1, 3, 1>; // the same with 128x64 chunks, 2 stages
constexpr int kEccCfgCount = 20;

}  // namespace stk
