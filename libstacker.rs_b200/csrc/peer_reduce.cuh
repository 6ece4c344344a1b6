// Multi-GPU exchange step of one stack, fused with the final divide: the B200 form of the reference's Rayon
// `try_reduce` of the per-task partial sums followed by MatExpr `/ n` (/root/reference/src/lib.rs:819-839;
// keypoint_match: :319-346).
//
// Every rank (one context per GPU; one process per GPU or all in one process) holds a partial stack.  Instead of
// a library reduce followed by a scale kernel on the root, ONE kernel per rank does both over NVLink peer
// memory, reduce-scatter style:
//
//   rank r owns slice r of the stack.  It (1) announces "my partial is complete" by a release store of the step
//   number into every peer's flag block, (2) waits until every peer has announced, (3) reads slice r of EVERY
//   rank's partial (peer loads through NVSwitch, 16 bytes per lane per peer, all peers in flight together), adds
//   them in rank order (deterministic, independent of which rank does it), multiplies by 1/n and (4) stores the
//   finished pixels straight into the ROOT's output buffer (peer stores), then (5) announces "slice r done".
//
// Traffic: a worker's inbound side carries the remote partials of its slice, the root's inbound side carries the
// finished stack (S bytes, irreducible when the result must end on the root).  With more than two ranks the root
// therefore takes NO slice (host side, stk_ecc_peer_reduce): a root slice would add (world-1) remote reads per
// pixel on the links that already carry the result — measured at world 4 with equal slices: 260 us, the root's
// inbound side carrying 1.5 S.  Without it every rank's inbound traffic is S, all ranks in parallel, and there is
// no separate 24N-byte scale pass on the root.  `peer_wait_done_kernel` closes the step on every rank: the root's
// output is complete and nobody still reads this rank's partial.
//
// Flags are 32-bit step counters (monotonic, wrap-safe compare) in a small cudaMalloc'ed block per context that
// peers map (CUDA IPC between processes, peer access inside one process).  Spins are bounded by %globaltimer:
// a missing peer sets kPeerError in the local block instead of hanging the GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ecc_iter.cuh"   // global_ns()

namespace stk {

constexpr int kMaxPeers = 16;
// layout of a flag block (uint32_t words)
constexpr int kPeerReady = 0;                 // [kMaxPeers] ready[p]: rank p's partial of step s is complete
constexpr int kPeerDone = kMaxPeers;          // [kMaxPeers] done[p] : rank p finished its slice of step s
constexpr int kPeerError = 2 * kMaxPeers;     // set to the step number when a spin timed out
constexpr int kPeerCounter = 2 * kMaxPeers + 1;   // block counter of the local reduce kernel
constexpr int kPeerFlagWords = 64;

struct PeerReduceParams {
  const float* partial[kMaxPeers];   // partial stack of every rank, index = rank (own entry: local pointer)
  uint32_t* flags[kMaxPeers];        // flag block of every rank
  float* out;                        // the root's output buffer as mapped on this device
  size_t begin, end;                 // this rank's slice, in floats; begin % 4 == 0
  int rank, world;
  uint32_t step;
  float scale;                       // float(1 / divisor), as the reference's MatExpr `/ n` evaluates it
  unsigned long long timeout_ns;
  int pre_waited;                    // 1: peer_announce_wait_kernel already announced and waited on this stream
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// true when *flag has reached `step` (wrap-safe) before the deadline
__device__ __forceinline__ bool spin_until(const uint32_t* flag, uint32_t step, unsigned long long timeout_ns) {
  if ((int32_t)(ld_acquire_sys(flag) - step) >= 0) return true;
  const unsigned long long t0 = global_ns();
  for (;;) {
    for (int i = 0; i < 64; ++i) {
      if ((int32_t)(ld_acquire_sys(flag) - step) >= 0) return true;
      __nanosleep(40);
    }
    if (global_ns() - t0 > timeout_ns) return false;
  }
}

__device__ __forceinline__ float4 ld_stream4(const float* p) {
  // read-once peer data: bypass L1 (coherent at the owner's L2), do not pollute
  return __ldcg(reinterpret_cast<const float4*>(p));
}

template <int WORLD>   // 0 = run-time world size
__global__ void __launch_bounds__(256) peer_reduce_scale_kernel(const PeerReduceParams p) {
  const int world = WORLD > 0 ? WORLD : p.world;
  uint32_t* mine = p.flags[p.rank];
  int ok = 1;
  if (p.pre_waited) {
    // announced and waited for by the one-warp kernel ahead of this one; a time-out there voids the step
    if (threadIdx.x == 0) ok = ld_acquire_sys(mine + kPeerError) == p.step ? 0 : 1;
  } else {
    // (1) my partial was completed by earlier work on this stream: tell everyone (block 0 only)
    if (blockIdx.x == 0 && threadIdx.x < world) {
      __threadfence_system();
      st_release_sys(p.flags[threadIdx.x] + kPeerReady + p.rank, p.step);
    }
    // (2) every block waits for every rank's announcement in the LOCAL flag block
    if (threadIdx.x < world) ok = spin_until(mine + kPeerReady + threadIdx.x, p.step, p.timeout_ns) ? 1 : 0;
  }
  ok = __syncthreads_and(ok);
  if (!ok) {
    if (threadIdx.x == 0) mine[kPeerError] = p.step;
  } else {
    // (3)+(4) slice: sum in rank order, scale, store into the root's buffer
    const size_t n4 = (p.end - p.begin) / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float* out = p.out + p.begin;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if constexpr (WORLD > 0) {
      // two float4 per thread and rank in flight (2*WORLD independent 16-byte loads, most of them over NVLink)
      const float* base[WORLD];
#pragma unroll
      for (int r = 0; r < WORLD; ++r) base[r] = p.partial[r] + p.begin;
      for (; i + stride < n4; i += 2 * stride) {
        float4 a[WORLD], b[WORLD];
#pragma unroll
        for (int r = 0; r < WORLD; ++r) { a[r] = ld_stream4(base[r] + 4 * i); b[r] = ld_stream4(base[r] + 4 * (i + stride)); }
        float4 s = a[0], t = b[0];
#pragma unroll
        for (int r = 1; r < WORLD; ++r) {
          s.x = __fadd_rn(s.x, a[r].x); s.y = __fadd_rn(s.y, a[r].y); s.z = __fadd_rn(s.z, a[r].z); s.w = __fadd_rn(s.w, a[r].w);
          t.x = __fadd_rn(t.x, b[r].x); t.y = __fadd_rn(t.y, b[r].y); t.z = __fadd_rn(t.z, b[r].z); t.w = __fadd_rn(t.w, b[r].w);
        }
        s.x = __fmul_rn(s.x, p.scale); s.y = __fmul_rn(s.y, p.scale); s.z = __fmul_rn(s.z, p.scale); s.w = __fmul_rn(s.w, p.scale);
        t.x = __fmul_rn(t.x, p.scale); t.y = __fmul_rn(t.y, p.scale); t.z = __fmul_rn(t.z, p.scale); t.w = __fmul_rn(t.w, p.scale);
        reinterpret_cast<float4*>(out)[i] = s;
        reinterpret_cast<float4*>(out)[i + stride] = t;
      }
    }
    for (; i < n4; i += stride) {
      float4 s = ld_stream4(p.partial[0] + p.begin + 4 * i);
      for (int r = 1; r < world; ++r) {
        const float4 a = ld_stream4(p.partial[r] + p.begin + 4 * i);
        s.x = __fadd_rn(s.x, a.x); s.y = __fadd_rn(s.y, a.y); s.z = __fadd_rn(s.z, a.z); s.w = __fadd_rn(s.w, a.w);
      }
      s.x = __fmul_rn(s.x, p.scale); s.y = __fmul_rn(s.y, p.scale); s.z = __fmul_rn(s.z, p.scale); s.w = __fmul_rn(s.w, p.scale);
      reinterpret_cast<float4*>(out)[i] = s;
    }
    // scalar tail of the slice (only the last rank's slice can have one)
    for (size_t k = p.begin + n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < p.end; k += stride) {
      float s = __ldcg(p.partial[0] + k);
      for (int r = 1; r < world; ++r) s = __fadd_rn(s, __ldcg(p.partial[r] + k));
      p.out[k] = __fmul_rn(s, p.scale);
    }
  }
  // (5) last block out announces "slice done" to every rank
  __threadfence_system();
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(mine + kPeerCounter, 1u);
    s_last = prev == gridDim.x - 1;
    if (s_last) mine[kPeerCounter] = 0;
  }
  __syncthreads();
  if (s_last && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + kPeerDone + p.rank, p.step);
  }
}

// (1) + (2) of the exchange as a kernel of its own (one warp): announce "my partial is complete" to every rank, then wait
// until every rank has.  A rank that reaches the exchange early spins HERE — 32 threads, not 592 blocks — while its lanes
// already work on the next stack; the reduce kernel behind it on the stream starts with every partial complete.
__global__ void peer_announce_wait_kernel(const PeerReduceParams p) {
  uint32_t* mine = p.flags[p.rank];
  if (threadIdx.x < p.world) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + kPeerReady + p.rank, p.step);
    if (!spin_until(mine + kPeerReady + threadIdx.x, p.step, p.timeout_ns)) mine[kPeerError] = p.step;
  }
}

// closes the step on this rank: every slice has landed in the root's buffer and no peer reads this rank's
// partial any more, so later work on the stream may overwrite it
__global__ void peer_wait_done_kernel(uint32_t* mine, int world, uint32_t step, unsigned long long timeout_ns) {
  if (threadIdx.x < world)
    if (!spin_until(mine + kPeerDone + threadIdx.x, step, timeout_ns)) mine[kPeerError] = step;
}

}  // namespace stk
