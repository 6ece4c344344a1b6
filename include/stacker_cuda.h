/*
 * stacker_cuda.h — C ABI of the B200 (sm_100a) align-and-stack library.
 *
 * This is the drop-in boundary for libstacker's ECC align-and-stack hot path.  The reference crate
 * (eadf/libstacker.rs) has no FFI of its own for this path: it reaches OpenCV through the `opencv`
 * crate's generated bindings.  Each entry point below replaces one group of those OpenCV calls; the
 * reference call site it stands for is cited as  file:line  into /root/reference.
 *
 * Conventions
 *   - every function returns an int status (STK_OK == 0); stk_last_error() returns a thread-local,
 *     human-readable description of the last failure on the calling thread.
 *   - the caller owns every host buffer; the library owns every device buffer.
 *   - plain pointers and sizes only; no C++/torch/OpenCV types.
 *   - images are 8-bit, interleaved, row-major, `pitch` bytes per row (cv::Mat::step), channel order
 *     as decoded by OpenCV (B,G,R[,A]).
 *   - a context is bound to ONE CUDA device.  Multi-GPU stacking = one context per device, each fed
 *     its shard of the frames, plus one sum-reduce of the partial stacks (NCCL) between
 *     stk_ecc_partial() and stk_ecc_finish_from() — see INTEGRATION.md.
 *   - there is NO CPU fallback: without a usable CUDA device every call fails with STK_ERR_CUDA.
 */
#ifndef STACKER_CUDA_H_
#define STACKER_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STK_ABI_VERSION 5

/* ---- status codes ------------------------------------------------------------------------- */
enum {
  STK_OK = 0,
  STK_ERR_BAD_ARG = 1,      /* -> StackerError::InvalidParams            (src/lib.rs:41-42)        */
  STK_ERR_CUDA = 2,         /* -> StackerError::ProcessingError          (src/lib.rs:43-44)        */
  STK_ERR_NOT_ENOUGH = 3,   /* -> StackerError::NotEnoughFiles           (src/lib.rs:31-32)        */
  STK_ERR_ECC_NOCONV = 4,   /* cv::Error::StsNoConv from findTransformECC -> StackerError::OpenCvError
                               (src/lib.rs:777 `?`): lambda denominator <= 0                       */
  STK_ERR_ECC_NAN = 5,      /* same, "NaN encountered"                                             */
  STK_ERR_CRITERIA = 6,     /* neither COUNT nor EPS set: CV_Assert in findTransformECC
                               -> StackerError::OpenCvError (src/utils.rs:159-170)                 */
  STK_ERR_STATE = 7,        /* call order violated (e.g. submit before set_reference)             */
  STK_ERR_UNSUPPORTED = 8,  /* -> StackerError::NotImplemented           (src/lib.rs:33-34)        */
  STK_ERR_NOMEM = 9
};

/* MotionType discriminants == opencv::video::MOTION_*            (src/lib.rs:603-609) */
enum { STK_MOTION_TRANSLATION = 0, STK_MOTION_EUCLIDEAN = 1, STK_MOTION_AFFINE = 2, STK_MOTION_HOMOGRAPHY = 3 };

/* TermCriteria_Type bits                                          (src/utils.rs:159-170) */
enum { STK_TERM_COUNT = 1, STK_TERM_EPS = 2 };

/* cv::BorderTypes accepted by the warp-only path (KeyPointMatchParameters::border_mode, src/lib.rs:66-68, :297-298).
   BORDER_TRANSPARENT (5) is STK_ERR_UNSUPPORTED: the reference warps into a fresh Mat, so "transparent" pixels would
   add uninitialised memory to the stack. */
enum { STK_BORDER_CONSTANT = 0, STK_BORDER_REPLICATE = 1, STK_BORDER_REFLECT = 2, STK_BORDER_WRAP = 3, STK_BORDER_REFLECT_101 = 4 };

typedef struct stk_ecc_ctx stk_ecc_ctx;

/* EccMatchParameters + TermCriteria + frame geometry            (src/lib.rs:611-623, src/utils.rs:159-170) */
typedef struct stk_ecc_config {
  int32_t width, height;      /* frame size; every frame of a stack has the size of frame 0        */
  int32_t channels;           /* 3 (BGR) or 4 (BGRA)                                               */
  int32_t motion_type;        /* STK_MOTION_*                                                      */
  int32_t criteria_type;      /* STK_TERM_COUNT | STK_TERM_EPS as built by src/utils.rs:161-168    */
  int32_t max_count;          /* used when COUNT is set, else OpenCV's 200                          */
  double  epsilon;            /* used when EPS is set, else -1 (run every iteration)                */
  int32_t gauss_filt_size;    /* odd, >= 1                                                         */
  int32_t device;             /* CUDA device ordinal, -1 = current device                          */
  int32_t lanes;              /* concurrent frame pipelines (streams) on the device; 0 = default   */
  int32_t seed_reference;     /* 1: the stack accumulator starts as the unwarped reference frame
                                 (src/lib.rs:752-754); 0 on the non-root shards of a multi-GPU stack */
  int32_t align;              /* 1: ECC alignment (ecc_match); 0: warp-only context (keypoint_match
                                 tail, src/lib.rs:289-350) — no reference planes are built          */
  int32_t ecc_width, ecc_height; /* 0,0: ECC runs on the full-size grey planes (ecc_match_no_scaling,
                                 src/lib.rs:719-847).  Otherwise ecc_match_scaling_down (src/lib.rs:849-1028):
                                 the grey planes are INTER_AREA-resized to this size (utils::scale_image,
                                 src/utils.rs:186-214 — use stk_scaled_size), ECC runs there, and the
                                 matrix is taken back to full resolution (src/lib.rs:941-958,
                                 src/utils.rs:218-248) before the full-size warp                     */
} stk_ecc_config;

/* per-frame outcome of findTransformECC                          (src/lib.rs:769-777) */
typedef struct stk_frame_result {
  int64_t tag;                /* caller's frame id                                                 */
  float   warp[9];            /* row-major 3x3; rows 0-1 are the 2x3 matrix for non-homography     */
  double  rho;                /* final correlation coefficient (the reference discards it)         */
  int32_t iterations;
  int32_t status;             /* STK_OK | STK_ERR_ECC_NOCONV | STK_ERR_ECC_NAN                     */
} stk_frame_result;

/* ---- library ------------------------------------------------------------------------------ */
int         stk_abi_version(void);
const char* stk_last_error(void);
int         stk_device_count(int* count);

/* utils::scale_image's size rule (src/utils.rs:186-200): the SMALLER dimension becomes `scale_down`, both
   new sizes are truncated (`as i32`).  Also applies ecc_match_scaling_down's validation
   (src/lib.rs:876-888): scale_down >= width or scale_down <= 10 -> STK_ERR_BAD_ARG.  On a landscape frame with
   height < scale_down < width the rule ENLARGES the planes, exactly as in the reference (cv::resize INTER_AREA then
   runs its 8-bit bilinear kernels in "area mode"; restated bit-exactly). */
int stk_scaled_size(int width, int height, float scale_down, int* scaled_width, int* scaled_height);

/* pinned host memory for decode targets: "JPEG decode stays on the host and feeds pinned,
   asynchronous uploads".  Replaces the Mat allocation inside imgcodecs::imread (src/utils.rs:132). */
int stk_pinned_alloc(void** ptr, size_t bytes);
int stk_pinned_free(void* ptr);

/* ---- ecc_match: context per (stack, device)            replaces src/lib.rs:719-847 ------------ */
int stk_ecc_create(const stk_ecc_config* cfg, stk_ecc_ctx** out);
int stk_ecc_destroy(stk_ecc_ctx* ctx);

/* frame 0: cvt_color + (inside find_transform_ecc) GaussianBlur + gradients of the INPUT image,
   done once per stack instead of once per call          (src/lib.rs:731-738, :769-772)
   *_device variants take a device pointer on the context's device.
   Asynchronous (no host synchronisation): lane 0 builds the plane behind whatever the lanes still have queued, the other
   lanes wait for it on the device.  A page-locked `bgr` is copied asynchronously and must stay unchanged until the next
   stk_ecc_sync / finish; a pageable one may be reused when the call returns. */
int stk_ecc_set_reference(stk_ecc_ctx* ctx, const uint8_t* bgr, size_t pitch);
int stk_ecc_set_reference_device(stk_ecc_ctx* ctx, const uint8_t* d_bgr, size_t pitch);

/* one non-reference frame: read_grey_and_f32's conversions, find_transform_ecc, warp_affine |
   warp_perspective, acc + warped                         (src/lib.rs:756-814)
   Asynchronous: returns once the frame is queued on a lane.  The host buffer may be reused as soon as
   the call returns unless it is pinned (stk_pinned_alloc), in which case it must stay valid until
   stk_ecc_sync()/finish.  Thread-safe.
   The frame's final warp + accumulate is DEFERRED: it is launched together with up to three other frames of the
   same lane (one pass over the f32 accumulator for four frames, in submission order — bit-identical to one launch
   per frame) when four are waiting, or at the next sync / results / partial / finish / exchange.  A DEVICE buffer
   (_device variants) must therefore stay valid and unmodified until one of those calls has returned, and whoever
   produced it must have finished writing it before the submit call (the library's streams are non-blocking), unless the
   producer's stream was announced with stk_ecc_set_input_stream. */
/* Stream contract for DEVICE buffers (ABI v5).  The library's lanes are non-blocking streams: by default nothing orders
   them behind the stream that produced a device frame.  After stk_ecc_set_input_stream(ctx, s, 1) every later *_device
   call (set_reference_device, submit_frame_device, submit_warp*_device) first records an event on `s` (a cudaStream_t
   of the context's device, NULL = the legacy default stream) and makes the lane that takes the frame wait for it — a
   device-side dependency, the host never blocks.  enabled = 0 restores the default.  This is what lets a torch / CuPy /
   NPP producer hand over a tensor it has only just queued the writes of (the Python mirror passes torch's current
   stream automatically).  Replaces nothing in the reference, whose Mats live on the host. */
int stk_ecc_set_input_stream(stk_ecc_ctx* ctx, void* cuda_stream, int enabled);

int stk_ecc_submit_frame(stk_ecc_ctx* ctx, const uint8_t* bgr, size_t pitch, int64_t tag);
int stk_ecc_submit_frame_pinned(stk_ecc_ctx* ctx, const uint8_t* pinned_bgr, size_t pitch, int64_t tag);
int stk_ecc_submit_frame_device(stk_ecc_ctx* ctx, const uint8_t* d_bgr, size_t pitch, int64_t tag);

/* Host feed for decode threads (replaces the Mat that imgcodecs::imread allocates, src/utils.rs:132, as the
   decode target): the context owns a ring of pinned frame buffers (2 per lane, dense rows, *pitch =
   width*channels).  acquire blocks until one is free; decode (imdecode_to / memcpy) into it from any thread
   without holding any library lock; submit_acquired queues the asynchronous upload + alignment and recycles
   the buffer when the upload has landed (do not touch it after the call); release gives an unused one back.
   stk_ecc_submit_frame on a pageable buffer is exactly acquire + row copy + submit_acquired. */
int stk_ecc_acquire_frame_buffer(stk_ecc_ctx* ctx, uint8_t** buf, size_t* pitch);
int stk_ecc_submit_acquired(stk_ecc_ctx* ctx, uint8_t* buf, int64_t tag);
int stk_ecc_release_frame_buffer(stk_ecc_ctx* ctx, uint8_t* buf);

/* keypoint_match tail: warp_perspective(img_f32, H, size, INTER_LINEAR, border_mode, border_value)
   + accumulate                                           (src/lib.rs:289-316)
   h is the 3x3 f64 matrix from find_homography (forward map, inverted internally like OpenCV). */
int stk_ecc_submit_warp(stk_ecc_ctx* ctx, const uint8_t* bgr, size_t pitch, const double h[9],
                        int border_mode, const double border_value[4], int64_t tag);
int stk_ecc_submit_warp_device(stk_ecc_ctx* ctx, const uint8_t* d_bgr, size_t pitch, const double h[9],
                               int border_mode, const double border_value[4], int64_t tag);

/* The 2x3 form: warp_affine(img_f32, M, size, INTER_LINEAR, border_mode, border_value) + accumulate — the final warp
   ecc_match applies for Translation / Euclidean / Affine (src/lib.rs:782-790), with a caller-supplied matrix.
   m is the row-major 2x3 f64 forward map (inverted internally like cv::warpAffine; coordinates in OpenCV's 10-bit
   fixed point).  Bit-identical to cv2.warpAffine on the CV_32F frame. */
int stk_ecc_submit_warp_affine(stk_ecc_ctx* ctx, const uint8_t* bgr, size_t pitch, const double m[6],
                               int border_mode, const double border_value[4], int64_t tag);
int stk_ecc_submit_warp_affine_device(stk_ecc_ctx* ctx, const uint8_t* d_bgr, size_t pitch, const double m[6],
                                      int border_mode, const double border_value[4], int64_t tag);

/* wait for every queued frame; returns the first per-frame error (STK_ERR_ECC_*) or STK_OK */
int stk_ecc_sync(stk_ecc_ctx* ctx);

/* per-frame results in submission order; *count receives how many were written (<= capacity) */
int stk_ecc_results(stk_ecc_ctx* ctx, stk_frame_result* out, int capacity, int* count);

/* Rayon try_reduce of the per-thread partial sums + `/ n`  (src/lib.rs:819-839, :339-346)
   stk_ecc_finish   : single-device stack: sum lanes, scale by 1/divisor, copy to host (out_pitch
                      bytes per row, width*channels floats per row).
   stk_ecc_partial  : multi-device: sum lanes in place and hand out the device pointer of this
                      device's partial stack (height*width*channels contiguous floats) so the caller's
                      plumbing can sum-reduce it across devices (ncclReduce / torch.distributed).
   stk_ecc_finish_from : scale `d_sum` (device, same layout; NULL = this context's partial) by
                      1/divisor and copy to host. */
int stk_ecc_finish(stk_ecc_ctx* ctx, int divisor, float* out, size_t out_pitch);
int stk_ecc_partial(stk_ecc_ctx* ctx, float** d_partial, size_t* n_floats);
int stk_ecc_finish_from(stk_ecc_ctx* ctx, const float* d_sum, int divisor, float* out, size_t out_pitch);
/* same, result left on the device (d_out: height*width*channels floats) */
int stk_ecc_finish_device(stk_ecc_ctx* ctx, const float* d_sum, int divisor, float* d_out);

/* ---- multi-GPU: the ONE exchange step, fused with the divide, over NVLink peer memory -----------------------
   Replaces Rayon's try_reduce of the partial sums + MatExpr `/ n` (src/lib.rs:819-839, :319-346) when the
   frames of a stack were sharded over several devices (one context per device, rank 0 = the context created with
   seed_reference = 1).  Reduce-scatter style: rank r sums slice r of EVERY rank's partial stack with peer loads
   over NVLink/NVSwitch (rank order, deterministic), scales by 1/divisor and stores the finished pixels straight
   into rank 0's output buffer; no library collective, no separate scale pass (csrc/peer_reduce.cuh).

   One process per GPU:  every rank calls stk_ecc_peer_export, the caller's plumbing all-gathers the handles
   (plain bytes: torch.distributed / MPI / a pipe), every rank calls stk_ecc_peer_connect with all of them.
   One process, several GPUs (the Rust crate): stk_ecc_peer_connect_local with the contexts in rank order.
   Then, per stack, EVERY rank calls stk_ecc_peer_reduce after submitting its frames (collective: same number of
   calls on every rank).  It is asynchronous and needs no host synchronisation: the lanes are joined on the device,
   summed, exchanged.  On rank 0 *d_out receives the device pointer of the finished stack (height*width*channels
   floats, library-owned, valid after stk_ecc_sync until the next peer_reduce); elsewhere NULL.  stk_ecc_sync
   afterwards returns per-frame ECC errors, and STK_ERR_CUDA if a rank did not show up within
   STK_PEER_TIMEOUT_MS (environment, default 30000) — the kernels never spin forever; raise it when ranks can reach the
   exchange further apart than that (e.g. decode imbalance).  connect resets the exchange step count and the flag
   block of the context: no rank may start an exchange before EVERY rank has returned from connect (a barrier or, as
   distributed.connect_peers does, a vote all ranks take part in).
   The exchange runs on a stream of its own behind every lane's queued work: a rank that is early waits in a one-warp
   kernel, and the frames of the NEXT stack may be submitted (after stk_ecc_reset + stk_ecc_set_reference) while it is in
   flight.  A world of ONE is legal (connect with the context's own handle): the exchange is then the lane sum fused with
   the divide, and the stacks of a single device queue back to back the same way. */
typedef struct stk_peer_handle { unsigned char bytes[256]; } stk_peer_handle;
int stk_ecc_peer_export(stk_ecc_ctx* ctx, stk_peer_handle* out);
int stk_ecc_peer_connect(stk_ecc_ctx* ctx, int rank, int world, const stk_peer_handle* handles /* [world] */);
int stk_ecc_peer_connect_local(stk_ecc_ctx* const* ctxs /* [world], rank order */, int world);
int stk_ecc_peer_reduce(stk_ecc_ctx* ctx, int divisor, const float** d_out);
/* Same exchange, but every rank KEEPS the finished pixels of its own slice (reduce-scatter): *d_slice = device
   pointer of the slice, [*begin, *begin + *count) its position in the height*width*channels stack.  For callers
   that want the stack in HOST memory: stk_ecc_peer_slice_to_host then queues the device-to-host copy of the slice
   into `out + begin` (`out` = the dense host stack, ideally pinned and, with one process per GPU, a shared
   mapping), so the result leaves over EVERY GPU's PCIe link at once instead of the root's alone (the copy of a
   4K stack is 2 ms over one link — at 8 GPUs as long as the alignment itself).  Complete after stk_ecc_sync.  With several contexts in ONE process, queue the
   reduce_scatter on every context before the first slice_to_host: an exchange waits for all ranks, and a copy into
   pageable host memory blocks the calling thread until its device's exchange is done. */
int stk_ecc_peer_reduce_scatter(stk_ecc_ctx* ctx, int divisor, const float** d_slice, size_t* begin, size_t* count);
int stk_ecc_peer_slice_to_host(stk_ecc_ctx* ctx, float* out);
int stk_ecc_peer_disconnect(stk_ecc_ctx* ctx);

/* start a new stack on the same context (same geometry/parameters): clears accumulators/results.
   Asynchronous (ABI v5): nothing waits on the host.  Whatever the previous stack still has queued — frames, an exchange,
   a slice copy — completes in stream order; the next stk_ecc_set_reference joins the lanes on the device before it
   overwrites the reference plane, and the next stack's first accumulator writes wait for an exchange in flight.  Results
   and the output of the previous stack must be read (stk_ecc_results / stk_ecc_sync + the exchange's pointer) before the
   NEXT exchange overwrites them, not before reset. */
int stk_ecc_reset(stk_ecc_ctx* ctx);

/* counters for benchmarks: kernels launched by this context since creation / last reset */
int stk_ecc_launch_count(stk_ecc_ctx* ctx, int64_t* launches);
/* per-stage device times.  With profiling on, every align submission is bracketed by CUDA events on its
   lane's stream: [prep | device ECC loop | warp+accumulate].  stk_ecc_stage_times (after a sync) returns the
   summed milliseconds per stage, the number of frames and the total ECC iterations they cover. */
int stk_ecc_set_profiling(stk_ecc_ctx* ctx, int enabled);
int stk_ecc_stage_times(stk_ecc_ctx* ctx, double ms[3], int64_t* frames, int64_t* iterations);

/* ---- single stages, exposed for verification and for callers that want one step only --------------- */
/* cvt_color(BGR2GRAY) + convertTo(f32) + GaussianBlur(k x k, sigma 0): the template/input plane
   findTransformECC builds from an 8-bit frame (src/utils.rs:136-142 + src/lib.rs:769).  Host in, host out
   (out_pitch bytes per row, width floats). */
int stk_prep_grey_blur(const uint8_t* bgr, size_t pitch, int width, int height, int channels, int ksize,
                       int device, float* out, size_t out_pitch);
/* cvt_color(BGR2GRAY) (skipped for channels == 1) + resize(INTER_AREA) to out_width x out_height: the plane
   utils::scale_image builds from the grey frame (src/utils.rs:186-214).  Host in, host out, 8-bit. */
int stk_grey_resize_area(const uint8_t* img, size_t pitch, int width, int height, int channels, int out_width,
                         int out_height, int device, uint8_t* out, size_t out_pitch);
/* ONE ECC iteration of `frame` (host, 8-bit) against the context's reference, starting from `warp_in`
   (row-major 3x3 f32): returns the kernel's reduced sums (layout documented in csrc/ecc_iter.cuh,
   *nv values written, capacity `cap`), the updated warp, rho and the loop status.  Does not touch the
   accumulators.  Test hook for the parity suite. */
int stk_ecc_debug_iteration(stk_ecc_ctx* ctx, const uint8_t* bgr, size_t pitch, const float warp_in[9],
                            double* totals, int cap, int* nv, float warp_out[9], double* rho, int* status);

/* Same call, but instead of the sums it returns %globaltimer stamps (ns) of the iteration kernel:
   for each tile [start, pixels done, partial written, -] then [cross-tile sum done, solve done, end,
   last block id]; *n_tiles receives the tile count.  `iters` launches are made back to back and the
   stamps of the last one are returned.  Profiling hook. */
int stk_ecc_debug_timing(stk_ecc_ctx* ctx, const uint8_t* bgr, size_t pitch, const float warp_in[9], int iters,
                         uint64_t* stamps, int cap, int* n_tiles);

/* ---- sharpness_tenengrad                               replaces src/lib.rs:1101-1147 ---------- */
/* grey: single-channel 8-bit (channels == 1) — or BGR/BGRA (channels 3/4), converted with
   cvt_color(BGR2GRAY) on the fly.  ksize in {1,3,5,7} else STK_ERR_BAD_ARG (src/lib.rs:1103-1107).
   *out = mean(gx^2 + gy^2), bit-identical to the CV_64F OpenCV pipeline. */
int stk_tenengrad(const uint8_t* img, size_t pitch, int width, int height, int channels, int ksize,
                  int device, double* out);
int stk_tenengrad_device(const uint8_t* d_img, size_t pitch, int width, int height, int channels,
                         int ksize, int device, double* out);
/* n same-sized frames resident on the device, frame i at d_imgs + i*frame_stride; out[n] */
int stk_tenengrad_batch_device(const uint8_t* d_imgs, size_t frame_stride, size_t pitch, int width,
                               int height, int channels, int ksize, int n, int device, double* out);

/* ---- the crate's other sharpness metrics, fused with Tenengrad(3) into ONE pass over the plane ---------
   out[4] = { LAPM  sharpness_modified_laplacian              src/lib.rs:1032-1068,
              LAPV  sharpness_variance_of_laplacian           src/lib.rs:1070-1090,
              TENG  sharpness_tenengrad(k_size = 3)           src/lib.rs:1101-1147,
              GLVN  sharpness_normalized_gray_level_variance  src/lib.rs:1151-1166 }
   — the order examples/main.rs:43-46 evaluates them in.  Same input rules as stk_tenengrad; every value is
   bit-identical to the CV_64F OpenCV pipeline on 8-bit input. */
int stk_sharpness_all(const uint8_t* img, size_t pitch, int width, int height, int channels, int device,
                      double out[4]);
/* n same-sized frames resident on the device, frame i at d_imgs + i*frame_stride; out[n*4] */
int stk_sharpness_all_batch_device(const uint8_t* d_imgs, size_t frame_stride, size_t pitch, int width, int height,
                                   int channels, int n, int device, double* out);

#ifdef __cplusplus
}
#endif
#endif /* STACKER_CUDA_H_ */
