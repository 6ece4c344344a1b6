// C++ host-side mirror of the libstacker crate's public API for the ECC align-and-stack path
// (/root/reference/src/lib.rs), on top of the C ABI in include/stacker_cuda.h.
//
// The reference's host language is Rust (see ../rust/ for the crate that binds the same ABI); Rust is not
// available in this image, so this mirror is what gets compiled and self-tested here.  Same names, argument
// meaning and error behaviour:
//   ecc_match            src/lib.rs:702-847      EccMatchParameters  src/lib.rs:611-623
//   keypoint_match tail  src/lib.rs:289-350      MotionType          src/lib.rs:603-609
//   sharpness_tenengrad  src/lib.rs:1101-1147    StackerError        src/lib.rs:27-45
// Decode stays on the host like in the reference (imgcodecs::imread, src/utils.rs:128-144).  C++ OpenCV is
// not installed here, so the decoder is a callback; the built-in one reads binary PNM (P6 -> BGR, P5).
#pragma once
#include <array>
#include <atomic>
#include <cstdint>
#include <filesystem>
#include <functional>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

namespace libstacker {

// ---- errors: one exception type per StackerError variant ---------------------------------------------
struct StackerError : std::runtime_error { using std::runtime_error::runtime_error; };
struct OpenCvError : StackerError { using StackerError::StackerError; };          // incl. ECC StsNoConv
struct NotEnoughFiles : StackerError { NotEnoughFiles() : StackerError("Not enough files") {} };
struct NotImplemented : StackerError { using StackerError::StackerError; };
struct IoError : StackerError { using StackerError::StackerError; };
struct InvalidParams : StackerError { explicit InvalidParams(const std::string& m) : StackerError("Invalid parameter(s) " + m) {} };
struct ProcessingError : StackerError { explicit ProcessingError(const std::string& m) : StackerError("Internal error " + m) {} };

// ---- parameter types -------------------------------------------------------------------------------
enum class MotionType : int { Translation = 0, Euclidean = 1, Affine = 2, Homography = 3 };   // OpenCV MOTION_*

struct EccMatchParameters {
  MotionType motion_type;
  std::optional<int> max_count;      // TermCriteria::max_count, COUNT flag when set
  std::optional<double> epsilon;     // TermCriteria::epsilon,   EPS flag when set
  int gauss_filt_size;
};

struct TermCriteria { int typ = 0; int max_count = 0; double epsilon = 0.0; };
enum : int { TERM_COUNT = 1, TERM_EPS = 2 };
TermCriteria term_criteria(const EccMatchParameters& p);     // src/utils.rs:159-170

struct KeyPointMatchParameters {     // defaults: src/utils.rs:250-261
  int method = 8;                    // calib3d::RANSAC
  double ransac_reproj_threshold = 3.0;
  float match_keep_ratio = 0.75f;
  float match_ratio = 0.8f;
  int border_mode = 0;               // core::BORDER_CONSTANT
  double border_value[4] = {0, 0, 0, 0};
};

// ---- images ----------------------------------------------------------------------------------------
struct ImageU8 {                     // decoded frame, interleaved, B,G,R[,A] order like cv::Mat from imread
  int width = 0, height = 0, channels = 0;
  std::vector<uint8_t> data;
};
struct ImageF32 {                    // the stacked result: CV_32FC3 in [0,1]
  int width = 0, height = 0, channels = 0;
  std::vector<float> data;
};
using Decoder = std::function<ImageU8(const std::filesystem::path&)>;
ImageU8 read_pnm(const std::filesystem::path& path);          // P6 (RGB -> stored as BGR) and P5

struct FrameAlignment { int64_t frame; float warp[9]; double rho; int iterations; };   // frame = index into `files`

// ---- the API ---------------------------------------------------------------------------------------
// scale_down_width: Some(width) = ecc_match_scaling_down (src/lib.rs:849-1028): ECC on INTER_AREA-downscaled
// greys, matrix rescaled to full resolution, full-size warp.
ImageF32 ecc_match(const std::vector<std::filesystem::path>& files, const EccMatchParameters& params,
                   std::optional<float> scale_down_width = std::nullopt, const Decoder& decode = read_pnm,
                   int device = -1, std::vector<FrameAlignment>* details = nullptr);

// The same stack sharded over several GPUs of one box from ONE process (what the Rust crate does with its Rayon
// pool): one context per device, frames dealt round-robin, one fused exchange + divide over NVLink peer memory
// (stk_ecc_peer_connect_local / stk_ecc_peer_reduce_scatter), every device copying its slice of the result out.
std::vector<int> all_devices();
ImageF32 ecc_match_on_devices(const std::vector<std::filesystem::path>& files, const EccMatchParameters& params,
                              const std::vector<int>& devices, std::optional<float> scale_down_width = std::nullopt,
                              const Decoder& decode = read_pnm, std::vector<FrameAlignment>* details = nullptr);

// The GPU tail of keypoint_match (src/lib.rs:289-350): frames[0] unwarped + warp_perspective(frames[i],
// homographies[i-1]) accumulated, divided by the frame count.  The feature stages stay OpenCV on the host.
ImageF32 stack_with_homographies(const std::vector<ImageU8>& frames, const std::vector<std::array<double, 9>>& homographies,
                                 const KeyPointMatchParameters& params = {}, int device = -1);

double sharpness_tenengrad(const ImageU8& grey, int k_size, int device = -1);
// {LAPM, LAPV, TENG(3), GLVN} in one pass (examples/main.rs:43-46), and the crate's three single-metric entry
// points (src/lib.rs:1032-1090, :1151-1166) on top of it
std::array<double, 4> sharpness_all(const ImageU8& grey, int device = -1);
double sharpness_modified_laplacian(const ImageU8& grey, int device = -1);
double sharpness_variance_of_laplacian(const ImageU8& grey, int device = -1);
double sharpness_normalized_gray_level_variance(const ImageU8& grey, int device = -1);

}  // namespace libstacker
