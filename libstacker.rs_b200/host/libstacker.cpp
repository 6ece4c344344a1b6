#include <array>
#include "libstacker.hpp"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <future>
#include <memory>
#include <thread>

#include "../../include/stacker_cuda.h"

namespace libstacker {

namespace {
[[noreturn]] void raise(int rc) {
  const std::string msg = stk_last_error();
  switch (rc) {
    case STK_ERR_ECC_NOCONV: case STK_ERR_ECC_NAN: case STK_ERR_CRITERIA: throw OpenCvError(msg);
    case STK_ERR_BAD_ARG: throw InvalidParams(msg);
    case STK_ERR_NOT_ENOUGH: throw NotEnoughFiles();
    case STK_ERR_UNSUPPORTED: throw NotImplemented(msg);
    default: throw ProcessingError(msg);
  }
}
void check(int rc) { if (rc != STK_OK) raise(rc); }

struct Ctx {                                   // RAII over stk_ecc_ctx
  stk_ecc_ctx* c = nullptr;
  explicit Ctx(const stk_ecc_config& cfg) { check(stk_ecc_create(&cfg, &c)); }
  ~Ctx() { stk_ecc_destroy(c); }
  Ctx(const Ctx&) = delete;
  Ctx& operator=(const Ctx&) = delete;
};

void check_colour(const ImageU8& f) {
  if (f.channels != 3 && f.channels != 4) throw OpenCvError("cvtColor(BGR2GRAY): input must have 3 or 4 channels");
  if (f.data.size() != (size_t)f.width * f.height * f.channels) throw OpenCvError("malformed frame");
}
}  // namespace

TermCriteria term_criteria(const EccMatchParameters& p) {
  TermCriteria t;
  if (p.max_count) { t.typ |= TERM_COUNT; t.max_count = *p.max_count; }
  if (p.epsilon) { t.typ |= TERM_EPS; t.epsilon = *p.epsilon; }
  return t;
}

ImageU8 read_pnm(const std::filesystem::path& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw OpenCvError("imread failed for " + path.string());   // the reference fails in the first OpenCV call
  std::string magic;
  in >> magic;
  auto next_int = [&]() {
    int v; char ch;
    while (in >> std::ws && in.peek() == '#') while (in.get(ch) && ch != '\n') {}
    if (!(in >> v)) throw OpenCvError("bad PNM header in " + path.string());
    return v;
  };
  if (magic != "P6" && magic != "P5") throw OpenCvError("unsupported image format in " + path.string());
  ImageU8 img;
  img.width = next_int(); img.height = next_int();
  const int maxv = next_int();
  if (maxv != 255) throw OpenCvError("findTransformECC: images must have 8uC1 type (16-bit PNM given)");
  in.get();
  img.channels = magic == "P6" ? 3 : 1;
  img.data.resize((size_t)img.width * img.height * img.channels);
  in.read(reinterpret_cast<char*>(img.data.data()), (std::streamsize)img.data.size());
  if (!in) throw OpenCvError("truncated PNM " + path.string());
  if (img.channels == 3) for (size_t i = 0; i + 2 < img.data.size(); i += 3) std::swap(img.data[i], img.data[i + 2]);  // RGB -> BGR
  return img;
}

ImageF32 ecc_match(const std::vector<std::filesystem::path>& files, const EccMatchParameters& params,
                   std::optional<float> scale_down_width, const Decoder& decode, int device,
                   std::vector<FrameAlignment>* details) {
  return ecc_match_on_devices(files, params, {device}, scale_down_width, decode, details);
}

std::vector<int> all_devices() {
  int n = 0;
  check(stk_device_count(&n));
  std::vector<int> d(n);
  for (int i = 0; i < n; ++i) d[i] = i;
  return d;
}

// One context per device.  Frame 0 goes to every device (each builds its own reference planes; only the first
// seeds its accumulator with it, src/lib.rs:752-754), the other frames are dealt round-robin, and the partial stacks
// meet in ONE exchange step fused with the `/ n` (src/lib.rs:819-839): stk_ecc_peer_reduce_scatter over NVLink
// peer memory, every device then copying its slice of the result into `out` over its own PCIe link.
ImageF32 ecc_match_on_devices(const std::vector<std::filesystem::path>& files, const EccMatchParameters& params,
                              const std::vector<int>& devices, std::optional<float> scale_down_width,
                              const Decoder& decode, std::vector<FrameAlignment>* details) {
  if (files.empty()) throw NotEnoughFiles();                                            // src/lib.rs:725
  if (devices.empty()) throw InvalidParams("no CUDA device given");
  const TermCriteria crit = term_criteria(params);
  ImageU8 first = decode(files[0]);
  check_colour(first);
  stk_ecc_config cfg{};
  cfg.width = first.width; cfg.height = first.height; cfg.channels = first.channels;
  if (scale_down_width) {
    // ecc_match_scaling_down (src/lib.rs:849-1028): validation (:876-888) and utils::scale_image's size rule
    int sw = 0, sh = 0;
    check(stk_scaled_size(first.width, first.height, *scale_down_width, &sw, &sh));
    cfg.ecc_width = sw; cfg.ecc_height = sh;
  }
  if (!crit.typ) throw OpenCvError("findTransformECC: criteria.type must have COUNT or EPS set");
  cfg.motion_type = (int)params.motion_type;
  cfg.criteria_type = crit.typ; cfg.max_count = crit.max_count; cfg.epsilon = crit.epsilon;
  cfg.gauss_filt_size = params.gauss_filt_size;
  cfg.lanes = 0; cfg.align = 1;
  const size_t ndev = devices.size();
  std::vector<std::unique_ptr<Ctx>> ctxs;
  for (size_t d = 0; d < ndev; ++d) {
    cfg.device = devices[d];
    cfg.seed_reference = d == 0 ? 1 : 0;
    ctxs.push_back(std::make_unique<Ctx>(cfg));
    check(stk_ecc_set_reference(ctxs[d]->c, first.data.data(), (size_t)first.width * first.channels));
  }
  if (ndev > 1) {
    std::vector<stk_ecc_ctx*> raw;
    for (auto& c : ctxs) raw.push_back(c->c);
    check(stk_ecc_peer_connect_local(raw.data(), (int)ndev));
  }
  // one decode task per frame (Rayon's into_par_iter, src/lib.rs:746-749); submission is thread-safe
  const size_t n = files.size();
  const unsigned workers = std::max(1u, std::min<unsigned>(8, std::thread::hardware_concurrency()));
  std::vector<std::future<void>> pool;
  std::atomic<size_t> next{1};
  for (unsigned w = 0; w < workers; ++w) {
    pool.push_back(std::async(std::launch::async, [&] {
      for (size_t i = next++; i < n; i = next++) {
        ImageU8 f = decode(files[i]);
        check_colour(f);
        if (f.width != first.width || f.height != first.height || f.channels != first.channels)
          throw OpenCvError("frame size differs from the first frame");
        check(stk_ecc_submit_frame(ctxs[i % ndev]->c, f.data.data(), (size_t)f.width * f.channels, (int64_t)i));
      }
    }));
  }
  for (auto& f : pool) f.get();                                                          // rethrows the first error
  ImageF32 out;
  out.width = first.width; out.height = first.height; out.channels = first.channels;
  out.data.resize((size_t)out.width * out.height * out.channels);
  if (ndev == 1) {
    check(stk_ecc_finish(ctxs[0]->c, (int)n, out.data.data(), (size_t)out.width * out.channels * sizeof(float)));
  } else {
    // queue the exchange on EVERY device first: a device's exchange kernel waits for all the others, and a copy
    // into pageable host memory blocks this thread until that device's exchange is done
    for (auto& c : ctxs) check(stk_ecc_peer_reduce_scatter(c->c, (int)n, nullptr, nullptr, nullptr));
    for (auto& c : ctxs) check(stk_ecc_peer_slice_to_host(c->c, out.data.data()));
    for (auto& c : ctxs) check(stk_ecc_sync(c->c));      // also surfaces per-frame ECC failures (src/lib.rs:777)
  }
  if (details) {
    details->clear();
    for (auto& c : ctxs) {
      std::vector<stk_frame_result> res(n);
      int count = 0;
      check(stk_ecc_results(c->c, res.data(), (int)n, &count));
      for (int i = 0; i < count; ++i) {
        FrameAlignment a;
        a.frame = res[i].tag;
        std::memcpy(a.warp, res[i].warp, sizeof a.warp);
        a.rho = res[i].rho; a.iterations = res[i].iterations;
        details->push_back(a);
      }
    }
    // submissions race on the decode pool: report in file order
    std::sort(details->begin(), details->end(), [](const FrameAlignment& x, const FrameAlignment& y) { return x.frame < y.frame; });
  }
  return out;
}

ImageF32 stack_with_homographies(const std::vector<ImageU8>& frames, const std::vector<std::array<double, 9>>& homographies,
                                 const KeyPointMatchParameters& params, int device) {
  if (frames.empty()) throw NotEnoughFiles();
  if (homographies.size() + 1 != frames.size()) throw InvalidParams("one homography per non-reference frame");
  const ImageU8& first = frames[0];
  check_colour(first);
  stk_ecc_config cfg{};
  cfg.width = first.width; cfg.height = first.height; cfg.channels = first.channels;
  cfg.device = device; cfg.seed_reference = 1; cfg.align = 0;
  Ctx ctx(cfg);
  check(stk_ecc_set_reference(ctx.c, first.data.data(), (size_t)first.width * first.channels));
  for (size_t i = 1; i < frames.size(); ++i) {
    check_colour(frames[i]);
    if (frames[i].width != first.width || frames[i].height != first.height || frames[i].channels != first.channels)
      throw NotImplemented("frames of differing size");         // the library would read first.height rows from this buffer
    check(stk_ecc_submit_warp(ctx.c, frames[i].data.data(), (size_t)frames[i].width * frames[i].channels,
                              homographies[i - 1].data(), params.border_mode, params.border_value, (int64_t)i));
  }
  ImageF32 out;
  out.width = first.width; out.height = first.height; out.channels = first.channels;
  out.data.resize((size_t)out.width * out.height * out.channels);
  check(stk_ecc_finish(ctx.c, (int)frames.size(), out.data.data(), (size_t)out.width * out.channels * sizeof(float)));
  return out;
}

double sharpness_tenengrad(const ImageU8& grey, int k_size, int device) {
  if (k_size != 1 && k_size != 3 && k_size != 5 && k_size != 7)
    throw InvalidParams("Kernel size must be 1, 3, 5, or 7");                            // src/lib.rs:1103-1107
  if (grey.channels != 1) throw OpenCvError("sharpness_tenengrad expects a single-channel image");
  double out = 0;
  check(stk_tenengrad(grey.data.data(), (size_t)grey.width, grey.width, grey.height, 1, k_size, device, &out));
  return out;
}

std::array<double, 4> sharpness_all(const ImageU8& grey, int device) {
  if (grey.channels != 1) throw OpenCvError("the sharpness metrics expect a single-channel image");
  std::array<double, 4> out{};
  check(stk_sharpness_all(grey.data.data(), (size_t)grey.width, grey.width, grey.height, 1, device, out.data()));
  return out;
}
double sharpness_modified_laplacian(const ImageU8& grey, int device) { return sharpness_all(grey, device)[0]; }
double sharpness_variance_of_laplacian(const ImageU8& grey, int device) { return sharpness_all(grey, device)[1]; }
double sharpness_normalized_gray_level_variance(const ImageU8& grey, int device) { return sharpness_all(grey, device)[3]; }

}  // namespace libstacker
