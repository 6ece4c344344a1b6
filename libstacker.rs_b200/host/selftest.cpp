// Self-test of the C++ host mirror.  `selftest cpu` exercises what needs no GPU (parameter semantics and
// error behaviour); `selftest gpu <dir>` stacks PNM frames written by tests/test_cpp_host.py and prints the
// recovered translations, the Tenengrad value and a checksum of the stack for the Python side to compare.
#include <array>
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <iostream>

#include "libstacker.hpp"
using namespace libstacker;

#define EXPECT(cond) do { if (!(cond)) { std::printf("FAILED: %s (line %d)\n", #cond, __LINE__); return 1; } } while (0)
template <class E, class F> bool throws(F&& f) { try { f(); } catch (const E&) { return true; } catch (...) { return false; } return false; }

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "cpu";
  // the reference's doctest: max_count None, epsilon Some(0.1) -> epsilon 0.1, typ == EPS (src/utils.rs:148-158)
  TermCriteria t = term_criteria({MotionType::Euclidean, std::nullopt, 0.1, 3});
  EXPECT(t.epsilon == 0.1 && t.typ == TERM_EPS && t.max_count == 0);
  t = term_criteria({MotionType::Affine, 5000, 1e-5, 5});
  EXPECT(t.typ == (TERM_COUNT | TERM_EPS) && t.max_count == 5000);
  EXPECT((int)MotionType::Homography == 3 && (int)MotionType::Translation == 0);
  KeyPointMatchParameters kd;
  EXPECT(kd.method == 8 && kd.ransac_reproj_threshold == 3.0 && kd.match_keep_ratio == 0.75f && kd.match_ratio == 0.8f);
  EXPECT(throws<NotEnoughFiles>([] { ecc_match({}, {MotionType::Affine, 10, 1e-3, 5}); }));
  EXPECT(throws<OpenCvError>([] { ecc_match({"/nonexistent/frame.ppm"}, {MotionType::Affine, 10, 1e-3, 5}); }));
  ImageU8 g; g.width = 8; g.height = 8; g.channels = 1; g.data.assign(64, 7);
  EXPECT(throws<InvalidParams>([&] { sharpness_tenengrad(g, 4); }));
  if (mode == "cpu") { std::printf("cpu selftest ok\n"); return 0; }

  EXPECT(argc > 2);
  const std::filesystem::path dir = argv[2];
  std::vector<std::filesystem::path> files;
  for (int i = 0;; ++i) {
    auto p = dir / ("f" + std::to_string(i) + ".ppm");
    if (!std::filesystem::exists(p)) break;
    files.push_back(p);
  }
  EXPECT(files.size() >= 2);
  std::vector<FrameAlignment> det;
  ImageF32 out = ecc_match(files, {MotionType::Translation, 200, 1e-6, 5}, std::nullopt, read_pnm, 0, &det);
  std::printf("frames %zu size %dx%dx%d\n", files.size(), out.width, out.height, out.channels);
  for (auto& a : det) std::printf("warp %.6f %.6f iters %d rho %.6f\n", a.warp[2], a.warp[5], a.iterations, a.rho);
  double sum = 0;
  for (float v : out.data) sum += v;
  std::printf("stack_sum %.6f\n", sum);
  // the same stack over every GPU of the box from this one process (skipped on a single-GPU box)
  const std::vector<int> devs = all_devices();
  if (devs.size() >= 2) {
    std::vector<FrameAlignment> det2;
    ImageF32 out2 = ecc_match_on_devices(files, {MotionType::Translation, 200, 1e-6, 5}, devs, std::nullopt, read_pnm, &det2);
    EXPECT(det2.size() == det.size());
    for (size_t i = 0; i < det.size(); ++i) EXPECT(std::memcmp(det[i].warp, det2[i].warp, sizeof det[i].warp) == 0);
    double worst = 0;
    for (size_t i = 0; i < out.data.size(); ++i) worst = std::max(worst, (double)std::fabs(out.data[i] - out2.data[i]));
    EXPECT(worst <= 1e-6);                                  // f32 summation order only
    std::printf("multi_gpu devices %zu max_abs_diff %.3g\n", devs.size(), worst);
  }
  ImageU8 grey = read_pnm(dir / "grey.pgm");
  std::printf("tenengrad %.17g\n", sharpness_tenengrad(grey, 3, 0));
  // neither COUNT nor EPS: CV_Assert inside findTransformECC -> OpenCvError
  EXPECT(throws<OpenCvError>([&] { ecc_match(files, {MotionType::Translation, std::nullopt, std::nullopt, 5}, std::nullopt, read_pnm, 0); }));
  std::printf("gpu selftest ok\n");
  return 0;
}
