"""Small end-to-end exercise of every kernel, for compute-sanitizer runs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
import synthetic as synth
pkg = ge.load_package()
w, h = 200, 136
for motion in (0, 1, 2, 3):
    frames = synth.Stack(w, h, 4, motion, seed=47 + motion).frames()
    out, res = pkg.ecc_match(frames, pkg.EccMatchParameters(pkg.MotionType(motion), 30, 1e-4, 5), None, device=0, return_details=True)
    print(motion, [r["iterations"] for r in res], float(out.mean()))
rng = np.random.default_rng(3)
fr4 = [rng.integers(0, 256, (90, 120, 4), dtype=np.uint8) for _ in range(2)]
hm = synth.random_warp(rng, 3, 120, 90); hm[:2, 2] += (30, -20)
with pkg.EccStack(120, 90, 4, None, device=0, lanes=1) as st:
    st.set_reference(fr4[0]); st.submit_warp(fr4[1], hm, 0, (0.25, 0.5, 0.75, 1.0)); print(float(st.finish(2).mean()))
print(pkg.sharpness_tenengrad(frames[0][..., 0].copy(), 3, device=0))
print(float(pkg.prep_grey_blur(frames[0], 5, device=0).mean()))
# round-1 additions: scale-down path (K0 generic + 2x2 paths), fused sharpness, larger interior frames (lean/packed paths)
fr = synth.Stack(480, 360, 3, 3, seed=65).frames()
for sd in (240.0, 180.0, 200.0):
    out, res = pkg.ecc_match(fr, pkg.EccMatchParameters(pkg.MotionType.Homography, 20, 1e-4, 5), sd, device=0, return_details=True)
    print("scale", sd, [r["iterations"] for r in res], float(out.mean()))
out, res = pkg.ecc_match(fr, pkg.EccMatchParameters(pkg.MotionType.Affine, 20, 1e-4, 3), 111.0, device=0, return_details=True)
print("affine scale", [r["iterations"] for r in res])
print(pkg.sharpness_all(fr[0][..., 1].copy(), device=0))
print(pkg.grey_resize_area(fr[0], 213, 160, device=0).mean(), pkg.grey_resize_area(fr[0][..., 0].copy(), 160, 120, device=0).mean())
big = synth.Stack(700, 520, 3, 3, seed=9).frames()
out, res = pkg.ecc_match(big, pkg.EccMatchParameters(pkg.MotionType.Homography, 25, 1e-5, 5), None, device=0, return_details=True)
print("big", [r["iterations"] for r in res], float(out.mean()))
