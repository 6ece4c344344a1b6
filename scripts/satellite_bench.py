"""Device times of the satellite kernels (Tenengrad, fused sharpness, INTER_AREA grey resize, scale-down ecc_match)
with CUDA events around the C-ABI calls on device-resident frames; GB/s on algorithmic bytes vs the measured HBM peak."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as ge
import synthetic as synth
pkg = ge.load_package()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6534.1


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3      # us


for (w, h, n) in [(3840, 2160, 32), (6000, 4000, 8)]:
    rng = np.random.default_rng(1)
    grey = torch.from_numpy(rng.integers(0, 256, (n, h, w), dtype=np.uint8)).cuda()
    bgr = torch.from_numpy(rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)).cuda()
    npx = w * h
    import ctypes as C
    lib = pkg._ffi.lib
    out = (C.c_double * (4 * n))()
    # the batch entry points include a cudaMalloc/memset/D2H of the sums per call: whole-call time, as a user sees it
    t = timed(lambda: lib.stk_tenengrad_batch_device(grey.data_ptr(), h * w, w, w, h, 1, 3, n, 0, out), reps=5)
    print(f"{w}x{h}: tenengrad k=3, grey, batch of {n}: {t / n:8.1f} us/frame  {npx / (t / n) / 1e3:7.0f} GB/s ({npx / (t / n) / 1e3 / peak:.0%} of HBM peak)")
    t = timed(lambda: lib.stk_sharpness_all_batch_device(grey.data_ptr(), h * w, w, w, h, 1, n, 0, out), reps=5)
    print(f"{w}x{h}: LAPM+LAPV+TENG+GLVN, grey, batch of {n}: {t / n:8.1f} us/frame  {npx / (t / n) / 1e3:7.0f} GB/s ({npx / (t / n) / 1e3 / peak:.0%})")
    t = timed(lambda: lib.stk_sharpness_all_batch_device(bgr.data_ptr(), h * w * 3, w * 3, w, h, 3, n, 0, out), reps=5)
    print(f"{w}x{h}: same from BGR (grey fused), batch of {n}: {t / n:8.1f} us/frame  {3 * npx / (t / n) / 1e3:7.0f} GB/s ({3 * npx / (t / n) / 1e3 / peak:.0%})")
    del grey, bgr

# scale-down ecc_match on the bench stack (16 frames of 4K, ECC at 960x540)
w, h, n = 3840, 2160, 16
st_ = synth.Stack(w, h, n, 3, seed=4)
frames = [torch.from_numpy(st_.frame(i)).cuda() for i in range(n)]
params = pkg.EccMatchParameters(pkg.MotionType.Homography, 5000, 1e-5, 5)
for sd in (None, 540.0):
    ecc_size = None if sd is None else pkg.scaled_size(w, h, sd)
    with pkg.EccStack(w, h, 3, params, device=0, lanes=4, ecc_size=ecc_size) as st:
        out_dev = torch.empty(h, w, 3, dtype=torch.float32, device="cuda")

        def step():
            st.reset()
            st.set_reference(frames[0])
            for i in range(1, n):
                st.submit(frames[i], tag=i)
            ptr, nfl = st.partial()
            st.finish_device(ptr, n, out_dev.data_ptr())
        t = timed(step, reps=5, warm=2)
        res = st.results()
        err = max(synth.corner_displacement(r["warp"], st_.truth[r["tag"]], w, h) for r in res)
        its = sum(r["iterations"] for r in res)
        print(f"ecc_match 16x4K Homography, scale_down={sd}: {t / 1e3:7.2f} ms per stack = {n / (t / 1e6):7.0f} frames/s, "
              f"{its} iterations, max corner error vs truth {err:.3f} px")
