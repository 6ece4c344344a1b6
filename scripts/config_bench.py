"""Device-resident throughput of every BASELINE.json config (frames already on the GPU as u8, contexts created),
CUDA events around whole stacks; the recovered warps are checked against the generator's ground truth.
Frame generation at 4K / 24 MPx is slow on the host, so configs 3-5 reuse a few generated frames."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as ge
import synthetic as synth
pkg = ge.load_package()


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def ecc_config(name, cfg, motion, n_total, n_gen, scale_down=None):
    st_ = synth.config_stack(cfg, n_frames=n_gen)
    w, h = st_.width, st_.height
    gen = [torch.from_numpy(st_.frame(i)).cuda() for i in range(n_gen)]
    frames = [gen[0]] + [gen[1 + (i % (n_gen - 1))] for i in range(n_total - 1)]
    params = pkg.EccMatchParameters(pkg.MotionType(motion), 5000, 1e-5, 5)
    ecc_size = None if scale_down is None else pkg.scaled_size(w, h, scale_down)
    out = torch.empty(h, w, 3, dtype=torch.float32, device="cuda")
    with pkg.EccStack(w, h, 3, params, device=0, lanes=4, ecc_size=ecc_size) as st:
        def step():
            st.reset()
            st.set_reference(frames[0])
            for i in range(1, n_total):
                st.submit(frames[i], tag=1 + ((i - 1) % (n_gen - 1)))
            ptr, _ = st.partial()
            st.finish_device(ptr, n_total, out.data_ptr())
        ms = timed(step)
        res = st.results()
    err = max(synth.corner_displacement(r["warp"] if motion == 3 else r["warp"][:2], st_.truth[r["tag"]], w, h) for r in res)
    its = sum(r["iterations"] for r in res)
    print(f"{name}: {n_total} x {w}x{h}: {ms:8.2f} ms per stack = {n_total / ms * 1e3:7.0f} frames/s, {its} ECC iterations, "
          f"max corner error vs truth {err:.3f} px", flush=True)
    return frames, st_


ecc_config("config 1  ecc_match Homography (examples/main.rs)", 1, 3, 5, 5)
ecc_config("config 1' ecc_match Homography, scale_down 400", 1, 3, 5, 5, 400.0)
ecc_config("config 2  ecc_match Euclidean", 2, 1, 16, 6)
frames3, st3 = ecc_config("config 3  ecc_match Affine (31 after drop-worst)", 3, 2, 31, 5)
# config 3's ranking step: Tenengrad(3) of 32 grey 4K planes, batch call
greys = torch.stack([f.float().mean(dim=2).to(torch.uint8) for f in frames3[:4]] * 8)
import ctypes as C
out = (C.c_double * 32)()
ms = timed(lambda: pkg._ffi.lib.stk_tenengrad_batch_device(greys.data_ptr(), 2160 * 3840, 3840, 3840, 2160, 1, 3, 32, 0, out))
print(f"config 3  sharpness_tenengrad(3) of 32 x 3840x2160 grey planes: {ms:8.2f} ms = {32 / ms * 1e3:7.0f} frames/s", flush=True)
del greys, frames3
ecc_config("config 4  ecc_match Homography (the bench workload, 16 of 64 frames)", 4, 3, 16, 5)
# config 5: keypoint_match tail at 6000x4000 — warp + accumulate with given homographies (host ORB stages excluded)
w, h, n = 6000, 4000, 16
rng = np.random.default_rng(5)
base = [torch.from_numpy(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).cuda() for _ in range(3)]
hs = [synth.random_warp(rng, 3, w, h) for _ in range(n - 1)]
out5 = torch.empty(h, w, 3, dtype=torch.float32, device="cuda")
with pkg.EccStack(w, h, 3, None, device=0, lanes=4) as st:
    def step5():
        st.reset()
        st.set_reference(base[0])
        for i in range(1, n):
            st.submit_warp(base[i % 3], hs[i - 1], tag=i)
        ptr, _ = st.partial()
        st.finish_device(ptr, n, out5.data_ptr())
    ms = timed(step5)
print(f"config 5  keypoint_match tail (warp_perspective + accumulate + / n): {n} x {w}x{h}: {ms:8.2f} ms per stack = "
      f"{n / ms * 1e3:7.0f} frames/s ({27 * w * h * (n - 1) / ms / 1e6:6.0f} GB/s on 27N bytes per frame)", flush=True)
