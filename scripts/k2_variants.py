"""K2 geometry / generation sweep on one B200: for every compiled variant of the ECC iteration kernel
(STK_ECC_GEN x STK_ECC_CFG, csrc/ecc_iter_v2.cuh) measure, on the same 4K Homography frames,

  one lane  : us per iteration = CUDA events on the lane stream around each frame's device loop / iterations
              (serial tail and relaunch gaps included) — what `roofline.us_per_launch` of bench.py reports;
  four lanes: device time of the whole stack / iterations after subtracting nothing (prep and warp included)
              and the stack's frames/s;

and check that every variant recovers the same warps (max corner displacement against variant gen1) with the
same iteration counts.  Usage:  python scripts/k2_variants.py [n_frames] [variant ...]   (variant = gen:cfg[:blocks per SM per launch])
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__ as ge
import synthetic as synth

pkg = ge.load_package()
w, h = 3840, 2160
n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
variants = sys.argv[2:] or ["1:0", "2:0", "2:1", "2:2", "2:3", "2:4", "2:5", "2:6", "2:7"]
dev = torch.device("cuda", 0)
src = synth.Stack(w, h, n, 3, seed=4)
frames = [torch.from_numpy(src.frame(i)).to(dev) for i in range(n)]
params = pkg.EccMatchParameters(pkg.MotionType.Homography, 5000, 1e-5, 5)
base = None
rows = []
for v in variants:
    parts = v.split(":")
    gen, cfg = parts[0], parts[1]
    os.environ["STK_ECC_GEN"], os.environ["STK_ECC_CFG"] = gen, cfg
    if len(parts) > 2:
        os.environ["STK_ECC_BLOCKS_PER_SM"] = parts[2]       # gen:cfg:blocks-per-SM-per-launch
    else:
        os.environ.pop("STK_ECC_BLOCKS_PER_SM", None)
    rec = {"variant": v}
    try:
        with pkg.EccStack(w, h, 3, params, device=0, lanes=1) as s1:
            for rep in range(2):
                s1.reset()
                s1.set_profiling(rep == 1)
                s1.set_reference(frames[0])
                for i in range(1, n):
                    s1.submit(frames[i], tag=i)
                s1.sync()
            t = s1.stage_times()
            res = s1.results()
        rec["one_lane_us_per_iter"] = 1e3 * t["loop_ms"] / t["iterations"]
        rec["iterations"] = t["iterations"]
        warps = {r["tag"]: r["warp"] for r in res}
        if base is None:
            base = warps
        rec["max_px_vs_first_variant"] = max(synth.corner_displacement(warps[k], base[k], w, h) for k in warps)
        with pkg.EccStack(w, h, 3, params, device=0, lanes=4) as s4:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            for rep in range(1 + reps):
                if rep == 1:
                    torch.cuda.synchronize()
                    e0.record()
                s4.reset()
                s4.set_reference(frames[0])
                for i in range(1, n):
                    s4.submit(frames[i], tag=i)
                s4.sync()
            torch.cuda.synchronize()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            it4 = sum(r["iterations"] for r in s4.results())
        rec["four_lane_ms_per_stack"] = ms
        rec["four_lane_us_per_iter_all_in"] = 1e3 * ms / it4
        rec["four_lane_frames_per_s"] = n / (ms * 1e-3)
    except Exception as e:  # a variant that does not fit the device (shared memory) is reported, not fatal
        rec["error"] = str(e)
    rows.append(rec)
    print(json.dumps(rec), flush=True)
