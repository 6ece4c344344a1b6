"""Where does one ECC iteration spend its time?  %globaltimer stamps from inside the kernel, per block:
start | first chunk landed (TMA) | pixel loop + folds done | partial written, and the last block's tail.
Usage: python scripts/timing_probe.py [gen:cfg ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
import synthetic as synth
pkg = ge.load_package()
w, h = 3840, 2160
st_ = synth.Stack(w, h, 2, 3, seed=4)
f0, f1 = st_.frame(0), st_.frame(1)
params = pkg.EccMatchParameters(pkg.MotionType.Homography, 50, 1e-5, 5)
for v in (sys.argv[1:] or ["1:0", "2:0", "2:2"]):
    os.environ["STK_ECC_GEN"], os.environ["STK_ECC_CFG"] = v.split(":")
    with pkg.EccStack(w, h, 3, params, device=0, lanes=1) as st:
        st.set_reference(f0)
        tiles, tail = st.debug_timing(f1, st_.truth[1].astype(np.float32), iters=4)
        t0 = tiles[:, 0].min()
        rel = (tiles.astype(np.int64) - int(t0)) / 1e3
        q = lambda a: "min %.1f med %.1f max %.1f" % (a.min(), np.median(a), a.max())
        print(f"== variant {v}: {len(tiles)} blocks")
        print("  block start        us:", q(rel[:, 0]))
        if v[0] == "2":
            print("  first chunk landed us after start:", q(rel[:, 3] - rel[:, 0]))
        print("  pixels+folds done  us:", q(rel[:, 1]))
        print("  block main loop duration us:", q(rel[:, 1] - rel[:, 0]))
        dur = rel[:, 1] - rel[:, 0]
        hist, edges = np.histogram(dur, bins=np.arange(16, 64, 4))
        print("  duration histogram (us: blocks):", " ".join(f"{int(e)}-{int(e) + 4}:{c}" for e, c in zip(edges, hist) if c))
        print("  durations by block id (us):", " ".join(str(int(round(float(d)))) for d in dur))
        tl = (tail[:3].astype(np.int64) - int(t0)) / 1e3
        print("  tail: cross-block sum done %.1f, solve done %.1f, end %.1f  (last block %d)" % (tl[0], tl[1], tl[2], int(tail[3])))
