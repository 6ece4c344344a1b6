"""Where does one ECC iteration spend its time?  %globaltimer stamps from inside the kernel."""
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
with pkg.EccStack(w, h, 3, params, device=0, lanes=1) as st:
    st.set_reference(f0)
    for it in (1, 4):
        tiles, tail = st.debug_timing(f1, st_.truth[1].astype(np.float32), iters=it)
        t0 = tiles[:, 0].min()
        rel = (tiles[:, :3].astype(np.int64) - int(t0)) / 1e3
        print(f"iters={it} tiles={len(tiles)}")
        print("  block start  us: min %.1f med %.1f max %.1f" % (rel[:, 0].min(), np.median(rel[:, 0]), rel[:, 0].max()))
        print("  pixels done  us: min %.1f med %.1f max %.1f" % (rel[:, 1].min(), np.median(rel[:, 1]), rel[:, 1].max()))
        print("  partial done us: min %.1f med %.1f max %.1f" % (rel[:, 2].min(), np.median(rel[:, 2]), rel[:, 2].max()))
        print("  main-loop duration per block us: min %.1f med %.1f max %.1f" % ((rel[:, 1] - rel[:, 0]).min(), np.median(rel[:, 1] - rel[:, 0]), (rel[:, 1] - rel[:, 0]).max()))
        tl = (tail[:3].astype(np.int64) - int(t0)) / 1e3
        print("  tail: cross-tile sum done %.1f, solve done %.1f, end %.1f  (last block %d)" % (tl[0], tl[1], tl[2], int(tail[3])))
        order = np.argsort(rel[:, 1])
        print("  slowest tiles:", [(int(i), round(float(rel[i, 1]), 1)) for i in order[-6:]])
        if it == 4:
            dur = rel[:, 1] - rel[:, 0]
            print("  durations by block:", " ".join(str(int(round(float(d)))) for d in dur))
