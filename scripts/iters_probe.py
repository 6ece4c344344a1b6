"""Per-frame ECC iteration counts of the benchmark stack (load-balance input for the multi-GPU sharding)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
import synthetic as synth
pkg = ge.load_package()
w, h, n = 3840, 2160, 64
st_ = synth.Stack(w, h, n, 3, seed=4)
params = pkg.EccMatchParameters(pkg.MotionType.Homography, 5000, 1e-5, 5)
with pkg.EccStack(w, h, 3, params, device=0, lanes=4) as st:
    st.set_reference(torch.from_numpy(st_.frame(0)).cuda())
    for i in range(1, n):
        st.submit(torch.from_numpy(st_.frame(i)).cuda(), tag=i)
    st.sync()
    res = sorted(st.results(), key=lambda r: r["tag"])
its = [r["iterations"] for r in res]
print("iterations per frame:", its)
for world in (2, 4, 8):
    loads = [sum(its[i - 1] for i in range(1, n) if (i - 1) % world == r) for r in range(world)]
    print(world, "ranks: iterations per rank", loads, "max/mean %.2f" % (max(loads) / (sum(loads) / world)))
