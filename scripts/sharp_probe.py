import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, ctypes as C
import __graft_entry__ as ge
pkg = ge.load_package(); lib = pkg._ffi.lib
for (w, h, n) in [(3840, 2160, 32), (3840, 2160, 8), (3840, 2160, 1), (6000, 4000, 8), (6016, 4000, 8), (6000, 4000, 1)]:
    grey = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device="cuda")
    out = (C.c_double * (4 * n))()
    for name, fn in (("teng", lambda: lib.stk_tenengrad_batch_device(grey.data_ptr(), h * w, w, w, h, 1, 3, n, 0, out)),
                     ("all", lambda: lib.stk_sharpness_all_batch_device(grey.data_ptr(), h * w, w, w, h, 1, n, 0, out))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): fn()
        torch.cuda.synchronize()
        print(w, h, n, name, "%.1f us/frame" % ((time.perf_counter() - t0) / 3 / n * 1e6))
    del grey
