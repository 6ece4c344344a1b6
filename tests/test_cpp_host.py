"""The C++ host mirror (libstacker.rs_b200/host): parameter/error semantics on CPU, a small stack on GPU."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "libstacker.rs_b200", "host")


@pytest.fixture(scope="module")
def selftest(pkg):
    subprocess.run(["make", "-C", HOST, "all"], check=True, capture_output=True)
    return os.path.join(HOST, "selftest")


def test_cpp_host_cpu_semantics(selftest):
    out = subprocess.run([selftest, "cpu"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cpu selftest ok" in out.stdout


def _write_pnm(path, arr):
    if arr.ndim == 3:
        h, w, _ = arr.shape
        with open(path, "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (w, h))
            f.write(np.ascontiguousarray(arr[..., ::-1]).tobytes())      # file stores RGB
    else:
        h, w = arr.shape
        with open(path, "wb") as f:
            f.write(b"P5\n%d %d\n255\n" % (w, h))
            f.write(np.ascontiguousarray(arr).tobytes())


@pytest.mark.gpu
def test_cpp_host_gpu_stack_matches_oracle(selftest, tmp_path):
    from oracle import restate as R, synth
    w, h = 256, 160
    frames = synth.Stack(w, h, 4, 0, seed=61).frames()
    for i, f in enumerate(frames):
        _write_pnm(tmp_path / f"f{i}.ppm", f)
    grey = R.bgr2gray_u8(frames[0])
    _write_pnm(tmp_path / "grey.pgm", grey)
    out = subprocess.run([selftest, "gpu", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.splitlines()
    warps = [tuple(float(v) for v in ln.split()[1:3]) for ln in lines if ln.startswith("warp")]
    want_stack, want_warps, _ = R.ecc_match(frames, 0, 200, 1e-6, 5)
    assert len(warps) == 3
    for (tx, ty), m in zip(warps, want_warps[1:]):
        assert abs(tx - m[0, 2]) < 0.05 and abs(ty - m[1, 2]) < 0.05, (tx, ty, m)
    stack_sum = float(next(ln for ln in lines if ln.startswith("stack_sum")).split()[1])
    want_sum = float(want_stack.astype(np.float64).sum())
    assert abs(stack_sum - want_sum) < 2e-4 * want_sum, (stack_sum, want_sum)
    teng = float(next(ln for ln in lines if ln.startswith("tenengrad")).split()[1])
    assert teng == R.sharpness_tenengrad(grey, 3)
    # on a box with several GPUs the self-test also shards the stack over all of them from its one process
    # (ecc_match_on_devices: peer_connect_local + reduce_scatter + slice_to_host) and checks it against the
    # single-device stack itself
    import torch
    if torch.cuda.device_count() >= 2:
        assert any(ln.startswith("multi_gpu devices") for ln in lines), out.stdout
