"""The multi-GPU exchange step (stk_ecc_peer_*, csrc/peer_reduce.cuh) against the single-context result.

The fused reduce-scatter + divide must give exactly what the reference's try_reduce + `/ n`
(/root/reference/src/lib.rs:819-839) gives for the same partial sums: the kernel adds the ranks' partial stacks in
rank order and multiplies by float(1/n), which is what `finish` does on one device with the lanes in place of the
ranks.  World size 1 runs on any GPU box; the 2-rank cases need two devices."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _stack(pkg, n=5, w=320, h=240, motion=2, seed=31):
    from oracle import synth
    frames = synth.Stack(w, h, n, motion, seed=seed).frames()
    params = pkg.EccMatchParameters(pkg.MotionType(motion), 50, 1e-4, 5)
    return frames, params


def _device_array(ptr, shape):
    import torch
    from_ptr = type("P", (), {})()
    n = int(np.prod(shape))
    from_ptr.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                         "strides": None}
    return torch.as_tensor(from_ptr, device="cuda").view(*shape).cpu().numpy()


def test_world_of_one_equals_finish(pkg):
    frames, params = _stack(pkg)
    h, w = frames[0].shape[:2]
    with pkg.EccStack(w, h, 3, params, device=0, lanes=3) as st:
        st.set_reference(frames[0])
        for i, f in enumerate(frames[1:], 1):
            st.submit(f, tag=i)
        want = st.finish(len(frames))
        # same stack again through the exchange path, twice (step counter, reset in between)
        st.peer_connect(0, 1, [st.peer_export()])
        for _ in range(2):
            st.reset()
            st.set_reference(frames[0])
            for i, f in enumerate(frames[1:], 1):
                st.submit(f, tag=i)
            ptr = st.peer_reduce(len(frames))
            st.sync()
            got = _device_array(ptr, (h, w, 3))
            assert np.array_equal(got, want)
        # reduce-scatter form: the (only) rank keeps the whole stack as its slice and copies it to the host
        st.reset()
        st.set_reference(frames[0])
        for i, f in enumerate(frames[1:], 1):
            st.submit(f, tag=i)
        d_slice, begin, count = st.peer_reduce_scatter(len(frames))
        assert (begin, count) == (0, h * w * 3)
        host = np.zeros((h, w, 3), np.float32)
        st.peer_slice_to_host(host.ctypes.data)
        st.sync()
        assert np.array_equal(host, want)
        st.peer_disconnect()
        with pytest.raises(pkg.StackerError):
            st.peer_slice_to_host(host.ctypes.data)


def test_stacks_back_to_back_without_host_sync(pkg):
    """reset / set_reference / the exchange never synchronise with the host (ABI v5): three DIFFERENT stacks are queued
    on one context one after the other, each result is copied out of the exchange stream into its own host array, and
    only then does the host synchronise.  Every stack must equal what the same frames give on a fresh context — the
    next stack's reference prep, ECC iterations and accumulator writes are ordered behind the previous stack on the
    device (ref_ready / drained / x_done events)."""
    from oracle import synth
    w, h, n = 320, 240, 5
    params = pkg.EccMatchParameters(pkg.MotionType.Affine, 50, 1e-4, 5)
    stacks = [synth.Stack(w, h, n, 2, seed=s).frames() for s in (41, 42, 43)]
    want = []
    for frames in stacks:
        with pkg.EccStack(w, h, 3, params, device=0, lanes=3) as st:
            st.set_reference(frames[0])
            for i, f in enumerate(frames[1:], 1):
                st.submit(f, tag=i)
            want.append(st.finish(n))
    assert not np.array_equal(want[0], want[1])
    outs = [np.zeros((h, w, 3), np.float32) for _ in stacks]
    with pkg.EccStack(w, h, 3, params, device=0, lanes=3) as st:
        st.peer_connect(0, 1, [st.peer_export()])
        for frames, out in zip(stacks, outs):
            st.reset()
            st.set_reference(frames[0])
            for i, f in enumerate(frames[1:], 1):
                st.submit(f, tag=i)
            st.peer_reduce_scatter(n)
            st.peer_slice_to_host(out.ctypes.data)          # asynchronous, on the exchange stream
        st.sync()
        for got, ref in zip(outs, want):
            assert np.array_equal(got, ref)
        st.peer_disconnect()


def test_peer_reduce_needs_connect(pkg):
    frames, params = _stack(pkg, n=2)
    h, w = frames[0].shape[:2]
    with pkg.EccStack(w, h, 3, params, device=0) as st:
        st.set_reference(frames[0])
        with pytest.raises(pkg.StackerError):
            st.peer_reduce(2)
        with pytest.raises(pkg.StackerError):
            st.peer_connect(0, 2, [st.peer_export()])          # one handle for a world of two


def test_two_devices_one_process(pkg):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    D = pkg.distributed
    frames, params = _stack(pkg, n=6)
    n = len(frames)
    h, w = frames[0].shape[:2]
    with pkg.EccStack(w, h, 3, params, device=0, lanes=1) as one:
        one.set_reference(frames[0])
        for i in range(1, n):
            one.submit(frames[i], tag=i)
        one.sync()
        single = one.finish(n)
    stacks = [pkg.EccStack(w, h, 3, params, device=r, lanes=2, seed_reference=(r == 0)) for r in range(2)]
    try:
        pkg.EccStack.peer_connect_local(stacks)
        for rep in range(2):
            for r, st in enumerate(stacks):
                st.reset()
                st.set_reference(frames[0])
                for i in D.shard_frames(n, r, 2):
                    st.submit(frames[i], tag=i)
            ptrs = [st.peer_reduce(n) for st in stacks]
            for st in stacks:
                st.sync()
            assert ptrs[1] is None
            import torch
            with torch.cuda.device(0):
                got = _device_array(ptrs[0], (h, w, 3))
            # same frames, same warps; only the f32 summation order differs from the single-device stack
            assert np.abs(got - single).max() <= 1e-6
    finally:
        for st in stacks:
            st.close()


def test_two_processes_ipc_equals_nccl():
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "scripts", "peer_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "peer exchange ok" in r.stdout


def test_python_ecc_match_on_two_devices(pkg, tmp_path):
    """The Python mirror's ecc_match(devices=[...]): one process, one context per device, frames dealt round-robin,
    exchange + divide + per-device copy-out — against the single-device call, for decoded arrays and for files."""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    import cv2
    frames, params = _stack(pkg, n=6)
    one, r1 = pkg.ecc_match(frames, params, device=0, return_details=True)
    two, r2 = pkg.ecc_match(frames, params, devices=[0, 1], return_details=True)
    assert [r["tag"] for r in r2] == [r["tag"] for r in r1] == [1, 2, 3, 4, 5]
    for a, b in zip(r1, r2):
        assert np.array_equal(a["warp"], b["warp"]) and a["iterations"] == b["iterations"]
    assert np.abs(one - two).max() <= 1e-6           # f32 summation order only
    paths = []
    for i, f in enumerate(frames):
        paths.append(str(tmp_path / f"f{i}.png"))
        assert cv2.imwrite(paths[-1], f)
    three = pkg.ecc_match(paths, params, devices=[0, 1])
    assert np.abs(three - one).max() <= 1e-6
    with pytest.raises(pkg.StackerError):
        pkg.ecc_match(frames, params, devices=[0, 0])


def test_python_keypoint_match_on_two_devices(pkg):
    """keypoint_match(devices=[...]) (BASELINE configs[4]: the warp + stack tail across GPUs): same drops and the
    same stack as the single-device call, up to f32 summation order."""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    pytest.importorskip("cv2")
    from oracle import synth
    frames = synth.Stack(800, 600, 6, 3, seed=31).frames()
    params = pkg.KeyPointMatchParameters(method=pkg.RANSAC, ransac_reproj_threshold=5.0, match_keep_ratio=0.8, match_ratio=0.9)
    d1, one = pkg.keypoint_match(frames, params, device=0)
    d2, two = pkg.keypoint_match(frames, params, devices=[0, 1])
    assert d1 == d2 == 0
    assert np.abs(one - two).max() <= 1e-6
