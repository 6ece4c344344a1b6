"""BASELINE configs[4] at scale: the device stage of keypoint_match (warp_perspective + accumulate + / n,
/root/reference/src/lib.rs:289-346) on 256 frames of 6000x4000 over 1/2/4/8 B200 — 18.4 GB of u8 frames resident
in HBM, sharded over the ranks, one fused exchange + divide over NVLink.

  python scripts/config5_scale.py                                   # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         scripts/config5_scale.py [--frames 256 --width 6000 --height 4000 --steps 3]

The host stages of keypoint_match (ORB, BFMatcher, findHomography: ~0.6 s per 24 MPx frame on the CPU, north-star keeps
them on the host) are NOT part of this measurement: the homographies are seeded random near-identity matrices of the
kind findHomography returns.  Frames are seeded random bytes generated on the device (no host rendering of 256 x 72 MB).
Prints one JSON line on rank 0: frames/s, GB/s per GPU on the 27N-bytes-per-frame accounting of SURVEY §8(d) and on the
bytes the batched kernel really moves (3N + 24N/k), and — unless --no-check — the comparison of the N-GPU stack with
the same stack redone on rank 0's single GPU."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import __graft_entry__ as ge
import synthetic as synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--width", type=int, default=6000)
    ap.add_argument("--height", type=int, default=4000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--lanes", type=int, default=4)
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = ge.load_package()
    D = pkg.distributed
    n, w, h = a.frames, a.width, a.height

    def frame(i):           # seeded per index: every rank (and the single-GPU check) sees the same bytes for frame i
        g = torch.Generator(device=dev)
        g.manual_seed(1000 + i)
        return torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device=dev, generator=g)

    rng = np.random.default_rng(5)
    hs = {i: synth.random_warp(rng, 3, w, h) for i in range(1, n)}
    mine = D.shard_frames(n, rank, world)
    frames = {0: frame(0)}
    for i in mine:
        frames[i] = frame(i)
    st = pkg.EccStack(w, h, 3, None, device=local, lanes=a.lanes, seed_reference=(rank == 0))
    use_peers = world > 1 and D.connect_peers(st)
    out = torch.empty(h, w, 3, dtype=torch.float32, device=dev)
    res = {}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        st.reset()
        st.set_reference(frames[0])
        for i in mine:
            st.submit_warp(frames[i], hs[i], tag=i)
        if use_peers:
            res["ptr"] = st.peer_reduce(n)
            return
        ptr, nfl = st.partial()
        if world > 1:
            part = torch.as_tensor(D.DevicePtrArray(ptr, nfl), device=dev)
            D.reduce_partial_stack(part, 0)
            torch.cuda.synchronize()
        if rank == 0:
            st.finish_device(ptr, n, out.data_ptr())

    for _ in range(a.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    if use_peers:
        st.sync()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())

    check = None
    if rank == 0:
        multi = (torch.as_tensor(D.DevicePtrArray(res["ptr"], h * w * 3), device=dev).view(h, w, 3) if use_peers else out).clone()
    if world > 1 and not a.no_check and rank == 0:
        with pkg.EccStack(w, h, 3, None, device=local, lanes=a.lanes, seed_reference=True) as s1:
            s1.set_reference(frames[0])
            for i in range(1, n):
                s1.submit_warp(frames[i] if i in frames else frame(i), hs[i], tag=i)
                if i not in frames and i % 8 == 0:
                    s1.sync()           # bounds the temporaries alive on the device
            ptr, _ = s1.partial()
            single = torch.empty_like(multi)
            s1.finish_device(ptr, n, single.data_ptr())
        check = {"max_abs_diff_8bit_vs_single_gpu": float((torch.round(multi * 255) - torch.round(single * 255)).abs().max().item()),
                 "max_abs_diff_f32_vs_single_gpu": float((multi - single).abs().max().item())}
    if use_peers:
        barrier()
        st.peer_disconnect()
    st.close()
    if rank == 0:
        npx = w * h
        k = 4
        line = {"workload": f"BASELINE configs[4] device stage: keypoint_match tail, {n} x {w}x{h} u8 BGR, warp_perspective + accumulate + / n",
                "n_gpus": world, "ms_per_stack": ms, "frames_per_s": n / (ms * 1e-3),
                "GBps_per_gpu_on_27N": 27.0 * npx * (n - 1) / world / (ms * 1e-3) / 1e9,
                "GBps_per_gpu_moved": (3.0 + 24.0 / k) * npx * (n - 1) / world / (ms * 1e-3) / 1e9,
                "resident_u8_GB": 3.0 * npx * n / 1e9, "exchange": "fused peer reduce + divide" if use_peers else ("NCCL reduce" if world > 1 else "none"),
                "stack_mean": float(multi.mean().item()), "check": check}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
