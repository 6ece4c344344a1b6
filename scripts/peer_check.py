"""torchrun --nproc-per-node N scripts/peer_check.py — one process per GPU: the fused peer exchange
(stk_ecc_peer_reduce over CUDA-IPC mappings) against the NCCL reduce + scale kernel it replaces, on the same
partial stacks; prints the time of both."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    import synthetic as synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = ge.load_package()
    D = pkg.distributed
    w, h, n, motion = int(os.environ.get("PC_W", 640)), int(os.environ.get("PC_H", 480)), 2 * world + 1, 2
    frames = synth.Stack(w, h, n, motion, seed=33).frames()
    params = pkg.EccMatchParameters(pkg.MotionType(motion), 50, 1e-4, 5)
    st = pkg.EccStack(w, h, 3, params, device=local, lanes=2, seed_reference=(rank == 0))

    def fill():
        st.reset()
        st.set_reference(frames[0])
        for i in D.shard_frames(n, rank, world):
            st.submit(frames[i], tag=i)

    # baseline: NCCL reduce + scale on the root
    fill()
    ptr, nfl = st.partial()
    part = torch.as_tensor(D.DevicePtrArray(ptr, nfl), device=dev)
    D.reduce_partial_stack(part, 0)
    torch.cuda.synchronize()
    want = None
    if rank == 0:
        out = torch.empty(h, w, 3, dtype=torch.float32, device=dev)
        st.finish_device(ptr, n, out.data_ptr())
        want = out.cpu().numpy()
    # fused exchange
    assert D.connect_peers(st), D.connect_peers.last_failure
    for rep in range(3):
        fill()
        d_out = st.peer_reduce(n)
        st.sync()
        if rank == 0:
            got = torch.as_tensor(D.DevicePtrArray(d_out, h * w * 3), device=dev).view(h, w, 3).cpu().numpy()
            # NCCL's reduce adds in its own (ring/tree) order: equal up to f32 summation order
            err = float(np.abs(got - want).max())
            assert err <= 1e-6, err
        else:
            assert d_out is None
    # reduce-scatter + every rank copying its slice into the host stack all ranks map
    shared = D.SharedHostStack((h, w, 3))
    for rep in range(2):
        fill()
        d_slice, begin, count = st.peer_reduce_scatter(n)
        assert (begin, begin + count) == tuple(D.scatter_bounds(h * w * 3, rank, world))
        st.peer_slice_to_host(shared.ptr)
        st.sync()
        dist.barrier()
        if rank == 0:
            err = float(np.abs(shared.array - want).max())
            assert err <= 1e-6, err
        dist.barrier()
    shared.close()
    # timing of the exchange alone (partials in place)
    for name in ("peer", "nccl"):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            if name == "peer":
                st.peer_reduce(n)
            else:
                D.reduce_partial_stack(part, 0)
                if rank == 0:
                    st.finish_device(ptr, n, out.data_ptr())
        st.sync()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            print(f"{name}: {1e6 * (time.perf_counter() - t0) / reps:.1f} us per exchange of {h * w * 12 / 1e6:.1f} MB, world {world}")
    dist.barrier()
    st.peer_disconnect()
    dist.barrier()
    st.close()
    if rank == 0:
        print("peer exchange ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
