"""Multi-GPU plumbing for one stack: one process per GPU (torch.distributed, NCCL over NVLink), frames
sharded across ranks, ONE exchange step of the partial stacks to rank 0 fused with the divide.

The exchange is the library's own reduce-scatter kernel over NVLink peer memory (csrc/peer_reduce.cuh,
`connect_peers` + `EccStack.peer_reduce`): torch.distributed only carries the 256-byte handles once.  The
NCCL form (`reduce_partial_stack` + `finish_device`) stays as the baseline it is measured against.

This is the B200 form of the reference's Rayon map-reduce (/root/reference/src/lib.rs:746-751, :819-839):
`try_fold` = each rank aligning + accumulating its own frames, `try_reduce` = the reduce, `/ n` = the final
scale on the root.  The frame -> rank assignment and the reduce are host logic and are exercised on CPU
with the gloo backend (tests/test_distributed_gloo.py)."""
from __future__ import annotations

from typing import List


def shard_frames(n_frames: int, rank: int, world: int) -> List[int]:
    """Indices (into the stack, frame 0 = reference excluded) this rank aligns: interleaved, so that every
    rank gets the same number of frames +-1 regardless of how the caller ordered them.  Frame i goes to rank
    i % world, so when the count does not divide evenly the short shard is rank 0's — the rank that also
    seeds the stack with frame 0 and does the final scale."""
    return [i for i in range(1, n_frames) if i % world == rank]


class DevicePtrArray:
    """Wraps a raw device pointer as a __cuda_array_interface__ provider so torch can view it without a
    copy (torch.as_tensor(DevicePtrArray(...), device=...))."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {
            "shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None}


def reduce_partial_stack(tensor, dst: int = 0, group=None):
    """In-place sum-reduce of each rank's partial stack onto `dst` (ncclReduce under the nccl backend)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return tensor


def slice_bounds(n_floats: int, rank: int, world: int):
    """[begin, end) of the stack (in floats) that `rank` reduces in the peer exchange — the rule
    stk_ecc_peer_reduce applies: slices in units of 4 floats, the last worker also takes the remainder; with
    more than two ranks the root (rank 0) takes no slice, because every finished pixel already has to enter
    the root over its inbound NVLink side."""
    workers = world - 1 if world > 2 else world
    w = rank - 1 if world > 2 else rank
    if w < 0:
        return 0, 0
    per = (n_floats // 4) // workers
    begin = w * per * 4
    end = n_floats if w == workers - 1 else (w + 1) * per * 4
    return begin, end


def bind_to_gpu_numa_node(device_index: int) -> bool:
    """Pin the calling process to the CPUs next to `device_index` (NVML's ideal CPU affinity for the GPU) so that
    the pinned frame buffers it allocates afterwards are first-touched on the GPU's own NUMA node.  With one
    process per GPU and every process on the default node, the uploads of all GPUs pull from ONE socket's memory
    (measured on an 8-GPU box: 155 GB/s aggregate, 19 GB/s per GPU instead of 50).  Returns True only when the
    affinity actually CHANGED (False when NVML is not available or reports no narrower CPU set); never raises."""
    try:
        import os
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = torch.cuda.get_device_properties(device_index).uuid
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) if (cpus & allowed) else set()
        if not cpus or cpus == allowed:
            return False          # nothing to narrow (e.g. a box that reports every GPU next to every CPU)
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def scatter_bounds(n_floats: int, rank: int, world: int):
    """[begin, end) of the slice `rank` keeps after stk_ecc_peer_reduce_scatter: equal slices in units of 4 floats
    for every rank (nothing converges on the root), the last rank also takes the remainder."""
    per = (n_floats // 4) // world
    begin = rank * per * 4
    end = n_floats if rank == world - 1 else (rank + 1) * per * 4
    return begin, end


def gather_handles(handle: bytes, group=None):
    """All ranks' peer handles, rank order (plain bytes through all_gather_object; works under gloo and nccl)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, bytes(handle), group=group)
    return out


def connect_peers(stack, group=None) -> bool:
    """Map every rank's partial stack into this rank's context (CUDA IPC).  Collective.  Returns False —
    on EVERY rank — if any rank could not map its peers (no NVLink/P2P between the devices, IPC forbidden by
    the container), so that all ranks take the same path afterwards."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    handles = gather_handles(stack.peer_export(), group)
    ok, why = True, ""
    try:
        stack.peer_connect(rank, world, handles)
    except Exception as e:  # the outcome is agreed on below; nothing is swallowed silently
        ok, why = False, str(e)
    votes = [None] * world
    dist.all_gather_object(votes, (ok, why), group=group)
    if all(v[0] for v in votes):
        return True
    if ok:
        stack.peer_disconnect()
    connect_peers.last_failure = "; ".join(f"rank {r}: {v[1]}" for r, v in enumerate(votes) if not v[0])
    return False


connect_peers.last_failure = ""


class SharedHostStack:
    """The finished stack in HOST memory that every rank of the node maps (POSIX shared memory, page-locked in
    each process): after `EccStack.peer_reduce_scatter` every rank copies its own slice out over its own PCIe
    link (`EccStack.peer_slice_to_host(shared.ptr)`), so the device-to-host copy of the result is spread over
    all GPUs instead of serialised on the root's link.  Collective constructor (raises RuntimeError on EVERY rank
    when /dev/shm has no room); `array` is the H x W x C f32 view (read it on any rank after every rank's sync()
    + a barrier)."""

    def __init__(self, shape, group=None, register: bool = True):
        import numpy as np
        import torch.distributed as dist
        from multiprocessing import resource_tracker, shared_memory
        self.group = group
        self.rank = dist.get_rank(group)
        nbytes = 4
        for d in shape:
            nbytes *= int(d)
        name = [None]
        self.shm = None
        if self.rank == 0:
            # a full /dev/shm does not fail at creation but with SIGBUS at first touch: check the space first
            try:
                import os
                vfs = os.statvfs("/dev/shm")
                if vfs.f_bavail * vfs.f_frsize >= nbytes + (64 << 20):
                    self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
                    name[0] = self.shm.name
            except Exception:
                self.shm = None
        dist.broadcast_object_list(name, src=0, group=group)
        if name[0] is None:          # the same outcome on every rank
            raise RuntimeError("no room for the shared host stack in /dev/shm")
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
            # only the creator owns the segment: keep this process's resource tracker from unlinking it at exit
            try:
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.nbytes = nbytes
        self.array = np.ndarray(tuple(int(d) for d in shape), np.float32, buffer=self.shm.buf)
        self.ptr = self.array.ctypes.data
        self.registered = False
        if register:
            import torch
            if torch.cuda.is_available():
                rc = torch.cuda.cudart().cudaHostRegister(self.ptr, nbytes, 0)
                if int(rc) != 0:
                    raise RuntimeError(f"cudaHostRegister(shared stack) failed: {rc}")
                self.registered = True
        dist.barrier(group)

    def close(self):
        import torch.distributed as dist
        if getattr(self, "shm", None) is None:
            return
        if self.registered:
            import torch
            torch.cuda.cudart().cudaHostUnregister(self.ptr)
            self.registered = False
        self.array = None
        dist.barrier(self.group)
        self.shm.close()
        if self.rank == 0:
            self.shm.unlink()
        self.shm = None


def agree_on_failure(err, group=None):
    """The reference aborts the WHOLE stack when one frame fails (`?` on find_transform_ecc,
    /root/reference/src/lib.rs:777).  With the frames sharded over ranks only the rank that owns the failing frame
    sees the error, so the ranks agree on it here (collective; `err` = this rank's exception or None): if any rank
    failed, EVERY rank raises — the failing ranks their own error, the others an OpenCvError naming it — and nobody
    enters a collective the failing rank will never reach."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if err is not None:
            raise err
        return
    mine = None if err is None else (dist.get_rank(group), type(err).__name__, str(err))
    votes = [None] * dist.get_world_size(group)
    dist.all_gather_object(votes, mine, group=group)
    first = next((v for v in votes if v is not None), None)
    if first is None:
        return
    if err is not None:
        raise err
    from .api import OpenCvError
    raise OpenCvError(f"the stack was aborted: rank {first[0]} failed with {first[1]}: {first[2]}")


def stack_on_ranks(stack, frames_by_index, n_frames: int, rank: int, world: int, device=None, peers: bool = False):
    """Run this rank's shard through `stack` (an EccStack whose reference is already set; rank 0's was
    created with seed_reference=True, the others with False), reduce, and return the final HxWxC f32 torch
    tensor on rank 0 (None elsewhere).  `frames_by_index[i]` is frame i as a host array or CUDA tensor.
    `peers=True` (after connect_peers succeeded) uses the fused peer-memory exchange; the returned tensor then
    views the library's output buffer (valid until the next exchange)."""
    import torch
    for i in shard_frames(n_frames, rank, world):
        stack.submit(frames_by_index[i], tag=i)
    from .api import StackerError
    err = None
    if peers:
        # the exchange is queued on every rank whatever happened to the frames (a failed frame contributes nothing),
        # so no rank is left waiting for a peer; the per-frame errors surface at sync() and are agreed on afterwards
        d_out = stack.peer_reduce(n_frames)
        try:
            stack.sync()
        except StackerError as e:
            err = e
        agree_on_failure(err)
        if rank != 0:
            return None
        n = stack.height * stack.width * stack.channels
        return torch.as_tensor(DevicePtrArray(d_out, n), device=device).view(stack.height, stack.width, stack.channels)
    ptr = n = None
    try:
        ptr, n = stack.partial()
    except StackerError as e:
        err = e
    agree_on_failure(err)        # before the reduce: a rank that failed must not leave the others inside the collective
    part = torch.as_tensor(DevicePtrArray(ptr, n), device=device)
    reduce_partial_stack(part, 0)
    if rank != 0:
        return None
    out = torch.empty(stack.height, stack.width, stack.channels, dtype=torch.float32, device=part.device)
    stack.finish_device(ptr, n_frames, out.data_ptr())
    return out
