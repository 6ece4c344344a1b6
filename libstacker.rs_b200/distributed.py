"""Multi-GPU plumbing for one stack: one process per GPU (torch.distributed, NCCL over NVLink), frames
sharded across ranks, ONE sum-reduce of the partial stacks to rank 0, then the divide.

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


def stack_on_ranks(stack, frames_by_index, n_frames: int, rank: int, world: int, device=None):
    """Run this rank's shard through `stack` (an EccStack whose reference is already set; rank 0's was
    created with seed_reference=True, the others with False), reduce, and return the final HxWxC f32 torch
    tensor on rank 0 (None elsewhere).  `frames_by_index[i]` is frame i as a host array or CUDA tensor."""
    import torch
    for i in shard_frames(n_frames, rank, world):
        stack.submit(frames_by_index[i], tag=i)
    ptr, n = stack.partial()
    part = torch.as_tensor(DevicePtrArray(ptr, n), device=device)
    reduce_partial_stack(part, 0)
    if rank != 0:
        return None
    out = torch.empty(stack.height, stack.width, stack.channels, dtype=torch.float32, device=part.device)
    stack.finish_device(ptr, n_frames, out.data_ptr())
    return out
