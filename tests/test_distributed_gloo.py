"""N > 1 host logic on CPU: frame sharding + the single sum-reduce + divide (world_size 2, gloo).

The per-rank compute is stood in by the oracle (tests may call it); what is under test is the plumbing in
libstacker.rs_b200/distributed.py that bench.py and a multi-GPU caller use unchanged with the nccl backend."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_frames, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    from oracle import restate as R, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = ge.load_package().distributed
    w, h, motion = 96, 64, 0
    frames = synth.Stack(w, h, n_frames, motion, seed=21).frames()
    grey0 = R.bgr2gray_u8(frames[0])
    # rank 0 seeds its partial stack with the unwarped reference frame (src/lib.rs:752-754)
    acc = R.to_f32_unit(frames[0]) if rank == 0 else np.zeros((h, w, 3), np.float32)
    mine = D.shard_frames(n_frames, rank, world)
    for i in mine:
        _, m, _ = R.find_transform_ecc(R.bgr2gray_u8(frames[i]), grey0, motion, R.term_criteria(30, 1e-4), 5)
        acc = acc + R.final_warp(frames[i], m, motion)
    part = torch.from_numpy(acc.reshape(-1).copy())
    D.reduce_partial_stack(part, 0)
    if rank == 0:
        out = (part.numpy() * np.float32(1.0 / n_frames)).reshape(h, w, 3)
        q.put((mine, out))
    else:
        q.put((mine, None))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 2])
def test_two_rank_stack_equals_single_rank(n_frames):
    import torch.multiprocessing as mp
    from oracle import restate as R, synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = sorted(sum((r[0] for r in results), []))
    assert shards == list(range(1, n_frames))            # every non-reference frame exactly once
    out = next(r[1] for r in results if r[1] is not None)
    frames = synth.Stack(96, 64, n_frames, 0, seed=21).frames()
    want, _, _ = R.ecc_match(frames, 0, 30, 1e-4, 5)
    assert np.abs(out - want).max() <= 1e-6              # f32 summation order only


def test_shard_frames_partition(pkg):
    D = pkg.distributed
    for n in (1, 2, 7, 64):
        for world in (1, 2, 4, 8):
            got = sorted(sum((D.shard_frames(n, r, world) for r in range(world)), []))
            assert got == list(range(1, n))
            sizes = [len(D.shard_frames(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_peer_slices_cover_the_stack(pkg):
    """stk_ecc_peer_reduce's slice rule (distributed.slice_bounds mirrors it): contiguous, 4-float aligned
    starts, every float exactly once, the tail on the last rank, no slice for the root of a world > 2."""
    D = pkg.distributed
    for n in (0, 3, 4, 17, 96 * 64 * 3, 3840 * 2160 * 3, 1001 * 3):
        for world in (1, 2, 3, 4, 8, 16):
            edges = [D.slice_bounds(n, r, world) for r in range(world)]
            if world > 2:
                assert edges[0] == (0, 0)
                edges = edges[1:]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (b0, e0), (b1, e1) in zip(edges, edges[1:]):
                assert e0 == b1 and b1 % 4 == 0 and b0 <= e0


def _handle_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = ge.load_package().distributed
    mine = bytes([rank]) * 256                       # stands in for EccStack.peer_export()
    q.put((rank, D.gather_handles(mine)))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_handles_are_gathered_in_rank_order():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_handle_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, handles in results:
        assert handles == [bytes([0]) * 256, bytes([1]) * 256]


def _shared_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = ge.load_package().distributed
    shape = (7, 5, 3)
    shared = D.SharedHostStack(shape, register=False)        # no CUDA here: the mapping itself is under test
    n = 7 * 5 * 3
    b, e = D.scatter_bounds(n, rank, world)
    shared.array.reshape(-1)[b:e] = np.arange(b, e, dtype=np.float32) + 1000 * rank    # "this rank's slice"
    dist.barrier()
    if rank == 0:
        q.put(shared.array.reshape(-1).copy())
    dist.barrier()
    shared.close()
    dist.destroy_process_group()


def test_shared_host_stack_collects_every_ranks_slice(pkg):
    import torch.multiprocessing as mp
    D = pkg.distributed
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shared_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 7 * 5 * 3
    want = np.empty(n, np.float32)
    for r in range(2):
        b, e = D.scatter_bounds(n, r, 2)
        want[b:e] = np.arange(b, e, dtype=np.float32) + 1000 * r
    assert np.array_equal(got, want)
    for world in (1, 2, 3, 8):
        edges = [D.scatter_bounds(n, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n and all(a[1] == b[0] for a, b in zip(edges, edges[1:]))


def _failure_worker(rank, world, port, failing_rank, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = ge.load_package()
    err = pkg.OpenCvError("findTransformECC: the algorithm stopped before its convergence") if rank == failing_rank else None
    try:
        pkg.distributed.agree_on_failure(err)
        q.put((rank, None))
    except pkg.StackerError as e:
        q.put((rank, str(e)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("failing_rank", [1, None])
def test_a_failed_frame_aborts_the_stack_on_every_rank(failing_rank):
    """src/lib.rs:777: one frame's ECC failure fails the whole call.  Sharded over ranks, every rank must raise (and
    none may be left inside the reduce): distributed.agree_on_failure, used by stack_on_ranks before the exchange."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_failure_worker, args=(r, 2, port, failing_rank, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if failing_rank is None:
        assert results == {0: None, 1: None}
    else:
        assert "stopped before its convergence" in results[0] and "rank 1" in results[0]
        assert "stopped before its convergence" in results[1]
