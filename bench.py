#!/usr/bin/env python
"""Benchmark of the ECC align-and-stack hot path (BASELINE.json metric: ECC-aligned+stacked frames/s at 4K).

  python bench.py --gpus N --steps K --warmup W          this repo's CUDA path
  python bench.py --impl reference ...                   the reference's CPU path (OpenCV through cv2, driven
                                                         call-for-call like /root/reference/src/lib.rs:719-847)

A "step" = one whole stack: BASELINE configs[3] — ecc_match MotionType::Homography, max_count 5000, eps 1e-5,
gauss_filt_size 5, on 64 synthetic 3840x2160 BGR frames (synthetic.py, seeds fixed) — reference prep + seed,
63 x {prep, device ECC loop, final warp + accumulate}, lane sum, [reduce over ranks], divide.
`value`  : frames/s with every frame already resident in HBM as u8 BGR (CUDA-event time, max over ranks).
`e2e`    : the same through the host-facing API with frames in pinned HOST memory and the stacked image copied
           back to the host every step.
N > 1    : the 64 frames are sharded over the ranks (strong scaling), one NCCL sum-reduce of the partial stacks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ECC-aligned+stacked frames/s at 4K"
UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # workload overrides (anything but the defaults is reported in config and is not the headline)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--motion", type=int, default=3)
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("STK_LANES", "4")))
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="frames in the CPU sample (0 = one per core, <= 32)")
    return ap.parse_args()


def workload_name(a):
    base = (f"ecc_match MotionType::{['Translation', 'Euclidean', 'Affine', 'Homography'][a.motion]}, "
            f"{a.frames} synthetic {a.width}x{a.height} BGR frames, max_count 5000, eps 1e-5, gauss_filt_size 5")
    if (a.frames, a.width, a.height, a.motion) == (64, 3840, 2160, 3):
        return "BASELINE configs[3]: " + base
    return "NON-HEADLINE override: " + base


def make_stack(a):
    import synthetic as synth
    return synth.Stack(a.width, a.height, a.frames, a.motion, seed=4)


# ---- clocks -----------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML: one sample at the
    start, one every 100 ms, one at the end).  NVML queries take a driver lock, so the period is kept long and
    only rank 0 samples: at 8 ranks x 50 Hz the polling itself produced multi-millisecond stragglers."""

    def __init__(self, device_index: int):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            import torch
            uuid = torch.cuda.get_device_properties(device_index).uuid
            self._nv = pynvml
            try:
                self._h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None

    def _loop(self):
        nv = self._nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self._h is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def _sample_once(self):
        try:
            self.samples.append(self._nv.nvmlDeviceGetClockInfo(self._h, self._nv.NVML_CLOCK_SM))
        except Exception:
            pass

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        if self._h is not None:
            self._sample_once()
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# ---- CPU reference arm --------------------------------------------------------------------------------------
def cpu_sample_run(a, n_sample: int):
    """The reference's CPU path on a bounded sample of the workload: frame 0 + (n_sample-1) frames of the
    same stack, one task per frame on all host cores.  Returns (frames/s, seconds, workers, n_sample)."""
    import cv2  # noqa: F401
    from oracle import cvref
    stack = make_stack(a)
    frames = [stack.frame(i) for i in range(n_sample)]
    workers = min(os.cpu_count() or 1, max(1, n_sample - 1))
    t0 = time.perf_counter()
    _, warps, _ = cvref.ecc_match(frames, a.motion, 5000, 1e-5, 5, workers=workers)
    dt = time.perf_counter() - t0
    cpu_sample_run.warps = {i: w for i, w in enumerate(warps) if w is not None}      # the checker's matrices, tag -> warp
    return n_sample / dt, dt, workers, n_sample


cpu_sample_run.warps = {}


def oracle_warp_error(a, res, tags):
    """max corner displacement (px) between this run's recovered matrices and the reference engine's
    (cv2.findTransformECC driven as src/lib.rs:769-777, oracle/cvref.py) for the given frame tags — the checker
    leg of the bench (outside every timed region)."""
    import synthetic as synth
    from oracle import cvref
    stack = make_stack(a)
    have = cpu_sample_run.warps
    g0 = None
    worst, n = 0.0, 0
    by_tag = {r["tag"]: r["warp"] for r in res}
    for t in tags:
        if t not in by_tag:
            continue
        if t in have:
            m_ref = have[t]
        else:
            import cv2
            if g0 is None:
                g0 = cv2.cvtColor(stack.frame(0), cv2.COLOR_BGR2GRAY)
            _, m_ref = cvref.align_frame(cv2.cvtColor(stack.frame(t), cv2.COLOR_BGR2GRAY), g0, a.motion,
                                         cvref.term_criteria(5000, 1e-5), 5)
        mine = by_tag[t] if a.motion == 3 else by_tag[t][:2]
        worst = max(worst, synth.corner_displacement(mine, m_ref, a.width, a.height))
        n += 1
    return worst, n


def default_cpu_sample(a):
    if a.cpu_sample > 0:
        return a.cpu_sample
    return min(a.frames, min(os.cpu_count() or 1, 32) + 1)


def run_reference(a, rank, world):
    if rank != 0:
        return
    try:
        import cv2
    except Exception as e:  # pragma: no cover
        print(json.dumps({"impl": "reference", "unavailable": f"cv2 (the reference's OpenCV engine) not importable: {e}"}))
        return
    n_sample = default_cpu_sample(a)
    vals, secs, workers = [], [], 1
    for i in range(a.warmup + a.steps):
        v, dt, workers, _ = cpu_sample_run(a, n_sample)
        if i >= a.warmup:
            vals.append(v)
            secs.append(dt)
    value = len(vals) * n_sample / sum(secs)
    sample = (f"frame 0 + {n_sample - 1} of the {a.frames} frames of the workload per step, frames in host memory, "
              f"ThreadPool({workers}) one task per frame + OpenCV's own pool ({cv2.getNumThreads()} threads), cv2 {cv2.__version__}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "host_cores": os.cpu_count()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample,
                         "note": "the Rust crate cannot be built here (no rustc/cargo, no C++ OpenCV); this is the same "
                                 "OpenCV engine (cv2) driven call-for-call as src/lib.rs:719-847 drives it"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---- this repo's arm ------------------------------------------------------------------------------------------
def run_b200(a, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    import synthetic as synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = ge.load_package()
    D = pkg.distributed
    # one process per GPU: allocate this rank's pinned frame buffers on the GPU's own NUMA node
    numa_bound = D.bind_to_gpu_numa_node(local_rank) if (world > 1 and os.environ.get("STK_NUMA_BIND", "1") != "0") else False

    n, w, h = a.frames, a.width, a.height
    n_px = w * h
    stack_src = make_stack(a)
    mine = D.shard_frames(n, rank, world)
    host = {0: stack_src.frame(0)}
    for i in mine:
        host[i] = stack_src.frame(i)
    dev_frames = {i: torch.from_numpy(f).to(dev) for i, f in host.items()}
    pinned = {i: torch.from_numpy(f).pin_memory() for i, f in host.items()}
    pinned_np = {i: t.numpy() for i, t in pinned.items()}
    params = pkg.EccMatchParameters(pkg.MotionType(a.motion), 5000, 1e-5, 5)
    st = pkg.EccStack(w, h, 3, params, device=local_rank, lanes=a.lanes, seed_reference=(rank == 0))
    out_dev = torch.empty(h, w, 3, dtype=torch.float32, device=dev)
    out_host = torch.empty(h, w, 3, dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    trace = os.environ.get("STK_BENCH_TRACE") == "1"

    # the exchange step: the library's fused reduce-scatter + divide over NVLink peer memory (default), or the
    # NCCL reduce + scale kernel it replaces (STK_REDUCE=nccl, and the fallback when peers cannot be mapped)
    def connect(stack):
        """Map the peers (N > 1, collective) — or, on one GPU, make the context a world of one: its "exchange" is the lane
        sum fused with the divide on the exchange stream, so that consecutive stacks queue back to back like at N > 1."""
        if world > 1:
            return D.connect_peers(stack)
        stack.peer_connect(0, 1, [stack.peer_export()])
        return True

    use_peers = False
    if os.environ.get("STK_REDUCE", "peer") != "nccl":
        use_peers = connect(st)
        if not use_peers and rank == 0:
            print(f"peer exchange unavailable, using NCCL: {D.connect_peers.last_failure}", file=sys.stderr)
    peer_out = {}
    shared_host = None
    if use_peers and world > 1:
        try:
            shared_host = D.SharedHostStack((h, w, 3))
        except RuntimeError as e:       # same outcome on every rank: fall back to the root's own copy-out
            if rank == 0:
                print(f"shared host stack unavailable ({e}); the root copies the whole stack out", file=sys.stderr)

    def peer_result():
        return torch.as_tensor(D.DevicePtrArray(peer_out["ptr"], h * w * 3), device=dev).view(h, w, 3)

    def step_resident():
        t = [time.perf_counter()]
        st.reset()
        st.set_reference(dev_frames[0])
        t.append(time.perf_counter())
        for i in mine:
            st.submit(dev_frames[i], tag=i)
        t.append(time.perf_counter())
        if use_peers:
            peer_out["ptr"] = st.peer_reduce(n)       # asynchronous; the next step's reset() joins the lanes
            if trace:
                torch.cuda.synchronize()
                t.append(time.perf_counter())
                print(f"TRACE rank{rank} " + " ".join(f"{k}={1e3 * (b - a):.3f}ms" for k, a, b in
                                                      zip(["reset+set_reference", "submit", "peer exchange"], t, t[1:])), file=sys.stderr)
            return
        ptr, nfl = st.partial()
        t.append(time.perf_counter())
        if world > 1:
            part = torch.as_tensor(D.DevicePtrArray(ptr, nfl), device=dev)
            D.reduce_partial_stack(part, 0)
            torch.cuda.synchronize()
        t.append(time.perf_counter())
        if rank == 0:
            st.finish_device(ptr, n, out_dev.data_ptr())
        t.append(time.perf_counter())
        if trace:
            names = ["reset+set_reference", "submit", "partial(sync+lane sum)", "reduce", "finish"]
            print(f"TRACE rank{rank} " + " ".join(f"{k}={1e3 * (b - a):.3f}ms" for k, a, b in zip(names, t, t[1:])), file=sys.stderr)

    # ---- end-to-end leg: host frames in, finished stack back on the host, through EccStack -------------------------
    # Two contexts take the stacks alternately: while one stack's tail (last alignments, exchange, copy-out of the result)
    # is still in flight, the uploads of the next stack are already queued on the other context — PCIe is full duplex and
    # the host never waits for a stack before feeding the next.  Every stack's result is read back inside the timed
    # region (the last one by the drain).  N > 1: frame 0 is uploaded by rank 0 only and reaches the other ranks over
    # NVLink (NCCL broadcast on torch's stream; set_reference orders itself behind that stream, ABI v5).
    e2e_slots = []

    def make_e2e_slots():
        stacks = [st]
        if os.environ.get("STK_E2E_PINGPONG", "1") != "0":
            st2 = pkg.EccStack(w, h, 3, params, device=local_rank, lanes=a.lanes, seed_reference=(rank == 0))
            ok2 = (not use_peers) or connect(st2)
            if ok2:
                stacks.append(st2)
            else:
                st2.close()
        for k, s_ in enumerate(stacks):
            sh = shared_host if k == 0 else None
            if use_peers and shared_host is not None and k > 0:
                try:
                    sh = D.SharedHostStack((h, w, 3))
                except RuntimeError:
                    sh = None
            e2e_slots.append({"st": s_, "shared": sh, "ref": torch.empty_like(dev_frames[0]) if world > 1 else None,
                              "out_dev": out_dev if k == 0 else torch.empty_like(out_dev),
                              "out_host": out_host if k == 0 else torch.empty(h, w, 3, dtype=torch.float32).pin_memory(), "ptr": None})
        if use_peers and any(sl["shared"] is None for sl in e2e_slots):      # the same copy-out form on every slot
            for sl in e2e_slots:
                sl["shared"] = None

    def e2e_queue(sl):
        c = sl["st"]
        c.reset()
        if world > 1:
            if rank == 0:
                sl["ref"].copy_(pinned[0], non_blocking=True)
            dist.broadcast(sl["ref"], src=0)
            c.set_reference(sl["ref"])
        else:
            c.set_reference(pinned_np[0])
        for i in mine:
            c.submit(pinned_np[i], tag=i, pinned=True)
        if use_peers and sl["shared"] is not None:
            # every rank keeps its slice of the finished stack and copies it out over its OWN PCIe link into the
            # host stack all ranks map; rank 0 owns the result once every rank's copy has landed
            c.peer_reduce_scatter(n)
            c.peer_slice_to_host(sl["shared"].ptr)
        elif use_peers:
            sl["ptr"] = c.peer_reduce(n)

    def e2e_complete(sl):
        c = sl["st"]
        if use_peers and sl["shared"] is not None:
            c.sync()
            dist.barrier()
            return
        if use_peers:
            c.sync()
            if rank == 0:
                sl["out_host"].copy_(torch.as_tensor(D.DevicePtrArray(sl["ptr"], h * w * 3), device=dev).view(h, w, 3), non_blocking=False)
            return
        ptr, nfl = c.partial()
        if world > 1:
            part = torch.as_tensor(D.DevicePtrArray(ptr, nfl), device=dev)
            D.reduce_partial_stack(part, 0)
            torch.cuda.synchronize()
        if rank == 0:
            c.finish_device(ptr, n, sl["out_dev"].data_ptr())
            sl["out_host"].copy_(sl["out_dev"], non_blocking=False)

    e2e_state = {"k": 0, "pending": None}

    def step_e2e():
        sl = e2e_slots[e2e_state["k"] % len(e2e_slots)]
        e2e_state["k"] += 1
        if e2e_state["pending"] is sl:            # one context only: finish the stack before reusing it
            e2e_complete(sl)
            e2e_state["pending"] = None
        e2e_queue(sl)
        prev, e2e_state["pending"] = e2e_state["pending"], sl
        if prev is not None:
            e2e_complete(prev)

    def drain_e2e():
        if e2e_state["pending"] is not None:
            e2e_complete(e2e_state["pending"])
            e2e_state["pending"] = None

    if world > 1 and not use_peers:
        # NCCL sets up channels lazily over its first collectives on a buffer (measured: one 13 ms reduce among
        # the first five); do that outside the timed region, on the very buffer the steps reduce
        step_resident()
        ptr0, nfl0 = st.partial()
        warm = torch.as_tensor(D.DevicePtrArray(ptr0, nfl0), device=dev)
        for _ in range(10):
            D.reduce_partial_stack(warm, 0)
        torch.cuda.synchronize()
        dist.barrier()

    def timed(fn, steps, warmup, sample_clocks, drain=None):
        for _ in range(warmup):
            fn()
        if drain:
            drain()
        # NVML set-up takes milliseconds: do it BEFORE the barrier, or rank 0 enters the timed region late and
        # every other rank's first exchange waits for it (measured: +1 ms per step on the max-over-ranks time)
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = st.launch_count()
        e0.record()
        t0 = time.perf_counter()
        launches = 0
        for _ in range(steps):
            fn()
            launches += st.launch_count()     # reset() zeroes the counter at the start of every step
        if drain:
            drain()                           # the last stack's result is read back INSIDE the timed region
        torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if use_peers:
            # the peer path never synchronises inside a step, so the device-launched iteration kernels of a step
            # are only accounted once its results are in: every step does the same work
            st.sync()
            launches = steps * st.launch_count()
        clocks = sampler.stop() if sampler else None
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        del l0
        return float(ms.item()), wall * 1e3, launches, clocks

    ms, wall_ms, launches, clocks = timed(step_resident, a.steps, a.warmup, True)
    value = n * a.steps / (ms / 1e3)

    # diagnostic (untimed): what each rank's own share costs with no exchange at the end — the spread over the ranks is
    # the frame-granular imbalance, max(rank_work_ms) against ms_per_step is what the exchange and the coupling cost
    rank_work_ms = None
    if world > 1:
        def own_share():
            st.reset()
            st.set_reference(dev_frames[0])
            for i in mine:
                st.submit(dev_frames[i], tag=i)
            st.sync()
        own_share()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            own_share()
        e1.record()
        torch.cuda.synchronize()
        own = torch.zeros(world, dtype=torch.float64, device=dev)
        own[rank] = e0.elapsed_time(e1) / 3
        dist.all_reduce(own)
        rank_work_ms = [round(float(v), 4) for v in own.tolist()]
        barrier()
    res = st.results()
    iters = [r["iterations"] for r in res]
    # the run must have aligned the frames for real: recovered warps vs the ground truth of the generator
    truth_err = max((synth.corner_displacement(r["warp"] if a.motion == 3 else r["warp"][:2],
                                               stack_src.truth[r["tag"]], w, h) for r in res), default=0.0)
    statuses = sorted({r["status"] for r in res})

    e2e = None
    shared_host_used = shared_host is not None
    if not a.skip_e2e:
        make_e2e_slots()
        shared_host_used = use_peers and e2e_slots[0]["shared"] is not None
        ems, _, _, _ = timed(step_e2e, a.steps, max(2, min(a.warmup, 2)), False, drain_e2e)
        h2d = sum(pinned_np[i].nbytes for i in ([0] if (rank == 0 or world == 1) else []) + mine)
        e2e = {"value": n * a.steps / (ems / 1e3), "unit": UNIT, "ms_per_step": ems / a.steps,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(out_host.numel() * 4) if rank == 0 else 0,
               "contexts": len(e2e_slots),
               "api": ("EccStack.set_reference/submit(pinned host frames)/peer_reduce_scatter/peer_slice_to_host: every rank "
                       "copies its slice of the stack into the shared pinned host stack" if shared_host_used else
                       "EccStack.set_reference/submit(pinned host frames)/partial/finish_device + D2H of the stack") +
                      (f"; {len(e2e_slots)} contexts take the stacks alternately (a stack's tail and copy-out overlap the next stack's uploads)"
                       if len(e2e_slots) > 1 else "") +
                      ("; frame 0 is uploaded by rank 0 only and broadcast over NVLink (NCCL)" if world > 1 else "")}
        if world > 1:
            tot = torch.tensor([float(h2d)], dtype=torch.float64, device=dev)
            dist.all_reduce(tot)
            e2e["h2d_bytes_per_step"] = int(tot.item())

    # bare host->device ceiling for this run's pinned frame buffers (all ranks at once, like the e2e leg)
    h2d_probe = None
    if not a.skip_e2e:
        scratch = torch.empty_like(dev_frames[0])
        order = [0] + mine

        def h2d_only():
            for i in order:
                scratch.copy_(pinned[i], non_blocking=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h2d_only()
        barrier()
        e0.record()
        for _ in range(3):
            h2d_only()
        e1.record()
        torch.cuda.synchronize()
        t_ms = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        h2d_probe = {"ms_per_step": float(t_ms.item()), "GBps_aggregate": e2e["h2d_bytes_per_step"] / (float(t_ms.item()) * 1e-3) / 1e9,
                     "frames_per_s_ceiling": n / (float(t_ms.item()) * 1e-3),
                     "how": "the same pinned frame buffers copied to the device with nothing else running, every rank at once, max over ranks"}
        e2e["frac_of_h2d_ceiling"] = e2e["value"] / h2d_probe["frames_per_s_ceiling"]
        e2e["h2d_probe"] = h2d_probe
        del scratch

    # the plugin call itself, one-shot, as a user of the reference would make it: ecc_match(frames, params, None)
    # (creates and destroys its context inside the call, like the reference allocates per call)
    e2e_api = None
    if world == 1 and not a.skip_e2e:
        frames_list = [pinned_np[i] for i in range(n)]
        api_out = out_host.numpy()
        pkg.ecc_match(frames_list, params, None, device=local_rank, pinned=True, out=api_out)      # warm-up
        t0 = time.perf_counter()
        reps = max(1, min(a.steps, 3))
        for _ in range(reps):
            pkg.ecc_match(frames_list, params, None, device=local_rank, pinned=True, out=api_out)
        dt = (time.perf_counter() - t0) / reps
        e2e_api = {"value": n / dt, "unit": UNIT, "ms_per_call": dt * 1e3,
                   "api": "ecc_match(list of pinned host arrays, EccMatchParameters, None, out=pinned) — one call per stack, "
                          "context creation and destruction inside the call, wall clock",
                   "stack_mean": float(api_out.mean(dtype=np.float64))}

    stack_mean = float((peer_result() if use_peers else out_dev).mean().item()) if rank == 0 else None
    # ---- correctness of THIS run (outside every timed region) ------------------------------------------------------
    # N > 1: rank 0 redoes the whole stack on its one device and compares the 8-bit stacks and every matrix with what
    # the ranks produced together; all N: one or more matrices against the reference engine (cv2, oracle/cvref.py)
    multi_check = None
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, [(r["tag"], r["warp"].tolist(), r["iterations"]) for r in res])
        if rank == 0:
            multi = (peer_result() if use_peers else out_dev).clone()
            with pkg.EccStack(w, h, 3, params, device=local_rank, lanes=a.lanes, seed_reference=True) as s1:
                s1.set_reference(dev_frames[0])
                for i in range(1, n):
                    fr = dev_frames[i] if i in dev_frames else torch.from_numpy(stack_src.frame(i)).to(dev)
                    s1.submit(fr, tag=i)
                    if i not in dev_frames:
                        s1.sync()             # the temporary frame is released right after its alignment
                single = torch.from_numpy(s1.finish(n)).to(dev)
                single_res = {r["tag"]: r for r in s1.results()}
            d8 = (torch.round(multi * 255.0) - torch.round(single * 255.0)).abs().max().item()
            worst, seen = 0.0, set()
            for shard in gathered:
                for tag, wm, it in shard:
                    seen.add(tag)
                    worst = max(worst, synth.corner_displacement(np.array(wm, np.float32), single_res[tag]["warp"], w, h))
            multi_check = {"max_abs_diff_8bit_vs_single_gpu": float(d8), "max_abs_diff_f32_vs_single_gpu": float((multi - single).abs().max().item()),
                           "max_corner_diff_vs_single_gpu_px": worst, "frames_covered": len(seen), "frames_expected": n - 1}
            del multi, single
    e2e_mean = None
    if rank == 0 and e2e_slots:
        # every slot holds a finished copy of the same stack: check the one written last
        sl = e2e_slots[(e2e_state["k"] - 1) % len(e2e_slots)]
        e2e_mean = float((sl["shared"].array if (use_peers and sl["shared"] is not None) else sl["out_host"].numpy()).mean(dtype=np.float64))
    if use_peers:
        barrier()                 # nobody unmaps while a peer may still be inside an exchange
        closed = set()
        for sl in e2e_slots:
            if sl["shared"] is not None and id(sl["shared"]) not in closed:
                closed.add(id(sl["shared"]))
                sl["shared"].close()
        if shared_host is not None and id(shared_host) not in closed:
            shared_host.close()
        for sl in e2e_slots:
            if sl["st"] is not st:
                sl["st"].peer_disconnect()
        st.peer_disconnect()
        barrier()
    for sl in e2e_slots:
        if sl["st"] is not st:
            sl["st"].close()
    if world > 1:
        pr = torch.zeros(world, dtype=torch.float64, device=dev)
        pr[rank] = float(sum(iters))
        dist.all_reduce(pr)
        iters_per_rank_pre = [int(v) for v in pr.tolist()]
    else:
        iters_per_rank_pre = [sum(iters)]
    # ---- roofline of the dominant kernel (the ECC iteration kernel), measured alone: 1 lane, CUDA events per stage ----
    st.close()
    roof = stages = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    if rank == 0:
        with pkg.EccStack(w, h, 3, params, device=local_rank, lanes=1, seed_reference=True) as s1:
            for rep in range(2):          # first pass warms up
                s1.reset()
                s1.set_profiling(rep == 1)
                s1.set_reference(dev_frames[0])
                for i in mine[:16]:
                    s1.submit(dev_frames[i], tag=i)
                s1.sync()
            t = s1.stage_times()
        if t["frames"] and t["iterations"]:
            per_iter_ms = t["loop_ms"] / t["iterations"]
            achieved = 8.0 * n_px / (per_iter_ms * 1e-3) / 1e9
            traffic = None
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "ecc_iter_traffic.json"))).get("dram_bytes_per_launch")
            except Exception:
                pass
            # in the timed step the lanes overlap one frame's serial tail and relaunch gap with other frames' pixel phases:
            # what one iteration costs THERE = (device time of the step - the other kernels' own times) / iterations
            other_ms = (t["prep_ms"] + t["warp_ms"]) / t["frames"] * (n - 1) / world
            step_ms = ms / a.steps
            iters_here = max(iters_per_rank_pre) if iters_per_rank_pre else 0
            in_step_us = 1e3 * (step_ms - other_ms) / iters_here if iters_here else None
            roof = {"bound": "hbm", "kernel": "ecc_iter_v2_kernel<Homography>" if a.motion == 3 else f"ecc_iter_v2_kernel<{a.motion}>",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "algorithmic_bytes_per_launch": 8 * n_px, "us_per_launch": per_iter_ms * 1e3,
                    "peak_source": peak_src,
                    "how": "ONE LANE (the conservative figure): CUDA events on the lane stream around each frame's device loop "
                           "(init + all iterations of the graph WHILE node, serial tail and relaunch gaps included) / iterations; "
                           f"{t['frames']} frames, {t['iterations']} iterations",
                    "in_step": None if in_step_us is None else {
                        "us_per_launch": in_step_us, "achieved": 8.0 * n_px / (in_step_us * 1e-6) / 1e9,
                        "frac": 8.0 * n_px / (in_step_us * 1e-6) / 1e9 / peak,
                        "dominant_kernel_ms_per_step": in_step_us * iters_here * 1e-3, "ms_per_step": step_ms,
                        "how": f"inside the timed {a.lanes}-lane step: (ms_per_step - (prep + warp) per frame measured on one lane x frames "
                               "of the slowest rank) / ECC iterations of the slowest rank; <= ms_per_step by construction"}}
            stages = {
                "prep": {"ms_per_frame": t["prep_ms"] / t["frames"], "GBps": 7.0 * n_px / (t["prep_ms"] / t["frames"] * 1e-3) / 1e9},
                "ecc_loop": {"ms_per_frame": t["loop_ms"] / t["frames"], "iterations_per_frame": t["iterations"] / t["frames"]},
                "warp_accumulate": {"ms_per_frame": t["warp_ms"] / t["frames"], "GBps": 27.0 * n_px / (t["warp_ms"] / t["frames"] * 1e-3) / 1e9},
            }

    if world > 1:
        it_t = torch.tensor([float(sum(iters))], dtype=torch.float64, device=dev)
        dist.all_reduce(it_t)
        total_iters = int(it_t.item())
        per_rank = torch.zeros(world, dtype=torch.float64, device=dev)
        per_rank[rank] = float(sum(iters))
        dist.all_reduce(per_rank)
        iters_per_rank = [int(v) for v in per_rank.tolist()]
        l_t = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
        dist.all_reduce(l_t)
        launches = int(l_t.item())
    else:
        total_iters = sum(iters)
        iters_per_rank = [total_iters]

    cpu = None
    if rank == 0 and world == 1 and not a.skip_cpu:
        try:
            import cv2
            n_sample = default_cpu_sample(a)
            v, dt, workers, _ = cpu_sample_run(a, n_sample)
            cpu = {"value": v, "unit": UNIT, "cores": workers, "kind": "port",
                   "sample": f"frame 0 + {n_sample - 1} frames of the same stack in {dt:.1f} s, one task per frame on "
                             f"{workers} threads + OpenCV pool ({cv2.getNumThreads()}), cv2 {cv2.__version__}; "
                             f"host has {os.cpu_count()} cores"}
        except Exception as e:  # pragma: no cover
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    oracle_check = None
    if rank == 0 and not a.skip_cpu:
        try:
            # N = 1: every frame of the CPU sample (its matrices are a by-product of the cpu_baseline leg); N > 1: one
            # frame through cv2.findTransformECC (~20 s on the host cores)
            tags = sorted(cpu_sample_run.warps) if cpu_sample_run.warps else [mine[0] if mine else 1]
            err, cnt = oracle_warp_error(a, res, tags)
            oracle_check = {"max_corner_error_vs_oracle_px": err, "frames": cnt,
                            "oracle": "cv2.findTransformECC driven as src/lib.rs:769-777 (oracle/cvref.py)"}
        except Exception as e:  # pragma: no cover
            oracle_check = {"error": str(e)}

    if rank == 0:
        # algorithmic bytes of the whole step: per frame (34 + 8K)N, per stack 43N per GPU
        alg_bytes = (34 * (n - 1) + 8 * total_iters + 43 * world) * n_px
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "lanes": a.lanes,
                       "l2": ("inputs larger than L2: every step reads all frames (%.2f GB u8) from HBM" % (n * n_px * 3 / 1e9)) if world == 1 else
                             ("inputs larger than L2 on every GPU: a step reads all %d frames (%.2f GB u8), %d MB per GPU against 126 MB of L2"
                              % (n, n * n_px * 3 / 1e9, round((len(mine) + 1) * n_px * 3 / 1e6))),
                       "parallelism": (f"frames sharded over {world} GPU(s), " + ("one fused reduce-scatter+divide kernel per rank over NVLink peer memory"
                                                                                   if use_peers else "one NCCL reduce")) if world > 1 else "1 GPU",
                       "numa_bound": bool(numa_bound), "ecc_iterations_per_step": total_iters, "ecc_iterations_per_rank": iters_per_rank, "rank_work_ms_no_exchange": rank_work_ms,
                       "step_pipelining": ("none (every step ends with a host synchronisation)" if not use_peers else
                                           "no host synchronisation inside the timed region: the exchange of step i runs on its own stream "
                                           "and overlaps the prep + ECC iterations of step i+1; that step's accumulator writes wait for it"), "wall_ms_per_step": wall_ms / a.steps},
            "whole_step": {"algorithmic_GBps_per_gpu": alg_bytes / (ms / a.steps * 1e-3) / 1e9 / world,
                           "frac_of_hbm_peak": alg_bytes / (ms / a.steps * 1e-3) / 1e9 / world / peak},
            "roofline": roof, "stages": stages, "cpu_baseline": cpu, "e2e": e2e, "e2e_api": e2e_api,
            "gpu_launches": launches, "clocks": clocks,
            "check": {"max_corner_error_vs_ground_truth_px": truth_err, "ecc_status_codes": statuses,
                      "stack_mean": stack_mean, "e2e_stack_mean": e2e_mean,
                      "oracle": oracle_check, "multi_gpu": multi_check},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if a.impl == "reference":
        run_reference(a, rank, world)
    else:
        run_b200(a, rank, local_rank, world)


if __name__ == "__main__":
    main()
