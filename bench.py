#!/usr/bin/env python
"""bench.py -- Mrays/s of the path-tracing hot path on BASELINE.json's headline workload.

Workload (BASELINE.json configs[1]): scenes/diamond_scene.json forced to 1920x1080, `path` integrator, max_depth 64,
spi 4, seed 0. One STEP = one `IRenderDevice::render()` call = one iteration = 4 samples per pixel over the whole
frame (8 294 400 camera rays and every bounce / shadow ray they spawn); 16 steps = the 64 spp of the reference's
own harness (scripts/Benchmark.py). Mrays/s = (camera + bounce + shadow rays) / time.

    python bench.py [--gpus N] [--steps K] [--warmup W]            this repository's CUDA device
    python bench.py --impl reference [...]                         CPU restatement of the reference's CPU device

For N > 1 it is launched by torchrun (one rank per GPU): the framebuffer is split into 32x32 tiles dealt along diagonals ((tx + ty) mod N)
to the ranks (scene replicated, no data-path collective), and the accumulation buffers are summed onto rank 0 with
one NCCL exchange at the end of the K steps (inside the timed region): every rank sends the pixels of its own tiles to rank 0 (ignis_b200/partition.py TileGather). Total work is fixed => "scaling": "strong".

Timed regions: `value` = K steps issued back to back with the scene resident in HBM, timed with CUDA events on the
device's render stream between barriers, max over ranks. `e2e` = the same K steps through the public host API
(B200Device.render + getFramebufferForHost over the C ABI): every step uploads its Settings and reads the whole
accumulated framebuffer back into pinned host memory (after the NCCL gather for N > 1).

The oracle (oracle/) is executed here only for `cpu_baseline` and `--impl reference`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SCENE = "scenes/diamond_scene.json"
# BASELINE.json configs; the default (headline, what the driver runs) is c2. The others are parity-test cases that can be timed the same way.
WORKLOADS = {
    "c2": dict(scene="scenes/diamond_scene.json", width=1920, height=1080, spi=4, name="diamond_scene", note="BASELINE.json configs[1]; 16 steps = 64 spp"),
    "c3": dict(scene="scenes/primitives.json", width=1920, height=1080, spi=4, name="primitives", note="BASELINE.json configs[2]; 64 steps = 256 spp"),
    "c4": dict(scene="scenes/synthetic_room.json", width=1920, height=1080, spi=4, name="synthetic_room",
               note="BASELINE.json configs[3] (4 GPUs): the Bitterli bedroom is not obtainable offline; stand-in with 1.8 M instanced triangles (tools/make_room_scene.py); 32 steps = 128 spp"),
    "c5": dict(scene="scenes/many_point_lights.json", width=3840, height=2160, spi=1, name="many_point_lights",
               note="BASELINE.json configs[4] (8 GPUs): hierarchy selector over 10 embedded point lights + sky, checkerboard, bump-mapped rough conductor; spi 1 = the reference's GPU policy at this size (Runtime.cpp:71-79); 512 steps = 512 spp"),
}
# algorithmic bytes per unit of work of the wavefront hand-off
B_PRIMARY, B_SHADOW, B_SPLAT = 216, 108, 24
# the same figure split by pipeline stage (DESIGN.md "Algorithmic bytes")
# dram__bytes_read.sum + dram__bytes_write.sum per k_turn_trace launch of this workload (ncu --set full, profiles/)
NCU_TRAFFIC_BYTES = 5.387e8
NCU_TRAFFIC_SOURCE = ("profiles/r6x_k_turn_trace_full.csv (ncu --set full of the final merged-tree kernel k_turn_trace<256,3,2,1,1>): dram read + write "
                      "of the three k_turn_trace launches of the first iteration (437 + 751 + 428 MB) / 3; later iterations also trace the carried "
                      "paths, hence the larger algorithmic figure")
# Thread instructions per traced ray of k_turn_trace on the headline workload (c2), from ncu: smsp__thread_inst_executed.sum of the four
# split-turn trace launches of one iteration (7.806 + 13.458 + 8.652 + 5.455 = 35.37 G) / the rays they traced (22.85 M primary + 10.83 M
# shadow; igb200_launch_profile) -- profiles/r5g_issue_roofline.csv. Warp instructions: 1914.9 M for the same launches.
NCU_THREAD_INST_PER_RAY = 1050.2
NCU_WARP_INST_PER_RAY = 56.85
B_STAGE = {"generate": 68, "traverse_primary": 60, "shade_read": 88, "shade_bounce_write": 68, "shade_shadow_write": 56,
           "traverse_secondary": 52}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 100 ms while the timed region runs (NVML)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _loop(self):
        nv = self._nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake", nv.nvmlClocksEventReasonApplicationsClocksSetting: "applications_clocks"}
        self._stop.wait(0.003)   # the first NVML query not on top of the first launches of the timed region
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def algorithmic_bytes(st):
    return B_PRIMARY * (st["CameraRayCount"] + st["BounceRayCount"]) + B_SHADOW * st["ShadowRayCount"] + B_SPLAT * st["Splats"]


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def build_timing_oracle():
    """The oracle compiled the way the reference compiles its CPU device (CMakeLists.txt:93-98: -O3 -march=native
    -ffast-math), on this box. Falls back to the parity build (-O2, no contraction) if that fails."""
    from oracle import oracle as o
    src = os.path.join(ROOT, "oracle", "oracle.cpp")
    # -march=native code must never travel between machines: always built on the spot, outside the tree
    import tempfile
    out = os.path.join(tempfile.gettempdir(), f"igb200_oracle_native_{os.getuid()}_{os.getpid()}.so")
    try:
        if True:
            subprocess.run(["g++", "-O3", "-march=native", "-ffast-math", "-std=c++17", "-fPIC", "-shared", "-o", out, src, "-lpthread"],
                           check=True, capture_output=True)
        o._LIB = None
        real_build = o.build
        o.build = lambda force=False: out
        try:
            o.lib()
        finally:
            o.build = real_build
        return "-O3 -march=native -ffast-math"
    except Exception:
        o._LIB = None
        o.lib()
        return "-O2 -ffp-contract=off (parity build)"


def cpu_render_steps(tables, w, h, spi, steps, first_iter=0):
    """Runs `steps` full-frame iterations of the oracle on all host threads; returns (seconds, rays)."""
    from oracle.oracle import Oracle
    o = Oracle(tables)
    fb = np.zeros((h, w, 3), np.float32)
    t0 = time.perf_counter()
    for it in range(steps):
        o.render(w, h, spi=spi, iteration=first_iter + it, fb=fb, use_bvh=2)   # SAH BVH4 + 4-triangle leaves, near child first: the reference's CPU tree
    dt = time.perf_counter() - t0
    rays = int(o.counters.sum())
    o.close()
    return dt, rays


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from ignis_b200.scene import load_scene
    from oracle import oracle as o
    flags = build_timing_oracle()
    cores = int(o.lib().igo_hardware_threads())
    tables = load_scene(os.path.join(ROOT, args.scene), args.width, args.height)
    w, h, spi = args.width, args.height, args.spi
    if args.warmup > 0:
        cpu_render_steps(tables, w, h, spi, args.warmup)
    dt, rays = cpu_render_steps(tables, w, h, spi, args.steps, first_iter=args.warmup)
    value = rays / dt / 1e6
    line = {"impl": "reference", "metric": metric_name(args), "value": value, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} full-frame iterations ({w}x{h}, spi {spi}) of the CPU restatement of the reference's CPU device "
                                       f"(oracle/oracle.cpp, scalar, 16x16 tiles, SAH BVH4 + 4-triangle leaves walked near child first as the reference's CPU device does, {cores} threads, built {flags}); the reference itself needs the AnyDSL JIT and cannot be built"},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "msamples_per_s": w * h * spi * args.steps / dt / 1e6}
    print(json.dumps(line))
    return 0


def metric_name(args):
    return f"Mrays/s (camera+bounce+shadow) @{args.width}x{args.height} {WORKLOADS[args.workload]['name']} path"


def workload_config(args, world):
    return {"workload": f"{args.scene} {args.width}x{args.height}, path integrator, spi {args.spi}, seed 0; 1 step = 1 render() iteration "
                        f"({args.spi} spp, {args.width * args.height * args.spi} camera rays); {WORKLOADS[args.workload]['note']}",
            "config_id": args.workload,
            "spi": args.spi, "width": args.width, "height": args.height,
            "parallelism": f"framebuffer tiles 32x32 dealt along diagonals over {world} GPU(s), scene replicated, one NCCL gather (send / recv inside the device library, igb200_comm_gather_framebuffer) of the ranks' tiles of the accumulation buffer onto rank 0",
            "l2": "no flush needed: every step streams its ray queues through HBM (> 800 MB per step per GPU at N=1, L2 is 126 MB)"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from ignis_b200.device import B200Device, Runtime
    from ignis_b200.partition import TILE
    from ignis_b200.scene import load_scene

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints on stdout (version banner, everything NCCL_DEBUG asks for): keep the real stdout for the one JSON line (see the end)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w, h, spi = args.width, args.height, args.spi
    tables = load_scene(os.path.join(ROOT, args.scene), w, h)
    rt = Runtime(tables, w, h, spi=spi, seed=0, cuda_device=local_rank)
    dev = rt.device
    if world > 1:
        # the exchange lives behind the C ABI (igb200_comm_*: NCCL send / recv of the ranks' tiles inside the device library);
        # torch.distributed only carries the 128-byte NCCL id to the ranks, the barriers and the statistics of this script
        ids = [B200Device.commUniqueId() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        dev.commInit(rank, world, ids[0], TILE)
    stream = torch.cuda.ExternalStream(dev.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_to_root(to_host=False):
        # disjoint tile support: the sum of the accumulation buffers is a gather of the per-rank tiles; rank 0 ends up
        # with the complete frame in a buffer of its own (its accumulation buffer keeps accumulating only its tiles)
        return dev.commGatherFramebuffer("", to_host=to_host)[1]

    # ---- warm-up (also sizes the ray queues and warms NCCL)
    for _ in range(max(args.warmup, 0)):
        rt.step()
    if world > 1:
        reduce_to_root()
    barrier()

    def timed(e2e: bool):
        rt.reset()
        dev.resetStatistics()
        rt.IterationCount = 0
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        frames = 0
        clk = ClockSampler(local_rank)   # NVML is initialised BEFORE the barrier: eight ranks doing it at once take milliseconds each, and a rank
        barrier()                        # that starts late keeps rank 0 waiting in the gather (r6f: 6.5 ms of "gather" at N = 8 were start skew)
        if world > 1:
            # ranks leave the barrier up to 2 ms apart (8 processes + their NCCL / NVML threads on 16 host cores, r6g); the exchange at the end makes
            # every rank wait for the last one, so the ranks agree on a start instant (CLOCK_MONOTONIC is shared on one box) and spin until it comes
            tt = torch.tensor([time.perf_counter()], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            go = float(tt[0]) + 0.002
            while time.perf_counter() < go:
                pass
        t0 = time.perf_counter()
        with clk:
            ev0.record(stream)
            host = None
            if e2e and args.e2e_mode == "stream":
                # every step's frame travels to pinned host memory (after the tile gather at N > 1) while later steps render; the timed
                # region ends when the LAST step's frame has arrived (igb200_frame_stream_*). All K frames are received inside it.
                for _ in range(args.steps):
                    rt.step()
                    while (f := dev.frameStreamNext(0)) is not None:
                        frames += 1
                        host = f[1]
                while (f := dev.frameStreamNext(2)) is not None:
                    frames += 1
                    host = f[1]
                if rank == 0:
                    assert frames == args.steps, (frames, args.steps)
            for _ in range(args.steps if not (e2e and args.e2e_mode == "stream") else 0):
                rt.step()
                if e2e:
                    if world > 1:
                        host = reduce_to_root(to_host=True)   # gather onto rank 0 + D2H of the complete frame there
                    else:
                        host = dev.getFramebufferForHost()   # D2H of the accumulated frame into pinned memory
            ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev_a.record(stream)      # the K iterations are issued (fused launches flush every `fuse` iterations)
            if not e2e:
                dev.sync()           # finishes the deferred tail of the last steps: all work of the K steps is inside the timed region
            ev_b.record(stream)
            if not e2e and world > 1:
                reduce_to_root()
            ev1.record(stream)
            stream.synchronize()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        dev_ms = ev0.elapsed_time(ev1)
        parts = [ev0.elapsed_time(ev_a), ev_a.elapsed_time(ev_b), ev_b.elapsed_time(ev1)]   # render launches | flush + drain | reduce
        ms = max(dev_ms, 0.0)
        st = dev.getStatistics()
        vals = torch.tensor([ms, wall_ms, st["CameraRayCount"], st["ShadowRayCount"], st["BounceRayCount"], st["Splats"], st["KernelLaunches"]],
                            dtype=torch.float64, device="cuda")
        per_rank = [parts + [ms, float(st["CameraRayCount"] + st["ShadowRayCount"] + st["BounceRayCount"]), t0 * 1e3]]   # t0: CLOCK_MONOTONIC, comparable across the ranks of one box
        if world > 1:
            mine = torch.tensor(per_rank[0], dtype=torch.float64, device="cuda")
            every = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(every, mine)
            per_rank = [[round(float(x), 3) for x in t_] for t_ in every]
            first = min(r[5] for r in per_rank)
            for r in per_rank: r[5] = round(r[5] - first, 3)
        else:
            per_rank[0][5] = 0.0
        if world > 1:
            mx = vals.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = vals.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, wall_ms = float(mx[0]), float(mx[1])
            tot = {"CameraRayCount": int(sm[2]), "ShadowRayCount": int(sm[3]), "BounceRayCount": int(sm[4]), "Splats": int(sm[5]), "KernelLaunches": int(sm[6])}
        else:
            tot = {k: st[k] for k in ("CameraRayCount", "ShadowRayCount", "BounceRayCount", "Splats", "KernelLaunches")}
        tot["TotalRays"] = tot["CameraRayCount"] + tot["ShadowRayCount"] + tot["BounceRayCount"]
        timed.per_rank = per_rank
        return ms, wall_ms, tot, clk.summary(), host

    ms, wall_ms, tot, clocks, _ = timed(e2e=False)
    per_rank = timed.per_rank
    shared_frames = False
    if args.e2e_mode == "stream":
        if world > 1 and not args.no_shared_frames:
            # the ranks of one box hand their tiles of every frame to rank 0 through shared pinned host memory (each GPU over its own PCIe link)
            # instead of gathering them on rank 0's GPU and copying K whole frames through ONE link (igb200_frame_stream_share)
            keys = [int.from_bytes(os.urandom(3), "little") + 0x1000000 if rank == 0 else None]
            dist.broadcast_object_list(keys, src=0)
            dev.frameStreamShare(keys[0])
        shared_frames = world > 1 and not args.no_shared_frames
        try:
            if shared_frames and rank == world - 1 and os.environ.get("IGB200_BENCH_FAIL_SHARED"):   # test hook: exercises the collective fall-back below
                raise RuntimeError("simulated failure to attach the shared segment")
            dev.frameStreamBegin(32 if world > 1 else 16)
            began = 1.0
        except Exception as e:   # noqa: BLE001 -- e.g. a box that refuses System V segments of this size
            began = 0.0
            print(f"[bench] rank {rank}: frame stream could not start ({e})", file=sys.stderr)
        if world > 1:
            ok = torch.tensor([began], dtype=torch.float64, device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            began_all = float(ok[0]) > 0
        else:
            began_all = began > 0
        if not began_all:
            if not shared_frames:
                raise RuntimeError("frame stream could not start")
            # every rank falls back together: frames gathered on rank 0's GPU and copied from there
            if began:
                dev.frameStreamEnd()
            dev.frameStreamShare(0)
            shared_frames = False
            dev.frameStreamBegin(32)
        for _ in range(2):                 # warm the streaming path: pinned frame buffers are allocated on first use
            rt.step()
        while dev.frameStreamNext(2) is not None:
            pass
    ms_e, wall_e, tot_e, _, host = timed(e2e=True)
    image_mean = float(np.asarray(host).mean() / args.steps) if host is not None else None
    if args.e2e_mode == "stream":
        dev.frameStreamEnd()

    # ---- per-kernel timing (separate pass, N = 1): every launch bracketed by CUDA events on the render stream
    kt = None
    if world == 1:
        rt.reset()
        dev.resetStatistics()
        dev.setOption("profile_kernels", 1)
        for _ in range(args.steps):
            rt.step()
        dev.sync()
        prof = dev.launchProfile()
        kt = dev.kernelTimes()
        kst = dev.getStatistics()
        dev.setOption("profile_kernels", 0)

    # ---- parity of the frames THIS configuration produces (same size, spi, seed, partition and exchange as the timed runs): P iterations
    # rendered again from iteration 0, brought to rank 0 exactly as in the timed region, and compared there with the oracle's parity build
    parity = None
    if args.parity_iters > 0:
        P = args.parity_iters
        rt.reset()
        dev.resetStatistics()
        rt.IterationCount = 0
        for _ in range(P):
            rt.step()
        frame = None
        if world > 1:
            frame = reduce_to_root(to_host=True)
            frame = None if frame is None else frame.copy()
        else:
            frame = dev.getFramebufferForHost().copy()
        pst = dev.getStatistics()
        cnt = torch.tensor([pst["CameraRayCount"], pst["ShadowRayCount"], pst["BounceRayCount"]], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        if rank == 0:
            from oracle.oracle import Oracle
            orc = Oracle(tables)
            ref = np.zeros((h, w, 3), np.float32)
            for it in range(P):
                orc.render(w, h, spi=spi, iteration=it, fb=ref)
            rel = float(np.linalg.norm((frame - ref).ravel()) / max(float(np.linalg.norm(ref.ravel())), 1e-30))
            parity = {"rel_l2": rel, "tolerance": 1e-4, "ok": bool(rel <= 1e-4), "iterations": P, "n_gpus": world,
                      "ray_counts": [int(x) for x in cnt.tolist()], "oracle_ray_counts": [int(x) for x in orc.counters],
                      "ray_counts_equal": [int(x) for x in cnt.tolist()] == [int(x) for x in orc.counters],
                      "against": "oracle/oracle.cpp (parity build -O2 -ffp-contract=off), same scene / size / spi / seed / iterations; at N > 1 the frame compared is the one "
                                 "gathered onto rank 0 from the ranks' tiles. diamond_scene has no reference image: the oracle's dielectric is pinned by physics tests (DESIGN.md 4)"}
            orc.close()

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        value = tot["TotalRays"] / (ms * 1e-3) / 1e6
        e2e_value = tot_e["TotalRays"] / (wall_e * 1e-3) / 1e6
        line = {"metric": metric_name(args), "value": value, "unit": "Mrays/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
                "clocks": clocks, "gpu_launches": tot["KernelLaunches"],
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 32, "d2h_bytes_per_step": w * h * 12,
                        "ms_per_step": wall_e / args.steps, "timed": "host wall clock between barriers, max over ranks",
                        "mode": (("every step's accumulated frame streamed to pinned host memory while later steps render (igb200_frame_stream_*), all K frames received inside the timed region"
                                  + ("; the frames live in shared pinned host memory that every rank writes its own tiles into (igb200_frame_stream_share), rank 0 takes each "
                                     "frame once all ranks have flagged it" if shared_frames else ("; every frame gathered onto rank 0's GPU (NCCL) and copied from there" if world > 1 else "")))
                                 if args.e2e_mode == "stream" else "render() + synchronous getFramebufferForHost every step")},
                "per_rank_ms": {"columns": ["issued launches", "flush + drain of the deferred tail", "NCCL gather of the tiles (on rank 0 incl. waiting for the last rank)", "total", "rays traced", "host start after the first rank (ms)"], "ranks": per_rank},
                "rays": tot, "msamples_per_s": w * h * spi * args.steps / (ms * 1e-3) / 1e6, "wall_ms_per_step": wall_ms / args.steps}
        step_bytes = algorithmic_bytes(tot)
        line["roofline_step"] = {"bound": "hbm", "achieved": step_bytes / (ms * 1e-3) / 1e9 / world, "peak": peak, "unit": "GB/s",
                                 "frac": step_bytes / (ms * 1e-3) / 1e9 / world / peak, "traffic": None,
                                 "what": f"whole wavefront step per GPU: {B_PRIMARY} B x primary + {B_SHADOW} B x shadow + {B_SPLAT} B x splat (SURVEY.md 8d), peak {peak_src}"}
        line["roofline"] = dict(line["roofline_step"], kernel="whole step, all kernels (per-kernel figures are measured at N = 1)")
        if kt is not None:
            # Dominant kernel of the step: k_turn_trace<256,3,2,1>, the trace phase of the split turns. Algorithmic bytes of a launch
            # (trace share): 40 B ray read + 20 B hit write per primary ray, 52 B per shadow ray, 24 B per splat.
            kern = prof["kernels"]
            total_ms = sum(v["ms"] for v in kern.values()) or 1.0
            w_ = prof["k_turn_trace_work"]
            n_l = max(kern["k_turn_trace"]["launches"], 1)
            avg_ms = kern["k_turn_trace"]["ms"] / n_l
            kbytes = (B_STAGE["traverse_primary"] * w_["primary"] + B_STAGE["traverse_secondary"] * w_["shadow"] + B_SPLAT * w_["splats"]) / n_l
            ach = kbytes / (avg_ms * 1e-3) / 1e9
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": NCU_TRAFFIC_BYTES,
                                "kernel": "k_turn_trace<256,3,2,1,1> (merged single-level tree staged in shared memory; closest-hit + any-hit/splat phase of a split wavefront turn)",
                                "avg_launch_ms": avg_ms, "launches": n_l, "share_of_step": kern["k_turn_trace"]["ms"] / total_ms,
                                "algorithmic_bytes_per_launch": kbytes, "rays_per_launch": (w_["primary"] + w_["shadow"]) / n_l,
                                "peak_source": peak_src, "traffic_source": NCU_TRAFFIC_SOURCE,
                                "timing": "CUDA events around every launch on the render stream, separate pass of the same K steps"}
            if args.workload == "c2" and w * h * spi == 1920 * 1080 * 4:
                # The ceiling this kernel actually runs against (DESIGN.md 3 "Roofline"): the scene sits in shared memory, so the trace phase is bound
                # by instruction issue, not by HBM. Peak = SMs x 4 schedulers x 32 lanes x clock (thread instructions / s); achieved = the ncu-measured
                # thread instructions per ray x the rays this run traced per second in this kernel. simt = lanes active per issued instruction.
                clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
                n_sm = torch.cuda.get_device_properties(0).multi_processor_count
                rays_s = (w_["primary"] + w_["shadow"]) / (kern["k_turn_trace"]["ms"] * 1e-3)
                peak_t = n_sm * 4 * 32 * clk
                line["issue_roofline"] = {"bound": "issue", "kernel": "k_turn_trace", "unit": "T thread-inst/s", "achieved": NCU_THREAD_INST_PER_RAY * rays_s / 1e12, "peak": peak_t / 1e12,
                                          "frac": NCU_THREAD_INST_PER_RAY * rays_s / peak_t, "thread_inst_per_ray": NCU_THREAD_INST_PER_RAY, "warp_inst_per_ray": NCU_WARP_INST_PER_RAY,
                                          "simt_lanes_per_inst": NCU_THREAD_INST_PER_RAY / NCU_WARP_INST_PER_RAY, "warp_issue_frac": NCU_WARP_INST_PER_RAY * rays_s / (n_sm * 4 * clk),
                                          "grays_per_s": rays_s / 1e9, "grays_per_s_at_full_simt_and_issue": peak_t / NCU_THREAD_INST_PER_RAY / 1e9,
                                          "source": "profiles/r5g_issue_roofline.csv (ncu smsp__thread_inst_executed.sum / smsp__inst_executed.sum of the split-turn trace launches of one iteration) x live ray rate"}
            line["kernels"] = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps, "share": v["ms"] / total_ms} for k, v in kern.items()}
            prim = kst["CameraRayCount"] + kst["BounceRayCount"]
            phase_bytes = {"trace": B_STAGE["traverse_primary"] * prim + B_STAGE["traverse_secondary"] * kst["ShadowRayCount"] + B_SPLAT * kst["Splats"],
                           "shade_generate": B_STAGE["generate"] * kst["CameraRayCount"] + B_STAGE["shade_read"] * prim
                                             + B_STAGE["shade_bounce_write"] * kst["BounceRayCount"] + B_STAGE["shade_shadow_write"] * kst["ShadowRayCount"]}
            total_k = sum(v["ms"] for v in kt.values()) or 1.0
            line["phase_ms"] = {k: {"ms_per_step": v["ms"] / args.steps, "share": v["ms"] / total_k,
                                    "algorithmic_GBps": phase_bytes[k] / max(v["ms"], 1e-9) / 1e6} for k, v in kt.items()}
        if image_mean is not None:
            line["image_mean"] = image_mean
        line["parity"] = parity

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as o
        flags = build_timing_oracle()
        cores = int(o.lib().igo_hardware_threads())
        n_it = args.cpu_iters
        dt, rays = cpu_render_steps(tables, w, h, spi, n_it)
        line["cpu_baseline"] = {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                "sample": f"{n_it} full-frame iteration(s) of the same workload ({w}x{h}, spi {spi}; {rays} rays, {dt:.1f} s) on {cores} threads, "
                                          f"CPU restatement of the reference's CPU device (BVH4+Tri4 ordered: SAH BVH4, 4-triangle leaves, near child first) built {flags}"}
    if rank == 0:
        if world > 1:
            # NCCL (with NCCL_DEBUG set) writes to fd 1 whenever it likes, also at exit: fd 1 stays pointed at stderr for the whole run
            # and the ONE JSON line goes to the real stdout
            os.write(saved_stdout, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line), flush=True)
    # tear down in dependency order: tensors that alias or were used on the device's stream go first, then the device
    # (which owns that stream), then NCCL; otherwise the allocator records events on a stream that no longer exists at exit
    torch.cuda.synchronize()
    del stream
    torch.cuda.synchronize()
    rt.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-shared-frames", action="store_true", help="e2e at N > 1: gather every frame on rank 0's GPU and copy it from there (the round-2 default before r6)")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="BASELINE.json config (default: the headline one, c2)")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--spi", type=int, default=0)
    ap.add_argument("--cpu-iters", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-mode", default="stream", choices=["stream", "sync"], help="end-to-end leg: streamed frames (default) or a synchronous read-back every step")
    ap.add_argument("--parity-iters", type=int, default=2, help="iterations of the parity check against the oracle (0: skip)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    args.scene = wl["scene"]
    args.width, args.height, args.spi = args.width or wl["width"], args.height or wl["height"], args.spi or wl["spi"]
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
