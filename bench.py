#!/usr/bin/env python
"""Headline benchmark: Mrays/s of the Intersector::trace() / trace_probe() hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic rays of the named workload
(BASELINE.json configs):

  c1: the reference's own CPU-runnable case -- Cornell box (32 triangles), 512 x 512 pinhole primaries
      (+ one ambient-occlusion probe per primary hit, extra figure); tiny: launch-latency bound on a GPU
  c2 (default, the 1-GPU configuration):  999 698-triangle displaced grid, 16 Mi coherent pinhole
      primaries + 16 Mi incoherent cosine-weighted bounce rays, closest hit
  c3: 9 999 392-triangle fBm terrain x 64 assembly instances, 32 Mi incoherent closest-hit rays
      (+ shadow probes reported as an extra figure)
  c4: 2 000 000 moving triangles (msc = 1), 16 Mi incoherent closest-hit rays with random times
  c5: the C3 scene, 1920 x 1080 x 64 spp synthetic path stream (camera ray + 3 cosine bounces + one
      shadow probe per vertex, ~1 G rays per frame) through the wavefront queues (asgpu_path_stream_*),
      32 x 32 tiles dealt to the ranks in Hilbert order: a step is one frame, scaling is STRONG

`value`  = rays / device time with rays already resident in HBM (CUDA events on the launch stream).
`e2e`    = the same step through the host-buffer C ABI call (asgpu_trace_host) with pinned host
           buffers: H2D of the rays and D2H of the hit records inside the timed region.
`roofline` = algorithmic bytes (measured node / triangle / instance visits per ray x record sizes
           + ray in + hit out) / kernel time, against the measured HBM copy bandwidth.
`cpu_baseline` = the reference's CPU path (oracle/_ref when present, else the oracle port) on all
           host cores over a bounded sample of the same rays.

With --gpus N > 1 (launched by torchrun, one rank per GPU) rank 0 builds and flattens the scene,
broadcasts the blob once over NCCL, and every rank traces its own shard of rays (weak scaling, no
data-path collective); time = max over ranks.

--impl reference times the CPU reference path alone (rank 0 only) and prints the same line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from appleseed_b200 import scenes  # noqa: E402
from appleseed_b200.scene import HIT_DTYPE, VIS_DIFFUSE, VIS_SHADOW, RayBatch  # noqa: E402

METRIC = "Mrays/s closest-hit (wavefront batches, rays resident in HBM)"
# Bytes the kernel fetches when a ray enters an assembly instance: 112 of the 128-byte ItemRecord
# (3 x 4 matrix + tree / visibility / id), 32 bytes of TreeDesc fields (node, triangle and pose
# offsets, node and slice counts) and the 4-byte item index of the wide top-level leaf.
INSTANCE_BYTES = 112 + 32 + 4
UNIT = "Mrays/s"
MI = 1 << 20


# ---------------------------------------------------------------------------------------------
# Workloads
# ---------------------------------------------------------------------------------------------

def workload_name(args) -> str:
    return {
        "c1": "C1: Cornell box (32 triangles), %d pinhole primary rays, closest hit (+ AO probes, extra)%.0s",
        "c2": "C2: 999698-triangle displaced grid, %d coherent primary + %d incoherent cosine bounce rays, closest hit",
        "c3": "C3: 9999392-triangle fBm terrain x 64 assembly instances, %d incoherent closest-hit rays (+ %d shadow probes, extra)",
        "c4": "C4: 2000000 moving triangles (msc=1), %d incoherent closest-hit rays with random time (+ %d probes, extra)",
        "c5": "C5: %dx%dx%d spp path stream (camera + 3 cosine bounces + 1 shadow probe per vertex) on the 9999392-triangle x 64-instance scene, wavefront queues, 32x32 tiles",
    }[args.workload] % ((args.width, args.height, args.spp) if args.workload == "c5" else (args.rays, args.rays))


def make_scene(args):
    if args.workload == "c1":
        return scenes.scene_c1()
    if args.workload == "c2":
        return scenes.scene_c2(args.res or 707)
    if args.workload in ("c3", "c5"):
        return scenes.scene_c3(args.res or 2236, 8)
    if args.workload == "c4":
        return scenes.scene_c4(args.res or 1000, 1)
    raise ValueError(args.workload)


def primary_rays_c2(n: int, rank: int) -> RayBatch:
    side = int(round(math.sqrt(n)))
    # Each rank looks at the terrain from its own camera position (its own image "tile").
    ang = 0.37 * rank
    eye = (0.35 * math.cos(ang), 3.0, 0.35 * math.sin(ang) + 0.001)
    return scenes.pinhole_rays(side, side, eye, (0.0, 0.0, 0.0), up=(0, 0, 1), film=0.025, focal=0.05)


def incoherent_rays(desc, n: int, seed: int, time: bool = False) -> RayBatch:
    """Origins uniform in the scene's bounding box (2 % margin), directions uniform on the sphere
    (SURVEY.md section 8(d), C3 / C4)."""
    lo, hi = scenes.scene_bbox(desc)
    ext = hi - lo
    return scenes.uniform_sphere_rays(n, lo - 0.02 * ext, hi + 0.02 * ext, seed, time=time)


# ---------------------------------------------------------------------------------------------
# Clocks sampling
# ---------------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU DURING the timed region: NVML polled
    from a thread every few milliseconds (nvidia-smi -lms as a fallback)."""
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int, period_s: float = 0.004):
        self.index = index
        self.period = period_s
        self.samples = []           # (sm_mhz, power_w, reasons bitmask)
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.proc = None
        self.lines = []
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: resolve through the PCI bus id of the CUDA device.
            import torch
            bus = getattr(torch.cuda.get_device_properties(self.index), "pci_bus_id", None)
            handle = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        handle = h
                        break
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml, self.handle = pynvml, handle
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                reasons = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                try:
                    power = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                except Exception:
                    power = None
                self.samples.append((float(mhz), power, int(reasons)))
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            n = self.nvml
            names = [("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap)]
            reasons = sorted({name for _, _, r in self.samples for name, bit in names if r & bit})
            sm = [x[0] for x in self.samples]
            pw = [x[1] for x in self.samples if x[1] is not None]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "power_w_max": max(pw) if pw else None, "reasons": reasons, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for k, name in enumerate(names):
                if p[3 + k].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


# ---------------------------------------------------------------------------------------------
# CPU reference arm
# ---------------------------------------------------------------------------------------------

def cpu_oracle():
    from oracle import oracle as orc
    try:
        orc.build("all")
    except Exception:
        pass
    if orc.available("asref"):
        return orc.Oracle("asref"), "reference"
    return orc.Oracle("orc"), "port"


def time_cpu(oscene, rays: RayBatch, probe: bool, threads: int, repeats: int = 1) -> float:
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        (oscene.trace_probe if probe else oscene.trace)(rays, threads=threads)
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_sample(args, desc, rank: int = 0):
    """The bounded sample of the workload's rays the CPU arm is timed on."""
    n = args.cpu_rays if args.cpu_rays > 0 else args.rays
    if args.workload == "c1":
        return scenes.rays_c1_primary(), 0
    if args.workload == "c2":
        half = n // 2
        prim = primary_rays_c2(args.rays // 2, rank)
        stride = max(1, len(prim) // half)
        coherent = prim.take(np.arange(0, len(prim), stride)[:half])
        # Bounce rays need hit points: the CPU arm bounces from the sampled primaries it traces itself.
        return coherent, half
    return incoherent_rays(desc, n, 1000 + rank, time=(args.workload == "c4")), 0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    desc = make_scene(args)
    oracle, kind = cpu_oracle()
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    oscene = oracle.scene(desc)
    build_s = time.perf_counter() - t0
    if args.workload == "c5":
        cfg = c5_config(args, desc)
        w, h = max(32, args.width // 2), max(32, args.height // 2)
        cpu_path_stream(desc, oscene, cfg, w // 4, h // 4, threads)
        ns, ts = [], []
        for k in range(args.steps):
            n, secs = cpu_path_stream(desc, oscene, cfg, w, h, threads, seed=7 + k)
            ns.append(n); ts.append(secs)
        ms = 1e3 * sum(ts) / len(ts)
        value = sum(ns) / sum(ts) / 1e6
        print(json.dumps({
            "impl": "reference", "metric": "Mrays/s closest-hit + shadow-probe (wavefront path stream, queues resident in HBM)",
            "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": 1, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "sample": "%dx%d x 1 spp of the path stream per step (%d rays)" % (w, h, ns[0]),
                       "scene_build_s": round(build_s, 2)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": "%dx%d x 1 spp of the path stream per step, %d threads" % (w, h, threads)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    sample, half = cpu_sample(args, desc)
    if args.workload == "c2":
        hits = oscene.trace(sample, threads=threads)
        mask, pts, nrm = scenes.hit_points_and_normals(desc, sample, hits)
        bounce = scenes.bounce_rays(pts, nrm, 1, flags=VIS_DIFFUSE)
        batches = [sample, bounce]
    else:
        batches = [sample]
    n_rays = sum(len(b) for b in batches)
    for _ in range(max(1, min(args.warmup, 1))):
        for b in batches:
            oscene.trace(b, threads=threads)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        for b in batches:
            oscene.trace(b, threads=threads)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = n_rays / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": "%d rays per step (bounded sample of the workload)" % n_rays,
                   "scene_build_s": round(build_s, 2)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "%d rays of the workload per step, %d threads" % (n_rays, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------

def pinned(a: np.ndarray):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a) if a.dtype != np.uint32 else np.ascontiguousarray(a).view(np.int32))
    return t.pin_memory()


class PinnedRays:
    """Host ray batch in pinned memory + its device copy."""

    def __init__(self, rays: RayBatch, device):
        from appleseed_b200.intersector import DeviceRays
        self.n = len(rays)
        f = lambda a: None if a is None else pinned(a)
        self.host_t = [f(rays.org), f(rays.dir), f(rays.tmin), f(rays.tmax), f(rays.time_absolute), f(rays.time_normalized), f(rays.flags)]
        np_of = lambda t, dt: None if t is None else t.numpy().view(dt)
        self.host = RayBatch(np_of(self.host_t[0], np.float64), np_of(self.host_t[1], np.float64), np_of(self.host_t[2], np.float64),
                             np_of(self.host_t[3], np.float64), np_of(self.host_t[4], np.float32), np_of(self.host_t[5], np.float32),
                             np_of(self.host_t[6], np.uint32))
        dev = [None if t is None else t.to(device, non_blocking=True) for t in self.host_t]
        self.dev = DeviceRays(*dev)
        self.bytes_in = self.n * self.host.bytes_per_ray


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from appleseed_b200.intersector import HIT_BYTES, HostTrees, Intersector, TraceContext, hits_from_tensor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- scene: built once on rank 0, replicated with ONE broadcast --------------------------
    desc = make_scene(args)
    t0 = time.perf_counter()
    build_s = flatten_s = bcast_s = 0.0
    if rank == 0:
        trees = HostTrees(desc, threads=0, build_device=local_rank if args.tree_build == "device" else None)
        build_s = trees.build_seconds
        t1 = time.perf_counter()
        ctx = TraceContext(device=local_rank, trees=trees)
        trees.close()
        flatten_s = time.perf_counter() - t1
    if world > 1:
        size = torch.tensor([ctx.blob_size if rank == 0 else 0], dtype=torch.int64, device=device)
        dist.broadcast(size, 0)
        blob = ctx.blob_tensor() if rank == 0 else torch.empty(int(size.item()), dtype=torch.uint8, device=device)
        barrier()
        t1 = time.perf_counter()
        dist.broadcast(blob, 0)
        torch.cuda.synchronize()
        bcast_s = time.perf_counter() - t1
        if rank != 0:
            ctx = TraceContext.from_blob(blob, adopt=True)
    isect = Intersector(ctx)
    info = ctx.info()

    # ---- rays (this rank's shard) ---------------------------------------------------------------
    n = args.rays
    batches = []          # (label, PinnedRays)
    hits_dev = torch.empty(n * HIT_BYTES, dtype=torch.uint8, device=device)
    if args.workload == "c2":
        prim = primary_rays_c2(n, rank)
        n = len(prim)
        p = PinnedRays(prim, device)
        isect.trace_device(p.dev, hits_dev)
        torch.cuda.synchronize()
        hits = hits_from_tensor(hits_dev, n)
        mask, pts, nrm = scenes.hit_points_and_normals(desc, prim, hits)
        if len(pts) < n:            # pad the bounce wavefront to n rays by wrapping around
            idx = np.resize(np.arange(len(pts)), n)
            pts, nrm = pts[idx], nrm[idx]
        bounce = scenes.bounce_rays(pts, nrm, 1 + rank, flags=VIS_DIFFUSE)
        batches = [("coherent_primary", p), ("incoherent_bounce", PinnedRays(bounce, device))]
        probe_batch = None
    elif args.workload == "c1":
        prim = scenes.rays_c1_primary()
        n = len(prim)
        hits_dev = torch.empty(n * HIT_BYTES, dtype=torch.uint8, device=device)
        p = PinnedRays(prim, device)
        isect.trace_device(p.dev, hits_dev)
        torch.cuda.synchronize()
        hits = hits_from_tensor(hits_dev, n)
        _, ao = scenes.rays_c1_ao(desc, prim, hits)
        batches = [("primary", p)]
        probe_batch = PinnedRays(ao, device)
    else:
        inc = incoherent_rays(desc, n, 2 + rank, time=(args.workload == "c4"))
        p = PinnedRays(inc, device)
        batches = [("incoherent", p)]
        isect.trace_device(p.dev, hits_dev)
        torch.cuda.synchronize()
        hits = hits_from_tensor(hits_dev, n)
        hit = hits["prim_type"] == 2
        pts = inc.org + np.where(hit, hits["t"], 0.0)[:, None] * inc.dir
        pts = pts - 1e-6 * inc.dir
        lo, hi = scenes.scene_bbox(desc)
        lights = np.array([[lo[0], hi[1] + 2.0, lo[2]], [hi[0], hi[1] + 2.0, lo[2]], [lo[0], hi[1] + 2.0, hi[2]], [hi[0], hi[1] + 2.0, hi[2]]])
        sh = scenes.shadow_rays(pts, lights, 3 + rank, flags=VIS_SHADOW)
        if args.workload == "c4":
            sh.time_absolute, sh.time_normalized = inc.time_absolute, inc.time_normalized
        probe_batch = PinnedRays(sh, device)
    del hits
    rays_per_step = sum(b.n for _, b in batches)
    hit_bufs = [torch.empty(b.n * HIT_BYTES, dtype=torch.uint8, device=device) for _, b in batches]
    torch.cuda.synchronize()

    def step_device():
        for (_, b), out in zip(batches, hit_bufs):
            isect.trace_device(b.dev, out)

    # ---- timed region: device-resident rays ----------------------------------------------------
    for _ in range(max(3, args.warmup)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    per_batch_ms = [0.0] * len(batches)
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in batches] for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    for s in range(args.steps):
        for k, ((_, b), out) in enumerate(zip(batches, hit_bufs)):
            ev[s][k][0].record()
            isect.trace_device(b.dev, out)
            ev[s][k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    for s in range(args.steps):
        for k in range(len(batches)):
            per_batch_ms[k] += ev[s][k][0].elapsed_time(ev[s][k][1]) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms_step = sum(per_batch_ms)
    launches = args.steps * len(batches)

    # ---- probes (extra figure) -------------------------------------------------------------------
    probe_ms = None
    if probe_batch is not None:
        occ = torch.empty(probe_batch.n, dtype=torch.uint8, device=device)
        for _ in range(3):
            isect.trace_probe_device(probe_batch.dev, occ)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            isect.trace_probe_device(probe_batch.dev, occ)
        e1.record()
        torch.cuda.synchronize()
        probe_ms = e0.elapsed_time(e1) / args.steps

    # ---- end to end through the host-buffer ABI call (pinned host memory) ---------------------
    host_hits = [torch.empty(b.n * HIT_BYTES, dtype=torch.uint8).pin_memory() for _, b in batches]
    host_hits_np = [h.numpy().view(HIT_DTYPE) for h in host_hits]

    def step_host():
        for (_, b), out in zip(batches, host_hits_np):
            isect.trace(b.host, out=out)

    step_host()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        step_host()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    h2d = sum(b.bytes_in for _, b in batches)
    d2h = sum(b.n * HIT_BYTES for _, b in batches)
    # The e2e results must be the device results.
    check = hits_from_tensor(hit_bufs[-1], batches[-1][1].n)
    assert check.tobytes() == host_hits_np[-1].tobytes(), "host-path results differ from device-path results"

    # ---- algorithmic bytes per ray (counters variant, untimed) ----------------------------------
    ctx.counters(reset=True)
    for (_, b), out in zip(batches, hit_bufs):
        isect.trace_device(b.dev, out, counters=True)
    c = ctx.counters(reset=True)
    r = max(1, c["rays"])
    per_ray = {
        "top_nodes": c["assembly_nodes_visited"] / r, "instances": c["instances_visited"] / r,
        "nodes": c["triangle_nodes_visited"] / r, "triangles": c["triangles_tested"] / r, "hit_rate": c["hits"] / r,
    }
    ray_in = h2d / rays_per_step
    bytes_per_ray = (per_ray["top_nodes"] + per_ray["nodes"]) * 80 + per_ray["triangles"] * 48 + per_ray["instances"] * INSTANCE_BYTES \
        + ray_in + HIT_BYTES
    if args.workload == "c4":
        bytes_per_ray += per_ray["triangles"] * 72       # two 36-byte poses per moving triangle

    # ---- aggregate over ranks (max time) --------------------------------------------------------
    t = torch.tensor([ms_step, e2e_ms, probe_ms or 0.0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, e2e_ms, probe_ms_max = (float(x) for x in t.cpu())
    total_rays = rays_per_step * world
    value = total_rays / (ms_step * 1e-3) / 1e6
    e2e_value = total_rays / (e2e_ms * 1e-3) / 1e6

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = bytes_per_ray * rays_per_step / (ms_step * 1e-3) / 1e9      # per GPU
        # DRAM traffic per launch: measured bytes per ray of the committed ncu capture of this
        # workload's launches (profiles/traffic.json, written by tools/ncu_traffic.py) x this run's
        # rays per launch, averaged over the step's launches; null when no capture is committed.
        traffic = None
        traffic_src = None
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                entry = json.load(open(prof)).get(args.workload)
                per_launch = [entry["launches"][label]["dram_bytes_per_ray"] * b.n for label, b in batches]
                traffic = sum(per_launch) / len(per_launch)
                traffic_src = "ncu --set full capture %s (dram__bytes_read+write per ray x rays per launch)" % entry["source"]
            except Exception:
                traffic = None
        algorithmic_per_launch = bytes_per_ray * rays_per_step / len(batches)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(args), "rays_per_step_per_gpu": rays_per_step,
                "batches": {label: {"rays": b.n, "ms": round(ms, 4), "mrays_s": round(b.n / ms / 1e3, 1)}
                            for (label, b), ms in zip(batches, per_batch_ms)},
                "kernel": "wide_kernel (8-wide quantised BVH, fp32 interval box tests, warp-cooperative exact fp64 triangle tests)",
                "l2": "inputs larger than L2 (%.0f MB of rays + %.0f MB scene blob per step vs 126 MB L2)" % (h2d / 1e6, info["blob_bytes"] / 1e6),
                "scene": {k: info[k] for k in ("triangle_count", "instance_count", "wide_node_count", "binary_node_count", "blob_bytes")},
                "scene_build_s": round(build_s, 2), "tree_build": args.tree_build, "flatten_upload_s": round(flatten_s, 2), "broadcast_s": round(bcast_s, 4),
                "per_ray": {k: round(v, 3) for k, v in per_ray.items()},
                "parallelism": "rays sharded by rank, scene replicated by one NCCL broadcast" if world > 1 else "1 GPU",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_launch": algorithmic_per_launch, "bytes_per_ray": round(bytes_per_ray, 1),
                         "peak_source": peak_src, "traffic_source": traffic_src,
                         "kernel": "wide_kernel<closest>",
                         "note": "algorithmic bytes are mostly served by L1/L2 (scene smaller than or comparable to the 126 MB L2): "
                                 "the kernel is instruction-issue bound, see profiles/README.md"},
            "clocks": clocks,
        }
        if probe_batch is not None:
            line["config"]["shadow_probe"] = {"rays": probe_batch.n, "ms": round(probe_ms_max, 4),
                                               "mrays_s": round(probe_batch.n * world / probe_ms_max / 1e3, 1)}
        if world == 1 and not args.no_cpu:
            oracle, kind = cpu_oracle()
            threads = os.cpu_count() or 1
            oscene = oracle.scene(desc)
            cpu_n = min(args.cpu_rays if args.cpu_rays > 0 else batches[-1][1].n, batches[-1][1].n)
            sample = batches[-1][1].host.slice(0, cpu_n)
            secs = time_cpu(oscene, sample, False, threads)
            cpu_hits = oscene.trace(sample.slice(0, min(cpu_n, 200000)), threads=threads)
            gpu_hits = host_hits_np[-1][: len(cpu_hits)]
            agree = float((cpu_hits["tri_slot"] == gpu_hits["tri_slot"]).mean())
            line["cpu_baseline"] = {"value": cpu_n / secs / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": "first %d rays of the '%s' batch, %d threads, %.1f s" % (cpu_n, batches[-1][0], threads, secs),
                                    "identity_agreement_with_gpu": agree}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()

# ---------------------------------------------------------------------------------------------
# C5: the wavefront path stream
# ---------------------------------------------------------------------------------------------

def c5_config(args, desc) -> dict:
    """Camera above one corner of the instance lattice looking at its centre; 4 point lights above
    the corners (the C3 lights)."""
    from appleseed_b200.wavefront import look_at
    lo, hi = scenes.scene_bbox(desc)
    centre = 0.5 * (lo + hi)
    diag = float(np.linalg.norm(hi - lo))
    eye = centre + np.array([0.32, 0.55, 0.45]) * diag
    lights = np.array([[lo[0], hi[1] + 2.0, lo[2]], [hi[0], hi[1] + 2.0, lo[2]], [lo[0], hi[1] + 2.0, hi[2]], [hi[0], hi[1] + 2.0, hi[2]]])
    return dict(width=args.width, height=args.height, spp=args.spp, camera_to_world=look_at(eye, centre), lights=lights,
                max_bounces=3, tile_size=32, seed=5, offset_eps=1.0e-6 * diag, parents=not getattr(args, "no_parents", False))


def cpu_path_stream(desc, oscene, cfg: dict, width: int, height: int, threads: int, seed: int = 7):
    """The same path stream restated on the host for the CPU arm: rays generated with numpy (same
    camera, same sampling distributions), traced by the reference's CPU path.  Returns (rays traced,
    seconds spent in the trace calls)."""
    cam = scenes.camera_rays(width, height, cfg["camera_to_world"], (0.025, 0.025 * height / width), 0.035)
    rays, total, secs = cam, 0, 0.0
    for depth in range(cfg["max_bounces"] + 1):
        t0 = time.perf_counter()
        hits = oscene.trace(rays, threads=threads)
        secs += time.perf_counter() - t0
        total += len(rays)
        mask, pts, nrm = scenes.hit_points_and_normals(desc, rays, hits)
        if len(pts) == 0:
            break
        org = pts + cfg["offset_eps"] * nrm
        sh = scenes.shadow_rays(org, np.asarray(cfg["lights"]), seed + 100 + depth)
        t0 = time.perf_counter()
        oscene.trace_probe(sh, threads=threads)
        secs += time.perf_counter() - t0
        total += len(sh)
        rays = scenes.bounce_rays(pts, nrm, seed + depth, flags=VIS_DIFFUSE, offset=cfg["offset_eps"])
    return total, secs


def run_gpu_c5(args):
    import torch
    import torch.distributed as dist
    from appleseed_b200.distributed import tile_ids_shard
    from appleseed_b200.intersector import HostTrees, TraceContext
    from appleseed_b200.wavefront import PathStream, PathStreamConfig

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    desc = make_scene(args)
    build_s = flatten_s = bcast_s = 0.0
    if rank == 0:
        trees = HostTrees(desc, threads=0, build_device=local_rank if args.tree_build == "device" else None)
        build_s = trees.build_seconds
        t1 = time.perf_counter()
        ctx = TraceContext(device=local_rank, trees=trees)
        trees.close()
        flatten_s = time.perf_counter() - t1
    if world > 1:
        size = torch.tensor([ctx.blob_size if rank == 0 else 0], dtype=torch.int64, device=device)
        dist.broadcast(size, 0)
        blob = ctx.blob_tensor() if rank == 0 else torch.empty(int(size.item()), dtype=torch.uint8, device=device)
        barrier()
        t1 = time.perf_counter()
        dist.broadcast(blob, 0)
        torch.cuda.synchronize()
        bcast_s = time.perf_counter() - t1
        if rank != 0:
            ctx = TraceContext.from_blob(blob, adopt=True)
    info = ctx.info()

    cfg = c5_config(args, desc)
    tiles = tile_ids_shard(args.width, args.height, world, rank, 32)
    ps = PathStream(ctx, PathStreamConfig(**cfg), queue_capacity=args.rays)

    def frame():
        ps.render(tiles)

    for _ in range(max(3, args.warmup) if args.spp <= 8 else 1):
        frame()
    torch.cuda.synchronize()
    ps.clear()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):
        ev[s][0].record()
        frame()
        ev[s][1].record()
    barrier()
    ms_step = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    st = ps.stats()
    rays_step = (st["camera_rays"] + st["bounce_rays"] + st["probe_rays"]) // args.steps
    closest_step = (st["camera_rays"] + st["bounce_rays"]) // args.steps
    launches = st["kernel_launches"]

    # ---- end to end: clear, tile list in, image out, through the public API (host buffers) ----
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        ps.clear()
        frame()
        image = ps.image()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    # One frame's accumulators of this rank's pixels; summed over ranks it does not depend on N.
    checksum = int(image.astype(np.uint64).sum())

    # ---- algorithmic bytes per ray (counters variant of the same frame, untimed) --------------
    ps.close()
    psc = PathStream(ctx, PathStreamConfig(**{**cfg, "counters": True}), queue_capacity=args.rays)
    ctx.counters(reset=True)
    psc.render(tiles[: max(1, len(tiles) // 8)])
    c = ctx.counters(reset=True)
    psc.close()
    r = max(1, c["rays"])
    per_ray = {"top_nodes": c["assembly_nodes_visited"] / r, "instances": c["instances_visited"] / r,
               "nodes": c["triangle_nodes_visited"] / r, "triangles": c["triangles_tested"] / r, "hit_rate": c["hits"] / r}
    closest_frac = closest_step / max(1, rays_step)
    bytes_per_ray = (per_ray["top_nodes"] + per_ray["nodes"]) * 80 + per_ray["triangles"] * 48 + per_ray["instances"] * INSTANCE_BYTES \
        + 72 + closest_frac * 40 + (1 - closest_frac) * 1

    t = torch.tensor([ms_step, e2e_ms], dtype=torch.float64, device=device)
    counts = torch.tensor([rays_step, closest_step, checksum, launches], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    ms_step, e2e_ms = (float(x) for x in t.cpu())
    total_rays, total_closest, checksum, launches = (int(x) for x in counts.cpu())
    value = total_rays / (ms_step * 1e-3) / 1e6
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = bytes_per_ray * (total_rays / world) / (ms_step * 1e-3) / 1e9
        line = {
            "metric": "Mrays/s closest-hit + shadow-probe (wavefront path stream, queues resident in HBM)", "value": value, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": 3 if args.spp <= 8 else 1,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(args), "rays_per_frame": total_rays, "closest_rays_per_frame": total_closest,
                "probe_rays_per_frame": total_rays - total_closest, "tiles_per_gpu": int(len(tiles)), "queue_capacity": args.rays,
                "next_ray_origin": "parent shading point, refined + offset on the device (ShadingPoint::refine_and_offset)" if cfg["parents"] else "hit point + eps * normal",
                "image_checksum": checksum,
                "l2": "inputs larger than L2 (%.0f MB scene blob, %.0f MB of queued rays per wavefront vs 126 MB L2)" % (info["blob_bytes"] / 1e6, args.rays * 72 / 1e6),
                "scene": {k: info[k] for k in ("triangle_count", "instance_count", "wide_node_count", "binary_node_count", "blob_bytes")},
                "scene_build_s": round(build_s, 2), "tree_build": args.tree_build, "flatten_upload_s": round(flatten_s, 2), "broadcast_s": round(bcast_s, 4),
                "per_ray": {k: round(v, 3) for k, v in per_ray.items()},
                "parallelism": "tiles dealt to ranks in Hilbert order, scene replicated by one NCCL broadcast" if world > 1 else "1 GPU",
            },
            "e2e": {"value": total_rays / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(len(tiles)) * 4 * world,
                    "d2h_bytes_per_step": args.width * args.height * 16 * world, "ms_per_step": e2e_ms},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "bytes_per_ray": round(bytes_per_ray, 1), "peak_source": peak_src, "kernel": "wide_kernel (closest + any hit)",
                         "note": "whole-frame figure: trace kernels plus the generate / shade / accumulate stages"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            oracle, kind = cpu_oracle()
            threads = os.cpu_count() or 1
            oscene = oracle.scene(desc)
            w = max(32, args.width // 4)
            h = max(32, args.height // 4)
            n, secs = cpu_path_stream(desc, oscene, cfg, w, h, threads)
            line["cpu_baseline"] = {"value": n / secs / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": "%dx%d x 1 spp of the same path stream (%d rays), %d threads, %.1f s" % (w, h, n, threads, secs)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--width", type=int, default=1920, help="c5: image width")
    ap.add_argument("--height", type=int, default=1080, help="c5: image height")
    ap.add_argument("--spp", type=int, default=64, help="c5: camera paths per pixel")
    ap.add_argument("--no-parents", action="store_true", help="c5: offset next origins by eps * normal instead of carrying parent shading points")
    ap.add_argument("--rays", type=int, default=0, help="rays per batch per GPU (default: the workload's)")
    ap.add_argument("--res", type=int, default=0, help="override the grid resolution (smaller scene for quick runs)")
    ap.add_argument("--cpu-rays", type=int, default=0, help="size of the CPU baseline sample (default: one whole batch)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--tree-build", default="sah", choices=["sah", "device"],
                    help="sah: the reference's sweep SAH on the host (default, result-identical trees); device: linear BVH built by lbvh.cu")
    args = ap.parse_args()
    if args.rays == 0:
        args.rays = {"c1": 512 * 512, "c2": 16 * MI, "c3": 32 * MI, "c4": 16 * MI, "c5": 16 * MI}[args.workload]
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_gpu_c5(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
