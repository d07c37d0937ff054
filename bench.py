#!/usr/bin/env python
"""Headline benchmark: Mrays/s of the Intersector::trace() / trace_probe() hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c3|c2|c4|c1] [--impl reference]

Workloads (BASELINE.json configs):

  c5 (default, at every N): the configuration the metric is quoted on -- "closest-hit & shadow-probe
      at 1/2/4/8 B200 vs CPU Intersector": the 9 999 392-triangle fBm terrain x 64 assembly instances
      (the C3 scene), 1920 x 1080 x 64 spp synthetic path stream (camera ray + 3 cosine bounces + one
      shadow probe per vertex, ~1 G rays per frame) through the wavefront queues
      (asgpu_path_stream_*), rays never leave the device; 32 x 32 tiles dealt to the ranks in
      Hilbert order.  A step is one frame; scaling is STRONG (the frame is fixed, ranks share it).
  c3: the same scene, 32 Mi incoherent closest-hit rays + 32 Mi shadow probes as flat batches
      (the north star's ">= 1 Grays/s incoherent closest-hit on a 10M-triangle instanced scene")
  c2: 999 698-triangle displaced grid, 16 Mi coherent pinhole primaries + 16 Mi incoherent
      cosine-weighted bounce rays, closest hit
  c4: 2 000 000 moving triangles, 16 Mi incoherent closest-hit rays with random times (+ probes);
      msc = 1 and 3 (the reference's mesh reader takes power-of-two pose counts) and the literal
      "2 motion segments" (release builds only)
  c1: the reference's own CPU-runnable case -- Cornell box (32 triangles), 512 x 512 pinhole
      primaries + one ambient-occlusion probe per primary hit

The default run at N = 1 times c5 as the step and then adds the figures of c3 (with its own
roofline, host-buffer e2e and CPU baseline), c2, c4 (msc 1, 2, 3) and c1 as keyed objects of the
same JSON line ("c3", "c2", "c4", "c1"); `--workload cX` makes one of them the step instead.  At
N > 1 the line adds "c3_host": the C3 closest-hit batch dealt to the ranks through the host-buffer
ABI call, with the H2D / D2H rates each GPU saw and the raw pinned-copy ceiling measured beside it.

`value`  = rays / device time (CUDA events on the launch stream), max over ranks.
`e2e`    = the same step through the public API with host buffers inside the timed region: for c5
           tile list in -> this rank's tiles of the image out (asgpu_path_stream_render + _read_tiles); for c1-c4 rays in ->
           hit records out (asgpu_trace_host) from pinned host memory.
`roofline` = algorithmic bytes (measured node / triangle / instance visits per ray x record sizes
           + ray in + hit out) of the dominant kernel's launches / its launch durations, measured live
           with CUDA events over the timed region, against the measured HBM copy bandwidth.
`cpu_baseline` = the reference's CPU path (oracle/_ref when present, else the oracle port) on all
           host cores over a bounded sample of the same workload.

With --gpus N > 1 (launched by torchrun, one rank per GPU) rank 0 builds and flattens the scene,
broadcasts the blob once over NCCL, and every rank traces its own tiles (c5) or rays (c1-c4); no
data-path collective; time = max over ranks.

--impl reference times the CPU reference path alone (rank 0 only) on the same workload and prints
the same line.
"""
from __future__ import annotations

import argparse
import hashlib
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

METRIC_BATCH = "Mrays/s closest-hit (wavefront batches, rays resident in HBM)"
METRIC_STREAM = "Mrays/s closest-hit + shadow-probe (wavefront path stream, queues resident in HBM)"
# Bytes the kernel fetches when a ray enters an assembly instance: 112 of the 128-byte ItemRecord
# (3 x 4 matrix + tree / visibility / id), 32 bytes of TreeDesc fields (node, triangle and pose
# offsets, node and slice counts) and the 4-byte item index of the wide top-level leaf.
INSTANCE_BYTES = 112 + 32 + 4
NODE_BYTES, TRI_BYTES, POSE_BYTES, HIT_OUT_BYTES = 80, 48, 72, 40
UNIT = "Mrays/s"
MI = 1 << 20
DEFAULT_RAYS = {"c1": 512 * 512, "c2": 16 * MI, "c3": 32 * MI, "c4": 16 * MI, "c5": 32 * MI}   # c5: queue capacity (rays per wavefront batch)


# ---------------------------------------------------------------------------------------------
# Workloads
# ---------------------------------------------------------------------------------------------

def workload_name(workload: str, args, msc: int = 1) -> str:
    rays = args.rays or DEFAULT_RAYS[workload]
    if workload == "c5":
        return ("C5: %dx%dx%d spp path stream (camera + 3 cosine bounces + 1 shadow probe per vertex) on the C3 scene "
                "(9999392-triangle fBm terrain x 64 assembly instances), wavefront queues, 32x32 tiles" % (args.width, args.height, args.spp))
    return {
        "c1": "C1: Cornell box (32 triangles), %d pinhole primary rays, closest hit (+ AO probes, extra)%.0s",
        "c2": "C2: 999698-triangle displaced grid, %d coherent primary + %d incoherent cosine bounce rays, closest hit",
        "c3": "C3: 9999392-triangle fBm terrain x 64 assembly instances, %d incoherent closest-hit rays (+ %d shadow probes, extra)",
        "c4": "C4: 2000000 moving triangles (msc=" + str(msc) + "), %d incoherent closest-hit rays with random time (+ %d probes, extra)",
    }[workload] % (rays, rays)


def shared_config(workload: str, args, world: int) -> dict:
    """The part of `config` that describes the WORKLOAD: identical in both arms."""
    cfg = {"workload": workload_name(workload, args, args.msc)}
    if workload == "c5":
        cfg.update({"width": args.width, "height": args.height, "spp": args.spp, "max_bounces": 3, "tile_size": 32,
                    "next_ray_origin": "hit point + eps * normal" if args.no_parents else "parent shading point, refined + offset (ShadingPoint::refine_and_offset)"})
    else:
        cfg["rays_per_batch"] = args.rays or DEFAULT_RAYS[workload]
    return cfg


def make_scene(workload: str, res: int = 0, msc: int = 1):
    if workload == "c1":
        return scenes.scene_c1()
    if workload == "c2":
        return scenes.scene_c2(res or 707)
    if workload in ("c3", "c5"):
        return scenes.scene_c3(res or 2236, 8)
    if workload == "c4":
        return scenes.scene_c4(res or 1000, msc)
    raise ValueError(workload)


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


def corner_lights(desc) -> np.ndarray:
    lo, hi = scenes.scene_bbox(desc)
    return np.array([[lo[0], hi[1] + 2.0, lo[2]], [hi[0], hi[1] + 2.0, lo[2]], [lo[0], hi[1] + 2.0, hi[2]], [hi[0], hi[1] + 2.0, hi[2]]])


def shadow_rays_from(desc, rays: RayBatch, hits: np.ndarray, seed: int) -> RayBatch:
    """C3 / C4 probes: from the hit points (pulled back 1e-6 along the ray) to one of 4 point lights,
    tmax = dist * (1 - 1e-6) (Tracer::trace_between, tracer.h:252-259); they keep their ray's time."""
    hit = hits["prim_type"] == 2
    pts = rays.org + np.where(hit, hits["t"], 0.0)[:, None] * rays.dir - 1e-6 * rays.dir
    sh = scenes.shadow_rays(pts, corner_lights(desc), seed, flags=VIS_SHADOW)
    if rays.time_normalized is not None:
        sh.time_absolute, sh.time_normalized = rays.time_absolute, rays.time_normalized
    return sh


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
# Shared helpers
# ---------------------------------------------------------------------------------------------

def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_sha() -> str:
    """Identifies the trace kernels a committed ncu capture was taken from."""
    h = hashlib.sha1()
    for name in ("kernels.cu", "traverse_core.h", "gpu_layout.h"):
        with open(os.path.join(ROOT, "appleseed_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:12]


def measured_traffic(workload: str, label: str):
    """DRAM bytes per ray of the committed ncu --set full capture of this workload's launch
    (profiles/traffic.json, written by tools/ncu_traffic.py) -- only when the capture was taken from
    the kernels this run executes (same source hash); else None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        entry = json.load(open(path))[workload]
        if entry.get("kernel_source_sha") != kernel_source_sha():
            return None, "profiles/traffic.json[%s] was captured from other kernel sources (%s)" % (workload, entry.get("kernel_source_sha"))
        return entry["launches"][label]["dram_bytes_per_ray"], "ncu --set full capture %s (dram__bytes_read+write per ray x rays per launch)" % entry["source"]
    except Exception:
        return None, None


def bytes_per_ray(per_ray: dict, ray_in: float, out_bytes: float, moving: bool) -> float:
    b = (per_ray["top_nodes"] + per_ray["nodes"]) * NODE_BYTES + per_ray["triangles"] * TRI_BYTES + per_ray["instances"] * INSTANCE_BYTES + ray_in + out_bytes
    if moving:
        b += per_ray["triangles"] * POSE_BYTES      # two 36-byte poses per moving triangle
    return b


def lane_shares(lp: dict) -> dict:
    """asgpu_lane_profile as percentages of the lane slots of the node-test rounds."""
    slots = max(1, 32 * lp["rounds"])
    out = {k: round(100.0 * lp[k] / slots, 1) for k in ("testing", "no_ray", "traversed", "held", "want_enter", "found_leaf", "nothing_to_fetch")}
    out.update({"rounds_per_iteration": round(lp["rounds"] / max(1, lp["iterations"]), 2),
                "lanes_per_batched_entry": round(lp["lanes_entered"] / max(1, lp["batched_entries"]), 1),
                "lanes_per_refill": round(lp["lanes_refilled"] / max(1, lp["refills"]), 1)})
    return out


def per_ray_of(c: dict) -> dict:
    r = max(1, c["rays"])
    return {"top_nodes": c["assembly_nodes_visited"] / r, "instances": c["instances_visited"] / r, "nodes": c["triangle_nodes_visited"] / r,
            "triangles": c["triangles_tested"] / r, "hit_rate": c["hits"] / r}


def cpu_oracle():
    from oracle import oracle as orc
    try:
        orc.build("all")
    except Exception:
        pass
    if orc.available("asref"):
        return orc.Oracle("asref"), "reference"
    return orc.Oracle("orc"), "port"


def pinned(a: np.ndarray):
    import torch
    a = np.ascontiguousarray(a)
    if not a.flags.writeable:
        a = a.copy()
    t = torch.from_numpy(a if a.dtype != np.uint32 else a.view(np.int32))
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


class Dist:
    """torch.distributed state of this process (world 1: no process group)."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.device)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.device)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    def sum(self, values):
        t = self.torch.tensor(values, dtype=self.torch.int64, device=self.device)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(x) for x in t.cpu()]

    def gather(self, values):
        """Every rank's list of floats, on every rank."""
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.device)
        if self.dist is None:
            return [[float(x) for x in t.cpu()]]
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [[float(x) for x in o.cpu()] for o in out]

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def build_context(D: Dist, desc, tree_build: str):
    """Scene built and flattened once on rank 0, replicated with ONE broadcast."""
    from appleseed_b200.intersector import HostTrees, TraceContext
    torch = D.torch
    timing = {"scene_build_s": 0.0, "flatten_upload_s": 0.0, "broadcast_s": 0.0, "tree_build": tree_build}
    ctx = None
    if D.rank == 0:
        trees = HostTrees(desc, threads=0, build_device=D.local_rank if tree_build == "device" else None)
        timing["scene_build_s"] = round(trees.build_seconds, 2)
        t1 = time.perf_counter()
        ctx = TraceContext(device=D.local_rank, trees=trees)
        trees.close()
        timing["flatten_upload_s"] = round(time.perf_counter() - t1, 2)
    if D.world > 1:
        size = torch.tensor([ctx.blob_size if D.rank == 0 else 0], dtype=torch.int64, device=D.device)
        D.dist.broadcast(size, 0)
        blob = ctx.blob_tensor() if D.rank == 0 else torch.empty(int(size.item()), dtype=torch.uint8, device=D.device)
        D.barrier()
        t1 = time.perf_counter()
        D.dist.broadcast(blob, 0)
        torch.cuda.synchronize()
        timing["broadcast_s"] = round(time.perf_counter() - t1, 4)
        timing["broadcast_GBps"] = round(blob.numel() / max(timing["broadcast_s"], 1e-9) / 1e9, 1)
        if D.rank != 0:
            ctx = TraceContext.from_blob(blob, adopt=True)
    return ctx, timing


# ---------------------------------------------------------------------------------------------
# Flat batches: C1 - C4
# ---------------------------------------------------------------------------------------------

def make_batches(workload: str, desc, isect, n: int, device, rank: int, world: int):
    """This rank's ray batches: [(label, PinnedRays)], probe batch or None.  C3 / C4: the batch of
    `n` rays is cut into `world` contiguous ranges (strong scaling); C2: one camera per rank (weak)."""
    import torch
    from appleseed_b200.intersector import HIT_BYTES, hits_from_tensor

    def traced(p):
        out = torch.empty(p.n * HIT_BYTES, dtype=torch.uint8, device=device)
        isect.trace_device(p.dev, out)
        torch.cuda.synchronize()
        return hits_from_tensor(out, p.n)

    if workload == "c2":
        prim = primary_rays_c2(n, rank)
        p = PinnedRays(prim, device)
        mask, pts, nrm = scenes.hit_points_and_normals(desc, prim, traced(p))
        if len(pts) < len(prim):            # pad the bounce wavefront by wrapping around
            idx = np.resize(np.arange(len(pts)), len(prim))
            pts, nrm = pts[idx], nrm[idx]
        bounce = scenes.bounce_rays(pts, nrm, 1 + rank, flags=VIS_DIFFUSE)
        return [("coherent_primary", p), ("incoherent_bounce", PinnedRays(bounce, device))], None
    if workload == "c1":
        prim = scenes.rays_c1_primary()
        p = PinnedRays(prim, device)
        _, ao = scenes.rays_c1_ao(desc, prim, traced(p))
        return [("primary", p)], PinnedRays(ao, device)
    inc = incoherent_rays(desc, n, 2, time=(workload == "c4"))
    if world > 1:
        lo, hi = n * rank // world, n * (rank + 1) // world
        inc = inc.slice(lo, hi)
    p = PinnedRays(inc, device)
    sh = shadow_rays_from(desc, inc, traced(p), 3 + rank)
    return [("incoherent", p)], PinnedRays(sh, device)


def time_batches(D: Dist, isect, batches, steps: int, warmup: int, probe: bool = False):
    """CUDA-event time per batch (ms, averaged over `steps`) after `warmup` untimed passes."""
    import torch
    from appleseed_b200.intersector import HIT_BYTES
    outs = [torch.empty(b.n * (1 if probe else HIT_BYTES), dtype=torch.uint8, device=D.device) for _, b in batches]
    run = (lambda b, o: isect.trace_probe_device(b.dev, o)) if probe else (lambda b, o: isect.trace_device(b.dev, o))
    for _ in range(warmup):
        for (_, b), o in zip(batches, outs):
            run(b, o)
    D.barrier()
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in batches] for _ in range(steps)]
    for s in range(steps):
        for k, ((_, b), o) in enumerate(zip(batches, outs)):
            ev[s][k][0].record()
            run(b, o)
            ev[s][k][1].record()
    D.barrier()
    ms = [sum(ev[s][k][0].elapsed_time(ev[s][k][1]) for s in range(steps)) / steps for k in range(len(batches))]
    return ms, outs


def host_e2e(D: Dist, isect, batches, steps: int, check=None):
    """The step through asgpu_trace_host: pinned host rays in, hit records out, copies inside."""
    import torch
    from appleseed_b200.intersector import HIT_BYTES, hits_from_tensor
    host_hits = [torch.empty(b.n * HIT_BYTES, dtype=torch.uint8).pin_memory() for _, b in batches]
    host_np = [h.numpy().view(HIT_DTYPE) for h in host_hits]

    def step():
        for (_, b), out in zip(batches, host_np):
            isect.trace(b.host, out=out)

    step()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    mine = (time.perf_counter() - t0) * 1e3 / steps
    D.barrier()
    if check is not None:       # the e2e results must be the device results
        assert hits_from_tensor(check, batches[-1][1].n).tobytes() == host_np[-1].tobytes(), "host-path results differ from device-path results"
    return mine, host_np


def batch_figures(D: Dist, workload: str, args, desc, ctx, isect, timing, steps: int, warmup: int, msc: int = 1, with_cpu: bool = True, with_e2e: bool = True):
    """Everything bench.py reports for one of the flat-batch workloads on this process group."""
    n = args.rays or DEFAULT_RAYS[workload]
    info = ctx.info()
    batches, probe_batch = make_batches(workload, desc, isect, n, D.device, D.rank, D.world)
    rays_step = sum(b.n for _, b in batches)
    sampler = ClockSampler(D.local_rank)
    if D.rank == 0:
        sampler.start()
    ms, outs = time_batches(D, isect, batches, steps, warmup)
    clocks = sampler.stop() if D.rank == 0 else None
    probe_ms = time_batches(D, isect, [("probe", probe_batch)], steps, 3, probe=True)[0][0] if probe_batch is not None else 0.0

    e2e_ms, host_np = (host_e2e(D, isect, batches, max(1, min(steps, 5)), check=outs[-1]) if with_e2e else (0.0, None))
    h2d = sum(b.bytes_in for _, b in batches)
    d2h = sum(b.n * HIT_OUT_BYTES for _, b in batches)

    # Algorithmic bytes per ray (counters variant of the same launches, untimed).
    ctx.counters(reset=True)
    for (_, b), o in zip(batches, outs):
        isect.trace_device(b.dev, o, counters=True)
    lanes = lane_shares(ctx.lane_profile()[0])
    per_ray = per_ray_of(ctx.counters(reset=True))
    probe_per_ray = None
    if probe_batch is not None:
        occ = D.torch.empty(probe_batch.n, dtype=D.torch.uint8, device=D.device)
        isect.trace_probe_device(probe_batch.dev, occ, counters=True)
        probe_per_ray = per_ray_of(ctx.counters(reset=True))
    ray_in = h2d / rays_step
    bpr = bytes_per_ray(per_ray, ray_in, HIT_OUT_BYTES, workload == "c4")

    ms_step, e2e_max, probe_max = D.max([sum(ms), e2e_ms, probe_ms])
    total_rays, total_probe = D.sum([rays_step, probe_batch.n if probe_batch is not None else 0])
    rates = D.gather([h2d / max(e2e_ms, 1e-9) / 1e6, d2h / max(e2e_ms, 1e-9) / 1e6])       # GB/s per GPU during its own e2e steps
    out = {
        "workload": workload_name(workload, args, msc), "value": total_rays / ms_step / 1e3, "unit": UNIT, "ms_per_step": ms_step,
        "rays_per_step": total_rays, "rays_per_step_per_gpu": rays_step,
        "batches": {label: {"rays": b.n, "ms": round(m, 4), "mrays_s": round(b.n / m / 1e3, 1)} for (label, b), m in zip(batches, ms)},
        "per_ray": {k: round(v, 3) for k, v in per_ray.items()}, "lane_slots_pct": lanes,
        "scene": {k: info[k] for k in ("triangle_count", "moving_triangle_count", "instance_count", "wide_node_count", "binary_node_count", "blob_bytes")},
        "timing": timing, "launches": steps * len(batches), "clocks": clocks,
        "l2": "inputs larger than L2 (%.0f MB of rays + %.0f MB scene blob per step vs 126 MB L2)" % (h2d / 1e6, info["blob_bytes"] / 1e6),
    }
    if probe_batch is not None:
        out["shadow_probe"] = {"rays": total_probe, "ms": round(probe_max, 4), "mrays_s": round(total_probe / probe_max / 1e3, 1),
                               "per_ray": {k: round(v, 3) for k, v in probe_per_ray.items()}}
    if with_e2e:
        out["e2e"] = {"value": total_rays / e2e_max / 1e3, "unit": UNIT, "h2d_bytes_per_step": h2d * D.world, "d2h_bytes_per_step": d2h * D.world, "ms_per_step": e2e_max,
                      "h2d_GBps_per_gpu": [round(r[0], 1) for r in rates], "d2h_GBps_per_gpu": [round(r[1], 1) for r in rates]}
    peak, peak_src = hbm_peak()
    dominant = max(range(len(batches)), key=lambda k: ms[k])
    label = batches[dominant][0]
    achieved = bpr * rays_step / (sum(ms) * 1e-3) / 1e9                 # per GPU, all launches of the step
    traffic_per_ray, traffic_src = measured_traffic(workload if workload != "c4" else "c4", label)
    out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                       "traffic": traffic_per_ray * batches[dominant][1].n if traffic_per_ray is not None else None,
                       "algorithmic_bytes_per_launch": bpr * rays_step / len(batches), "bytes_per_ray": round(bpr, 1),
                       "peak_source": peak_src, "traffic_source": traffic_src, "kernel": "wide_kernel<closest>",
                       "launch_ms": round(sum(ms) / len(batches), 4),
                       "note": "algorithmic bytes are mostly served by L1/L2: the kernel is instruction-issue bound, see profiles/README.md"}
    if D.rank == 0 and D.world == 1 and with_cpu and not args.no_cpu:
        oracle, kind = cpu_oracle()
        threads = os.cpu_count() or 1
        oscene = oracle.scene(desc)
        src = batches[-1][1]
        cpu_n = min(args.cpu_rays if args.cpu_rays > 0 else 4 * MI, src.n)
        sample = src.host.slice(0, cpu_n)
        t0 = time.perf_counter()
        cpu_hits = oscene.trace(sample, threads=threads)
        secs = time.perf_counter() - t0
        gpu_hits = host_np[-1][:cpu_n] if host_np is not None else None
        agree = float((cpu_hits["tri_slot"] == gpu_hits["tri_slot"]).mean()) if gpu_hits is not None else None
        identical = float((cpu_hits.view(np.uint8).reshape(cpu_n, -1) == gpu_hits.view(np.uint8).reshape(cpu_n, -1)).all(axis=1).mean()) if gpu_hits is not None else None
        out["cpu_baseline"] = {"value": cpu_n / secs / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
                               "sample": "first %d rays of the '%s' batch, %d threads, %.1f s" % (cpu_n, batches[-1][0], threads, secs),
                               "identity_agreement_with_gpu": agree, "records_byte_identical_with_gpu": identical}
        if probe_batch is not None:
            ps = probe_batch.host.slice(0, min(cpu_n, probe_batch.n))
            t0 = time.perf_counter()
            oscene.trace_probe(ps, threads=threads)
            out["cpu_baseline"]["shadow_probe_value"] = len(ps) / (time.perf_counter() - t0) / 1e6
        oscene.close()
    return out


def run_gpu_batches(args):
    from appleseed_b200.intersector import Intersector
    D = Dist()
    desc = make_scene(args.workload, args.res, args.msc)
    ctx, timing = build_context(D, desc, args.tree_build)
    isect = Intersector(ctx)
    f = batch_figures(D, args.workload, args, desc, ctx, isect, timing, args.steps, max(3, args.warmup), msc=args.msc)
    if D.rank == 0:
        cfg = shared_config(args.workload, args, D.world)
        cfg.update({"l2": f["l2"], "parallelism": ("rays dealt to ranks in contiguous ranges, scene replicated by one NCCL broadcast" if D.world > 1 else "1 GPU")})
        line = {"metric": METRIC_BATCH, "value": f["value"], "unit": UNIT, "n_gpus": D.world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": f["ms_per_step"], "higher_is_better": True, "scaling": "weak" if args.workload in ("c1", "c2") else "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "detail": {k: f[k] for k in ("rays_per_step", "rays_per_step_per_gpu", "batches", "per_ray", "scene", "timing") if k in f},
                "e2e": f["e2e"], "gpu_launches": f["launches"], "roofline": f["roofline"], "clocks": f["clocks"]}
        if "shadow_probe" in f:
            line["detail"]["shadow_probe"] = f["shadow_probe"]
        if "cpu_baseline" in f:
            line["cpu_baseline"] = f["cpu_baseline"]
        print(json.dumps(line))
    D.close()


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
    return dict(width=args.width, height=args.height, spp=args.spp, camera_to_world=look_at(eye, centre), lights=corner_lights(desc),
                max_bounces=3, tile_size=32, seed=5, offset_eps=1.0e-6 * diag, parents=not getattr(args, "no_parents", False))


def cpu_path_stream(desc, oscene, cfg: dict, width: int, height: int, threads: int, seed: int = 7):
    """The same path stream restated on the host for the CPU arm: rays generated with numpy (same
    camera, same sampling distributions), traced by the reference's CPU path -- with parent shading
    points refined and offset per hit (ShadingPoint::refine_and_offset) when the stream carries them,
    as Intersector::trace(ray, shading_point, parent) does.  Returns (rays traced, seconds spent in
    the reference path: trace / trace_probe / refine_and_offset calls)."""
    cam = scenes.camera_rays(width, height, cfg["camera_to_world"], (0.025, 0.025 * height / width), 0.035)
    parents = cfg.get("parents", False)
    rays, par, total, secs = cam, None, 0, 0.0
    for depth in range(cfg["max_bounces"] + 1):
        t0 = time.perf_counter()
        hits = oscene.trace(rays, threads=threads) if par is None else oscene.trace_parents(rays, par, threads=threads)
        secs += time.perf_counter() - t0
        total += len(rays)
        mask, pts, nrm = scenes.hit_points_and_normals(desc, rays, hits)
        if len(pts) == 0:
            break
        if parents:
            t0 = time.perf_counter()
            refined = oscene.refine_offset(rays, hits, threads=threads)[hits["prim_type"] == 2]
            secs += time.perf_counter() - t0
            org = pts
        else:
            refined, org = None, pts + cfg["offset_eps"] * nrm
        sh = scenes.shadow_rays(org, np.asarray(cfg["lights"]), seed + 100 + depth)
        t0 = time.perf_counter()
        if parents:
            oscene.trace_probe_parents(sh, refined, threads=threads)
        else:
            oscene.trace_probe(sh, threads=threads)
        secs += time.perf_counter() - t0
        total += len(sh)
        rays = scenes.bounce_rays(pts, nrm, seed + depth, flags=VIS_DIFFUSE, offset=0.0 if parents else cfg["offset_eps"])
        par = refined
    return total, secs


def host_copy_ceiling(D: Dist, nbytes: int = 1 << 30, reps: int = 3):
    """Raw pinned-memory copy rates with every rank copying at once (GB/s per GPU): the ceiling any
    host-buffer path on this box lives under."""
    torch = D.torch
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dev = torch.empty(nbytes, dtype=torch.uint8, device=D.device)
    out = []
    for direction in ("h2d", "d2h", "both"):
        s2 = torch.cuda.Stream()
        host2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory() if direction == "both" else None
        dev2 = torch.empty(nbytes, dtype=torch.uint8, device=D.device) if direction == "both" else None
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if direction in ("h2d", "both"):
                dev.copy_(host, non_blocking=True)
            if direction == "d2h":
                host.copy_(dev, non_blocking=True)
            if direction == "both":
                with torch.cuda.stream(s2):
                    host2.copy_(dev2, non_blocking=True)
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        D.barrier()
        out.append(nbytes * reps / secs / 1e9)
    return out      # h2d alone, d2h alone, h2d while d2h runs


def run_gpu_c5(args):
    from appleseed_b200.distributed import tile_ids_shard
    from appleseed_b200.intersector import Intersector
    from appleseed_b200.wavefront import PathStream, PathStreamConfig
    D = Dist()
    torch = D.torch
    desc = make_scene("c5", args.res)
    ctx, timing = build_context(D, desc, args.tree_build)
    info = ctx.info()
    capacity = args.rays or DEFAULT_RAYS["c5"]
    cfg = c5_config(args, desc)
    tiles = tile_ids_shard(args.width, args.height, D.world, D.rank, 32)
    ps = PathStream(ctx, PathStreamConfig(**cfg), queue_capacity=capacity)
    warmup = max(3, args.warmup)

    for _ in range(warmup):
        ps.render(tiles)
    torch.cuda.synchronize()
    ps.clear()
    ps.set_profiling(True)          # CUDA events around every launch of the timed frames
    D.barrier()
    sampler = ClockSampler(D.local_rank)
    if D.rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    D.barrier()
    for s in range(args.steps):
        ev[s][0].record()
        ps.render(tiles)
        ev[s][1].record()
    D.barrier()
    ms_step = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    clocks = sampler.stop() if D.rank == 0 else None
    st = ps.stats()
    prof = ps.profile()
    ps.set_profiling(False)
    rays_step = (st["camera_rays"] + st["bounce_rays"] + st["probe_rays"]) // args.steps
    closest_step = (st["camera_rays"] + st["bounce_rays"]) // args.steps
    probe_step = st["probe_rays"] // args.steps
    launches = st["kernel_launches"]

    # ---- end to end: clear, tile list in, image out, through the public API (host buffers) ----
    # Every rank reads back the pixels of ITS tiles (asgpu_path_stream_read_tiles): 1 / N of the frame.
    ts = cfg["tile_size"]
    image_host = torch.empty((len(tiles), ts, ts, 4), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    D.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        ps.clear()
        ps.render(tiles)
        image = ps.image_tiles(tiles, out=image_host)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    D.barrier()
    assert int(ps.image().astype(np.uint64).sum()) == int(image.astype(np.uint64).sum()), "tile read-back differs from the frame"
    # One frame's accumulators of this rank's pixels; summed over ranks it does not depend on N.
    checksum = int(image.astype(np.uint64).sum())

    # ---- algorithmic bytes per ray (counters variant of an eighth of the same tiles, untimed) --
    ps.close()
    psc = PathStream(ctx, PathStreamConfig(**{**cfg, "counters": True}), queue_capacity=capacity)
    ctx.counters(reset=True)
    psc.render(tiles[: max(1, len(tiles) // 8)])
    lp_closest, lp_probe = ctx.lane_profile()
    c_closest, c_probe = ctx.counters_by_kind(reset=True)
    psc.close()
    pr_closest, pr_probe = per_ray_of(c_closest), per_ray_of(c_probe)
    ray_in = 76.0       # org, dir, tmin, tmax (64) + two times (8) + flags (4); parents ride in their own array
    parent_bytes = 80.0 if cfg["parents"] else 0.0
    bpr_closest = bytes_per_ray(pr_closest, ray_in + parent_bytes, HIT_OUT_BYTES, False)
    bpr_probe = bytes_per_ray(pr_probe, ray_in + parent_bytes, 1.0, False)

    ms_step, e2e_ms = D.max([ms_step, e2e_ms])
    total_rays, total_closest, total_probe, checksum, launches, tiles_total = D.sum([rays_step, closest_step, probe_step, checksum, launches, len(tiles)])
    value = total_rays / (ms_step * 1e-3) / 1e6
    line = None
    if D.rank == 0:
        peak, peak_src = hbm_peak()
        # The dominant kernel: wide_kernel<closest> over the stream's closest-hit queues.  Its launches
        # of the timed frames were bracketed by CUDA events on the launch stream (rank 0's).
        closest_launch_ms = prof["closest_ms"] / max(1, prof["closest_launches"])
        closest_rays_per_launch = closest_step * args.steps / max(1, prof["closest_launches"])
        achieved = bpr_closest * closest_rays_per_launch / (closest_launch_ms * 1e-3) / 1e9
        traffic_per_ray, traffic_src = measured_traffic("c5", "closest")
        frame_bytes = bpr_closest * closest_step + bpr_probe * probe_step
        cfgd = shared_config("c5", args, D.world)
        cfgd.update({
            "l2": "inputs larger than L2 (%.0f MB scene blob, %.0f MB of queued rays per wavefront vs 126 MB L2)" % (info["blob_bytes"] / 1e6, capacity * 76 / 1e6),
            "parallelism": "tiles dealt to ranks in Hilbert order, scene replicated by one NCCL broadcast" if D.world > 1 else "1 GPU"})
        line = {
            "metric": METRIC_STREAM, "value": value, "unit": UNIT, "n_gpus": D.world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfgd,
            "detail": {
                "rays_per_frame": total_rays, "closest_rays_per_frame": total_closest, "probe_rays_per_frame": total_probe,
                "tiles_per_gpu": int(len(tiles)), "queue_capacity": capacity, "image_checksum": checksum,
                "scene": {k: info[k] for k in ("triangle_count", "instance_count", "wide_node_count", "binary_node_count", "blob_bytes")},
                "timing": timing,
                "per_ray_closest": {k: round(v, 3) for k, v in pr_closest.items()}, "per_ray_probe": {k: round(v, 3) for k, v in pr_probe.items()},
                "lane_slots_closest_pct": lane_shares(lp_closest), "lane_slots_probe_pct": lane_shares(lp_probe),
                "rank0_device_ms_per_frame": {"closest_trace": round(prof["closest_ms"] / args.steps, 3), "probe_trace": round(prof["probe_ms"] / args.steps, 3),
                                              "refine_and_offset": round(prof["refine_ms"] / args.steps, 3),
                                              "generate_shade_accumulate": round(prof["stage_ms"] / args.steps, 3),
                                              "note": "refine_and_offset runs inside shade_kernel (0 = no kernel of its own; ASGPU_FUSE_REFINE=0 splits it out)"},
                "rank0_closest_mrays_s": round(closest_step / (prof["closest_ms"] / args.steps) / 1e3, 1) if prof["closest_ms"] else None,
                "rank0_probe_mrays_s": round(probe_step / (prof["probe_ms"] / args.steps) / 1e3, 1) if prof["probe_ms"] else None,
            },
            "e2e": {"value": total_rays / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(len(tiles)) * 4 * D.world,
                    "d2h_bytes_per_step": int(tiles_total) * ts * ts * 16, "ms_per_step": e2e_ms,
                    "note": "tile list in, this rank's tiles of the image out through asgpu_path_stream_render + _read_tiles; the rays are generated, traced and shaded on the device"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic_per_ray * closest_rays_per_launch if traffic_per_ray is not None else None,
                         "algorithmic_bytes_per_launch": bpr_closest * closest_rays_per_launch, "bytes_per_ray": round(bpr_closest, 1),
                         "launch_ms": round(closest_launch_ms, 4), "launches_timed": prof["closest_launches"],
                         "peak_source": peak_src, "traffic_source": traffic_src, "kernel": "wide_kernel<closest> (path stream closest-hit queues)",
                         "whole_frame": {"achieved": frame_bytes / (ms_step * 1e-3) / 1e9, "frac": frame_bytes / (ms_step * 1e-3) / 1e9 / peak,
                                         "bytes_per_probe_ray": round(bpr_probe, 1),
                                         "note": "rank 0's closest + probe bytes of one frame over the whole step, stage kernels included in the time"},
                         "note": "algorithmic bytes are mostly served by L1/L2: the kernel is instruction-issue bound, see profiles/README.md"},
            "clocks": clocks,
        }
    # ---- extra figures -----------------------------------------------------------------------
    if not args.no_extras:
        isect = Intersector(ctx)
        xargs = argparse.Namespace(**{**vars(args), "rays": 0})
        if D.world == 1:
            c3 = batch_figures(D, "c3", xargs, desc, ctx, isect, timing, args.extra_steps, 3)
            if line is not None:
                line["c3"] = c3
        else:
            # C3's closest-hit batch dealt to the ranks through the HOST-buffer call, and what the box
            # lets through when every GPU copies at once.
            c3 = batch_figures(D, "c3", xargs, desc, ctx, isect, timing, args.extra_steps, 3, with_cpu=False)
            ceiling = D.gather(host_copy_ceiling(D))
            if line is not None:
                line["c3_host"] = {"workload": c3["workload"], "device_value": c3["value"], "e2e": c3["e2e"],
                                   "pinned_copy_ceiling_GBps_per_gpu": {"h2d_alone": [round(c[0], 1) for c in ceiling], "d2h_alone": [round(c[1], 1) for c in ceiling],
                                                                        "h2d_with_d2h": [round(c[2], 1) for c in ceiling]},
                                   "bytes_per_ray_over_pcie": {"in": c3["e2e"]["h2d_bytes_per_step"] / c3["rays_per_step"], "out": HIT_OUT_BYTES}}
    if line is not None and D.world == 1 and not args.no_cpu:
        oracle, kind = cpu_oracle()
        threads = os.cpu_count() or 1
        oscene = oracle.scene(desc)
        w, h = args.cpu_width or args.width, args.cpu_height or args.height
        n, secs = cpu_path_stream(desc, oscene, cfg, w, h, threads)
        line["cpu_baseline"] = {"value": n / secs / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
                                "sample": "%dx%d x 1 spp of the same path stream (%d rays), %d threads, %.1f s" % (w, h, n, threads, secs)}
        oscene.close()
    ctx.close()
    if not args.no_extras and D.world == 1:
        # The other single-GPU configurations, each on its own scene.
        small = max(3, args.extra_steps)
        for name, key, msc in (("c2", "c2", 1), ("c4", "msc1", 1), ("c4", "msc2", 2), ("c4", "msc3", 3), ("c1", "c1", 1)):
            d2 = make_scene(name, 0, msc)
            ctx2, timing2 = build_context(D, d2, args.tree_build)
            xargs = argparse.Namespace(**{**vars(args), "rays": 0, "cpu_rays": 2 * MI})
            f = batch_figures(D, name, xargs, d2, ctx2, Intersector(ctx2), timing2, small, 3, msc=msc)
            ctx2.close()
            if name == "c4":
                line.setdefault("c4", {})[key] = f
                if msc == 2:
                    f["note"] = "the literal '2 motion segments' of BASELINE.json: release builds only (meshobjectreader.cpp:870-878 takes power-of-two pose counts)"
            else:
                line[key] = f
    if line is not None:
        print(json.dumps(line))
    D.close()


# ---------------------------------------------------------------------------------------------
# CPU reference arm
# ---------------------------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    desc = make_scene(args.workload, args.res, args.msc)
    oracle, kind = cpu_oracle()
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    oscene = oracle.scene(desc)
    build_s = time.perf_counter() - t0
    cfg = shared_config(args.workload, args, world)
    warmup = max(0, args.warmup)
    if args.workload == "c5":
        scfg = c5_config(args, desc)
        w, h = args.cpu_width or args.width, args.cpu_height or args.height
        sample = "%dx%d x 1 spp of the %d spp path stream per step" % (w, h, args.spp)
        for k in range(warmup):
            cpu_path_stream(desc, oscene, scfg, w, h, threads, seed=100 + k)
        ns, ts = [], []
        for k in range(args.steps):
            n, secs = cpu_path_stream(desc, oscene, scfg, w, h, threads, seed=7 + k)
            ns.append(n); ts.append(secs)
        ms = 1e3 * sum(ts) / len(ts)
        value = sum(ns) / sum(ts) / 1e6
        metric, scaling = METRIC_STREAM, "strong"
        sample += " (%d rays)" % ns[0]
    else:
        n = args.rays or DEFAULT_RAYS[args.workload]
        # The identical ray set of the GPU arm (rank 0's), bounded by --cpu-rays when it is given.
        if args.workload == "c2":
            prim = primary_rays_c2(n, 0)
            hits = oscene.trace(prim, threads=threads)
            mask, pts, nrm = scenes.hit_points_and_normals(desc, prim, hits)
            if len(pts) < len(prim):
                idx = np.resize(np.arange(len(pts)), len(prim))
                pts, nrm = pts[idx], nrm[idx]
            batches = [prim, scenes.bounce_rays(pts, nrm, 1, flags=VIS_DIFFUSE)]
        elif args.workload == "c1":
            batches = [scenes.rays_c1_primary()]
        else:
            batches = [incoherent_rays(desc, n, 2, time=(args.workload == "c4"))]
        if args.cpu_rays > 0:
            batches = [b.slice(0, min(len(b), args.cpu_rays)) for b in batches]
        n_rays = sum(len(b) for b in batches)
        sample = "%d rays per step (%s)" % (n_rays, "the whole ray set of the GPU arm" if args.cpu_rays <= 0 else "the first --cpu-rays rays of every batch")
        for _ in range(warmup):
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
        metric, scaling = METRIC_BATCH, "weak" if args.workload in ("c1", "c2") else "strong"
    cfg.update({"l2": "n/a (CPU arm)", "parallelism": "%d host threads over contiguous ray ranges" % threads})
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg, "detail": {"sample": sample, "scene_build_s": round(build_s, 2)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample + ", %d threads" % threads},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--width", type=int, default=1920, help="c5: image width")
    ap.add_argument("--height", type=int, default=1080, help="c5: image height")
    ap.add_argument("--spp", type=int, default=64, help="c5: camera paths per pixel")
    ap.add_argument("--msc", type=int, default=1, help="c4: motion segment count")
    ap.add_argument("--no-parents", action="store_true", help="c5: offset next origins by eps * normal instead of carrying parent shading points")
    ap.add_argument("--rays", type=int, default=0, help="rays per batch (c5: queue capacity) -- default: the workload's")
    ap.add_argument("--res", type=int, default=0, help="override the grid resolution (smaller scene for quick runs)")
    ap.add_argument("--cpu-rays", type=int, default=0, help="size of the CPU sample (c1-c4; default: 4 Mi rays beside the GPU arm, the whole ray set for --impl reference)")
    ap.add_argument("--cpu-width", type=int, default=0, help="c5: width of the 1-spp CPU sample (default: the image's)")
    ap.add_argument("--cpu-height", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="c5: only the step, none of the c3 / c2 / c4 / c1 figures")
    ap.add_argument("--extra-steps", type=int, default=5, help="timed passes of every extra figure")
    ap.add_argument("--tree-build", default="sah", choices=["sah", "device"],
                    help="sah: the reference's sweep SAH on the host (default, result-identical trees); device: linear BVH built by lbvh.cu")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_gpu_c5(args)
    else:
        run_gpu_batches(args)


if __name__ == "__main__":
    main()
