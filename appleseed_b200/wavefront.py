"""Wavefront ray queues and the synthetic path stream (include/asgpu.h, "Wavefront ray queues").

Host-side mirror of the restructured trace loop: where the reference renders one sample by
recursion (``GenericSampleRenderer::render_sample``, renderer/kernel/rendering/generic/
genericsamplerenderer.cpp:164-299 -> ``PathTracer`` -> ``Intersector::trace`` /
``Tracer::trace_between``), ``PathStream.render`` pushes whole tiles through device-resident
closest-hit and shadow-probe queues.  All work happens in ``libasgpu.so``; this module only holds
handles and converts arrays.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .intersector import AsgpuError, TraceContext, _check
from .scene import HIT_DTYPE, PARENT_DTYPE, RayBatch


class RayQueue:
    """A device-resident ray queue (``asgpu_ray_queue``)."""

    def __init__(self, ctx: TraceContext, capacity: int):
        self.ctx = ctx
        self.lib = ctx.lib
        self.handle = self.lib.asgpu_queue_create(ctx.handle, capacity)
        if not self.handle:
            raise AsgpuError("asgpu_queue_create failed: " + _lib.last_error())
        self.capacity = int(self.lib.asgpu_queue_capacity(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.asgpu_queue_destroy(self.handle)
            self.handle = None

    __del__ = close

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.ctx.device).cuda_stream)

    def reset(self):
        _check(self.lib.asgpu_queue_reset(self.handle, self._stream()), "asgpu_queue_reset")

    def __len__(self) -> int:
        n = C.c_uint64(0)
        _check(self.lib.asgpu_queue_count(self.handle, self._stream(), C.byref(n)), "asgpu_queue_count")
        return int(n.value)

    def push(self, rays: RayBatch, path_ids: Optional[np.ndarray] = None):
        cr = rays.to_c()
        ids = None if path_ids is None else np.ascontiguousarray(path_ids, dtype=np.uint32)
        _check(self.lib.asgpu_queue_push_host(self.handle, C.byref(cr), None if ids is None else ids.ctypes.data, len(rays), self._stream()),
               "asgpu_queue_push_host")

    def trace(self, hits: "torch.Tensor", exact: bool = False):
        flags = _lib.TRACE_EXACT if exact else 0
        _check(self.lib.asgpu_trace_queue(self.ctx.handle, self.handle, hits.data_ptr(), flags, self._stream()), "asgpu_trace_queue")

    def trace_probe(self, occluded: "torch.Tensor", exact: bool = False):
        flags = _lib.TRACE_EXACT if exact else 0
        _check(self.lib.asgpu_trace_probe_queue(self.ctx.handle, self.handle, occluded.data_ptr(), flags, self._stream()), "asgpu_trace_probe_queue")


@dataclass
class PathStreamConfig:
    """``asgpu_path_stream_desc``: the synthetic path stream of BASELINE.json configs[4]."""
    width: int
    height: int
    spp: int
    camera_to_world: np.ndarray            # 4 x 4 or 3 x 4, row-major
    lights: np.ndarray                     # k x 3 point lights, 1 <= k <= 8
    max_bounces: int = 3
    tile_size: int = 32
    seed: int = 5
    film_width: float = 0.025
    film_height: Optional[float] = None    # default: film_width * height / width
    focal_length: float = 0.035
    offset_eps: float = 1.0e-6
    exact: bool = False
    counters: bool = False                 # accumulate the scene's traversal counters (slower)
    parents: bool = False                  # child rays carry their refined parent shading point (ASGPU_STREAM_PARENTS)
    shutter_open: float = 0.0              # ray time: one normalized time per camera path, absolute = lerp(open, close, t);
    shutter_close: float = 0.0             # open == close: every ray at time 0

    def to_c(self) -> "_lib.PathStreamDesc":
        d = _lib.PathStreamDesc()
        d.width, d.height, d.spp, d.max_bounces, d.tile_size = self.width, self.height, self.spp, self.max_bounces, self.tile_size
        lights = np.asarray(self.lights, dtype=np.float64).reshape(-1, 3)
        d.light_count = lights.shape[0]
        d.trace_flags = (_lib.TRACE_EXACT if self.exact else 0) | (_lib.TRACE_COUNTERS if self.counters else 0)
        d.stream_flags = _lib.STREAM_PARENTS if self.parents else 0
        d.seed = self.seed
        m = np.asarray(self.camera_to_world, dtype=np.float64).reshape(-1, 4)[:3]
        for i, v in enumerate(m.reshape(-1)):
            d.camera_to_world[i] = float(v)
        d.film_width = self.film_width
        d.film_height = self.film_height if self.film_height is not None else self.film_width * self.height / self.width
        d.focal_length = self.focal_length
        for k in range(min(8, lights.shape[0])):
            for a in range(3):
                d.lights[k][a] = float(lights[k, a])
        d.offset_eps = self.offset_eps
        d.shutter_open, d.shutter_close = self.shutter_open, self.shutter_close
        return d


@dataclass
class CapturedWavefront:
    kind: str                  # "closest" | "probe"
    depth: int
    rays: RayBatch
    path_ids: np.ndarray
    results: np.ndarray        # HIT_DTYPE records or uint8 occlusion flags
    parents: Optional[np.ndarray] = None       # PARENT_DTYPE records the rays carried


def look_at(origin, target, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """Camera-to-world matrix of a camera at ``origin`` looking at ``target`` (camera looks down -Z)."""
    origin = np.asarray(origin, dtype=np.float64)
    z = origin - np.asarray(target, dtype=np.float64)
    z /= np.linalg.norm(z)
    x = np.cross(np.asarray(up, dtype=np.float64), z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, origin
    return m


class PathStream:
    """``asgpu_path_stream``: tiles in, per-pixel accumulators out, everything in between on the GPU."""

    def __init__(self, ctx: TraceContext, config: PathStreamConfig, queue_capacity: int = 16 << 20):
        self.ctx = ctx
        self.lib = ctx.lib
        self.config = config
        d = config.to_c()
        self.handle = self.lib.asgpu_path_stream_create(ctx.handle, C.byref(d), queue_capacity)
        if not self.handle:
            raise AsgpuError("asgpu_path_stream_create failed: " + _lib.last_error())
        self.tile_count = int(self.lib.asgpu_path_stream_tile_count(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.asgpu_path_stream_destroy(self.handle)
            self.handle = None

    __del__ = close

    def render(self, tiles: Optional[Sequence[int]] = None):
        """Enqueue the given tiles (default: the whole frame) on torch's current stream."""
        import torch
        t = np.arange(self.tile_count, dtype=np.uint32) if tiles is None else np.ascontiguousarray(tiles, dtype=np.uint32)
        stream = C.c_void_p(torch.cuda.current_stream(self.ctx.device).cuda_stream)
        _check(self.lib.asgpu_path_stream_render(self.handle, t.ctypes.data if len(t) else None, len(t), stream), "asgpu_path_stream_render")

    def image(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Per-pixel accumulators (height x width x 4 uint32).  ``out``: a caller-owned array to fill
        (page-locked memory makes the device-to-host copy several times faster)."""
        if out is None:
            out = np.empty((self.config.height, self.config.width, 4), dtype=np.uint32)
        assert out.dtype == np.uint32 and out.size == self.config.height * self.config.width * 4 and out.flags.c_contiguous
        _check(self.lib.asgpu_path_stream_read_image(self.handle, out.ctypes.data), "asgpu_path_stream_read_image")
        return out

    def image_tiles(self, tiles: Sequence[int], out: Optional[np.ndarray] = None) -> np.ndarray:
        """Accumulators of the listed tiles only (len(tiles) x tile_size x tile_size x 4 uint32, zero
        outside the image): what a process that rendered a share of the frame reads back."""
        t = np.ascontiguousarray(tiles, dtype=np.uint32)
        ts = self.config.tile_size
        if out is None:
            out = np.empty((len(t), ts, ts, 4), dtype=np.uint32)
        assert out.dtype == np.uint32 and out.size == len(t) * ts * ts * 4 and out.flags.c_contiguous
        _check(self.lib.asgpu_path_stream_read_tiles(self.handle, t.ctypes.data if len(t) else None, len(t), out.ctypes.data), "asgpu_path_stream_read_tiles")
        return out.reshape(len(t), ts, ts, 4)

    def scatter_tiles(self, tiles: Sequence[int], pixels: np.ndarray, image: Optional[np.ndarray] = None) -> np.ndarray:
        """Writes tile pixels (image_tiles) into a height x width x 4 image (host side)."""
        h, w, ts = self.config.height, self.config.width, self.config.tile_size
        if image is None:
            image = np.zeros((h, w, 4), dtype=np.uint32)
        tiles_x = (w + ts - 1) // ts
        for k, tile in enumerate(np.asarray(tiles, dtype=np.int64)):
            x0, y0 = int(tile % tiles_x) * ts, int(tile // tiles_x) * ts
            x1, y1 = min(x0 + ts, w), min(y0 + ts, h)
            image[y0:y1, x0:x1] = pixels[k, : y1 - y0, : x1 - x0]
        return image

    def clear(self):
        _check(self.lib.asgpu_path_stream_clear(self.handle), "asgpu_path_stream_clear")

    def stats(self) -> dict:
        v = _lib.PathStreamStats()
        _check(self.lib.asgpu_path_stream_get_stats(self.handle, C.byref(v)), "asgpu_path_stream_get_stats")
        return v.as_dict()

    def capture(self, max_rays: int):
        _check(self.lib.asgpu_path_stream_capture(self.handle, max_rays), "asgpu_path_stream_capture")

    def captured(self) -> List[CapturedWavefront]:
        out = []
        for k in range(int(self.lib.asgpu_path_stream_capture_count(self.handle))):
            kind, depth = C.c_int(0), C.c_uint32(0)
            n = int(self.lib.asgpu_path_stream_capture_get(self.handle, k, C.byref(kind), C.byref(depth), *([None] * 8)))
            if n < 0:
                raise AsgpuError("asgpu_path_stream_capture_get failed: " + _lib.last_error())
            org, dirs = np.empty((n, 3)), np.empty((n, 3))
            tmin, tmax = np.empty(n), np.empty(n)
            flags, ids = np.empty(n, dtype=np.uint32), np.empty(n, dtype=np.uint32)
            res = np.empty(n, dtype=HIT_DTYPE if kind.value == 0 else np.uint8)
            par = np.zeros(n, dtype=PARENT_DTYPE)
            ptr = lambda a: a.ctypes.data if n else None
            self.lib.asgpu_path_stream_capture_get(self.handle, k, None, None, ptr(org), ptr(dirs), ptr(tmin), ptr(tmax), ptr(flags), ptr(ids), ptr(res), ptr(par))
            ta, tn = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
            self.lib.asgpu_path_stream_capture_get_times(self.handle, k, ptr(ta), ptr(tn))
            out.append(CapturedWavefront("closest" if kind.value == 0 else "probe", int(depth.value),
                                         RayBatch(org, dirs, tmin, tmax, ta, tn, flags), ids, res, par))
        return out

    def set_profiling(self, enabled: bool = True):
        """Record CUDA events around every launch of the following render calls."""
        _check(self.lib.asgpu_path_stream_set_profiling(self.handle, 1 if enabled else 0), "asgpu_path_stream_set_profiling")

    def profile(self) -> dict:
        """Device time by kind of launch since the last clear() (needs set_profiling)."""
        v = _lib.PathStreamProfile()
        _check(self.lib.asgpu_path_stream_get_profile(self.handle, C.byref(v)), "asgpu_path_stream_get_profile")
        return v.as_dict()
