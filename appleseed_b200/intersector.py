"""Host-side mirror of the reference interface for the path, over the C ABI.

Names follow the reference (renderer/kernel/intersection):

* ``TraceContext``  <-> ``renderer::TraceContext`` (tracecontext.h:52): owns the acceleration
  structures of a scene; ``update()`` there == construction here (build the reference-format
  trees, flatten them, upload the blob).  Immutable afterwards, shareable between intersectors.
* ``Intersector``   <-> ``renderer::Intersector`` (intersector.h:66-145): ``trace`` (closest hit)
  and ``trace_probe`` (any hit), but on BATCHES of rays (wavefront queues) instead of one ray per
  call.  Results are ``asgpu_hit`` records (``scene.HIT_DTYPE``) mirroring the ShadingPoint
  primary block.

torch is used only for device memory, streams and (elsewhere) torch.distributed; every compute
call goes through ``libasgpu.so``.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from .scene import HIT_DTYPE, PARENT_DTYPE, CRays, RayBatch, SceneDesc

HIT_BYTES = HIT_DTYPE.itemsize


class AsgpuError(RuntimeError):
    pass


def _check(rc: int, what: str):
    if rc != 0:
        raise AsgpuError("%s failed (%d): %s" % (what, rc, _lib.last_error()))


class HostTrees:
    """Reference-format trees: the reference's sweep SAH on the host (``asgpu_trees_build``), or,
    with ``build_device`` = a CUDA device ordinal, triangle-tree topology built on that device as a
    linear BVH (``asgpu_trees_build_on_device``)."""

    def __init__(self, desc: SceneDesc, threads: int = 0, build_device: Optional[int] = None, keys=None):
        """``keys``: {assembly instance index: scene.InstanceKeys} -- animated assembly instances
        (``asgpu_trees_build_animated``: motion bounding boxes and interpolator segments as the
        reference's TransformSequence computes them)."""
        self.lib = _lib.load()
        self._cdesc, self._keep = desc.to_c()
        if keys:
            assert build_device is None, "animated instances are built by the host builder"
            from .scene import CInstanceKeys
            arr = (CInstanceKeys * len(desc.assembly_instances))()
            for i, k in keys.items():
                arr[i].times, arr[i].local_to_parent, arr[i].parent_to_local = k.times.ctypes.data, k.local_to_parent.ctypes.data, k.parent_to_local.ctypes.data
                arr[i].key_count = len(k.times)
            self._keep += [arr, keys]
            self.handle = self.lib.asgpu_trees_build_animated(C.byref(self._cdesc), C.cast(arr, C.c_void_p), threads)
        elif build_device is None:
            self.handle = self.lib.asgpu_trees_build(C.byref(self._cdesc), threads)
        else:
            self.handle = self.lib.asgpu_trees_build_on_device(C.byref(self._cdesc), threads, int(build_device))
        if not self.handle:
            raise AsgpuError("asgpu_trees_build failed: " + _lib.last_error())

    def close(self):
        if getattr(self, "handle", None):
            self.lib.asgpu_trees_destroy(self.handle)
            self.handle = None

    __del__ = close

    @property
    def build_seconds(self) -> float:
        return float(self.lib.asgpu_trees_build_seconds(self.handle))

    @property
    def device_seconds(self) -> float:
        return float(self.lib.asgpu_trees_device_seconds(self.handle))

    @property
    def triangle_tree_count(self) -> int:
        return int(self.lib.asgpu_trees_triangle_tree_count(self.handle))

    def triangle_tree_view(self, i: int) -> "_lib.TriangleTreeView":
        v = _lib.TriangleTreeView()
        _check(self.lib.asgpu_trees_get_triangle_tree(self.handle, i, C.byref(v)), "asgpu_trees_get_triangle_tree")
        return v

    def source_geometry(self, i: int) -> "_lib.SourceGeometry":
        v = _lib.SourceGeometry()
        _check(self.lib.asgpu_trees_get_source_geometry(self.handle, i, C.byref(v)), "asgpu_trees_get_source_geometry")
        return v

    def assembly_tree_view(self) -> "_lib.AssemblyTreeView":
        v = _lib.AssemblyTreeView()
        _check(self.lib.asgpu_trees_get_assembly_tree(self.handle, C.byref(v)), "asgpu_trees_get_assembly_tree")
        return v

    @staticmethod
    def _bytes(ptr, n):
        if not ptr or n == 0:
            return np.zeros(0, dtype=np.uint8)
        return np.frombuffer((C.c_uint8 * n).from_address(ptr), dtype=np.uint8).copy()

    def triangle_tree(self, i: int) -> dict:
        v = self.triangle_tree_view(i)
        return {
            "nodes": self._bytes(v.nodes, v.node_count * 128),
            "node_bboxes": self._bytes(v.node_bboxes, v.node_bbox_count * 48).view(np.float64),
            "leaf_data": self._bytes(v.leaf_data, v.leaf_data_size),
            "triangle_keys": self._bytes(v.triangle_keys, v.triangle_key_count * 12),
            "static_triangle_count": int(v.static_triangle_count),
            "moving_triangle_count": int(v.moving_triangle_count),
        }

    def assembly_tree(self) -> dict:
        v = self.assembly_tree_view()
        n = int(v.item_count)
        items = [v.items[i] for i in range(n)]
        return {
            "nodes": self._bytes(v.nodes, v.node_count * 128),
            "item_assembly_instance": np.array([it.assembly_instance for it in items], dtype=np.uint32),
            "item_tree": np.array([it.triangle_tree for it in items], dtype=np.uint32),
        }


@dataclass
class DeviceRays:
    """A ray batch resident in HBM (torch CUDA tensors, one per ShadingRay field)."""
    org: "torch.Tensor"
    dir: "torch.Tensor"
    tmin: "torch.Tensor"
    tmax: "torch.Tensor"
    time_absolute: Optional["torch.Tensor"] = None
    time_normalized: Optional["torch.Tensor"] = None
    flags: Optional["torch.Tensor"] = None

    def __len__(self):
        return int(self.tmin.shape[0])

    @staticmethod
    def from_host(rays: RayBatch, device, non_blocking: bool = False) -> "DeviceRays":
        import torch

        def up(a, dt):
            if a is None:
                return None
            t = torch.from_numpy(a)
            if dt is torch.uint32:
                t = torch.from_numpy(a.view(np.int32))
            return t.to(device, non_blocking=non_blocking)

        return DeviceRays(up(rays.org, None), up(rays.dir, None), up(rays.tmin, None), up(rays.tmax, None),
                          up(rays.time_absolute, None), up(rays.time_normalized, None), up(rays.flags, torch.uint32))

    def to_c(self) -> CRays:
        r = CRays()
        r.org = self.org.data_ptr()
        r.dir = self.dir.data_ptr()
        r.tmin = self.tmin.data_ptr()
        r.tmax = self.tmax.data_ptr()
        r.time_absolute = self.time_absolute.data_ptr() if self.time_absolute is not None else None
        r.time_normalized = self.time_normalized.data_ptr() if self.time_normalized is not None else None
        r.flags = self.flags.data_ptr() if self.flags is not None else None
        return r

    @property
    def bytes_per_ray(self) -> int:
        b = 64
        for t in (self.time_absolute, self.time_normalized, self.flags):
            b += 4 if t is not None else 0
        return b


class TraceContext:
    """Owns the flattened scene on one GPU."""

    def __init__(self, desc: Optional[SceneDesc] = None, device: int = 0, flags: int = _lib.SCENE_DEFAULT,
                 threads: int = 0, trees: Optional[HostTrees] = None, _handle=None, source_geometry: bool = True,
                 filters: Optional[dict] = None):
        """``filters``: {(triangle tree index, object instance index): scene.IntersectionFilter} --
        ``TriangleTree::m_intersection_filters`` (cut-out geometry, closest hit only)."""
        self.lib = _lib.load()
        self.device = device
        self._borrowed_blob = None
        if _handle is not None:
            self.handle = _handle
            self.build_seconds = 0.0
        else:
            own = trees is None
            if own:
                if desc is None:
                    raise ValueError("TraceContext needs a SceneDesc or HostTrees")
                trees = HostTrees(desc, threads)
            n = trees.triangle_tree_count
            views = (_lib.TriangleTreeView * max(1, n))()
            sources = (_lib.SourceGeometry * max(1, n))()
            keep = []
            for i in range(n):
                views[i] = trees.triangle_tree_view(i)
                sources[i] = trees.source_geometry(i)
                mine = {o: f for (t, o), f in (filters or {}).items() if t == i}
                if mine:
                    arr = (_lib.CIntersectionFilter * max(1, sources[i].object_count))()
                    for o, f in mine.items():
                        arr[o], k = f.to_c()
                        keep.append(k)
                    keep.append(arr)
                    sources[i].filters = C.cast(arr, C.c_void_p)
            top = trees.assembly_tree_view()
            self.handle = self.lib.asgpu_scene_create_ex(views, n, C.byref(top), sources if source_geometry else None, flags, device)
            self.build_seconds = trees.build_seconds
            if own:
                trees.close()
        if not self.handle:
            raise AsgpuError("scene creation failed: " + _lib.last_error())

    @classmethod
    def from_tree_views(cls, tree_views, top_view, device: int = 0, flags: int = _lib.SCENE_DEFAULT, sources=None) -> "TraceContext":
        """Flatten reference-format trees supplied by the caller (``asgpu_scene_create``; with
        ``sources``, one ``_lib.SourceGeometry`` per tree: ``asgpu_scene_create_ex``)."""
        lib = _lib.load()
        n = len(tree_views)
        arr = (_lib.TriangleTreeView * max(1, n))(*tree_views)
        if sources is None:
            handle = lib.asgpu_scene_create(arr, n, C.byref(top_view), flags, device)
        else:
            src = (_lib.SourceGeometry * max(1, n))(*sources)
            handle = lib.asgpu_scene_create_ex(arr, n, C.byref(top_view), src, flags, device)
        if not handle:
            raise AsgpuError("asgpu_scene_create failed: " + _lib.last_error())
        return cls(device=device, _handle=handle)

    @classmethod
    def from_blob(cls, blob: "torch.Tensor", adopt: bool = True) -> "TraceContext":
        """Adopt a blob received from another rank (uint8 CUDA tensor)."""
        lib = _lib.load()
        device = blob.device.index
        handle = lib.asgpu_scene_import_blob(blob.data_ptr(), blob.numel(), device, 1 if adopt else 0)
        if not handle:
            raise AsgpuError("asgpu_scene_import_blob failed: " + _lib.last_error())
        ctx = cls(device=device, _handle=handle)
        if adopt:
            ctx._borrowed_blob = blob      # keep the memory alive
        return ctx

    def close(self):
        if getattr(self, "handle", None):
            self.lib.asgpu_scene_destroy(self.handle)
            self.handle = None

    __del__ = close

    @property
    def blob_size(self) -> int:
        return int(self.lib.asgpu_scene_blob_size(self.handle))

    def blob_tensor(self) -> "torch.Tensor":
        """A uint8 CUDA tensor holding a copy of the scene blob (the broadcast payload)."""
        import torch
        n = self.blob_size
        out = torch.empty(n, dtype=torch.uint8, device="cuda:%d" % self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _check(self.lib.asgpu_scene_export_blob(self.handle, out.data_ptr(), n, C.c_void_p(stream)), "asgpu_scene_export_blob")
        return out

    def info(self) -> dict:
        v = _lib.SceneInfo()
        _check(self.lib.asgpu_scene_get_info(self.handle, C.byref(v)), "asgpu_scene_get_info")
        return v.as_dict()

    def counters(self, reset: bool = False) -> dict:
        v = _lib.Counters()
        _check(self.lib.asgpu_get_counters(self.handle, C.byref(v), 1 if reset else 0), "asgpu_get_counters")
        return v.as_dict()

    def lane_profile(self):
        """(closest-hit, any-hit) ``asgpu_lane_profile`` dicts of the counters launches so far."""
        a, b = _lib.LaneProfile(), _lib.LaneProfile()
        _check(self.lib.asgpu_get_lane_profile(self.handle, C.byref(a), C.byref(b)), "asgpu_get_lane_profile")
        return a.as_dict(), b.as_dict()

    def counters_by_kind(self, reset: bool = False):
        """(closest-hit counters, any-hit counters) -- the same traversal statistics kept apart."""
        a, b = _lib.Counters(), _lib.Counters()
        _check(self.lib.asgpu_get_counters_by_kind(self.handle, C.byref(a), C.byref(b), 1 if reset else 0), "asgpu_get_counters_by_kind")
        return a.as_dict(), b.as_dict()


class Intersector:
    """Batched ``trace`` / ``trace_probe`` on a ``TraceContext``."""

    def __init__(self, trace_context: TraceContext):
        self.ctx = trace_context
        self.lib = trace_context.lib

    # -- host buffers (copies inside the call) ---------------------------------------------------

    @staticmethod
    def _flags(exact: bool, counters: bool, sort: bool) -> int:
        return (_lib.TRACE_EXACT if exact else 0) | (_lib.TRACE_COUNTERS if counters else 0) | (_lib.TRACE_SORT if sort else 0)

    def trace(self, rays: RayBatch, exact: bool = False, counters: bool = False, out: Optional[np.ndarray] = None, sort: bool = False) -> np.ndarray:
        n = len(rays)
        hits = np.empty(n, dtype=HIT_DTYPE) if out is None else out
        cr = rays.to_c()
        flags = self._flags(exact, counters, sort)
        _check(self.lib.asgpu_trace_host(self.ctx.handle, C.byref(cr), n, hits.ctypes.data if n else None, flags), "asgpu_trace_host")
        return hits

    def trace_probe(self, rays: RayBatch, exact: bool = False, counters: bool = False, out: Optional[np.ndarray] = None, sort: bool = False) -> np.ndarray:
        n = len(rays)
        occ = np.empty(n, dtype=np.uint8) if out is None else out
        cr = rays.to_c()
        flags = self._flags(exact, counters, sort)
        _check(self.lib.asgpu_trace_probe_host(self.ctx.handle, C.byref(cr), n, occ.ctypes.data if n else None, flags), "asgpu_trace_probe_host")
        return occ

    # -- device buffers (no copies; enqueued on torch's current stream) --------------------------

    def trace_device(self, rays: DeviceRays, hits: "torch.Tensor", exact: bool = False, counters: bool = False, sort: bool = False):
        """``hits``: uint8 CUDA tensor of n * 40 bytes receiving ``asgpu_hit`` records."""
        import torch
        n = len(rays)
        assert hits.numel() * hits.element_size() >= n * HIT_BYTES
        cr = rays.to_c()
        flags = self._flags(exact, counters, sort)
        stream = torch.cuda.current_stream(self.ctx.device).cuda_stream
        _check(self.lib.asgpu_trace(self.ctx.handle, C.byref(cr), n, hits.data_ptr(), flags, C.c_void_p(stream)), "asgpu_trace")

    def trace_probe_device(self, rays: DeviceRays, occluded: "torch.Tensor", exact: bool = False, counters: bool = False, sort: bool = False):
        import torch
        n = len(rays)
        assert occluded.numel() >= n
        cr = rays.to_c()
        flags = self._flags(exact, counters, sort)
        stream = torch.cuda.current_stream(self.ctx.device).cuda_stream
        _check(self.lib.asgpu_trace_probe(self.ctx.handle, C.byref(cr), n, occluded.data_ptr(), flags, C.c_void_p(stream)), "asgpu_trace_probe")


    # -- parent shading points (refine_and_offset + the parent origin rule) ----------------------

    def refine_and_offset_device(self, rays: DeviceRays, hits: "torch.Tensor", parents: "torch.Tensor"):
        """``parents``: uint8 CUDA tensor of n * 80 bytes receiving ``asgpu_parent`` records."""
        import torch
        n = len(rays)
        assert parents.numel() * parents.element_size() >= n * PARENT_DTYPE.itemsize
        cr = rays.to_c()
        stream = torch.cuda.current_stream(self.ctx.device).cuda_stream
        _check(self.lib.asgpu_refine_and_offset(self.ctx.handle, C.byref(cr), hits.data_ptr(), n, parents.data_ptr(), C.c_void_p(stream)),
               "asgpu_refine_and_offset")

    def trace_with_parents_device(self, rays: DeviceRays, parents: "torch.Tensor", hits: "torch.Tensor", exact: bool = False):
        import torch
        cr = rays.to_c()
        stream = torch.cuda.current_stream(self.ctx.device).cuda_stream
        _check(self.lib.asgpu_trace_with_parents(self.ctx.handle, C.byref(cr), parents.data_ptr(), len(rays), hits.data_ptr(),
                                                 self._flags(exact, False, False), C.c_void_p(stream)), "asgpu_trace_with_parents")

    def trace_probe_with_parents_device(self, rays: DeviceRays, parents: "torch.Tensor", occluded: "torch.Tensor", exact: bool = False):
        import torch
        cr = rays.to_c()
        stream = torch.cuda.current_stream(self.ctx.device).cuda_stream
        _check(self.lib.asgpu_trace_probe_with_parents(self.ctx.handle, C.byref(cr), parents.data_ptr(), len(rays), occluded.data_ptr(),
                                                       self._flags(exact, False, False), C.c_void_p(stream)), "asgpu_trace_probe_with_parents")

    def support_planes_device(self, rays: DeviceRays, hits: "torch.Tensor", planes: "torch.Tensor"):
        """``planes``: float64 CUDA tensor of n * 9 values receiving ``m_triangle_support_plane``
        (v0, e0, e1) of every hit (``asgpu_get_support_planes``)."""
        import torch
        n = len(rays)
        assert planes.dtype == torch.float64 and planes.numel() >= n * 9
        cr = rays.to_c()
        stream = torch.cuda.current_stream(self.ctx.device).cuda_stream
        _check(self.lib.asgpu_get_support_planes(self.ctx.handle, C.byref(cr), hits.data_ptr(), n, planes.data_ptr(), C.c_void_p(stream)),
               "asgpu_get_support_planes")

    def support_planes(self, rays: RayBatch, hits: np.ndarray) -> np.ndarray:
        import torch
        dev = "cuda:%d" % self.ctx.device
        d = DeviceRays.from_host(rays, dev)
        h = torch.from_numpy(np.ascontiguousarray(hits).view(np.uint8)).to(dev)
        out = torch.empty(max(1, len(rays)) * 9, dtype=torch.float64, device=dev)
        self.support_planes_device(d, h, out)
        torch.cuda.synchronize()
        return out.cpu().numpy()[: len(rays) * 9].reshape(-1, 9).copy()

    # Host-array conveniences over the three calls above (upload, run, download).
    def refine_and_offset(self, rays: RayBatch, hits: np.ndarray) -> np.ndarray:
        import torch
        dev = "cuda:%d" % self.ctx.device
        d = DeviceRays.from_host(rays, dev)
        h = torch.from_numpy(np.ascontiguousarray(hits).view(np.uint8)).to(dev)
        out = torch.empty(max(1, len(rays)) * PARENT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        self.refine_and_offset_device(d, h, out)
        torch.cuda.synchronize()
        return out.cpu().numpy()[: len(rays) * PARENT_DTYPE.itemsize].view(PARENT_DTYPE).copy()

    def trace_with_parents(self, rays: RayBatch, parents: np.ndarray, exact: bool = False) -> np.ndarray:
        import torch
        dev = "cuda:%d" % self.ctx.device
        d = DeviceRays.from_host(rays, dev)
        p = torch.from_numpy(np.ascontiguousarray(parents).view(np.uint8)).to(dev)
        out = torch.empty(max(1, len(rays)) * HIT_BYTES, dtype=torch.uint8, device=dev)
        self.trace_with_parents_device(d, p, out, exact=exact)
        torch.cuda.synchronize()
        return hits_from_tensor(out, len(rays))

    def trace_probe_with_parents(self, rays: RayBatch, parents: np.ndarray, exact: bool = False) -> np.ndarray:
        import torch
        dev = "cuda:%d" % self.ctx.device
        d = DeviceRays.from_host(rays, dev)
        p = torch.from_numpy(np.ascontiguousarray(parents).view(np.uint8)).to(dev)
        out = torch.empty(max(1, len(rays)), dtype=torch.uint8, device=dev)
        self.trace_probe_with_parents_device(d, p, out, exact=exact)
        torch.cuda.synchronize()
        return out.cpu().numpy()[: len(rays)].copy()

    def sort_rays(self, rays: DeviceRays):
        """``asgpu_sort_rays``: (order, keys) int32 CUDA tensors -- the permutation by ascending
        origin / direction Morton key and the sorted keys."""
        import torch
        n = len(rays)
        dev = "cuda:%d" % self.ctx.device
        order = torch.empty(n, dtype=torch.int32, device=dev)
        keys = torch.empty(n, dtype=torch.int32, device=dev)
        cr = rays.to_c()
        stream = torch.cuda.current_stream(self.ctx.device).cuda_stream
        _check(self.lib.asgpu_sort_rays(self.ctx.handle, C.byref(cr), n, order.data_ptr(), keys.data_ptr(), C.c_void_p(stream)), "asgpu_sort_rays")
        return order, keys


def hits_from_tensor(t: "torch.Tensor", n: int) -> np.ndarray:
    """View a device hit buffer as ``HIT_DTYPE`` records on the host."""
    return t.cpu().numpy().view(np.uint8)[: n * HIT_BYTES].view(HIT_DTYPE).copy()
