"""ctypes loader for the TEST-ONLY host build of the traversal core (tests/hostsim/hostsim.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from appleseed_b200.scene import HIT_DTYPE, CRays, CSceneDesc, RayBatch, SceneDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
SCENE_EXACT, SCENE_WIDE = 1, 2


def load():
    override = os.environ.get("HOSTSIM_LIB")        # e.g. an AddressSanitizer build (tests/fuzz_views.py)
    if not override:
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
    lib = C.CDLL(override or os.path.join(_HERE, "libhostsim.so"))
    lib.hostsim_scene_create.restype = C.c_void_p
    lib.hostsim_scene_create.argtypes = [C.POINTER(CSceneDesc), C.c_uint32, C.c_int]
    lib.hostsim_scene_create_filtered.restype = C.c_void_p
    lib.hostsim_scene_create_filtered.argtypes = [C.POINTER(CSceneDesc), C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hostsim_scene_create_lbvh.restype = C.c_void_p
    lib.hostsim_scene_create_lbvh.argtypes = [C.POINTER(CSceneDesc), C.c_uint32, C.c_int]
    lib.hostsim_scene_create_ploc.restype = C.c_void_p
    lib.hostsim_scene_create_ploc.argtypes = [C.POINTER(CSceneDesc), C.c_uint32, C.c_int, C.c_int]
    lib.hostsim_scene_create_views.restype = C.c_void_p
    lib.hostsim_scene_create_views.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.hostsim_scene_destroy.argtypes = [C.c_void_p]
    lib.hostsim_last_error.restype = C.c_char_p
    lib.hostsim_blob_size.restype = C.c_size_t
    lib.hostsim_blob_size.argtypes = [C.c_void_p]
    lib.hostsim_blob_data.restype = C.c_void_p
    lib.hostsim_blob_data.argtypes = [C.c_void_p]
    lib.hostsim_trace.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
    lib.hostsim_trace_probe.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
    lib.hostsim_trace_parents.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
    lib.hostsim_trace_probe_parents.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
    lib.hostsim_refine_offset.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.hostsim_support_planes.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_void_p, C.c_size_t, C.c_void_p]
    return lib


class SimScene:
    def __init__(self, lib, desc: SceneDesc, flags=SCENE_EXACT | SCENE_WIDE, threads=4, filters=None, lbvh=False, ploc=0):
        """``filters``: {(triangle tree index, object instance index): scene.IntersectionFilter};
        ``lbvh`` / ``ploc`` (= search radius): triangle trees as asgpu_trees_build_on_device makes them
        (topology from a sequential host run of lbvh_core.h / ploc_core.h)."""
        from appleseed_b200.scene import CIntersectionFilter
        self.lib = lib
        self._cdesc, self._keep = desc.to_c()
        if lbvh or ploc:
            assert not filters
            if ploc:
                self.handle = lib.hostsim_scene_create_ploc(C.byref(self._cdesc), flags, threads, int(ploc))
            else:
                self.handle = lib.hostsim_scene_create_lbvh(C.byref(self._cdesc), flags, threads)
            if not self.handle:
                raise RuntimeError(lib.hostsim_last_error().decode())
            return
        items = sorted((filters or {}).items())
        n = len(items)
        trees = np.array([k[0] for k, _ in items], dtype=np.uint32)
        objs = np.array([k[1] for k, _ in items], dtype=np.uint32)
        arr = (CIntersectionFilter * max(1, n))()
        for i, (_, f) in enumerate(items):
            arr[i], keep = f.to_c()
            self._keep.append(keep)
        self.handle = lib.hostsim_scene_create_filtered(C.byref(self._cdesc), flags, threads, n, trees.ctypes.data if n else None,
                                                        objs.ctypes.data if n else None, C.cast(arr, C.c_void_p))
        if not self.handle:
            raise RuntimeError(lib.hostsim_last_error().decode())

    @classmethod
    def from_views(cls, lib, tree_views, top_view, keep, flags=SCENE_EXACT | SCENE_WIDE, sources=None):
        """The product's flattener on caller-supplied reference-format trees (``_lib`` view structs);
        ``sources``: one ``_lib.SourceGeometry`` per tree (needed by refine_offset) or None."""
        from appleseed_b200 import _lib
        self = cls.__new__(cls)
        self.lib = lib
        self._keep = [tree_views, top_view, keep]
        arr = (_lib.TriangleTreeView * max(1, len(tree_views)))(*tree_views)
        self._keep.append(arr)
        src = None
        if sources is not None:
            src = (_lib.SourceGeometry * max(1, len(sources)))(*sources)
            self._keep += [sources, src]
        self.handle = lib.hostsim_scene_create_views(C.cast(arr, C.c_void_p), len(tree_views), C.cast(C.pointer(top_view), C.c_void_p),
                                                     C.cast(src, C.c_void_p) if src is not None else None, flags)
        if not self.handle:
            raise RuntimeError(lib.hostsim_last_error().decode())
        return self

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.hostsim_scene_destroy(self.handle)
            self.handle = None

    def blob(self) -> np.ndarray:
        """A copy of the flattened scene image."""
        n = int(self.lib.hostsim_blob_size(self.handle))
        return np.frombuffer((C.c_uint8 * n).from_address(self.lib.hostsim_blob_data(self.handle)), dtype=np.uint8).copy()

    def trace(self, rays: RayBatch, wide: bool):
        out = np.zeros(len(rays), dtype=HIT_DTYPE)
        cnt = np.zeros(6, dtype=np.uint64)
        cr = rays.to_c()
        self.lib.hostsim_trace(self.handle, C.byref(cr), len(rays), out.ctypes.data, int(wide), cnt.ctypes.data)
        return out, cnt

    def refine_offset(self, rays: RayBatch, hits: np.ndarray) -> np.ndarray:
        from appleseed_b200.scene import PARENT_DTYPE
        out = np.zeros(len(rays), dtype=PARENT_DTYPE)
        cr = rays.to_c()
        hits = np.ascontiguousarray(hits)
        self.lib.hostsim_refine_offset(self.handle, C.byref(cr), hits.ctypes.data, len(rays), out.ctypes.data)
        return out

    def support_planes(self, rays: RayBatch, hits: np.ndarray) -> np.ndarray:
        out = np.zeros((len(rays), 9), dtype=np.float64)
        cr = rays.to_c()
        hits = np.ascontiguousarray(hits)
        self.lib.hostsim_support_planes(self.handle, C.byref(cr), hits.ctypes.data, len(rays), out.ctypes.data)
        return out

    def trace_parents(self, rays: RayBatch, parents: np.ndarray, wide: bool) -> np.ndarray:
        out = np.zeros(len(rays), dtype=HIT_DTYPE)
        cr = rays.to_c()
        parents = np.ascontiguousarray(parents)
        self.lib.hostsim_trace_parents(self.handle, C.byref(cr), parents.ctypes.data, len(rays), out.ctypes.data, int(wide))
        return out

    def trace_probe_parents(self, rays: RayBatch, parents: np.ndarray, wide: bool) -> np.ndarray:
        out = np.zeros(len(rays), dtype=np.uint8)
        cr = rays.to_c()
        parents = np.ascontiguousarray(parents)
        self.lib.hostsim_trace_probe_parents(self.handle, C.byref(cr), parents.ctypes.data, len(rays), out.ctypes.data, int(wide))
        return out

    def trace_probe(self, rays: RayBatch, wide: bool):
        out = np.zeros(len(rays), dtype=np.uint8)
        cnt = np.zeros(6, dtype=np.uint64)
        cr = rays.to_c()
        self.lib.hostsim_trace_probe(self.handle, C.byref(cr), len(rays), out.ctypes.data, int(wide), cnt.ctypes.data)
        return out, cnt
