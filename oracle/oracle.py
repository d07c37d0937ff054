"""ctypes front-end for the two CPU checkers (oracle_api.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py -- never by appleseed_b200/.

``Oracle("orc")``   -> oracle/liboracle.so      (self-contained restatement, "port")
``Oracle("asref")`` -> oracle/_ref/libasref.so  (the reference's own headers, "reference")
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

from appleseed_b200.scene import HIT_DTYPE, PARENT_DTYPE, CInstanceKeys, CItemMotion, CRays, CSceneDesc, RayBatch, SceneDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "orc": os.path.join(_HERE, "liboracle.so"),
    "asref": os.path.join(_HERE, "_ref", "libasref.so"),
}


class TriangleTreeView(C.Structure):
    _fields_ = [
        ("nodes", C.c_void_p),
        ("node_bboxes", C.c_void_p),
        ("leaf_data", C.c_void_p),
        ("triangle_keys", C.c_void_p),
        ("node_count", C.c_uint64),
        ("node_bbox_count", C.c_uint64),
        ("leaf_data_size", C.c_uint64),
        ("triangle_key_count", C.c_uint64),
        ("static_triangle_count", C.c_uint64),
        ("moving_triangle_count", C.c_uint64),
    ]


class AssemblyTreeView(C.Structure):
    _fields_ = [
        ("nodes", C.c_void_p),
        ("item_assembly_instance", C.c_void_p),
        ("item_tree", C.c_void_p),
        ("node_count", C.c_uint64),
        ("item_count", C.c_uint64),
    ]


class Counters(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64),
        ("assembly_nodes_visited", C.c_uint64),
        ("instances_visited", C.c_uint64),
        ("triangle_nodes_visited", C.c_uint64),
        ("triangles_tested", C.c_uint64),
        ("hits", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def build(which: str = "all") -> None:
    """Compile the checkers (``make -C oracle``).  ``_ref`` is only rebuilt where the
    reference tree exists; elsewhere the prebuilt library that travelled is used."""
    target = {"all": "all", "orc": "liboracle.so", "asref": "ref"}[which]
    subprocess.run(["make", "-s", "-C", _HERE, target], check=True, stdout=sys.stderr)      # keep stdout for bench.py's JSON line


def available(prefix: str) -> bool:
    return os.path.exists(_PATHS[prefix])


def _view_bytes(ptr, nbytes):
    if not ptr or nbytes == 0:
        return np.zeros(0, dtype=np.uint8)
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=np.uint8).copy()


class Oracle:
    def __init__(self, prefix: str = "orc", path: str = None):
        """``path``: another library exporting the same interface (tests/adaptor's build of ref_driver.cpp)."""
        if prefix not in _PATHS:
            raise ValueError(prefix)
        path = path or _PATHS[prefix]
        if not os.path.exists(path):
            if prefix == "orc":
                build("orc")
            else:
                raise FileNotFoundError(path)
        self.prefix = prefix
        self.lib = C.CDLL(path)
        L, p = self.lib, prefix
        self._create = getattr(L, p + "_scene_create")
        self._create.restype = C.c_void_p
        self._create.argtypes = [C.POINTER(CSceneDesc)]
        self._destroy = getattr(L, p + "_scene_destroy")
        self._destroy.argtypes = [C.c_void_p]
        self._tree_count = getattr(L, p + "_tree_count")
        self._tree_count.argtypes = [C.c_void_p]
        self._tree_count.restype = C.c_int
        self._get_tt = getattr(L, p + "_get_triangle_tree")
        self._get_tt.argtypes = [C.c_void_p, C.c_int, C.POINTER(TriangleTreeView)]
        self._get_at = getattr(L, p + "_get_assembly_tree")
        self._get_at.argtypes = [C.c_void_p, C.POINTER(AssemblyTreeView)]
        self._trace = getattr(L, p + "_trace")
        self._trace.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_size_t, C.c_void_p, C.c_int, C.POINTER(Counters)]
        self._probe = getattr(L, p + "_trace_probe")
        self._probe.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_size_t, C.c_void_p, C.c_int, C.POINTER(Counters)]
        self._refine = getattr(L, p + "_refine_offset")
        self._refine.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        self._planes = getattr(L, p + "_support_planes")
        self._planes.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        self._trace_par = getattr(L, p + "_trace_parents")
        self._trace_par.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        self._probe_par = getattr(L, p + "_trace_probe_parents")
        self._probe_par.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        self._set_filter = getattr(L, p + "_set_filter")
        self._set_filter.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        if prefix == "asref":
            self._create_animated = L.asref_scene_create_animated
            self._create_animated.restype = C.c_void_p
            self._create_animated.argtypes = [C.POINTER(CSceneDesc), C.c_void_p]
            self._item_motion = L.asref_get_item_motion
            self._item_motion.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(CItemMotion)]
            self._item_p2l = L.asref_get_item_parent_to_local
            self._item_p2l.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
            self._trace_planes = L.asref_trace_planes
            self._trace_planes.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        if prefix == "orc":
            self._two = L.orc_two_nearest
            self._two.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        for name, argt, rest in (
            ("_kat_ray_triangle", [C.c_void_p] * 5 + [C.c_double, C.c_double, C.c_void_p], C.c_int),
            ("_kat_ray_triangle_bool", [C.c_void_p] * 5 + [C.c_double, C.c_double], C.c_int),
            ("_kat_ray_aabb", [C.c_void_p] * 4 + [C.c_double, C.c_double, C.c_void_p], C.c_int),
            ("_kat_ray_aabb_ex", [C.c_int] + [C.c_void_p] * 4 + [C.c_double, C.c_double, C.c_void_p], C.c_int),
            ("_kat_ray_info", [C.c_void_p] * 3, None),
        ):
            f = getattr(L, p + name)
            f.argtypes = argt
            f.restype = rest

    # -- scenes -------------------------------------------------------------------------------

    def scene(self, desc: SceneDesc, keys=None) -> "OracleScene":
        """``keys``: {assembly instance index: scene.InstanceKeys} -- animated assembly instances
        (asref only: it links the reference's own TransformSequence)."""
        return OracleScene(self, desc, keys)

    # -- known-answer entry points ------------------------------------------------------------

    @staticmethod
    def _v(a):
        return np.ascontiguousarray(a, dtype=np.float64)

    def kat_ray_triangle(self, v0, v1, v2, org, dir, tmin=0.0, tmax=np.finfo(np.float64).max):
        a = [self._v(x) for x in (v0, v1, v2, org, dir)]
        tuv = np.zeros(3)
        hit = getattr(self.lib, self.prefix + "_kat_ray_triangle")(
            *[x.ctypes.data for x in a], tmin, tmax, tuv.ctypes.data)
        return bool(hit), tuv

    def kat_ray_triangle_bool(self, v0, v1, v2, org, dir, tmin=0.0, tmax=np.finfo(np.float64).max):
        a = [self._v(x) for x in (v0, v1, v2, org, dir)]
        return bool(getattr(self.lib, self.prefix + "_kat_ray_triangle_bool")(
            *[x.ctypes.data for x in a], tmin, tmax))

    def kat_ray_aabb(self, bmin, bmax, org, dir, tmin=0.0, tmax=np.finfo(np.float64).max):
        a = [self._v(x) for x in (bmin, bmax, org, dir)]
        t = np.zeros(1)
        hit = getattr(self.lib, self.prefix + "_kat_ray_aabb")(*[x.ctypes.data for x in a], tmin, tmax, t.ctypes.data)
        return bool(hit), float(t[0])

    def kat_ray_aabb_ex(self, mode, bmin, bmax, org, dir, tmin, tmax, io):
        """rayaabb.h: mode 0 = intersect(ray, info, bbox), 1 = intersect(..., distance) with io = [distance],
        2 = clip(ray, ...) with io = [ray tmin, ray tmax]; io is updated in place; returns the bool."""
        a = [self._v(x) for x in (bmin, bmax, org, dir)]
        buf = np.array(io, dtype=np.float64)
        hit = getattr(self.lib, self.prefix + "_kat_ray_aabb_ex")(mode, *[x.ctypes.data for x in a], tmin, tmax, buf.ctypes.data)
        return bool(hit), buf

    def kat_ray_info(self, dir):
        d = self._v(dir)
        rcp = np.zeros(3)
        sgn = np.zeros(3, dtype=np.uint32)
        getattr(self.lib, self.prefix + "_kat_ray_info")(d.ctypes.data, rcp.ctypes.data, sgn.ctypes.data)
        return rcp, sgn


class OracleScene:
    def __init__(self, oracle: Oracle, desc: SceneDesc, keys=None):
        self.oracle = oracle
        self.desc = desc
        self._cdesc, self._keep = desc.to_c()
        if keys:
            if oracle.prefix != "asref":
                raise ValueError("animated assembly instances are checked against the reference-header build (asref) only")
            arr = (CInstanceKeys * len(desc.assembly_instances))()
            for i, k in keys.items():
                arr[i].times, arr[i].local_to_parent, arr[i].parent_to_local = k.times.ctypes.data, k.local_to_parent.ctypes.data, k.parent_to_local.ctypes.data
                arr[i].key_count = len(k.times)
            self._keep += [arr, keys]
            self.handle = oracle._create_animated(C.byref(self._cdesc), C.cast(arr, C.c_void_p))
        else:
            self.handle = oracle._create(C.byref(self._cdesc))
        if not self.handle:
            raise RuntimeError("oracle scene_create failed")

    def close(self):
        if self.handle:
            self.oracle._destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def tree_count(self) -> int:
        return self.oracle._tree_count(self.handle)

    def triangle_tree(self, index: int) -> dict:
        """Copy of one reference-format triangle tree (the bytes an in-tree flattener sees)."""
        v = TriangleTreeView()
        self.oracle._get_tt(self.handle, index, C.byref(v))
        return {
            "nodes": _view_bytes(v.nodes, v.node_count * 128),
            "node_bboxes": _view_bytes(v.node_bboxes, v.node_bbox_count * 48).view(np.float64),
            "leaf_data": _view_bytes(v.leaf_data, v.leaf_data_size),
            "triangle_keys": _view_bytes(v.triangle_keys, v.triangle_key_count * 12),
            "static_triangle_count": int(v.static_triangle_count),
            "moving_triangle_count": int(v.moving_triangle_count),
        }

    def assembly_tree(self) -> dict:
        v = AssemblyTreeView()
        self.oracle._get_at(self.handle, C.byref(v))
        return {
            "nodes": _view_bytes(v.nodes, v.node_count * 128),
            "item_assembly_instance": _view_bytes(v.item_assembly_instance, v.item_count * 4).view(np.uint32),
            "item_tree": _view_bytes(v.item_tree, v.item_count * 4).view(np.uint32),
        }

    def item_motion(self, item: int):
        """(key times, key parent_to_local [k, 16], segments [k - 1, 20]) of tree-order item ``item`` or None."""
        m = CItemMotion()
        self.oracle._item_motion(self.handle, item, C.byref(m))
        k = int(m.key_count)
        if k < 2:
            return None
        return (_view_bytes(m.key_times, k * 4).view(np.float32).copy(), _view_bytes(m.key_parent_to_local, k * 128).view(np.float64).reshape(k, 16).copy(),
                _view_bytes(m.segments, (k - 1) * 160).view(np.float64).reshape(k - 1, 20).copy())

    def item_parent_to_local(self, item: int) -> np.ndarray:
        out = np.zeros(16)
        self.oracle._item_p2l(self.handle, item, out.ctypes.data)
        return out

    def trace(self, rays: RayBatch, threads: int = 1, counters: bool = False):
        n = len(rays)
        out = np.zeros(n, dtype=HIT_DTYPE)
        cr = rays.to_c()
        cnt = Counters()
        self.oracle._trace(self.handle, C.byref(cr), n, out.ctypes.data, threads, C.byref(cnt))
        return (out, cnt.as_dict()) if counters else out

    def trace_probe(self, rays: RayBatch, threads: int = 1, counters: bool = False):
        n = len(rays)
        out = np.zeros(n, dtype=np.uint8)
        cr = rays.to_c()
        cnt = Counters()
        self.oracle._probe(self.handle, C.byref(cr), n, out.ctypes.data, threads, C.byref(cnt))
        return (out, cnt.as_dict()) if counters else out

    def refine_offset(self, rays: RayBatch, hits: np.ndarray, threads: int = 1) -> np.ndarray:
        """ShadingPoint::refine_and_offset for every hit: PARENT_DTYPE records (orc_parent)."""
        n = len(rays)
        out = np.zeros(n, dtype=PARENT_DTYPE)
        cr = rays.to_c()
        hits = np.ascontiguousarray(hits)
        self.oracle._refine(self.handle, C.byref(cr), hits.ctypes.data, n, out.ctypes.data, threads)
        return out

    def support_planes(self, rays: RayBatch, hits: np.ndarray, threads: int = 1) -> np.ndarray:
        """m_triangle_support_plane (v0, e0, e1) of every hit, recomputed from leaf slot + ray time."""
        out = np.zeros((len(rays), 9), dtype=np.float64)
        cr = rays.to_c()
        hits = np.ascontiguousarray(hits)
        self.oracle._planes(self.handle, C.byref(cr), hits.ctypes.data, len(rays), out.ctypes.data, threads)
        return out

    def trace_planes(self, rays: RayBatch, threads: int = 1):
        """(hits, planes): Intersector::trace with the support plane the traversal itself stored in the
        ShadingPoint (oracle/_ref only)."""
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        planes = np.zeros((len(rays), 9), dtype=np.float64)
        cr = rays.to_c()
        self.oracle._trace_planes(self.handle, C.byref(cr), len(rays), hits.ctypes.data, planes.ctypes.data, threads)
        return hits, planes

    def trace_parents(self, rays: RayBatch, parents: np.ndarray, threads: int = 1) -> np.ndarray:
        n = len(rays)
        out = np.zeros(n, dtype=HIT_DTYPE)
        cr = rays.to_c()
        parents = np.ascontiguousarray(parents)
        self.oracle._trace_par(self.handle, C.byref(cr), parents.ctypes.data, n, out.ctypes.data, threads)
        return out

    def trace_probe_parents(self, rays: RayBatch, parents: np.ndarray, threads: int = 1) -> np.ndarray:
        n = len(rays)
        out = np.zeros(n, dtype=np.uint8)
        cr = rays.to_c()
        parents = np.ascontiguousarray(parents)
        self.oracle._probe_par(self.handle, C.byref(cr), parents.ctypes.data, n, out.ctypes.data, threads)
        return out

    def set_filter(self, assembly: int, object_instance: int, flt) -> None:
        """Attach an ``appleseed_b200.scene.IntersectionFilter`` to one object instance (copied)."""
        c, keep = flt.to_c()
        self.oracle._set_filter(self.handle, assembly, object_instance, C.byref(c))

    def two_nearest(self, rays: RayBatch, threads: int = 1):
        n = len(rays)
        t1 = np.zeros(n)
        t2 = np.zeros(n)
        cr = rays.to_c()
        self.oracle._two(self.handle, C.byref(cr), n, t1.ctypes.data, t2.ctypes.data, threads)
        return t1, t2
