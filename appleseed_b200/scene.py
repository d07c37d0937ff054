"""Scene description mirrored from appleseed's entities on the Intersector path.

This is host-side marshalling only: plain numpy arrays packed into the C structs of
``include/asgpu.h`` (``asgpu_mesh``, ``asgpu_object_instance``, ``asgpu_assembly``,
``asgpu_assembly_instance``, ``asgpu_scene_desc``).  The entities follow the reference:

* ``Mesh``             <-> ``StaticTriangleTess`` (renderer/kernel/tessellation/statictessellation.h)
* ``ObjectInstance``   <-> ``ObjectInstance``     (renderer/modeling/scene/objectinstance.h)
* ``Assembly``         <-> ``Assembly`` + its ``acceleration_structure`` parameters
                           (renderer/kernel/intersection/triangletree.cpp:404-407,538-540)
* ``AssemblyInstance`` <-> ``AssemblyInstance`` with a single-key ``TransformSequence``
                           (renderer/utility/transformsequence.h:185-210); nested assemblies are
                           flattened by the caller exactly as ``collect_assembly_instances``
                           does (renderer/kernel/intersection/assemblytree.cpp:111-152).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

# renderer/modeling/scene/visibilityflags.h:50-64
VIS_INVISIBLE = 0
VIS_CAMERA = 1 << 0
VIS_LIGHT = 1 << 1
VIS_SHADOW = 1 << 2
VIS_TRANSPARENCY = 1 << 3
VIS_PROBE = 1 << 4
VIS_DIFFUSE = 1 << 5
VIS_GLOSSY = 1 << 6
VIS_SPECULAR = 1 << 7
VIS_SUBSURFACE = 1 << 8
VIS_NPR = 1 << 9
VIS_ALL = 0xFFFFFFFF

MISS = 0xFFFFFFFF


class CMesh(C.Structure):
    _fields_ = [
        ("vertices", C.c_void_p),
        ("triangles", C.c_void_p),
        ("triangle_pa", C.c_void_p),
        ("vertex_poses", C.c_void_p),
        ("vertex_count", C.c_uint32),
        ("triangle_count", C.c_uint32),
        ("motion_segment_count", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class CObjectInstance(C.Structure):
    _fields_ = [
        ("local_to_parent", C.c_double * 16),
        ("parent_to_local", C.c_double * 16),
        ("mesh_index", C.c_uint32),
        ("vis_flags", C.c_uint32),
    ]


class CAssembly(C.Structure):
    _fields_ = [
        ("object_instances", C.POINTER(CObjectInstance)),
        ("object_instance_count", C.c_uint32),
        ("max_leaf_size", C.c_uint32),
        ("interior_node_traversal_cost", C.c_float),
        ("triangle_intersection_cost", C.c_float),
        ("time", C.c_double),
    ]


class CAssemblyInstance(C.Structure):
    _fields_ = [
        ("local_to_parent", C.c_double * 16),
        ("parent_to_local", C.c_double * 16),
        ("assembly_index", C.c_uint32),
        ("vis_flags", C.c_uint32),
    ]


class CSceneDesc(C.Structure):
    _fields_ = [
        ("meshes", C.POINTER(CMesh)),
        ("assemblies", C.POINTER(CAssembly)),
        ("assembly_instances", C.POINTER(CAssemblyInstance)),
        ("mesh_count", C.c_uint32),
        ("assembly_count", C.c_uint32),
        ("assembly_instance_count", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class CRays(C.Structure):
    _fields_ = [
        ("org", C.c_void_p),
        ("dir", C.c_void_p),
        ("tmin", C.c_void_p),
        ("tmax", C.c_void_p),
        ("time_absolute", C.c_void_p),
        ("time_normalized", C.c_void_p),
        ("flags", C.c_void_p),
    ]


# Hit record, identical to asgpu_hit in include/asgpu.h (40 bytes).
HIT_DTYPE = np.dtype(
    [
        ("t", "<f8"),
        ("u", "<f4"),
        ("v", "<f4"),
        ("assembly_instance", "<u4"),
        ("object_instance_index", "<u4"),
        ("primitive_index", "<u4"),
        ("tri_slot", "<u4"),
        ("motion_segment", "<u4"),
        ("prim_type", "<u4"),
    ]
)
assert HIT_DTYPE.itemsize == 40

# asgpu_parent (include/asgpu.h): what Intersector::trace reads of the parent ShadingPoint.
PARENT_DTYPE = np.dtype(
    [("assembly_instance", np.uint32), ("reserved", np.uint32), ("front", np.float64, 3), ("back", np.float64, 3),
     ("geo_normal", np.float64, 3)], align=True)
assert PARENT_DTYPE.itemsize == 80


def _mat(m) -> np.ndarray:
    m = np.ascontiguousarray(np.asarray(m, dtype=np.float64).reshape(4, 4))
    return m


def invert_transform(m: np.ndarray) -> np.ndarray:
    """Inverse of a 4x4 matrix in float64 (Transformd holds both directions,
    foundation/math/transform.h:135-136; the caller supplies both)."""
    return np.linalg.inv(_mat(m))


@dataclass
class Mesh:
    vertices: np.ndarray                      # (nv, 3) float32, object space
    triangles: np.ndarray                     # (nt, 3) uint32
    triangle_pa: Optional[np.ndarray] = None  # (nt,) uint16
    vertex_poses: Optional[np.ndarray] = None  # (nv, msc, 3) float32

    def __post_init__(self):
        self.vertices = np.ascontiguousarray(self.vertices, dtype=np.float32).reshape(-1, 3)
        self.triangles = np.ascontiguousarray(self.triangles, dtype=np.uint32).reshape(-1, 3)
        if self.triangle_pa is not None:
            self.triangle_pa = np.ascontiguousarray(self.triangle_pa, dtype=np.uint16).reshape(-1)
            assert self.triangle_pa.shape[0] == self.triangles.shape[0]
        if self.vertex_poses is not None:
            self.vertex_poses = np.ascontiguousarray(self.vertex_poses, dtype=np.float32)
            assert self.vertex_poses.ndim == 3 and self.vertex_poses.shape[0] == self.vertices.shape[0]
            assert self.vertex_poses.shape[2] == 3

    @property
    def motion_segment_count(self) -> int:
        return 0 if self.vertex_poses is None else int(self.vertex_poses.shape[1])


@dataclass
class ObjectInstance:
    mesh_index: int
    local_to_parent: np.ndarray = field(default_factory=lambda: np.eye(4))
    vis_flags: int = VIS_ALL
    parent_to_local: Optional[np.ndarray] = None

    def __post_init__(self):
        self.local_to_parent = _mat(self.local_to_parent)
        self.parent_to_local = (
            invert_transform(self.local_to_parent) if self.parent_to_local is None else _mat(self.parent_to_local)
        )


@dataclass
class Assembly:
    object_instances: List[ObjectInstance]
    max_leaf_size: int = 2                       # intersectionsettings.h:71
    interior_node_traversal_cost: float = 1.0    # intersectionsettings.h:72
    triangle_intersection_cost: float = 1.0      # intersectionsettings.h:73
    time: float = 0.5                            # triangletree.cpp:406


@dataclass
class AssemblyInstance:
    assembly_index: int
    local_to_parent: np.ndarray = field(default_factory=lambda: np.eye(4))
    vis_flags: int = VIS_ALL
    parent_to_local: Optional[np.ndarray] = None

    def __post_init__(self):
        self.local_to_parent = _mat(self.local_to_parent)
        self.parent_to_local = (
            invert_transform(self.local_to_parent) if self.parent_to_local is None else _mat(self.parent_to_local)
        )


@dataclass
class SceneDesc:
    meshes: List[Mesh]
    assemblies: List[Assembly]
    assembly_instances: List[AssemblyInstance]

    def to_c(self):
        """Pack into a ``CSceneDesc``.  Returns ``(desc, keepalive)``; the keepalive list owns
        every buffer the C struct points to."""
        keep = []
        cm = (CMesh * max(1, len(self.meshes)))()
        for i, m in enumerate(self.meshes):
            cm[i].vertices = m.vertices.ctypes.data
            cm[i].triangles = m.triangles.ctypes.data
            cm[i].triangle_pa = m.triangle_pa.ctypes.data if m.triangle_pa is not None else None
            cm[i].vertex_poses = m.vertex_poses.ctypes.data if m.vertex_poses is not None else None
            cm[i].vertex_count = m.vertices.shape[0]
            cm[i].triangle_count = m.triangles.shape[0]
            cm[i].motion_segment_count = m.motion_segment_count
            keep += [m.vertices, m.triangles, m.triangle_pa, m.vertex_poses]
        ca = (CAssembly * max(1, len(self.assemblies)))()
        for i, a in enumerate(self.assemblies):
            co = (CObjectInstance * max(1, len(a.object_instances)))()
            for j, o in enumerate(a.object_instances):
                co[j].local_to_parent[:] = o.local_to_parent.reshape(-1).tolist()
                co[j].parent_to_local[:] = o.parent_to_local.reshape(-1).tolist()
                co[j].mesh_index = o.mesh_index
                co[j].vis_flags = o.vis_flags & 0xFFFFFFFF
            keep.append(co)
            ca[i].object_instances = C.cast(co, C.POINTER(CObjectInstance))
            ca[i].object_instance_count = len(a.object_instances)
            ca[i].max_leaf_size = a.max_leaf_size
            ca[i].interior_node_traversal_cost = a.interior_node_traversal_cost
            ca[i].triangle_intersection_cost = a.triangle_intersection_cost
            ca[i].time = a.time
        ci = (CAssemblyInstance * max(1, len(self.assembly_instances)))()
        for i, inst in enumerate(self.assembly_instances):
            ci[i].local_to_parent[:] = inst.local_to_parent.reshape(-1).tolist()
            ci[i].parent_to_local[:] = inst.parent_to_local.reshape(-1).tolist()
            ci[i].assembly_index = inst.assembly_index
            ci[i].vis_flags = inst.vis_flags & 0xFFFFFFFF
        desc = CSceneDesc()
        desc.meshes = C.cast(cm, C.POINTER(CMesh))
        desc.assemblies = C.cast(ca, C.POINTER(CAssembly))
        desc.assembly_instances = C.cast(ci, C.POINTER(CAssemblyInstance))
        desc.mesh_count = len(self.meshes)
        desc.assembly_count = len(self.assemblies)
        desc.assembly_instance_count = len(self.assembly_instances)
        keep += [cm, ca, ci]
        return desc, keep


class CAlphaMask(C.Structure):
    _fields_ = [("bits", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class CIntersectionFilter(C.Structure):
    _fields_ = [("object_mask", CAlphaMask), ("material_masks", C.POINTER(CAlphaMask)), ("material_mask_count", C.c_uint32),
                ("reserved", C.c_uint32), ("uv", C.c_void_p)]


def pack_mask(opaque: np.ndarray) -> np.ndarray:
    """A boolean image [height, width] as foundation::BitMask2 storage (bitmask.h:
    bits[y * ((width + 7) / 8) + x / 8] bit (x & 7))."""
    opaque = np.asarray(opaque, dtype=bool)
    h, w = opaque.shape
    padded = np.zeros((h, (w + 7) // 8 * 8), dtype=bool)
    padded[:, :w] = opaque
    return np.ascontiguousarray(np.packbits(padded, axis=1, bitorder="little"))


@dataclass
class IntersectionFilter:
    """``renderer::IntersectionFilter`` of one object instance (intersectionfilter.h): an optional
    object alpha mask, optional per-material alpha masks (indexed by the triangle's
    primitive-attribute index) and three UV pairs per triangle.  Masks are boolean images
    [height, width], True = opaque."""
    uv: np.ndarray                                  # (triangle_count, 3, 2) float32
    object_mask: Optional[np.ndarray] = None
    material_masks: Optional[List[Optional[np.ndarray]]] = None

    def to_c(self):
        keep = []
        f = CIntersectionFilter()

        def mask(m):
            c = CAlphaMask()
            if m is not None:
                bits = pack_mask(m)
                keep.append(bits)
                c.bits, c.width, c.height = bits.ctypes.data, m.shape[1], m.shape[0]
            return c

        f.object_mask = mask(self.object_mask)
        mats = self.material_masks or []
        arr = (CAlphaMask * max(1, len(mats)))(*[mask(m) for m in mats])
        keep.append(arr)
        f.material_masks = C.cast(arr, C.POINTER(CAlphaMask))
        f.material_mask_count = len(mats)
        uv = np.ascontiguousarray(self.uv, dtype=np.float32).reshape(-1)
        keep.append(uv)
        f.uv = uv.ctypes.data
        return f, keep


class CInstanceKeys(C.Structure):
    _fields_ = [("times", C.c_void_p), ("local_to_parent", C.c_void_p), ("parent_to_local", C.c_void_p),
                ("key_count", C.c_uint32), ("reserved", C.c_uint32)]


class CItemMotion(C.Structure):
    """asgpu_item_motion / orc_item_motion."""
    _fields_ = [("key_times", C.c_void_p), ("key_parent_to_local", C.c_void_p), ("segments", C.c_void_p),
                ("key_count", C.c_uint32), ("reserved", C.c_uint32)]


@dataclass
class InstanceKeys:
    """Keys of an animated assembly instance (``TransformSequence``): ascending times and one
    local-to-parent matrix per key."""
    times: np.ndarray                      # (k,) float32
    local_to_parent: np.ndarray            # (k, 4, 4) float64

    def __post_init__(self):
        self.times = np.ascontiguousarray(self.times, dtype=np.float32).reshape(-1)
        self.local_to_parent = np.ascontiguousarray(self.local_to_parent, dtype=np.float64).reshape(-1, 4, 4)
        self.parent_to_local = np.ascontiguousarray(np.linalg.inv(self.local_to_parent))
        assert len(self.times) == len(self.local_to_parent) and np.all(np.diff(self.times) > 0)


@dataclass
class RayBatch:
    """The ShadingRay fields the path consumes (renderer/kernel/shading/shadingray.h:99-109,
    foundation/math/ray.h:68-71), one array per field."""

    org: np.ndarray                                # (n, 3) float64
    dir: np.ndarray                                # (n, 3) float64
    tmin: np.ndarray                               # (n,) float64
    tmax: np.ndarray                               # (n,) float64
    time_absolute: Optional[np.ndarray] = None     # (n,) float32
    time_normalized: Optional[np.ndarray] = None   # (n,) float32
    flags: Optional[np.ndarray] = None             # (n,) uint32

    def __post_init__(self):
        self.org = np.ascontiguousarray(self.org, dtype=np.float64).reshape(-1, 3)
        n = self.org.shape[0]
        self.dir = np.ascontiguousarray(self.dir, dtype=np.float64).reshape(-1, 3)
        self.tmin = np.ascontiguousarray(np.broadcast_to(np.asarray(self.tmin, dtype=np.float64), (n,)))
        self.tmax = np.ascontiguousarray(np.broadcast_to(np.asarray(self.tmax, dtype=np.float64), (n,)))
        if self.time_absolute is not None:
            self.time_absolute = np.ascontiguousarray(self.time_absolute, dtype=np.float32).reshape(n)
        if self.time_normalized is not None:
            self.time_normalized = np.ascontiguousarray(self.time_normalized, dtype=np.float32).reshape(n)
        if self.flags is not None:
            self.flags = np.ascontiguousarray(self.flags, dtype=np.uint32).reshape(n)
        assert self.dir.shape[0] == n

    def __len__(self) -> int:
        return self.org.shape[0]

    def slice(self, lo: int, hi: int) -> "RayBatch":
        s = slice(lo, hi)
        return RayBatch(
            self.org[s], self.dir[s], self.tmin[s], self.tmax[s],
            None if self.time_absolute is None else self.time_absolute[s],
            None if self.time_normalized is None else self.time_normalized[s],
            None if self.flags is None else self.flags[s],
        )

    def take(self, idx) -> "RayBatch":
        return RayBatch(
            self.org[idx], self.dir[idx], self.tmin[idx], self.tmax[idx],
            None if self.time_absolute is None else self.time_absolute[idx],
            None if self.time_normalized is None else self.time_normalized[idx],
            None if self.flags is None else self.flags[idx],
        )

    def to_c(self) -> CRays:
        r = CRays()
        r.org = self.org.ctypes.data
        r.dir = self.dir.ctypes.data
        r.tmin = self.tmin.ctypes.data
        r.tmax = self.tmax.ctypes.data
        r.time_absolute = self.time_absolute.ctypes.data if self.time_absolute is not None else None
        r.time_normalized = self.time_normalized.ctypes.data if self.time_normalized is not None else None
        r.flags = self.flags.ctypes.data if self.flags is not None else None
        return r

    @property
    def bytes_per_ray(self) -> int:
        b = 48 + 16
        b += 4 if self.time_absolute is not None else 0
        b += 4 if self.time_normalized is not None else 0
        b += 4 if self.flags is not None else 0
        return b
