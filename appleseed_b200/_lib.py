"""Loader of the engine's C-ABI shared library (include/asgpu.h).

The library is built in-tree (``appleseed_b200/libasgpu.so``) by ``__graft_entry__.build()`` /
``make -C appleseed_b200/csrc``.  There is no fallback: if it is missing or fails to load, every
entry point of the package raises.
"""
from __future__ import annotations

import ctypes as C
import os

from .scene import CRays, CSceneDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libasgpu.so")

# Every symbol include/asgpu.h declares.
EXPORTS = [
    "asgpu_trees_build", "asgpu_trees_destroy", "asgpu_trees_triangle_tree_count",
    "asgpu_trees_get_triangle_tree", "asgpu_trees_get_assembly_tree", "asgpu_trees_build_seconds",
    "asgpu_scene_create", "asgpu_scene_create_from_desc", "asgpu_scene_destroy",
    "asgpu_scene_blob_size", "asgpu_scene_blob_device_ptr", "asgpu_scene_export_blob", "asgpu_scene_import_blob",
    "asgpu_scene_get_info",
    "asgpu_trace", "asgpu_trace_probe", "asgpu_trace_host", "asgpu_trace_probe_host",
    "asgpu_get_counters", "asgpu_last_error", "asgpu_version",
]

SCENE_EXACT = 1 << 0
SCENE_WIDE = 1 << 1
SCENE_DEFAULT = SCENE_EXACT | SCENE_WIDE
TRACE_EXACT = 1 << 0
TRACE_COUNTERS = 1 << 1
TRACE_SORT = 1 << 2


class TriangleTreeView(C.Structure):
    _fields_ = [
        ("nodes", C.c_void_p), ("node_bboxes", C.c_void_p), ("leaf_data", C.c_void_p), ("triangle_keys", C.c_void_p),
        ("node_count", C.c_uint64), ("node_bbox_count", C.c_uint64), ("leaf_data_size", C.c_uint64),
        ("triangle_key_count", C.c_uint64), ("static_triangle_count", C.c_uint64), ("moving_triangle_count", C.c_uint64),
    ]


class AssemblyItem(C.Structure):
    _fields_ = [
        ("parent_to_local", C.c_double * 16), ("assembly_instance", C.c_uint32), ("triangle_tree", C.c_uint32),
        ("vis_flags", C.c_uint32), ("reserved", C.c_uint32),
    ]


class AssemblyTreeView(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("items", C.POINTER(AssemblyItem)), ("node_count", C.c_uint64), ("item_count", C.c_uint64)]


class SceneInfo(C.Structure):
    _fields_ = [
        ("blob_bytes", C.c_uint64), ("triangle_tree_count", C.c_uint64), ("instance_count", C.c_uint64),
        ("triangle_count", C.c_uint64), ("moving_triangle_count", C.c_uint64), ("binary_node_count", C.c_uint64),
        ("wide_node_count", C.c_uint64), ("binary_node_bytes", C.c_uint64), ("wide_node_bytes", C.c_uint64),
        ("triangle_bytes", C.c_uint64), ("flags", C.c_uint32), ("wide_stack_depth", C.c_uint32),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_ if k != "reserved"}


class Counters(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64), ("assembly_nodes_visited", C.c_uint64), ("instances_visited", C.c_uint64),
        ("triangle_nodes_visited", C.c_uint64), ("triangles_tested", C.c_uint64), ("hits", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("reserved", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_ if k != "reserved"}


_lib = None


def load() -> C.CDLL:
    """Load libasgpu.so and declare the prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "appleseed_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    lib.asgpu_last_error.restype = C.c_char_p
    lib.asgpu_version.restype = C.c_int
    lib.asgpu_trees_build.restype = C.c_void_p
    lib.asgpu_trees_build.argtypes = [P(CSceneDesc), C.c_int]
    lib.asgpu_trees_destroy.argtypes = [C.c_void_p]
    lib.asgpu_trees_triangle_tree_count.argtypes = [C.c_void_p]
    lib.asgpu_trees_get_triangle_tree.argtypes = [C.c_void_p, C.c_int, P(TriangleTreeView)]
    lib.asgpu_trees_get_assembly_tree.argtypes = [C.c_void_p, P(AssemblyTreeView)]
    lib.asgpu_trees_build_seconds.restype = C.c_double
    lib.asgpu_trees_build_seconds.argtypes = [C.c_void_p]
    lib.asgpu_scene_create.restype = C.c_void_p
    lib.asgpu_scene_create.argtypes = [P(TriangleTreeView), C.c_uint32, P(AssemblyTreeView), C.c_uint32, C.c_int]
    lib.asgpu_scene_create_from_desc.restype = C.c_void_p
    lib.asgpu_scene_create_from_desc.argtypes = [P(CSceneDesc), C.c_uint32, C.c_int, C.c_int]
    lib.asgpu_scene_destroy.argtypes = [C.c_void_p]
    lib.asgpu_scene_blob_size.restype = C.c_size_t
    lib.asgpu_scene_blob_size.argtypes = [C.c_void_p]
    lib.asgpu_scene_blob_device_ptr.restype = C.c_void_p
    lib.asgpu_scene_blob_device_ptr.argtypes = [C.c_void_p]
    lib.asgpu_scene_export_blob.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.asgpu_scene_import_blob.restype = C.c_void_p
    lib.asgpu_scene_import_blob.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    lib.asgpu_scene_get_info.argtypes = [C.c_void_p, P(SceneInfo)]
    lib.asgpu_trace.argtypes = [C.c_void_p, P(CRays), C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.asgpu_trace_probe.argtypes = [C.c_void_p, P(CRays), C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.asgpu_trace_host.argtypes = [C.c_void_p, P(CRays), C.c_size_t, C.c_void_p, C.c_uint32]
    lib.asgpu_trace_probe_host.argtypes = [C.c_void_p, P(CRays), C.c_size_t, C.c_void_p, C.c_uint32]
    lib.asgpu_get_counters.argtypes = [C.c_void_p, P(Counters), C.c_int]
    _lib = lib
    return lib


def last_error() -> str:
    return load().asgpu_last_error().decode("utf-8", "replace")
