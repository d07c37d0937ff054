"""Loader of the engine's C-ABI shared library (include/asgpu.h).

The library is built in-tree (``appleseed_b200/libasgpu.so``) by ``__graft_entry__.build()`` /
``make -C appleseed_b200/csrc``.  There is no fallback: if it is missing or fails to load, every
entry point of the package raises.
"""
from __future__ import annotations

import ctypes as C
import os

from .scene import CIntersectionFilter, CRays, CSceneDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
# ASGPU_LIB: another build of the same library (kernel experiments: tools/build_variant.sh).
LIB_PATH = os.environ.get("ASGPU_LIB") or os.path.join(_HERE, "libasgpu.so")

# Every symbol include/asgpu.h declares.
EXPORTS = [
    "asgpu_trees_build", "asgpu_trees_destroy", "asgpu_trees_triangle_tree_count",
    "asgpu_trees_get_triangle_tree", "asgpu_trees_get_assembly_tree", "asgpu_trees_build_seconds",
    "asgpu_scene_create", "asgpu_scene_create_from_desc", "asgpu_scene_destroy",
    "asgpu_scene_blob_size", "asgpu_scene_blob_device_ptr", "asgpu_scene_export_blob", "asgpu_scene_import_blob",
    "asgpu_scene_get_info",
    "asgpu_trace", "asgpu_trace_probe", "asgpu_trace_host", "asgpu_trace_probe_host",
    "asgpu_get_counters", "asgpu_last_error", "asgpu_version", "asgpu_sort_rays",
    "asgpu_trees_get_source_geometry", "asgpu_scene_create_ex", "asgpu_refine_and_offset",
    "asgpu_trace_with_parents", "asgpu_trace_probe_with_parents",
    "asgpu_queue_create", "asgpu_queue_destroy", "asgpu_queue_capacity", "asgpu_queue_device_arrays",
    "asgpu_queue_reset", "asgpu_queue_count", "asgpu_queue_push_host", "asgpu_trace_queue", "asgpu_trace_probe_queue",
    "asgpu_path_stream_create", "asgpu_path_stream_destroy", "asgpu_path_stream_tile_count", "asgpu_path_stream_render",
    "asgpu_path_stream_read_image", "asgpu_path_stream_clear", "asgpu_path_stream_get_stats",
    "asgpu_path_stream_capture", "asgpu_path_stream_capture_count", "asgpu_path_stream_capture_get",
    "asgpu_trees_build_on_device", "asgpu_trees_device_seconds",
    "asgpu_get_counters_by_kind", "asgpu_get_lane_profile", "asgpu_get_support_planes", "asgpu_pin_host", "asgpu_unpin_host", "asgpu_reload_tuning",
    "asgpu_path_stream_capture_get_times", "asgpu_path_stream_set_profiling", "asgpu_path_stream_get_profile",
    "asgpu_path_stream_read_tiles", "asgpu_trees_build_animated",
]

SCENE_EXACT = 1 << 0
SCENE_WIDE = 1 << 1
SCENE_DEFAULT = SCENE_EXACT | SCENE_WIDE
TRACE_EXACT = 1 << 0
TRACE_COUNTERS = 1 << 1
TRACE_SORT = 1 << 2
STREAM_PARENTS = 1 << 0


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
    _fields_ = [("nodes", C.c_void_p), ("items", C.POINTER(AssemblyItem)), ("node_count", C.c_uint64), ("item_count", C.c_uint64),
                ("item_motion", C.c_void_p)]


class SourceGeometry(C.Structure):
    _fields_ = [("objects", C.c_void_p), ("object_count", C.c_uint32), ("reserved", C.c_uint32), ("filters", C.c_void_p)]


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


class LaneProfile(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("rounds", "testing", "no_ray", "traversed", "held", "want_enter", "found_leaf", "nothing_to_fetch",
                                          "iterations", "batched_entries", "lanes_entered", "refills", "lanes_refilled")] + [("reserved", C.c_uint64 * 3)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_ if k != "reserved"}


class PathStreamDesc(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("spp", C.c_uint32), ("max_bounces", C.c_uint32),
        ("tile_size", C.c_uint32), ("light_count", C.c_uint32), ("trace_flags", C.c_uint32), ("stream_flags", C.c_uint32),
        ("seed", C.c_uint64), ("camera_to_world", C.c_double * 12),
        ("film_width", C.c_double), ("film_height", C.c_double), ("focal_length", C.c_double),
        ("lights", (C.c_double * 3) * 8), ("offset_eps", C.c_double),
        ("shutter_open", C.c_float), ("shutter_close", C.c_float),
    ]


class PathStreamProfile(C.Structure):
    _fields_ = [
        ("closest_ms", C.c_double), ("probe_ms", C.c_double), ("refine_ms", C.c_double), ("stage_ms", C.c_double),
        ("closest_launches", C.c_uint64), ("probe_launches", C.c_uint64),
    ]

    def as_dict(self):
        return {k: (float if "ms" in k else int)(getattr(self, k)) for k, _ in self._fields_}


class PathStreamStats(C.Structure):
    _fields_ = [
        ("camera_rays", C.c_uint64), ("bounce_rays", C.c_uint64), ("probe_rays", C.c_uint64),
        ("surface_hits", C.c_uint64), ("escaped", C.c_uint64), ("unoccluded", C.c_uint64),
        ("wavefronts", C.c_uint64), ("kernel_launches", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


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
    lib.asgpu_trees_build_animated.restype = C.c_void_p
    lib.asgpu_trees_build_animated.argtypes = [P(CSceneDesc), C.c_void_p, C.c_int]
    lib.asgpu_trees_build_on_device.restype = C.c_void_p
    lib.asgpu_trees_build_on_device.argtypes = [P(CSceneDesc), C.c_int, C.c_int]
    lib.asgpu_trees_destroy.argtypes = [C.c_void_p]
    lib.asgpu_trees_triangle_tree_count.argtypes = [C.c_void_p]
    lib.asgpu_trees_get_triangle_tree.argtypes = [C.c_void_p, C.c_int, P(TriangleTreeView)]
    lib.asgpu_trees_get_assembly_tree.argtypes = [C.c_void_p, P(AssemblyTreeView)]
    lib.asgpu_trees_build_seconds.restype = C.c_double
    lib.asgpu_trees_build_seconds.argtypes = [C.c_void_p]
    lib.asgpu_trees_device_seconds.restype = C.c_double
    lib.asgpu_trees_device_seconds.argtypes = [C.c_void_p]
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
    lib.asgpu_get_counters_by_kind.argtypes = [C.c_void_p, P(Counters), P(Counters), C.c_int]
    lib.asgpu_get_lane_profile.argtypes = [C.c_void_p, P(LaneProfile), P(LaneProfile)]
    lib.asgpu_trees_get_source_geometry.argtypes = [C.c_void_p, C.c_int, P(SourceGeometry)]
    lib.asgpu_scene_create_ex.restype = C.c_void_p
    lib.asgpu_scene_create_ex.argtypes = [P(TriangleTreeView), C.c_uint32, P(AssemblyTreeView), P(SourceGeometry), C.c_uint32, C.c_int]
    lib.asgpu_refine_and_offset.argtypes = [C.c_void_p, P(CRays), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.asgpu_trace_with_parents.argtypes = [C.c_void_p, P(CRays), C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.asgpu_trace_probe_with_parents.argtypes = [C.c_void_p, P(CRays), C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.asgpu_sort_rays.argtypes = [C.c_void_p, P(CRays), C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.asgpu_queue_create.restype = C.c_void_p
    lib.asgpu_queue_create.argtypes = [C.c_void_p, C.c_size_t]
    lib.asgpu_queue_destroy.argtypes = [C.c_void_p]
    lib.asgpu_queue_capacity.restype = C.c_size_t
    lib.asgpu_queue_capacity.argtypes = [C.c_void_p]
    lib.asgpu_queue_device_arrays.argtypes = [C.c_void_p, P(CRays), P(C.c_void_p), P(C.c_void_p)]
    lib.asgpu_queue_reset.argtypes = [C.c_void_p, C.c_void_p]
    lib.asgpu_queue_count.argtypes = [C.c_void_p, C.c_void_p, P(C.c_uint64)]
    lib.asgpu_queue_push_host.argtypes = [C.c_void_p, P(CRays), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.asgpu_trace_queue.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.asgpu_trace_probe_queue.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.asgpu_path_stream_create.restype = C.c_void_p
    lib.asgpu_path_stream_create.argtypes = [C.c_void_p, P(PathStreamDesc), C.c_size_t]
    lib.asgpu_path_stream_destroy.argtypes = [C.c_void_p]
    lib.asgpu_path_stream_tile_count.restype = C.c_uint32
    lib.asgpu_path_stream_tile_count.argtypes = [C.c_void_p]
    lib.asgpu_path_stream_render.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.asgpu_path_stream_read_image.argtypes = [C.c_void_p, C.c_void_p]
    lib.asgpu_path_stream_read_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.asgpu_path_stream_clear.argtypes = [C.c_void_p]
    lib.asgpu_path_stream_get_stats.argtypes = [C.c_void_p, P(PathStreamStats)]
    lib.asgpu_path_stream_capture.argtypes = [C.c_void_p, C.c_size_t]
    lib.asgpu_path_stream_capture_count.argtypes = [C.c_void_p]
    lib.asgpu_path_stream_capture_get.restype = C.c_longlong
    lib.asgpu_path_stream_capture_get.argtypes = [C.c_void_p, C.c_int, P(C.c_int), P(C.c_uint32)] + [C.c_void_p] * 8
    lib.asgpu_path_stream_capture_get_times.restype = C.c_longlong
    lib.asgpu_path_stream_capture_get_times.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.asgpu_path_stream_set_profiling.argtypes = [C.c_void_p, C.c_int]
    lib.asgpu_path_stream_get_profile.argtypes = [C.c_void_p, P(PathStreamProfile)]
    lib.asgpu_get_support_planes.argtypes = [C.c_void_p, P(CRays), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.asgpu_pin_host.argtypes = [C.c_void_p, C.c_size_t]
    lib.asgpu_unpin_host.argtypes = [C.c_void_p]
    lib.asgpu_reload_tuning.restype = None
    lib.asgpu_reload_tuning.argtypes = []
    _lib = lib
    return lib


def last_error() -> str:
    return load().asgpu_last_error().decode("utf-8", "replace")
