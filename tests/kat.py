"""Known-answer vectors transcribed from the reference's own unit tests.

Shared by the CPU-oracle tests and the GPU parity tests so both are pinned to the same cases.
"""
import numpy as np

from appleseed_b200 import scenes
from appleseed_b200.scene import (VIS_CAMERA, Assembly, AssemblyInstance, Mesh, ObjectInstance,
                                  RayBatch, SceneDesc)

DBL_MAX = float(np.finfo(np.float64).max)

# foundation/meta/tests/test_intersection_raytriangle.cpp:46-51 (fixture), :60-110 (cases).
TRI = ([0.5, 0.0, 0.5], [-0.5, 0.0, 0.5], [-0.5, 0.0, -0.5])
# (name, org, dir, tmin, tmax, expect_hit, expect (t, u, v) or None)
RAY_TRIANGLE = [
    ("tmin_equal_hit_distance_true",  [-0.2, 1.0, 0.2], [0.0, -1.0, 0.0], 1.0, 10.0,    True,  (1.0, None, None)),
    ("tmax_equal_hit_distance_false", [-0.2, 1.0, 0.2], [0.0, -1.0, 0.0], 0.0, 1.0,     False, None),
    ("quad_diagonal",                 [0.0, 1.0, 0.0],  [0.0, -1.0, 0.0], 0.0, DBL_MAX, True,  (1.0, 0.0, 0.5)),
]

# foundation/meta/tests/test_intersection_rayaabb.cpp: all 34 cases.  First the 17 three-argument
# cases (:47-214; the distance column is what the 4-argument overload returns on the same input).
# (name, bmin, bmax, org, dir, tmin, tmax, hit, distance)
_U = ([-1.0] * 3, [1.0] * 3)
_P = ([0.0] * 3, [1.0] * 3)
RAY_AABB = [
    ("not_piercing",        *_U, [2, 0, 2],  [0, 0, -1], 0.0, DBL_MAX, False, None),
    ("pos_x_face_middle",   *_U, [2, 0, 0],  [-1, 0, 0], 0.0, DBL_MAX, True,  1.0),
    ("neg_x_face_middle",   *_U, [-2, 0, 0], [1, 0, 0],  0.0, DBL_MAX, True,  1.0),
    ("pos_y_face_middle",   *_U, [0, 2, 0],  [0, -1, 0], 0.0, DBL_MAX, True,  1.0),
    ("neg_y_face_middle",   *_U, [0, -2, 0], [0, 1, 0],  0.0, DBL_MAX, True,  1.0),
    ("pos_z_face_middle",   *_U, [0, 0, 2],  [0, 0, -1], 0.0, DBL_MAX, True,  1.0),
    ("neg_z_face_middle",   *_U, [0, 0, -2], [0, 0, 1],  0.0, DBL_MAX, True,  1.0),
    ("embedded_pos_x_face", *_P, [1, 0, 2],  [0, 0, -1], 0.0, DBL_MAX, True,  1.0),
    ("embedded_neg_x_face", *_P, [0, 0, 2],  [0, 0, -1], 0.0, DBL_MAX, True,  1.0),
    ("embedded_pos_y_face", *_P, [2, 1, 0],  [-1, 0, 0], 0.0, DBL_MAX, True,  1.0),
    ("embedded_neg_y_face", *_P, [2, 0, 0],  [-1, 0, 0], 0.0, DBL_MAX, True,  1.0),
    ("embedded_pos_z_face", *_P, [0, 2, 1],  [0, -1, 0], 0.0, DBL_MAX, True,  1.0),
    ("embedded_neg_z_face", *_P, [0, 2, 0],  [0, -1, 0], 0.0, DBL_MAX, True,  1.0),
    ("tmin_equal_hit",      *_U, [0, 0, 2],  [0, 0, -1], 3.0, 10.0,    True,  3.0),
    ("tmax_equal_hit",      *_U, [0, 0, 2],  [0, 0, -1], 0.0, 1.0,     False, None),
    ("tmin_larger_than_hit", *_U, [0, 0, 2], [0, 0, -1], 3.1, 10.0,    False, None),
    ("tmax_smaller_than_hit", *_U, [0, 0, 2], [0, 0, -1], 0.0, 0.9,    False, None),
]

# The 11 four-argument cases (:217-346: the distance argument starts at 42 and must be left alone on
# a miss) and the 6 clip cases (:349-419: ray.m_tmin / m_tmax updated on a hit, untouched on a miss).
# (name, bmin, bmax, org, dir, tmin, tmax, hit, distance after the call)
RAY_AABB_DISTANCE = [
    ("not_piercing_distance_unchanged", *_U, [2, 0, 2], [0, 0, -1], 0.0, DBL_MAX, False, 42.0),
    ("embedded_pos_x_face_distance", *_P, [1, 0, 2], [0, 0, -1], 0.0, DBL_MAX, True, 1.0),
    ("embedded_neg_x_face_distance", *_P, [0, 0, 2], [0, 0, -1], 0.0, DBL_MAX, True, 1.0),
    ("embedded_pos_y_face_distance", *_P, [2, 1, 0], [-1, 0, 0], 0.0, DBL_MAX, True, 1.0),
    ("embedded_neg_y_face_distance", *_P, [2, 0, 0], [-1, 0, 0], 0.0, DBL_MAX, True, 1.0),
    ("embedded_pos_z_face_distance", *_P, [0, 2, 1], [0, -1, 0], 0.0, DBL_MAX, True, 1.0),
    ("embedded_neg_z_face_distance", *_P, [0, 2, 0], [0, -1, 0], 0.0, DBL_MAX, True, 1.0),
    ("tmin_equal_hit_distance", *_U, [0, 0, 2], [0, 0, -1], 3.0, 10.0, True, 3.0),
    ("tmax_equal_hit_distance_unchanged", *_U, [0, 0, 2], [0, 0, -1], 0.0, 1.0, False, 42.0),
    ("tmin_larger_than_hit_distance_unchanged", *_U, [0, 0, 2], [0, 0, -1], 3.1, 10.0, False, 42.0),
    ("tmax_smaller_than_hit_distance_unchanged", *_U, [0, 0, 2], [0, 0, -1], 0.0, 0.9, False, 42.0),
]
# (name, bmin, bmax, org, dir, tmin, tmax, hit, ray tmin after, ray tmax after)
RAY_AABB_CLIP = [
    ("clip_not_piercing", *_U, [2, 0, 2], [0, 0, -1], 0.0, DBL_MAX, False, 0.0, DBL_MAX),
    ("clip_pos_z_face_middle", *_U, [0, 0, 2], [0, 0, -1], 0.0, DBL_MAX, True, 1.0, 3.0),
    ("clip_tmin_equal_hit", *_U, [0, 0, 2], [0, 0, -1], 3.0, 10.0, True, 3.0, 3.0),
    ("clip_tmax_equal_hit", *_U, [0, 0, 2], [0, 0, -1], 0.0, 1.0, False, 0.0, 1.0),
    ("clip_tmin_larger_than_hit", *_U, [0, 0, 2], [0, 0, -1], 3.1, 10.0, False, 3.1, 10.0),
    ("clip_tmax_smaller_than_hit", *_U, [0, 0, 2], [0, 0, -1], 0.0, 0.9, False, 0.0, 0.9),
]

# foundation/meta/tests/test_ray.cpp:59-83.
RAY_INFO = ([-2.0, 0.0, 2.0], [-0.5, np.inf, 0.5], [0, 1, 1])


def unit_quad() -> Mesh:
    """renderer/meta/tests/test_tracer.cpp:172-194: two triangles in the x = 0 plane, +-0.5."""
    v = np.array([[0.0, -0.5, -0.5], [0.0, 0.5, -0.5], [0.0, 0.5, 0.5], [0.0, -0.5, 0.5]], dtype=np.float32)
    return Mesh(v, np.array([[0, 1, 2], [2, 3, 0]], dtype=np.uint32))


def tracer_scene(xs, scale=1.0) -> SceneDesc:
    """Unit quads instanced at x in ``xs`` through assembly instances (test_tracer.cpp:421-456,
    :954-981); ``scale`` scales the whole assembly instance (:1016-1059)."""
    insts = [AssemblyInstance(0, scenes.scaling(scale) @ scenes.translation(x, 0.0, 0.0)) for x in xs]
    return SceneDesc([unit_quad()], [Assembly([ObjectInstance(0)])], insts)


def x_ray(tmax=DBL_MAX, flags=VIS_CAMERA) -> RayBatch:
    return RayBatch(np.zeros((1, 3)), np.array([[1.0, 0.0, 0.0]]), 0.0, tmax, flags=np.array([flags], dtype=np.uint32))


def empty_bbox_scene() -> SceneDesc:
    """renderer/meta/tests/test_intersector.cpp:61-90: an assembly whose only object has a
    bounding box [-1, 1]^3 and no geometry."""
    v = np.array([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]], dtype=np.float32)
    mesh = Mesh(v, np.zeros((0, 3), dtype=np.uint32))
    return SceneDesc([mesh], [Assembly([ObjectInstance(0)])], [AssemblyInstance(0)])


def empty_bbox_ray() -> RayBatch:
    """test_intersector.cpp:118-126: tmax ends inside the assembly."""
    return RayBatch(np.array([[0.0, 0.0, 2.0]]), np.array([[0.0, 0.0, -1.0]]), 0.0, 2.0,
                    flags=np.array([VIS_CAMERA], dtype=np.uint32))
