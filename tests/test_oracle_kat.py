"""Pins the CPU oracle(s) to the reference's own known-answer tests (CPU only)."""
import numpy as np
import pytest

import kat


@pytest.fixture(params=["orc", "asref"])
def oracle(request):
    return request.getfixturevalue(request.param)


@pytest.mark.parametrize("case", kat.RAY_TRIANGLE, ids=[c[0] for c in kat.RAY_TRIANGLE])
def test_ray_triangle(oracle, case):
    _, org, d, tmin, tmax, expect_hit, tuv = case
    hit, got = oracle.kat_ray_triangle(*kat.TRI, org, d, tmin, tmax)
    assert hit == expect_hit
    assert oracle.kat_ray_triangle_bool(*kat.TRI, org, d, tmin, tmax) == expect_hit
    if tuv is not None:
        for g, e in zip(got, tuv):
            if e is not None:
                assert abs(g - e) <= 1e-14


@pytest.mark.parametrize("case", kat.RAY_AABB, ids=[c[0] for c in kat.RAY_AABB])
def test_ray_aabb(oracle, case):
    _, bmin, bmax, org, d, tmin, tmax, expect_hit, dist = case
    hit, t = oracle.kat_ray_aabb(bmin, bmax, org, d, tmin, tmax)
    assert hit == expect_hit
    if dist is not None:
        assert abs(t - dist) <= 1e-14


@pytest.mark.parametrize("case", kat.RAY_AABB, ids=[c[0] for c in kat.RAY_AABB])
def test_ray_aabb_three_argument_overload(oracle, case):
    _, bmin, bmax, org, d, tmin, tmax, expect_hit, _ = case
    hit, _ = oracle.kat_ray_aabb_ex(0, bmin, bmax, org, d, tmin, tmax, [0.0])
    assert hit == expect_hit


@pytest.mark.parametrize("case", kat.RAY_AABB_DISTANCE, ids=[c[0] for c in kat.RAY_AABB_DISTANCE])
def test_ray_aabb_distance_contract(oracle, case):
    _, bmin, bmax, org, d, tmin, tmax, expect_hit, dist = case
    hit, io = oracle.kat_ray_aabb_ex(1, bmin, bmax, org, d, tmin, tmax, [42.0])
    assert hit == expect_hit
    if hit:
        assert abs(io[0] - dist) <= 1e-14
    else:
        assert io[0] == 42.0                 # EXPECT_EQ(42.0, distance)


@pytest.mark.parametrize("case", kat.RAY_AABB_CLIP, ids=[c[0] for c in kat.RAY_AABB_CLIP])
def test_ray_aabb_clip(oracle, case):
    _, bmin, bmax, org, d, tmin, tmax, expect_hit, tmin_after, tmax_after = case
    hit, io = oracle.kat_ray_aabb_ex(2, bmin, bmax, org, d, tmin, tmax, [tmin, tmax])
    assert hit == expect_hit
    if hit:
        assert abs(io[0] - tmin_after) <= 1e-14 and abs(io[1] - tmax_after) <= 1e-14
    else:
        assert io[0] == tmin_after and io[1] == tmax_after


def test_ray_info(oracle):
    d, rcp, sgn = kat.RAY_INFO
    got_rcp, got_sgn = oracle.kat_ray_info(d)
    assert got_rcp[0] == rcp[0] and np.isposinf(got_rcp[1]) and got_rcp[2] == rcp[2]
    assert list(got_sgn) == sgn


def test_negative_zero_direction_has_sign_zero(oracle):
    # ray.h:313-321: sgn = (1 / dir >= 0); 1 / -0.0 = -inf -> 0.
    rcp, sgn = oracle.kat_ray_info([-0.0, 1.0, 1.0])
    assert np.isneginf(rcp[0]) and sgn[0] == 0


def test_tracer_quad_hit_at_two(oracle):
    # test_tracer.cpp:421-439: quad instanced at x = 2, ray (0,0,0) -> +x hits at distance 2.
    s = oracle.scene(kat.tracer_scene([2.0]))
    h = s.trace(kat.x_ray())
    assert h["prim_type"][0] == 2 and h["t"][0] == 2.0 and h["assembly_instance"][0] == 0
    # :441-456 opaque occluder => probe reports a hit.
    assert s.trace_probe(kat.x_ray())[0] == 1


def test_tracer_two_quads_nearest_then_unoccluded(oracle):
    # test_tracer.cpp:954-981: planes at x = 2 and x = 4: nearest is 2.0; a probe from beyond the
    # first plane to the second with tmax = dist * (1 - 1e-6) (tracer.h:249-259) is unoccluded.
    s = oracle.scene(kat.tracer_scene([2.0, 4.0]))
    h = s.trace(kat.x_ray())
    assert h["t"][0] == 2.0 and h["assembly_instance"][0] == 0
    from appleseed_b200.scene import RayBatch
    probe = RayBatch(np.array([[2.0 + 1e-9, 0.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]), 0.0, 2.0 * (1.0 - 1.0e-6))
    assert s.trace_probe(probe)[0] == 0


def test_tracer_scaled_assembly_instance(oracle):
    # test_tracer.cpp:1016-1059: the assembly instance scaled by 0.5 => distance 1.0; t is shared
    # between world and instance space because the local direction is not renormalised.
    s = oracle.scene(kat.tracer_scene([2.0], scale=0.5))
    h = s.trace(kat.x_ray())
    assert h["prim_type"][0] == 2 and abs(h["t"][0] - 1.0) <= 1e-15


def test_intersector_empty_bbox(oracle):
    # test_intersector.cpp:116-147.
    s = oracle.scene(kat.empty_bbox_scene())
    h = s.trace(kat.empty_bbox_ray())
    assert h["prim_type"][0] == 0 and h["assembly_instance"][0] == 0xFFFFFFFF and h["t"][0] == 2.0
    assert s.trace_probe(kat.empty_bbox_ray())[0] == 0


def test_interval_is_tmin_inclusive_tmax_exclusive(oracle):
    # ray.h:49-53 through the whole two-level path.
    s = oracle.scene(kat.tracer_scene([2.0]))
    from appleseed_b200.scene import RayBatch
    r = RayBatch(np.zeros((2, 3)), np.array([[1.0, 0, 0], [1.0, 0, 0]]), np.array([2.0, 0.0]), np.array([10.0, 2.0]))
    h = s.trace(r)
    assert h["prim_type"][0] == 2 and h["prim_type"][1] == 0
    p = s.trace_probe(r)
    assert list(p) == [1, 0]


def test_visibility_flags(oracle):
    # assemblytree.cpp:629 (instance flags) and triangletree.cpp:1389 (object-instance flags).
    from appleseed_b200.scene import (VIS_CAMERA, VIS_SHADOW, Assembly, AssemblyInstance, ObjectInstance, SceneDesc)
    from appleseed_b200 import scenes
    desc = SceneDesc(
        [kat.unit_quad()],
        [Assembly([ObjectInstance(0, vis_flags=VIS_CAMERA)]), Assembly([ObjectInstance(0)])],
        [AssemblyInstance(0, scenes.translation(2, 0, 0)), AssemblyInstance(1, scenes.translation(3, 0, 0), vis_flags=VIS_SHADOW)])
    s = oracle.scene(desc)
    assert s.trace(kat.x_ray(flags=VIS_CAMERA))["t"][0] == 2.0          # quad 0 visible to camera rays
    h = s.trace(kat.x_ray(flags=VIS_SHADOW))                              # only instance 1 visible
    assert h["t"][0] == 3.0 and h["assembly_instance"][0] == 1
    assert s.trace(kat.x_ray(flags=1 << 5))["prim_type"][0] == 0          # diffuse rays see nothing


def test_node_packing(asref):
    # foundation/meta/tests/test_bvh.cpp:62-75 (TestStorageAndRetrievalOf3DBoundingBoxes) on the
    # reference's own bvh::Node<AABB3d>, plus the raw layout the product's AsNode / flattener assume
    # (appleseed_b200/csrc/as_format.h; bvh_node.h:100-107, 141-162).
    import ctypes as C
    left = np.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    right = np.array([7.0, 8.0, 9.0, 10.0, 11.0, 12.0])
    back, raw = np.zeros(12), np.zeros(128, dtype=np.uint8)
    f = asref.lib.asref_kat_node_pack
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    f(left.ctypes.data, right.ctypes.data, 42, back.ctypes.data, raw.ctypes.data)
    assert np.array_equal(back[:6], left) and np.array_equal(back[6:], right)          # EXPECT_EQ(LeftBBox, get_left_bbox()) ...
    words = raw.view(np.uint32)
    assert words[0] == 0xFFFFFFFF and words[1] == 42                                      # interior marker, first child
    boxes = raw[32:].view(np.float64)
    for axis in range(3):                                                                 # [minL minR maxL maxR] per axis
        assert list(boxes[axis * 4: axis * 4 + 4]) == [left[axis], right[axis], left[3 + axis], right[3 + axis]]
    # The product's flattener reads exactly this node: a root over two empty leaves.
    from appleseed_b200 import _lib
    from hostsim import hostsim
    nodes = np.zeros(3 * 128, dtype=np.uint8)
    f(left.ctypes.data, right.ctypes.data, 1, back.ctypes.data, raw.ctypes.data)
    nodes[:128] = raw
    nodes[128 + 32: 128 + 36] = 0xFF                                                       # leaves: no items, payload "in node"
    nodes[256 + 32: 256 + 36] = 0xFF
    view = _lib.TriangleTreeView()
    view.nodes, view.node_count = nodes.ctypes.data, 3
    item = (_lib.AssemblyItem * 1)()
    item[0].parent_to_local[:] = np.eye(4).reshape(-1).tolist()
    item[0].assembly_instance, item[0].triangle_tree, item[0].vis_flags = 0, 0, 0xFFFFFFFF
    top_nodes = np.zeros(128, dtype=np.uint8)
    top_nodes.view(np.uint32)[0] = 1                                                       # one leaf holding item 0
    top = _lib.AssemblyTreeView()
    top.nodes, top.items, top.node_count, top.item_count = top_nodes.ctypes.data, C.cast(item, C.POINTER(_lib.AssemblyItem)), 1, 1
    s = hostsim.SimScene.from_views(hostsim.load(), [view], top, [nodes, top_nodes, item])
    rays = kat.x_ray()
    hits, counters = s.trace(rays, wide=False)
    assert hits["prim_type"][0] == 0 and int(counters[3]) >= 1                              # the root was visited, nothing to hit


def test_bitmask_storage(asref):
    # foundation/meta/tests/test_bitmask.cpp:101-128 (StressTest: 17 x 9, 1000 random set(x, y, value)
    # calls, the mask must equal a plain bool array after each) on the reference's own BitMask2 -- and
    # its raw storage must be the byte layout asgpu_alpha_mask::bits documents, i.e. what
    # scene.pack_mask produces for every filter test of the product.
    import ctypes as C
    from appleseed_b200.scene import pack_mask
    width, height, count = 17, 9, 1000
    rng = np.random.default_rng(7)
    xs = rng.integers(0, width, count).astype(np.uint32)
    ys = rng.integers(0, height, count).astype(np.uint32)
    vs = rng.integers(0, 2, count).astype(np.uint8)
    f = asref.lib.asref_kat_bitmask
    f.restype = None
    f.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    values = np.zeros((height, width), dtype=np.uint8)
    for n in (0, 1, 10, 500, count):                                   # the reference checks after every call; a few prefixes here
        values[:] = 0
        for i in range(n):
            values[ys[i], xs[i]] = vs[i]
        got = np.zeros(width * height, dtype=np.uint8)
        storage = np.zeros(((width + 7) // 8) * height, dtype=np.uint8)
        f(width, height, xs.ctypes.data, ys.ctypes.data, vs.ctypes.data, n, got.ctypes.data, storage.ctypes.data)
        assert np.array_equal(got.reshape(height, width), values)
        assert storage.tobytes() == np.ascontiguousarray(pack_mask(values.astype(bool))).tobytes()
    assert values.sum() > 30
