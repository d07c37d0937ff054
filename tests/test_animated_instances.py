"""Animated assembly instances (multi-key TransformSequence): the traversal evaluates the instance
transform at the ray's absolute time (assemblytree.cpp:635-639).

The checker is oracle/_ref, which links the reference's OWN renderer/utility/transformsequence.cpp
(evaluate, interpolate, motion bounding boxes) and foundation's TransformInterpolator / fast_slerp.
The product consumes the reference-format trees that checker builds (as an in-tree integration
would pass the live AssemblyTree) plus the interpolator data of every animated item.

CPU tier: the host build of the product's flattener + traversal code; GPU tier: the kernels."""
import ctypes as C

import numpy as np
import pytest

import cases
from appleseed_b200 import _lib, scenes
from appleseed_b200.scene import CItemMotion, InstanceKeys


def animated_case(seed=4):
    desc, rays, probes = cases.case_c3()
    rng = np.random.default_rng(seed)
    for r in (rays, probes):
        r.time_absolute = (rng.random(len(r)) * 1.3 - 0.15).astype(np.float32)       # also before the first / after the last key
        r.time_absolute[:64] = np.float32(0.4)                                          # exactly on a key
        r.time_normalized = np.zeros(len(r), dtype=np.float32)
    keys = {}
    for i, (a, b) in {0: (0.6, 1.3), 3: (-0.9, 0.4), 5: (0.2, 2.5), 7: (1.1, -1.7)}.items():
        base = desc.assembly_instances[i].local_to_parent
        mats = [base,
                scenes.translation(0.3, 0.1, -0.2) @ base @ scenes.rotation_y(a),
                scenes.translation(0.5, 0.3, 0.1) @ base @ scenes.rotation_y(b) @ scenes.scaling(1.2, 0.9, 1.1)]
        keys[i] = InstanceKeys([0.0, 0.4, 1.0] if i != 7 else [0.1, 0.4, 0.85], np.stack(mats))
    keys[2] = InstanceKeys([0.0, 1.0], np.stack([desc.assembly_instances[2].local_to_parent,
                                                   scenes.translation(0.0, 0.4, 0.0) @ desc.assembly_instances[2].local_to_parent]))
    return desc, rays, probes, keys


def product_views(asref, oscene, desc):
    """Reference-format trees + item motion of an asref scene as the product's view structs."""
    from oracle.oracle import AssemblyTreeView, TriangleTreeView
    keep, views = [], []
    for i in range(oscene.tree_count):
        v = TriangleTreeView()
        asref._get_tt(oscene.handle, i, C.byref(v))
        w = _lib.TriangleTreeView()
        for f, _t in _lib.TriangleTreeView._fields_:
            setattr(w, f, getattr(v, f))
        views.append(w)
    av = AssemblyTreeView()
    asref._get_at(oscene.handle, C.byref(av))
    n = int(av.item_count)
    inst = np.frombuffer((C.c_uint32 * n).from_address(av.item_assembly_instance), dtype=np.uint32)
    tree = np.frombuffer((C.c_uint32 * n).from_address(av.item_tree), dtype=np.uint32)
    items = (_lib.AssemblyItem * max(1, n))()
    motion = (CItemMotion * max(1, n))()
    for k in range(n):
        ai = desc.assembly_instances[int(inst[k])]
        items[k].parent_to_local[:] = oscene.item_parent_to_local(k).tolist()
        items[k].assembly_instance = int(inst[k])
        items[k].triangle_tree = int(tree[k])
        items[k].vis_flags = ai.vis_flags & 0xFFFFFFFF
        asref._item_motion(oscene.handle, k, C.byref(motion[k]))       # pointers into the checker's own storage
    top = _lib.AssemblyTreeView()
    top.nodes, top.items, top.node_count, top.item_count = av.nodes, C.cast(items, C.POINTER(_lib.AssemblyItem)), av.node_count, n
    top.item_motion = C.cast(motion, C.c_void_p)
    keep += [items, motion, oscene]
    return views, top, keep


def check(ref, got_exact, got_wide, probes_ref, pexact, pwide):
    assert got_exact.tobytes() == ref.tobytes()
    assert np.array_equal(pexact, probes_ref)
    # Wide kernels: identical except exact-t ties (there are none in this scene, but the rule stands).
    same = (got_wide.view(np.uint8).reshape(len(ref), -1) == ref.view(np.uint8).reshape(len(ref), -1)).all(axis=1)
    for k in ("t", "u", "v", "prim_type"):
        assert np.array_equal(got_wide[k][~same], ref[k][~same]), k
    assert (~same).sum() <= 5
    assert (pwide != probes_ref).sum() <= 2


def test_animation_changes_the_picture(asref):
    desc, rays, probes, keys = animated_case()
    static, moving = asref.scene(desc).trace(rays, threads=4), asref.scene(desc, keys=keys).trace(rays, threads=4)
    assert (static["t"] != moving["t"]).sum() > 1000
    at_key = rays.slice(0, 64)                                     # time 0.4 = the middle key: evaluate == that key's transform
    assert asref.scene(desc, keys=keys).trace(at_key, threads=1).tobytes() == moving[:64].tobytes()


def test_host_build_of_product_code(asref):
    from hostsim import hostsim
    desc, rays, probes, keys = animated_case()
    o = asref.scene(desc, keys=keys)
    views, top, keep = product_views(asref, o, desc)
    sim = hostsim.SimScene.from_views(hostsim.load(), views, top, keep)
    check(o.trace(rays, threads=4), sim.trace(rays, wide=False)[0], sim.trace(rays, wide=True)[0],
          o.trace_probe(probes, threads=4), sim.trace_probe(probes, wide=False)[0], sim.trace_probe(probes, wide=True)[0])


@pytest.mark.gpu
def test_kernels_with_animated_instances(asref):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from appleseed_b200.intersector import AsgpuError, Intersector, TraceContext
    desc, rays, probes, keys = animated_case()
    o = asref.scene(desc, keys=keys)
    views, top, keep = product_views(asref, o, desc)
    ctx = TraceContext.from_tree_views(views, top)
    isect = Intersector(ctx)
    ref = o.trace(rays, threads=4)
    check(ref, isect.trace(rays, exact=True), isect.trace(rays), o.trace_probe(probes, threads=4),
          isect.trace_probe(probes, exact=True), isect.trace_probe(probes))
    with pytest.raises(AsgpuError, match="animated|source geometry"):
        isect.refine_and_offset(rays, ref)


def _random_rigid(rng, spread):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    r = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    m = np.eye(4)
    m[:3, :3] = r @ np.diag(rng.uniform(0.5, 1.6, size=3))
    m[:3, 3] = rng.uniform(-spread, spread, size=3)
    return m


@pytest.mark.parametrize("seed", list(range(12)))
def test_random_animated_scenes_on_the_host_build(asref, seed):
    """Random static scenes (tests/cases.py::random_scene) whose assembly instances get 2-5 random
    keys (arbitrary rotations, non-uniform scales, translations, uneven key times)."""
    from hostsim import hostsim
    desc, rays = cases.random_scene(seed, moving=False)
    rng = np.random.default_rng(500 + seed)
    rays.time_absolute = (rng.random(len(rays)) * 1.4 - 0.2).astype(np.float32)
    rays.time_normalized = np.zeros(len(rays), dtype=np.float32)
    keys = {}
    for i in range(len(desc.assembly_instances)):
        if rng.random() < 0.7:
            k = int(rng.integers(2, 6))
            times = np.sort(rng.random(k)).astype(np.float32)
            times += np.arange(k, dtype=np.float32) * np.float32(1e-3)             # strictly ascending
            mats = [desc.assembly_instances[i].local_to_parent] + [_random_rigid(rng, 2.5) for _ in range(k - 1)]
            keys[i] = InstanceKeys(times, np.stack(mats))
    if not keys:
        pytest.skip("no animated instance drawn")
    o = asref.scene(desc, keys=keys)
    views, top, keep = product_views(asref, o, desc)
    sim = hostsim.SimScene.from_views(hostsim.load(), views, top, keep)
    ref = o.trace(rays, threads=2)
    assert sim.trace(rays, wide=False)[0].tobytes() == ref.tobytes()
    wide = sim.trace(rays, wide=True)[0]
    same = (wide.view(np.uint8).reshape(len(ref), -1) == ref.view(np.uint8).reshape(len(ref), -1)).all(axis=1)
    for k in ("t", "u", "v", "prim_type"):                     # coincident triangles / instances tie exactly
        assert np.array_equal(wide[k][~same], ref[k][~same]), k
    pref = o.trace_probe(rays, threads=2)
    assert np.array_equal(sim.trace_probe(rays, wide=False)[0], pref)
    assert (sim.trace_probe(rays, wide=True)[0] != pref).sum() <= 2


@pytest.mark.gpu
@pytest.mark.parametrize("seed", list(range(8)))
def test_random_animated_scenes_on_the_kernels(asref, seed):
    from appleseed_b200.intersector import Intersector, TraceContext
    desc, rays = cases.random_scene(seed, n_rays=20000, moving=False)
    rng = np.random.default_rng(500 + seed)
    rays.time_absolute = (rng.random(len(rays)) * 1.4 - 0.2).astype(np.float32)
    rays.time_normalized = np.zeros(len(rays), dtype=np.float32)
    keys = {}
    for i in range(len(desc.assembly_instances)):
        if rng.random() < 0.7:
            k = int(rng.integers(2, 6))
            times = np.sort(rng.random(k)).astype(np.float32)
            times += np.arange(k, dtype=np.float32) * np.float32(1e-3)
            mats = [desc.assembly_instances[i].local_to_parent] + [_random_rigid(rng, 2.5) for _ in range(k - 1)]
            keys[i] = InstanceKeys(times, np.stack(mats))
    if not keys:
        pytest.skip("no animated instance drawn")
    o = asref.scene(desc, keys=keys)
    views, top, keep = product_views(asref, o, desc)
    isect = Intersector(TraceContext.from_tree_views(views, top))
    ref = o.trace(rays, threads=4)
    assert isect.trace(rays, exact=True).tobytes() == ref.tobytes()
    wide = isect.trace(rays)
    same = (wide.view(np.uint8).reshape(len(ref), -1) == ref.view(np.uint8).reshape(len(ref), -1)).all(axis=1)
    for k in ("t", "u", "v", "prim_type"):
        assert np.array_equal(wide[k][~same], ref[k][~same]), k
    pref = o.trace_probe(rays, threads=4)
    assert np.array_equal(isect.trace_probe(rays, exact=True), pref)
    assert (isect.trace_probe(rays) != pref).sum() <= 2
