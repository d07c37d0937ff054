"""asgpu_trees_build_animated: the assembly tree over ANIMATED assembly instances, built by the
product (motion_bounds.cpp: TransformInterpolator segments and TransformSequence::to_parent's motion
bounding boxes restated) against oracle/_ref, which links the reference's own
renderer/utility/transformsequence.cpp and foundation/math/transform.h:

* the assembly tree (nodes with their child boxes, item order) byte for byte;
* keys and interpolator segments (scale, quaternion, translation at both ends) bit for bit;
* rays through the product-built scene (host build of the product's flattener and traversal code)
  give the records of the reference traversal.
"""
import ctypes as C

import numpy as np
import pytest

import cases
from appleseed_b200 import _lib, scenes
from appleseed_b200.intersector import HostTrees
from appleseed_b200.scene import CItemMotion, InstanceKeys
from hostsim import hostsim
from test_animated_instances import animated_case, check


def product_motion(top, k):
    m = C.cast(top.item_motion, C.POINTER(CItemMotion))[k]
    n = int(m.key_count)
    if n < 2:
        return None
    as_np = lambda ptr, count, dt: np.frombuffer((C.c_uint8 * (count * np.dtype(dt).itemsize)).from_address(ptr), dtype=dt).copy()
    return as_np(m.key_times, n, np.float32), as_np(m.key_parent_to_local, n * 16, np.float64).reshape(n, 16), as_np(m.segments, (n - 1) * 20, np.float64).reshape(n - 1, 20)


def compare_trees(asref, desc, keys):
    o = asref.scene(desc, keys)
    trees = HostTrees(desc, threads=4, keys=keys)
    top = trees.assembly_tree_view()
    want = o.assembly_tree()
    n = int(top.item_count)
    assert n == len(want["item_assembly_instance"])
    got_nodes = np.frombuffer((C.c_uint8 * (int(top.node_count) * 128)).from_address(top.nodes), dtype=np.uint8)
    assert got_nodes.tobytes() == want["nodes"].tobytes()                       # same boxes => same SAH tree, same child boxes
    animated = 0
    for k in range(n):
        assert int(top.items[k].assembly_instance) == int(want["item_assembly_instance"][k])
        assert int(top.items[k].triangle_tree) == int(want["item_tree"][k])
        ref = o.item_motion(k)
        got = product_motion(top, k) if top.item_motion else None
        assert (ref is None) == (got is None)
        if ref is not None:
            animated += 1
            for a, b in zip(ref, got):
                assert a.tobytes() == b.tobytes()
    assert animated == sum(1 for i, kk in keys.items() if len(kk.times) >= 2 and desc.assemblies[desc.assembly_instances[i].assembly_index].object_instances)
    return o, trees


def test_assembly_tree_over_animated_instances_is_the_references(asref):
    desc, rays, probes, keys = animated_case()
    o, trees = compare_trees(asref, desc, keys)
    # ... and it is not the tree of the same scene standing still.
    top, still = trees.assembly_tree_view(), HostTrees(desc, threads=4).assembly_tree_view()
    nodes = lambda v: np.frombuffer((C.c_uint8 * (int(v.node_count) * 128)).from_address(v.nodes), dtype=np.uint8).tobytes()
    assert nodes(top) != nodes(still) and not still.item_motion


def _random_rigid(rng, spread):
    axis = rng.normal(size=3); axis /= np.linalg.norm(axis)
    angle = rng.uniform(-3.0, 3.0)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)
    m = np.eye(4); m[:3, :3] = R; m[:3, 3] = rng.uniform(-spread, spread, 3)
    return m


@pytest.mark.parametrize("seed", list(range(12)))
def test_random_key_sequences(asref, seed):
    """Rotations about arbitrary axes up to +-3 rad per segment (several Newton sub-intervals),
    non-uniform and mirrored scalings, 2..5 keys, pure translations (angle == 0: no extremum search)
    and identical keys."""
    rng = np.random.default_rng(900 + seed)
    desc, rays = cases.random_scene(300 + seed, n_rays=800)
    keys = {}
    for i in range(len(desc.assembly_instances)):
        if rng.random() < 0.25:
            continue
        n = int(rng.integers(2, 6))
        times = np.sort(rng.choice(np.linspace(0.0, 1.0, 41), size=n, replace=False))
        base = desc.assembly_instances[i].local_to_parent
        mats = [base]
        kind = rng.integers(0, 4)
        for _ in range(n - 1):
            if kind == 0:
                mats.append(scenes.translation(*rng.uniform(-1, 1, 3)) @ base)
            elif kind == 1:
                mats.append(mats[-1].copy())
            else:
                s = rng.uniform(0.6, 1.5, 3) * (np.array([-1.0, 1.0, 1.0]) if (kind == 3 and rng.random() < 0.5) else 1.0)
                mats.append(_random_rigid(rng, 1.5) @ base @ scenes.scaling(*s))
        keys[i] = InstanceKeys(times, np.stack(mats))
    if not keys:
        keys[0] = InstanceKeys([0.0, 1.0], np.stack([desc.assembly_instances[0].local_to_parent, _random_rigid(rng, 1.0) @ desc.assembly_instances[0].local_to_parent]))
    o, trees = compare_trees(asref, desc, keys)


def test_traversal_through_product_built_animated_trees(asref):
    """The trees of asgpu_trees_build_animated through the product's flattener and traversal code
    (host build) against the reference traversal of the same animated scene."""
    sim = hostsim.load()
    desc, rays, probes, keys = animated_case()
    o, trees = compare_trees(asref, desc, keys)
    views = [trees.triangle_tree_view(i) for i in range(trees.triangle_tree_count)]
    top = trees.assembly_tree_view()
    s = hostsim.SimScene.from_views(sim, views, top, [trees])
    ref, pref = o.trace(rays, threads=4), o.trace_probe(probes, threads=4)
    check(ref, s.trace(rays, wide=False)[0], s.trace(rays, wide=True)[0], pref, s.trace_probe(probes, wide=False)[0], s.trace_probe(probes, wide=True)[0])


def test_bad_keys_are_refused():
    desc, _, _ = cases.case_c3()
    base = desc.assembly_instances[0].local_to_parent
    k = InstanceKeys([0.0, 1.0], np.stack([base, base]))
    k.times = np.array([0.5, 0.5], dtype=np.float32)
    from appleseed_b200.intersector import AsgpuError
    with pytest.raises(AsgpuError, match="ascend"):
        HostTrees(desc, keys={0: k})
    k = InstanceKeys([0.0, 1.0], np.stack([base, base]))
    k.local_to_parent[1, 0, 0] = np.nan
    with pytest.raises(AsgpuError, match="finite"):
        HostTrees(desc, keys={0: k})


@pytest.mark.gpu
def test_kernels_on_product_built_animated_trees(asref):
    """Description + keys -> asgpu_trees_build_animated -> asgpu_scene_create_ex -> both kernels,
    refine_and_offset included (the trees carry their source geometry), against the reference."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from appleseed_b200.intersector import Intersector, TraceContext
    desc, rays, probes, keys = animated_case()
    o = asref.scene(desc, keys=keys)
    isect = Intersector(TraceContext(trees=HostTrees(desc, keys=keys), device=0))
    ref = o.trace(rays, threads=4)
    check(ref, isect.trace(rays, exact=True), isect.trace(rays), o.trace_probe(probes, threads=4),
          isect.trace_probe(probes, exact=True), isect.trace_probe(probes))
    assert isect.refine_and_offset(rays, ref).tobytes() == o.refine_offset(rays, ref, threads=4).tobytes()
