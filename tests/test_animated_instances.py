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
    with pytest.raises(AsgpuError, match="source geometry"):        # asgpu_scene_create: no source geometry
        isect.refine_and_offset(rays, ref)


def product_sources(desc):
    """Source geometry per triangle tree (what asgpu_scene_create_ex takes), from the product's own
    host builder: the triangle trees do not depend on the instance animation."""
    from appleseed_b200.intersector import HostTrees
    trees = HostTrees(desc)
    return [trees.source_geometry(i) for i in range(trees.triangle_tree_count)], trees


def child_rays(rays, hits, seed):
    """Bounce rays from the un-offset world hit points, at their parent ray's time."""
    from appleseed_b200.scene import RayBatch
    h = hits["prim_type"] == 2
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(int(h.sum()), 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = rays.org[h] + hits["t"][h][:, None] * rays.dir[h]
    return h, RayBatch(pts, d, 0.0, np.finfo(np.float64).max, time_absolute=rays.time_absolute[h],
                       time_normalized=rays.time_normalized[h], flags=rays.flags[h] if rays.flags is not None else None)


def test_refine_and_offset_with_animated_instances_host_build(asref):
    """ShadingPoint::refine_and_offset works in the space of m_assembly_instance_transform, which the
    traversal set to the sequence evaluated at the ray time (assemblytree.cpp:738-739)."""
    from hostsim import hostsim
    desc, rays, probes, keys = animated_case()
    o = asref.scene(desc, keys=keys)
    views, top, keep = product_views(asref, o, desc)
    sources, trees = product_sources(desc)
    sim = hostsim.SimScene.from_views(hostsim.load(), views, top, keep + [trees], sources=sources)
    hits = o.trace(rays, threads=4)
    ref = o.refine_offset(rays, hits, threads=4)
    assert sim.refine_offset(rays, hits).tobytes() == ref.tobytes()
    # The animation matters: the same hits refined under the first key's transform differ.
    frozen = asref.scene(desc).refine_offset(rays, hits, threads=4)
    assert (frozen["front"] != ref["front"]).any(axis=1).sum() > 500
    # Child rays with their parents, through both traversals of the product code.
    h, bounce = child_rays(rays, hits, 11)
    par = ref[h]
    want = o.trace_parents(bounce, par, threads=4)
    assert sim.trace_parents(bounce, par, wide=False).tobytes() == want.tobytes()
    wide = sim.trace_parents(bounce, par, wide=True)
    assert (wide["t"] != want["t"]).sum() <= 5
    assert int(((want["prim_type"] == 2) & (want["t"] < 1e-9)).sum()) == 0                 # no self-intersection
    assert int(((o.trace(bounce, threads=4)["t"] < 1e-9)).sum()) > 0.2 * len(bounce)      # ... which a parentless child ray has


@pytest.mark.gpu
def test_refine_and_offset_with_animated_instances_on_the_kernels(asref):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from appleseed_b200.intersector import Intersector, TraceContext
    desc, rays, probes, keys = animated_case()
    o = asref.scene(desc, keys=keys)
    views, top, keep = product_views(asref, o, desc)
    sources, trees = product_sources(desc)
    isect = Intersector(TraceContext.from_tree_views(views, top, sources=sources))
    hits = isect.trace(rays, exact=True)
    assert hits.tobytes() == o.trace(rays, threads=4).tobytes()
    ref = o.refine_offset(rays, hits, threads=4)
    assert isect.refine_and_offset(rays, hits).tobytes() == ref.tobytes()
    h, bounce = child_rays(rays, hits, 11)
    want = o.trace_parents(bounce, ref[h], threads=4)
    assert isect.trace_with_parents(bounce, ref[h], exact=True).tobytes() == want.tobytes()
    assert (isect.trace_with_parents(bounce, ref[h])["t"] != want["t"]).sum() <= 5


# renderer/meta/tests/test_transformsequence.cpp:335-406 (TwoTransformsFixture): keys at times 1 and 3,
# translations (1,2,3) and (4,5,6); evaluate(0) = the first key, evaluate(4) = the last,
# evaluate(2) = the interpolator at 0.5 = translation (2.5, 3.5, 4.5) (EXPECT_FEQ).  Seen through the
# path: a unit quad in the x = 0 plane of the instance, rays along +x from (0, y, z) of the key.
TS_KAT = [(0.0, (1.0, 2.0, 3.0), True), (1.0, (1.0, 2.0, 3.0), True), (4.0, (4.0, 5.0, 6.0), True), (3.0, (4.0, 5.0, 6.0), True),
          (2.0, (2.5, 3.5, 4.5), False)]


def transformsequence_kat():
    import kat
    from appleseed_b200.scene import Assembly, AssemblyInstance, ObjectInstance, RayBatch, SceneDesc
    first, second = scenes.translation(1.0, 2.0, 3.0), scenes.translation(4.0, 5.0, 6.0)
    desc = SceneDesc([kat.unit_quad()], [Assembly([ObjectInstance(0)])], [AssemblyInstance(0, first)])
    # (:384-406 sets the keys in reverse order: TransformSequence::prepare sorts them, upstream of the
    # boundary -- the engine receives keys in time order.)
    keys = {0: InstanceKeys([1.0, 3.0], np.stack([first, second]))}
    org = np.array([[0.0, p[1], p[2]] for _, p, _ in TS_KAT])
    rays = RayBatch(org, np.tile([1.0, 0.0, 0.0], (len(TS_KAT), 1)), 0.0, np.finfo(np.float64).max,
                    time_absolute=np.array([t for t, _, _ in TS_KAT], dtype=np.float32), time_normalized=np.zeros(len(TS_KAT), dtype=np.float32))
    return desc, keys, rays


def check_transformsequence_kat(hits):
    assert np.all(hits["prim_type"] == 2)
    for h, (_, p, exact) in zip(hits, TS_KAT):
        assert h["t"] == p[0] if exact else abs(h["t"] - p[0]) <= 1e-14 * p[0]
        assert abs(h["u"] + h["v"] - 0.5) <= 1e-6 or abs(h["u"] - 0.5) <= 1e-6 or abs(h["v"] - 0.5) <= 1e-6     # the quad's centre lies on its diagonal


def test_transformsequence_known_answers(asref):
    from hostsim import hostsim
    desc, keys, rays = transformsequence_kat()
    o = asref.scene(desc, keys=keys)
    ref = o.trace(rays)
    check_transformsequence_kat(ref)
    views, top, keep = product_views(asref, o, desc)
    sim = hostsim.SimScene.from_views(hostsim.load(), views, top, keep)
    assert sim.trace(rays, wide=False)[0].tobytes() == ref.tobytes()
    check_transformsequence_kat(sim.trace(rays, wide=True)[0])
    assert list(o.trace_probe(rays)) == [1] * len(rays) == list(sim.trace_probe(rays, wide=True)[0])


@pytest.mark.gpu
def test_transformsequence_known_answers_on_the_kernels(asref):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from appleseed_b200.intersector import Intersector, TraceContext
    desc, keys, rays = transformsequence_kat()
    o = asref.scene(desc, keys=keys)
    views, top, keep = product_views(asref, o, desc)
    isect = Intersector(TraceContext.from_tree_views(views, top))
    assert isect.trace(rays, exact=True).tobytes() == o.trace(rays).tobytes()
    check_transformsequence_kat(isect.trace(rays))
    assert list(isect.trace_probe(rays)) == [1] * len(rays) == list(isect.trace_probe(rays, exact=True))


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
    if not keys:                # the draw animated no instance (one seed in eight): animate the first one
        times = np.array([0.2, 0.7], dtype=np.float32)
        keys[0] = InstanceKeys(times, np.stack([desc.assembly_instances[0].local_to_parent, _random_rigid(rng, 2.5)]))
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
    if not keys:                # the draw animated no instance (one seed in eight): animate the first one
        times = np.array([0.2, 0.7], dtype=np.float32)
        keys[0] = InstanceKeys(times, np.stack([desc.assembly_instances[0].local_to_parent, _random_rigid(rng, 2.5)]))
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


# foundation/meta/tests/test_transform.cpp:240-301 (TransformInterpolator), seen through the path: a unit
# quad in the x = 1 plane of the instance (object instance translated by (1, 0, 0)), rays along +x.
def interpolator_kat(first, second):
    import kat
    from appleseed_b200.scene import Assembly, AssemblyInstance, ObjectInstance, SceneDesc
    desc = SceneDesc([kat.unit_quad()], [Assembly([ObjectInstance(0, scenes.translation(1.0, 0.0, 0.0))])], [AssemblyInstance(0, first)])
    return desc, {0: InstanceKeys([0.0, 1.0], np.stack([first, second]))}


def _x_rays(points, time):
    from appleseed_b200.scene import RayBatch
    org = np.array(points, dtype=np.float64)
    return RayBatch(org, np.tile([1.0, 0.0, 0.0], (len(org), 1)), 0.0, np.finfo(np.float64).max,
                    time_absolute=np.full(len(org), time, dtype=np.float32), time_normalized=np.zeros(len(org), dtype=np.float32))


def _both(asref, desc, keys, rays):
    from hostsim import hostsim
    o = asref.scene(desc, keys=keys)
    ref = o.trace(rays)
    views, top, keep = product_views(asref, o, desc)
    sim = hostsim.SimScene.from_views(hostsim.load(), views, top, keep)
    assert sim.trace(rays, wide=False)[0].tobytes() == ref.tobytes()
    wide = sim.trace(rays, wide=True)[0]
    assert np.array_equal(wide["prim_type"], ref["prim_type"]) and np.array_equal(wide["t"], ref["t"])
    return ref


def test_interpolator_scaling_known_answer(asref):
    # :240-251: identity -> scaling (3, 5, 0.6), evaluated at 0.5, has scaling (2, 3, 0.8) (EXPECT_FEQ).
    desc, keys = interpolator_kat(np.eye(4), scenes.scaling(3.0, 5.0, 0.6))
    rays = _x_rays([[0, 0, 0], [0, 1.49, 0], [0, 1.51, 0], [0, 0, 0.39], [0, 0, 0.41]], 0.5)
    h = _both(asref, desc, keys, rays)
    assert list(h["prim_type"]) == [2, 2, 0, 2, 0]                      # the quad spans +-0.5 * 3 in y and +-0.5 * 0.8 in z
    assert np.all(np.abs(h["t"][h["prim_type"] == 2] - 2.0) <= 1e-14)   # and sits at x = 1 * 2


def test_interpolator_mirroring_known_answers(asref):
    # :253-270: identical mirrored from / to matrices -> the input transform (EXPECT_FEQ).
    mirror = np.array([[-1.0, 0, 0, 4.0], [0, 1.0, 0, 6.0], [0, 0, 1.0, 8.0], [0, 0, 0, 1.0]])
    desc, keys = interpolator_kat(mirror, mirror)
    h = _both(asref, desc, keys, _x_rays([[0, 6, 8], [0, 6.49, 8.49], [0, 6.51, 8]], 0.5))
    assert list(h["prim_type"]) == [2, 2, 0] and np.all(np.abs(h["t"][:2] - 3.0) <= 1e-14)     # x = 4 - 1
    # :272-301: real-world matrices with mirroring, evaluated at 0.02432: a valid transform (its two
    # halves are inverses of each other to 1e-9) -- here: a ray aimed at the image of the quad's
    # centre under the evaluated local_to_parent hits it in the middle.
    a = np.array([[-0.99702130594062099, 0.077087478205260004, -0.0025112020321310003, 0.76365516185966298],
                  [0.077096829499781999, 0.99701625983943298, -0.0039071886356200000, 146.27250157945070],
                  [-0.0022025165107730001, 0.0040890970283240001, 0.99998953373952104, 8.9871358181588690],
                  [0.0, 0.0, 0.0, 1.0]])
    b = np.array([[-0.99665231953250000, 0.081673207457143002, -0.0036906925368350003, 0.65271732495429602],
                  [0.081677812416244000, 0.99665822601435605, -0.0011578872622540000, 145.34519371726373],
                  [-0.0035839072864200005, 0.0014554931967890000, 0.99999257746504999, 9.0022056937678983],
                  [0.0, 0.0, 0.0, 1.0]])
    desc, keys = interpolator_kat(a, b)
    t = np.float32(0.024320000000000008)
    # The centre of the quad in world space (the keys are 2.4 % apart in time and nearly equal, so a
    # linear blend of its two images locates it well within the quad): shoot at it from 10 units away.
    centre_guess = (1.0 - t) * (a @ [1, 0, 0, 1])[:3] + t * (b @ [1, 0, 0, 1])[:3]
    rays = _x_rays([[centre_guess[0] - 10.0, centre_guess[1], centre_guess[2]]], t)
    rays.dir[0] = [1.0, 0.0, 0.0]
    h = _both(asref, desc, keys, rays)
    assert h["prim_type"][0] == 2 and abs(h["t"][0] - 10.0) < 1e-3
    assert abs(h["u"][0] + h["v"][0] - 0.5) < 5e-3 or abs(max(h["u"][0], h["v"][0]) - 0.5) < 5e-3       # near the quad's centre (on its diagonal)
