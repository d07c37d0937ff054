"""Triangle trees built on the device (asgpu_trees_build_on_device):
parallel locally-ordered clustering (ploc.cu, the default) or a linear BVH (lbvh.cu,
ASGPU_DEVICE_BUILD=lbvh) over the Morton order instead of the reference's sweep SAH.  The TREE differs from the
reference's, the hit records must not: every valid BVH over the same triangles yields the same
nearest hit with bit-identical t, u, v (the triangle test is the reference's), exact-t ties aside.

CPU tier: a sequential host run of the same per-node code (ploc_core.h / lbvh_core.h) + the product's emission
into the reference node format, flattener and both traversals (tests/hostsim), against the oracle.
GPU tier: the kernels of ploc.cu / lbvh.cu through the C ABI; their tree must be the one the sequential run
builds (same counters and byte-identical records from the exact kernels), and results must meet the
parity rule against the oracle."""
import numpy as np
import pytest

import cases
import parity
from appleseed_b200 import scenes
from appleseed_b200.scene import Assembly, AssemblyInstance, Mesh, ObjectInstance, RayBatch, SceneDesc
from hostsim import hostsim


@pytest.fixture(scope="module")
def sim():
    return hostsim.load()


ALGORITHMS = {"ploc": dict(ploc=16), "ploc_r4": dict(ploc=4), "lbvh": dict(lbvh=True)}


def device_trees(desc, algorithm):
    """asgpu_trees_build_on_device with the algorithm the environment selects (read per call)."""
    import os
    from appleseed_b200.intersector import HostTrees
    saved = {k: os.environ.get(k) for k in ("ASGPU_DEVICE_BUILD", "ASGPU_PLOC_RADIUS")}
    try:
        os.environ.pop("ASGPU_DEVICE_BUILD", None); os.environ.pop("ASGPU_PLOC_RADIUS", None)
        if algorithm == "lbvh":
            os.environ["ASGPU_DEVICE_BUILD"] = "lbvh"
        elif algorithm == "ploc_r4":
            os.environ["ASGPU_PLOC_RADIUS"] = "4"
        return HostTrees(desc, build_device=0)
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


def check_against_oracle(o, rays, probes, exact, wide, pexact, pwide):
    ref = o.trace(rays, threads=4)
    se = parity.compare_hits(o, rays, exact, ref)
    sw = parity.compare_hits(o, rays, wide, ref)
    # Same triangles, same arithmetic: where the identity agrees, t / u / v are bit-identical.
    for got in (exact, wide):
        same = (got["prim_type"] == ref["prim_type"]) & (got["primitive_index"] == ref["primitive_index"]) & \
               (got["object_instance_index"] == ref["object_instance_index"]) & (got["assembly_instance"] == ref["assembly_instance"])
        for k in ("t", "u", "v", "motion_segment"):
            assert np.array_equal(got[k][same], ref[k][same]), k
    pref = o.trace_probe(probes, threads=4)
    parity.compare_probes(o, probes, pexact, pref)
    parity.compare_probes(o, probes, pwide, pref)
    return se, sw


@pytest.mark.parametrize("algorithm", list(ALGORITHMS))
@pytest.mark.parametrize("name", list(cases.CASES))
def test_device_topology_on_the_host_build(sim, orc, name, algorithm):
    desc, rays, probes = cases.CASES[name]()
    s = hostsim.SimScene(sim, desc, **ALGORITHMS[algorithm])
    check_against_oracle(orc.scene(desc), rays, probes, s.trace(rays, wide=False)[0], s.trace(rays, wide=True)[0],
                         s.trace_probe(probes, wide=False)[0], s.trace_probe(probes, wide=True)[0])


@pytest.mark.parametrize("algorithm", ["ploc", "lbvh"])
@pytest.mark.parametrize("seed", list(range(8)))
def test_device_topology_on_random_scenes(sim, orc, seed, algorithm):
    desc, rays = cases.random_scene(100 + seed, n_rays=3000)
    probes = rays
    s = hostsim.SimScene(sim, desc, **ALGORITHMS[algorithm])
    check_against_oracle(orc.scene(desc), rays, probes, s.trace(rays, wide=False)[0], s.trace(rays, wide=True)[0],
                         s.trace_probe(probes, wide=False)[0], s.trace_probe(probes, wide=True)[0])


def test_device_trees_differ_from_the_sweep_tree_but_not_in_results(sim, orc):
    desc, rays, _ = cases.case_c2()
    a, b, c = hostsim.SimScene(sim, desc), hostsim.SimScene(sim, desc, lbvh=True), hostsim.SimScene(sim, desc, ploc=16)
    (ha, ca), (hb, cb), (hc, cc) = a.trace(rays, wide=False), b.trace(rays, wide=False), c.trace(rays, wide=False)
    assert int(ca[3]) != int(cb[3]) and int(ca[3]) != int(cc[3])        # different trees: different node visits ...
    assert int(cb[3]) < 3 * int(ca[3])                                  # ... of comparable quality on a regular mesh
    assert int(cc[3]) < int(cb[3])                                      # clustering by surface area beats the Morton splits
    assert int(cc[3]) < 1.15 * int(ca[3])                               # ... and, with the sweep SAH on top, stays close to the sweep tree
    print("binary node visits: sweep SAH %d, clustering + sweep top %d, linear %d" % (int(ca[3]), int(cc[3]), int(cb[3])))
    for h in (hb, hc):
        for k in ("t", "u", "v", "primitive_index", "prim_type"):
            assert np.array_equal(ha[k], h[k]), k


def coincident_scene(copies):
    """`copies` identical triangles (equal Morton keys: the hierarchy tells them apart by position)
    in front of a few distinct ones."""
    v = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 1.0], [0.0, 1.0, 1.0],
                  [3.0, 3.0, 2.0], [4.0, 3.0, 2.0], [3.0, 4.0, 2.0]], dtype=np.float32)
    tris = [[0, 1, 2]] * copies + [[3, 4, 5], [6, 7, 8]]
    return SceneDesc([Mesh(v, np.array(tris, dtype=np.uint32))], [Assembly([ObjectInstance(0)])], [AssemblyInstance(0)])


@pytest.mark.parametrize("algorithm", ["ploc", "lbvh"])
@pytest.mark.parametrize("copies", [1, 2, 3, 37])
def test_coincident_triangles(sim, orc, copies, algorithm):
    desc = coincident_scene(copies)
    rng = np.random.default_rng(copies)
    org = np.column_stack([rng.uniform(-0.2, 1.2, 400), rng.uniform(-0.2, 1.2, 400), np.full(400, 5.0)])
    rays = RayBatch(org, np.tile([0.0, 0.0, -1.0], (400, 1)), 0.0, np.finfo(np.float64).max)
    s = hostsim.SimScene(sim, desc, **ALGORITHMS[algorithm])
    ref = orc.scene(desc).trace(rays, threads=2)
    for wide in (False, True):
        got = s.trace(rays, wide=wide)[0]
        for k in ("t", "u", "v", "prim_type"):                          # which of the identical copies wins is a tie
            assert np.array_equal(got[k], ref[k]), k
    assert (ref["prim_type"] == 2).sum() > 50


def test_tiny_trees_skip_the_device(sim, orc):
    # At most max_leaf_size triangles: a single leaf, no topology to build.
    desc, rays, probes = cases.case_cornell()
    one = SceneDesc([Mesh(desc.meshes[0].vertices, desc.meshes[0].triangles[:2])], [Assembly([ObjectInstance(0)])], [AssemblyInstance(0)])
    for how in (dict(lbvh=True), dict(ploc=16)):
        s = hostsim.SimScene(sim, one, **how)
        assert s.trace(rays, wide=False)[0].tobytes() == orc.scene(one).trace(rays, threads=2).tobytes()


def test_device_build_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from appleseed_b200.intersector import AsgpuError, HostTrees
    desc, _, _ = cases.case_cornell()
    with pytest.raises(AsgpuError, match="device"):
        HostTrees(desc, build_device=0)


@pytest.mark.gpu
@pytest.mark.parametrize("algorithm", list(ALGORITHMS))
@pytest.mark.parametrize("name", list(cases.CASES))
def test_device_build_matches_the_sequential_run_and_the_oracle(sim, orc, name, algorithm):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from appleseed_b200.intersector import Intersector, TraceContext
    desc, rays, probes = cases.CASES[name]()
    trees = device_trees(desc, algorithm)
    isect = Intersector(TraceContext(trees=trees, device=0))
    s = hostsim.SimScene(sim, desc, **ALGORITHMS[algorithm])
    isect.ctx.counters(reset=True)
    exact = isect.trace(rays, exact=True, counters=True)
    cnt = isect.ctx.counters(reset=True)
    want, wcnt = s.trace(rays, wide=False)
    assert exact.tobytes() == want.tobytes()                            # same tree as the sequential run: same slots, same order
    assert [cnt[k] for k in ("assembly_nodes_visited", "instances_visited", "triangle_nodes_visited", "triangles_tested")] == [int(x) for x in wcnt[1:5]]
    assert np.array_equal(isect.trace_probe(probes, exact=True), s.trace_probe(probes, wide=False)[0])
    check_against_oracle(orc.scene(desc), rays, probes, exact, isect.trace(rays),
                         isect.trace_probe(probes, exact=True), isect.trace_probe(probes))


@pytest.mark.gpu
@pytest.mark.parametrize("algorithm", ["ploc", "lbvh"])
def test_device_build_at_scale(orc, algorithm):
    """One million triangles: the device-built tree finds what the sweep-SAH tree finds."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from appleseed_b200.intersector import Intersector, TraceContext
    desc = scenes.scene_c2()
    _, rays, _ = cases.case_c2(n=200000)
    sah = Intersector(TraceContext(desc, device=0, flags=2))
    trees = device_trees(desc, algorithm)
    lin = Intersector(TraceContext(trees=trees, device=0, flags=2))
    a, b = sah.trace(rays), lin.trace(rays)
    same = a["primitive_index"] == b["primitive_index"]
    assert same.mean() > 0.9999
    for k in ("t", "u", "v", "prim_type"):
        assert np.array_equal(a[k][same], b[k][same]), k
    assert np.array_equal(a["prim_type"], b["prim_type"])
    np.testing.assert_allclose(a["t"][~same], b["t"][~same], rtol=1e-6)
    print("build seconds: sweep SAH %.3f, device %s %.3f" % (sah.ctx.build_seconds, algorithm, trees.build_seconds))
