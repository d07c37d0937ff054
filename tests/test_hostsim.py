"""CPU-only check of the engine's traversal LOGIC: the same per-ray code the CUDA kernels run
(appleseed_b200/csrc/traverse_core.h), compiled for the host by tests/hostsim, against the oracle.

* EXACT layout: hit records and probe results byte-identical, and identical traversal counters
  (same visit order as the reference).
* WIDE layout: the north-star parity rule (tests/parity.py).
The GPU tests (-m gpu) repeat these comparisons on the real kernels through the C ABI."""
import numpy as np
import pytest

import cases
import kat
import parity
from hostsim import hostsim


@pytest.fixture(scope="module")
def sim():
    return hostsim.load()


@pytest.mark.parametrize("name", list(cases.CASES))
def test_exact_layout_is_bit_identical(sim, orc, name):
    desc, rays, probes = cases.CASES[name]()
    o = orc.scene(desc)
    s = hostsim.SimScene(sim, desc)
    ref, cref = o.trace(rays, threads=4, counters=True)
    got, cnt = s.trace(rays, wide=False)
    assert got.tobytes() == ref.tobytes()
    assert [int(x) for x in cnt[1:5]] == [cref[k] for k in ("assembly_nodes_visited", "instances_visited", "triangle_nodes_visited", "triangles_tested")]
    pref = o.trace_probe(probes, threads=4)
    pgot, _ = s.trace_probe(probes, wide=False)
    assert np.array_equal(pgot, pref)


@pytest.mark.parametrize("name", list(cases.CASES))
def test_wide_layout_meets_parity_rule(sim, orc, name):
    desc, rays, probes = cases.CASES[name]()
    o = orc.scene(desc)
    s = hostsim.SimScene(sim, desc)
    got, _ = s.trace(rays, wide=True)
    parity.compare_hits(o, rays, got, o.trace(rays, threads=4))
    pgot, _ = s.trace_probe(probes, wide=True)
    parity.compare_probes(o, probes, pgot, o.trace_probe(probes, threads=4))


def test_wide_only_scene(sim, orc):
    desc, rays, _ = cases.case_c3()
    s = hostsim.SimScene(sim, desc, flags=hostsim.SCENE_WIDE)
    o = orc.scene(desc)
    parity.compare_hits(o, rays, s.trace(rays, wide=True)[0], o.trace(rays, threads=4))


def test_axis_parallel_and_negative_tmin_rays(sim, orc):
    # Directions with zero components (rcp = +-inf, NaN plane distances) and rays whose interval
    # starts before the origin exercise the interval arithmetic of the wide box test.
    from appleseed_b200 import scenes
    from appleseed_b200.scene import RayBatch
    desc = scenes.scene_c2(40)
    o = orc.scene(desc)
    s = hostsim.SimScene(sim, desc)
    rng = np.random.default_rng(3)
    n = 6000
    org = np.stack([rng.uniform(-1.2, 1.2, n), rng.uniform(-0.5, 1.0, n), rng.uniform(-1.2, 1.2, n)], 1)
    axes = np.array([[0, -1, 0], [0, 1, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1], [0, -1, -0.0], [-0.0, -1, 0]], dtype=np.float64)
    d = axes[rng.integers(0, len(axes), n)]
    # snap some origins exactly onto grid lines so that plane distances are exactly zero
    org[::3, 0] = np.round(org[::3, 0] * 20) / 20
    org[::3, 2] = np.round(org[::3, 2] * 20) / 20
    rays = RayBatch(org, d, rng.choice([0.0, -0.7, -3.0], n), rng.choice([scenes.DBL_MAX, 2.0], n))
    ref = o.trace(rays, threads=4)
    assert s.trace(rays, wide=False)[0].tobytes() == ref.tobytes()
    parity.compare_hits(o, rays, s.trace(rays, wide=True)[0], ref)
    parity.compare_probes(o, rays, s.trace_probe(rays, wide=True)[0], o.trace_probe(rays, threads=4))


def test_reference_known_answers(sim):
    s = hostsim.SimScene(sim, kat.tracer_scene([2.0, 4.0]))
    for wide in (False, True):
        h, _ = s.trace(kat.x_ray(), wide=wide)
        assert h["t"][0] == 2.0 and h["assembly_instance"][0] == 0 and h["prim_type"][0] == 2
        assert s.trace_probe(kat.x_ray(), wide=wide)[0][0] == 1
    s = hostsim.SimScene(sim, kat.tracer_scene([2.0], scale=0.5))
    for wide in (False, True):
        assert abs(s.trace(kat.x_ray(), wide=wide)[0]["t"][0] - 1.0) <= 1e-15
    s = hostsim.SimScene(sim, kat.empty_bbox_scene())
    for wide in (False, True):
        h, _ = s.trace(kat.empty_bbox_ray(), wide=wide)
        assert h["prim_type"][0] == 0 and h["t"][0] == 2.0
        assert s.trace_probe(kat.empty_bbox_ray(), wide=wide)[0][0] == 0


@pytest.mark.parametrize("name", ["c3", "mixed"])
def test_extreme_and_non_finite_rays(sim, orc, name):
    desc, rays, _ = cases.CASES[name]()
    o, s = orc.scene(desc), hostsim.SimScene(sim, desc)
    ok = cases.extreme_rays(rays.slice(0, 6000), cases.EXTREME_OK, 5)
    ref = o.trace(ok, threads=4)
    assert s.trace(ok, wide=False)[0].tobytes() == ref.tobytes()
    parity.compare_hits(o, ok, s.trace(ok, wide=True)[0], ref)
    pref = o.trace_probe(ok, threads=4)
    assert np.array_equal(s.trace_probe(ok, wide=False)[0], pref)
    parity.compare_probes(o, ok, s.trace_probe(ok, wide=True)[0], pref)
    assert (ref["prim_type"] == 2).sum() > 300
    # No meaning as rays: the exact layout still repeats the reference bit for bit, the wide one terminates.
    bad = cases.extreme_rays(rays.slice(0, 6000), cases.EXTREME_UNDEFINED, 6)
    assert s.trace(bad, wide=False)[0].tobytes() == o.trace(bad, threads=4).tobytes()
    assert np.array_equal(s.trace_probe(bad, wide=False)[0], o.trace_probe(bad, threads=4))
    assert len(s.trace(bad, wide=True)[0]) == len(bad) and len(s.trace_probe(bad, wide=True)[0]) == len(bad)


@pytest.mark.parametrize("name", ["c3", "c4_msc2", "mixed"])
def test_flattened_scene_does_not_depend_on_the_thread_count(sim, name, monkeypatch):
    """The host builder and the flattener run on several threads (triangle collection, leaf payloads,
    node decoding, the wide collapse level by level): chunks land at positions fixed by prefix sums, so
    the blob is the same byte for byte whatever ASGPU_HOST_THREADS says."""
    from appleseed_b200 import scenes
    desc = {"c3": lambda: scenes.scene_c3(120, 3), "c4_msc2": lambda: scenes.scene_c4(160, 2), "mixed": lambda: cases.case_mixed()[0]}[name]()
    blobs = []
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("ASGPU_HOST_THREADS", threads)
        blobs.append(hostsim.SimScene(sim, desc, threads=int(threads)).blob())
    assert blobs[0].tobytes() == blobs[1].tobytes() == blobs[2].tobytes()
    assert len(blobs[0]) > 100000
