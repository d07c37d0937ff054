"""The optional coherence sort (ASGPU_TRACE_SORT / asgpu_sort_rays): it may only change the ORDER
in which rays are processed, never a result."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from appleseed_b200 import intersector
    return intersector


@pytest.mark.parametrize("name", ["c2", "c3", "c4_msc1", "mixed"])
def test_sorted_trace_is_byte_identical(engine, name):
    desc, rays, probes = cases.CASES[name]()
    ctx = engine.TraceContext(desc, device=0)
    isect = engine.Intersector(ctx)
    assert isect.trace(rays, sort=True).tobytes() == isect.trace(rays).tobytes()
    assert np.array_equal(isect.trace_probe(probes, sort=True), isect.trace_probe(probes))
    assert isect.trace(rays, sort=True, exact=True).tobytes() == isect.trace(rays, exact=True).tobytes()


@pytest.mark.parametrize("n", [1, 31, 2048, 2049, 100003])
def test_sort_is_a_stable_permutation_by_key(engine, n):
    import torch
    from appleseed_b200 import scenes
    desc = scenes.scene_c2(16)
    ctx = engine.TraceContext(desc, device=0)
    isect = engine.Intersector(ctx)
    lo, hi = scenes.scene_bbox(desc)
    rays = scenes.uniform_sphere_rays(n, lo - 1.0, hi + 1.0, seed=n)
    if n > 40:
        rays.org[7] = np.nan                    # NaNs must not break the sort
        rays.dir[11] = np.inf
    dev = engine.DeviceRays.from_host(rays, "cuda:0")
    order, keys = isect.sort_rays(dev)
    torch.cuda.synchronize()
    order, keys = order.cpu().numpy().view(np.uint32), keys.cpu().numpy().view(np.uint32)
    assert np.array_equal(np.sort(order), np.arange(n, dtype=np.uint32))
    assert np.all(np.diff(keys.astype(np.int64)) >= 0) and keys.max() < (1 << 24)
    same = np.nonzero(np.diff(keys.astype(np.int64)) == 0)[0]
    assert np.all(order[same] < order[same + 1])              # stable: ties keep the input order
    if n > 1000:
        # Rays that are neighbours after the sort start close together.
        o = rays.org[order]
        ok = np.isfinite(o).all(axis=1)
        near = np.linalg.norm(np.diff(o[ok], axis=0), axis=1).mean()
        far = np.linalg.norm(np.diff(rays.org[np.isfinite(rays.org).all(axis=1)], axis=0), axis=1).mean()
        assert near < 0.5 * far


def test_sorted_traces_in_flight_on_several_streams(engine):
    """Two sorted traces on two streams at once: the sort scratch is a stream-ordered allocation of
    each call, never shared through the scene handle."""
    import torch
    desc, rays, _ = cases.case_c3(n=200000)
    ctx = engine.TraceContext(desc, device=0)
    isect = engine.Intersector(ctx)
    a, b = rays.slice(0, 100000), rays.slice(100000, 200000)
    want_a, want_b = isect.trace(a), isect.trace(b)
    da, db = engine.DeviceRays.from_host(a, "cuda:0"), engine.DeviceRays.from_host(b, "cuda:0")
    ha = torch.empty(len(a) * engine.HIT_BYTES, dtype=torch.uint8, device="cuda:0")
    hb = torch.empty(len(b) * engine.HIT_BYTES, dtype=torch.uint8, device="cuda:0")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(4):
        ha.zero_(); hb.zero_()
        torch.cuda.synchronize()
        with torch.cuda.stream(s1):
            isect.trace_device(da, ha, sort=True)
        with torch.cuda.stream(s2):
            isect.trace_device(db, hb, sort=True)
        torch.cuda.synchronize()
        assert engine.hits_from_tensor(ha, len(a)).tobytes() == want_a.tobytes()
        assert engine.hits_from_tensor(hb, len(b)).tobytes() == want_b.tobytes()
