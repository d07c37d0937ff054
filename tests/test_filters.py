"""Alpha-mask intersection filters (renderer::IntersectionFilter, cut-out geometry): closest-hit
candidates on transparent texels are skipped, shadow probes ignore filters (triangletree.cpp:
1404-1411, 1455-1462 vs 1506-1603).

CPU tier: restatement == the checker built on the reference's BitMask2 / Vector2f / clamp /
truncate.  GPU tier: exact kernels byte-identical to the oracle, wide kernels under the parity
rule, on a static scene, an instanced scene and a scene with moving triangles and material masks."""
import numpy as np
import pytest

import cases
import parity
from appleseed_b200.scene import IntersectionFilter


def _uv_of(mesh):
    v = mesh.vertices[mesh.triangles]
    lo, hi = mesh.vertices.min(axis=0), mesh.vertices.max(axis=0)
    ext = np.maximum(hi - lo, 1e-6)
    # Slightly outside [0, 1] on purpose: the lookup clamps.
    return np.stack([(v[..., 0] - lo[0]) / ext[0] * 1.1 - 0.05, (v[..., 2] - lo[2]) / ext[2] * 1.1 - 0.05], axis=-1).astype(np.float32)


def _masks(seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:37, 0:53]
    checker = (xx // 3 + yy // 2) % 2 == 0
    stripes = xx % 5 != 0
    noise = rng.random((19, 11)) > 0.35
    return checker, stripes, noise


def filtered_cases():
    checker, stripes, noise = _masks(3)
    out = {}
    desc, rays, probes = cases.case_c2()
    out["c2_object_mask"] = (desc, rays, probes, {(0, 0): IntersectionFilter(_uv_of(desc.meshes[0]), object_mask=checker)})
    desc, rays, probes = cases.case_c3()
    out["c3_instanced"] = (desc, rays, probes, {(0, 0): IntersectionFilter(_uv_of(desc.meshes[0]), object_mask=noise)})
    desc, rays, probes = cases.case_mixed()
    # Assembly 0: object instances 0 and 1 share mesh 0 (pa = index % 5); only instance 1 is filtered,
    # with material masks for pa 1 and 3 and none for the others.  Assembly 1: moving mesh, object mask.
    out["mixed_material_masks"] = (desc, rays, probes, {
        (0, 1): IntersectionFilter(_uv_of(desc.meshes[0]), material_masks=[None, stripes, None, noise]),
        (1, 0): IntersectionFilter(_uv_of(desc.meshes[1]), object_mask=checker, material_masks=[stripes]),
    })
    return out


def _attach(oscene, desc, filters):
    # (tree index, object instance) -> (assembly index, object instance): trees follow the assemblies with geometry.
    with_geometry = [a for a, asm in enumerate(desc.assemblies) if len(asm.object_instances)]
    for (tree, oi), f in filters.items():
        oscene.set_filter(with_geometry[tree], oi, f)


@pytest.mark.parametrize("name", ["c2_object_mask", "c3_instanced", "mixed_material_masks"])
def test_restatement_equals_reference_headers(orc, asref, name):
    desc, rays, probes, filters = filtered_cases()[name]
    o, r = orc.scene(desc), asref.scene(desc)
    plain = o.trace(rays, threads=4)
    pplain = o.trace_probe(probes, threads=4)
    _attach(o, desc, filters); _attach(r, desc, filters)
    a, b = o.trace(rays, threads=4), r.trace(rays, threads=4)
    assert a.tobytes() == b.tobytes()
    changed = int((a["t"] != plain["t"]).sum())
    assert changed > 200                                         # the filter really cuts holes
    assert (a["prim_type"] == 2).sum() < (plain["prim_type"] == 2).sum()
    # Probes ignore filters.
    assert np.array_equal(o.trace_probe(probes, threads=4), pplain)
    assert np.array_equal(r.trace_probe(probes, threads=4), pplain)


@pytest.mark.parametrize("name", ["c2_object_mask", "c3_instanced", "mixed_material_masks"])
def test_host_build_of_product_code_with_filters(orc, name):
    """The product's flattener (filter tables in the blob) and filter_accept, compiled for the host."""
    from hostsim import hostsim
    desc, rays, probes, filters = filtered_cases()[name]
    o = orc.scene(desc)
    pplain = o.trace_probe(probes, threads=4)
    _attach(o, desc, filters)
    ref = o.trace(rays, threads=4)
    sim = hostsim.SimScene(hostsim.load(), desc, filters=filters)
    assert sim.trace(rays, wide=False)[0].tobytes() == ref.tobytes()
    parity.compare_hits(o, rays, sim.trace(rays, wide=True)[0], ref)
    assert np.array_equal(sim.trace_probe(probes, wide=False)[0], pplain)
    parity.compare_probes(o, probes, sim.trace_probe(probes, wide=True)[0], pplain)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2_object_mask", "c3_instanced", "mixed_material_masks"])
def test_kernels_with_filters(orc, name):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from appleseed_b200.intersector import Intersector, TraceContext
    desc, rays, probes, filters = filtered_cases()[name]
    o = orc.scene(desc)
    pplain = o.trace_probe(probes, threads=4)
    _attach(o, desc, filters)
    ref = o.trace(rays, threads=4)
    isect = Intersector(TraceContext(desc, device=0, filters=filters))
    assert isect.trace(rays, exact=True).tobytes() == ref.tobytes()
    parity.compare_hits(o, rays, isect.trace(rays), ref)
    assert np.array_equal(isect.trace_probe(probes, exact=True), pplain)
    parity.compare_probes(o, probes, isect.trace_probe(probes), pplain)
    # Without filters the same scene gives the unfiltered result (the filter data is per scene).
    plain = Intersector(TraceContext(desc, device=0))
    assert plain.trace(rays, exact=True).tobytes() != ref.tobytes()
