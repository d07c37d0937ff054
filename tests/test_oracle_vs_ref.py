"""Differential test: the self-contained restatement (oracle/oracle.cpp) against the reference's
own headers (oracle/_ref).  Trees are compared field by field and hit records byte for byte.
Skipped where oracle/_ref is unavailable."""
import numpy as np
import pytest

import cases

# foundation/math/bvh/bvh_node.h:100-107
NODE = np.dtype([("item_count", "<u4"), ("index", "<u4"), ("lbi", "<u4"), ("lbc", "<u4"), ("rbi", "<u4"),
                 ("rbc", "<u4"), ("pad", "<u4", 2), ("bbox", "<f8", 12)])
KEY_BYTES = [0, 1, 2, 3, 4, 5, 8, 9, 10, 11]      # TriangleKey minus its 2-byte hole


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def compare_tree(a, b, triangle_tree):
    na, nb = a["nodes"].view(NODE), b["nodes"].view(NODE)
    assert na.shape == nb.shape
    assert np.array_equal(na["item_count"], nb["item_count"]) and np.array_equal(na["index"], nb["index"])
    interior = na["item_count"] == 0xFFFFFFFF
    assert same_bits(na["bbox"][interior], nb["bbox"][interior])
    if not triangle_tree:
        return
    assert same_bits(na["bbox"][~interior], nb["bbox"][~interior])          # in-node leaf payloads
    assert np.array_equal(na["lbc"][interior], nb["lbc"][interior])
    assert np.array_equal(na["rbc"][interior], nb["rbc"][interior])
    m = interior & (na["lbc"] > 1)
    assert np.array_equal(na["lbi"][m], nb["lbi"][m])
    m = interior & (na["rbc"] > 1)
    assert np.array_equal(na["rbi"][m], nb["rbi"][m])
    assert same_bits(a["node_bboxes"], b["node_bboxes"])
    assert same_bits(a["leaf_data"], b["leaf_data"])
    assert same_bits(a["triangle_keys"].reshape(-1, 12)[:, KEY_BYTES], b["triangle_keys"].reshape(-1, 12)[:, KEY_BYTES])
    assert a["static_triangle_count"] == b["static_triangle_count"]
    assert a["moving_triangle_count"] == b["moving_triangle_count"]


@pytest.mark.parametrize("name", list(cases.CASES))
def test_trees_and_hits_identical(orc, asref, name):
    desc, rays, probes = cases.CASES[name]()
    a, b = asref.scene(desc), orc.scene(desc)
    assert a.tree_count == b.tree_count
    for i in range(a.tree_count):
        compare_tree(a.triangle_tree(i), b.triangle_tree(i), True)
    ta, tb = a.assembly_tree(), b.assembly_tree()
    compare_tree(ta, tb, False)
    assert np.array_equal(ta["item_assembly_instance"], tb["item_assembly_instance"])
    assert np.array_equal(ta["item_tree"], tb["item_tree"])
    assert a.trace(rays, threads=4).tobytes() == b.trace(rays, threads=4).tobytes()
    assert a.trace_probe(probes, threads=4).tobytes() == b.trace_probe(probes, threads=4).tobytes()


def test_larger_static_tree_identical(orc, asref):
    from appleseed_b200 import scenes
    desc = scenes.scene_c2(220)
    a, b = asref.scene(desc), orc.scene(desc)
    compare_tree(a.triangle_tree(0), b.triangle_tree(0), True)
