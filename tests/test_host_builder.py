"""The product's host builder (asgpu_trees_build, through the C ABI) must produce the same
reference-format trees as the oracle restatement and, where available, as the reference's own
headers (oracle/_ref).  CPU only."""
import numpy as np
import pytest

import cases
from appleseed_b200 import scenes
from appleseed_b200.intersector import AsgpuError, HostTrees
from test_oracle_vs_ref import compare_tree


def check_against(oracle, desc, threads):
    a = oracle.scene(desc)
    b = HostTrees(desc, threads=threads)
    assert a.tree_count == b.triangle_tree_count
    for i in range(a.tree_count):
        compare_tree(a.triangle_tree(i), b.triangle_tree(i), True)
    ta, tb = a.assembly_tree(), b.assembly_tree()
    compare_tree(ta, tb, False)
    assert np.array_equal(ta["item_assembly_instance"], tb["item_assembly_instance"])
    assert np.array_equal(ta["item_tree"], tb["item_tree"])


@pytest.mark.parametrize("name", list(cases.CASES))
@pytest.mark.parametrize("threads", [1, 4])
def test_trees_match_oracle(orc, name, threads):
    check_against(orc, cases.CASES[name]()[0], threads)


@pytest.mark.parametrize("name", ["cornell", "c3", "c4_msc3", "mixed"])
def test_trees_match_reference_headers(asref, name):
    check_against(asref, cases.CASES[name]()[0], 3)


def test_parallel_subtree_stitching_matches_serial_order(orc):
    # Large enough that the builder defers many subtrees to worker threads and stitches them back
    # into the reference's depth-first node order.
    desc = scenes.scene_c2(260)
    check_against(orc, desc, 8)
    check_against(orc, desc, 1)


@pytest.mark.parametrize("threads", [6, 16, 48])
def test_chunked_sweeps_of_big_ranges_match_the_serial_ones(orc, asref, threads, monkeypatch):
    """Ranges above ASGPU_PARALLEL_SWEEP_MIN items (262 144 by default: the first levels of a
    multi-million-triangle tree) are swept, repartitioned and bounded in chunks by several threads
    per axis, and the two halves of a top node are expanded concurrently.  With the threshold
    lowered the same code runs over several levels of a 135 k-triangle mesh: the tree must stay the
    reference's byte for byte, whatever the thread count (= chunk boundaries)."""
    monkeypatch.setenv("ASGPU_PARALLEL_SWEEP_MIN", "3000")
    desc = scenes.scene_c2(260)
    check_against(orc, desc, threads)
    check_against(asref, scenes.scene_c4(120, msc=2), threads)


def test_chunked_sweeps_at_the_default_threshold(orc):
    check_against(orc, scenes.scene_c2(380), 12)        # 288 800 triangles: the root range is above the default threshold


def test_many_instances_top_level_tree(orc):
    check_against(orc, scenes.scene_c3(8, 12), 2)


def test_rejects_malformed_input():
    from appleseed_b200.scene import Assembly, AssemblyInstance, Mesh, ObjectInstance, SceneDesc
    mesh = Mesh(np.zeros((3, 3), dtype=np.float32), np.array([[0, 1, 7]], dtype=np.uint32))
    with pytest.raises(AsgpuError, match="out of range"):
        HostTrees(SceneDesc([mesh], [Assembly([ObjectInstance(0)])], [AssemblyInstance(0)]))
    ok = Mesh(np.eye(3, dtype=np.float32), np.array([[0, 1, 2]], dtype=np.uint32))
    with pytest.raises(AsgpuError, match="assembly instance"):
        HostTrees(SceneDesc([ok], [Assembly([ObjectInstance(0)])], [AssemblyInstance(3)]))


def test_collection_of_larger_meshes_and_dropped_triangles(orc):
    # Moving triangles (pose-major vertices) and a static mesh with triangles that the collection
    # drops (degenerate ones), which shifts every later triangle's vertex offset, at sizes where the
    # builder works with several threads.  (A chunked, concurrent collection was tried against this
    # test: byte-identical, but no faster than the serial walk -- it is copy bound -- and reverted.)
    check_against(orc, scenes.scene_c4(200, msc=3), 8)
    desc = scenes.scene_c2(200)
    t = desc.meshes[0].triangles
    t[::7, 2] = t[::7, 1]
    check_against(orc, desc, 8)
    check_against(orc, desc, 1)
