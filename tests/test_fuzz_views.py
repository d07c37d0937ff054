"""The product's flattener must survive corrupted reference-format trees (an in-tree integration
passes it live pointers): every mutation ends in an error message or in a blob that passed the
product's own validation and traces without leaving its arrays.  The fuzzer (tests/fuzz_views.py)
runs in a subprocess so that a crash or a hang fails this test instead of the test runner.  For a
stronger check run it against an AddressSanitizer build of tests/hostsim (HOSTSIM_LIB=..., with
LD_PRELOAD=libasan): 900 mutations were clean that way when this test was written.  It found one
bug on its first day: scenes with the exact layout only uploaded the assembly tree's nodes
unchecked (flatten.cpp: check_hierarchy now runs on every tree whatever the layout flags)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("seed", [21, 22])
def test_corrupted_trees_are_rejected_or_harmless(seed):
    r = subprocess.run([sys.executable, os.path.join(HERE, "fuzz_views.py"), str(seed), "150"], capture_output=True, text=True, timeout=600)
    tail = "\n".join(r.stdout.splitlines()[-3:])
    assert r.returncode == 0, "fuzzer died (rc %d) after: %s\n%s" % (r.returncode, tail, r.stderr[-2000:])
    last = r.stdout.splitlines()[-1]
    assert last.startswith("fuzz done: 150 mutations"), last
    rejected = int(last.split(",")[1].split()[0])
    assert 50 <= rejected <= 150, last              # most corruptions are caught; a few are harmless (e.g. a box entry)


@pytest.mark.gpu
def test_corrupted_assembly_tree_is_rejected_through_the_c_abi():
    """The case the fuzzer found, through asgpu_scene_create: an exact-layout-only scene whose
    assembly tree points back at its own root must not reach the kernels."""
    import numpy as np
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    sys.path.insert(0, HERE)
    import cases
    import fuzz_views
    from appleseed_b200 import _lib
    from appleseed_b200.intersector import AsgpuError, Intersector, TraceContext
    desc, rays, _ = cases.case_c3(res=12, lattice=2, n=300)
    tt, top = fuzz_views.snapshot(desc)
    views, a = fuzz_views.views_of(tt, top)
    good = Intersector(TraceContext.from_tree_views(views, a, flags=_lib.SCENE_EXACT)).trace(rays, exact=True)
    assert (good["prim_type"] == 2).sum() > 10
    nodes = top["nodes"].view(np.uint32).reshape(-1, fuzz_views.NODE_U32)
    inner = np.nonzero(nodes[:, 0] == 0xFFFFFFFF)[0]
    nodes[inner[-1], 1] = 0                                     # a cycle through the root
    with pytest.raises(AsgpuError, match="parent-before-child"):
        TraceContext.from_tree_views(views, a, flags=_lib.SCENE_EXACT)
    nodes[inner[-1], 1] = 0x7FFFFFF0                            # far outside the array
    with pytest.raises(AsgpuError, match="parent-before-child|out of range"):
        TraceContext.from_tree_views(views, a, flags=_lib.SCENE_EXACT)


@pytest.mark.parametrize("seed", [23])
def test_hostile_scene_descriptions_are_rejected_or_harmless(seed):
    """tests/fuzz_desc.py: non-finite / huge / denormal vertices and matrices, degenerate and
    duplicated triangles, odd leaf sizes and costs, both tree builders.  Found one bug: coordinates
    near FLT_MAX overflowed every SAH cost and the sweep recursed for ever on an empty left half
    (tree_builder.cpp: split() now makes a leaf when no candidate exists)."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "fuzz_desc.py"), str(seed), "50"], capture_output=True, text=True, timeout=900)
    tail = "\n".join(r.stdout.splitlines()[-3:])
    assert r.returncode == 0, "fuzzer died (rc %d) after: %s\n%s" % (r.returncode, tail, r.stderr[-2000:])
    last = r.stdout.splitlines()[-1]
    assert last.startswith("fuzz done: 50 scenes"), last
    assert int(last.split(",")[2].split()[0]) >= 15, last        # and most scenes are still accepted and traced
