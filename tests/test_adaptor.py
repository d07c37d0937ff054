"""The product's C++ adaptor (include/asgpu_adaptor.hpp) compiled against the reference's own headers
(tests/adaptor/adaptor_check.cpp): the flattener views it makes from reference-type trees, and the
ShadingPoints it makes from asgpu_hit records + support planes, against what the reference traversal
leaves in its ShadingPoint (primary block, shading/shadingpoint.h:289-302; written at
assemblytree.cpp:733-744 and by Intersector::make_triangle_shading_point, intersector.cpp:240-271).

CPU tier: the hit records / planes come from the host build of the product's kernels (tests/hostsim)
over the adaptor's views.  GPU tier: from asgpu_trace / asgpu_get_support_planes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cases
from appleseed_b200 import _lib
from appleseed_b200.scene import CRays

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "adaptor", "libadaptorcheck.so")


@pytest.fixture(scope="module")
def adaptor():
    """(library, Oracle bound to the library's own build of ref_driver.cpp)."""
    from oracle.oracle import Oracle
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "adaptor")], check=False)
    if not os.path.exists(LIB):
        pytest.skip("tests/adaptor/libadaptorcheck.so not available (built only where /root/reference exists)")
    lib = C.CDLL(LIB)
    lib.adaptor_views_create.restype = C.c_void_p
    lib.adaptor_views_create.argtypes = [C.c_void_p]
    lib.adaptor_views_destroy.argtypes = [C.c_void_p]
    lib.adaptor_views_trees.restype = C.POINTER(_lib.TriangleTreeView)
    lib.adaptor_views_trees.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    lib.adaptor_views_top.restype = C.POINTER(_lib.AssemblyTreeView)
    lib.adaptor_views_top.argtypes = [C.c_void_p]
    lib.adaptor_check_shading_points.restype = C.c_longlong
    lib.adaptor_check_shading_points.argtypes = [C.c_void_p, C.POINTER(CRays), C.c_size_t, C.c_void_p, C.c_void_p, C.c_int,
                                                 C.POINTER(C.c_longlong), C.c_char_p, C.c_size_t]
    return lib, Oracle("asref", path=LIB)


class Views:
    """The adaptor's views of a reference-side scene (valid while this object and the scene live)."""

    def __init__(self, lib, oscene):
        self.lib, self.oscene = lib, oscene
        self.handle = lib.adaptor_views_create(oscene.handle)
        n = C.c_uint32(0)
        p = lib.adaptor_views_trees(self.handle, C.byref(n))
        self.trees = [p[i] for i in range(n.value)]
        self.top = lib.adaptor_views_top(self.handle).contents

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.adaptor_views_destroy(self.handle)
            self.handle = None


def check_points(lib, oscene, rays, hits, planes, exact_identity=True):
    cr = rays.to_c()
    hits = np.ascontiguousarray(hits)
    planes = np.ascontiguousarray(planes, dtype=np.float64)
    ties = C.c_longlong(0)
    msg = C.create_string_buffer(512)
    bad = lib.adaptor_check_shading_points(oscene.handle, C.byref(cr), len(rays), hits.ctypes.data, planes.ctypes.data,
                                           1 if exact_identity else 0, C.byref(ties), msg, 512)
    return int(bad), int(ties.value), msg.value.decode()


def animated_scene(oracle):
    from test_animated_instances import animated_case
    desc, rays, probes, keys = animated_case()
    return desc, rays, oracle.scene(desc, keys=keys)


SCENES = ["cornell", "c3", "c4_msc3", "mixed", "animated"]


def make_scene(oracle, name):
    if name == "animated":
        return animated_scene(oracle)
    desc, rays, _ = cases.CASES[name]()
    return desc, rays, oracle.scene(desc)


@pytest.mark.parametrize("name", SCENES)
def test_adaptor_on_the_host_build_of_the_kernels(adaptor, name):
    from hostsim import hostsim
    lib, oracle = adaptor
    desc, rays, oscene = make_scene(oracle, name)
    v = Views(lib, oscene)
    sim = hostsim.SimScene.from_views(hostsim.load(), v.trees, v.top, v)
    ref = oscene.trace(rays, threads=4)
    hits = sim.trace(rays, wide=False)[0]
    assert hits.tobytes() == ref.tobytes()                  # the adaptor's views describe the reference's trees
    planes = sim.support_planes(rays, hits)
    bad, ties, msg = check_points(lib, oscene, rays, hits, planes)
    assert bad == 0 and ties == 0, msg
    assert int((hits["prim_type"] == 2).sum()) > 100
    # The check does notice a wrong record: shift one hit's distance by an ulp, swap a plane.
    wrong = hits.copy()
    k = int(np.nonzero(hits["prim_type"] == 2)[0][0])
    wrong["t"][k] = np.nextafter(wrong["t"][k], np.inf)
    assert check_points(lib, oscene, rays, wrong, planes)[0] == 1
    wrong_planes = planes.copy()
    wrong_planes[k, 3:6], wrong_planes[k, 6:9] = planes[k, 6:9], planes[k, 3:6]
    assert check_points(lib, oscene, rays, hits, wrong_planes)[0] == 1
    # The throughput layout through the same views: only exact-t ties may pick another triangle.
    wide = sim.trace(rays, wide=True)[0]
    bad, ties, msg = check_points(lib, oscene, rays, wide, sim.support_planes(rays, wide), exact_identity=False)
    assert bad == 0 and ties <= 5, msg


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCENES)
def test_adaptor_on_the_gpu(adaptor, name):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from appleseed_b200.intersector import Intersector, TraceContext
    lib, oracle = adaptor
    desc, rays, oscene = make_scene(oracle, name)
    v = Views(lib, oscene)
    isect = Intersector(TraceContext.from_tree_views(v.trees, v.top))
    for exact in (True, False):
        hits = isect.trace(rays, exact=exact)
        planes = isect.support_planes(rays, hits)
        bad, ties, msg = check_points(lib, oscene, rays, hits, planes, exact_identity=exact)
        assert bad == 0 and ties <= 5, msg
