"""The C-ABI shared library loads on a machine without a GPU, exports every symbol that
include/asgpu.h declares, and fails loudly (no CPU fallback) when asked to compute without CUDA."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from appleseed_b200 import _lib, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "asgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(asgpu_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTS) == names
    assert lib.asgpu_version() == 2


def test_struct_sizes_match_header():
    from appleseed_b200.scene import HIT_DTYPE, CAssemblyInstance, CMesh, CObjectInstance, CRays
    assert HIT_DTYPE.itemsize == 40
    assert C.sizeof(CMesh) == 48 and C.sizeof(CObjectInstance) == 264 and C.sizeof(CAssemblyInstance) == 264
    assert C.sizeof(CRays) == 56
    assert C.sizeof(_lib.AssemblyItem) == 144 and C.sizeof(_lib.TriangleTreeView) == 80
    assert C.sizeof(_lib.PathStreamDesc) == 368 and C.sizeof(_lib.PathStreamStats) == 64      # gcc sizeof of the header's structs
    assert C.sizeof(_lib.PathStreamProfile) == 48
    import fuzz_views
    assert C.sizeof(fuzz_views.SourceObject) == 168


def test_null_arguments_are_rejected_with_a_message():
    lib = _lib.load()
    assert not lib.asgpu_trees_build(None, 1)
    assert "null" in _lib.last_error()
    assert lib.asgpu_trace(None, None, 10, None, 0, None) < 0
    assert "scene" in _lib.last_error()


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from appleseed_b200.intersector import AsgpuError, TraceContext
    with pytest.raises(AsgpuError, match="(?i)cuda|device|driver"):
        TraceContext(scenes.scene_c2(8))


def test_product_does_not_reference_the_oracle():
    # The product tree must not import, include, link or call anything under oracle/.
    pkg = os.path.join(ROOT, "appleseed_b200")
    pattern = re.compile(r"import\s+oracle|from\s+oracle|#include\s*[\"<][^\">]*oracle|liboracle|libasref|\borc_\w+\s*\(|\basref_\w+\s*\(")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cpp", ".cu", "Makefile")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert not pattern.search(text), os.path.join(base, f)
    out = os.popen("ldd %s" % _lib.LIB_PATH).read()
    assert "oracle" not in out and "asref" not in out
    syms = os.popen("nm -D %s" % _lib.LIB_PATH).read()
    assert "orc_" not in syms and "asref_" not in syms and "hostsim" not in syms


def test_header_is_plain_c(tmp_path):
    """include/asgpu.h is the whole boundary: it must compile as C99 and as C++ on its own."""
    import subprocess
    src = tmp_path / "use.c"
    src.write_text('#include "asgpu.h"\nint main(void) { asgpu_hit h; asgpu_parent p; (void) h; (void) p; return asgpu_version() == ASGPU_VERSION ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src)], check=True)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)], check=True)
    # ... and link against the library from C.
    exe = tmp_path / "use"
    subprocess.run(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), _lib.LIB_PATH, "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0
