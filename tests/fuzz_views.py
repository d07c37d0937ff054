"""Fuzzer for the product's flattener (appleseed_b200/csrc/flatten.cpp) on CORRUPTED reference-format
trees: an in-tree integration hands it live pointers into the renderer's arrays, so malformed input
must end in an error message (or in a valid blob that traces without leaving its arrays) -- never in
a crash, a hang or an out-of-bounds access.  Run as a subprocess by tests/test_fuzz_views.py (a crash
would take the test runner down); prints one summary line.

usage: python fuzz_views.py <seed> <mutations>"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from appleseed_b200 import _lib  # noqa: E402
from appleseed_b200.intersector import HostTrees  # noqa: E402
from hostsim import hostsim  # noqa: E402

class SourceObject(C.Structure):       # asgpu_source_object
    _fields_ = [("vertices", C.c_void_p), ("triangles", C.c_void_p), ("vertex_count", C.c_uint32), ("triangle_count", C.c_uint32),
                ("triangle_stride", C.c_uint32), ("motion_segment_count", C.c_uint32), ("parent_to_local", C.c_double * 16),
                ("vertex_poses", C.c_void_p)]


NODE_U32 = 32       # 128-byte node = 32 words: item_count, index, 4 motion box words, 2 pad, 24 words of boxes / user data


def snapshot(desc):
    """Private, mutable copies of every array of the product's own host trees."""
    trees = HostTrees(desc)
    tt = [trees.triangle_tree(i) for i in range(trees.triangle_tree_count)]
    v = trees.assembly_tree_view()
    items = (_lib.AssemblyItem * max(1, int(v.item_count)))()
    for i in range(int(v.item_count)):
        C.memmove(C.byref(items[i]), C.byref(v.items[i]), C.sizeof(_lib.AssemblyItem))
    top = {"nodes": HostTrees._bytes(v.nodes, v.node_count * 128), "items": items, "item_count": int(v.item_count)}
    # Source geometry: private copies of the object records; the vertex / index arrays they point at
    # stay owned by `trees`, which therefore has to outlive the snapshot.
    top["sources"] = []
    for i in range(trees.triangle_tree_count):
        g = trees.source_geometry(i)
        objs = (SourceObject * max(1, g.object_count))()
        if g.object_count:
            C.memmove(objs, g.objects, C.sizeof(SourceObject) * g.object_count)
        top["sources"].append((objs, int(g.object_count)))
    top["keep"] = trees
    return tt, top


def views_of(tt, top):
    views = []
    for t in tt:
        w = _lib.TriangleTreeView()
        w.nodes = t["nodes"].ctypes.data
        w.node_bboxes = t["node_bboxes"].ctypes.data if len(t["node_bboxes"]) else None
        w.leaf_data = t["leaf_data"].ctypes.data if len(t["leaf_data"]) else None
        w.triangle_keys = t["triangle_keys"].ctypes.data if len(t["triangle_keys"]) else None
        w.node_count = len(t["nodes"]) // 128
        w.node_bbox_count = len(t["node_bboxes"]) // 6
        w.leaf_data_size = len(t["leaf_data"])
        w.triangle_key_count = len(t["triangle_keys"]) // 12
        w.static_triangle_count = t["static_triangle_count"]
        w.moving_triangle_count = t["moving_triangle_count"]
        views.append(w)
    a = _lib.AssemblyTreeView()
    a.nodes = top["nodes"].ctypes.data
    a.items = C.cast(top["items"], C.POINTER(_lib.AssemblyItem))
    a.node_count = len(top["nodes"]) // 128
    a.item_count = top["item_count"]
    a.item_motion = None
    return views, a


def sources_of(top):
    out = []
    for objs, count in top["sources"]:
        g = _lib.SourceGeometry()
        g.objects = C.cast(objs, C.c_void_p)
        g.object_count = count
        g.reserved = 0
        g.filters = None
        out.append(g)
    return out


def mutate(rng, tt, top):
    """One random corruption; returns a short description."""
    interesting = [0, 1, 2, 3, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFE, 0xFFFFFFFF]
    kind = int(rng.integers(0, 12))
    t = tt[int(rng.integers(0, len(tt)))]
    if kind >= 10 and top.get("sources"):           # source geometry: counts that no longer cover the keys, odd matrices
        k = int(rng.integers(0, len(top["sources"])))
        objs, count = top["sources"][k]
        if count and rng.random() < 0.7:
            o = objs[int(rng.integers(0, count))]
            which = int(rng.integers(0, 3))
            if which == 0:
                o.vertex_count = int(rng.integers(0, max(1, o.vertex_count)))
                return "source vertex_count = %d" % o.vertex_count
            if which == 1:
                o.triangle_count = int(rng.integers(0, max(1, o.triangle_count)))
                return "source triangle_count = %d" % o.triangle_count
            o.parent_to_local[int(rng.integers(0, 16))] = float(rng.choice([np.nan, np.inf, 0.0, 1e300]))
            return "source matrix entry odd"
        top["sources"][k] = (objs, int(rng.integers(0, count + 1)))
        return "source object_count = %d" % top["sources"][k][1]
    nodes = t["nodes"].view(np.uint32).reshape(-1, NODE_U32)
    n = nodes.shape[0]
    value = int(rng.choice(interesting)) if rng.random() < 0.5 else int(rng.integers(0, max(2, 2 * n)))
    if kind <= 2:                                   # a header word of a triangle-tree node
        i, w = int(rng.integers(0, n)), int(rng.integers(0, 6))
        nodes[i, w] = value
        return "tree node %d word %d = %#x" % (i, w, value)
    if kind == 3:                                   # leaf user data: payload offset / vis / msc words
        leaves = np.nonzero(nodes[:, 0] != 0xFFFFFFFF)[0]
        i, w = int(rng.choice(leaves)), 8 + int(rng.integers(0, 4))
        nodes[i, w] = value
        return "leaf %d user word %d = %#x" % (i, w - 8, value)
    if kind == 4:                                   # a child pointer that closes a cycle
        inner = np.nonzero(nodes[:, 0] == 0xFFFFFFFF)[0]
        if len(inner):
            i = int(rng.choice(inner))
            nodes[i, 1] = int(rng.integers(0, i + 1))
            return "tree node %d points back at %d" % (i, nodes[i, 1])
    if kind == 5 and len(t["leaf_data"]) >= 8:      # spilled payload header
        words = t["leaf_data"][: len(t["leaf_data"]) // 4 * 4].view(np.uint32)
        i = int(rng.integers(0, len(words)))
        words[i] = value
        return "leaf_data word %d = %#x" % (i, value)
    if kind == 6:                                   # declared sizes that disagree with the arrays
        which = rng.choice(["static_triangle_count", "moving_triangle_count"])
        t[which] = int(rng.choice([0, 1, 2 ** 31, 2 ** 40]))
        return "%s = %d" % (which, t[which])
    if kind == 7:                                   # a non-finite or inverted child box
        inner = np.nonzero(nodes[:, 0] == 0xFFFFFFFF)[0]
        if len(inner):
            i = int(rng.choice(inner))
            boxes = t["nodes"].view(np.float64).reshape(-1, 16)
            boxes[i, 4 + int(rng.integers(0, 12))] = float(rng.choice([np.nan, np.inf, -np.inf, 1e300, -1e300]))
            return "tree node %d box entry non-finite / huge" % i
    if kind == 8 and top["item_count"]:             # assembly items
        it = top["items"][int(rng.integers(0, top["item_count"]))]
        if rng.random() < 0.5:
            it.triangle_tree = value
            return "item triangle_tree = %#x" % value
        it.parent_to_local[int(rng.integers(0, 16))] = float(rng.choice([np.nan, np.inf, 0.0]))
        return "item matrix entry non-finite / zero"
    top_nodes = top["nodes"].view(np.uint32).reshape(-1, NODE_U32)          # a header word of an assembly-tree node
    i, w = int(rng.integers(0, top_nodes.shape[0])), int(rng.integers(0, 2))
    top_nodes[i, w] = value
    return "assembly node %d word %d = %#x" % (i, w, value)


def main():
    seed, count = int(sys.argv[1]), int(sys.argv[2])
    rng = np.random.default_rng(seed)
    sim = hostsim.load()
    sources = [cases.case_mixed(n=300), cases.case_c3(res=12, lattice=2, n=300), cases.case_c4(msc=3, res=10, n=300)]
    pristine = [(snapshot(desc), rays) for desc, rays, _ in sources]
    rejected = accepted = 0
    for k in range(count):
        (tt0, top0), rays = pristine[k % len(pristine)]
        tt = [dict(t, nodes=t["nodes"].copy(), node_bboxes=t["node_bboxes"].copy(), leaf_data=t["leaf_data"].copy(),
                   triangle_keys=t["triangle_keys"].copy()) for t in tt0]
        items = (_lib.AssemblyItem * max(1, top0["item_count"]))()
        C.memmove(items, top0["items"], C.sizeof(items))
        top = {"nodes": top0["nodes"].copy(), "items": items, "item_count": top0["item_count"], "sources": []}
        for objs0, count0 in top0["sources"]:
            objs = (SourceObject * max(1, count0))()
            C.memmove(objs, objs0, C.sizeof(objs))
            top["sources"].append((objs, count0))
        what = [mutate(rng, tt, top) for _ in range(int(rng.integers(1, 4)))]
        views, a = views_of(tt, top)
        sys.stdout.write("%d: %s\n" % (k, "; ".join(what)))
        sys.stdout.flush()
        try:
            flags = int(rng.choice([hostsim.SCENE_EXACT, hostsim.SCENE_WIDE, hostsim.SCENE_EXACT | hostsim.SCENE_WIDE]))
            with_sources = bool(rng.integers(0, 2))
            s = hostsim.SimScene.from_views(sim, views, a, [tt, top], flags=flags, sources=sources_of(top) if with_sources else None)
        except RuntimeError:
            rejected += 1
            continue
        accepted += 1
        # Accepted: the blob passed the product's own validation, so the traversals must stay inside it
        # and terminate.
        if flags & hostsim.SCENE_EXACT:
            hits = s.trace(rays, wide=False)[0]
            s.trace_probe(rays, wide=False)
            # refine_and_offset: static triangles only, and only when every tree has source geometry
            # (asgpu_refine_and_offset refuses anything else before the kernel is launched: api.cu).
            if with_sources and k % len(pristine) != 2 and all(count > 0 for _, count in top["sources"]):
                s.refine_offset(rays, hits)
        if flags & hostsim.SCENE_WIDE:
            s.trace(rays, wide=True)
            s.trace_probe(rays, wide=True)
    print("fuzz done: %d mutations, %d rejected, %d accepted" % (count, rejected, accepted))


if __name__ == "__main__":
    main()
