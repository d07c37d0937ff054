"""Fuzzer for the product's host builder + flattener (tree_builder.cpp, flatten.cpp) on hostile scene
DESCRIPTIONS: non-finite / huge / denormal vertices and matrices, degenerate and duplicated
triangles, odd leaf sizes and costs, empty meshes and assemblies.  Every scene must end in an error
message or in a blob that passed validation and traces inside its arrays -- never in a crash or a
hang.  Run as a subprocess by tests/test_fuzz_views.py; prints one summary line.

usage: python fuzz_desc.py <seed> <scenes>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from appleseed_b200 import scenes  # noqa: E402
from appleseed_b200.scene import Assembly, AssemblyInstance, Mesh, ObjectInstance, SceneDesc  # noqa: E402
from hostsim import hostsim  # noqa: E402

ODD = [np.nan, np.inf, -np.inf, 3.0e38, -3.0e38, 1.0e30, 1.0e-45, 0.0, -0.0]


def hostile_scene(rng):
    what = []
    res = int(rng.integers(2, 14))
    base = scenes.grid_mesh(res, "sines")
    v, t = base.vertices.copy(), base.triangles.copy()
    poses = None
    k = int(rng.integers(0, 9))
    if k == 0:                                                  # odd values in some vertices
        n = int(rng.integers(1, 6))
        value = np.float32(ODD[int(rng.integers(0, len(ODD)))])
        v[rng.integers(0, len(v), n), rng.integers(0, 3, n)] = value
        what.append("%d vertex components = %r" % (n, float(value)))
    elif k == 1:                                                # all vertices the same / collinear
        v[:] = v[0] if rng.random() < 0.5 else np.outer(np.linspace(0, 1, len(v)), [1.0, 2.0, 3.0])
        what.append("degenerate mesh")
    elif k == 2:                                                # many copies of one triangle
        t = np.tile(t[:1], (int(rng.integers(3, 200)), 1))
        what.append("%d copies of one triangle" % len(t))
    elif k == 3:                                                # no triangles, or no vertices at all
        t = t[:0]
        if rng.random() < 0.5:
            v = v[:0]
        what.append("empty mesh")
    elif k == 4:                                                # huge coordinates
        v *= np.float32(rng.choice([1e20, 1e30, 1e36, 1e-30, 1e-40]))
        what.append("scaled mesh")
    elif k == 5:                                                # moving mesh with odd poses
        msc = int(rng.integers(1, 4))
        poses = np.repeat(v[:, None, :], msc, axis=1).astype(np.float32)
        poses += rng.normal(size=poses.shape).astype(np.float32) * np.float32(rng.choice([0.0, 0.01, 1e30]))
        if rng.random() < 0.4:
            poses[int(rng.integers(0, len(v))), 0, 0] = np.float32(ODD[int(rng.integers(0, len(ODD)))])
        what.append("moving mesh, msc %d" % msc)
    mesh = Mesh(v, t, vertex_poses=poses) if poses is not None else Mesh(v, t)

    def matrix():
        m = scenes.translation(*rng.uniform(-2, 2, 3)) @ scenes.rotation_y(float(rng.uniform(0, 6.28))) @ scenes.scaling(float(rng.choice([1.0, 0.5, 1e-12, 1e12])))
        if rng.random() < 0.15:
            m[int(rng.integers(0, 3)), int(rng.integers(0, 4))] = ODD[int(rng.integers(0, len(ODD)))]
            what.append("odd matrix entry")
        if rng.random() < 0.05:
            m[:3, :3] = 0.0
            what.append("singular matrix")
        return m

    def safe_inverse(m):
        try:
            inv = np.linalg.inv(m)
            return inv if np.isfinite(inv).all() or rng.random() < 0.5 else np.eye(4)
        except np.linalg.LinAlgError:
            return np.eye(4) if rng.random() < 0.5 else np.full((4, 4), np.nan)

    def instance(cls):
        m = matrix()
        return cls(0, m, parent_to_local=safe_inverse(m))

    ois = [instance(ObjectInstance) for _ in range(int(rng.integers(0, 3)))]
    if not ois:
        what.append("assembly without object instances")
    asm = Assembly(ois, max_leaf_size=int(rng.choice([0, 1, 2, 3, 64, 100000])),
                   interior_node_traversal_cost=float(rng.choice([1.0, 0.0, -1.0, np.inf, np.nan])),
                   triangle_intersection_cost=float(rng.choice([1.0, 0.0, -1.0, np.inf, np.nan])))
    insts = [instance(AssemblyInstance) for _ in range(int(rng.integers(0, 4)))]
    return SceneDesc([mesh], [asm], insts), "; ".join(what) or "plain"


def main():
    seed, count = int(sys.argv[1]), int(sys.argv[2])
    rng = np.random.default_rng(seed)
    sim = hostsim.load()
    rays = scenes.uniform_sphere_rays(200, np.array([-3.0, -3.0, -3.0]), np.array([3.0, 3.0, 3.0]), seed, time=True)
    rejected = accepted = 0
    for k in range(count):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                desc, what = hostile_scene(rng)
            except (np.linalg.LinAlgError, AssertionError, ValueError) as e:        # the Python scene classes refused it
                sys.stdout.write("%d: refused by the harness (%s)\n" % (k, type(e).__name__))
                rejected += 1
                continue
        lbvh = bool(rng.integers(0, 2))
        sys.stdout.write("%d: %s%s\n" % (k, what, " [lbvh]" if lbvh else ""))
        sys.stdout.flush()
        try:
            s = hostsim.SimScene(sim, desc, lbvh=lbvh, threads=2)
        except RuntimeError:
            rejected += 1
            continue
        accepted += 1
        for wide in (False, True):
            s.trace(rays, wide=wide)
            s.trace_probe(rays, wide=wide)
    print("fuzz done: %d scenes, %d rejected, %d accepted" % (count, rejected, accepted))


if __name__ == "__main__":
    main()
