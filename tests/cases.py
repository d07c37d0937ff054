"""Small seeded parity cases (scene + closest-hit rays + probe rays) shared by the oracle tests,
the golden-vector generator and the GPU parity tests.  Sizes are chosen so the CPU oracle
finishes each case in well under a second."""
import numpy as np

from appleseed_b200 import scenes
from appleseed_b200.scene import (VIS_ALL, VIS_CAMERA, VIS_DIFFUSE, VIS_SHADOW, Assembly, AssemblyInstance,
                                  Mesh, ObjectInstance, RayBatch, SceneDesc)


def _concat(a: RayBatch, b: RayBatch) -> RayBatch:
    def cat(x, y, fill, dt):
        if x is None and y is None:
            return None
        x = np.full(len(a), fill, dtype=dt) if x is None else x
        y = np.full(len(b), fill, dtype=dt) if y is None else y
        return np.concatenate([x, y])
    return RayBatch(np.concatenate([a.org, b.org]), np.concatenate([a.dir, b.dir]),
                    np.concatenate([a.tmin, b.tmin]), np.concatenate([a.tmax, b.tmax]),
                    cat(a.time_absolute, b.time_absolute, 0.0, np.float32),
                    cat(a.time_normalized, b.time_normalized, 0.0, np.float32),
                    cat(a.flags, b.flags, VIS_ALL, np.uint32))


def _shadow_from(desc, rays, seed, n_lights=4):
    lo, hi = scenes.scene_bbox(desc)
    rng = np.random.default_rng(seed)
    lights = lo + rng.random((n_lights, 3)) * (hi - lo)
    lights[:, 1] = hi[1] + 0.5
    pts = rays.org + 0.3 * rays.dir
    return scenes.shadow_rays(pts, lights, seed + 1)


def case_cornell():
    desc = scenes.scene_c1()
    prim = scenes.rays_c1_primary()
    sub = prim.take(np.arange(0, len(prim), 7))
    # Shadow probes between random points inside the box and points just under the ceiling.
    lo, hi = scenes.scene_bbox(desc)
    rng = np.random.default_rng(11)
    pts = lo + rng.random((20000, 3)) * (hi - lo)
    lights = lo + rng.random((4, 3)) * (hi - lo)
    lights[:, 1] = lo[1] + 0.97 * (hi[1] - lo[1])
    return desc, sub, scenes.shadow_rays(pts, lights, 12)


def case_c2(res=96, n=20000):
    desc = scenes.scene_c2(res)
    cam = scenes.pinhole_rays(96, 96, (0.1, 3.0, 0.05), (0.0, 0.0, 0.0), up=(0, 0, 1), focal=0.05)
    lo, hi = scenes.scene_bbox(desc)
    inc = scenes.uniform_sphere_rays(n, lo - 0.1, hi + np.array([0.1, 0.6, 0.1]), 5)
    rays = _concat(cam, inc)
    return desc, rays, _shadow_from(desc, inc, 12)


def case_c3(res=64, lattice=3, n=30000):
    desc = scenes.scene_c3(res, lattice)
    lo, hi = scenes.scene_bbox(desc)
    rays = scenes.uniform_sphere_rays(n, lo - 0.1, hi + np.array([0.1, 1.0, 0.1]), 6)
    return desc, rays, _shadow_from(desc, rays, 13)


def case_c4(msc=1, res=48, n=30000):
    desc = scenes.scene_c4(res, msc)
    lo, hi = scenes.scene_bbox(desc)
    rays = scenes.uniform_sphere_rays(n, lo - 0.1, hi + np.array([0.1, 0.5, 0.1]), 7, time=True)
    sh = _shadow_from(desc, rays, 14)
    sh.time_absolute = rays.time_absolute
    sh.time_normalized = rays.time_normalized
    return desc, rays, sh


def case_mixed(n=20000):
    """Two assemblies (one static with two object instances of different visibility, one moving
    with msc = 3), five assembly instances with mixed visibility, degenerate triangles, a mesh with
    no triangles, and ray flags drawn from several ray types."""
    g = scenes.grid_mesh(24, "sines")
    tris = g.triangles.copy()
    tris[5] = [3, 3, 7]                    # zero area: dropped at build time (triangletree.cpp:141)
    tris[40] = [9, 10, 9]
    static_mesh = Mesh(g.vertices, tris, triangle_pa=(np.arange(tris.shape[0]) % 5).astype(np.uint16))
    moving = scenes.moving_grid_mesh(16, 3, seed=9)
    empty = Mesh(np.array([[-1, -1, -1], [1, 1, 1]], dtype=np.float32), np.zeros((0, 3), dtype=np.uint32))
    a0 = Assembly([
        ObjectInstance(0, scenes.translation(0, 0, 0), VIS_ALL),
        ObjectInstance(0, scenes.translation(0.0, 0.4, 0.0) @ scenes.rotation_y(0.3) @ scenes.scaling(0.5), VIS_CAMERA | VIS_DIFFUSE),
        ObjectInstance(2),
    ])
    a1 = Assembly([ObjectInstance(1, scenes.scaling(0.8))])
    a2 = Assembly([])                       # skipped: no object instances (assemblytree.cpp:133-134)
    insts = [
        AssemblyInstance(0, scenes.translation(-1.5, 0, 0)),
        AssemblyInstance(1, scenes.translation(1.5, 0.1, 0.2) @ scenes.rotation_y(0.7)),
        AssemblyInstance(2, scenes.translation(0, 2, 0)),
        AssemblyInstance(0, scenes.translation(0.3, 0.8, 0.1) @ scenes.scaling(1.3, 0.7, 1.1), VIS_SHADOW | VIS_DIFFUSE),
        AssemblyInstance(1, scenes.translation(0, -0.9, 0) @ scenes.scaling(2.0), VIS_CAMERA),
    ]
    desc = SceneDesc([static_mesh, moving, empty], [a0, a1, a2], insts)
    lo, hi = scenes.scene_bbox(desc)
    rays = scenes.uniform_sphere_rays(n, lo - 0.2, hi + 0.2, 8, time=True)
    rng = np.random.default_rng(21)
    rays.flags = rng.choice(np.array([VIS_CAMERA, VIS_SHADOW, VIS_DIFFUSE, 1 << 6, VIS_ALL], dtype=np.uint32), size=n)
    rays.tmin = rng.choice(np.array([0.0, 0.0, 0.25]), size=n)
    rays.tmax = rng.choice(np.array([scenes.DBL_MAX, 1.5, 4.0]), size=n)
    probes = rays.take(np.arange(n))
    return desc, rays, probes


CASES = {
    "cornell": case_cornell,
    "c2": case_c2,
    "c3": case_c3,
    "c4_msc1": lambda: case_c4(1),
    "c4_msc2": lambda: case_c4(2),
    "c4_msc3": lambda: case_c4(3),
    "mixed": case_mixed,
}
