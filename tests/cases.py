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


# ---------------------------------------------------------------------------------------------
# Random scenes: triangle soups, arbitrary affine object / assembly instances (rotations about
# any axis, non-uniform and negative scales = handedness swaps), random leaf sizes and SAH costs,
# random visibility masks, overlapping and coincident geometry (exact-t ties), rays with random
# intervals and flags.  Used by the differential tests of both tiers.
# ---------------------------------------------------------------------------------------------

def _random_affine(rng, spread, allow_flip=True):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    r = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    s = rng.uniform(0.4, 1.8, size=3)
    if allow_flip and rng.random() < 0.3:
        s[rng.integers(0, 3)] *= -1.0
    m = np.eye(4)
    m[:3, :3] = r @ np.diag(s)
    m[:3, 3] = rng.uniform(-spread, spread, size=3)
    return m


def _random_mesh(rng, moving):
    kind = rng.integers(0, 3)
    if kind == 0:           # soup of small random triangles
        nt = int(rng.integers(20, 400))
        centres = rng.uniform(-1, 1, size=(nt, 1, 3))
        v = (centres + rng.normal(scale=rng.uniform(0.02, 0.3), size=(nt, 3, 3))).reshape(-1, 3)
        t = np.arange(nt * 3, dtype=np.uint32).reshape(nt, 3)
    elif kind == 1:         # displaced grid, some triangles duplicated (coincident geometry)
        g = scenes.grid_mesh(int(rng.integers(3, 14)), "sines")
        v, t = g.vertices.copy(), g.triangles.copy()
        dup = rng.integers(0, len(t), size=max(1, len(t) // 10))
        t = np.concatenate([t, t[dup]])
    else:                   # a few large triangles spanning everything (big leaves, deep overlap)
        nt = int(rng.integers(1, 12))
        v = rng.uniform(-1.5, 1.5, size=(nt * 3, 3))
        t = np.arange(nt * 3, dtype=np.uint32).reshape(nt, 3)
    v = v.astype(np.float32)
    poses = None
    if moving:
        msc = int(rng.choice([1, 2, 3]))
        drift = rng.normal(scale=0.08, size=(1, msc, 3)).cumsum(axis=1) + rng.normal(scale=0.02, size=(len(v), msc, 3))
        poses = (v[:, None, :] + drift).astype(np.float32)
    return Mesh(v, t, triangle_pa=(np.arange(len(t)) % 7).astype(np.uint16), vertex_poses=poses)


def random_scene(seed, n_rays=6000, moving=True):
    rng = np.random.default_rng(1000 + seed)
    vis_choices = np.array([VIS_ALL, VIS_ALL, VIS_CAMERA | VIS_DIFFUSE, VIS_SHADOW, VIS_ALL & ~VIS_SHADOW], dtype=np.uint64)
    n_static = int(rng.integers(1, 4))
    meshes = [_random_mesh(rng, False) for _ in range(n_static)]
    if moving and rng.random() < 0.6:
        meshes.append(_random_mesh(rng, True))
    assemblies = []
    for _ in range(int(rng.integers(1, 4))):
        # One triangle tree holds either static or moving meshes with a single msc (the reference's
        # Debug build asserts a uniform pose count per tree, triangletree.cpp:824).
        want_moving = len(meshes) > n_static and rng.random() < 0.4
        pool = [len(meshes) - 1] if want_moving else list(range(n_static))
        ois = [ObjectInstance(int(rng.choice(pool)), _random_affine(rng, 0.8), int(rng.choice(vis_choices)))
               for _ in range(int(rng.integers(1, 4)))]
        assemblies.append(Assembly(ois, max_leaf_size=int(rng.choice([1, 2, 2, 4, 8])),
                                   interior_node_traversal_cost=float(rng.choice([1.0, 1.0, 2.5])),
                                   triangle_intersection_cost=float(rng.choice([1.0, 1.0, 0.5]))))
    insts = [AssemblyInstance(int(rng.integers(0, len(assemblies))), _random_affine(rng, 2.5), int(rng.choice(vis_choices)))
             for _ in range(int(rng.integers(1, 7)))]
    if rng.random() < 0.3:
        insts.append(AssemblyInstance(insts[0].assembly_index, insts[0].local_to_parent.copy()))       # coincident instance: exact ties
    desc = SceneDesc(meshes, assemblies, insts)
    lo, hi = scenes.scene_bbox(desc)
    ext = np.maximum(hi - lo, 1e-3)
    rays = scenes.uniform_sphere_rays(n_rays, lo - 0.1 * ext, hi + 0.1 * ext, 77 + seed, time=True)
    rays.flags = rng.choice(np.array([VIS_CAMERA, VIS_SHADOW, VIS_DIFFUSE, VIS_ALL], dtype=np.uint32), size=n_rays)
    diag = float(np.linalg.norm(ext))
    rays.tmin = rng.choice(np.array([0.0, 0.0, 0.1 * diag, -0.2 * diag]), size=n_rays)
    rays.tmax = rng.choice(np.array([scenes.DBL_MAX, scenes.DBL_MAX, 0.5 * diag, 1.2 * diag]), size=n_rays)
    rays.dir[::11] *= rng.uniform(0.2, 5.0)                # directions need not be unit length (instance space)
    return desc, rays


# Extreme but finite ray fields, and infinities where the reference's arithmetic stays meaningful
# (an infinite origin misses everything, an infinite tmax is "no limit", tmin = +inf is an empty
# interval): the wide kernels must still meet the parity rule on these.
EXTREME_OK = {"org": [0.0, -0.0, 5e-324, -5e-324, 1e-300, 1e308, -1e308, np.inf, -np.inf],
              "dir": [0.0, -0.0, 5e-324, -5e-324, 1e-300, 1e308, -1e308],
              "tmin": [0.0, -0.0, 5e-324, -5e-324, 1e-300, 1e308, -1e308, np.inf],
              "tmax": [0.0, -0.0, 5e-324, -5e-324, 1e-300, 1e308, -1e308, np.inf, -np.inf]}
# Values with no meaning for a ray (NaN anywhere, an infinite direction, tmin = -inf): the exact
# kernels still repeat the reference's arithmetic bit for bit; the wide kernels only promise to terminate.
EXTREME_UNDEFINED = {"org": [np.nan], "dir": [np.nan, np.inf, -np.inf], "tmin": [np.nan, -np.inf], "tmax": [np.nan]}


def extreme_rays(rays: RayBatch, table: dict, seed: int) -> RayBatch:
    """`rays` with one field of every ray replaced by a value of `table`."""
    rng = np.random.default_rng(seed)
    f = {"org": rays.org.copy(), "dir": rays.dir.copy(), "tmin": rays.tmin.copy(), "tmax": rays.tmax.copy()}
    names = sorted(table)
    for k in range(len(rays)):
        name = names[int(rng.integers(0, len(names)))]
        value = table[name][int(rng.integers(0, len(table[name])))]
        if f[name].ndim == 2:
            f[name][k, int(rng.integers(0, 3))] = value
        else:
            f[name][k] = value
    return RayBatch(f["org"], f["dir"], f["tmin"], f["tmax"], rays.time_absolute, rays.time_normalized, rays.flags)
