"""Deterministic synthetic scenes and ray streams for the BASELINE.json configurations.

Everything here is seeded numpy; the same arrays feed the GPU engine and the CPU checkers so that
both consume identical bytes.

Scene recipes:
  C1  Cornell box (32 triangles) -- fixture ``tests/golden/cornell_box.npz`` made from the
      reference's OBJ meshes by ``tests/golden/make_cornell_fixture.py``
  C2  ``grid_mesh(707)``    999 698 triangles, tessellated like ``create_triangles``
      (renderer/modeling/object/meshobjectprimitives.cpp:440-452), sum-of-sines displacement
  C3  ``grid_mesh(2236)``   9 999 392 triangles, fBm displacement, 64 assembly instances
  C4  ``grid_mesh(1000)``   2 000 000 moving triangles, ``msc`` motion segments
"""

from __future__ import annotations

import math
from typing import Tuple

import numpy as np

from .scene import (VIS_ALL, VIS_CAMERA, VIS_DIFFUSE, VIS_PROBE, VIS_SHADOW, Assembly,
                    AssemblyInstance, Mesh, ObjectInstance, RayBatch, SceneDesc)

DBL_MAX = float(np.finfo(np.float64).max)


# ---------------------------------------------------------------------------------------------
# Meshes
# ---------------------------------------------------------------------------------------------

def grid_topology(res_u: int, res_v: int) -> np.ndarray:
    """Triangle indices of a (res_u x res_v)-quad grid with the reference's vertex numbering
    ``(res_u + 1) * j + i`` and winding ``(v3, v1, v0), (v3, v2, v1)``
    (meshobjectprimitives.cpp:427-452)."""
    j, i = np.meshgrid(np.arange(res_v, dtype=np.int64), np.arange(res_u, dtype=np.int64), indexing="ij")
    v0 = (res_u + 1) * j + i
    v1 = v0 + 1
    v2 = (res_u + 1) * (j + 1) + i + 1
    v3 = (res_u + 1) * (j + 1) + i
    tris = np.stack([np.stack([v3, v1, v0], -1), np.stack([v3, v2, v1], -1)], axis=2)
    return tris.reshape(-1, 3).astype(np.uint32)


def _grid_xz(res_u: int, res_v: int) -> Tuple[np.ndarray, np.ndarray]:
    # fit<size_t, float>(i, 0, res, 0, 1) then mapped to [-1, 1]; float arithmetic like the reference.
    s = (np.arange(res_u + 1, dtype=np.float32) / np.float32(res_u)).astype(np.float32)
    t = (np.arange(res_v + 1, dtype=np.float32) / np.float32(res_v)).astype(np.float32)
    x = (s * np.float32(2.0) - np.float32(1.0)).astype(np.float32)
    z = (t * np.float32(2.0) - np.float32(1.0)).astype(np.float32)
    zz, xx = np.meshgrid(z, x, indexing="ij")
    return xx.reshape(-1), zz.reshape(-1)


def sines_height(x: np.ndarray, z: np.ndarray) -> np.ndarray:
    """Fixed sum of sines used by C2 (and as the base shape of C4)."""
    x = x.astype(np.float64)
    z = z.astype(np.float64)
    h = (0.10 * np.sin(3.0 * x + 0.5) * np.cos(2.0 * z - 0.3)
         + 0.05 * np.sin(7.0 * x - 1.1) * np.sin(5.0 * z + 0.7)
         + 0.02 * np.cos(17.0 * x + 13.0 * z)
         + 0.008 * np.sin(41.0 * x - 29.0 * z + 0.2))
    return h.astype(np.float32)


def _value_noise(x: np.ndarray, z: np.ndarray, seed: int) -> np.ndarray:
    """Smooth value noise on the integer lattice (hash -> [0,1), smoothstep interpolation)."""
    def hash2(ix, iz):
        h = (ix.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
             + iz.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
             + np.uint64(seed) * np.uint64(0x165667B19E3779F9))
        h ^= h >> np.uint64(29)
        h *= np.uint64(0xBF58476D1CE4E5B9)
        h ^= h >> np.uint64(32)
        return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    x0 = np.floor(x)
    z0 = np.floor(z)
    fx = x - x0
    fz = z - z0
    ix = x0.astype(np.int64) + (1 << 20)
    iz = z0.astype(np.int64) + (1 << 20)
    sx = fx * fx * (3.0 - 2.0 * fx)
    sz = fz * fz * (3.0 - 2.0 * fz)
    with np.errstate(over="ignore"):
        a = hash2(ix, iz)
        b = hash2(ix + 1, iz)
        c = hash2(ix, iz + 1)
        d = hash2(ix + 1, iz + 1)
    return (a * (1 - sx) + b * sx) * (1 - sz) + (c * (1 - sx) + d * sx) * sz


def fbm_height(x: np.ndarray, z: np.ndarray, seed: int = 2, octaves: int = 7) -> np.ndarray:
    """Seeded fBm used by the C3 terrain."""
    x = x.astype(np.float64)
    z = z.astype(np.float64)
    h = np.zeros_like(x)
    amp, freq = 0.25, 2.0
    for o in range(octaves):
        h += amp * (_value_noise(x * freq + 17.0 * o, z * freq - 11.0 * o, seed + o) - 0.5)
        amp *= 0.5
        freq *= 2.03
    return h.astype(np.float32)


def grid_mesh(res: int, height: str = "sines", seed: int = 2) -> Mesh:
    x, z = _grid_xz(res, res)
    if height == "sines":
        y = sines_height(x, z)
    elif height == "fbm":
        y = fbm_height(x, z, seed)
    elif height == "flat":
        y = np.zeros_like(x)
    else:
        raise ValueError(height)
    verts = np.stack([x, y, z], axis=1).astype(np.float32)
    return Mesh(verts, grid_topology(res, res))


def moving_grid_mesh(res: int, msc: int, seed: int = 3) -> Mesh:
    """C4: base pose = sines grid; pose k = base + k * delta(x, z), a smooth seeded field."""
    base = grid_mesh(res, "sines")
    x = base.vertices[:, 0].astype(np.float64)
    z = base.vertices[:, 2].astype(np.float64)
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0.0, 2.0 * math.pi, size=6)
    dx = 0.010 * np.sin(5.0 * z + ph[0]) * np.cos(3.0 * x + ph[1])
    dy = 0.020 * np.sin(4.0 * x + ph[2]) * np.sin(6.0 * z + ph[3])
    dz = 0.010 * np.cos(5.0 * x + ph[4]) * np.sin(2.0 * z + ph[5])
    delta = np.stack([dx, dy, dz], axis=1)
    poses = np.empty((base.vertices.shape[0], msc, 3), dtype=np.float32)
    for m in range(msc):
        poses[:, m, :] = (base.vertices.astype(np.float64) + (m + 1) * delta).astype(np.float32)
    return Mesh(base.vertices, base.triangles, vertex_poses=poses)


# ---------------------------------------------------------------------------------------------
# Transforms
# ---------------------------------------------------------------------------------------------

def translation(tx, ty, tz) -> np.ndarray:
    m = np.eye(4)
    m[:3, 3] = (tx, ty, tz)
    return m


def scaling(sx, sy=None, sz=None) -> np.ndarray:
    sy = sx if sy is None else sy
    sz = sx if sz is None else sz
    return np.diag([sx, sy, sz, 1.0]).astype(np.float64)


def rotation_y(angle: float) -> np.ndarray:
    c, s = math.cos(angle), math.sin(angle)
    m = np.eye(4)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    return m


# ---------------------------------------------------------------------------------------------
# Scenes
# ---------------------------------------------------------------------------------------------

def single_mesh_scene(mesh: Mesh, vis_flags: int = VIS_ALL) -> SceneDesc:
    return SceneDesc([mesh], [Assembly([ObjectInstance(0, np.eye(4), vis_flags)])], [AssemblyInstance(0)])


def scene_c2(res: int = 707) -> SceneDesc:
    return single_mesh_scene(grid_mesh(res, "sines"))


def scene_c3(res: int = 2236, lattice: int = 8, seed: int = 2) -> SceneDesc:
    """One fBm terrain assembly, ``lattice**2`` assembly instances:
    translate . rotateY(k*pi/32) . uniform scale in {0.75, 1, 1.25}."""
    mesh = grid_mesh(res, "fbm", seed)
    instances = []
    scales = (0.75, 1.0, 1.25)
    k = 0
    for gz in range(lattice):
        for gx in range(lattice):
            m = translation(2.6 * (gx - (lattice - 1) / 2.0), 0.05 * ((gx * 7 + gz * 3) % 5), 2.6 * (gz - (lattice - 1) / 2.0)) \
                @ rotation_y(k * math.pi / 32.0) @ scaling(scales[k % 3])
            instances.append(AssemblyInstance(0, m))
            k += 1
    return SceneDesc([mesh], [Assembly([ObjectInstance(0)])], instances)


def scene_c4(res: int = 1000, msc: int = 1, seed: int = 3) -> SceneDesc:
    return single_mesh_scene(moving_grid_mesh(res, msc, seed))


def scene_bbox(desc: SceneDesc) -> Tuple[np.ndarray, np.ndarray]:
    """World-space bounds of a scene (float64, loose: transformed corners of per-mesh bounds)."""
    lo = np.full(3, np.inf)
    hi = np.full(3, -np.inf)
    for inst in desc.assembly_instances:
        asm = desc.assemblies[inst.assembly_index]
        for oi in asm.object_instances:
            mesh = desc.meshes[oi.mesh_index]
            pts = mesh.vertices.astype(np.float64)
            mlo, mhi = pts.min(0), pts.max(0)
            if mesh.vertex_poses is not None:
                pp = mesh.vertex_poses.reshape(-1, 3).astype(np.float64)
                mlo, mhi = np.minimum(mlo, pp.min(0)), np.maximum(mhi, pp.max(0))
            corners = np.array([[x, y, z, 1.0] for x in (mlo[0], mhi[0]) for y in (mlo[1], mhi[1]) for z in (mlo[2], mhi[2])])
            w = (inst.local_to_parent @ oi.local_to_parent @ corners.T).T[:, :3]
            lo, hi = np.minimum(lo, w.min(0)), np.maximum(hi, w.max(0))
    return lo, hi


# ---------------------------------------------------------------------------------------------
# Rays
# ---------------------------------------------------------------------------------------------

def _normalize(v: np.ndarray) -> np.ndarray:
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def pinhole_rays(width: int, height: int, origin, target, up=(0.0, 1.0, 0.0), film=0.025, focal=0.035,
                 flags: int = VIS_CAMERA) -> RayBatch:
    """Pinhole primaries through pixel centres, like ``PinholeCamera::spawn_ray``
    (renderer/modeling/camera/pinholecamera.cpp:159-195, perspectivecamera.cpp:204-211):
    ndc = ((x + .5)/w, (y + .5)/h); camera-space target ((.5 - ndc.x) * film_w, (ndc.y - .5) * film_h, focal),
    negated, rotated to world space and normalised."""
    origin = np.asarray(origin, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    zaxis = origin - target
    zaxis /= np.linalg.norm(zaxis)
    xaxis = np.cross(np.asarray(up, dtype=np.float64), zaxis)
    xaxis /= np.linalg.norm(xaxis)
    yaxis = np.cross(zaxis, xaxis)
    px, py = np.meshgrid(np.arange(width), np.arange(height), indexing="xy")
    ndx = (px.reshape(-1) + 0.5) / width
    ndy = (py.reshape(-1) + 0.5) / height
    film_w = film
    film_h = film * height / width
    cx = (0.5 - ndx) * film_w
    cy = (ndy - 0.5) * film_h
    cz = np.full_like(cx, focal)
    d = -(cx[:, None] * xaxis[None] + cy[:, None] * yaxis[None] + cz[:, None] * zaxis[None])
    d = _normalize(d)
    n = d.shape[0]
    return RayBatch(np.broadcast_to(origin, (n, 3)).copy(), d, np.zeros(n), np.full(n, DBL_MAX),
                    flags=np.full(n, flags, dtype=np.uint32))


def sample_hemisphere_cosine(s: np.ndarray) -> np.ndarray:
    """foundation/math/sampling/mappings.h:299-314: cosine-weighted direction about +Y."""
    phi = 2.0 * math.pi * s[:, 0]
    cos_theta = np.sqrt(1.0 - s[:, 1])
    sin_theta = np.sqrt(s[:, 1])
    return np.stack([np.cos(phi) * sin_theta, cos_theta, np.sin(phi) * sin_theta], axis=1)


def _basis_from_normal(n: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    a = np.where(np.abs(n[:, 0:1]) > 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    u = _normalize(np.cross(a, n))
    v = np.cross(n, u)
    return u, v


def bounce_rays(org: np.ndarray, normal: np.ndarray, seed: int, flags: int = VIS_DIFFUSE,
                tmax: float = DBL_MAX, offset: float = 1.0e-6) -> RayBatch:
    """Cosine-weighted rays about ``normal`` from ``org`` (offset along the normal; the
    parent == nullptr convention of SURVEY.md section 8(d))."""
    rng = np.random.default_rng(seed)
    s = rng.random((org.shape[0], 2))
    local = sample_hemisphere_cosine(s)
    u, v = _basis_from_normal(normal)
    d = _normalize(local[:, 0:1] * u + local[:, 1:2] * normal + local[:, 2:3] * v)
    o = org + offset * normal
    n = o.shape[0]
    return RayBatch(o, d, np.zeros(n), np.full(n, tmax), flags=np.full(n, flags, dtype=np.uint32))


def uniform_sphere_rays(n: int, lo: np.ndarray, hi: np.ndarray, seed: int, flags: int = VIS_DIFFUSE,
                        time: bool = False) -> RayBatch:
    """Origins uniform in the box [lo, hi], directions uniform on the sphere."""
    rng = np.random.default_rng(seed)
    o = lo[None] + rng.random((n, 3)) * (hi - lo)[None]
    z = 1.0 - 2.0 * rng.random(n)
    phi = 2.0 * math.pi * rng.random(n)
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    d = _normalize(np.stack([r * np.cos(phi), z, r * np.sin(phi)], axis=1))
    tn = rng.random(n, dtype=np.float32) if time else None
    if tn is not None:
        tn = np.minimum(tn, np.float32(1.0) - np.float32(2.0 ** -24))  # m_normalized in [0, 1)
    return RayBatch(o, d, np.zeros(n), np.full(n, DBL_MAX), time_absolute=tn, time_normalized=tn,
                    flags=np.full(n, flags, dtype=np.uint32))


def shadow_rays(points: np.ndarray, lights: np.ndarray, seed: int, flags: int = VIS_SHADOW) -> RayBatch:
    """One shadow probe per point towards one of ``lights`` with tmax = dist * (1 - 1e-6)
    (renderer/kernel/lighting/tracer.h:252-259)."""
    rng = np.random.default_rng(seed)
    li = rng.integers(0, lights.shape[0], size=points.shape[0])
    v = lights[li] - points
    dist = np.linalg.norm(v, axis=1)
    d = v / dist[:, None]
    n = points.shape[0]
    return RayBatch(points, d, np.zeros(n), dist * (1.0 - 1.0e-6), flags=np.full(n, flags, dtype=np.uint32))


def hit_points_and_normals(desc: SceneDesc, rays: RayBatch, hits: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """World-space hit points and geometric normals (flipped against the ray) of closest-hit
    records; returns (mask, points, normals) for the rays that hit.  Static meshes only."""
    mask = hits["prim_type"] == 2
    idx = np.nonzero(mask)[0]
    pts = rays.org[idx] + hits["t"][idx, None] * rays.dir[idx]
    nrm = np.zeros_like(pts)
    for ai in np.unique(hits["assembly_instance"][idx]):
        inst = desc.assembly_instances[int(ai)]
        asm = desc.assemblies[inst.assembly_index]
        sel_a = hits["assembly_instance"][idx] == ai
        for oi_index in np.unique(hits["object_instance_index"][idx][sel_a]):
            oi = asm.object_instances[int(oi_index)]
            mesh = desc.meshes[oi.mesh_index]
            sel = sel_a & (hits["object_instance_index"][idx] == oi_index)
            tri = mesh.triangles[hits["primitive_index"][idx][sel]]
            m = inst.local_to_parent @ oi.local_to_parent
            v = mesh.vertices.astype(np.float64)
            p0 = v[tri[:, 0]] @ m[:3, :3].T + m[:3, 3]
            p1 = v[tri[:, 1]] @ m[:3, :3].T + m[:3, 3]
            p2 = v[tri[:, 2]] @ m[:3, :3].T + m[:3, 3]
            nn = np.cross(p1 - p0, p2 - p0)
            nn /= np.maximum(np.linalg.norm(nn, axis=1, keepdims=True), 1e-300)
            nrm[sel] = nn
    flip = np.sum(nrm * rays.dir[idx], axis=1) > 0.0
    nrm[flip] *= -1.0
    return mask, pts, nrm


# ---------------------------------------------------------------------------------------------
# C1: Cornell box
# ---------------------------------------------------------------------------------------------

def _golden_dir() -> str:
    import os
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_cornell() -> dict:
    import os
    return dict(np.load(os.path.join(_golden_dir(), "cornell_box.npz")))


def scene_c1() -> SceneDesc:
    """The 8 meshes (32 triangles) of the reference's Cornell box, one assembly, one identity
    assembly instance; object-instance transforms are the scene file's diag(0.00100000004749700)."""
    fx = load_cornell()
    n = len(fx["names"])
    meshes = [Mesh(fx["vertices_%d" % i], fx["triangles_%d" % i]) for i in range(n)]
    ois = [ObjectInstance(i, fx["transform_%d" % i]) for i in range(n)]
    return SceneDesc(meshes, [Assembly(ois)], [AssemblyInstance(0)])


def camera_rays(width: int, height: int, camera_matrix: np.ndarray, film=(0.025, 0.025), focal: float = 0.035,
                flags: int = VIS_CAMERA) -> RayBatch:
    """``PinholeCamera::spawn_ray`` (pinholecamera.cpp:159-195) through pixel centres:
    org = translation of the camera matrix; dir = normalize(M3x3 . -ndc_to_camera(ndc)) with
    ``ndc_to_camera`` = ((0.5 - x) * film_w, (y - 0.5) * film_h, focal) (perspectivecamera.cpp:204-211)."""
    m = np.asarray(camera_matrix, dtype=np.float64).reshape(4, 4)
    px, py = np.meshgrid(np.arange(width), np.arange(height), indexing="xy")
    ndx = (px.reshape(-1) + 0.5) / width
    ndy = (py.reshape(-1) + 0.5) / height
    cam = -np.stack([(0.5 - ndx) * film[0], (ndy - 0.5) * film[1], np.full_like(ndx, focal)], axis=1)
    d = _normalize(cam @ m[:3, :3].T)
    n = d.shape[0]
    org = np.broadcast_to(m[:3, 3], (n, 3)).copy()
    return RayBatch(org, d, np.zeros(n), np.full(n, DBL_MAX), flags=np.full(n, flags, dtype=np.uint32))


def rays_c1_primary() -> RayBatch:
    fx = load_cornell()
    w, h = (int(x) for x in fx["resolution"])
    return camera_rays(w, h, fx["camera_matrix"], tuple(fx["film_dimensions"]), float(fx["focal_length"]))


def rays_c1_ao(desc: SceneDesc, primary: RayBatch, hits: np.ndarray, seed: int = 0xA55EED) -> Tuple[np.ndarray, RayBatch]:
    """One cosine-weighted ambient-occlusion probe per primary hit
    (renderer/kernel/shading/ambientocclusion.h:56-113): tmin = 0, tmax = 1.0 (ao_surface_shader
    default max_distance), flags ProbeRay, origin offset 1e-6 * scene diagonal along the normal."""
    mask, pts, nrm = hit_points_and_normals(desc, primary, hits)
    lo, hi = scene_bbox(desc)
    diag = float(np.linalg.norm(hi - lo))
    rays = bounce_rays(pts, nrm, seed, flags=VIS_PROBE, tmax=1.0, offset=1.0e-6 * diag)
    return mask, rays
