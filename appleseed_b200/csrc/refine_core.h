//
// refine_core.h -- ShadingPoint::refine_and_offset for one hit, shared by the CUDA kernel
// (refine.cu) and the TEST-ONLY host build (tests/hostsim), like traverse_core.h.
//
//   refine_space_ray = assembly_instance_transform.to_local(ray);  org += tmax * dir
//                                                       (the transform the traversal stored in the
//                                                       ShadingPoint, assemblytree.cpp:738-739: for an
//                                                       animated instance the one evaluated at the
//                                                       ray's absolute time)
//   org = refine(org, dir, plane)                       two Newton steps onto the triangle's
//                                                       support plane (refining.h:97-113,
//                                                       raytrianglemt.h:300-309)
//   n   = faceforward(object_instance.normal_to_parent(cross(v1 - v0, v2 - v0)), dir)
//                                                       source vertices, float cross product
//                                                       (renderer/utility/triangle.h:57-64); for a
//                                                       deforming mesh the vertices interpolated between
//                                                       the two poses around the ray time
//                                                       (fetch_triangle_source_geometry,
//                                                       shadingpoint.cpp:186-256)
//   front / back = adaptive_offset(org, normalize(n))   ulp steps of doubling size until the point
//                                                       is off the plane (refining.h:168-221)
//
// Every fp64 operation is an explicit round-to-nearest operation in the reference's order (no FMA
// contraction): the records are bit-identical to the CPU oracle's.
//
#pragma once

#include "traverse_core.h"

namespace asgpu
{

#if ASGPU_DEVICE_CODE
ASGPU_HD double dsqrt(double a) { return __dsqrt_rn(a); }
ASGPU_HD unsigned long long double_bits(double a) { return static_cast<unsigned long long>(__double_as_longlong(a)); }
ASGPU_HD double bits_double(unsigned long long b) { return __longlong_as_double(static_cast<long long>(b)); }
#else
inline double dsqrt(double a) { volatile double r = std::sqrt(a); return r; }
inline unsigned long long double_bits(double a) { unsigned long long b; std::memcpy(&b, &a, 8); return b; }
inline double bits_double(unsigned long long b) { double a; std::memcpy(&a, &b, 8); return a; }
#endif

// TriangleMTSupportPlane::intersect (raytrianglemt.h:300-309).
ASGPU_HD double plane_intersect(const TriD& tri, const double org[3], const double dir[3])
{
    const double tvec[3] = { dsub(org[0], tri.v0[0]), dsub(org[1], tri.v0[1]), dsub(org[2], tri.v0[2]) };
    double qvec[3], pvec[3];
    cross_d(tvec, tri.e0, qvec);
    cross_d(dir, tri.e1, pvec);
    return ddiv(dot_d(tri.e1, qvec), dot_d(tri.e0, pvec));
}

// adaptive_offset_point_step (refining.h:197-221).
ASGPU_HD void offset_step(double p[3], const double n[3], const long long mag)
{
    const double Threshold = 1.0e-25;
    for (int i = 0; i < 3; ++i)
    {
        if ((p[i] < 0.0 ? -p[i] : p[i]) < Threshold) p[i] = dadd(p[i], dmul(n[i], Threshold));
        else
        {
            const unsigned long long pi = double_bits(p[i]), ni = double_bits(n[i]);
            const long long step = ((pi ^ ni) >> 63) ? -mag : mag;
            p[i] = bits_double(pi + static_cast<unsigned long long>(step));
        }
    }
}

// adaptive_offset_point (refining.h:176-195).
ASGPU_HD void offset_point(const TriD& tri, const double p[3], const double n[3], double out[3])
{
    long long mag = 8;
    out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
    for (int i = 0; i < 64; ++i)
    {
        offset_step(out, n, mag);
        if (plane_intersect(tri, out, n) < 0.0) break;
        mag *= 2;
    }
}

// Source vertices of triangle `primitive` of one object instance (object space, float):
// fetch_triangle_source_geometry, shadingpoint.cpp:186-256.  Static mesh: m_vertices.  Deforming
// mesh: base_time = time_normalized * motion_segment_count (a FLOAT product: float x size_t),
// previous pose = m_vertices for base_index 0 else pose base_index - 1, next pose = pose base_index,
// v = previous * (1 - frac) + next * frac in float, two roundings per term.
ASGPU_HD void source_vertices(const uint8_t* blob, const uint8_t* so, const uint32_t primitive, const float time_normalized, float v[3][3])
{
    const uint2 ov = load8(so + offsetof(SrcObject, vertices)), ot = load8(so + offsetof(SrcObject, triangles));
    const uint8_t* verts = blob + (static_cast<uint64_t>(ov.x) | (static_cast<uint64_t>(ov.y) << 32));
    const uint8_t* tris = blob + (static_cast<uint64_t>(ot.x) | (static_cast<uint64_t>(ot.y) << 32)) + static_cast<uint64_t>(primitive) * 12;
    const uint32_t idx[3] = { load4(tris), load4(tris + 4), load4(tris + 8) };
    const uint32_t msc = load4(so + offsetof(SrcObject, motion_segment_count));
    if (msc == 0)
    {
        for (int c = 0; c < 3; ++c)
            for (int k = 0; k < 3; ++k) v[c][k] = u2f(load4(verts + static_cast<uint64_t>(idx[c]) * 12 + k * 4));
        return;
    }
    const uint2 op = load8(so + offsetof(SrcObject, poses));
    const uint8_t* poses = blob + (static_cast<uint64_t>(op.x) | (static_cast<uint64_t>(op.y) << 32));
    const double base_time = static_cast<double>(fmul(time_normalized, static_cast<float>(msc)));
    const uint32_t base_index = static_cast<uint32_t>(base_time);
    const float frac = static_cast<float>(dsub(base_time, static_cast<double>(base_index)));
    const float omf = fsub(1.0f, frac);
    for (int c = 0; c < 3; ++c)
    {
        const uint8_t* prev = base_index == 0 ? verts + static_cast<uint64_t>(idx[c]) * 12
                                              : poses + (static_cast<uint64_t>(idx[c]) * msc + (base_index - 1)) * 12;
        const uint8_t* next = poses + (static_cast<uint64_t>(idx[c]) * msc + base_index) * 12;
        for (int k = 0; k < 3; ++k)
            v[c][k] = fadd(fmul(u2f(load4(prev + k * 4)), omf), fmul(u2f(load4(next + k * 4)), frac));
    }
}

// One hit.  `item` = ItemRecord index of the hit's assembly instance; `time_absolute` = the ray's
// absolute time (read by animated instances only), `time_normalized` its normalized time (read for
// moving triangles only); writes the 80-byte asgpu_parent record (id, pad, front, back, geo_normal)
// as ten 8-byte words.
ASGPU_HD void refine_offset_one(const SceneView& s, const double world_org[3], const double world_dir[3], const float time_absolute, const float time_normalized,
                                const double t, const uint32_t item, const uint32_t object_instance, const uint32_t primitive, const uint32_t slot, double* dst)
{
    // refine_space_ray = to_local(ray), moved to the hit point.
    const uint8_t* ip = s.blob + s.items + static_cast<uint64_t>(item) * sizeof(ItemRecord);
    const uint4 meta = load16(ip + 96);
    double p[3], dir[3];
    instance_org_dir_at(s.blob, ip, meta.w, time_absolute, world_org, world_dir, p, dir);
    for (int k = 0; k < 3; ++k) p[k] = dadd(p[k], dmul(dir[k], t));

    // Support plane = the hit triangle (stored, or interpolated at the ray time), widened to double.
    TreeDesc td; load_tree_desc(s, meta.x, td);
    TriD tri;
    hit_triangle(s.blob + td.tris + static_cast<uint64_t>(slot) * sizeof(TriRecord), s.blob + td.poses, time_normalized, tri);
    for (int step = 0; step < 2; ++step)
    {
        const double tt = plane_intersect(tri, p, dir);
        for (int k = 0; k < 3; ++k) p[k] = dadd(p[k], dmul(dir[k], tt));
    }

    // Geometric normal from the source vertices.
    double nrm[3];
    {
        const uint8_t* so = s.blob + td.src_objects + static_cast<uint64_t>(object_instance) * sizeof(SrcObject);
        float v[3][3];
        source_vertices(s.blob, so, primitive, time_normalized, v);
        const float a[3] = { fsub(v[1][0], v[0][0]), fsub(v[1][1], v[0][1]), fsub(v[1][2], v[0][2]) };
        const float b[3] = { fsub(v[2][0], v[0][0]), fsub(v[2][1], v[0][1]), fsub(v[2][2], v[0][2]) };
        const double nf[3] = {
            static_cast<double>(fsub(fmul(a[1], b[2]), fmul(b[1], a[2]))),
            static_cast<double>(fsub(fmul(a[2], b[0]), fmul(b[2], a[0]))),
            static_cast<double>(fsub(fmul(a[0], b[1]), fmul(b[0], a[1]))) };
        // normal_to_parent: column k of parent_to_local's 3 x 3 block (transform.h:446-463).
        for (int k = 0; k < 3; ++k)
            nrm[k] = dadd(dadd(dmul(load_f64(so + k * 8), nf[0]), dmul(load_f64(so + (3 + k) * 8), nf[1])), dmul(load_f64(so + (6 + k) * 8), nf[2]));
        if (!(dot_d(nrm, dir) < 0.0)) { nrm[0] = -nrm[0]; nrm[1] = -nrm[1]; nrm[2] = -nrm[2]; }
    }

    // adaptive_offset: n = normalize(n) = n * (1 / norm) (vector.h:638-642, 768-773).
    const double rcp = ddiv(1.0, dsqrt(dot_d(nrm, nrm)));
    const double un[3] = { dmul(nrm[0], rcp), dmul(nrm[1], rcp), dmul(nrm[2], rcp) };
    const double mn[3] = { -un[0], -un[1], -un[2] };
    double front[3], back[3];
    offset_point(tri, p, un, front);
    offset_point(tri, p, mn, back);

    dst[0] = bits_double(static_cast<unsigned long long>(meta.z));
    for (int k = 0; k < 3; ++k) { dst[1 + k] = front[k]; dst[4 + k] = back[k]; dst[7 + k] = nrm[k]; }
}

}   // namespace asgpu
