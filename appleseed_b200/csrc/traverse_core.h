//
// traverse_core.h -- per-ray traversal and intersection, shared by the CUDA kernels (kernels.cu)
// and by a TEST-ONLY host build (tests/hostsim) that runs the very same code on the CPU so the
// logic can be checked against the oracle without a GPU.  The shipped library only contains the
// CUDA instantiation; there is no CPU execution path in the product.
//
// Two traversals over the two layouts of gpu_layout.h:
//
//   exact_trace<ANY>  1:1 restatement of the reference's two-level traversal in fp64 on the EXACT
//                     layout: generic scalar top-level loop (bvh_intersector.h:139-258) with
//                     AssemblyLeaf[Probe]Visitor (assemblytree.cpp:604-838, 845-1054), SSE2-order
//                     bottom-level loops (bvh_intersector.h:472-616, 623-890) with
//                     TriangleLeaf[Probe]Visitor (triangletree.cpp:1352-1603).  Same visit order,
//                     same operations, no FMA: results are bit-identical to the reference's.
//
//   wide_trace<ANY>   throughput traversal on the WIDE layout: 8-wide nodes, quantised child
//                     boxes tested in fp32 INTERVAL arithmetic (directed rounding) so that a box is
//                     never missed when the reference's fp64 slab test on the tighter binary box
//                     would pass; candidate triangles then go through the same exact fp64
//                     Moeller-Trumbore test as above.  The accepted set of triangles is therefore a
//                     superset of the reference's and the nearest accepted one is reported.
//
// All fp64 arithmetic that must match the reference uses explicit round-to-nearest intrinsics
// (never contracted into FMAs).
//
#pragma once

#include "../../include/asgpu.h"
#include "gpu_layout.h"

#include <cstdint>

#if defined(__CUDA_ARCH__)
    #define ASGPU_DEVICE_CODE 1
#else
    #define ASGPU_DEVICE_CODE 0
#endif

#if defined(__CUDACC__)
    #define ASGPU_HD __host__ __device__ __forceinline__
#else
    #define ASGPU_HD inline
#endif

#if !ASGPU_DEVICE_CODE
    #include <cfenv>
    #include <cmath>
    #include <cstring>
#endif

namespace asgpu
{

// ------------------------------------------------------------------------------------------
// Arithmetic primitives.
// ------------------------------------------------------------------------------------------

#if ASGPU_DEVICE_CODE

ASGPU_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
ASGPU_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
ASGPU_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
ASGPU_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
ASGPU_HD float  fmul(float a, float b)   { return __fmul_rn(a, b); }
ASGPU_HD float  fadd(float a, float b)   { return __fadd_rn(a, b); }
ASGPU_HD float  fsub(float a, float b)   { return __fsub_rn(a, b); }
ASGPU_HD float  fsub_dn(float a, float b) { return __fsub_rd(a, b); }
ASGPU_HD float  fsub_up(float a, float b) { return __fsub_ru(a, b); }
ASGPU_HD float  fmul_dn(float a, float b) { return __fmul_rd(a, b); }
ASGPU_HD float  fmul_up(float a, float b) { return __fmul_ru(a, b); }
ASGPU_HD float  fma_dn(float a, float b, float c) { return __fmaf_rd(a, b, c); }
ASGPU_HD float  fma_up(float a, float b, float c) { return __fmaf_ru(a, b, c); }
ASGPU_HD float  d2f_dn(double a) { return __double2float_rd(a); }
ASGPU_HD float  d2f_up(double a) { return __double2float_ru(a); }
// Bounds of 1 / a for a >= 0 from the hardware approximation (rcp.approx.f32, relative error
// <= 2^-23 per the PTX ISA; the margin allows 2^-21) instead of a correctly rounded division:
// frcp_dn(a) <= 1 / a <= frcp_up(a).  The lower bound is kept finite (and 0 where the quotient
// would be subnormal); frcp_dn(0) = frcp_up(0) = +inf.
ASGPU_HD float  rcp_approx(float a) { float r; asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
ASGPU_HD float  frcp_dn(float a)
{
    if (a == 0.0f) return __int_as_float(0x7F800000);
    if (a > 4.0e37f) return 0.0f;
    const float r = __fmul_rd(rcp_approx(a), 0.99999952316284f);        // 1 - 2^-21
    return r == __int_as_float(0x7F800000) ? 3.402823466e38f : r;       // a NaN stays a NaN
}
ASGPU_HD float  frcp_up(float a)
{
    return fmaxf(__fmul_ru(rcp_approx(a), 1.00000047683716f), 3.0e-38f);    // 1 + 2^-21; NaN -> 3e-38 only for a NaN ray
}
ASGPU_HD bool   sign_bit(double a) { return __double2hiint(a) < 0; }
ASGPU_HD float  fmin_nan(float a, float b) { return fminf(a, b); }     // returns the non-NaN operand
ASGPU_HD float  fmax_nan(float a, float b) { return fmaxf(a, b); }
ASGPU_HD float  bits_to_float(uint32_t u) { return __uint_as_float(u); }
ASGPU_HD uint32_t float_to_bits(float f) { return __float_as_uint(f); }
ASGPU_HD int    popc(uint32_t x) { return __popc(x); }
ASGPU_HD int    high_bit(uint32_t x) { return 31 - __clz(x); }
ASGPU_HD uint4  load16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
ASGPU_HD uint2  load8(const void* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
ASGPU_HD uint32_t load4(const void* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
ASGPU_HD double load_f64(const void* p) { return __ldg(reinterpret_cast<const double*>(p)); }

#else

struct uint4_h { uint32_t x, y, z, w; };
struct uint2_h { uint32_t x, y; };
#if !defined(__CUDACC__)
typedef uint4_h uint4;
typedef uint2_h uint2;
#endif

// Host stand-ins (tests/hostsim is compiled with -ffp-contract=off -frounding-math).
struct RoundScope
{
    int saved;
    explicit RoundScope(int mode) : saved(fegetround()) { fesetround(mode); }
    ~RoundScope() { fesetround(saved); }
};
inline double dmul(double a, double b) { volatile double r = a * b; return r; }
inline double dadd(double a, double b) { volatile double r = a + b; return r; }
inline double dsub(double a, double b) { volatile double r = a - b; return r; }
inline double ddiv(double a, double b) { volatile double r = a / b; return r; }
inline float  fmul(float a, float b)   { volatile float r = a * b; return r; }
inline float  fadd(float a, float b)   { volatile float r = a + b; return r; }
inline float  fsub(float a, float b)   { volatile float r = a - b; return r; }
inline float  fsub_dn(float a, float b) { RoundScope s(FE_DOWNWARD); volatile float x = a, y = b; volatile float r = x - y; return r; }
inline float  fsub_up(float a, float b) { RoundScope s(FE_UPWARD);   volatile float x = a, y = b; volatile float r = x - y; return r; }
inline float  fmul_dn(float a, float b) { RoundScope s(FE_DOWNWARD); volatile float x = a, y = b; volatile float r = x * y; return r; }
inline float  fmul_up(float a, float b) { RoundScope s(FE_UPWARD);   volatile float x = a, y = b; volatile float r = x * y; return r; }
inline float  fma_dn(float a, float b, float c) { RoundScope s(FE_DOWNWARD); volatile float x = a, y = b, z = c; volatile float r = std::fmaf(x, y, z); return r; }
inline float  fma_up(float a, float b, float c) { RoundScope s(FE_UPWARD);   volatile float x = a, y = b, z = c; volatile float r = std::fmaf(x, y, z); return r; }
inline float  d2f_dn(double a) { RoundScope s(FE_DOWNWARD); volatile double x = a; volatile float r = static_cast<float>(x); return r; }
inline float  d2f_up(double a) { RoundScope s(FE_UPWARD);   volatile double x = a; volatile float r = static_cast<float>(x); return r; }
inline float  frcp_dn(float a) { RoundScope s(FE_DOWNWARD); volatile float x = a, one = 1.0f; volatile float r = one / x; return r; }
inline float  frcp_up(float a) { RoundScope s(FE_UPWARD);   volatile float x = a, one = 1.0f; volatile float r = one / x; return r; }
inline bool   sign_bit(double a) { return std::signbit(a); }
inline float  fmin_nan(float a, float b) { return std::fmin(a, b); }
inline float  fmax_nan(float a, float b) { return std::fmax(a, b); }
inline float  bits_to_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t float_to_bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline int    popc(uint32_t x) { return __builtin_popcount(x); }
inline int    high_bit(uint32_t x) { return 31 - __builtin_clz(x); }
inline uint4  load16(const void* p) { uint4 r; std::memcpy(&r, p, 16); return r; }
inline uint2  load8(const void* p) { uint2 r; std::memcpy(&r, p, 8); return r; }
inline uint32_t load4(const void* p) { uint32_t r; std::memcpy(&r, p, 4); return r; }
inline double load_f64(const void* p) { double r; std::memcpy(&r, p, 8); return r; }

#endif

ASGPU_HD float u2f(uint32_t u) { return bits_to_float(u); }

// ------------------------------------------------------------------------------------------
// Scene view, ray, hit.
// ------------------------------------------------------------------------------------------

struct SceneView
{
    const uint8_t*  blob;
    uint64_t        trees, items, top_nodes, top_wnodes, top_witems;
    uint32_t        tree_count, item_count, top_node_count, top_wnode_count;
    uint32_t        wide_stack_need;
    uint32_t        has_motion;         // some tree has moving triangles (time-sliced child planes)
    uint32_t        has_filters;        // some tree carries intersection filters
    uint32_t        has_animated;       // some assembly instance has a multi-key transform sequence
};

struct Ray
{
    double      org[3], dir[3];
    double      tmin, tmax;
    float       time_absolute, time_normalized;
    uint32_t    flags;
};

struct Hit
{
    float       u, v;
    uint32_t    item;           // ItemRecord index, 0xFFFFFFFF = none
    uint32_t    slot;           // reference leaf-order slot
    uint32_t    segment;
};

struct Stats
{
    uint32_t    top_nodes, instances, nodes, triangles;
};

ASGPU_HD void load_ray(const asgpu_rays& rays, const size_t i, Ray& r)
{
#if ASGPU_DEVICE_CODE
    r.org[0] = __ldg(rays.org + i * 3); r.org[1] = __ldg(rays.org + i * 3 + 1); r.org[2] = __ldg(rays.org + i * 3 + 2);
    r.dir[0] = __ldg(rays.dir + i * 3); r.dir[1] = __ldg(rays.dir + i * 3 + 1); r.dir[2] = __ldg(rays.dir + i * 3 + 2);
    r.tmin = __ldg(rays.tmin + i);
    r.tmax = __ldg(rays.tmax + i);
    r.time_absolute = rays.time_absolute ? __ldg(rays.time_absolute + i) : 0.0f;
    r.time_normalized = rays.time_normalized ? __ldg(rays.time_normalized + i) : 0.0f;
    r.flags = rays.flags ? __ldg(rays.flags + i) : 0xFFFFFFFFu;
#else
    for (int a = 0; a < 3; ++a) { r.org[a] = rays.org[i * 3 + a]; r.dir[a] = rays.dir[i * 3 + a]; }
    r.tmin = rays.tmin[i];
    r.tmax = rays.tmax[i];
    r.time_absolute = rays.time_absolute ? rays.time_absolute[i] : 0.0f;
    r.time_normalized = rays.time_normalized ? rays.time_normalized[i] : 0.0f;
    r.flags = rays.flags ? rays.flags[i] : 0xFFFFFFFFu;
#endif
}

// compute_assembly_instance_ray (assemblytree.cpp:556-596) with Transform::vector_to_local /
// point_to_local (transform.h:311-344, 381-400): products accumulated left to right; the
// instance matrix is affine so w == 1 and no division happens.
ASGPU_HD void instance_org_dir(const uint8_t* item, const double worg[3], const double wdir[3], double lorg[3], double ldir[3])
{
    double m[12];
#if ASGPU_DEVICE_CODE
    #pragma unroll
#endif
    for (int k = 0; k < 12; ++k) m[k] = load_f64(item + k * 8);
#if ASGPU_DEVICE_CODE
    #pragma unroll
#endif
    for (int r = 0; r < 3; ++r)
    {
        const double* row = m + r * 4;
        ldir[r] = dadd(dadd(dmul(row[0], wdir[0]), dmul(row[1], wdir[1])), dmul(row[2], wdir[2]));
        lorg[r] = dadd(dadd(dadd(dmul(row[0], worg[0]), dmul(row[1], worg[1])), dmul(row[2], worg[2])), row[3]);
    }
}

// Animated assembly instance: world -> instance rows of the transform at the ray's absolute time.
// TransformSequence::evaluate (transformsequence.h:185-210): the first / last key outside the key
// range, else interpolate (transformsequence.cpp:331-356: binary search over the key times,
// t = (time - begin) / (end - begin) in float) and TransformInterpolator::evaluate
// (transform.h:694-789: fast_slerp of the rotations, quaternion.h:500-511, lerp of scale and
// translation, analytic inverse).  fp64, the reference's operation order, no contraction.
ASGPU_HD void animated_item_matrix(const uint8_t* blob, const uint8_t* item, const uint32_t key_count, const float time, double m[12])
{
    const uint2 mo = load8(item + 112);
    const uint8_t* motion = blob + (static_cast<uint64_t>(mo.x) | (static_cast<uint64_t>(mo.y) << 32));
    const uint8_t* keys = motion + (static_cast<uint64_t>(key_count) * 4 + 15) / 16 * 16;
    const uint8_t* segments = keys + static_cast<uint64_t>(key_count) * 96;
    const float first = u2f(load4(motion)), last = u2f(load4(motion + static_cast<uint64_t>(key_count - 1) * 4));
    if (time <= first || time >= last)
    {
        const uint8_t* k = time <= first ? keys : keys + static_cast<uint64_t>(key_count - 1) * 96;
        for (int e = 0; e < 12; ++e) m[e] = load_f64(k + e * 8);
        return;
    }
    uint32_t begin = 0, end = key_count;
    while (end - begin > 1)
    {
        const uint32_t mid = (begin + end) / 2;
        if (time < u2f(load4(motion + static_cast<uint64_t>(mid) * 4))) end = mid; else begin = mid;
    }
    const float begin_time = u2f(load4(motion + static_cast<uint64_t>(begin) * 4)), end_time = u2f(load4(motion + static_cast<uint64_t>(end) * 4));
#if ASGPU_DEVICE_CODE
    const double t = static_cast<double>(__fdiv_rn(fsub(time, begin_time), fsub(end_time, begin_time)));
#else
    volatile float tf = fsub(time, begin_time) / fsub(end_time, begin_time);
    const double t = static_cast<double>(tf);
#endif
    const uint8_t* sg = segments + static_cast<uint64_t>(begin) * 160;
    double s0[3], q0[4], t0[3], s1[3], q1[4], t1[3];
    for (int a = 0; a < 3; ++a)
    {
        s0[a] = load_f64(sg + a * 8);          t0[a] = load_f64(sg + (7 + a) * 8);
        s1[a] = load_f64(sg + (10 + a) * 8);   t1[a] = load_f64(sg + (17 + a) * 8);
    }
    for (int a = 0; a < 4; ++a) { q0[a] = load_f64(sg + (3 + a) * 8); q1[a] = load_f64(sg + (13 + a) * 8); }

    // fast_slerp: d = dot(p, q) = p.s * q.s + dot(p.v, q.v) (dot accumulates from 0).
    double dv = dadd(0.0, dmul(q0[1], q1[1])); dv = dadd(dv, dmul(q0[2], q1[2])); dv = dadd(dv, dmul(q0[3], q1[3]));
    const double d = dadd(dmul(q0[0], q1[0]), dv);
    const double pa = dadd(1.0904, dmul(d, dadd(-3.2452, dmul(d, dadd(3.55645, dmul(d, -1.43519))))));
    const double pb = dadd(0.848013, dmul(d, dadd(-1.06021, dmul(d, 0.215638))));
    const double u = dsub(t, 1.0), v = dsub(t, 0.5);
    const double k = dadd(dmul(dmul(pa, v), v), pb);
    const double w = dadd(dmul(dmul(dmul(k, u), v), t), t);
    // normalize(lerp(p, q, w)): lerp = (1 - w) * p + w * q, normalize = q * (1 / sqrt(dot(q, q))).
    const double omw = dsub(1.0, w);
    double q[4];
    for (int a = 0; a < 4; ++a) q[a] = dadd(dmul(q0[a], omw), dmul(q1[a], w));
    double nv = dadd(0.0, dmul(q[1], q[1])); nv = dadd(nv, dmul(q[2], q[2])); nv = dadd(nv, dmul(q[3], q[3]));
    const double nn = dadd(dmul(q[0], q[0]), nv);
#if ASGPU_DEVICE_CODE
    const double rn = ddiv(1.0, __dsqrt_rn(nn));
#else
    volatile double sq = std::sqrt(nn);
    const double rn = ddiv(1.0, sq);
#endif
    const double qs = dmul(q[0], rn), qx = dmul(q[1], rn), qy = dmul(q[2], rn), qz = dmul(q[3], rn);

    const double rtx = dadd(qx, qx), rty = dadd(qy, qy), rtz = dadd(qz, qz);
    const double twx = dmul(rtx, qs), twy = dmul(rty, qs), twz = dmul(rtz, qs);
    const double txx = dmul(rtx, qx), txy = dmul(rty, qx), txz = dmul(rtz, qx);
    const double tyy = dmul(rty, qy), tyz = dmul(rtz, qy), tzz = dmul(rtz, qz);
    // parent_to_local rotation block (the transpose of local_to_parent's).
    m[0] = dsub(1.0, dadd(tyy, tzz));  m[1] = dadd(txy, twz);             m[2] = dsub(txz, twy);
    m[4] = dsub(txy, twz);             m[5] = dsub(1.0, dadd(txx, tzz));  m[6] = dadd(tyz, twx);
    m[8] = dadd(txz, twy);             m[9] = dsub(tyz, twx);             m[10] = dsub(1.0, dadd(txx, tyy));
    // s = lerp(s0, s1, t); rows scaled by 1 / s.
    const double omt = dsub(1.0, t);
    for (int r = 0; r < 3; ++r)
    {
        const double sr = dadd(dmul(omt, s0[r]), dmul(t, s1[r]));
        const double rcp = ddiv(1.0, sr);
        m[r * 4 + 0] = dmul(m[r * 4 + 0], rcp); m[r * 4 + 1] = dmul(m[r * 4 + 1], rcp); m[r * 4 + 2] = dmul(m[r * 4 + 2], rcp);
    }
    // p = lerp(t0, t1, t); translation column = -(row . p).
    double p[3];
    for (int a = 0; a < 3; ++a) p[a] = dadd(dmul(omt, t0[a]), dmul(t, t1[a]));
    for (int r = 0; r < 3; ++r)
        m[r * 4 + 3] = -dadd(dadd(dmul(m[r * 4 + 0], p[0]), dmul(m[r * 4 + 1], p[1])), dmul(m[r * 4 + 2], p[2]));
}

// instance_org_dir for an item that may be animated (key_count = the item's meta word 3).
ASGPU_HD void instance_org_dir_at(const uint8_t* blob, const uint8_t* item, const uint32_t key_count, const float time_absolute,
                                  const double worg[3], const double wdir[3], double lorg[3], double ldir[3])
{
    if (key_count < 2) { instance_org_dir(item, worg, wdir, lorg, ldir); return; }
    double m[12];
    animated_item_matrix(blob, item, key_count, time_absolute, m);
    for (int r = 0; r < 3; ++r)
    {
        const double* row = m + r * 4;
        ldir[r] = dadd(dadd(dmul(row[0], wdir[0]), dmul(row[1], wdir[1])), dmul(row[2], wdir[2]));
        lorg[r] = dadd(dadd(dadd(dmul(row[0], worg[0]), dmul(row[1], worg[1])), dmul(row[2], worg[2])), row[3]);
    }
}

ASGPU_HD void to_instance_space(const uint8_t* blob, const uint8_t* item, const uint32_t key_count, const Ray& world, Ray& local)
{
    instance_org_dir_at(blob, item, key_count, world.time_absolute, world.org, world.dir, local.org, local.dir);
    local.tmin = world.tmin;
    local.tmax = world.tmax;
    local.time_absolute = world.time_absolute;
    local.time_normalized = world.time_normalized;
    local.flags = world.flags;
}

// compute_assembly_instance_ray with a parent shading point (assemblytree.cpp:565-576): inside the
// assembly instance that holds the parent's hit, the child ray starts from the parent's offset
// point -- front when dot(geo_normal, dir) > 0, else back (ShadingPoint::get_offset_point,
// shadingpoint.h:604-613; dot accumulates from 0, vector.h:745-753).  `parent` points at this ray's
// asgpu_parent record (80 bytes: id, pad, front, back, geo_normal) or is null.
ASGPU_HD void parent_origin(const uint8_t* parent, const uint32_t assembly_instance, const double ldir[3], double lorg[3])
{
    if (parent == nullptr || load4(parent) != assembly_instance) return;
    double d = dadd(0.0, dmul(load_f64(parent + 56), ldir[0]));
    d = dadd(d, dmul(load_f64(parent + 64), ldir[1]));
    d = dadd(d, dmul(load_f64(parent + 72), ldir[2]));
    const uint8_t* src = parent + (d > 0.0 ? 8 : 32);
    lorg[0] = load_f64(src); lorg[1] = load_f64(src + 8); lorg[2] = load_f64(src + 16);
}

// ------------------------------------------------------------------------------------------
// Exact Moeller-Trumbore (raytrianglemt.h:148-268) on a float triangle widened to double
// (raytrianglemt.h:139-146).  cross: vector.h:1239-1246; dot accumulates from 0: vector.h:745-753.
// ------------------------------------------------------------------------------------------

struct TriD { double v0[3], e0[3], e1[3]; };

ASGPU_HD void cross_d(const double a[3], const double b[3], double r[3])
{
    r[0] = dsub(dmul(a[1], b[2]), dmul(b[1], a[2]));
    r[1] = dsub(dmul(a[2], b[0]), dmul(b[2], a[0]));
    r[2] = dsub(dmul(a[0], b[1]), dmul(b[0], a[1]));
}

ASGPU_HD double dot_d(const double a[3], const double b[3])
{
    double r = dadd(0.0, dmul(a[0], b[0]));
    r = dadd(r, dmul(a[1], b[1]));
    r = dadd(r, dmul(a[2], b[2]));
    return r;
}

// Returns true when the ray hits within [tmin, tmax); WITH_TUV also produces the scaled t, u, v.
template <bool WITH_TUV>
ASGPU_HD bool mt_test(const TriD& tri, const Ray& ray, double& t, double& u, double& v)
{
    double pvec[3]; cross_d(ray.dir, tri.e1, pvec);
    const double det = dot_d(tri.e0, pvec);
    const double tvec[3] = { dsub(ray.org[0], tri.v0[0]), dsub(ray.org[1], tri.v0[1]), dsub(ray.org[2], tri.v0[2]) };
    double qvec[3];
    double tt, uu, vv;
    if (det > 0.0)
    {
        uu = dot_d(tvec, pvec);
        if (uu < 0.0 || uu > det) return false;
        cross_d(tvec, tri.e0, qvec);
        vv = dot_d(ray.dir, qvec);
        if (vv < 0.0 || dadd(uu, vv) > det) return false;
        tt = dot_d(tri.e1, qvec);
        if (tt >= dmul(ray.tmax, det) || tt < dmul(ray.tmin, det)) return false;
    }
    else
    {
        uu = dot_d(tvec, pvec);
        if (uu > 0.0 || uu < det) return false;
        cross_d(tvec, tri.e0, qvec);
        vv = dot_d(ray.dir, qvec);
        if (vv > 0.0 || dadd(uu, vv) < det) return false;
        tt = dot_d(tri.e1, qvec);
        if (tt <= dmul(ray.tmax, det) || tt > dmul(ray.tmin, det)) return false;
    }
    if (WITH_TUV)
    {
        const double rcp_det = ddiv(1.0, det);
        t = dmul(tt, rcp_det);
        u = dmul(uu, rcp_det);
        v = dmul(vv, rcp_det);
    }
    return true;
}

// Loads the triangle of one record for this ray's time.  Returns false when the triangle is
// invisible to the ray (triangletree.cpp:1389-1393, 1426-1430).  Moving triangles follow
// triangletree.cpp:1432-1451 (closest hit: float product) / :1570-1585 (probe: double product).
// MOTION = false: the caller knows the scene has no moving triangle (the pose code is compiled out).
template <bool ANY, bool MOTION = true>
ASGPU_HD bool fetch_triangle(const uint8_t* record, const uint8_t* poses, const Ray& ray, TriD& tri, uint32_t& slot, uint32_t& segment)
{
    const uint4 a = load16(record);
    const uint4 b = load16(record + 16);
    const uint4 c = load16(record + 32);
    const uint32_t vis = c.y;
    slot = c.z;
    segment = 0;
    if (!(vis & ray.flags)) return false;
    float f[9];
    if (!MOTION || c.w == 0)
    {
        f[0] = u2f(a.x); f[1] = u2f(a.y); f[2] = u2f(a.z); f[3] = u2f(a.w);
        f[4] = u2f(b.x); f[5] = u2f(b.y); f[6] = u2f(b.z); f[7] = u2f(b.w);
        f[8] = u2f(c.x);
    }
    else
    {
        const uint32_t msc = a.x;
        double base_time;
        if (ANY) base_time = dmul(static_cast<double>(ray.time_normalized), static_cast<double>(msc));
        else base_time = static_cast<double>(fmul(ray.time_normalized, static_cast<float>(msc)));
        const uint32_t base_index = static_cast<uint32_t>(base_time);
        const float frac = static_cast<float>(dsub(base_time, static_cast<double>(base_index)));
        const float omf = fsub(1.0f, frac);
        const uint8_t* p = poses + (static_cast<uint64_t>(c.w - 1) + static_cast<uint64_t>(base_index) * 9) * 4;
        float v[9];
#if ASGPU_DEVICE_CODE
        #pragma unroll
#endif
        for (int k = 0; k < 9; ++k)
        {
            const float p0 = u2f(load4(p + k * 4));
            const float p1 = u2f(load4(p + 36 + k * 4));
            v[k] = fadd(fmul(p0, omf), fmul(p1, frac));
        }
        // TriangleMT<float>(v0, v1, v2): edges in float (raytrianglemt.h:128-137).
        f[0] = v[0]; f[1] = v[1]; f[2] = v[2];
        f[3] = fsub(v[3], v[0]); f[4] = fsub(v[4], v[1]); f[5] = fsub(v[5], v[2]);
        f[6] = fsub(v[6], v[0]); f[7] = fsub(v[7], v[1]); f[8] = fsub(v[8], v[2]);
        segment = base_index;
    }
#if ASGPU_DEVICE_CODE
    #pragma unroll
#endif
    for (int k = 0; k < 3; ++k)
    {
        tri.v0[k] = static_cast<double>(f[k]);
        tri.e0[k] = static_cast<double>(f[3 + k]);
        tri.e1[k] = static_cast<double>(f[6 + k]);
    }
    return true;
}

// The triangle the closest-hit visitor leaves behind for a hit (m_hit_triangle, triangletree.cpp:
// 1413, 1468-1469): the leaf's stored triangle, or for a moving triangle the one interpolated at
// the ray's normalized time (m_interpolated_triangle, a member: triangletree.h:232), widened to
// double by TriangleReader.  read_hit_triangle_data (:1483-1499) initialises the ShadingPoint's
// support plane from it.  `record` = the slot's TriRecord of the EXACT layout.
ASGPU_HD void hit_triangle(const uint8_t* record, const uint8_t* poses, const float time_normalized, TriD& tri)
{
    Ray r;
    r.flags = 0xFFFFFFFFu;
    r.time_normalized = time_normalized;
    uint32_t slot, segment;
    fetch_triangle<false>(record, poses, r, tri, slot, segment);
}

// ------------------------------------------------------------------------------------------
// IntersectionFilter::accept (intersectionfilter.h:169-205) with AlphaMask::is_opaque (:103-112):
// float arithmetic in the reference's order (Vector2f * float, sums left to right, clamp, truncate).
// `tree` points at the TreeDesc of a tree with filters; `slot` is the reference leaf slot.
// ------------------------------------------------------------------------------------------

ASGPU_HD uint64_t load_u64(const void* p)
{
    const uint2 w = load8(p);
    return static_cast<uint64_t>(w.x) | (static_cast<uint64_t>(w.y) << 32);
}

ASGPU_HD bool mask_is_opaque(const uint8_t* blob, const uint8_t* mask, const float ux, const float uy)
{
    const uint64_t bits = load_u64(mask);
    const uint2 dim = load8(mask + 8);
    const float max_x = fsub(static_cast<float>(dim.x), 1.0f), max_y = fsub(static_cast<float>(dim.y), 1.0f);
    float fx = fmul(ux, static_cast<float>(dim.x)), fy = fmul(uy, static_cast<float>(dim.y));
    fx = fx < 0.0f ? 0.0f : fx > max_x ? max_x : fx;
    fy = fy < 0.0f ? 0.0f : fy > max_y ? max_y : fy;
    const uint64_t ix = static_cast<uint64_t>(fx), iy = static_cast<uint64_t>(fy);
    const uint8_t byte = blob[bits + iy * ((dim.x + 7u) / 8u) + ix / 8u];
    return ((byte >> (ix & 7u)) & 1u) != 0;
}

ASGPU_HD bool filter_accept(const uint8_t* blob, const uint8_t* tree, const uint32_t slot, const double u, const double v)
{
    if (u != u || v != v) return true;
    const uint2 key = load8(blob + load_u64(tree + offsetof(TreeDesc, keys)) + static_cast<uint64_t>(slot) * sizeof(HitKey));
    if (key.x >= load4(tree + offsetof(TreeDesc, filter_count))) return true;
    const uint8_t* fr = blob + load_u64(tree + offsetof(TreeDesc, filters)) + static_cast<uint64_t>(key.x) * sizeof(FilterRecord);
    const uint64_t uv_off = load_u64(fr + offsetof(FilterRecord, uv));
    if (uv_off == 0) return true;                                           // no filter on this object instance
    const uint8_t* pap = blob + load_u64(tree + offsetof(TreeDesc, key_pa)) + static_cast<uint64_t>(slot) * 2;
    const uint32_t pa = static_cast<uint32_t>(pap[0]) | (static_cast<uint32_t>(pap[1]) << 8);
    const uint8_t* obj_mask = fr + offsetof(FilterRecord, object_mask);
    const bool has_obj = load_u64(obj_mask) != 0;
    const uint8_t* mtl_mask = nullptr;
    if (pa < load4(fr + offsetof(FilterRecord, material_mask_count)))
    {
        mtl_mask = blob + load_u64(fr + offsetof(FilterRecord, material_masks)) + static_cast<uint64_t>(pa) * sizeof(MaskRecord);
        if (load_u64(mtl_mask) == 0) mtl_mask = nullptr;
    }
    if (!has_obj && mtl_mask == nullptr) return true;
    const float fu = static_cast<float>(u), fv = static_cast<float>(v);
    const float w = fsub(fsub(1.0f, fu), fv);
    const uint8_t* t = blob + uv_off + static_cast<uint64_t>(key.y) * 24;
    float uv[2];
#if ASGPU_DEVICE_CODE
    #pragma unroll
#endif
    for (int c = 0; c < 2; ++c)
    {
        float x = fmul(u2f(load4(t + c * 4)), w);
        x = fadd(x, fmul(u2f(load4(t + 8 + c * 4)), fu));
        x = fadd(x, fmul(u2f(load4(t + 16 + c * 4)), fv));
        uv[c] = x;
    }
    if (has_obj && !mask_is_opaque(blob, obj_mask, uv[0], uv[1])) return false;
    if (mtl_mask != nullptr) return mask_is_opaque(blob, mtl_mask, uv[0], uv[1]);
    return true;
}

// ------------------------------------------------------------------------------------------
// EXACT traversal.
// ------------------------------------------------------------------------------------------

// minmax.h:171-180 -- identical selection to _mm_max_pd / _mm_min_pd (second operand on NaN).
ASGPU_HD double ssemax(double a, double b) { return a > b ? a : b; }
ASGPU_HD double ssemin(double a, double b) { return a < b ? a : b; }

struct RayInfoD { double rcp[3]; int sgn[3]; };

// RayInfo (ray.h:313-321).
ASGPU_HD void make_ray_info(const Ray& r, RayInfoD& info)
{
#if ASGPU_DEVICE_CODE
    #pragma unroll
#endif
    for (int a = 0; a < 3; ++a)
    {
        info.rcp[a] = ddiv(1.0, r.dir[a]);
        info.sgn[a] = info.rcp[a] >= 0.0 ? 1 : 0;
    }
}

// One child of an interior node; box[] = [minL minR maxL maxR] x axis (bvh_intersector.h:519-536).
ASGPU_HD bool slab_child(const double box[12], const int side, const Ray& ray, const RayInfoD& info,
                         const double ray_tmin, const double ray_tmax, double& tmin_out)
{
    double l1[3], l2[3];
#if ASGPU_DEVICE_CODE
    #pragma unroll
#endif
    for (int a = 0; a < 3; ++a)
    {
        const double near_plane = info.sgn[a] ? box[a * 4 + side] : box[a * 4 + 2 + side];
        const double far_plane  = info.sgn[a] ? box[a * 4 + 2 + side] : box[a * 4 + side];
        l1[a] = dmul(info.rcp[a], dsub(near_plane, ray.org[a]));
        l2[a] = dmul(info.rcp[a], dsub(far_plane, ray.org[a]));
    }
    const double tmin = ssemax(l1[2], ssemax(l1[1], ssemax(l1[0], ray_tmin)));
    const double tmax = ssemin(l2[2], ssemin(l2[1], ssemin(l2[0], ray_tmax)));
    tmin_out = tmin;
    return !(tmin > tmax || tmax < ray_tmin || tmin >= ray_tmax);
}

// Bottom level on one triangle tree.  `ray` is the instance-space ray; its tmax shrinks on hits
// (closest hit).  Returns true on the first hit for ANY.
template <bool ANY, bool COUNT>
ASGPU_HD bool exact_triangle_tree(const uint8_t* blob, const TreeDesc& td, const uint8_t* td_ptr, Ray& ray, Hit& hit, const uint32_t item, Stats& stats)
{
    RayInfoD info; make_ray_info(ray, info);
    const bool motion = td.moving > 0;
    const double ray_time = static_cast<double>(ray.time_normalized);
    const uint8_t* bnodes = blob + td.bnodes;
    const uint8_t* poses = blob + td.poses;

    uint32_t stack[64];
    int sp = 0;
    uint32_t node = 0;
    double rtmax = ray.tmax;

    while (true)
    {
        if (COUNT) ++stats.nodes;
        const uint8_t* np = bnodes + static_cast<uint64_t>(node) * sizeof(BNodeF);
        const uint4 w0 = load16(np);
        const uint32_t index = w0.x, item_count = w0.y;
        if (item_count == InteriorMark)
        {
            const uint4 w1 = load16(np + 16), w2 = load16(np + 32), w3 = load16(np + 48);
            double box[12];
            box[0] = u2f(w0.z); box[1] = u2f(w0.w);
            box[2] = u2f(w1.x); box[3] = u2f(w1.y); box[4] = u2f(w1.z); box[5] = u2f(w1.w);
            box[6] = u2f(w2.x); box[7] = u2f(w2.y); box[8] = u2f(w2.z); box[9] = u2f(w2.w);
            box[10] = u2f(w3.x); box[11] = u2f(w3.y);

            if (motion)
            {
                // Children with several motion boxes are interpolated at the ray time
                // (bvh_intersector.h:675-836): a * w1 + b * w2, w2 = t - trunc(t), w1 = 1 - w2.
                const uint4 mn = load16(blob + td.mnodes + static_cast<uint64_t>(node) * sizeof(MNode));
                const uint32_t idx[2] = { mn.x, mn.z }, cnt[2] = { mn.y, mn.w };
                for (int side = 0; side < 2; ++side)
                {
                    const uint32_t segments = cnt[side] - 1;
                    if (segments == 0) continue;
                    const double t = dmul(ray_time, static_cast<double>(segments));
                    const int prev = static_cast<int>(t);
                    const double w2 = dsub(t, static_cast<double>(prev));
                    const double w1 = dsub(1.0, w2);
                    const uint8_t* mb = blob + td.mboxes + (static_cast<uint64_t>(idx[side]) + prev) * sizeof(MBox);
                    for (int a = 0; a < 3; ++a)
                    {
                        const double lo0 = u2f(load4(mb + (a * 2) * 4)), hi0 = u2f(load4(mb + (a * 2 + 1) * 4));
                        const double lo1 = u2f(load4(mb + 24 + (a * 2) * 4)), hi1 = u2f(load4(mb + 24 + (a * 2 + 1) * 4));
                        box[a * 4 + side] = dadd(dmul(lo0, w1), dmul(lo1, w2));
                        box[a * 4 + 2 + side] = dadd(dmul(hi0, w1), dmul(hi1, w2));
                    }
                }
            }

            double tmin0, tmin1;
            const bool hit_left = slab_child(box, 0, ray, info, ray.tmin, rtmax, tmin0);
            const bool hit_right = slab_child(box, 1, ray, info, ray.tmin, rtmax, tmin1);
            if (hit_left != hit_right) { node = index + (hit_right ? 1 : 0); continue; }
            if (hit_left)
            {
                // Near child first; ties go right first (bvh_intersector.h:551-562).
                const int far_index = tmin0 < tmin1 ? 1 : 0;
                stack[sp++] = index + far_index;
                node = index + 1 - far_index;
                continue;
            }
            if (sp == 0) break;
            node = stack[--sp];
            continue;
        }

        // Leaf: TriangleLeafVisitor::visit / TriangleLeafProbeVisitor::visit.
        for (uint32_t j = 0; j < item_count; ++j)
        {
            if (COUNT) ++stats.triangles;
            TriD tri; uint32_t slot, segment;
            if (!fetch_triangle<ANY>(blob + td.tris + static_cast<uint64_t>(index + j) * sizeof(TriRecord), poses, ray, tri, slot, segment))
                continue;
            double t, u, v;
            if (mt_test<!ANY>(tri, ray, t, u, v))
            {
                if (ANY) return true;
                // Optionally filter intersections (triangletree.cpp:1404-1411; closest hit only).
                if (td.filter_count != 0 && !filter_accept(blob, td_ptr, slot, u, v)) continue;
                ray.tmax = t;
                hit.u = static_cast<float>(u);
                hit.v = static_cast<float>(v);
                hit.item = item;
                hit.slot = slot;
                hit.segment = segment;
            }
        }
        if (rtmax > ray.tmax) rtmax = ray.tmax;
        if (sp == 0) break;
        node = stack[--sp];
    }
    return false;
}

ASGPU_HD void load_tree_desc(const SceneView& s, const uint32_t tree, TreeDesc& td)
{
    const uint8_t* p = s.blob + s.trees + static_cast<uint64_t>(tree) * sizeof(TreeDesc);
    uint64_t* dst = reinterpret_cast<uint64_t*>(&td);
#if ASGPU_DEVICE_CODE
    #pragma unroll
#endif
    for (int k = 0; k < static_cast<int>(sizeof(TreeDesc) / 8); ++k)
    {
        const uint2 w = load8(p + k * 8);
        dst[k] = static_cast<uint64_t>(w.x) | (static_cast<uint64_t>(w.y) << 32);
    }
}

// Top level: generic scalar intersector + assembly leaf visitors.  On return ray.tmax is the hit
// distance (closest hit).  Returns true when something was hit.
template <bool ANY, bool COUNT>
ASGPU_HD bool exact_trace(const SceneView& s, Ray& ray, Hit& hit, Stats& stats, const uint8_t* parent = nullptr)
{
    hit.item = 0xFFFFFFFFu; hit.slot = 0; hit.segment = 0; hit.u = hit.v = 0.0f;
    RayInfoD info; make_ray_info(ray, info);
    const uint8_t* top = s.blob + s.top_nodes;

    uint32_t stack[64];
    int sp = 0;
    uint32_t node = 0;
    double ray_tmax = ray.tmax;

    while (true)
    {
        if (COUNT) ++stats.top_nodes;
        const uint8_t* np = top + static_cast<uint64_t>(node) * sizeof(BNodeD);
        const uint2 head = load8(np);
        const uint32_t item_count = head.x, index = head.y;
        if (item_count == InteriorMark)
        {
            double box[12];
#if ASGPU_DEVICE_CODE
            #pragma unroll
#endif
            for (int k = 0; k < 12; ++k) box[k] = load_f64(np + 32 + k * 8);
            // rayaabb.h:228-252 reads the LIVE ray.tmax; the loop also requires tmin < ray_tmax
            // (bvh_intersector.h:177-186).  tmin_out = ssemax(ray.tmin, tmin) == tmin.
            double tmin0, tmin1;
            const bool hit_left = slab_child(box, 0, ray, info, ray.tmin, ray.tmax, tmin0) && tmin0 < ray_tmax;
            const bool hit_right = slab_child(box, 1, ray, info, ray.tmin, ray.tmax, tmin1) && tmin1 < ray_tmax;
            if (hit_left != hit_right) { node = index + (hit_right ? 1 : 0); continue; }
            if (hit_left)
            {
                const int far_index = tmin0 < tmin1 ? 1 : 0;
                stack[sp++] = index + far_index;
                node = index + 1 - far_index;
                continue;
            }
            if (sp == 0) break;
            node = stack[--sp];
            continue;
        }

        for (uint32_t i = 0; i < item_count; ++i)
        {
            const uint32_t item = index + i;
            const uint8_t* ip = s.blob + s.items + static_cast<uint64_t>(item) * sizeof(ItemRecord);
            const uint4 meta = load16(ip + 96);     // tree, vis_flags, assembly_instance, pad
            if (!(meta.y & ray.flags)) continue;
            if (COUNT) ++stats.instances;
            Ray local;
            to_instance_space(s.blob, ip, meta.w, ray, local);
            parent_origin(parent, meta.z, local.dir, local.org);
            if (meta.x == 0xFFFFFFFFu) continue;
            TreeDesc td; load_tree_desc(s, meta.x, td);
            // The instance gets its own record (AssemblyLeafVisitor's asm_inst_shading_point) and is
            // committed only when it is strictly nearer (assemblytree.cpp:727-744): a triangle that
            // re-computes exactly the current t in a coincident instance must not replace the hit.
            Hit local_hit;
            local_hit.item = 0xFFFFFFFFu; local_hit.slot = 0; local_hit.segment = 0; local_hit.u = local_hit.v = 0.0f;
            const bool found = exact_triangle_tree<ANY, COUNT>(s.blob, td, s.blob + s.trees + static_cast<uint64_t>(meta.x) * sizeof(TreeDesc),
                                                               local, local_hit, item, stats);
            if (ANY) { if (found) return true; }
            else if (local_hit.item != 0xFFFFFFFFu && local.tmax < ray.tmax) { ray.tmax = local.tmax; hit = local_hit; }
        }
        if (ray_tmax > ray.tmax) ray_tmax = ray.tmax;
        if (sp == 0) break;
        node = stack[--sp];
    }
    return !ANY && hit.item != 0xFFFFFFFFu;
}

// ------------------------------------------------------------------------------------------
// WIDE traversal.
// ------------------------------------------------------------------------------------------

// Interval data for conservative fp32 box tests of one ray in one space.  The ray parameter is
// shifted by `shift` = min(tmin, 0) so that the tested interval starts at a non-negative value;
// with that, products of negative plane distances never matter and one multiply per plane is
// enough.  Per axis the frame is mirrored when the direction is negative, so the near plane is
// always the "lo" one in that frame:
//   d_near >= A_lo + q_near * s   (rounded down),   tau_near = d_near * rn   (rounded down)
//   d_far  <= A_hi + q_far  * s   (rounded up),     tau_far  = d_far  * rf   (rounded up)
// with [rn, rf] enclosing |1 / dir|.
struct WideRay
{
    float       o_lo[3], o_hi[3];   // origin interval (already mirrored per axis)
    float       rn[3], rf[3];       // |1 / dir| interval
    float       tmin_f, tmax_f;     // shifted parameter interval, rounded outward
    uint32_t    oct;                // bit a (0-2) set: direction negative along axis a; bit 3: some component is exactly 0
    double      shift;
};

ASGPU_HD void make_wide_ray(const double org[3], const double dir[3], const double tmin, const double tmax, WideRay& w)
{
    w.shift = tmin < 0.0 ? tmin : 0.0;
    w.oct = 0;
#if ASGPU_DEVICE_CODE
    #pragma unroll
#endif
    for (int a = 0; a < 3; ++a)
    {
        // Same sign convention as RayInfo (ray.h:313-321), which tests 1 / dir >= 0: negative for
        // dir < 0, -0.0 (1 / -0.0 = -inf) and NaN.
        const double d = dir[a];
        const bool neg = !(d > 0.0 || (d == 0.0 && !sign_bit(d)));
        if (neg) w.oct |= 1u << a;
        if (d == 0.0) w.oct |= 8u;
        // [rn, rf] encloses |1 / dir|: directed fp32 reciprocals of an fp32 enclosure of |dir|
        // (no fp64 division).  |dir| = 0 gives rn = rf = +inf, a NaN gives NaNs (dropped by the
        // min/max of the box test = "no constraint").
        const double mag = fabs(d);                 // also clears the sign of -0.0
#if ASGPU_DEVICE_CODE
        {
            // One conversion and one rcp.approx: |magf / mag - 1| <= 2^-24 for a normal magf and
            // rcp.approx.f32 is within 2^-23 (PTX ISA), so r * (1 -+ 2^-21), rounded outward,
            // brackets 1 / mag.  Outside the normal range the bounds are set by hand.
            const float magf = __double2float_rn(mag);
            float rn, rf;
            if (magf >= 1.17549435e-38f && magf <= 4.0e37f)
            {
                const float r = rcp_approx(magf);
                rn = __fmul_rd(r, 0.99999952316284f);
                rf = __fmul_ru(r, 1.00000047683716f);
            }
            else if (mag == 0.0) rn = rf = __int_as_float(0x7F800000);
            else if (magf > 4.0e37f) { rn = 0.0f; rf = 3.0e-38f; }         // 1 / mag < 2.5e-38
            else if (magf == magf) { rn = 8.0e37f; rf = __int_as_float(0x7F800000); }     // 1 / mag > 8.5e37
            else { rn = magf; rf = 3.0e-38f; }                              // NaN direction
            w.rn[a] = rn;
            w.rf[a] = rf;
        }
#else
        w.rn[a] = frcp_dn(d2f_up(mag));
        w.rf[a] = frcp_up(d2f_dn(mag));
#endif
        const double o = org[a];
        float lo, hi;
        if (w.shift != 0.0)
        {
            // Shifted origin o + shift * dir, widened by a bound on its fp64 rounding error.
            const double step = dmul(w.shift, d);
            const double moved = dadd(o, step);
            const double mag2 = dadd(o < 0.0 ? -o : o, step < 0.0 ? -step : step);
            const double err = dmul(mag2, 4.5e-16);
            lo = d2f_dn(dsub(moved, err));
            hi = d2f_up(dadd(moved, err));
        }
        else { lo = d2f_dn(o); hi = d2f_up(o); }
        // Mirrored frame: x' = -x, so the interval becomes [-hi, -lo].
        w.o_lo[a] = neg ? -hi : lo;
        w.o_hi[a] = neg ? -lo : hi;
    }
    w.tmin_f = d2f_dn(dsub(tmin, w.shift));
    w.tmax_f = d2f_up(dsub(tmax, w.shift));
}

ASGPU_HD void make_wide_ray(const Ray& r, WideRay& w) { make_wide_ray(r.org, r.dir, r.tmin, r.tmax, w); }

ASGPU_HD void shrink_wide_ray(const Ray& r, WideRay& w)
{
    w.tmax_f = d2f_up(dsub(r.tmax, w.shift));
}

// Byte k of `word` as the float 1 + byte * 2^-15: the byte dropped into the mantissa of 1.0f by
// ONE byte permute (no integer-to-float conversion, no subtraction).  The box test works with
// plane coefficients pre-scaled by 2^15 and offsets pre-reduced by them, see wide_node_test_form.
// `one` holds the bits of 1.0f; the kernels pass it through a kernel argument so that it stays in
// a register and the permute selector is the instruction's immediate.
const uint32_t UnitBits = 0x3F800000u;

ASGPU_HD float byte_to_unit(const uint32_t word, const int k, const uint32_t one)
{
#if ASGPU_DEVICE_CODE
    return __uint_as_float(__byte_perm(word, one, 0x7604 | (k << 4)));
#else
    return bits_to_float(one | (((word >> (8 * k)) & 0xFFu) << 8));
#endif
}

// Tests the 8 children of a wide node.  Returns (internal-hit bits << 24 | imask) in `nmask` (the
// internal-hit bits permuted so that the highest set bit is the child to visit first for this
// ray's octant) and the triangle-hit bits in `tmask`; fills child_base / tri_base.
//
// Per axis, in the mirrored frame, a child plane sits at p + q * s (exact) and the entry / exit
// parameters are bounded by
//     tau_near >= (a_lo + q_near * s) * rn,    tau_far <= (a_hi + q_far * s) * rf
// with a_lo <= p - o <= a_hi.  EXPANDED = false evaluates exactly that (one FMA and one multiply
// per plane, directed rounding).  EXPANDED = true distributes the reciprocal once per node,
//     tau_near >= q_near * (s * rn) + a_lo * rn,    tau_far <= q_far * (s * rf) + a_hi * rf,
// one FMA per plane; every rounding still goes outward, so both forms never miss a box the
// reference's fp64 slab test on the tighter binary box would enter.  The expanded form produces
// NaNs (= "no constraint", conservative but useless) when a reciprocal is infinite, so rays with
// an exactly zero direction component (oct bit 3) take the other form.
//
// The quantised byte q enters the FMA as m = 1 + q * 2^-15 (byte_to_unit): with S = s * 2^15
// (exact, a power of two), q * s + a = m * S + (a - S), so the offset is reduced by S once per
// node and axis (rounded outward; S is only 2^15 grid steps, so this costs 2^-8 of a grid step of
// tightness) and no per-plane conversion remains.
//
// `np` is the wide node (origin, exponents, imask, bases, meta in its first 32 bytes), `qp` its 48
// bytes of quantised child planes: the node's own (np + 32, bounding the whole motion) or the
// WSlice of the ray's time slice in a tree with moving triangles.
template <bool EXPANDED>
ASGPU_HD void wide_node_test_form(const uint8_t* np, const uint8_t* qp, const WideRay& w, const uint32_t one, uint32_t& child_base, uint32_t& tri_base, uint32_t& nmask, uint32_t& tmask)
{
    const uint4 n0 = load16(np), n1 = load16(np + 16), n2 = load16(qp), n3 = load16(qp + 16), n4 = load16(qp + 32);
    child_base = n1.x;
    tri_base = n1.y;
    const uint32_t imask = n0.w >> 24;
    const uint32_t meta_lo = n1.z, meta_hi = n1.w;
    // Quantised planes: qlo x (n2.x, n2.y) y (n2.z, n2.w) z (n3.x, n3.y); qhi x (n3.z, n3.w) y (n4.x, n4.y) z (n4.z, n4.w).
    const uint32_t qlo[3][2] = { { n2.x, n2.y }, { n2.z, n2.w }, { n3.x, n3.y } };
    const uint32_t qhi[3][2] = { { n3.z, n3.w }, { n4.x, n4.y }, { n4.z, n4.w } };
    const float origin[3] = { u2f(n0.x), u2f(n0.y), u2f(n0.z) };

    float cn[3], cf[3], bn[3], bf[3];       // coefficients and offsets of the near / far plane FMAs
    uint32_t qn[3][2], qf[3][2];
#if ASGPU_DEVICE_CODE
    #pragma unroll
#endif
    for (int a = 0; a < 3; ++a)
    {
        const bool neg = (w.oct >> a) & 1;
        // S = 2^(e + 15); the flattener keeps e + 15 below the exponent range.
        const float scale15 = u2f((((n0.w >> (8 * a)) & 0xFF) + 15u) << 23);
        const float p = neg ? -origin[a] : origin[a];       // node origin in the mirrored frame
        const float s15 = neg ? -scale15 : scale15;
        const float a_lo = fsub_dn(p, w.o_hi[a]);
        const float a_hi = fsub_up(p, w.o_lo[a]);
        if (EXPANDED)
        {
            cn[a] = fmul_dn(s15, w.rn[a]); bn[a] = fma_dn(a_lo, w.rn[a], -cn[a]);
            cf[a] = fmul_up(s15, w.rf[a]); bf[a] = fma_up(a_hi, w.rf[a], -cf[a]);
        }
        else { cn[a] = s15; cf[a] = s15; bn[a] = fsub_dn(a_lo, s15); bf[a] = fsub_up(a_hi, s15); }
        qn[a][0] = neg ? qhi[a][0] : qlo[a][0]; qn[a][1] = neg ? qhi[a][1] : qlo[a][1];
        qf[a][0] = neg ? qlo[a][0] : qhi[a][0]; qf[a][1] = neg ? qlo[a][1] : qhi[a][1];
    }

    uint32_t hitmask = 0;
#if ASGPU_DEVICE_CODE
    #pragma unroll
#endif
    for (int k = 0; k < 8; ++k)
    {
        float tn = w.tmin_f, tf = w.tmax_f;
#if ASGPU_DEVICE_CODE
        #pragma unroll
#endif
        for (int a = 0; a < 3; ++a)
        {
            const float mn = byte_to_unit(qn[a][k >> 2], k & 3, one);
            const float mf = byte_to_unit(qf[a][k >> 2], k & 3, one);
            if (EXPANDED)
            {
                tn = fmax_nan(tn, fma_dn(mn, cn[a], bn[a]));
                tf = fmin_nan(tf, fma_up(mf, cf[a], bf[a]));
            }
            else
            {
                tn = fmax_nan(tn, fmul_dn(fma_dn(mn, cn[a], bn[a]), w.rn[a]));
                tf = fmin_nan(tf, fmul_up(fma_up(mf, cf[a], bf[a]), w.rf[a]));
            }
        }
        // meta: 0 = empty; internal child 0x20 | (24 + k); leaf (unary count << 5) | first slot.
        const uint32_t meta = ((k < 4 ? meta_lo : meta_hi) >> (8 * (k & 3))) & 0xFF;
        const uint32_t bits = (meta >> 5) << (meta & 31);
        hitmask |= tn <= tf ? bits : 0u;
    }

    // Internal hits sit at bit 24 + k; move child k to bit 24 + (k ^ (7 - octant)).
    uint32_t x = hitmask >> 24;
    const uint32_t oct_inv = 7 - (w.oct & 7);
    if (oct_inv & 1) x = ((x & 0xAAu) >> 1) | ((x & 0x55u) << 1);
    if (oct_inv & 2) x = ((x & 0xCCu) >> 2) | ((x & 0x33u) << 2);
    if (oct_inv & 4) x = ((x & 0xF0u) >> 4) | ((x & 0x0Fu) << 4);
    nmask = (x << 24) | imask;
    tmask = hitmask & 0x00FFFFFFu;
}

ASGPU_HD void wide_node_test(const uint8_t* np, const uint8_t* qp, const WideRay& w, const uint32_t one, uint32_t& child_base, uint32_t& tri_base, uint32_t& nmask, uint32_t& tmask)
{
    if (w.oct & 8) wide_node_test_form<false>(np, qp, w, one, child_base, tri_base, nmask, tmask);
    else wide_node_test_form<true>(np, qp, w, one, child_base, tri_base, nmask, tmask);
}

// Time slice of a ray in a tree with `slices` slices: floor(time * slices), clamped.  The slices
// overlap by 1e-3 of their width (flatten.cpp), far more than the rounding of this product.
ASGPU_HD uint32_t time_slice(const float time_normalized, const uint32_t slices)
{
    const float x = fmul(time_normalized, static_cast<float>(slices));
    uint32_t j = x > 0.0f ? static_cast<uint32_t>(x) : 0u;
    return j < slices ? j : slices - 1;
}

const uint32_t WideStackSize = WideStackMax;       // host driver; the kernels pick a depth per scene

ASGPU_HD void load_ray_org_dir(const asgpu_rays& rays, const size_t i, Ray& r)
{
#if ASGPU_DEVICE_CODE
    r.org[0] = __ldg(rays.org + i * 3); r.org[1] = __ldg(rays.org + i * 3 + 1); r.org[2] = __ldg(rays.org + i * 3 + 2);
    r.dir[0] = __ldg(rays.dir + i * 3); r.dir[1] = __ldg(rays.dir + i * 3 + 1); r.dir[2] = __ldg(rays.dir + i * 3 + 2);
#else
    for (int a = 0; a < 3; ++a) { r.org[a] = rays.org[i * 3 + a]; r.dir[a] = rays.dir[i * 3 + a]; }
#endif
}

// The wide traversal as an explicit per-ray state machine: begin() loads a ray, step() performs one
// iteration (fetch + test one wide node, intersect its pending leaf items, choose what comes next)
// and returns true when the ray is finished.  The kernel interleaves step() with a warp-level
// refill of finished lanes from the ray queue, so a warp keeps most of its lanes busy even when
// ray lengths differ wildly.
//
// Only the CURRENT-space origin and direction are kept in registers: in world space they are the
// ray's own, inside an instance they are the transformed ones, and the world ones are re-read from
// the ray arrays when the traversal returns to world space.  tmin / tmax / time / flags are shared
// between spaces because the instance direction is not renormalised (assemblytree.cpp:556-596).
//
// `stack` holds WideStackSize entries per thread, element i of this thread at stack[i * stride].
template <bool ANY, bool COUNT>
struct WideTraversal
{
    Ray             ray;
    WideRay         wr;
    const uint8_t*  wnodes;
    const uint8_t*  qbase;              // child planes of node i at qbase + i * qstride
    uint32_t        qstride;
    const uint8_t*  wtris;
    const uint8_t*  poses;
    const uint8_t*  filter_tree;        // TreeDesc of the current tree when it has intersection filters
    const uint8_t*  parent;             // this ray's asgpu_parent record or null (set after begin())
    uint2           ngroup, tgroup;
    uint32_t        fetch;              // wide node to fetch next, 0xFFFFFFFF = none
    uint32_t        sp;
    uint32_t        cur_item;           // 0xFFFFFFFF while in world space
    Hit             hit;

    ASGPU_HD void begin(const SceneView& s, const asgpu_rays& rays, const size_t index)
    {
        load_ray(rays, index, ray);
        make_wide_ray(ray, wr);
        wnodes = s.blob + s.top_wnodes;
        qbase = wnodes + 32; qstride = sizeof(WNode);
        wtris = nullptr;
        poses = nullptr;
        filter_tree = nullptr;
        parent = nullptr;
        ngroup.x = 0; ngroup.y = 0;
        tgroup.x = 0; tgroup.y = 0;
        fetch = s.top_wnode_count != 0 ? 0u : 0xFFFFFFFFu;
        sp = 0;
        cur_item = 0xFFFFFFFFu;
        hit.item = 0xFFFFFFFFu; hit.slot = 0; hit.segment = 0; hit.u = hit.v = 0.0f;
    }

    ASGPU_HD bool found() const { return hit.item != 0xFFFFFFFFu; }

    // Returns true when the traversal of this ray is complete.
    ASGPU_HD bool step(const SceneView& s, const asgpu_rays& rays, const size_t index, Stats& stats, uint2* stack, const uint32_t stride)
    {
        const bool in_instance = cur_item != 0xFFFFFFFFu;

        if (fetch != 0xFFFFFFFFu)
        {
            if (COUNT) { if (in_instance) ++stats.nodes; else ++stats.top_nodes; }
            if (ngroup.y & 0xFF000000u) { stack[sp * stride] = ngroup; ++sp; }
            uint32_t child_base, tri_base, nmask, tmask;
            wide_node_test(wnodes + static_cast<uint64_t>(fetch) * sizeof(WNode), qbase + static_cast<uint64_t>(fetch) * qstride, wr, UnitBits, child_base, tri_base, nmask, tmask);
            ngroup.x = child_base; ngroup.y = nmask;
            tgroup.x = tri_base; tgroup.y = tmask;
            fetch = 0xFFFFFFFFu;
        }

        // Leaf items of the current node.
        while (tgroup.y)
        {
            const int bit = high_bit(tgroup.y);
            tgroup.y &= ~(1u << bit);
            if (in_instance)
            {
                if (COUNT) ++stats.triangles;
                TriD tri; uint32_t slot, segment;
                if (!fetch_triangle<ANY>(wtris + static_cast<uint64_t>(tgroup.x + bit) * sizeof(TriRecord), poses, ray, tri, slot, segment))
                    continue;
                double t, u, v;
                if (mt_test<!ANY>(tri, ray, t, u, v))
                {
                    if (!ANY && filter_tree != nullptr && !filter_accept(s.blob, filter_tree, slot, u, v)) continue;
                    hit.item = cur_item;
                    if (ANY) return true;
                    ray.tmax = t;
                    shrink_wide_ray(ray, wr);
                    hit.u = static_cast<float>(u);
                    hit.v = static_cast<float>(v);
                    hit.slot = slot;
                    hit.segment = segment;
                }
            }
            else
            {
                // Assembly instance: AssemblyLeafVisitor::visit (assemblytree.cpp:604-744).
                const uint32_t item = load4(s.blob + s.top_witems + static_cast<uint64_t>(tgroup.x + bit) * 4);
                const uint8_t* ip = s.blob + s.items + static_cast<uint64_t>(item) * sizeof(ItemRecord);
                const uint4 meta = load16(ip + 96);
                if (!(meta.y & ray.flags) || meta.x == 0xFFFFFFFFu) continue;
                if (COUNT) ++stats.instances;
                // Save the world-space traversal state, then descend.
                if (ngroup.y & 0xFF000000u) { stack[sp * stride] = ngroup; ++sp; }
                if (tgroup.y) { stack[sp * stride] = tgroup; ++sp; }
                uint2 sentinel; sentinel.x = 0xFFFFFFFFu; sentinel.y = 0;
                stack[sp * stride] = sentinel; ++sp;
                Ray local;
                to_instance_space(s.blob, ip, meta.w, ray, local);
                parent_origin(parent, meta.z, local.dir, local.org);
                ray = local;
                make_wide_ray(ray, wr);
                TreeDesc td; load_tree_desc(s, meta.x, td);
                wnodes = s.blob + td.wnodes;
                if (td.wslice_count != 0)
                {
                    qbase = s.blob + td.wslices + static_cast<uint64_t>(time_slice(ray.time_normalized, td.wslice_count)) * sizeof(WSlice);
                    qstride = td.wslice_count * static_cast<uint32_t>(sizeof(WSlice));
                }
                else { qbase = wnodes + 32; qstride = sizeof(WNode); }
                wtris = s.blob + td.wtris;
                poses = s.blob + td.poses;
                filter_tree = td.filter_count != 0 ? s.blob + s.trees + static_cast<uint64_t>(meta.x) * sizeof(TreeDesc) : nullptr;
                cur_item = item;
                ngroup.y = 0; tgroup.y = 0;
                if (td.wnode_count != 0) fetch = 0;
                return false;
            }
        }

        // Next internal child of the current group, nearest octant slot first.
        if (ngroup.y & 0xFF000000u)
        {
            const int bit = high_bit(ngroup.y);
            ngroup.y &= ~(1u << bit);
            const uint32_t k = static_cast<uint32_t>(bit - 24) ^ (7 - (wr.oct & 7));
            fetch = ngroup.x + popc(ngroup.y & 0xFFu & ((1u << k) - 1u));
            return false;
        }

        if (sp == 0) return true;
        --sp;
        const uint2 top = stack[sp * stride];
        if (top.x == 0xFFFFFFFFu && top.y == 0)
        {
            // Back to world space: the world origin and direction come from the ray arrays again.
            load_ray_org_dir(rays, index, ray);
            make_wide_ray(ray, wr);
            wnodes = s.blob + s.top_wnodes;
            qbase = wnodes + 32; qstride = sizeof(WNode);
            cur_item = 0xFFFFFFFFu;
            ngroup.y = 0; tgroup.y = 0;
            return false;
        }
        if (top.y & 0xFF000000u) { ngroup = top; tgroup.y = 0; }
        else { tgroup = top; ngroup.y = 0; }
        return false;
    }
};

// One ray start to finish (host simulation; the kernel drives WideTraversal itself).
template <bool ANY, bool COUNT>
ASGPU_HD bool wide_trace(const SceneView& s, const asgpu_rays& rays, const size_t index, Ray& out_ray, Hit& hit, Stats& stats, uint2* stack, const uint32_t stride,
                         const uint8_t* parent = nullptr)
{
    WideTraversal<ANY, COUNT> tr;
    tr.begin(s, rays, index);
    tr.parent = parent;
    while (!tr.step(s, rays, index, stats, stack, stride)) {}
    out_ray = tr.ray;
    hit = tr.hit;
    return tr.found();
}

}   // namespace asgpu
