//
// refine.cu -- ShadingPoint::refine_and_offset on the device.
//
// For every closest-hit record the kernel produces what the reference keeps in the ShadingPoint
// for its child rays (shadingpoint.cpp:362-425, triangle branch, RENDERER_ADAPTIVE_OFFSET):
//
//   refine_space_ray = assembly_instance_transform.to_local(ray);  org += tmax * dir
//   org = refine(org, dir, plane)                       two Newton steps onto the triangle's
//                                                       support plane (refining.h:97-113,
//                                                       raytrianglemt.h:300-309)
//   n   = faceforward(object_instance.normal_to_parent(cross(v1 - v0, v2 - v0)), dir)
//                                                       source vertices, float cross product
//                                                       (renderer/utility/triangle.h:57-64)
//   front / back = adaptive_offset(org, normalize(n))   ulp steps of doubling size until the point
//                                                       is off the plane (refining.h:168-221)
//
// Every fp64 operation is an explicit round-to-nearest intrinsic in the reference's order (no FMA
// contraction): the records are bit-identical to the CPU oracle's.  One thread per hit; the work
// is a few hundred dependent fp64 operations and five dependent loads per hit, i.e. latency bound
// and tiny next to the trace that produced the hits.
//

#include "kernels.h"
#include "traverse_core.h"

#include <cuda_runtime.h>

namespace asgpu
{

namespace
{

const int RefineThreads = 128;

__device__ __forceinline__ double plane_intersect(const TriD& tri, const double org[3], const double dir[3])
{
    const double tvec[3] = { dsub(org[0], tri.v0[0]), dsub(org[1], tri.v0[1]), dsub(org[2], tri.v0[2]) };
    double qvec[3], pvec[3];
    cross_d(tvec, tri.e0, qvec);
    cross_d(dir, tri.e1, pvec);
    return ddiv(dot_d(tri.e1, qvec), dot_d(tri.e0, pvec));
}

// adaptive_offset_point_step (refining.h:197-221).
__device__ __forceinline__ void offset_step(double p[3], const double n[3], const long long mag)
{
    const double Threshold = 1.0e-25;
    #pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        if (fabs(p[i]) < Threshold) p[i] = dadd(p[i], dmul(n[i], Threshold));
        else
        {
            const unsigned long long pi = static_cast<unsigned long long>(__double_as_longlong(p[i]));
            const unsigned long long ni = static_cast<unsigned long long>(__double_as_longlong(n[i]));
            const long long step = ((pi ^ ni) >> 63) ? -mag : mag;
            p[i] = __longlong_as_double(static_cast<long long>(pi + static_cast<unsigned long long>(step)));
        }
    }
}

// adaptive_offset_point (refining.h:176-195).
__device__ __forceinline__ void offset_point(const TriD& tri, const double p[3], const double n[3], double out[3])
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

__global__ void __launch_bounds__(RefineThreads)
refine_offset_kernel(const SceneView s, const asgpu_rays rays, const asgpu_hit* __restrict__ hits, const unsigned long long n_host,
                     const unsigned long long* n_dev, const bool raw_item, const uint32_t* __restrict__ id_to_item, const uint32_t id_count,
                     asgpu_parent* out)
{
    const unsigned long long n = n_dev ? min(*n_dev, n_host) : n_host;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    {
        const unsigned long long* hw = reinterpret_cast<const unsigned long long*>(hits + i);
        const unsigned long long w0 = hw[0], w2 = hw[2], w3 = hw[3], w4 = hw[4];
        double* dst = reinterpret_cast<double*>(out + i);
        uint32_t item = static_cast<uint32_t>(w2);
        bool valid = static_cast<uint32_t>(w4 >> 32) == 2;
        if (valid && !raw_item)
        {
            valid = item < id_count;
            if (valid) { item = id_to_item[item]; valid = item != 0xFFFFFFFFu; }
        }
        if (!valid)
        {
            reinterpret_cast<unsigned long long*>(dst)[0] = 0xFFFFFFFFull;
            #pragma unroll
            for (int k = 1; k < 10; ++k) dst[k] = 0.0;
            continue;
        }
        const double t = __longlong_as_double(static_cast<long long>(w0));
        const uint32_t object_instance = static_cast<uint32_t>(w2 >> 32);
        const uint32_t primitive = static_cast<uint32_t>(w3), slot = static_cast<uint32_t>(w3 >> 32);

        // refine_space_ray = to_local(ray), moved to the hit point.
        const uint8_t* ip = s.blob + s.items + static_cast<uint64_t>(item) * sizeof(ItemRecord);
        const uint4 meta = load16(ip + 96);
        Ray world;
        load_ray_org_dir(rays, i, world);
        double p[3], dir[3];
        instance_org_dir(ip, world.org, world.dir, p, dir);
        #pragma unroll
        for (int k = 0; k < 3; ++k) p[k] = dadd(p[k], dmul(dir[k], t));

        // Support plane = the triangle the leaf stores, widened to double.
        TreeDesc td; load_tree_desc(s, meta.x, td);
        TriD tri;
        {
            const uint8_t* rec = s.blob + td.tris + static_cast<uint64_t>(slot) * sizeof(TriRecord);
            const uint4 a = load16(rec), b = load16(rec + 16), c = load16(rec + 32);
            tri.v0[0] = u2f(a.x); tri.v0[1] = u2f(a.y); tri.v0[2] = u2f(a.z);
            tri.e0[0] = u2f(a.w); tri.e0[1] = u2f(b.x); tri.e0[2] = u2f(b.y);
            tri.e1[0] = u2f(b.z); tri.e1[1] = u2f(b.w); tri.e1[2] = u2f(c.x);
        }
        #pragma unroll
        for (int step = 0; step < 2; ++step)
        {
            const double tt = plane_intersect(tri, p, dir);
            #pragma unroll
            for (int k = 0; k < 3; ++k) p[k] = dadd(p[k], dmul(dir[k], tt));
        }

        // Geometric normal from the source vertices.
        double nrm[3];
        {
            const uint8_t* so = s.blob + td.src_objects + static_cast<uint64_t>(object_instance) * sizeof(SrcObject);
            const uint2 ov = load8(so + offsetof(SrcObject, vertices)), ot = load8(so + offsetof(SrcObject, triangles));
            const uint8_t* verts = s.blob + (static_cast<uint64_t>(ov.x) | (static_cast<uint64_t>(ov.y) << 32));
            const uint8_t* tris = s.blob + (static_cast<uint64_t>(ot.x) | (static_cast<uint64_t>(ot.y) << 32)) + static_cast<uint64_t>(primitive) * 12;
            const uint32_t i0 = load4(tris), i1 = load4(tris + 4), i2 = load4(tris + 8);
            float v[3][3];
            const uint32_t idx[3] = { i0, i1, i2 };
            #pragma unroll
            for (int c = 0; c < 3; ++c)
                #pragma unroll
                for (int k = 0; k < 3; ++k) v[c][k] = u2f(load4(verts + static_cast<uint64_t>(idx[c]) * 12 + k * 4));
            const float a[3] = { fsub(v[1][0], v[0][0]), fsub(v[1][1], v[0][1]), fsub(v[1][2], v[0][2]) };
            const float b[3] = { fsub(v[2][0], v[0][0]), fsub(v[2][1], v[0][1]), fsub(v[2][2], v[0][2]) };
            const double nf[3] = {
                static_cast<double>(fsub(fmul(a[1], b[2]), fmul(b[1], a[2]))),
                static_cast<double>(fsub(fmul(a[2], b[0]), fmul(b[2], a[0]))),
                static_cast<double>(fsub(fmul(a[0], b[1]), fmul(b[0], a[1]))) };
            // normal_to_parent: column k of parent_to_local's 3 x 3 block (transform.h:446-463).
            #pragma unroll
            for (int k = 0; k < 3; ++k)
                nrm[k] = dadd(dadd(dmul(load_f64(so + k * 8), nf[0]), dmul(load_f64(so + (3 + k) * 8), nf[1])), dmul(load_f64(so + (6 + k) * 8), nf[2]));
            if (!(dot_d(nrm, dir) < 0.0)) { nrm[0] = -nrm[0]; nrm[1] = -nrm[1]; nrm[2] = -nrm[2]; }
        }

        // adaptive_offset: n = normalize(n) = n * (1 / norm) (vector.h:638-642, 768-773).
        const double rcp = ddiv(1.0, __dsqrt_rn(dot_d(nrm, nrm)));
        const double un[3] = { dmul(nrm[0], rcp), dmul(nrm[1], rcp), dmul(nrm[2], rcp) };
        const double mn[3] = { -un[0], -un[1], -un[2] };
        double front[3], back[3];
        offset_point(tri, p, un, front);
        offset_point(tri, p, mn, back);

        reinterpret_cast<unsigned long long*>(dst)[0] = static_cast<unsigned long long>(meta.z);
        #pragma unroll
        for (int k = 0; k < 3; ++k) { dst[1 + k] = front[k]; dst[4 + k] = back[k]; dst[7 + k] = nrm[k]; }
    }
}

}   // anonymous namespace

int launch_refine_offset(const SceneView& scene, const asgpu_rays& rays, const asgpu_hit* hits, const size_t n, const unsigned long long* n_dev,
                         const bool raw_item, const uint32_t* id_to_item, const uint32_t id_count, asgpu_parent* parents,
                         const int sm_count, void* stream)
{
    if (n == 0) return 0;
    long long grid = static_cast<long long>(sm_count) * 8;
    const long long needed = static_cast<long long>((n + RefineThreads - 1) / RefineThreads);
    if (grid > needed) grid = needed;
    refine_offset_kernel<<<static_cast<unsigned>(grid), RefineThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        scene, rays, hits, n, n_dev, raw_item, id_to_item, id_count, parents);
    return static_cast<int>(cudaGetLastError());
}

}   // namespace asgpu
