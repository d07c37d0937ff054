//
// refine.cu -- ShadingPoint::refine_and_offset on the device.
//
// For every closest-hit record the kernel produces what the reference keeps in the ShadingPoint
// for its child rays (shadingpoint.cpp:362-425, triangle branch, RENDERER_ADAPTIVE_OFFSET); the
// per-hit arithmetic is refine_core.h.  One thread per hit; the work is a few hundred dependent
// fp64 operations and five dependent loads per hit, i.e. latency bound and small next to the
// trace that produced the hits (6 % of a C5 frame).
//

#include "kernels.h"
#include "refine_core.h"

#include <cuda_runtime.h>

namespace asgpu
{

namespace
{

const int RefineThreads = 128;

// 8 resident CTAs (64 registers, a few spills): the kernel waits on dependent loads, more warps in
// flight beat fewer spills (2.19 -> 1.86 ms per 8-spp C5 frame, profiles/README.md).
__global__ void __launch_bounds__(RefineThreads, 8)
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
        Ray world;
        load_ray_org_dir(rays, i, world);
        const float time_absolute = rays.time_absolute ? __ldg(rays.time_absolute + i) : 0.0f;
        const float time_normalized = rays.time_normalized ? __ldg(rays.time_normalized + i) : 0.0f;
        refine_offset_one(s, world.org, world.dir, time_absolute, time_normalized, t, item, static_cast<uint32_t>(w2 >> 32), static_cast<uint32_t>(w3), static_cast<uint32_t>(w3 >> 32), dst);
    }
}

// ShadingPoint::m_triangle_support_plane of every hit (read_hit_triangle_data, triangletree.cpp:
// 1483-1499): nine doubles v0, e0, e1; zeros for a miss.  72 B written and one 48-byte record (+ two
// 36-byte poses for a moving triangle) read per hit: a streaming kernel.
__global__ void __launch_bounds__(RefineThreads)
support_plane_kernel(const SceneView s, const asgpu_rays rays, const asgpu_hit* __restrict__ hits, const unsigned long long n,
                     const bool raw_item, const uint32_t* __restrict__ id_to_item, const uint32_t id_count, double* __restrict__ planes)
{
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    {
        const unsigned long long* hw = reinterpret_cast<const unsigned long long*>(hits + i);
        const unsigned long long w2 = hw[2], w3 = hw[3], w4 = hw[4];
        double* dst = planes + i * 9;
        uint32_t item = static_cast<uint32_t>(w2);
        bool valid = static_cast<uint32_t>(w4 >> 32) == 2;
        if (valid && !raw_item)
        {
            valid = item < id_count;
            if (valid) { item = id_to_item[item]; valid = item != 0xFFFFFFFFu; }
        }
        TriD tri;
        #pragma unroll
        for (int k = 0; k < 3; ++k) tri.v0[k] = tri.e0[k] = tri.e1[k] = 0.0;
        if (valid)
        {
            const uint8_t* ip = s.blob + s.items + static_cast<uint64_t>(item) * sizeof(ItemRecord);
            const uint32_t tree = load4(ip + 96);
            const uint8_t* tp = s.blob + s.trees + static_cast<uint64_t>(tree) * sizeof(TreeDesc);
            const uint64_t tris = load_u64(tp + offsetof(TreeDesc, tris)), poses = load_u64(tp + offsetof(TreeDesc, poses));
            const float time_normalized = rays.time_normalized ? __ldg(rays.time_normalized + i) : 0.0f;
            hit_triangle(s.blob + tris + (w3 >> 32) * sizeof(TriRecord), s.blob + poses, time_normalized, tri);
        }
        #pragma unroll
        for (int k = 0; k < 3; ++k) { dst[k] = tri.v0[k]; dst[3 + k] = tri.e0[k]; dst[6 + k] = tri.e1[k]; }
    }
}

}   // anonymous namespace

int launch_support_planes(const SceneView& scene, const asgpu_rays& rays, const asgpu_hit* hits, const size_t n, const bool raw_item,
                          const uint32_t* id_to_item, const uint32_t id_count, double* planes, const int sm_count, void* stream)
{
    if (n == 0) return 0;
    long long grid = static_cast<long long>(sm_count) * 8;
    const long long needed = static_cast<long long>((n + RefineThreads - 1) / RefineThreads);
    if (grid > needed) grid = needed;
    support_plane_kernel<<<static_cast<unsigned>(grid), RefineThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        scene, rays, hits, n, raw_item, id_to_item, id_count, planes);
    return static_cast<int>(cudaGetLastError());
}

int launch_refine_offset(const SceneView& scene, const asgpu_rays& rays, const asgpu_hit* hits, const size_t n, const unsigned long long* n_dev,
                         const bool raw_item, const uint32_t* id_to_item, const uint32_t id_count, asgpu_parent* parents,
                         const int sm_count, void* stream)
{
    if (n == 0) return 0;
    long long grid = static_cast<long long>(sm_count) * 16;
    const long long needed = static_cast<long long>((n + RefineThreads - 1) / RefineThreads);
    if (grid > needed) grid = needed;
    refine_offset_kernel<<<static_cast<unsigned>(grid), RefineThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        scene, rays, hits, n, n_dev, raw_item, id_to_item, id_count, parents);
    return static_cast<int>(cudaGetLastError());
}

}   // namespace asgpu
