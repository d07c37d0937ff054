//
// wavefront.cu -- device-resident ray queues and the wavefront path stream built on them
// (include/asgpu.h, "Wavefront ray queues").
//
// The reference renders a sample by recursion on one thread: GenericSampleRenderer::render_sample
// (renderer/kernel/rendering/generic/genericsamplerenderer.cpp:164-299) spawns a camera ray,
// Intersector::trace finds the first vertex, the path tracer (renderer/kernel/lighting/
// pathtracer.h) samples a bounce and calls trace again, and at every vertex the direct-lighting
// code asks Tracer::trace_between (renderer/kernel/lighting/tracer.h:252-259) for the visibility
// of a light sample.  Here the same work is a sequence of STAGES over millions of rays:
//
//     generate_kernel           camera rays of a batch of tiles           -> closest-hit queue A
//     for every depth:
//       wide_kernel (closest)   trace queue A                             -> hit records
//       shade_kernel            hit point, geometric normal, next origin; -> shadow-probe queue P
//                               one light sample + one cosine bounce      -> closest-hit queue B
//       wide_kernel (any hit)   trace queue P                             -> occlusion flags
//       accumulate_kernel       per-pixel accumulators
//       swap(A, B)
//
// Queues live in HBM (SoA, the arrays of asgpu_rays + a path id per ray) and carry their ray count
// in device memory: the trace kernels read it there, so no stage waits for the host.  Rays are
// appended with one atomicAdd per warp (ballot + popc).  Every random number is a counter-based
// hash of (seed, pixel, sample, depth, dimension): images are identical whatever the batching,
// the order of rays in a queue or the number of GPUs that shared the tiles.
//
// These stage kernels are streaming (HBM-bandwidth) kernels; the trace kernels of kernels.cu stay
// the hot spot (about 80 % of a frame's device time; shade 10 %, refine 6 %, profiles/README.md).
//

#include "api_internal.h"
#include "kernels.h"
#include "refine_core.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

using namespace asgpu;

// ------------------------------------------------------------------------------------------
// Handles.
// ------------------------------------------------------------------------------------------

struct asgpu_ray_queue
{
    asgpu_scene*        scene = nullptr;
    size_t              capacity = 0;
    double*             org = nullptr;
    double*             dir = nullptr;
    double*             tmin = nullptr;
    double*             tmax = nullptr;
    float*              time_absolute = nullptr;    // ShadingRay::Time of every ray (shadingray.h:68-85)
    float*              time_normalized = nullptr;
    uint32_t*           flags = nullptr;
    uint32_t*           path = nullptr;
    unsigned long long* count = nullptr;
    asgpu_parent*       parents = nullptr;      // only for the queues of a path stream with ASGPU_STREAM_PARENTS
};

namespace
{

struct QueueView
{
    asgpu_parent*       parents;
    double*             org;
    double*             dir;
    double*             tmin;
    double*             tmax;
    float*              time_absolute;
    float*              time_normalized;
    uint32_t*           flags;
    uint32_t*           path;
    unsigned long long* count;
    unsigned long long  capacity;
};

QueueView view_of(const asgpu_ray_queue* q)
{
    QueueView v;
    v.org = q->org; v.dir = q->dir; v.tmin = q->tmin; v.tmax = q->tmax;
    v.time_absolute = q->time_absolute; v.time_normalized = q->time_normalized;
    v.flags = q->flags; v.path = q->path; v.count = q->count; v.capacity = q->capacity; v.parents = q->parents;
    return v;
}

asgpu_rays rays_of(const asgpu_ray_queue* q)
{
    asgpu_rays r;
    r.org = q->org; r.dir = q->dir; r.tmin = q->tmin; r.tmax = q->tmax;
    r.time_absolute = q->time_absolute; r.time_normalized = q->time_normalized; r.flags = q->flags;
    return r;
}

struct StreamParams
{
    uint32_t    width, height, spp, max_bounces, tile_size, tiles_x, light_count, pad;
    uint64_t    seed;
    double      cam[12];
    double      film_w, film_h, focal;
    double      lights[8][3];
    double      eps;
    float       shutter_open, shutter_close;
};

enum { StatCamera = 0, StatBounce, StatProbe, StatHits, StatEscaped, StatUnoccluded, StatCount };

const int StageThreads = 256;

// ------------------------------------------------------------------------------------------
// Device helpers.
// ------------------------------------------------------------------------------------------

// Counter-based random numbers: two rounds of the splitmix64 finaliser over
// (seed, path, depth, dimension).
__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

__device__ __forceinline__ double rng(const uint64_t seed, const uint32_t path, const uint32_t depth, const uint32_t dim)
{
    const uint64_t key = (static_cast<uint64_t>(path) << 16) | (static_cast<uint64_t>(depth) << 8) | dim;
    const uint64_t x = mix64(mix64(seed ^ 0x9E3779B97F4A7C15ull) + key * 0xD1342543DE82EF95ull);
    return static_cast<double>(x >> 11) * (1.0 / 9007199254740992.0);      // [0, 1)
}

// Appends one ray per emitting lane: one atomicAdd per warp.  Returns the slot or ~0 when the
// queue is full (cannot happen for queues sized by asgpu_path_stream_create).
__device__ __forceinline__ unsigned long long enqueue_slot(const QueueView& q, const bool emit)
{
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, emit);
    if (mask == 0) return ~0ull;
    const unsigned lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == static_cast<unsigned>(leader)) base = atomicAdd(q.count, static_cast<unsigned long long>(__popc(mask)));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    const unsigned long long slot = base + __popc(mask & ((1u << lane) - 1u));
    return emit && slot < q.capacity ? slot : ~0ull;
}

__device__ __forceinline__ void write_ray(const QueueView& q, const unsigned long long slot, const double o[3], const double d[3],
                                          const double tmin, const double tmax, const float time_absolute, const float time_normalized,
                                          const uint32_t flags, const uint32_t path)
{
    q.time_absolute[slot] = time_absolute;
    q.time_normalized[slot] = time_normalized;
    q.org[slot * 3] = o[0]; q.org[slot * 3 + 1] = o[1]; q.org[slot * 3 + 2] = o[2];
    q.dir[slot * 3] = d[0]; q.dir[slot * 3 + 1] = d[1]; q.dir[slot * 3 + 2] = d[2];
    q.tmin[slot] = tmin;
    q.tmax[slot] = tmax;
    q.flags[slot] = flags;
    q.path[slot] = path;
}

// Copies one 80-byte parent record (or marks "no parent").
__device__ __forceinline__ void write_parent(const QueueView& q, const unsigned long long slot, const asgpu_parent* src)
{
    if (q.parents == nullptr) return;
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(q.parents + slot);
    if (src == nullptr) { dst[0] = 0xFFFFFFFFull; return; }
    const unsigned long long* s8 = reinterpret_cast<const unsigned long long*>(src);
    #pragma unroll
    for (int k = 0; k < 10; ++k) dst[k] = s8[k];
}

__device__ __forceinline__ void add_stats(unsigned long long* stats, const unsigned (&local)[StatCount])
{
    const unsigned lane = threadIdx.x & 31;
    #pragma unroll
    for (int k = 0; k < StatCount; ++k)
    {
        unsigned v = local[k];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (lane == 0 && v) atomicAdd(stats + k, static_cast<unsigned long long>(v));
    }
}

__device__ __forceinline__ void normalize3(double v[3])
{
    const double inv = rsqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
}

// ------------------------------------------------------------------------------------------
// Stage kernels.
// ------------------------------------------------------------------------------------------

// Camera rays of `tile_count` tiles: PinholeCamera::spawn_ray (pinholecamera.cpp:159-195) with
// ndc_to_camera of perspectivecamera.cpp:204-211, one jittered sample position per path.
// Path id = pixel * spp + sample.  Consecutive threads take consecutive samples of one pixel,
// then the next pixel of the tile row: a warp's rays are as coherent as the image allows.
__global__ void __launch_bounds__(StageThreads)
generate_kernel(const StreamParams p, const uint32_t* __restrict__ tiles, const uint32_t tile_count, const QueueView out, unsigned long long* stats)
{
    const unsigned long long per_tile = static_cast<unsigned long long>(p.tile_size) * p.tile_size * p.spp;
    const unsigned long long total = per_tile * tile_count;
    const unsigned long long padded = (total + 31ull) & ~31ull;
    unsigned local[StatCount] = { 0, 0, 0, 0, 0, 0 };
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < padded;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    {
        bool emit = false;
        uint32_t path = 0;
        float time_absolute = 0.0f, time_normalized = 0.0f;
        double o[3] = { 0.0, 0.0, 0.0 }, d[3] = { 0.0, 0.0, 0.0 };
        if (i < total)
        {
            const uint32_t tile = tiles[i / per_tile];
            const uint32_t r = static_cast<uint32_t>(i % per_tile);
            const uint32_t in_tile = r / p.spp, sample = r % p.spp;
            const uint32_t x = (tile % p.tiles_x) * p.tile_size + in_tile % p.tile_size;
            const uint32_t y = (tile / p.tiles_x) * p.tile_size + in_tile / p.tile_size;
            if (x < p.width && y < p.height)
            {
                emit = true;
                path = (y * p.width + x) * p.spp + sample;
                const double ndx = (x + rng(p.seed, path, 0, 0)) / p.width;
                const double ndy = (y + rng(p.seed, path, 0, 1)) / p.height;
                const double c[3] = { -((0.5 - ndx) * p.film_w), -((ndy - 0.5) * p.film_h), -p.focal };
                #pragma unroll
                for (int k = 0; k < 3; ++k)
                {
                    d[k] = p.cam[k * 4] * c[0] + p.cam[k * 4 + 1] * c[1] + p.cam[k * 4 + 2] * c[2];
                    o[k] = p.cam[k * 4 + 3];
                }
                normalize3(d);
                if (p.shutter_open != p.shutter_close)
                {
                    // One time sample per camera path: a float in [0, 1) (24 random bits), then
                    // ShadingRay::Time::create_with_normalized_time (shadingray.h:230-239).
                    time_normalized = static_cast<float>(static_cast<uint32_t>(rng(p.seed, path, 0, 3) * 16777216.0)) * (1.0f / 16777216.0f);
                    time_absolute = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, time_normalized), p.shutter_open), __fmul_rn(time_normalized, p.shutter_close));
                }
            }
        }
        const unsigned long long slot = enqueue_slot(out, emit);
        if (slot != ~0ull)
        {
            write_ray(out, slot, o, d, 0.0, DBL_MAX, time_absolute, time_normalized, ASGPU_VIS_CAMERA, path);
            write_parent(out, slot, nullptr);
            ++local[StatCamera];
        }
    }
    add_stats(stats, local);
}

// One path vertex per hit: hit point, face-forwarded world-space geometric normal (from the
// 36-byte triangle the leaf stores, normal_to_parent = transpose of parent_to_local), next origin
// = point + eps * normal (the parent == nullptr convention of Intersector::trace), then
//   * one shadow probe towards a light chosen uniformly, tmax = distance * (1 - 1e-6)
//     (Tracer::trace_between, tracer.h:252-259), flags ShadowRay;
//   * when depth < max_bounces, one cosine-weighted bounce (mappings.h:299-314), flags DiffuseRay.
// hit.assembly_instance holds the ItemRecord index here (raw_item launches).
// FUSED: ShadingPoint::refine_and_offset of the hit runs here (refine_core.h, the very function of
// refine_offset_kernel) instead of in a kernel of its own: the 80-byte parent record goes straight
// into the two child queues, hits and rays are read once.
__device__ __noinline__ void refine_offset_call(const SceneView* s, const double* org_dir, const float time_absolute, const float time_normalized,
                                                const double t, const uint32_t item, const uint32_t object_instance, const uint32_t primitive,
                                                const uint32_t slot, double* dst)
{
    refine_offset_one(*s, org_dir, org_dir + 3, time_absolute, time_normalized, t, item, object_instance, primitive, slot, dst);
}

template <bool FUSED>
__global__ void __launch_bounds__(StageThreads, 3)
shade_kernel(const StreamParams p, const SceneView s, const QueueView in, const asgpu_hit* __restrict__ hits, const asgpu_parent* refined,
             const QueueView probes, const QueueView next, const uint32_t depth, uint32_t* image, unsigned long long* stats)
{
    const unsigned long long n = min(*in.count, in.capacity);
    const unsigned long long padded = (n + 31ull) & ~31ull;
    const bool bounce = depth < p.max_bounces;
    unsigned local[StatCount] = { 0, 0, 0, 0, 0, 0 };
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < padded;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    {
        bool vertex = false;
        uint32_t path = 0, pixel = 0, identity = 0;
        float time_absolute = 0.0f, time_normalized = 0.0f;     // child rays inherit their path's time (pathtracer.h:764)
        double o[3] = { 0.0, 0.0, 0.0 }, nrm[3] = { 0.0, 1.0, 0.0 };
        double parent_words[10];                                 // FUSED: this vertex's asgpu_parent record
        const bool with_parents = FUSED || refined != nullptr;
        if (i < n)
        {
            path = in.path[i];
            time_absolute = in.time_absolute[i];
            time_normalized = in.time_normalized[i];
            pixel = path / p.spp;
            const unsigned long long* hw = reinterpret_cast<const unsigned long long*>(hits + i);
            const unsigned long long w0 = hw[0], w2 = hw[2], w3 = hw[3], w4 = hw[4];
            if (static_cast<uint32_t>(w4 >> 32) == 0) ++local[StatEscaped];
            else
            {
                vertex = true;
                const double t = __longlong_as_double(static_cast<long long>(w0));
                const uint32_t item = static_cast<uint32_t>(w2), object_instance = static_cast<uint32_t>(w2 >> 32);
                const uint32_t primitive = static_cast<uint32_t>(w3), slot = static_cast<uint32_t>(w3 >> 32);
                double dir[3];
                if (FUSED)
                {
                    double org_dir[6];
                    #pragma unroll
                    for (int k = 0; k < 3; ++k) { org_dir[k] = in.org[i * 3 + k]; org_dir[3 + k] = in.dir[i * 3 + k]; }
                    refine_offset_call(&s, org_dir, time_absolute, time_normalized, t, item, object_instance, primitive, slot, parent_words);
                    #pragma unroll
                    for (int k = 0; k < 3; ++k) { dir[k] = org_dir[3 + k]; o[k] = org_dir[k] + t * dir[k]; }
                }
                else
                {
                    #pragma unroll
                    for (int k = 0; k < 3; ++k) { dir[k] = in.dir[i * 3 + k]; o[k] = in.org[i * 3 + k] + t * dir[k]; }
                }
                // Geometric normal.
                const uint8_t* ip = s.blob + s.items + static_cast<uint64_t>(item) * sizeof(ItemRecord);
                const uint4 meta = load16(ip + 96);
                const uint8_t* tp = s.blob + s.trees + static_cast<uint64_t>(meta.x) * sizeof(TreeDesc);
                // The hit triangle: the leaf's, or for a moving triangle the one interpolated at the ray time.
                TriD tri;
                hit_triangle(s.blob + load_u64(tp + offsetof(TreeDesc, tris)) + static_cast<uint64_t>(slot) * sizeof(TriRecord),
                             s.blob + load_u64(tp + offsetof(TreeDesc, poses)), time_normalized, tri);
                const double* e0 = tri.e0;
                const double* e1 = tri.e1;
                const double nl[3] = { e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0] };
                // normal_to_parent = transpose of parent_to_local; an animated instance: its transform at the ray time.
                double m[12];
                if (meta.w >= 2) animated_item_matrix(s.blob, ip, meta.w, time_absolute, m);
                else
                {
                    #pragma unroll
                    for (int k = 0; k < 12; ++k) m[k] = load_f64(ip + k * 8);
                }
                #pragma unroll
                for (int k = 0; k < 3; ++k)
                    nrm[k] = m[k] * nl[0] + m[4 + k] * nl[1] + m[8 + k] * nl[2];
                normalize3(nrm);
                if (nrm[0] * dir[0] + nrm[1] * dir[1] + nrm[2] * dir[2] > 0.0) { nrm[0] = -nrm[0]; nrm[1] = -nrm[1]; nrm[2] = -nrm[2]; }
                // Next origin: the hit point itself when the child rays carry the refined parent record,
                // else offset along the normal.
                if (!with_parents)
                {
                    #pragma unroll
                    for (int k = 0; k < 3; ++k) o[k] += p.eps * nrm[k];
                }
                identity = primitive * 2654435761u + object_instance * 0x9E3779B1u + meta.z * 0x85EBCA6Bu + slot;
                ++local[StatHits];
            }
        }

        // Per-pixel accumulators.  Consecutive rays of a queue are mostly samples of the same pixel
        // (64 per pixel): the lanes of a pixel are summed in the warp first (match + reduce), one
        // atomic per pixel, counter and warp instead of one per ray.
        {
            const unsigned live = __ballot_sync(0xFFFFFFFFu, i < n);
            if (i < n)
            {
                const unsigned peers = __match_any_sync(live, pixel);
                const unsigned hits_here = __reduce_add_sync(peers, vertex ? 1u : 0u);
                const unsigned escaped_here = __popc(peers) - hits_here;
                const unsigned identity_sum = __reduce_add_sync(peers, identity);
                if ((threadIdx.x & 31) == static_cast<unsigned>(__ffs(peers) - 1))
                {
                    uint32_t* px = image + static_cast<size_t>(pixel) * 4;
                    if (hits_here) { atomicAdd(px + 0, hits_here); atomicAdd(px + 3, identity_sum); }
                    if (escaped_here) atomicAdd(px + 2, escaped_here);
                }
            }
        }

        // Shadow probe.
        {
            double d[3] = { 0.0, 0.0, 0.0 };
            double tmax = 0.0;
            if (vertex)
            {
                uint32_t k = static_cast<uint32_t>(rng(p.seed, path, depth + 1, 2) * p.light_count);
                if (k >= p.light_count) k = p.light_count - 1;
                d[0] = p.lights[k][0] - o[0]; d[1] = p.lights[k][1] - o[1]; d[2] = p.lights[k][2] - o[2];
                const double dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                const double inv = 1.0 / dist;
                d[0] *= inv; d[1] *= inv; d[2] *= inv;
                tmax = dist * (1.0 - 1.0e-6);
            }
            const unsigned long long slot = enqueue_slot(probes, vertex);
            if (slot != ~0ull)
            {
                write_ray(probes, slot, o, d, 0.0, tmax, time_absolute, time_normalized, ASGPU_VIS_SHADOW, path);
                write_parent(probes, slot, FUSED ? reinterpret_cast<const asgpu_parent*>(parent_words) : (refined ? refined + i : nullptr));
                ++local[StatProbe];
            }
        }

        // Bounce.
        if (bounce)
        {
            double d[3] = { 0.0, 0.0, 0.0 };
            if (vertex)
            {
                const double s0 = rng(p.seed, path, depth + 1, 0), s1 = rng(p.seed, path, depth + 1, 1);
                // The local direction in float (it is renormalised in double below; the sampling
                // pattern does not need more, the unit length does).
                float sin_phi, cos_phi;
                sincospif(2.0f * static_cast<float>(s0), &sin_phi, &cos_phi);
                const float cos_theta = sqrtf(1.0f - static_cast<float>(s1)), sin_theta = sqrtf(static_cast<float>(s1));
                // Orthonormal frame (u, nrm, v).
                double aux[3] = { 1.0, 0.0, 0.0 };
                if (fabs(nrm[0]) > 0.9) { aux[0] = 0.0; aux[1] = 1.0; }
                double u[3] = { aux[1] * nrm[2] - aux[2] * nrm[1], aux[2] * nrm[0] - aux[0] * nrm[2], aux[0] * nrm[1] - aux[1] * nrm[0] };
                normalize3(u);
                const double v[3] = { nrm[1] * u[2] - nrm[2] * u[1], nrm[2] * u[0] - nrm[0] * u[2], nrm[0] * u[1] - nrm[1] * u[0] };
                const double lx = cos_phi * sin_theta, ly = cos_theta, lz = sin_phi * sin_theta;
                #pragma unroll
                for (int k = 0; k < 3; ++k) d[k] = lx * u[k] + ly * nrm[k] + lz * v[k];
                normalize3(d);
            }
            const unsigned long long slot = enqueue_slot(next, vertex);
            if (slot != ~0ull)
            {
                write_ray(next, slot, o, d, 0.0, DBL_MAX, time_absolute, time_normalized, ASGPU_VIS_DIFFUSE, path);
                write_parent(next, slot, FUSED ? reinterpret_cast<const asgpu_parent*>(parent_words) : (refined ? refined + i : nullptr));
                ++local[StatBounce];
            }
        }
    }
    add_stats(stats, local);
}

__global__ void __launch_bounds__(StageThreads)
accumulate_kernel(const StreamParams p, const QueueView probes, const uint8_t* __restrict__ occluded, uint32_t* image, unsigned long long* stats)
{
    const unsigned long long n = min(*probes.count, probes.capacity);
    unsigned local[StatCount] = { 0, 0, 0, 0, 0, 0 };
    const unsigned long long padded = (n + 31ull) & ~31ull;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < padded;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    {
        const bool lit = i < n && occluded[i] == 0;
        const unsigned live = __ballot_sync(0xFFFFFFFFu, lit);
        if (lit)
        {
            const uint32_t pixel = probes.path[i] / p.spp;
            const unsigned peers = __match_any_sync(live, pixel);
            if ((threadIdx.x & 31) == static_cast<unsigned>(__ffs(peers) - 1)) atomicAdd(image + static_cast<size_t>(pixel) * 4 + 1, static_cast<uint32_t>(__popc(peers)));
            ++local[StatUnoccluded];
        }
    }
    add_stats(stats, local);
}

__global__ void push_count_kernel(unsigned long long* count, const unsigned long long add) { *count += add; }

__global__ void fill_path_ids_kernel(uint32_t* path, const unsigned long long* count, const unsigned long long n)
{
    const unsigned long long base = *count;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
        path[base + i] = static_cast<uint32_t>(base + i);
}

struct Captured
{
    int                     kind = 0;
    uint32_t                depth = 0;
    std::vector<double>     org, dir, tmin, tmax;
    std::vector<float>      time_absolute, time_normalized;
    std::vector<uint32_t>   flags, path;
    std::vector<uint8_t>    results;
    std::vector<asgpu_parent> parents;
};

}   // anonymous namespace

struct asgpu_path_stream
{
    asgpu_scene*            scene = nullptr;
    asgpu_path_stream_desc  desc;
    StreamParams            params;
    uint32_t                tiles_x = 0, tiles_y = 0;
    size_t                  tiles_per_batch = 0;
    asgpu_ray_queue*        qa = nullptr;
    asgpu_ray_queue*        qb = nullptr;
    asgpu_ray_queue*        qp = nullptr;
    asgpu_hit*              hits = nullptr;
    asgpu_parent*           refined = nullptr;      // refine_and_offset of the current wavefront's hits (ASGPU_STREAM_PARENTS)
    bool                    fuse_refine = false;    // ... computed inside shade_kernel instead (ASGPU_FUSE_REFINE=1)
    uint8_t*                occluded = nullptr;
    uint32_t*               image = nullptr;
    uint32_t*               tiles_dev = nullptr;
    uint32_t*               tile_pixels = nullptr;  // staging of asgpu_path_stream_read_tiles
    size_t                  tile_pixels_bytes = 0;
    unsigned long long*     stats_dev = nullptr;
    unsigned long long*     cursors = nullptr;      // ray-queue cursors of the trace launches (ring)
    uint64_t                cursor_next = 0;
    uint64_t                wavefronts = 0;
    uint64_t                launches = 0;
    std::vector<uint32_t>   item_ids;               // ItemRecord index -> caller's assembly instance id
    bool                    capture_armed = false;
    size_t                  capture_budget = 0;
    std::vector<Captured>   captured;
    // Optional per-launch timing (asgpu_path_stream_set_profiling): event pairs recorded on the
    // launch stream around every kernel, by kind: 0 closest-hit trace, 1 probe trace, 2 refine, 3 stage.
    bool                    profiling = false;
    std::vector<cudaEvent_t> event_pool;
    size_t                  events_used = 0;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> timed;
};

namespace
{

int stage_grid(const asgpu_scene* scene) { return scene->sm_count * 8; }

// Event pair around one launch when profiling is on (kind: 0 closest-hit trace, 1 probe trace,
// 2 refine, 3 generate / shade / accumulate).  Events are pooled: no allocation in steady state.
cudaEvent_t next_event(asgpu_path_stream* ps)
{
    if (ps->events_used == ps->event_pool.size())
    {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        ps->event_pool.push_back(e);
    }
    return ps->event_pool[ps->events_used++];
}

struct TimedLaunch
{
    asgpu_path_stream*  ps;
    cudaStream_t        stream;
    int                 kind;
    cudaEvent_t         start = nullptr;
    TimedLaunch(asgpu_path_stream* ps_, const int kind_, cudaStream_t stream_) : ps(ps_), stream(stream_), kind(kind_)
    {
        if (ps->profiling && (start = next_event(ps)) != nullptr) cudaEventRecord(start, stream);
    }
    ~TimedLaunch()
    {
        if (start == nullptr) return;
        cudaEvent_t stop = next_event(ps);
        if (stop == nullptr) return;
        cudaEventRecord(stop, stream);
        ps->timed.push_back(std::make_pair(kind, std::make_pair(start, stop)));
    }
};

int capture_wavefront(asgpu_path_stream* ps, const asgpu_ray_queue* q, const int kind, const uint32_t depth, cudaStream_t stream)
{
    ASGPU_CUDA(cudaStreamSynchronize(stream), "cudaStreamSynchronize(capture)");
    unsigned long long n = 0;
    ASGPU_CUDA(cudaMemcpy(&n, q->count, 8, cudaMemcpyDeviceToHost), "cudaMemcpy(queue count)");
    n = std::min<unsigned long long>(n, q->capacity);
    if (n > ps->capture_budget) return fail(ASGPU_E_INVALID, "capture budget exceeded");
    ps->capture_budget -= n;
    ps->captured.emplace_back();
    Captured& c = ps->captured.back();
    c.kind = kind; c.depth = depth;
    c.org.resize(n * 3); c.dir.resize(n * 3); c.tmin.resize(n); c.tmax.resize(n); c.flags.resize(n); c.path.resize(n);
    c.time_absolute.resize(n); c.time_normalized.resize(n);
    c.results.resize(n * (kind == 0 ? sizeof(asgpu_hit) : 1));
    c.parents.clear();
    if (n == 0) return ASGPU_OK;
    ASGPU_CUDA(cudaMemcpy(c.org.data(), q->org, n * 24, cudaMemcpyDeviceToHost), "capture");
    ASGPU_CUDA(cudaMemcpy(c.dir.data(), q->dir, n * 24, cudaMemcpyDeviceToHost), "capture");
    ASGPU_CUDA(cudaMemcpy(c.tmin.data(), q->tmin, n * 8, cudaMemcpyDeviceToHost), "capture");
    ASGPU_CUDA(cudaMemcpy(c.tmax.data(), q->tmax, n * 8, cudaMemcpyDeviceToHost), "capture");
    ASGPU_CUDA(cudaMemcpy(c.time_absolute.data(), q->time_absolute, n * 4, cudaMemcpyDeviceToHost), "capture");
    ASGPU_CUDA(cudaMemcpy(c.time_normalized.data(), q->time_normalized, n * 4, cudaMemcpyDeviceToHost), "capture");
    ASGPU_CUDA(cudaMemcpy(c.flags.data(), q->flags, n * 4, cudaMemcpyDeviceToHost), "capture");
    ASGPU_CUDA(cudaMemcpy(c.path.data(), q->path, n * 4, cudaMemcpyDeviceToHost), "capture");
    c.parents.resize(n);
    if (q->parents) ASGPU_CUDA(cudaMemcpy(c.parents.data(), q->parents, n * sizeof(asgpu_parent), cudaMemcpyDeviceToHost), "capture");
    else for (asgpu_parent& pr : c.parents) { std::memset(&pr, 0, sizeof(pr)); pr.assembly_instance = ASGPU_MISS; }
    if (kind == 0)
    {
        ASGPU_CUDA(cudaMemcpy(c.results.data(), ps->hits, n * sizeof(asgpu_hit), cudaMemcpyDeviceToHost), "capture");
        // Raw ItemRecord indices back to the caller's instance ids (what asgpu_trace reports).
        asgpu_hit* h = reinterpret_cast<asgpu_hit*>(c.results.data());
        for (unsigned long long i = 0; i < n; ++i)
            if (h[i].prim_type != 0 && h[i].assembly_instance < ps->item_ids.size()) h[i].assembly_instance = ps->item_ids[h[i].assembly_instance];
    }
    else ASGPU_CUDA(cudaMemcpy(c.results.data(), ps->occluded, n, cudaMemcpyDeviceToHost), "capture");
    return ASGPU_OK;
}

int trace_queue(asgpu_scene* scene, asgpu_ray_queue* q, asgpu_hit* hits, uint8_t* occluded, const bool any_hit, const uint32_t flags,
                unsigned long long* cursor, const bool raw_item, void* stream, const bool use_parents = true)
{
    const bool wide = (flags & ASGPU_TRACE_EXACT) == 0;
    if (flags & ASGPU_TRACE_SORT) return fail(ASGPU_E_UNSUPPORTED, "ASGPU_TRACE_SORT needs the ray count on the host: not available for queues");
    if (wide && !(scene->header.flags & ASGPU_SCENE_WIDE)) return fail(ASGPU_E_INVALID, "scene was created without the wide layout");
    if (!wide && !(scene->header.flags & ASGPU_SCENE_EXACT)) return fail(ASGPU_E_INVALID, "scene was created without the exact layout");
    const asgpu_rays rays = rays_of(q);
    const int err = launch_trace(scene->view, rays, q->capacity, hits, occluded, any_hit, wide, cursor,
                                 (flags & ASGPU_TRACE_COUNTERS) ? scene->counters : nullptr, nullptr, scene->sm_count, stream, q->count, raw_item,
                                 use_parents ? q->parents : nullptr);
    if (err != 0) return fail_cuda(static_cast<cudaError_t>(err), "kernel launch");
    ++scene->launches;
    return ASGPU_OK;
}

}   // anonymous namespace

extern "C" {

// ---- queues -------------------------------------------------------------------------------

asgpu_ray_queue* asgpu_queue_create(asgpu_scene* scene, size_t capacity)
{
    if (!scene || capacity == 0) { fail(ASGPU_E_INVALID, "queue needs a scene and a capacity"); return nullptr; }
    if (capacity > 0xFFFFFFFFull) { fail(ASGPU_E_UNSUPPORTED, "queue capacity above 2^32 - 1 rays"); return nullptr; }
    asgpu_ray_queue* q = new (std::nothrow) asgpu_ray_queue();
    if (!q) { fail(ASGPU_E_NOMEM, "out of host memory"); return nullptr; }
    q->scene = scene;
    q->capacity = capacity;
    cudaSetDevice(scene->device);
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&q->org, capacity * 24);
    if (e == cudaSuccess) e = cudaMalloc(&q->dir, capacity * 24);
    if (e == cudaSuccess) e = cudaMalloc(&q->tmin, capacity * 8);
    if (e == cudaSuccess) e = cudaMalloc(&q->tmax, capacity * 8);
    if (e == cudaSuccess) e = cudaMalloc(&q->time_absolute, capacity * 4);
    if (e == cudaSuccess) e = cudaMalloc(&q->time_normalized, capacity * 4);
    if (e == cudaSuccess) e = cudaMalloc(&q->flags, capacity * 4);
    if (e == cudaSuccess) e = cudaMalloc(&q->path, capacity * 4);
    if (e == cudaSuccess) e = cudaMalloc(&q->count, 8);
    if (e == cudaSuccess) e = cudaMemset(q->count, 0, 8);
    if (e != cudaSuccess) { fail_cuda(e, "cudaMalloc(queue)"); asgpu_queue_destroy(q); return nullptr; }
    return q;
}

void asgpu_queue_destroy(asgpu_ray_queue* q)
{
    if (!q) return;
    cudaSetDevice(q->scene->device);
    cudaFree(q->org); cudaFree(q->dir); cudaFree(q->tmin); cudaFree(q->tmax);
    cudaFree(q->time_absolute); cudaFree(q->time_normalized);
    cudaFree(q->flags); cudaFree(q->path); cudaFree(q->count); cudaFree(q->parents);
    delete q;
}

size_t asgpu_queue_capacity(const asgpu_ray_queue* q) { return q ? q->capacity : 0; }

int asgpu_queue_device_arrays(asgpu_ray_queue* q, asgpu_rays* rays, uint32_t** path_ids, uint64_t** count)
{
    if (!q) return fail(ASGPU_E_INVALID, "null queue");
    if (rays) *rays = rays_of(q);
    if (path_ids) *path_ids = q->path;
    if (count) *count = reinterpret_cast<uint64_t*>(q->count);
    return ASGPU_OK;
}

int asgpu_queue_reset(asgpu_ray_queue* q, void* stream)
{
    if (!q) return fail(ASGPU_E_INVALID, "null queue");
    ASGPU_CUDA(cudaSetDevice(q->scene->device), "cudaSetDevice");
    ASGPU_CUDA(cudaMemsetAsync(q->count, 0, 8, static_cast<cudaStream_t>(stream)), "cudaMemsetAsync(queue count)");
    return ASGPU_OK;
}

int asgpu_queue_count(asgpu_ray_queue* q, void* stream, uint64_t* count)
{
    if (!q || !count) return fail(ASGPU_E_INVALID, "null argument");
    ASGPU_CUDA(cudaSetDevice(q->scene->device), "cudaSetDevice");
    unsigned long long n = 0;
    ASGPU_CUDA(cudaMemcpyAsync(&n, q->count, 8, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)), "cudaMemcpyAsync(queue count)");
    ASGPU_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)), "cudaStreamSynchronize");
    *count = std::min<unsigned long long>(n, q->capacity);
    return ASGPU_OK;
}

int asgpu_queue_push_host(asgpu_ray_queue* q, const asgpu_rays* rays, const uint32_t* path_ids, size_t n, void* stream_)
{
    if (!q) return fail(ASGPU_E_INVALID, "null queue");
    if (n == 0) return ASGPU_OK;
    if (!rays || !rays->org || !rays->dir || !rays->tmin || !rays->tmax) return fail(ASGPU_E_INVALID, "ray batch misses a mandatory array");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    uint64_t have = 0;
    const int rc = asgpu_queue_count(q, stream_, &have);
    if (rc != ASGPU_OK) return rc;
    if (have + n > q->capacity) return fail(ASGPU_E_INVALID, "queue overflow");
    ASGPU_CUDA(cudaMemcpyAsync(q->org + have * 3, rays->org, n * 24, cudaMemcpyHostToDevice, stream), "H2D org");
    ASGPU_CUDA(cudaMemcpyAsync(q->dir + have * 3, rays->dir, n * 24, cudaMemcpyHostToDevice, stream), "H2D dir");
    ASGPU_CUDA(cudaMemcpyAsync(q->tmin + have, rays->tmin, n * 8, cudaMemcpyHostToDevice, stream), "H2D tmin");
    ASGPU_CUDA(cudaMemcpyAsync(q->tmax + have, rays->tmax, n * 8, cudaMemcpyHostToDevice, stream), "H2D tmax");
    if (rays->time_absolute) ASGPU_CUDA(cudaMemcpyAsync(q->time_absolute + have, rays->time_absolute, n * 4, cudaMemcpyHostToDevice, stream), "H2D time");
    else ASGPU_CUDA(cudaMemsetAsync(q->time_absolute + have, 0, n * 4, stream), "time");
    if (rays->time_normalized) ASGPU_CUDA(cudaMemcpyAsync(q->time_normalized + have, rays->time_normalized, n * 4, cudaMemcpyHostToDevice, stream), "H2D time");
    else ASGPU_CUDA(cudaMemsetAsync(q->time_normalized + have, 0, n * 4, stream), "time");
    if (rays->flags) ASGPU_CUDA(cudaMemcpyAsync(q->flags + have, rays->flags, n * 4, cudaMemcpyHostToDevice, stream), "H2D flags");
    else ASGPU_CUDA(cudaMemsetAsync(q->flags + have, 0xFF, n * 4, stream), "flags");
    if (path_ids) ASGPU_CUDA(cudaMemcpyAsync(q->path + have, path_ids, n * 4, cudaMemcpyHostToDevice, stream), "H2D path ids");
    else fill_path_ids_kernel<<<64, 256, 0, stream>>>(q->path, q->count, n);
    push_count_kernel<<<1, 1, 0, stream>>>(q->count, n);
    ASGPU_CUDA(cudaGetLastError(), "queue push");
    return ASGPU_OK;
}

int asgpu_trace_queue(asgpu_scene* scene, asgpu_ray_queue* queue, asgpu_hit* hits, uint32_t flags, void* stream)
{
    if (!scene || !queue || !hits) return fail(ASGPU_E_INVALID, "null argument");
    ASGPU_CUDA(cudaSetDevice(scene->device), "cudaSetDevice");
    return trace_queue(scene, queue, hits, nullptr, false, flags, scene->queue + (scene->queue_next++ % QueueRing), false, stream);
}

int asgpu_trace_probe_queue(asgpu_scene* scene, asgpu_ray_queue* queue, uint8_t* occluded, uint32_t flags, void* stream)
{
    if (!scene || !queue || !occluded) return fail(ASGPU_E_INVALID, "null argument");
    ASGPU_CUDA(cudaSetDevice(scene->device), "cudaSetDevice");
    return trace_queue(scene, queue, nullptr, occluded, true, flags, scene->queue + (scene->queue_next++ % QueueRing), false, stream);
}

// ---- path stream --------------------------------------------------------------------------

asgpu_path_stream* asgpu_path_stream_create(asgpu_scene* scene, const asgpu_path_stream_desc* desc, size_t queue_capacity)
{
    if (!scene || !desc) { fail(ASGPU_E_INVALID, "null argument"); return nullptr; }
    if (desc->width == 0 || desc->height == 0 || desc->spp == 0 || desc->tile_size == 0) { fail(ASGPU_E_INVALID, "empty image, sample count or tile"); return nullptr; }
    if (desc->light_count == 0 || desc->light_count > 8) { fail(ASGPU_E_INVALID, "light_count must be 1..8"); return nullptr; }
    if (desc->max_bounces > 200) { fail(ASGPU_E_INVALID, "max_bounces above 200"); return nullptr; }
    if (static_cast<uint64_t>(desc->width) * desc->height * desc->spp > 0xFFFFFFFFull) { fail(ASGPU_E_UNSUPPORTED, "more than 2^32 - 1 paths per frame"); return nullptr; }
    if (!(desc->shutter_open <= desc->shutter_close)) { fail(ASGPU_E_INVALID, "shutter_open must not exceed shutter_close"); return nullptr; }
    if (!(scene->header.flags & ASGPU_SCENE_EXACT)) { fail(ASGPU_E_INVALID, "the path stream needs the per-slot triangle records of the exact layout"); return nullptr; }
    const bool with_parents = (desc->stream_flags & ASGPU_STREAM_PARENTS) != 0;
    if (with_parents && !scene->has_source) { fail(ASGPU_E_INVALID, "ASGPU_STREAM_PARENTS needs a scene with source geometry (asgpu_scene_create_ex)"); return nullptr; }
    const size_t per_tile = static_cast<size_t>(desc->tile_size) * desc->tile_size * desc->spp;
    if (queue_capacity < per_tile) { fail(ASGPU_E_INVALID, "queue capacity below one tile's paths"); return nullptr; }

    asgpu_path_stream* ps = new (std::nothrow) asgpu_path_stream();
    if (!ps) { fail(ASGPU_E_NOMEM, "out of host memory"); return nullptr; }
    ps->scene = scene;
    ps->desc = *desc;
    ps->tiles_x = (desc->width + desc->tile_size - 1) / desc->tile_size;
    ps->tiles_y = (desc->height + desc->tile_size - 1) / desc->tile_size;
    ps->tiles_per_batch = queue_capacity / per_tile;
    const size_t capacity = ps->tiles_per_batch * per_tile;

    StreamParams& p = ps->params;
    std::memset(&p, 0, sizeof(p));
    p.width = desc->width; p.height = desc->height; p.spp = desc->spp; p.max_bounces = desc->max_bounces;
    p.tile_size = desc->tile_size; p.tiles_x = ps->tiles_x; p.light_count = desc->light_count; p.seed = desc->seed;
    std::memcpy(p.cam, desc->camera_to_world, sizeof(p.cam));
    p.film_w = desc->film_width; p.film_h = desc->film_height; p.focal = desc->focal_length;
    std::memcpy(p.lights, desc->lights, sizeof(p.lights));
    p.eps = desc->offset_eps;
    p.shutter_open = desc->shutter_open; p.shutter_close = desc->shutter_close;

    cudaSetDevice(scene->device);
    ps->qa = asgpu_queue_create(scene, capacity);
    ps->qb = asgpu_queue_create(scene, capacity);
    ps->qp = asgpu_queue_create(scene, capacity);
    cudaError_t e = cudaSuccess;
    const size_t pixels = static_cast<size_t>(desc->width) * desc->height;
    if (with_parents && ps->qa && ps->qb && ps->qp)
    {
        if (e == cudaSuccess) e = cudaMalloc(&ps->qa->parents, capacity * sizeof(asgpu_parent));
        if (e == cudaSuccess) e = cudaMalloc(&ps->qb->parents, capacity * sizeof(asgpu_parent));
        if (e == cudaSuccess) e = cudaMalloc(&ps->qp->parents, capacity * sizeof(asgpu_parent));
        if (e == cudaSuccess) e = cudaMalloc(&ps->refined, capacity * sizeof(asgpu_parent));
        // refine_and_offset inside shade_kernel (the default: -10 % on refine + shade together, +3 % on
        // a C5 frame, profiles/r2/r2_fuse.log); ASGPU_FUSE_REFINE=0 keeps the kernel of its own.
        ps->fuse_refine = true;
        if (const char* fuse = getenv("ASGPU_FUSE_REFINE")) ps->fuse_refine = atoi(fuse) != 0;
    }
    if (e == cudaSuccess) e = cudaMalloc(&ps->hits, capacity * sizeof(asgpu_hit));
    if (e == cudaSuccess) e = cudaMalloc(&ps->occluded, capacity);
    if (e == cudaSuccess) e = cudaMalloc(&ps->image, pixels * 16);
    if (e == cudaSuccess) e = cudaMemset(ps->image, 0, pixels * 16);
    if (e == cudaSuccess) e = cudaMalloc(&ps->tiles_dev, static_cast<size_t>(ps->tiles_x) * ps->tiles_y * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ps->stats_dev, StatCount * 8);
    if (e == cudaSuccess) e = cudaMemset(ps->stats_dev, 0, StatCount * 8);
    if (e == cudaSuccess) e = cudaMalloc(&ps->cursors, QueueRing * 8);
    std::vector<ItemRecord> items(scene->header.item_count);
    if (e == cudaSuccess && !items.empty())
        e = cudaMemcpy(items.data(), scene->blob + scene->header.items, items.size() * sizeof(ItemRecord), cudaMemcpyDeviceToHost);
    if (!ps->qa || !ps->qb || !ps->qp || e != cudaSuccess)
    {
        if (e != cudaSuccess) fail_cuda(e, "cudaMalloc(path stream)");
        asgpu_path_stream_destroy(ps);
        return nullptr;
    }
    ps->item_ids.resize(items.size());
    for (size_t i = 0; i < items.size(); ++i) ps->item_ids[i] = items[i].assembly_instance;
    return ps;
}

void asgpu_path_stream_destroy(asgpu_path_stream* ps)
{
    if (!ps) return;
    cudaSetDevice(ps->scene->device);
    asgpu_queue_destroy(ps->qa); asgpu_queue_destroy(ps->qb); asgpu_queue_destroy(ps->qp);
    cudaFree(ps->hits); cudaFree(ps->refined); cudaFree(ps->occluded); cudaFree(ps->image); cudaFree(ps->tiles_dev); cudaFree(ps->tile_pixels);
    cudaFree(ps->stats_dev); cudaFree(ps->cursors);
    for (cudaEvent_t e : ps->event_pool) cudaEventDestroy(e);
    delete ps;
}

uint32_t asgpu_path_stream_tile_count(const asgpu_path_stream* ps) { return ps ? ps->tiles_x * ps->tiles_y : 0; }

int asgpu_path_stream_render(asgpu_path_stream* ps, const uint32_t* tiles, size_t tile_count, void* cuda_stream)
{
    if (!ps) return fail(ASGPU_E_INVALID, "null path stream");
    if (tile_count == 0) return ASGPU_OK;
    if (!tiles) return fail(ASGPU_E_INVALID, "null tile list");
    const uint32_t all_tiles = ps->tiles_x * ps->tiles_y;
    if (tile_count > all_tiles) return fail(ASGPU_E_INVALID, "more tiles than the frame has");
    for (size_t i = 0; i < tile_count; ++i)
        if (tiles[i] >= all_tiles) return fail(ASGPU_E_INVALID, "tile index out of range");
    asgpu_scene* scene = ps->scene;
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    ASGPU_CUDA(cudaSetDevice(scene->device), "cudaSetDevice");
    // The tile list is small (4 bytes per 32 x 32 x spp paths): a synchronous upload keeps the
    // caller's array free to go away; everything after it is asynchronous on `stream`.
    ASGPU_CUDA(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    ASGPU_CUDA(cudaMemcpy(ps->tiles_dev, tiles, tile_count * 4, cudaMemcpyHostToDevice), "cudaMemcpy(tiles)");

    const int grid = stage_grid(scene);
    const uint32_t flags = ps->desc.trace_flags & (ASGPU_TRACE_EXACT | ASGPU_TRACE_COUNTERS);
    const QueueView vp = view_of(ps->qp);
    for (size_t begin = 0; begin < tile_count; begin += ps->tiles_per_batch)
    {
        const uint32_t count = static_cast<uint32_t>(std::min(ps->tiles_per_batch, tile_count - begin));
        asgpu_ray_queue* qa = ps->qa;
        asgpu_ray_queue* qb = ps->qb;
        ASGPU_CUDA(cudaMemsetAsync(qa->count, 0, 8, stream), "cudaMemsetAsync");
        {
            TimedLaunch timed(ps, 3, stream);
            generate_kernel<<<grid, StageThreads, 0, stream>>>(ps->params, ps->tiles_dev + begin, count, view_of(qa), ps->stats_dev);
        }
        ASGPU_CUDA(cudaGetLastError(), "generate_kernel");
        ++ps->launches;
        for (uint32_t depth = 0; depth <= ps->desc.max_bounces; ++depth)
        {
            int rc;
            {
                TimedLaunch timed(ps, 0, stream);
                // Camera rays have no parent shading point: their launch does not look at the (all "none") records.
                rc = trace_queue(scene, qa, ps->hits, nullptr, false, flags, ps->cursors + (ps->cursor_next++ % QueueRing), true, stream, depth != 0);
            }
            if (rc != ASGPU_OK) return rc;
            ++ps->launches; ++ps->wavefronts;
            if (ps->capture_armed && (rc = capture_wavefront(ps, qa, 0, depth, stream)) != ASGPU_OK) return rc;
            ASGPU_CUDA(cudaMemsetAsync(qb->count, 0, 8, stream), "cudaMemsetAsync");
            ASGPU_CUDA(cudaMemsetAsync(ps->qp->count, 0, 8, stream), "cudaMemsetAsync");
            if (ps->refined && !ps->fuse_refine)
            {
                TimedLaunch timed(ps, 2, stream);
                const int er = launch_refine_offset(scene->view, rays_of(qa), ps->hits, qa->capacity, qa->count, true, nullptr, 0, ps->refined, scene->sm_count, stream);
                if (er != 0) return fail_cuda(static_cast<cudaError_t>(er), "refine_offset_kernel");
                ++ps->launches;
            }
            {
                TimedLaunch timed(ps, 3, stream);
                if (ps->refined && ps->fuse_refine)
                    shade_kernel<true><<<grid, StageThreads, 0, stream>>>(ps->params, scene->view, view_of(qa), ps->hits, nullptr, vp, view_of(qb), depth, ps->image, ps->stats_dev);
                else
                    shade_kernel<false><<<grid, StageThreads, 0, stream>>>(ps->params, scene->view, view_of(qa), ps->hits, ps->refined, vp, view_of(qb), depth, ps->image, ps->stats_dev);
            }
            ASGPU_CUDA(cudaGetLastError(), "shade_kernel");
            ++ps->launches;
            {
                TimedLaunch timed(ps, 1, stream);
                rc = trace_queue(scene, ps->qp, nullptr, ps->occluded, true, flags, ps->cursors + (ps->cursor_next++ % QueueRing), true, stream);
            }
            if (rc != ASGPU_OK) return rc;
            ++ps->launches;
            if (ps->capture_armed && (rc = capture_wavefront(ps, ps->qp, 1, depth, stream)) != ASGPU_OK) return rc;
            {
                TimedLaunch timed(ps, 3, stream);
                accumulate_kernel<<<grid, StageThreads, 0, stream>>>(ps->params, vp, ps->occluded, ps->image, ps->stats_dev);
            }
            ASGPU_CUDA(cudaGetLastError(), "accumulate_kernel");
            ++ps->launches;
            std::swap(qa, qb);
        }
    }
    ps->capture_armed = false;
    return ASGPU_OK;
}

int asgpu_path_stream_read_image(asgpu_path_stream* ps, uint32_t* accum)
{
    if (!ps || !accum) return fail(ASGPU_E_INVALID, "null argument");
    ASGPU_CUDA(cudaSetDevice(ps->scene->device), "cudaSetDevice");
    ASGPU_CUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    ASGPU_CUDA(cudaMemcpy(accum, ps->image, static_cast<size_t>(ps->desc.width) * ps->desc.height * 16, cudaMemcpyDeviceToHost), "cudaMemcpy(image)");
    return ASGPU_OK;
}

namespace
{

// Pixels of the listed tiles, tile after tile (tile_size x tile_size x 4 words each, row-major inside
// the tile; pixels of an edge tile that fall outside the image are zero).
__global__ void gather_tiles_kernel(const uint32_t* image, const uint32_t* tiles, const size_t tile_count, const uint32_t width, const uint32_t height,
                                    const uint32_t tile_size, const uint32_t tiles_x, uint4* out)
{
    const size_t per_tile = static_cast<size_t>(tile_size) * tile_size;
    const size_t total = tile_count * per_tile;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    {
        const uint32_t tile = tiles[i / per_tile];
        const uint32_t in_tile = static_cast<uint32_t>(i % per_tile);
        const uint32_t x = (tile % tiles_x) * tile_size + in_tile % tile_size;
        const uint32_t y = (tile / tiles_x) * tile_size + in_tile / tile_size;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (x < width && y < height) v = reinterpret_cast<const uint4*>(image)[static_cast<size_t>(y) * width + x];
        out[i] = v;
    }
}

}   // anonymous namespace

int asgpu_path_stream_read_tiles(asgpu_path_stream* ps, const uint32_t* tiles, size_t tile_count, uint32_t* accum)
{
    if (!ps || !accum || (!tiles && tile_count != 0)) return fail(ASGPU_E_INVALID, "null argument");
    const uint32_t all_tiles = ps->tiles_x * ps->tiles_y;
    if (tile_count > all_tiles) return fail(ASGPU_E_INVALID, "more tiles than the image has");
    for (size_t i = 0; i < tile_count; ++i)
        if (tiles[i] >= all_tiles) return fail(ASGPU_E_INVALID, "tile index out of range");
    if (tile_count == 0) return ASGPU_OK;
    ASGPU_CUDA(cudaSetDevice(ps->scene->device), "cudaSetDevice");
    const size_t bytes = tile_count * ps->desc.tile_size * ps->desc.tile_size * 16;
    if (bytes > ps->tile_pixels_bytes)
    {
        cudaFree(ps->tile_pixels);
        ps->tile_pixels = nullptr; ps->tile_pixels_bytes = 0;
        ASGPU_CUDA(cudaMalloc(&ps->tile_pixels, bytes), "cudaMalloc(tile pixels)");
        ps->tile_pixels_bytes = bytes;
    }
    // Legacy default stream: ordered after the render calls whatever (blocking) stream they used,
    // like the cudaDeviceSynchronize of asgpu_path_stream_read_image.
    ASGPU_CUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    ASGPU_CUDA(cudaMemcpy(ps->tiles_dev, tiles, tile_count * 4, cudaMemcpyHostToDevice), "cudaMemcpy(tiles)");
    gather_tiles_kernel<<<stage_grid(ps->scene), StageThreads>>>(ps->image, ps->tiles_dev, tile_count, ps->desc.width, ps->desc.height,
                                                                ps->desc.tile_size, ps->tiles_x, reinterpret_cast<uint4*>(ps->tile_pixels));
    ASGPU_CUDA(cudaGetLastError(), "gather_tiles_kernel");
    ASGPU_CUDA(cudaMemcpy(accum, ps->tile_pixels, bytes, cudaMemcpyDeviceToHost), "cudaMemcpy(tile pixels)");
    return ASGPU_OK;
}

int asgpu_path_stream_clear(asgpu_path_stream* ps)
{
    if (!ps) return fail(ASGPU_E_INVALID, "null path stream");
    ASGPU_CUDA(cudaSetDevice(ps->scene->device), "cudaSetDevice");
    ASGPU_CUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    ASGPU_CUDA(cudaMemset(ps->image, 0, static_cast<size_t>(ps->desc.width) * ps->desc.height * 16), "cudaMemset(image)");
    ASGPU_CUDA(cudaMemset(ps->stats_dev, 0, StatCount * 8), "cudaMemset(stats)");
    ps->wavefronts = 0;
    ps->launches = 0;
    ps->timed.clear();
    ps->events_used = 0;
    return ASGPU_OK;
}

int asgpu_path_stream_set_profiling(asgpu_path_stream* ps, int enabled)
{
    if (!ps) return fail(ASGPU_E_INVALID, "null path stream");
    ps->profiling = enabled != 0;
    return ASGPU_OK;
}

int asgpu_path_stream_get_profile(asgpu_path_stream* ps, asgpu_path_stream_profile* out)
{
    if (!ps || !out) return fail(ASGPU_E_INVALID, "null argument");
    ASGPU_CUDA(cudaSetDevice(ps->scene->device), "cudaSetDevice");
    ASGPU_CUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    std::memset(out, 0, sizeof(*out));
    for (const auto& t : ps->timed)
    {
        float ms = 0.0f;
        ASGPU_CUDA(cudaEventElapsedTime(&ms, t.second.first, t.second.second), "cudaEventElapsedTime");
        if (t.first == 0) { out->closest_ms += ms; ++out->closest_launches; }
        else if (t.first == 1) { out->probe_ms += ms; ++out->probe_launches; }
        else if (t.first == 2) out->refine_ms += ms;
        else out->stage_ms += ms;
    }
    return ASGPU_OK;
}

int asgpu_path_stream_get_stats(asgpu_path_stream* ps, asgpu_path_stream_stats* out)
{
    if (!ps || !out) return fail(ASGPU_E_INVALID, "null argument");
    ASGPU_CUDA(cudaSetDevice(ps->scene->device), "cudaSetDevice");
    ASGPU_CUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    unsigned long long v[StatCount];
    ASGPU_CUDA(cudaMemcpy(v, ps->stats_dev, sizeof(v), cudaMemcpyDeviceToHost), "cudaMemcpy(stats)");
    out->camera_rays = v[StatCamera]; out->bounce_rays = v[StatBounce]; out->probe_rays = v[StatProbe];
    out->surface_hits = v[StatHits]; out->escaped = v[StatEscaped]; out->unoccluded = v[StatUnoccluded];
    out->wavefronts = ps->wavefronts;
    out->kernel_launches = ps->launches;
    return ASGPU_OK;
}

int asgpu_path_stream_capture(asgpu_path_stream* ps, size_t max_rays)
{
    if (!ps) return fail(ASGPU_E_INVALID, "null path stream");
    ps->captured.clear();
    ps->capture_armed = max_rays != 0;
    ps->capture_budget = max_rays;
    return ASGPU_OK;
}

int asgpu_path_stream_capture_count(const asgpu_path_stream* ps) { return ps ? static_cast<int>(ps->captured.size()) : 0; }

long long asgpu_path_stream_capture_get(const asgpu_path_stream* ps, int k, int* kind, uint32_t* depth,
                                        double* org, double* dir, double* tmin, double* tmax, uint32_t* flags,
                                        uint32_t* path_ids, void* results, asgpu_parent* parents)
{
    if (!ps || k < 0 || k >= static_cast<int>(ps->captured.size())) return fail(ASGPU_E_INVALID, "no such captured wavefront");
    const Captured& c = ps->captured[k];
    const size_t n = c.tmin.size();
    if (kind) *kind = c.kind;
    if (depth) *depth = c.depth;
    if (org) std::memcpy(org, c.org.data(), n * 24);
    if (dir) std::memcpy(dir, c.dir.data(), n * 24);
    if (tmin) std::memcpy(tmin, c.tmin.data(), n * 8);
    if (tmax) std::memcpy(tmax, c.tmax.data(), n * 8);
    if (flags) std::memcpy(flags, c.flags.data(), n * 4);
    if (path_ids) std::memcpy(path_ids, c.path.data(), n * 4);
    if (results) std::memcpy(results, c.results.data(), c.results.size());
    if (parents && !c.parents.empty()) std::memcpy(parents, c.parents.data(), c.parents.size() * sizeof(asgpu_parent));
    return static_cast<long long>(n);
}

long long asgpu_path_stream_capture_get_times(const asgpu_path_stream* ps, int k, float* time_absolute, float* time_normalized)
{
    if (!ps || k < 0 || k >= static_cast<int>(ps->captured.size())) return fail(ASGPU_E_INVALID, "no such captured wavefront");
    const Captured& c = ps->captured[k];
    const size_t n = c.tmin.size();
    if (time_absolute) std::memcpy(time_absolute, c.time_absolute.data(), n * 4);
    if (time_normalized) std::memcpy(time_normalized, c.time_normalized.data(), n * 4);
    return static_cast<long long>(n);
}

}   // extern "C"
