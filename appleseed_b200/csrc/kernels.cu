//
// kernels.cu -- sm_100a kernels of the intersection engine.
//
// One kernel template, instantiated for {closest hit, any hit} x {EXACT, WIDE} x {counters}:
// persistent CTAs (a multiple of the SM count) whose warps pull 32-ray chunks from a global ray
// queue (an atomic cursor), one ray per lane, traversal state in registers and, for the wide
// traversal, a per-thread stack in shared memory laid out [entry][thread] so that a warp's
// accesses to one stack level hit 32 distinct banks.  Node and triangle records are fetched with
// 16-byte read-only loads (ld.global.nc.v4).  B200 has no RT cores and the work is not a dense
// contraction, so tensor cores are not involved; the bound is the memory system (gpu_layout.h
// states the record sizes that make up the algorithmic bytes per ray).
//
// The per-ray logic lives in traverse_core.h.
//

#include "kernels.h"
#include "traverse_core.h"

#include <cuda_runtime.h>

namespace asgpu
{

namespace
{

const int BlockThreads = 128;

struct KernelArgs
{
    SceneView           scene;
    asgpu_rays          rays;
    unsigned long long  n;
    asgpu_hit*          hits;
    uint8_t*            occluded;
    unsigned long long* queue;          // ray queue cursor
    unsigned long long* counters;       // asgpu_counters layout, or nullptr
    const uint32_t*     order;          // optional permutation: ray processed at position i is order[i]
};

__device__ __forceinline__ void store_hit(asgpu_hit* out, const SceneView& s, const Ray& ray, const Hit& hit, const bool found)
{
    // 40-byte record written as five 8-byte stores.
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(out);
    uint32_t assembly_instance = 0xFFFFFFFFu, object_instance = 0, primitive = 0, slot = 0, segment = 0, prim_type = 0;
    float u = 0.0f, v = 0.0f;
    if (found)
    {
        const uint8_t* ip = s.blob + s.items + static_cast<uint64_t>(hit.item) * sizeof(ItemRecord);
        const uint4 meta = load16(ip + 96);
        assembly_instance = meta.z;
        // read_hit_triangle_data (triangletree.cpp:1483-1499): identity from the key of the hit slot.
        const uint8_t* tp = s.blob + s.trees + static_cast<uint64_t>(meta.x) * sizeof(TreeDesc);
        const uint2 keys_off = load8(tp + offsetof(TreeDesc, keys));
        const uint64_t keys = static_cast<uint64_t>(keys_off.x) | (static_cast<uint64_t>(keys_off.y) << 32);
        const uint2 key = load8(s.blob + keys + static_cast<uint64_t>(hit.slot) * sizeof(HitKey));
        object_instance = key.x;
        primitive = key.y;
        slot = hit.slot;
        segment = hit.segment;
        prim_type = 2;
        u = hit.u; v = hit.v;
    }
    dst[0] = static_cast<unsigned long long>(__double_as_longlong(ray.tmax));
    dst[1] = static_cast<unsigned long long>(__float_as_uint(u)) | (static_cast<unsigned long long>(__float_as_uint(v)) << 32);
    dst[2] = static_cast<unsigned long long>(assembly_instance) | (static_cast<unsigned long long>(object_instance) << 32);
    dst[3] = static_cast<unsigned long long>(primitive) | (static_cast<unsigned long long>(slot) << 32);
    dst[4] = static_cast<unsigned long long>(segment) | (static_cast<unsigned long long>(prim_type) << 32);
}

// Lanes of a warp whose ray has finished are refilled from the queue as soon as at least
// RefillThreshold of them are idle (or all are), so one long ray does not hold 31 lanes hostage.
const int RefillThreshold = 8;

template <bool ANY, bool COUNT>
__device__ __forceinline__ void finish_ray(const KernelArgs& args, const unsigned long long i, const WideTraversal<ANY, COUNT>& tr)
{
    if (ANY) args.occluded[i] = tr.found() ? 1 : 0;
    else store_hit(args.hits + i, args.scene, tr.ray, tr.hit, tr.found());
}

template <bool ANY, bool WIDE, bool COUNT>
__global__ void __launch_bounds__(BlockThreads)
trace_kernel(const KernelArgs args)
{
    __shared__ uint2 wide_stack[WIDE ? WideStackSize * BlockThreads : 1];

    const unsigned lane = threadIdx.x & 31;
    Stats stats; stats.top_nodes = stats.instances = stats.nodes = stats.triangles = 0;
    unsigned rays_done = 0, hits_found = 0;

    if (WIDE)
    {
        WideTraversal<ANY, COUNT> tr;
        unsigned long long index = 0;
        bool active = false;
        bool exhausted = false;         // warp-uniform: the queue has no more rays
        uint2* stack = wide_stack + threadIdx.x;

        for (;;)
        {
            const unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
            if (idle != 0 && !exhausted && (__popc(idle) >= RefillThreshold || idle == 0xFFFFFFFFu))
            {
                // Warp-aggregated pull: one atomic for all idle lanes.
                const int leader = __ffs(idle) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(args.queue, static_cast<unsigned long long>(__popc(idle)));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                if (!active)
                {
                    const unsigned long long pos = base + __popc(idle & ((1u << lane) - 1u));
                    if (pos < args.n)
                    {
                        index = args.order ? args.order[pos] : pos;
                        tr.begin(args.scene, args.rays, index);
                        active = true;
                    }
                }
                if (base + __popc(idle) >= args.n) exhausted = true;
            }
            else if (idle == 0xFFFFFFFFu) break;        // nothing active and nothing left to fetch
            if (exhausted && __ballot_sync(0xFFFFFFFFu, active) == 0) break;

            if (active)
            {
                if (tr.step(args.scene, args.rays, index, stats, stack, BlockThreads))
                {
                    finish_ray<ANY, COUNT>(args, index, tr);
                    active = false;
                    if (COUNT) { ++rays_done; hits_found += tr.found() ? 1 : 0; }
                }
            }
        }
    }
    else
    {
        for (;;)
        {
            // Warp-level pull from the ray queue.
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(args.queue, 32ull);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (base >= args.n) break;
            const unsigned long long pos = base + lane;
            if (pos < args.n)
            {
                const unsigned long long i = args.order ? args.order[pos] : pos;
                Ray ray;
                load_ray(args.rays, i, ray);
                Hit hit;
                const bool found = exact_trace<ANY, COUNT>(args.scene, ray, hit, stats);
                if (ANY) args.occluded[i] = found ? 1 : 0;
                else store_hit(args.hits + i, args.scene, ray, hit, found);
                if (COUNT) { ++rays_done; hits_found += found ? 1 : 0; }
            }
        }
    }

    if (COUNT)
    {
        // Warp-reduce, then one atomic per counter per warp.
        unsigned vals[6] = { rays_done, stats.top_nodes, stats.instances, stats.nodes, stats.triangles, hits_found };
        #pragma unroll
        for (int k = 0; k < 6; ++k)
        {
            unsigned v = vals[k];
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            if (lane == 0 && v) atomicAdd(args.counters + k, static_cast<unsigned long long>(v));
        }
    }
}

template <bool ANY, bool WIDE, bool COUNT>
cudaError_t launch(const KernelArgs& args, const int grid, cudaStream_t stream)
{
    trace_kernel<ANY, WIDE, COUNT><<<grid, BlockThreads, 0, stream>>>(args);
    return cudaGetLastError();
}

}   // anonymous namespace

int launch_trace(
    const SceneView&    scene,
    const asgpu_rays&   rays,
    const size_t        n,
    asgpu_hit*          hits,
    uint8_t*            occluded,
    const bool          any_hit,
    const bool          wide,
    unsigned long long* queue,
    unsigned long long* counters,
    const uint32_t*     order,
    const int           sm_count,
    void*               stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    KernelArgs args;
    args.scene = scene;
    args.rays = rays;
    args.n = n;
    args.hits = hits;
    args.occluded = occluded;
    args.queue = queue;
    args.counters = counters;
    args.order = order;

    cudaError_t err = cudaMemsetAsync(queue, 0, sizeof(unsigned long long), stream);
    if (err != cudaSuccess) return static_cast<int>(err);

    // Persistent grid: a multiple of the SM count, no larger than the work.
    int blocks_per_sm = 0;
    const void* fn = nullptr;
    #define ASGPU_PICK(A, W, C) (const void*)trace_kernel<A, W, C>
    const bool count = counters != nullptr;
    if (any_hit) fn = wide ? (count ? ASGPU_PICK(true, true, true) : ASGPU_PICK(true, true, false))
                           : (count ? ASGPU_PICK(true, false, true) : ASGPU_PICK(true, false, false));
    else fn = wide ? (count ? ASGPU_PICK(false, true, true) : ASGPU_PICK(false, true, false))
                   : (count ? ASGPU_PICK(false, false, true) : ASGPU_PICK(false, false, false));
    #undef ASGPU_PICK
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, fn, BlockThreads, 0);
    if (err != cudaSuccess) return static_cast<int>(err);
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    long long grid = static_cast<long long>(sm_count) * blocks_per_sm;
    const long long needed = static_cast<long long>((n + BlockThreads - 1) / BlockThreads);
    if (grid > needed) grid = needed;
    if (grid < 1) grid = 1;

    if (any_hit)
    {
        if (wide) err = count ? launch<true, true, true>(args, (int)grid, stream) : launch<true, true, false>(args, (int)grid, stream);
        else      err = count ? launch<true, false, true>(args, (int)grid, stream) : launch<true, false, false>(args, (int)grid, stream);
    }
    else
    {
        if (wide) err = count ? launch<false, true, true>(args, (int)grid, stream) : launch<false, true, false>(args, (int)grid, stream);
        else      err = count ? launch<false, false, true>(args, (int)grid, stream) : launch<false, false, false>(args, (int)grid, stream);
    }
    return static_cast<int>(err);
}

}   // namespace asgpu
