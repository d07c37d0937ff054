//
// kernels.cu -- sm_100a kernels of the intersection engine.
//
// Two kernel templates, each instantiated for {closest hit, any hit} x {counters}: persistent
// CTAs (a multiple of the SM count) whose warps pull rays from a global ray queue (an atomic
// cursor), one ray per lane.  Node and triangle records are fetched with 16-byte read-only loads
// (ld.global.nc.v4).  B200 has no RT cores and the work is not a dense contraction, so tensor
// cores are not involved (gpu_layout.h states the record sizes that make up the algorithmic bytes
// per ray).  Measured limiter: instruction issue, not memory (profiles/README.md).
//
//   trace_kernel   EXACT layout: the reference's traversal operation for operation, one ray per
//                  lane start to finish (the arbiter path).
//   wide_kernel    WIDE layout, the throughput path.  A warp advances its 32 rays together: per
//                  loop iteration each lane tests up to 3-4 8-wide nodes (fp32 interval arithmetic,
//                  stopping at the first node with leaf triangles), lanes that reached an assembly
//                  instance enter it together, then the warp gathers the (ray, triangle)
//                  candidates of all its lanes into a shared-memory queue and tests them 32 at a
//                  time with the exact fp64 test, whichever lane a candidate came from.  The fp64
//                  ray, the hit record and the traversal stack ([entry][thread], conflict free) live
//                  in shared memory; registers only hold the fp32 interval form of the ray and the
//                  traversal cursor.  Instantiations: {closest, any hit} x {counters} x stack depth
//                  {16, 24, 64} x LEAN (one static instance) x FILTERS (alpha masks, closest only).
//
// The per-ray building blocks live in traverse_core.h.
//

#include "kernels.h"
#include "traverse_core.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <mutex>

namespace asgpu
{

namespace
{

const int BlockThreads = 128;
const int LaneProfileWords = 16;      // ... followed by two lane-profile banks (closest, any hit) of the wide kernels' COUNT instantiations
const int CounterBankWords = 8;       // counters: one asgpu_counters-shaped bank for closest-hit launches, one for any-hit launches

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
    const unsigned long long* n_dev;    // optional: the ray count lives in device memory (wavefront queues)
    uint32_t            raw_item;           // hit records carry the ItemRecord index instead of the caller's instance id
    const asgpu_parent* parents;        // optional parent shading point per ray (assemblytree.cpp:565-576)
    int                 refill_threshold;   // idle lanes that trigger a pull from the ray queue
    int                 flush_threshold;    // queued candidates that trigger a test batch
    int                 stall_threshold;    // lanes idle or waiting for the queue that trigger one
    uint32_t            unit_bits;          // bits of 1.0f (see byte_to_unit)
    int                 prefetch;           // start the fetch of the next node as soon as it is chosen
    int                 max_steps;          // node tests per lane and loop iteration
    int                 step_threshold;     // lanes that must be able to advance for another step
    int                 enter_late;         // enter assembly instances once per iteration, all lanes together
    int                 enter_threshold;    // ... once this many lanes want to
    int                 enter_min_busy;     // ... or fewer lanes than this can advance without entering
};

__device__ __forceinline__ void store_hit(asgpu_hit* out, const SceneView& s, const double t, const Hit& hit, const bool found, const bool raw_item)
{
    // 40-byte record written as five 8-byte stores.
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(out);
    uint32_t assembly_instance = 0xFFFFFFFFu, object_instance = 0, primitive = 0, slot = 0, segment = 0, prim_type = 0;
    float u = 0.0f, v = 0.0f;
    if (found)
    {
        const uint8_t* ip = s.blob + s.items + static_cast<uint64_t>(hit.item) * sizeof(ItemRecord);
        const uint4 meta = load16(ip + 96);
        assembly_instance = raw_item ? hit.item : meta.z;
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
    dst[0] = static_cast<unsigned long long>(__double_as_longlong(t));
    dst[1] = static_cast<unsigned long long>(__float_as_uint(u)) | (static_cast<unsigned long long>(__float_as_uint(v)) << 32);
    dst[2] = static_cast<unsigned long long>(assembly_instance) | (static_cast<unsigned long long>(object_instance) << 32);
    dst[3] = static_cast<unsigned long long>(primitive) | (static_cast<unsigned long long>(slot) << 32);
    dst[4] = static_cast<unsigned long long>(segment) | (static_cast<unsigned long long>(prim_type) << 32);
}

__device__ __forceinline__ void flush_counters(unsigned long long* counters, const unsigned lane, const unsigned rays_done, const Stats& stats, const unsigned hits_found)
{
    // Warp-reduce, then one atomic per counter per warp.
    unsigned vals[6] = { rays_done, stats.top_nodes, stats.instances, stats.nodes, stats.triangles, hits_found };
    #pragma unroll
    for (int k = 0; k < 6; ++k)
    {
        unsigned v = vals[k];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (lane == 0 && v) atomicAdd(counters + k, static_cast<unsigned long long>(v));
    }
}

// ------------------------------------------------------------------------------------------
// EXACT layout.
// ------------------------------------------------------------------------------------------

template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(BlockThreads)
trace_kernel(const KernelArgs args)
{
    const unsigned lane = threadIdx.x & 31;
    // A device-side count (wavefront queues) is clamped to the capacity of the arrays: producers
    // keep counting past it when a queue overflows.
    const unsigned long long n = args.n_dev ? min(*args.n_dev, args.n) : args.n;
    Stats stats; stats.top_nodes = stats.instances = stats.nodes = stats.triangles = 0;
    unsigned rays_done = 0, hits_found = 0;

    for (;;)
    {
        // Warp-level pull from the ray queue.
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(args.queue, 32ull);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const unsigned long long pos = base + lane;
        if (pos < n)
        {
            const unsigned long long i = args.order ? args.order[pos] : pos;
            Ray ray;
            load_ray(args.rays, i, ray);
            Hit hit;
            const bool found = exact_trace<ANY, COUNT>(args.scene, ray, hit, stats,
                                                       args.parents ? reinterpret_cast<const uint8_t*>(args.parents + i) : nullptr);
            if (ANY) args.occluded[i] = found ? 1 : 0;
            else store_hit(args.hits + i, args.scene, ray.tmax, hit, found, args.raw_item != 0);
            if (COUNT) { ++rays_done; hits_found += found ? 1 : 0; }
        }
    }

    if (COUNT) flush_counters(args.counters + (ANY ? CounterBankWords : 0), lane, rays_done, stats, hits_found);
}

// ------------------------------------------------------------------------------------------
// WIDE layout.
// ------------------------------------------------------------------------------------------

// Lanes of a warp whose ray has finished are refilled from the queue as soon as at least
// RefillThreshold of them are idle (or all are), so one long ray does not hold 31 lanes hostage.
const int QueueSlots = 128;         // per warp: < 32 left over + one step's candidates (more are pushed in rounds)
const uint32_t None = 0xFFFFFFFFu;

// Shared memory of one CTA, every per-thread array laid out [component][thread].  Arrays that only
// the MOTION / EXTRAS instantiations use shrink to one element elsewhere: what the CTAs do not take
// as shared memory is the SM's L1 cache, which the node fetches want.
template <int STACK, bool MOTION, bool EXTRAS, bool LEAN, bool ANY>
struct WideShared
{
    uint2               stack[STACK][BlockThreads];
    double              ray[8][BlockThreads];           // current-space org, dir; tmin; tmax (= closest t so far)
    unsigned long long  tri_base[BlockThreads];         // blob offset of the current tree's wide triangle records
    unsigned long long  pose_base[MOTION ? BlockThreads : 1];   // blob offset of its pose pool
    unsigned long long  best[ANY ? 1 : BlockThreads];   // ordered key of the nearest hit of the running batch
    unsigned long long  queue[BlockThreads / 32][QueueSlots];
    uint32_t            flags[BlockThreads];
    float               time_n[MOTION ? BlockThreads : 1];
    float               hit_u[ANY ? 1 : BlockThreads], hit_v[ANY ? 1 : BlockThreads];
    uint32_t            hit_slot[ANY ? 1 : BlockThreads], hit_item[BlockThreads], hit_segment[MOTION && !ANY ? BlockThreads : 1];
    uint32_t            cur_item[BlockThreads];
    float               world_ray[12][LEAN ? 1 : BlockThreads];    // world-space interval ray (o_lo, o_hi, rn, rf), saved while inside an instance
    uint32_t            world_oct[LEAN ? 1 : BlockThreads];
    uint32_t            key_base[ANY ? 1 : BlockThreads];       // blob offset / SectionAlign of the current tree's hit keys (closest hit: prefetch of the winner's key)
    unsigned long long  filter_tree[EXTRAS ? BlockThreads : 1]; // blob offset of the current tree's TreeDesc when it has intersection filters, else 0
};

// Starts the fetch of one wide node (80 bytes, 16-byte aligned: at most two 128-byte lines) into L1.
__device__ __forceinline__ void prefetch_node(const uint8_t* p)
{
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p + 64));
}

// Monotone map from double to uint64 (for atomicMin on t).
__device__ __forceinline__ unsigned long long ordered_key(const double t)
{
    const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(t));
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// Out of line on purpose (rare path, must not cost the traversal loop registers): the ray in the
// space of an ANIMATED assembly instance.  io[0..5] = world org, dir in; instance org, dir out.
__device__ __noinline__ void animated_instance_ray(const uint8_t* blob, const uint8_t* item, const uint32_t key_count, const float time_absolute, double* io)
{
    double lorg[3], ldir[3];
    instance_org_dir_at(blob, item, key_count, time_absolute, io, io + 3, lorg, ldir);
    io[0] = lorg[0]; io[1] = lorg[1]; io[2] = lorg[2];
    io[3] = ldir[0]; io[4] = ldir[1]; io[5] = ldir[2];
}

// Out of line on purpose: the filter is a rare path (scenes with cut-out geometry only) and must
// not cost the traversal loop registers.
__device__ __noinline__ bool filter_accept_call(const uint8_t* blob, const uint8_t* tree, const uint32_t slot, const double u, const double v)
{
    return filter_accept(blob, tree, slot, u, v);
}

// Tests `count` (<= 32) queued candidates, one per lane.  Entry = (index of the triangle record in
// the source lane's current tree) << 5 | source lane.  Closest hit: the nearest accepted candidate of each source lane
// updates that lane's tmax and hit record (ties: lowest queue position).  Any hit: marks the lane.
template <bool ANY, bool COUNT, bool EXTRAS, bool MOTION, typename Shared>
__device__ __forceinline__ void test_candidates(
    Shared& sm, const uint8_t* blob, const unsigned long long* entries, const unsigned count,
    const unsigned lane, const unsigned warp_thread0, Stats& stats)
{
    bool hit = false;
    double t = 0.0, u = 0.0, v = 0.0;
    uint32_t slot = 0, segment = 0;
    unsigned st = warp_thread0;
    if (!ANY) sm.best[ANY ? 0 : warp_thread0 + lane] = ~0ull;
    if (lane < count)
    {
        const unsigned long long e = entries[lane];
        st = warp_thread0 + static_cast<unsigned>(e & 31u);
        Ray ray;
        #pragma unroll
        for (int k = 0; k < 3; ++k) { ray.org[k] = sm.ray[k][st]; ray.dir[k] = sm.ray[3 + k][st]; }
        ray.tmin = sm.ray[6][st];
        ray.tmax = sm.ray[7][st];
        ray.flags = sm.flags[st];
        ray.time_normalized = MOTION ? sm.time_n[MOTION ? st : 0] : 0.0f;
        ray.time_absolute = 0.0f;
        if (COUNT) ++stats.triangles;
        TriD tri;
        if (fetch_triangle<ANY, MOTION>(blob + sm.tri_base[st] + (e >> 5) * sizeof(TriRecord), MOTION ? blob + sm.pose_base[MOTION ? st : 0] : nullptr, ray, tri, slot, segment))
            hit = mt_test<!ANY>(tri, ray, t, u, v);
        // Optionally filter intersections (triangletree.cpp:1404-1411; closest hit only).
        if (!ANY && EXTRAS && hit && sm.filter_tree[EXTRAS ? st : 0] != 0) hit = filter_accept_call(blob, blob + sm.filter_tree[EXTRAS ? st : 0], slot, u, v);
    }
    if (ANY)
    {
        if (hit) sm.hit_item[st] = sm.cur_item[st];
        __syncwarp();
        return;
    }
    const unsigned hits = __ballot_sync(0xFFFFFFFFu, hit);
    if (hits == 0) return;
    __syncwarp();                                   // best[] initialised
    const unsigned long long key = ordered_key(t);
    if (hit) atomicMin(&sm.best[ANY ? 0 : st], key);
    __syncwarp();
    const bool win = hit && sm.best[ANY ? 0 : st] == key;
    const unsigned winners = __ballot_sync(0xFFFFFFFFu, win);
    if (win)
    {
        const unsigned peers = __match_any_sync(winners, st);
        if (lane == static_cast<unsigned>(__ffs(peers) - 1))
        {
            sm.ray[7][st] = t;
            sm.hit_u[ANY ? 0 : st] = static_cast<float>(u);
            sm.hit_v[ANY ? 0 : st] = static_cast<float>(v);
            sm.hit_slot[ANY ? 0 : st] = slot;
            if (MOTION) sm.hit_segment[MOTION && !ANY ? st : 0] = segment;
            sm.hit_item[st] = sm.cur_item[st];
            // The identity of the hit (store_hit reads its HitKey when the ray retires: one random
            // 8-byte read at the end of a chain of dependent loads) starts its way up from DRAM now.
            asm volatile("prefetch.global.L2 [%0];" :: "l"(blob + static_cast<uint64_t>(sm.key_base[ANY ? 0 : st]) * SectionAlign + static_cast<uint64_t>(slot) * sizeof(HitKey)));
        }
    }
    __syncwarp();
}

// Back from an instance: the world-space interval ray again, as it was saved at instance entry (12
// floats and the octant in shared memory); only the parameter interval is new (tmax has shrunk).
template <typename Shared>
__device__ __forceinline__ void restore_world_ray(const double tmin, const double tmax, const Shared& sm, const unsigned tid, WideRay& w)
{
    w.shift = tmin < 0.0 ? tmin : 0.0;
    w.oct = sm.world_oct[tid];
    #pragma unroll
    for (int a = 0; a < 3; ++a)
    {
        w.o_lo[a] = sm.world_ray[a][tid];
        w.o_hi[a] = sm.world_ray[3 + a][tid];
        w.rn[a] = sm.world_ray[6 + a][tid];
        w.rf[a] = sm.world_ray[9 + a][tid];
    }
    w.tmin_f = d2f_dn(dsub(tmin, w.shift));
    w.tmax_f = d2f_up(dsub(tmax, w.shift));
}

// LEAN = the scene has one assembly instance and no moving triangles (e.g. C2): nothing ever comes
// back to world space and no tree has time slices, so the instantiation drops the batched instance
// entry, the saved world-space ray and the indirection to the child planes (measured: 3-5 %).
// EXTRAS = the scene has intersection filters (cut-out geometry) or animated assembly instances:
// only then does the instantiation contain the alpha-mask lookup (closest hit) and the evaluation
// of a transform sequence at the ray time (both cost registers: -3 % when compiled into the common
// kernel).
// MOTION = some tree of the scene has moving triangles: only then do the instantiations contain the
// pose interpolation of the triangle stage and the time slices of the child planes (their registers
// cost the static scenes nothing: C3 +2 % closest hit, +4 % probes).
template <bool ANY, bool COUNT, int STACK, int MINB, bool LEAN, bool EXTRAS, bool MOTION>
__global__ void __launch_bounds__(BlockThreads, MINB)
wide_kernel(const KernelArgs args)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef WideShared<STACK, MOTION, EXTRAS, LEAN, ANY> Shared;
    Shared& sm = *reinterpret_cast<Shared*>(smem_raw);

    const unsigned tid = threadIdx.x;
    const unsigned lane = tid & 31;
    const unsigned warp_thread0 = tid - lane;
    const unsigned lanes_below = (1u << lane) - 1u;
    unsigned long long* queue = sm.queue[tid >> 5];
    const SceneView& s = args.scene;
    const uint8_t* blob = s.blob;
    const unsigned long long n = args.n_dev ? min(*args.n_dev, args.n) : args.n;     // clamped: see trace_kernel
    uint2* stack = &sm.stack[0][tid];
    const uint32_t stride = BlockThreads;

    Stats stats; stats.top_nodes = stats.instances = stats.nodes = stats.triangles = 0;
    unsigned rays_done = 0, hits_found = 0;
    // COUNT only: where the lane slots of the node-test rounds go (asgpu_get_lane_profile).
    unsigned prof[LaneProfileWords];
    #pragma unroll
    for (int k = 0; k < LaneProfileWords; ++k) prof[k] = 0;

    // Per-lane traversal cursor.
    WideRay w;
    const uint8_t* wnodes = blob;
    const uint8_t* qbase = blob;        // quantised child planes of node i: qbase + i * qstride (the node's own, or its time slice)
    uint32_t qstride = sizeof(WNode);
    uint2 ngroup, tgroup;               // pending internal children / pending instances (world space)
    ngroup.x = ngroup.y = tgroup.x = tgroup.y = 0;
    uint32_t fetch = None;              // wide node to test next
    uint32_t sp = 0;
    auto push = [&](const uint2 v) { stack[sp * stride] = v; ++sp; };
    auto pop = [&]() -> uint2 { --sp; return stack[sp * stride]; };
    uint32_t cur_item = None;           // None while in world space
    unsigned long long index = 0;
    bool active = false;                // this lane owns a ray
    bool traversed = false;             // ... whose traversal is complete (it may still wait for queued candidates)
    bool waiting = false;               // ... and which has candidates in the queue
    bool exhausted = false;             // warp-uniform: the ray queue has no more rays
    unsigned queued = 0;                // warp-uniform: candidates waiting in the queue

    // Entering an assembly instance (AssemblyLeafVisitor::visit, assemblytree.cpp:604-744): ~250
    // instructions of ray set-up.  Scenes with one instance do it in line (every ray enters right after
    // the root test, so the lanes of a refill do it together anyway); instanced scenes collect the
    // lanes that reach an instance during an iteration and enter together after the last round.
    auto enter_instance = [&]()
    {
        // Next assembly instance: AssemblyLeafVisitor::visit (assemblytree.cpp:604-744).
        const int bit = high_bit(tgroup.y);
        tgroup.y &= ~(1u << bit);
        const uint32_t item = load4(blob + s.top_witems + static_cast<uint64_t>(tgroup.x + bit) * 4);
        const uint8_t* ip = blob + s.items + static_cast<uint64_t>(item) * sizeof(ItemRecord);
        const uint4 meta = load16(ip + 96);
        if ((meta.y & sm.flags[tid]) && meta.x != None)
        {
            if (COUNT) ++stats.instances;
            // Save the world-space cursor, then descend.  When nothing is left to do in
            // world space there is nothing to come back to: no sentinel, the ray ends
            // with the instance.
            if (ngroup.y & 0xFF000000u) push(ngroup);
            if (tgroup.y) push(tgroup);
            if (sp != 0)
            {
                uint2 sentinel; sentinel.x = None; sentinel.y = 0;
                push(sentinel);
            }
            Ray world;
            load_ray_org_dir(args.rays, index, world);
            double lorg[3], ldir[3];
            if (EXTRAS && meta.w >= 2)
            {
                // Animated instance: its transform at the ray's absolute time (assemblytree.cpp:635-639).
                const float time_absolute = args.rays.time_absolute ? __ldg(args.rays.time_absolute + index) : 0.0f;
                double io[6] = { world.org[0], world.org[1], world.org[2], world.dir[0], world.dir[1], world.dir[2] };
                animated_instance_ray(blob, ip, meta.w, time_absolute, io);
                #pragma unroll
                for (int k = 0; k < 3; ++k) { lorg[k] = io[k]; ldir[k] = io[3 + k]; }
            }
            else instance_org_dir(ip, world.org, world.dir, lorg, ldir);
            if (args.parents) parent_origin(reinterpret_cast<const uint8_t*>(args.parents + index), meta.z, ldir, lorg);
            #pragma unroll
            for (int k = 0; k < 3; ++k) { sm.ray[k][tid] = lorg[k]; sm.ray[3 + k][tid] = ldir[k]; }
            if (!LEAN && sp != 0)
            {
                // Something is left to do in world space: keep the expensive part of the world ray.
                #pragma unroll
                for (int k = 0; k < 3; ++k)
                {
                    sm.world_ray[k][tid] = w.o_lo[k]; sm.world_ray[3 + k][tid] = w.o_hi[k];
                    sm.world_ray[6 + k][tid] = w.rn[k]; sm.world_ray[9 + k][tid] = w.rf[k];
                }
                sm.world_oct[tid] = w.oct;
            }
            make_wide_ray(lorg, ldir, sm.ray[6][tid], sm.ray[7][tid], w);
            const uint8_t* tp = blob + s.trees + static_cast<uint64_t>(meta.x) * sizeof(TreeDesc);
            const uint2 o_nodes = load8(tp + offsetof(TreeDesc, wnodes));
            const uint2 o_tris = load8(tp + offsetof(TreeDesc, wtris));
            const uint2 o_poses = load8(tp + offsetof(TreeDesc, poses));
            const uint32_t wnode_count = load4(tp + offsetof(TreeDesc, wnode_count));
            const uint32_t wslice_count = load4(tp + offsetof(TreeDesc, wslice_count));
            wnodes = blob + (static_cast<uint64_t>(o_nodes.x) | (static_cast<uint64_t>(o_nodes.y) << 32));
            if (MOTION && wslice_count != 0)
            {
                // Tree with moving triangles: child planes of this ray's time slice.
                const uint2 o_slices = load8(tp + offsetof(TreeDesc, wslices));
                qbase = blob + (static_cast<uint64_t>(o_slices.x) | (static_cast<uint64_t>(o_slices.y) << 32))
                      + static_cast<uint64_t>(time_slice(sm.time_n[MOTION ? tid : 0], wslice_count)) * sizeof(WSlice);
                qstride = wslice_count * static_cast<uint32_t>(sizeof(WSlice));
            }
            else { qbase = wnodes + 32; qstride = sizeof(WNode); }
            sm.tri_base[tid] = static_cast<uint64_t>(o_tris.x) | (static_cast<uint64_t>(o_tris.y) << 32);
            if (!ANY)
            {
                const uint2 o_keys = load8(tp + offsetof(TreeDesc, keys));
                sm.key_base[ANY ? 0 : tid] = static_cast<uint32_t>((static_cast<uint64_t>(o_keys.x) | (static_cast<uint64_t>(o_keys.y) << 32)) / SectionAlign);   // sections are SectionAlign-aligned, blobs < 1 TB
            }
            if (MOTION) sm.pose_base[MOTION ? tid : 0] = static_cast<uint64_t>(o_poses.x) | (static_cast<uint64_t>(o_poses.y) << 32);
            sm.cur_item[tid] = item;
            if (EXTRAS)
                sm.filter_tree[EXTRAS ? tid : 0] = load4(tp + offsetof(TreeDesc, filter_count)) != 0 ? s.trees + static_cast<uint64_t>(meta.x) * sizeof(TreeDesc) : 0ull;
            cur_item = item;
            ngroup.y = 0; tgroup.y = 0;
            if (wnode_count != 0) fetch = 0;
        }
    };

    for (;;)
    {
        // ---- refill idle lanes from the ray queue (one atomic per warp) -------------------------
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (idle != 0 && !exhausted && (__popc(idle) >= args.refill_threshold || idle == 0xFFFFFFFFu))
        {
            const int leader = __ffs(idle) - 1;
            if (COUNT) { prof[11] += 1; prof[12] += __popc(idle); }
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(args.queue, static_cast<unsigned long long>(__popc(idle)));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (!active)
            {
                const unsigned long long pos = base + __popc(idle & lanes_below);
                if (pos < n)
                {
                    index = args.order ? args.order[pos] : pos;
                    Ray ray;
                    load_ray(args.rays, index, ray);
                    sm.ray[6][tid] = ray.tmin;
                    sm.ray[7][tid] = ray.tmax;
                    sm.flags[tid] = ray.flags;
                    if (MOTION) sm.time_n[MOTION ? tid : 0] = ray.time_normalized;
                    sm.hit_item[tid] = None;
                    if (!ANY)
                    {
                        sm.hit_slot[ANY ? 0 : tid] = 0; if (MOTION) sm.hit_segment[MOTION && !ANY ? tid : 0] = 0;
                        sm.hit_u[ANY ? 0 : tid] = 0.0f; sm.hit_v[ANY ? 0 : tid] = 0.0f;
                    }
                    sm.cur_item[tid] = None;
                    make_wide_ray(ray.org, ray.dir, ray.tmin, ray.tmax, w);
                    wnodes = blob + s.top_wnodes;
                    qbase = wnodes + 32; qstride = sizeof(WNode);
                    ngroup.y = 0; tgroup.y = 0;
                    fetch = s.top_wnode_count != 0 ? 0u : None;
                    sp = 0;
                    cur_item = None;
                    active = true; traversed = false; waiting = false;
                }
            }
            if (base + __popc(idle) >= n) exhausted = true;
        }
        else if (idle == 0xFFFFFFFFu) break;            // nothing active and nothing left to fetch
        if (exhausted && __ballot_sync(0xFFFFFFFFu, active) == 0) break;

        // ---- one traversal step per active lane ---------------------------------------------------
        // Several steps per iteration: a lane keeps walking internal nodes until it finds leaf
        // triangles (most node tests find none), so the warp-level bookkeeping below (refill,
        // gather, flush, pick-up) is paid once per `max_steps` node tests instead of once per test.
        // The extra rounds only run while enough lanes can still advance.
        if (COUNT) prof[8] += 1;                        // loop iterations
        uint32_t pending = 0, tri_first = 0;            // leaf triangles found by this iteration's node tests
        bool held = false;                              // this lane cannot advance before the queue drains
        bool want_enter = false;                        // this lane's next move is to enter an assembly instance
        for (int round = 0; ; ++round)
        {
        if (COUNT)
        {
            // One round = 32 lane slots: how many test a node, and why the others do not.
            const bool can = active && !traversed && !held && !want_enter && pending == 0;
            prof[0] += 1;
            prof[1] += __popc(__ballot_sync(0xFFFFFFFFu, can && fetch != None));
            prof[2] += __popc(__ballot_sync(0xFFFFFFFFu, !active));
            prof[3] += __popc(__ballot_sync(0xFFFFFFFFu, active && traversed));
            prof[4] += __popc(__ballot_sync(0xFFFFFFFFu, active && !traversed && held));
            prof[5] += __popc(__ballot_sync(0xFFFFFFFFu, active && !traversed && !held && want_enter));
            prof[6] += __popc(__ballot_sync(0xFFFFFFFFu, active && !traversed && !held && !want_enter && pending != 0));
            prof[7] += __popc(__ballot_sync(0xFFFFFFFFu, can && fetch == None));
        }
        if (active && !traversed && !held && !want_enter && pending == 0)
        {
            if (fetch != None)
            {
                if (COUNT) { if (cur_item != None) ++stats.nodes; else ++stats.top_nodes; }
                if (ngroup.y & 0xFF000000u) push(ngroup);
                uint32_t child_base, tri_base, nmask, tmask;
                const uint8_t* np = wnodes + static_cast<uint64_t>(fetch) * sizeof(WNode);
                wide_node_test(np, !MOTION ? np + 32 : qbase + static_cast<uint64_t>(fetch) * qstride, w, args.unit_bits, child_base, tri_base, nmask, tmask);
                ngroup.x = child_base; ngroup.y = nmask;
                if (cur_item != None) { pending = tmask; tri_first = tri_base; }
                else { tgroup.x = tri_base; tgroup.y = tmask; }
                fetch = None;
            }

            // Choose what comes next (and start fetching it: the node is in L1 by the time it is tested).
            if (fetch == None)
            {
                if (tgroup.y && waiting)
                {
                    // Queued candidates still refer to this lane's instance-space ray: it must not
                    // be replaced by the next instance's before they are tested.
                    held = true;
                }
                else if (tgroup.y)
                {
                    if (!LEAN && args.enter_late) want_enter = true;    // performed below, with the other lanes of the warp that got here
                    else enter_instance();
                }
                else
                {
                    if (!(ngroup.y & 0xFF000000u))
                    {
                        if (sp == 0) traversed = true;
                        else
                        {
                            uint2 top = pop();
                            if (top.x == None && top.y == 0)
                            {
                                // Back to world space: the world ray comes from the ray arrays again.
                                // (Candidates of the instance still in the queue carry all they need.)
                                if (LEAN)
                                {
                                    Ray world;
                                    load_ray_org_dir(args.rays, index, world);
                                    make_wide_ray(world.org, world.dir, sm.ray[6][tid], sm.ray[7][tid], w);
                                }
                                else restore_world_ray(sm.ray[6][tid], sm.ray[7][tid], sm, tid, w);
                                wnodes = blob + s.top_wnodes;
                                qbase = wnodes + 32; qstride = sizeof(WNode);
                                cur_item = None;
                                top = pop();    // what was left to do in world space (there is no sentinel without it)
                            }
                            if (top.y & 0xFF000000u) ngroup = top;
                            else tgroup = top;
                        }
                    }
                    if (ngroup.y & 0xFF000000u)
                    {
                        // Next internal child of the current group, nearest octant slot first.
                        const int bit = high_bit(ngroup.y);
                        ngroup.y &= ~(1u << bit);
                        const uint32_t k = static_cast<uint32_t>(bit - 24) ^ (7 - (w.oct & 7));
                        fetch = ngroup.x + popc(ngroup.y & 0xFFu & ((1u << k) - 1u));
                    }
                    else if (!LEAN && tgroup.y)
                    {
                        // An instance group came off the stack: decided now, not one idle round later
                        // (the entry itself always goes through the batched path below).  Leaf triangles
                        // found by this very round's node test are not in the queue yet but refer to the
                        // instance the lane is leaving, just like queued ones: the entry must wait for both.
                        if (waiting || pending != 0) held = true;
                        else want_enter = true;
                    }
                }
            }
            if (args.prefetch && fetch != None) prefetch_node(wnodes + static_cast<uint64_t>(fetch) * sizeof(WNode));
        }
        if (round + 1 >= args.max_steps) break;
        if (__popc(__ballot_sync(0xFFFFFFFFu, active && !traversed && !held && !want_enter && pending == 0)) < args.step_threshold) break;
        }

        // ---- enter assembly instances: every lane that reached one in this iteration, together ------
        if (!LEAN)
        {
            const unsigned want = __ballot_sync(0xFFFFFFFFu, want_enter);
            if (want != 0)
            {
                // Enter when enough lanes wait for it, or when too few lanes could go on without.
                bool go = __popc(want) >= args.enter_threshold;
                if (!go) go = __popc(__ballot_sync(0xFFFFFFFFu, active && !traversed && !held && !want_enter)) < args.enter_min_busy;
                if (COUNT && go) { prof[9] += 1; prof[10] += __popc(want); }
                if (go && want_enter) enter_instance();
            }
        }

        // ---- gather this step's triangle candidates -------------------------------------------------
        bool tested = false;                            // warp-uniform: a batch ran in this iteration
        if (__ballot_sync(0xFFFFFFFFu, pending != 0) != 0)
        {
            // Exclusive prefix sum of the lanes' candidate counts = their places in the queue.
            const unsigned mine = __popc(pending);
            unsigned upto = mine;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const unsigned v = __shfl_up_sync(0xFFFFFFFFu, upto, o);
                if (lane >= o) upto += v;
            }
            const unsigned total = __shfl_sync(0xFFFFFFFFu, upto, 31);
            if (mine != 0) waiting = true;
            __syncwarp();                               // the ray / tree data of new lanes is visible
            if (queued + total <= QueueSlots)
            {
                unsigned pos = queued + upto - mine;
                const unsigned long long base = (static_cast<unsigned long long>(tri_first) << 5) | lane;
                while (pending != 0)
                {
                    const int bit = high_bit(pending);
                    pending &= ~(1u << bit);
                    queue[pos++] = base + (static_cast<unsigned long long>(bit) << 5);
                }
                queued += total;
                __syncwarp();
                while (queued >= 32)
                {
                    queued -= 32;
                    test_candidates<ANY, COUNT, EXTRAS, MOTION>(sm, blob, queue + queued, 32, lane, warp_thread0, stats);
                    tested = true;
                }
            }
            else
            {
                // More candidates than the queue holds (rare): one per lane and round.
                unsigned pushers = __ballot_sync(0xFFFFFFFFu, pending != 0);
                do
                {
                    if (pending != 0)
                    {
                        const int bit = high_bit(pending);
                        pending &= ~(1u << bit);
                        queue[queued + __popc(pushers & lanes_below)] = (static_cast<unsigned long long>(tri_first + bit) << 5) | lane;
                    }
                    queued += __popc(pushers);
                    __syncwarp();
                    if (queued >= 32)
                    {
                        queued -= 32;
                        test_candidates<ANY, COUNT, EXTRAS, MOTION>(sm, blob, queue + queued, 32, lane, warp_thread0, stats);
                        tested = true;
                    }
                    pushers = __ballot_sync(0xFFFFFFFFu, pending != 0);
                } while (pushers != 0);
            }
        }

        // ---- flush the queue when it is full enough or too many lanes wait for it -------------------
        if (queued != 0)
        {
            const unsigned stalled = __ballot_sync(0xFFFFFFFFu, !active || held || (traversed && waiting));
            if (queued >= args.flush_threshold || __popc(stalled) >= args.stall_threshold)
            {
                test_candidates<ANY, COUNT, EXTRAS, MOTION>(sm, blob, queue, queued, lane, warp_thread0, stats);
                queued = 0;
                tested = true;
            }
        }
        if (queued == 0) waiting = false;
        // ---- pick up the results, retire finished rays ---------------------------------------------
        if (active)
        {
            if (tested)
            {
                if (ANY) { if (sm.hit_item[tid] != None) traversed = true; }    // retires once its queued candidates are gone
                else
                {
                    const double tmin = sm.ray[6][tid];
                    w.tmax_f = d2f_up(dsub(sm.ray[7][tid], tmin < 0.0 ? tmin : 0.0));
                }
            }
            if (traversed && !waiting)
            {
                const bool found = sm.hit_item[tid] != None;
                if (ANY) args.occluded[index] = found ? 1 : 0;
                else
                {
                    Hit hit;
                    hit.u = sm.hit_u[ANY ? 0 : tid]; hit.v = sm.hit_v[ANY ? 0 : tid];
                    hit.item = sm.hit_item[tid]; hit.slot = sm.hit_slot[ANY ? 0 : tid]; hit.segment = MOTION ? sm.hit_segment[MOTION && !ANY ? tid : 0] : 0u;
                    store_hit(args.hits + index, s, sm.ray[7][tid], hit, found, args.raw_item != 0);
                }
                if (COUNT) { ++rays_done; hits_found += found ? 1 : 0; }
                active = false;
            }
        }
    }

    if (COUNT)
    {
        flush_counters(args.counters + (ANY ? CounterBankWords : 0), lane, rays_done, stats, hits_found);
        if (lane == 0)
        {
            #pragma unroll
            for (int k = 0; k < LaneProfileWords; ++k)
                if (prof[k]) atomicAdd(args.counters + 2 * CounterBankWords + (ANY ? LaneProfileWords : 0) + k, static_cast<unsigned long long>(prof[k]));
        }
    }
}

// Scheduling knobs of the wide kernel.  Defaults are the measured optima (profiles/README.md, sweeps
// r1/quick_steps*.log): scenes with a single assembly instance pay the ray set-up once per ray and
// prefer fuller refills and more steps per iteration; instanced scenes re-do the set-up on every
// instance entry and prefer to refill earlier.  ASGPU_REFILL / ASGPU_FLUSH / ASGPU_STALL /
// ASGPU_STEPS / ASGPU_STEPTHR / ASGPU_PREFETCH override them for experiments.
struct Tuning { int refill, flush, stall, prefetch, steps, step_threshold, enter_late, enter_threshold, enter_min_busy; };

Tuning read_tuning(const bool single_instance)
{
    Tuning v = { 8, 16, 16, 0, 3, 8, 1, 6, 8 };              // prefetch: measured neutral (C2) to -5 % (C3 probes), off
    if (single_instance) { v.refill = 22; v.flush = 24; v.stall = 24; v.steps = 4; v.enter_late = 0; }
    if (const char* e = getenv("ASGPU_REFILL")) v.refill = atoi(e);
    if (const char* e = getenv("ASGPU_FLUSH")) v.flush = atoi(e);
    if (const char* e = getenv("ASGPU_STALL")) v.stall = atoi(e);
    if (const char* e = getenv("ASGPU_PREFETCH")) v.prefetch = atoi(e);
    if (const char* e = getenv("ASGPU_STEPS")) v.steps = atoi(e);
    if (const char* e = getenv("ASGPU_STEPTHR")) v.step_threshold = atoi(e);
    if (const char* e = getenv("ASGPU_ENTERLATE")) v.enter_late = atoi(e);
    if (const char* e = getenv("ASGPU_ENTERTHR")) v.enter_threshold = atoi(e);
    if (const char* e = getenv("ASGPU_ENTERBUSY")) v.enter_min_busy = atoi(e);
    if (v.refill < 1) v.refill = 1;
    if (v.flush < 1) v.flush = 1;
    if (v.stall < 1) v.stall = 1;
    if (v.stall > 32) v.stall = 32;         // 32 stalled lanes = nobody can advance: must flush
    if (v.steps < 1) v.steps = 1;
    return v;
}

// The environment is read once (and again after asgpu_reload_tuning()), not once per launch: a C1
// launch is 44 us of device time and a C5 frame is hundreds of launches.
std::mutex g_tuning_mutex;
bool g_tuning_loaded = false;
Tuning g_tuning[2];
bool g_lean_allowed = true;

Tuning tuning(const bool single_instance, bool& lean_allowed)
{
    std::lock_guard<std::mutex> lock(g_tuning_mutex);
    if (!g_tuning_loaded)
    {
        g_tuning[0] = read_tuning(false);
        g_tuning[1] = read_tuning(true);
        g_lean_allowed = getenv("ASGPU_NO_LEAN") == nullptr;
        g_tuning_loaded = true;
    }
    lean_allowed = g_lean_allowed;
    return g_tuning[single_instance ? 1 : 0];
}

// Occupancy-sized persistent launch of one instantiation.  The opt-in to large dynamic shared
// memory and the occupancy query happen once per (instantiation, device), not per launch.
const int MaxDevices = 64;

template <typename Kernel>
cudaError_t launch_persistent(Kernel kernel, const KernelArgs& args, const size_t smem, const int sm_count, cudaStream_t stream)
{
    static std::atomic<int> cached_blocks[MaxDevices];      // zero-initialised: 0 = not queried yet
    int device = 0;
    cudaError_t err = cudaGetDevice(&device);
    if (err != cudaSuccess) return err;
    int blocks_per_sm = device < MaxDevices ? cached_blocks[device].load(std::memory_order_acquire) : 0;
    if (blocks_per_sm == 0)
    {
        if (smem > 48 * 1024)
        {
            err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (err != cudaSuccess) return err;
        }
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, BlockThreads, smem);
        if (err != cudaSuccess) return err;
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        if (device < MaxDevices) cached_blocks[device].store(blocks_per_sm, std::memory_order_release);
    }
    // Persistent grid: a multiple of the SM count, no larger than the work.
    long long grid = static_cast<long long>(sm_count) * blocks_per_sm;
    const long long needed = static_cast<long long>((args.n + BlockThreads - 1) / BlockThreads);
    if (grid > needed) grid = needed;
    if (grid < 1) grid = 1;
    kernel<<<static_cast<unsigned>(grid), BlockThreads, smem, stream>>>(args);
    return cudaGetLastError();
}

// 5 resident CTAs per SM (<= 102 registers): the measured optimum; 6 CTAs at 80 registers spill
// and run ~10 % slower (profiles/README.md).
const int WideMinBlocks = 5;

template <int STACK, bool LEAN, bool EXTRAS, bool MOTION>
cudaError_t launch_wide(const KernelArgs& args, const bool any_hit, const bool count, const int sm_count, cudaStream_t stream)
{
    const size_t smem = any_hit ? sizeof(WideShared<STACK, MOTION, EXTRAS, LEAN, true>) : sizeof(WideShared<STACK, MOTION, EXTRAS, LEAN, false>);
    // (Shadow probes ignore intersection filters -- TriangleLeafProbeVisitor has none -- but do see animated instances.)
    if (any_hit) return count ? launch_persistent(wide_kernel<true, true, STACK, WideMinBlocks, LEAN, EXTRAS, MOTION>, args, smem, sm_count, stream)
                              : launch_persistent(wide_kernel<true, false, STACK, WideMinBlocks, LEAN, EXTRAS, MOTION>, args, smem, sm_count, stream);
    return count ? launch_persistent(wide_kernel<false, true, STACK, WideMinBlocks, LEAN, EXTRAS, MOTION>, args, smem, sm_count, stream)
                 : launch_persistent(wide_kernel<false, false, STACK, WideMinBlocks, LEAN, EXTRAS, MOTION>, args, smem, sm_count, stream);
}

template <bool LEAN, bool EXTRAS, bool MOTION>
cudaError_t launch_wide_depth(const KernelArgs& args, const uint32_t stack_need, const bool any_hit, const bool count, const int sm_count, cudaStream_t stream)
{
    // The traversal stack lives in shared memory; its depth is the scene's (flatten.cpp computes
    // the bound), rounded up to one of the compiled variants.
    if (stack_need <= 16) return launch_wide<16, LEAN, EXTRAS, MOTION>(args, any_hit, count, sm_count, stream);
    if (stack_need <= 24) return launch_wide<24, LEAN, EXTRAS, MOTION>(args, any_hit, count, sm_count, stream);
    return launch_wide<WideStackMax, LEAN, EXTRAS, MOTION>(args, any_hit, count, sm_count, stream);
}

}   // anonymous namespace

void reload_tuning()
{
    std::lock_guard<std::mutex> lock(g_tuning_mutex);
    g_tuning_loaded = false;
}

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
    void*               stream_,
    const unsigned long long* n_dev,
    const bool          raw_item,
    const asgpu_parent* parents)
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
    args.n_dev = n_dev;
    args.raw_item = raw_item ? 1u : 0u;
    args.parents = parents;
    args.unit_bits = UnitBits;
    bool lean_allowed = true;
    const Tuning knobs = tuning(scene.item_count <= 1, lean_allowed);
    args.refill_threshold = knobs.refill;
    args.flush_threshold = knobs.flush;
    args.stall_threshold = knobs.stall;
    args.prefetch = knobs.prefetch;
    args.max_steps = knobs.steps;
    args.step_threshold = knobs.step_threshold;
    args.enter_late = knobs.enter_late;
    args.enter_threshold = knobs.enter_threshold;
    args.enter_min_busy = knobs.enter_min_busy;

    cudaError_t err = cudaMemsetAsync(queue, 0, sizeof(unsigned long long), stream);
    if (err != cudaSuccess) return static_cast<int>(err);

    const bool count = counters != nullptr;
    if (wide)
    {
        const bool extras = scene.has_filters || scene.has_animated;
        const bool lean = scene.item_count <= 1 && !scene.has_motion && !extras && lean_allowed;
        if (lean) err = launch_wide_depth<true, false, false>(args, scene.wide_stack_need, any_hit, count, sm_count, stream);
        else if (extras) err = launch_wide_depth<false, true, true>(args, scene.wide_stack_need, any_hit, count, sm_count, stream);
        else if (scene.has_motion) err = launch_wide_depth<false, false, true>(args, scene.wide_stack_need, any_hit, count, sm_count, stream);
        else err = launch_wide_depth<false, false, false>(args, scene.wide_stack_need, any_hit, count, sm_count, stream);
    }
    else if (any_hit) err = count ? launch_persistent(trace_kernel<true, true>, args, 0, sm_count, stream)
                                  : launch_persistent(trace_kernel<true, false>, args, 0, sm_count, stream);
    else err = count ? launch_persistent(trace_kernel<false, true>, args, 0, sm_count, stream)
                     : launch_persistent(trace_kernel<false, false>, args, 0, sm_count, stream);
    return static_cast<int>(err);
}

}   // namespace asgpu
