//
// api.cu -- the extern "C" layer of include/asgpu.h.
//
// Host side of the drop-in boundary: owns the device blob, validates arguments, launches the
// kernels of kernels.cu, and implements the host-buffer entry points as a chunked three-stream
// pipeline (H2D copy / kernel / D2H copy of consecutive chunks overlap).  No torch types, no
// exceptions across the ABI.
//

#include "../../include/asgpu.h"
#include "flatten.h"
#include "api_internal.h"
#include "kernels.h"
#include "tree_builder.h"
#include "parallel.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace asgpu;

namespace asgpu
{

thread_local std::string g_last_error;

int fail(const int code, const std::string& message)
{
    g_last_error = message;
    return code;
}

int fail_cuda(const cudaError_t err, const char* what)
{
    g_last_error = std::string(what) + ": " + cudaGetErrorString(err);
    return ASGPU_E_CUDA;
}

}   // namespace asgpu

struct asgpu_trees
{
    HostTrees trees;
};

namespace
{

// Device counters: a closest-hit and an any-hit bank of traversal counters, then the two lane profiles.
const size_t CounterBytes = 2 * sizeof(asgpu_counters) + 2 * sizeof(asgpu_lane_profile);

void make_view(asgpu_scene* s)
{
    s->view.blob = s->blob;
    s->view.trees = s->header.trees;
    s->view.items = s->header.items;
    s->view.top_nodes = s->header.top_nodes;
    s->view.top_wnodes = s->header.top_wnodes;
    s->view.top_witems = s->header.top_witems;
    s->view.tree_count = s->header.tree_count;
    s->view.item_count = s->header.item_count;
    s->view.top_node_count = s->header.top_node_count;
    s->view.top_wnode_count = s->header.top_wnode_count;
    s->view.wide_stack_need = s->header.wide_stack_need;
    s->view.has_motion = s->header.moving_triangle_count != 0 ? 1u : 0u;
    s->view.has_filters = (s->header.flags & BlobHasFilters) ? 1u : 0u;
    s->view.has_animated = (s->header.flags & BlobHasAnimatedInstances) ? 1u : 0u;
    // Source geometry present for every tree?  (Small table: read it back.)
    s->has_source = s->header.tree_count != 0;
    std::vector<TreeDesc> descs(s->header.tree_count);
    if (!descs.empty() && cudaMemcpy(descs.data(), s->blob + s->header.trees, descs.size() * sizeof(TreeDesc), cudaMemcpyDeviceToHost) != cudaSuccess)
        s->has_source = false;
    for (const TreeDesc& d : descs) if (d.src_objects == 0) s->has_source = false;
}

int init_device_side(asgpu_scene* s)
{
    ASGPU_CUDA(cudaSetDevice(s->device), "cudaSetDevice");
    cudaDeviceProp prop;
    ASGPU_CUDA(cudaGetDeviceProperties(&prop, s->device), "cudaGetDeviceProperties");
    s->sm_count = prop.multiProcessorCount;
    ASGPU_CUDA(cudaMalloc(&s->queue, QueueRing * sizeof(unsigned long long)), "cudaMalloc(queue)");
    ASGPU_CUDA(cudaMalloc(&s->counters, CounterBytes), "cudaMalloc(counters)");
    ASGPU_CUDA(cudaMemset(s->counters, 0, CounterBytes), "cudaMemset(counters)");
    return ASGPU_OK;
}

void free_staging(asgpu_scene* s)
{
    for (Staging& st : s->staging)
    {
        cudaFree(st.org); cudaFree(st.dir); cudaFree(st.tmin); cudaFree(st.tmax);
        cudaFree(st.time_absolute); cudaFree(st.time_normalized); cudaFree(st.flags);
        cudaFree(st.hits); cudaFree(st.occluded); cudaFree(st.queue); cudaFree(st.sort_ws);
        if (st.stream) cudaStreamDestroy(st.stream);
        st = Staging();
    }
    s->staging_ready = false;
}

int allocate_staging(asgpu_scene* s)
{
    for (Staging& st : s->staging)
    {
        ASGPU_CUDA(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking), "cudaStreamCreate");
        ASGPU_CUDA(cudaMalloc(&st.org, host_chunk_rays() * 24), "cudaMalloc(staging)");
        ASGPU_CUDA(cudaMalloc(&st.dir, host_chunk_rays() * 24), "cudaMalloc(staging)");
        ASGPU_CUDA(cudaMalloc(&st.tmin, host_chunk_rays() * 8), "cudaMalloc(staging)");
        ASGPU_CUDA(cudaMalloc(&st.tmax, host_chunk_rays() * 8), "cudaMalloc(staging)");
        ASGPU_CUDA(cudaMalloc(&st.time_absolute, host_chunk_rays() * 4), "cudaMalloc(staging)");
        ASGPU_CUDA(cudaMalloc(&st.time_normalized, host_chunk_rays() * 4), "cudaMalloc(staging)");
        ASGPU_CUDA(cudaMalloc(&st.flags, host_chunk_rays() * 4), "cudaMalloc(staging)");
        ASGPU_CUDA(cudaMalloc(&st.hits, host_chunk_rays() * sizeof(asgpu_hit)), "cudaMalloc(staging)");
        ASGPU_CUDA(cudaMalloc(&st.occluded, host_chunk_rays()), "cudaMalloc(staging)");
        ASGPU_CUDA(cudaMalloc(&st.queue, 64), "cudaMalloc(staging)");
        ASGPU_CUDA(cudaMalloc(&st.sort_ws, ray_sort_workspace_bytes(host_chunk_rays()) + host_chunk_rays() * 4), "cudaMalloc(staging)");
    }
    return ASGPU_OK;
}

int ensure_staging(asgpu_scene* s)
{
    if (s->staging_ready) return ASGPU_OK;
    const int rc = allocate_staging(s);
    if (rc != ASGPU_OK)
    {
        // Nothing half-allocated survives: the next call starts from scratch instead of
        // overwriting (and leaking) the streams and buffers that did get created.
        const std::string message = g_last_error;
        free_staging(s);
        g_last_error = message;
        return rc;
    }
    s->staging_ready = true;
    return ASGPU_OK;
}

int check_trace_args(asgpu_scene* scene, const asgpu_rays* rays, const size_t n, const void* out, const uint32_t flags, bool& wide)
{
    if (!scene) return fail(ASGPU_E_INVALID, "null scene");
    if (n == 0) return ASGPU_OK;
    if (!rays || !rays->org || !rays->dir || !rays->tmin || !rays->tmax) return fail(ASGPU_E_INVALID, "ray batch misses a mandatory array");
    if (!out) return fail(ASGPU_E_INVALID, "null output array");
    wide = (flags & ASGPU_TRACE_EXACT) == 0;
    if (wide && !(scene->header.flags & ASGPU_SCENE_WIDE)) return fail(ASGPU_E_INVALID, "scene was created without the wide layout");
    if (!wide && !(scene->header.flags & ASGPU_SCENE_EXACT)) return fail(ASGPU_E_INVALID, "scene was created without the exact layout");
    return ASGPU_OK;
}

int trace_device(asgpu_scene* scene, const asgpu_rays* rays, const size_t n, asgpu_hit* hits, uint8_t* occluded,
                 const bool any_hit, const uint32_t flags, unsigned long long* queue, void* stream, const asgpu_parent* parents = nullptr)
{
    bool wide = true;
    const int rc = check_trace_args(scene, rays, n, any_hit ? static_cast<const void*>(occluded) : static_cast<const void*>(hits), flags, wide);
    if (rc != ASGPU_OK || n == 0) return rc;
    ASGPU_CUDA(cudaSetDevice(scene->device), "cudaSetDevice");
    // ASGPU_TRACE_SORT: the permutation and the sort workspace are stream-ordered allocations of
    // THIS call (cudaMallocAsync / cudaFreeAsync on the caller's stream), so any number of sorted
    // traces may be in flight on different streams or threads without sharing scratch memory.
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    uint8_t* scratch = nullptr;
    const uint32_t* order = nullptr;
    if (flags & ASGPU_TRACE_SORT)
    {
        if (n > 0xFFFFFFFFull) return fail(ASGPU_E_UNSUPPORTED, "ASGPU_TRACE_SORT handles at most 2^32 - 1 rays per call");
        const size_t ws_bytes = (ray_sort_workspace_bytes(n) + 255) / 256 * 256;
        ASGPU_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), ws_bytes + n * sizeof(uint32_t), cs), "cudaMallocAsync(sort scratch)");
        uint32_t* call_order = reinterpret_cast<uint32_t*>(scratch + ws_bytes);
        const int es = launch_ray_sort(*rays, n, nullptr, call_order, nullptr, scratch, scene->sm_count, stream);
        if (es != 0) { cudaFreeAsync(scratch, cs); return fail_cuda(static_cast<cudaError_t>(es), "ray sort launch"); }
        scene->launches += ray_sort_launch_count();
        order = call_order;
    }
    const int err = launch_trace(scene->view, *rays, n, hits, occluded, any_hit, wide, queue,
                                 (flags & ASGPU_TRACE_COUNTERS) ? scene->counters : nullptr, order, scene->sm_count, stream,
                                 nullptr, false, parents);
    if (scratch) cudaFreeAsync(scratch, cs);        // ordered after the trace kernel on `stream`
    if (err != 0) return fail_cuda(static_cast<cudaError_t>(err), "kernel launch");
    ++scene->launches;
    return ASGPU_OK;
}

int trace_host(asgpu_scene* scene, const asgpu_rays* rays, const size_t n, asgpu_hit* hits, uint8_t* occluded,
               const bool any_hit, const uint32_t flags)
{
    bool wide = true;
    int rc = check_trace_args(scene, rays, n, any_hit ? static_cast<const void*>(occluded) : static_cast<const void*>(hits), flags, wide);
    if (rc != ASGPU_OK || n == 0) return rc;
    std::lock_guard<std::mutex> lock(scene->mutex);
    ASGPU_CUDA(cudaSetDevice(scene->device), "cudaSetDevice");
    rc = ensure_staging(scene);
    if (rc != ASGPU_OK) return rc;

    // One chunk: H2D of its rays, optional sort, trace, D2H of its results, all on one of the
    // staging streams.  The copies are asynchronous when the caller's arrays are page-locked
    // (asgpu_pin_host); from pageable memory the driver stages them synchronously.
    auto run_chunk = [&](const size_t begin, const size_t count, Staging& st) -> int
    {
        // Stream order guarantees the previous use of this staging slot has drained.
        ASGPU_CUDA(cudaMemcpyAsync(st.org, rays->org + begin * 3, count * 24, cudaMemcpyHostToDevice, st.stream), "H2D org");
        ASGPU_CUDA(cudaMemcpyAsync(st.dir, rays->dir + begin * 3, count * 24, cudaMemcpyHostToDevice, st.stream), "H2D dir");
        ASGPU_CUDA(cudaMemcpyAsync(st.tmin, rays->tmin + begin, count * 8, cudaMemcpyHostToDevice, st.stream), "H2D tmin");
        ASGPU_CUDA(cudaMemcpyAsync(st.tmax, rays->tmax + begin, count * 8, cudaMemcpyHostToDevice, st.stream), "H2D tmax");
        asgpu_rays dev;
        dev.org = st.org; dev.dir = st.dir; dev.tmin = st.tmin; dev.tmax = st.tmax;
        dev.time_absolute = nullptr; dev.time_normalized = nullptr; dev.flags = nullptr;
        if (rays->time_absolute)
        {
            ASGPU_CUDA(cudaMemcpyAsync(st.time_absolute, rays->time_absolute + begin, count * 4, cudaMemcpyHostToDevice, st.stream), "H2D time");
            dev.time_absolute = st.time_absolute;
        }
        if (rays->time_normalized)
        {
            ASGPU_CUDA(cudaMemcpyAsync(st.time_normalized, rays->time_normalized + begin, count * 4, cudaMemcpyHostToDevice, st.stream), "H2D time");
            dev.time_normalized = st.time_normalized;
        }
        if (rays->flags)
        {
            ASGPU_CUDA(cudaMemcpyAsync(st.flags, rays->flags + begin, count * 4, cudaMemcpyHostToDevice, st.stream), "H2D flags");
            dev.flags = st.flags;
        }
        const uint32_t* order = nullptr;
        if (flags & ASGPU_TRACE_SORT)
        {
            uint32_t* chunk_order = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(st.sort_ws) + ray_sort_workspace_bytes(host_chunk_rays()));
            const int es = launch_ray_sort(dev, count, nullptr, chunk_order, nullptr, st.sort_ws, scene->sm_count, st.stream);
            if (es != 0) return fail_cuda(static_cast<cudaError_t>(es), "ray sort launch");
            scene->launches += ray_sort_launch_count();
            order = chunk_order;
        }
        const int err = launch_trace(scene->view, dev, count, st.hits, st.occluded, any_hit, wide, st.queue,
                                     (flags & ASGPU_TRACE_COUNTERS) ? scene->counters : nullptr, order, scene->sm_count, st.stream);
        if (err != 0) return fail_cuda(static_cast<cudaError_t>(err), "kernel launch");
        ++scene->launches;
        if (any_hit)
            ASGPU_CUDA(cudaMemcpyAsync(occluded + begin, st.occluded, count, cudaMemcpyDeviceToHost, st.stream), "D2H occluded");
        else
            ASGPU_CUDA(cudaMemcpyAsync(hits + begin, st.hits, count * sizeof(asgpu_hit), cudaMemcpyDeviceToHost, st.stream), "D2H hits");
        return ASGPU_OK;
    };

    size_t chunk_index = 0;
    for (size_t begin = 0; begin < n && rc == ASGPU_OK; begin += host_chunk_rays(), ++chunk_index)
        rc = run_chunk(begin, std::min(host_chunk_rays(), n - begin), scene->staging[chunk_index % HostStreams]);
    // Also on an error: earlier chunks may still be copying into the caller's arrays, which must
    // not be released before those copies have drained.
    const std::string message = g_last_error;
    for (Staging& st : scene->staging)
    {
        const cudaError_t e = cudaStreamSynchronize(st.stream);
        if (e != cudaSuccess && rc == ASGPU_OK) rc = fail_cuda(e, "cudaStreamSynchronize");
        else if (rc != ASGPU_OK) g_last_error = message;
    }
    return rc;
}

}   // anonymous namespace

size_t asgpu::host_chunk_rays()
{
    static const size_t value = []() -> size_t
    {
        size_t v = size_t(1) << 20;
        if (const char* e = getenv("ASGPU_HOST_CHUNK")) { const long long x = atoll(e); if (x > 0) v = static_cast<size_t>(x); }
        return std::min(size_t(4) << 20, std::max(size_t(64) << 10, v));
    }();
    return value;
}

int asgpu::ensure_id_table(asgpu_scene* scene)
{
    if (scene->id_to_item) return ASGPU_OK;
    std::vector<ItemRecord> items(scene->header.item_count);
    if (items.empty()) return fail(ASGPU_E_INVALID, "scene without assembly instances");
    ASGPU_CUDA(cudaMemcpy(items.data(), scene->blob + scene->header.items, items.size() * sizeof(ItemRecord), cudaMemcpyDeviceToHost), "cudaMemcpy(items)");
    uint32_t max_id = 0;
    for (const ItemRecord& it : items) max_id = std::max(max_id, it.assembly_instance);
    if (max_id >= (1u << 24)) return fail(ASGPU_E_UNSUPPORTED, "assembly-instance ids above 2^24 - 1");
    std::vector<uint32_t> table(size_t(max_id) + 1, 0xFFFFFFFFu);
    for (size_t i = 0; i < items.size(); ++i)
    {
        if (table[items[i].assembly_instance] != 0xFFFFFFFFu) return fail(ASGPU_E_INVALID, "assembly-instance ids are not unique");
        table[items[i].assembly_instance] = static_cast<uint32_t>(i);
    }
    ASGPU_CUDA(cudaMalloc(&scene->id_to_item, table.size() * 4), "cudaMalloc(id table)");
    ASGPU_CUDA(cudaMemcpy(scene->id_to_item, table.data(), table.size() * 4, cudaMemcpyHostToDevice), "cudaMemcpy(id table)");
    scene->id_count = static_cast<uint32_t>(table.size());
    return ASGPU_OK;
}

namespace
{

// Host image -> device through two page-locked staging buffers: a pageable cudaMemcpy of a
// multi-gigabyte blob runs at a fraction of the link rate; here several threads fill one buffer
// while the other one is on its way.
cudaError_t upload_blob(uint8_t* device_dst, const uint8_t* host_src, const size_t bytes)
{
    const size_t Chunk = size_t(64) << 20;
    if (bytes <= Chunk) return cudaMemcpy(device_dst, host_src, bytes, cudaMemcpyHostToDevice);
    uint8_t* staging[2] = { nullptr, nullptr };
    cudaEvent_t done[2] = { nullptr, nullptr };
    cudaStream_t stream = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k)
    {
        e = cudaMallocHost(reinterpret_cast<void**>(&staging[k]), Chunk);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming);
    }
    size_t index = 0;
    for (size_t begin = 0; begin < bytes && e == cudaSuccess; begin += Chunk, ++index)
    {
        const int k = static_cast<int>(index & 1);
        const size_t n = std::min(Chunk, bytes - begin);
        if (index >= 2) e = cudaEventSynchronize(done[k]);          // the previous copy out of this buffer has finished
        if (e != cudaSuccess) break;
        parallel_memcpy(staging[k], host_src + begin, n, host_threads());
        e = cudaMemcpyAsync(device_dst + begin, staging[k], n, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaEventRecord(done[k], stream);
    }
    if (stream) { const cudaError_t s = cudaStreamSynchronize(stream); if (e == cudaSuccess) e = s; }
    for (int k = 0; k < 2; ++k) { if (staging[k]) cudaFreeHost(staging[k]); if (done[k]) cudaEventDestroy(done[k]); }
    if (stream) cudaStreamDestroy(stream);
    return e;
}

asgpu_scene* adopt_blob_image(const HostBlob& image, const int device)
{
    {
        // Structural check of what the flattener produced (offsets and counts of every table).
        std::string error;
        const int rc = validate_blob(image.data(), image.size(), error);
        if (rc != ASGPU_OK) { fail(rc, "flattened scene failed validation: " + error); return nullptr; }
    }
    asgpu_scene* s = new (std::nothrow) asgpu_scene();
    if (!s) { fail(ASGPU_E_NOMEM, "out of host memory"); return nullptr; }
    s->device = device;
    std::memcpy(&s->header, image.data(), sizeof(BlobHeader));
    s->blob_bytes = image.size();
    if (init_device_side(s) != ASGPU_OK) { asgpu_scene_destroy(s); return nullptr; }
    cudaError_t e = cudaMalloc(&s->blob, s->blob_bytes);
    if (e != cudaSuccess) { fail_cuda(e, "cudaMalloc(blob)"); asgpu_scene_destroy(s); return nullptr; }
    e = upload_blob(s->blob, image.data(), s->blob_bytes);
    if (e != cudaSuccess) { fail_cuda(e, "cudaMemcpy(blob)"); asgpu_scene_destroy(s); return nullptr; }
    make_view(s);
    return s;
}

}   // anonymous namespace

extern "C" {

const char* asgpu_last_error(void) { return g_last_error.c_str(); }

int asgpu_version(void) { return ASGPU_VERSION; }

// ---- host builder -------------------------------------------------------------------------

asgpu_trees* asgpu_trees_build(const asgpu_scene_desc* desc, int threads)
{
    if (!desc) { fail(ASGPU_E_INVALID, "null scene description"); return nullptr; }
    asgpu_trees* t = new (std::nothrow) asgpu_trees();
    if (!t) { fail(ASGPU_E_NOMEM, "out of host memory"); return nullptr; }
    std::string error;
    bool ok = false;
    try { ok = build_host_trees(*desc, threads, t->trees, error); }
    catch (const std::exception& e) { error = e.what(); }
    if (!ok) { fail(ASGPU_E_INVALID, error); delete t; return nullptr; }
    return t;
}

asgpu_trees* asgpu_trees_build_animated(const asgpu_scene_desc* desc, const asgpu_instance_keys* keys, int threads)
{
    if (!desc) { fail(ASGPU_E_INVALID, "null scene description"); return nullptr; }
    asgpu_trees* t = new (std::nothrow) asgpu_trees();
    if (!t) { fail(ASGPU_E_NOMEM, "out of host memory"); return nullptr; }
    std::string error;
    bool ok = false;
    try { ok = build_host_trees(*desc, threads, t->trees, error, nullptr, nullptr, keys); }
    catch (const std::exception& e) { error = e.what(); }
    if (!ok) { fail(ASGPU_E_INVALID, error); delete t; return nullptr; }
    return t;
}

asgpu_trees* asgpu_trees_build_on_device(const asgpu_scene_desc* desc, int threads, int device)
{
    if (!desc) { fail(ASGPU_E_INVALID, "null scene description"); return nullptr; }
    int device_count = 0;
    if (cudaGetDeviceCount(&device_count) != cudaSuccess || device < 0 || device >= device_count)
    { fail(ASGPU_E_CUDA, "asgpu_trees_build_on_device: no such CUDA device"); return nullptr; }
    asgpu_trees* t = new (std::nothrow) asgpu_trees();
    if (!t) { fail(ASGPU_E_NOMEM, "out of host memory"); return nullptr; }
    std::string error;
    bool ok = false;
    // ASGPU_DEVICE_BUILD = lbvh: the linear BVH (Morton splits); default: locally-ordered clustering.
    const char* algorithm = getenv("ASGPU_DEVICE_BUILD");
    const bool linear = algorithm != nullptr && std::strcmp(algorithm, "lbvh") == 0;
    try { ok = build_host_trees(*desc, threads, t->trees, error, linear ? lbvh_topology_device : ploc_topology_device, &device); }
    catch (const std::exception& e) { error = e.what(); }
    if (!ok) { fail(error.compare(0, 18, "device tree build:") == 0 ? ASGPU_E_CUDA : ASGPU_E_INVALID, error); delete t; return nullptr; }
    return t;
}

void asgpu_trees_destroy(asgpu_trees* trees) { delete trees; }

int asgpu_trees_triangle_tree_count(const asgpu_trees* trees)
{
    return trees ? static_cast<int>(trees->trees.triangle_trees.size()) : fail(ASGPU_E_INVALID, "null trees");
}

int asgpu_trees_get_triangle_tree(const asgpu_trees* trees, int index, asgpu_triangle_tree_view* out)
{
    if (!trees || !out) return fail(ASGPU_E_INVALID, "null argument");
    if (index < 0 || index >= static_cast<int>(trees->trees.triangle_trees.size())) return fail(ASGPU_E_INVALID, "triangle tree index out of range");
    const HostTriangleTree& t = *trees->trees.triangle_trees[index];
    out->nodes = t.nodes.data();
    out->node_bboxes = t.node_bboxes.empty() ? nullptr : t.node_bboxes.data();
    out->leaf_data = t.leaf_data.empty() ? nullptr : t.leaf_data.data();
    out->triangle_keys = t.keys.empty() ? nullptr : t.keys.data();
    out->node_count = t.nodes.size();
    out->node_bbox_count = t.node_bboxes.size() / 6;
    out->leaf_data_size = t.leaf_data.size();
    out->triangle_key_count = t.keys.size();
    out->static_triangle_count = t.static_triangle_count;
    out->moving_triangle_count = t.moving_triangle_count;
    return ASGPU_OK;
}

int asgpu_trees_get_assembly_tree(const asgpu_trees* trees, asgpu_assembly_tree_view* out)
{
    if (!trees || !out) return fail(ASGPU_E_INVALID, "null argument");
    const HostAssemblyTree& t = trees->trees.assembly_tree;
    out->nodes = t.nodes.data();
    out->items = t.items.empty() ? nullptr : t.items.data();
    out->node_count = t.nodes.size();
    out->item_count = t.items.size();
    out->item_motion = t.item_motion.empty() ? nullptr : t.item_motion.data();      // asgpu_trees_build_animated
    return ASGPU_OK;
}

double asgpu_trees_build_seconds(const asgpu_trees* trees) { return trees ? trees->trees.build_seconds : 0.0; }
double asgpu_trees_device_seconds(const asgpu_trees* trees) { return trees ? trees->trees.topology_seconds : 0.0; }

int asgpu_trees_get_source_geometry(const asgpu_trees* trees, int index, asgpu_source_geometry* out)
{
    if (!trees || !out) return fail(ASGPU_E_INVALID, "null argument");
    if (index < 0 || index >= static_cast<int>(trees->trees.triangle_trees.size())) return fail(ASGPU_E_INVALID, "triangle tree index out of range");
    const HostTriangleTree& t = *trees->trees.triangle_trees[index];
    out->objects = t.source_objects.empty() ? nullptr : t.source_objects.data();
    out->object_count = static_cast<uint32_t>(t.source_objects.size());
    out->reserved = 0;
    return ASGPU_OK;
}

// ---- GPU scene ----------------------------------------------------------------------------

asgpu_scene* asgpu_scene_create(
    const asgpu_triangle_tree_view* triangle_trees,
    uint32_t                        triangle_tree_count,
    const asgpu_assembly_tree_view* assembly_tree,
    uint32_t                        flags,
    int                             device)
{
    return asgpu_scene_create_ex(triangle_trees, triangle_tree_count, assembly_tree, nullptr, flags, device);
}

asgpu_scene* asgpu_scene_create_ex(
    const asgpu_triangle_tree_view* triangle_trees,
    uint32_t                        triangle_tree_count,
    const asgpu_assembly_tree_view* assembly_tree,
    const asgpu_source_geometry*    sources,
    uint32_t                        flags,
    int                             device)
{
    if (!assembly_tree) { fail(ASGPU_E_INVALID, "null assembly tree"); return nullptr; }
    HostBlob image;
    std::string error;
    int rc;
    try { rc = flatten_scene(triangle_trees, triangle_tree_count, *assembly_tree, sources, flags ? flags : ASGPU_SCENE_DEFAULT, image, error); }
    catch (const std::exception& e) { rc = ASGPU_E_NOMEM; error = e.what(); }
    if (rc != ASGPU_OK) { fail(rc, error); return nullptr; }
    return adopt_blob_image(image, device);
}

asgpu_scene* asgpu_scene_create_from_desc(const asgpu_scene_desc* desc, uint32_t flags, int device, int threads)
{
    asgpu_trees* trees = asgpu_trees_build(desc, threads);
    if (!trees) return nullptr;
    std::vector<asgpu_triangle_tree_view> views(trees->trees.triangle_trees.size());
    std::vector<asgpu_source_geometry> sources(views.size());
    for (size_t i = 0; i < views.size(); ++i)
    {
        asgpu_trees_get_triangle_tree(trees, static_cast<int>(i), &views[i]);
        asgpu_trees_get_source_geometry(trees, static_cast<int>(i), &sources[i]);
    }
    asgpu_assembly_tree_view top;
    asgpu_trees_get_assembly_tree(trees, &top);
    asgpu_scene* scene = asgpu_scene_create_ex(views.empty() ? nullptr : views.data(), static_cast<uint32_t>(views.size()), &top,
                                               sources.empty() ? nullptr : sources.data(), flags, device);
    asgpu_trees_destroy(trees);
    return scene;
}

void asgpu_scene_destroy(asgpu_scene* scene)
{
    if (!scene) return;
    cudaSetDevice(scene->device);
    free_staging(scene);
    if (scene->owns_blob) cudaFree(scene->blob);
    cudaFree(scene->queue);
    cudaFree(scene->counters);
    cudaFree(scene->id_to_item);
    delete scene;
}

size_t asgpu_scene_blob_size(const asgpu_scene* scene) { return scene ? scene->blob_bytes : 0; }

const void* asgpu_scene_blob_device_ptr(const asgpu_scene* scene) { return scene ? scene->blob : nullptr; }

int asgpu_scene_export_blob(const asgpu_scene* scene, void* device_dst, size_t capacity, void* stream)
{
    if (!scene || !device_dst) return fail(ASGPU_E_INVALID, "null argument");
    if (capacity < scene->blob_bytes) return fail(ASGPU_E_INVALID, "destination too small for the scene blob");
    ASGPU_CUDA(cudaSetDevice(scene->device), "cudaSetDevice");
    ASGPU_CUDA(cudaMemcpyAsync(device_dst, scene->blob, scene->blob_bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)), "cudaMemcpyAsync(blob)");
    return ASGPU_OK;
}

asgpu_scene* asgpu_scene_import_blob(const void* device_blob, size_t size, int device, int adopt)
{
    if (!device_blob || size < sizeof(BlobHeader)) { fail(ASGPU_E_INVALID, "blob too small"); return nullptr; }
    asgpu_scene* s = new (std::nothrow) asgpu_scene();
    if (!s) { fail(ASGPU_E_NOMEM, "out of host memory"); return nullptr; }
    s->device = device;
    s->blob_bytes = size;
    if (init_device_side(s) != ASGPU_OK) { asgpu_scene_destroy(s); return nullptr; }
    // The payload usually arrives by a collective on some stream of the caller's (the NCCL broadcast
    // of distributed.replicate_scene): everything enqueued on the device must have landed before the
    // header is read.  Import happens once per scene, a device-wide wait is cheap here.
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fail_cuda(e, "cudaDeviceSynchronize"); asgpu_scene_destroy(s); return nullptr; }
    // Only the header and the small tables are inspected on the host -- with the same structural
    // checks the flattener's own output goes through, so a truncated or corrupted payload is
    // refused here instead of being dereferenced by the kernels.
    {
        const uint8_t* src = static_cast<const uint8_t*>(device_blob);
        std::string error;
        const int rc = validate_blob_tables(
            [src, size](uint64_t offset, void* dst, size_t bytes) -> bool
            {
                if (offset > size || bytes > size - offset) return false;
                return cudaMemcpy(dst, src + offset, bytes, cudaMemcpyDeviceToHost) == cudaSuccess;
            }, size, error);
        if (rc != ASGPU_OK) { fail(rc, "imported blob failed validation: " + error); asgpu_scene_destroy(s); return nullptr; }
    }
    e = cudaMemcpy(&s->header, device_blob, sizeof(BlobHeader), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { fail_cuda(e, "cudaMemcpy(blob header)"); asgpu_scene_destroy(s); return nullptr; }
    if (adopt)
    {
        s->blob = const_cast<uint8_t*>(static_cast<const uint8_t*>(device_blob));
        s->owns_blob = false;
    }
    else
    {
        e = cudaMalloc(&s->blob, size);
        if (e != cudaSuccess) { fail_cuda(e, "cudaMalloc(blob)"); asgpu_scene_destroy(s); return nullptr; }
        e = cudaMemcpy(s->blob, device_blob, size, cudaMemcpyDeviceToDevice);
        if (e != cudaSuccess) { fail_cuda(e, "cudaMemcpy(blob)"); asgpu_scene_destroy(s); return nullptr; }
    }
    make_view(s);
    return s;
}

int asgpu_scene_get_info(const asgpu_scene* scene, asgpu_scene_info* out)
{
    if (!scene || !out) return fail(ASGPU_E_INVALID, "null argument");
    const BlobHeader& h = scene->header;
    out->blob_bytes = scene->blob_bytes;
    out->triangle_tree_count = h.tree_count;
    out->instance_count = h.item_count;
    out->triangle_count = h.triangle_count;
    out->moving_triangle_count = h.moving_triangle_count;
    out->binary_node_count = h.binary_node_count;
    out->wide_node_count = h.wide_node_count;
    out->binary_node_bytes = h.binary_node_bytes;
    out->wide_node_bytes = h.wide_node_bytes;
    out->triangle_bytes = h.triangle_bytes;
    out->flags = h.flags;
    out->wide_stack_depth = h.wide_stack_need;
    return ASGPU_OK;
}

// ---- tracing ------------------------------------------------------------------------------

int asgpu_trace(asgpu_scene* scene, const asgpu_rays* rays, size_t n, asgpu_hit* hits, uint32_t flags, void* stream)
{
    return trace_device(scene, rays, n, hits, nullptr, false, flags, scene ? scene->queue + (scene->queue_next++ % QueueRing) : nullptr, stream);
}

int asgpu_trace_probe(asgpu_scene* scene, const asgpu_rays* rays, size_t n, uint8_t* occluded, uint32_t flags, void* stream)
{
    return trace_device(scene, rays, n, nullptr, occluded, true, flags, scene ? scene->queue + (scene->queue_next++ % QueueRing) : nullptr, stream);
}

int asgpu_trace_host(asgpu_scene* scene, const asgpu_rays* rays, size_t n, asgpu_hit* hits, uint32_t flags)
{
    return trace_host(scene, rays, n, hits, nullptr, false, flags);
}

int asgpu_trace_probe_host(asgpu_scene* scene, const asgpu_rays* rays, size_t n, uint8_t* occluded, uint32_t flags)
{
    return trace_host(scene, rays, n, nullptr, occluded, true, flags);
}

int asgpu_trace_with_parents(asgpu_scene* scene, const asgpu_rays* rays, const asgpu_parent* parents, size_t n, asgpu_hit* hits, uint32_t flags, void* stream)
{
    if (n != 0 && !parents) return fail(ASGPU_E_INVALID, "null parent array");
    return trace_device(scene, rays, n, hits, nullptr, false, flags, scene ? scene->queue + (scene->queue_next++ % QueueRing) : nullptr, stream, parents);
}

int asgpu_trace_probe_with_parents(asgpu_scene* scene, const asgpu_rays* rays, const asgpu_parent* parents, size_t n, uint8_t* occluded, uint32_t flags, void* stream)
{
    if (n != 0 && !parents) return fail(ASGPU_E_INVALID, "null parent array");
    return trace_device(scene, rays, n, nullptr, occluded, true, flags, scene ? scene->queue + (scene->queue_next++ % QueueRing) : nullptr, stream, parents);
}

int asgpu_refine_and_offset(asgpu_scene* scene, const asgpu_rays* rays, const asgpu_hit* hits, size_t n, asgpu_parent* parents, void* stream)
{
    if (!scene) return fail(ASGPU_E_INVALID, "null scene");
    if (n == 0) return ASGPU_OK;
    if (!rays || !rays->org || !rays->dir || !hits || !parents) return fail(ASGPU_E_INVALID, "null argument");
    if (!(scene->header.flags & ASGPU_SCENE_EXACT)) return fail(ASGPU_E_INVALID, "refine_and_offset needs the per-slot triangle records of the exact layout");
    if (!scene->has_source) return fail(ASGPU_E_INVALID, "the scene was created without source geometry (asgpu_scene_create_ex)");
    ASGPU_CUDA(cudaSetDevice(scene->device), "cudaSetDevice");
    const int rt = ensure_id_table(scene);
    if (rt != ASGPU_OK) return rt;
    const int err = launch_refine_offset(scene->view, *rays, hits, n, nullptr, false, scene->id_to_item, scene->id_count, parents, scene->sm_count, stream);
    if (err != 0) return fail_cuda(static_cast<cudaError_t>(err), "kernel launch");
    ++scene->launches;
    return ASGPU_OK;
}

int asgpu_get_support_planes(asgpu_scene* scene, const asgpu_rays* rays, const asgpu_hit* hits, size_t n, double* planes, void* stream)
{
    if (!scene) return fail(ASGPU_E_INVALID, "null scene");
    if (n == 0) return ASGPU_OK;
    if (!rays || !hits || !planes) return fail(ASGPU_E_INVALID, "null argument");
    if (!(scene->header.flags & ASGPU_SCENE_EXACT)) return fail(ASGPU_E_INVALID, "support planes need the per-slot triangle records of the exact layout");
    ASGPU_CUDA(cudaSetDevice(scene->device), "cudaSetDevice");
    const int rt = ensure_id_table(scene);
    if (rt != ASGPU_OK) return rt;
    const int err = launch_support_planes(scene->view, *rays, hits, n, false, scene->id_to_item, scene->id_count, planes, scene->sm_count, stream);
    if (err != 0) return fail_cuda(static_cast<cudaError_t>(err), "kernel launch");
    ++scene->launches;
    return ASGPU_OK;
}

int asgpu_pin_host(void* ptr, size_t bytes)
{
    if (!ptr || bytes == 0) return fail(ASGPU_E_INVALID, "null range");
    ASGPU_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable), "cudaHostRegister");
    return ASGPU_OK;
}

int asgpu_unpin_host(void* ptr)
{
    if (!ptr) return fail(ASGPU_E_INVALID, "null range");
    ASGPU_CUDA(cudaHostUnregister(ptr), "cudaHostUnregister");
    return ASGPU_OK;
}

void asgpu_reload_tuning(void) { reload_tuning(); }

int asgpu_sort_rays(asgpu_scene* scene, const asgpu_rays* rays, size_t n, uint32_t* order, uint32_t* keys, void* stream)
{
    if (!scene) return fail(ASGPU_E_INVALID, "null scene");
    if (n == 0) return ASGPU_OK;
    if (!rays || !rays->org || !rays->dir || !order) return fail(ASGPU_E_INVALID, "null argument");
    if (n > 0xFFFFFFFFull) return fail(ASGPU_E_UNSUPPORTED, "at most 2^32 - 1 rays per sort");
    ASGPU_CUDA(cudaSetDevice(scene->device), "cudaSetDevice");
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    void* ws = nullptr;
    ASGPU_CUDA(cudaMallocAsync(&ws, ray_sort_workspace_bytes(n), cs), "cudaMallocAsync(sort workspace)");
    const int es = launch_ray_sort(*rays, n, nullptr, order, keys, ws, scene->sm_count, stream);
    cudaFreeAsync(ws, cs);
    if (es != 0) return fail_cuda(static_cast<cudaError_t>(es), "ray sort launch");
    scene->launches += ray_sort_launch_count();
    return ASGPU_OK;
}

int asgpu_get_counters_by_kind(asgpu_scene* scene, asgpu_counters* closest, asgpu_counters* probe, int reset)
{
    if (!scene) return fail(ASGPU_E_INVALID, "null argument");
    ASGPU_CUDA(cudaSetDevice(scene->device), "cudaSetDevice");
    ASGPU_CUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    asgpu_counters banks[2];
    ASGPU_CUDA(cudaMemcpy(banks, scene->counters, sizeof(banks), cudaMemcpyDeviceToHost), "cudaMemcpy(counters)");
    for (asgpu_counters& b : banks) { b.kernel_launches = scene->launches; b.reserved = 0; }
    if (closest) *closest = banks[0];
    if (probe) *probe = banks[1];
    if (reset)
    {
        ASGPU_CUDA(cudaMemset(scene->counters, 0, CounterBytes), "cudaMemset(counters)");
        scene->launches = 0;
    }
    return ASGPU_OK;
}

int asgpu_get_lane_profile(asgpu_scene* scene, asgpu_lane_profile* closest, asgpu_lane_profile* probe)
{
    if (!scene) return fail(ASGPU_E_INVALID, "null argument");
    ASGPU_CUDA(cudaSetDevice(scene->device), "cudaSetDevice");
    ASGPU_CUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    asgpu_lane_profile banks[2];
    static_assert(sizeof(banks) == CounterBytes - 2 * sizeof(asgpu_counters), "lane profile banks");
    ASGPU_CUDA(cudaMemcpy(banks, reinterpret_cast<const uint8_t*>(scene->counters) + 2 * sizeof(asgpu_counters), sizeof(banks), cudaMemcpyDeviceToHost), "cudaMemcpy(lane profile)");
    if (closest) *closest = banks[0];
    if (probe) *probe = banks[1];
    return ASGPU_OK;
}

int asgpu_get_counters(asgpu_scene* scene, asgpu_counters* out, int reset)
{
    if (!scene || !out) return fail(ASGPU_E_INVALID, "null argument");
    asgpu_counters closest, probe;
    const int rc = asgpu_get_counters_by_kind(scene, &closest, &probe, reset);
    if (rc != ASGPU_OK) return rc;
    out->rays = closest.rays + probe.rays;
    out->assembly_nodes_visited = closest.assembly_nodes_visited + probe.assembly_nodes_visited;
    out->instances_visited = closest.instances_visited + probe.instances_visited;
    out->triangle_nodes_visited = closest.triangle_nodes_visited + probe.triangle_nodes_visited;
    out->triangles_tested = closest.triangles_tested + probe.triangles_tested;
    out->hits = closest.hits + probe.hits;
    out->kernel_launches = closest.kernel_launches;
    out->reserved = 0;
    return ASGPU_OK;
}

}   // extern "C"
