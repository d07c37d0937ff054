//
// api_internal.h -- what the translation units behind the C ABI (api.cu, wavefront.cu) share:
// the scene handle, error reporting.
//
#pragma once

#include "../../include/asgpu.h"
#include "gpu_layout.h"
#include "traverse_core.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <mutex>
#include <string>

namespace asgpu
{

// Records the thread-local error string returned by asgpu_last_error() and returns `code`.
int fail(int code, const std::string& message);
int fail_cuda(cudaError_t err, const char* what);

#define ASGPU_CUDA(call, what) do { const cudaError_t e_ = (call); if (e_ != cudaSuccess) return ::asgpu::fail_cuda(e_, what); } while (0)

// Rays per chunk of the host-buffer entry points (ASGPU_HOST_CHUNK overrides, 64 Ki .. 4 Mi).
size_t host_chunk_rays();
const int HostStreams = 3;
const uint64_t QueueRing = 256;

// Device staging buffers of one stream of the host-buffer entry points.
struct Staging
{
    cudaStream_t    stream = nullptr;
    double*         org = nullptr;
    double*         dir = nullptr;
    double*         tmin = nullptr;
    double*         tmax = nullptr;
    float*          time_absolute = nullptr;
    float*          time_normalized = nullptr;
    uint32_t*       flags = nullptr;
    asgpu_hit*      hits = nullptr;
    uint8_t*        occluded = nullptr;
    unsigned long long* queue = nullptr;
    void*           sort_ws = nullptr;      // coherence-sort workspace + permutation of one chunk
};

}   // namespace asgpu

struct asgpu_scene
{
    int                 device = 0;
    int                 sm_count = 0;
    uint8_t*            blob = nullptr;         // device
    bool                owns_blob = true;
    size_t              blob_bytes = 0;
    asgpu::BlobHeader   header;
    asgpu::SceneView    view;
    unsigned long long* queue = nullptr;        // device, ring of QueueRing cursors (one per launch)
    std::atomic<uint64_t> queue_next{0};    // trace calls may come from several threads (one stream each)
    unsigned long long* counters = nullptr;     // device, asgpu_counters layout (first 6 words)
    std::atomic<uint64_t> launches{0};
    std::mutex          mutex;
    asgpu::Staging      staging[asgpu::HostStreams];
    bool                staging_ready = false;
    bool                has_source = false;     // every triangle tree carries source geometry
    uint32_t*           id_to_item = nullptr;   // device: caller's assembly-instance id -> ItemRecord index (built on demand)
    uint32_t            id_count = 0;
};

namespace asgpu
{
// Builds scene->id_to_item (needs unique instance ids below 2^24).
int ensure_id_table(asgpu_scene* scene);
}
