//
// kernels.h -- launch interface between the C ABI (api.cu) and the kernels (kernels.cu).
//
#pragma once

#include "../../include/asgpu.h"
#include "traverse_core.h"

namespace asgpu
{

// Enqueues one trace over `n` rays on `stream`.  Exactly one kernel of this library is launched
// (plus a memset of the queue cursor).  Returns a cudaError_t value (0 = success).
int launch_trace(
    const SceneView&    scene,
    const asgpu_rays&   rays,       // device pointers
    size_t              n,
    asgpu_hit*          hits,       // closest hit output (device) or nullptr
    uint8_t*            occluded,   // any hit output (device) or nullptr
    bool                any_hit,
    bool                wide,
    unsigned long long* queue,      // device: ray queue cursor (8 bytes)
    unsigned long long* counters,   // device: asgpu_counters or nullptr
    const uint32_t*     order,      // device: optional ray permutation
    int                 sm_count,
    void*               stream,
    const unsigned long long* n_dev = nullptr,  // device: when set, the ray count is read from here (n = capacity bound)
    bool                raw_item = false,       // hit records carry the ItemRecord index instead of the caller's instance id
    const asgpu_parent* parents = nullptr);     // device: optional parent shading point per ray

// Forgets the cached ASGPU_* scheduling knobs: the next launch reads the environment again.
void reload_tuning();

// ShadingPoint::refine_and_offset for n hits (refine.cu).  `raw_item`: hits carry ItemRecord
// indices (wavefront launches); the parents written always carry the caller's instance id.
int launch_refine_offset(
    const SceneView&    scene,
    const asgpu_rays&   rays,
    const asgpu_hit*    hits,
    size_t              n,
    const unsigned long long* n_dev,
    bool                raw_item,
    const uint32_t*     id_to_item,     // device: caller's instance id -> ItemRecord index (unused for raw hits)
    uint32_t            id_count,
    asgpu_parent*       parents,
    int                 sm_count,
    void*               stream);

// ShadingPoint::m_triangle_support_plane for n hits (refine.cu): 9 doubles per hit.
int launch_support_planes(
    const SceneView&    scene,
    const asgpu_rays&   rays,           // only time_normalized is read
    const asgpu_hit*    hits,
    size_t              n,
    bool                raw_item,
    const uint32_t*     id_to_item,
    uint32_t            id_count,
    double*             planes,
    int                 sm_count,
    void*               stream);

// Coherence sort (sort.cu): fills `order` with the permutation that sorts the rays by their
// origin / direction Morton key.  Returns a cudaError_t value.
size_t ray_sort_workspace_bytes(size_t n);
int ray_sort_launch_count();
int launch_ray_sort(
    const asgpu_rays&   rays,       // device pointers
    size_t              n,
    const unsigned long long* n_dev,    // optional device-side count (<= n)
    uint32_t*           order,      // device: n entries
    uint32_t*           keys_out,   // device: n sorted keys, or nullptr
    void*               workspace,  // device: ray_sort_workspace_bytes(n)
    int                 sm_count,
    void*               stream);

}   // namespace asgpu
