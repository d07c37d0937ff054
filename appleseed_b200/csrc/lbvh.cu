//
// lbvh.cu -- triangle-tree topology built on the device.
//
// Replaces the reference's single-threaded sweep SAH (bvh_sahpartitioner.h:99-170 driven by
// bvh_builder.h:163-229; ~0.8 s per million triangles) for callers that prefer build speed to the
// reference's exact tree: Morton keys of the build boxes -> radix sort -> every interior node of
// the binary radix tree in parallel (lbvh_core.h) -> boxes bottom-up, one thread per leaf, the
// second thread to reach a node merges its children.  All kernels stream over flat arrays (one
// coalesced pass each, HBM bound); the sort is cub::DeviceRadixSort (library code off the trace
// path).  The host turns the topology into reference-format nodes (tree_builder.cpp: emit_lbvh).
//

#include "kernels.h"
#include "lbvh_core.h"
#include "tree_builder.h"

#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#include <string>

namespace asgpu
{

namespace
{

const int LbvhThreads = 256;

struct Float3 { float v[3]; };

__global__ void __launch_bounds__(LbvhThreads)
lbvh_keys_kernel(const float* __restrict__ boxes, const uint32_t n, const Float3 origin, const Float3 scale, uint64_t* __restrict__ keys, uint32_t* __restrict__ ids)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        float box[6];
        #pragma unroll
        for (int k = 0; k < 6; ++k) box[k] = __ldg(boxes + size_t(i) * 6 + k);
        keys[i] = lbvh_morton(box, origin.v, scale.v);
        ids[i] = i;
    }
}

__global__ void __launch_bounds__(LbvhThreads)
lbvh_hierarchy_kernel(const uint64_t* __restrict__ keys, const uint32_t n, uint32_t* __restrict__ left, uint32_t* __restrict__ right,
                      uint32_t* __restrict__ first, uint32_t* __restrict__ last, uint32_t* __restrict__ node_parent, uint32_t* __restrict__ leaf_parent)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += gridDim.x * blockDim.x)
    {
        uint32_t l, r, f, e;
        lbvh_node(keys, n, i, l, r, f, e);
        left[i] = l; right[i] = r; first[i] = f; last[i] = e;
        if (l & LbvhLeafFlag) leaf_parent[l & ~LbvhLeafFlag] = i; else node_parent[l] = i;
        if (r & LbvhLeafFlag) leaf_parent[r & ~LbvhLeafFlag] = i; else node_parent[r] = i;
    }
}

__device__ __forceinline__ void load_box(const float* p, float box[6])
{
    #pragma unroll
    for (int k = 0; k < 6; ++k) box[k] = __ldcg(p + k);       // written by another thread: bypass L1
}

// One thread per leaf climbs towards the root; at every node the first arrival stops, the second
// (which knows both children are complete) writes the node's box and goes on.
__global__ void __launch_bounds__(LbvhThreads)
lbvh_refit_kernel(const float* __restrict__ boxes, const uint32_t* __restrict__ order, const uint32_t n, const uint32_t* __restrict__ left,
                  const uint32_t* __restrict__ right, const uint32_t* __restrict__ node_parent, const uint32_t* __restrict__ leaf_parent,
                  uint32_t* arrivals, float* node_boxes)
{
    for (uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x; leaf < n; leaf += gridDim.x * blockDim.x)
    {
        uint32_t node = leaf_parent[leaf];
        for (;;)
        {
            __threadfence();
            if (atomicAdd(arrivals + node, 1u) == 0u) break;
            __threadfence();
            float a[6], b[6];
            const uint32_t l = left[node], r = right[node];
            if (l & LbvhLeafFlag) load_box(boxes + size_t(order[l & ~LbvhLeafFlag]) * 6, a); else load_box(node_boxes + size_t(l) * 6, a);
            if (r & LbvhLeafFlag) load_box(boxes + size_t(order[r & ~LbvhLeafFlag]) * 6, b); else load_box(node_boxes + size_t(r) * 6, b);
            #pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                __stcg(node_boxes + size_t(node) * 6 + k, fminf(a[k], b[k]));
                __stcg(node_boxes + size_t(node) * 6 + 3 + k, fmaxf(a[3 + k], b[3 + k]));
            }
            if (node == 0) break;
            node = node_parent[node];
        }
    }
}

struct DeviceBuffers
{
    void* ptrs[16];
    int count = 0;
    ~DeviceBuffers() { for (int i = 0; i < count; ++i) cudaFree(ptrs[i]); }
    template <typename T> bool alloc(T*& p, const size_t elements)
    {
        void* q = nullptr;
        if (cudaMalloc(&q, (elements ? elements : 1) * sizeof(T)) != cudaSuccess) return false;
        ptrs[count++] = q;
        p = static_cast<T*>(q);
        return true;
    }
};

}   // anonymous namespace

// LbvhTopologyFn of the product: `context` points at the int device ordinal.
bool lbvh_topology_device(const float* boxes, size_t n, const float root_lo[3], const float root_hi[3], void* context,
                          LbvhTopology& out, std::string& error)
{
    const int device = context ? *static_cast<const int*>(context) : 0;
    auto cuda_failed = [&error](const cudaError_t e, const char* what) -> bool
    {
        if (e == cudaSuccess) return false;
        error = std::string("device tree build: ") + what + ": " + cudaGetErrorString(e);
        return true;
    };
    if (n < 2 || n >= LbvhLeafFlag) { error = "device tree build: item count out of range"; return false; }
    if (cuda_failed(cudaSetDevice(device), "cudaSetDevice")) return false;
    cudaDeviceProp prop;
    if (cuda_failed(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return false;

    Float3 origin, scale;
    for (int a = 0; a < 3; ++a)
    {
        const float extent = 2.0f * (root_hi[a] - root_lo[a]);
        origin.v[a] = 2.0f * root_lo[a];
        scale.v[a] = extent > 0.0f ? 2097152.0f / extent : 0.0f;
    }

    const uint32_t count = static_cast<uint32_t>(n);
    DeviceBuffers mem;
    float *d_boxes, *d_node_boxes;
    uint64_t *d_keys, *d_keys_sorted;
    uint32_t *d_ids, *d_order, *d_left, *d_right, *d_first, *d_last, *d_node_parent, *d_leaf_parent, *d_arrivals;
    void* d_temp = nullptr;
    size_t temp_bytes = 0;
    if (cuda_failed(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, static_cast<const uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr),
                                                    static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), static_cast<int>(count), 0, 63),
                    "radix sort sizing")) return false;
    uint8_t* d_temp_bytes;
    if (!mem.alloc(d_boxes, n * 6) || !mem.alloc(d_node_boxes, (n - 1) * 6) || !mem.alloc(d_keys, n) || !mem.alloc(d_keys_sorted, n) ||
        !mem.alloc(d_ids, n) || !mem.alloc(d_order, n) || !mem.alloc(d_left, n - 1) || !mem.alloc(d_right, n - 1) || !mem.alloc(d_first, n - 1) ||
        !mem.alloc(d_last, n - 1) || !mem.alloc(d_node_parent, n - 1) || !mem.alloc(d_leaf_parent, n) || !mem.alloc(d_arrivals, n - 1) ||
        !mem.alloc(d_temp_bytes, temp_bytes))
    { error = "device tree build: out of device memory"; return false; }
    d_temp = d_temp_bytes;

    cudaStream_t stream = nullptr;      // the build is synchronous: the legacy stream orders everything
    const int grid = static_cast<int>(std::min<size_t>((n + LbvhThreads - 1) / LbvhThreads, size_t(prop.multiProcessorCount) * 8));
    if (cuda_failed(cudaMemcpyAsync(d_boxes, boxes, n * 24, cudaMemcpyHostToDevice, stream), "H2D boxes")) return false;
    if (cuda_failed(cudaMemsetAsync(d_arrivals, 0, (n - 1) * 4, stream), "memset")) return false;
    lbvh_keys_kernel<<<grid, LbvhThreads, 0, stream>>>(d_boxes, count, origin, scale, d_keys, d_ids);
    if (cuda_failed(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, d_keys, d_keys_sorted, d_ids, d_order, static_cast<int>(count), 0, 63, stream),
                    "radix sort")) return false;
    lbvh_hierarchy_kernel<<<grid, LbvhThreads, 0, stream>>>(d_keys_sorted, count, d_left, d_right, d_first, d_last, d_node_parent, d_leaf_parent);
    lbvh_refit_kernel<<<grid, LbvhThreads, 0, stream>>>(d_boxes, d_order, count, d_left, d_right, d_node_parent, d_leaf_parent, d_arrivals, d_node_boxes);
    if (cuda_failed(cudaGetLastError(), "kernel launch")) return false;

    out.order.resize(n); out.left.resize(n - 1); out.right.resize(n - 1); out.first.resize(n - 1); out.last.resize(n - 1);
    out.node_boxes.resize((n - 1) * 6);
    if (cuda_failed(cudaMemcpyAsync(out.order.data(), d_order, n * 4, cudaMemcpyDeviceToHost, stream), "D2H order") ||
        cuda_failed(cudaMemcpyAsync(out.left.data(), d_left, (n - 1) * 4, cudaMemcpyDeviceToHost, stream), "D2H left") ||
        cuda_failed(cudaMemcpyAsync(out.right.data(), d_right, (n - 1) * 4, cudaMemcpyDeviceToHost, stream), "D2H right") ||
        cuda_failed(cudaMemcpyAsync(out.first.data(), d_first, (n - 1) * 4, cudaMemcpyDeviceToHost, stream), "D2H first") ||
        cuda_failed(cudaMemcpyAsync(out.last.data(), d_last, (n - 1) * 4, cudaMemcpyDeviceToHost, stream), "D2H last") ||
        cuda_failed(cudaMemcpyAsync(out.node_boxes.data(), d_node_boxes, (n - 1) * 24, cudaMemcpyDeviceToHost, stream), "D2H boxes") ||
        cuda_failed(cudaStreamSynchronize(stream), "synchronize"))
        return false;
    return true;
}

int lbvh_launch_count() { return 3; }      // keys, hierarchy, refit (+ the library sort's own kernels)

}   // namespace asgpu
