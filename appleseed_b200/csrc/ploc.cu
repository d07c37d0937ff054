//
// ploc.cu -- triangle-tree topology built on the device by parallel locally-ordered clustering
// (algorithm and arithmetic: ploc_core.h).
//
// Morton keys of the build boxes -> radix sort (the kernels of lbvh.cu) -> rounds of
//   nearest   every cluster finds its best neighbour within `radius` positions (the boxes of a
//             CTA's window are staged in shared memory once: 2 * radius area evaluations each
//             read them from there, not from HBM),
//   fate      mutual pairs are found, every cluster gets (survives, creates a node) bits,
//   scan      one exclusive 64-bit prefix sum gives every survivor its place in the next round and
//             every new node its index (cub::DeviceScan: library code off the trace path),
//   merge     survivors move to their place; the left partner of a pair writes the new node,
// until n / 8 clusters are left; the top of the tree over those is built on the host by the product's
// sweep SAH (build_cluster_top: agglomeration is good at the bottom of a tree and poor at its top);
// then one pass per level of the top and per round, last round first, hands the leaf ranges down
// so that every node covers a contiguous range of the final order (what emit_lbvh lays out in the
// reference's node format).  Node indices are given out from n - 2 downwards, in cluster order
// within a round: the root is node 0, parents have lower indices than their children, and the tree
// is a pure function of the input (no atomics decide anything).
//
// All kernels stream over flat arrays; a round costs ~100 bytes per cluster and the cluster count
// falls by about a third per round, so the whole build moves ~300 bytes per triangle.
//
#include "kernels.h"
#include "lbvh_core.h"
#include "ploc_core.h"
#include "tree_builder.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

namespace asgpu
{

namespace
{

const int PlocThreads = 256;

struct Float3 { float v[3]; };

__global__ void __launch_bounds__(PlocThreads)
ploc_keys_kernel(const float* __restrict__ boxes, const uint32_t n, const Float3 origin, const Float3 scale, uint64_t* __restrict__ keys, uint32_t* __restrict__ ids)
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

// Round 0 input: cluster i = the leaf at sorted position i.
__global__ void __launch_bounds__(PlocThreads)
ploc_init_kernel(const float* __restrict__ boxes, const uint32_t* __restrict__ order, const uint32_t n, float* __restrict__ cbox, uint32_t* __restrict__ cref,
                 uint32_t* __restrict__ ccount)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float* src = boxes + size_t(order[i]) * 6;
        #pragma unroll
        for (int k = 0; k < 6; ++k) cbox[size_t(i) * 6 + k] = __ldg(src + k);
        cref[i] = i | LbvhLeafFlag;
        ccount[i] = 1;
    }
}

// One CTA handles PlocThreads consecutive clusters; the boxes of [first - radius, first +
// PlocThreads + radius) are staged in shared memory.
__global__ void __launch_bounds__(PlocThreads)
ploc_nearest_kernel(const float* __restrict__ cbox, const uint32_t count, const int radius, uint32_t* __restrict__ nearest)
{
    __shared__ float tile[(PlocThreads + 2 * PlocMaxRadius) * 6];
    for (uint32_t first = blockIdx.x * PlocThreads; first < count; first += gridDim.x * PlocThreads)
    {
        const long long lo = static_cast<long long>(first) - radius;
        const int span = PlocThreads + 2 * radius;
        for (int k = threadIdx.x; k < span * 6; k += PlocThreads)
        {
            const long long c = lo + k / 6;
            tile[k] = (c >= 0 && c < static_cast<long long>(count)) ? __ldg(cbox + size_t(c) * 6 + k % 6) : 0.0f;
        }
        __syncthreads();
        const uint32_t i = first + threadIdx.x;
        if (i < count)
        {
            const float* base = tile;
            auto box_at = [base, lo](const uint32_t k) -> const float* { return base + (static_cast<long long>(k) - lo) * 6; };
            nearest[i] = ploc_nearest(box_at, count, i, radius);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(PlocThreads)
ploc_fate_kernel(const uint32_t* __restrict__ nearest, const uint32_t count, unsigned long long* __restrict__ fate)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        fate[i] = ploc_fate(nearest, i);
}

// `offsets` = exclusive prefix sum of `fate`: low word = place among the survivors, high word =
// rank among the round's new nodes.  New node r of the round gets index next_node - r.
__global__ void __launch_bounds__(PlocThreads)
ploc_merge_kernel(const float* __restrict__ cbox, const uint32_t* __restrict__ cref, const uint32_t* __restrict__ ccount, const uint32_t* __restrict__ nearest,
                  const unsigned long long* __restrict__ fate, const unsigned long long* __restrict__ offsets, const uint32_t count, const uint32_t next_node,
                  float* __restrict__ obox, uint32_t* __restrict__ oref, uint32_t* __restrict__ ocount,
                  uint32_t* __restrict__ left, uint32_t* __restrict__ right, uint32_t* __restrict__ leaves, float* __restrict__ node_boxes)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        const unsigned long long f = fate[i];
        if (!(f & 1ull)) continue;
        const unsigned long long o = offsets[i];
        const uint32_t place = static_cast<uint32_t>(o);
        float box[6];
        #pragma unroll
        for (int k = 0; k < 6; ++k) box[k] = cbox[size_t(i) * 6 + k];
        uint32_t ref = cref[i], leaf_count = ccount[i];
        if (f >> 32)
        {
            const uint32_t j = nearest[i];
            const uint32_t node = next_node - static_cast<uint32_t>(o >> 32);
            #pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                box[k] = ploc_min(box[k], cbox[size_t(j) * 6 + k]);
                box[3 + k] = ploc_max(box[3 + k], cbox[size_t(j) * 6 + 3 + k]);
            }
            left[node] = ref;
            right[node] = cref[j];
            leaf_count += ccount[j];
            leaves[node] = leaf_count;
            #pragma unroll
            for (int k = 0; k < 6; ++k) node_boxes[size_t(node) * 6 + k] = box[k];
            ref = node;
        }
        #pragma unroll
        for (int k = 0; k < 6; ++k) obox[size_t(place) * 6 + k] = box[k];
        oref[place] = ref;
        ocount[place] = leaf_count;
    }
}

// Hands the leaf range of the nodes [begin, end) (one round's nodes: no two are related) down to
// their children; a leaf child's final position is the start of its range.
__global__ void __launch_bounds__(PlocThreads)
ploc_ranges_kernel(const uint32_t begin, const uint32_t end, uint32_t* __restrict__ left, uint32_t* __restrict__ right, const uint32_t* __restrict__ leaves,
                   uint32_t* __restrict__ first, uint32_t* __restrict__ last, const uint32_t* __restrict__ sorted_ids, uint32_t* __restrict__ order)
{
    for (uint32_t node = begin + blockIdx.x * blockDim.x + threadIdx.x; node < end; node += gridDim.x * blockDim.x)
    {
        const uint32_t f = first[node], e = last[node];
        const uint32_t l = left[node], r = right[node];
        const uint32_t left_leaves = (l & LbvhLeafFlag) ? 1u : leaves[l];
        if (l & LbvhLeafFlag) { order[f] = sorted_ids[l & ~LbvhLeafFlag]; left[node] = f | LbvhLeafFlag; }
        else { first[l] = f; last[l] = f + left_leaves - 1; }
        if (r & LbvhLeafFlag) { order[e] = sorted_ids[r & ~LbvhLeafFlag]; right[node] = e | LbvhLeafFlag; }
        else { first[r] = f + left_leaves; last[r] = e; }
    }
}

// All device arrays of one build in ONE allocation (22 cudaMalloc / cudaFree pairs cost 0.03 - 0.2 s
// of driver time per build, more than the kernels): sizes are collected first, then carved out.
struct DeviceArena
{
    struct Want { void** where; size_t bytes; };
    std::vector<Want> wants;
    void* base = nullptr;
    ~DeviceArena() { cudaFree(base); }
    template <typename T> void want(T*& p, const size_t elements)
    {
        wants.push_back(Want{ reinterpret_cast<void**>(&p), ((elements ? elements : 1) * sizeof(T) + 255) / 256 * 256 });
    }
    bool commit()
    {
        size_t total = 0;
        for (const Want& w : wants) total += w.bytes;
        if (cudaMalloc(&base, total) != cudaSuccess) { base = nullptr; return false; }
        size_t at = 0;
        for (const Want& w : wants) { *w.where = static_cast<uint8_t*>(base) + at; at += w.bytes; }
        return true;
    }
};

}   // anonymous namespace

int ploc_radius()
{
    int radius = 16;
    if (const char* e = getenv("ASGPU_PLOC_RADIUS")) radius = atoi(e);
    return radius < 1 ? 1 : (radius > PlocMaxRadius ? PlocMaxRadius : radius);
}

// LbvhTopologyFn of the product (default of asgpu_trees_build_on_device): `context` points at the
// int device ordinal.
bool ploc_topology_device(const float* boxes, size_t n, const float root_lo[3], const float root_hi[3], void* context,
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
    const int radius = ploc_radius();

    Float3 origin, scale;
    for (int a = 0; a < 3; ++a)
    {
        const float extent = 2.0f * (root_hi[a] - root_lo[a]);
        origin.v[a] = 2.0f * root_lo[a];
        scale.v[a] = extent > 0.0f ? 2097152.0f / extent : 0.0f;
    }

    const uint32_t count = static_cast<uint32_t>(n);
    const auto t_begin = std::chrono::steady_clock::now();
    DeviceArena mem;
    float *d_boxes, *d_node_boxes, *d_cbox[2];
    uint64_t *d_keys, *d_keys_sorted;
    unsigned long long *d_fate, *d_offsets;
    uint32_t *d_ids, *d_sorted_ids, *d_order, *d_left, *d_right, *d_leaves, *d_first, *d_last, *d_nearest, *d_cref[2], *d_ccount[2];
    uint8_t* d_temp;
    size_t sort_bytes = 0, scan_bytes = 0;
    if (cuda_failed(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, static_cast<const uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr),
                                                    static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), static_cast<int>(count), 0, 63),
                    "radix sort sizing") ||
        cuda_failed(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, static_cast<const unsigned long long*>(nullptr), static_cast<unsigned long long*>(nullptr),
                                                  static_cast<int>(count)), "scan sizing"))
        return false;
    const size_t temp_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    mem.want(d_boxes, n * 6); mem.want(d_node_boxes, (n - 1) * 6); mem.want(d_cbox[0], n * 6); mem.want(d_cbox[1], n * 6);
    mem.want(d_keys, n); mem.want(d_keys_sorted, n); mem.want(d_fate, n); mem.want(d_offsets, n);
    mem.want(d_ids, n); mem.want(d_sorted_ids, n); mem.want(d_order, n); mem.want(d_left, n - 1); mem.want(d_right, n - 1);
    mem.want(d_leaves, n - 1); mem.want(d_first, n - 1); mem.want(d_last, n - 1); mem.want(d_nearest, n);
    mem.want(d_cref[0], n); mem.want(d_cref[1], n); mem.want(d_ccount[0], n); mem.want(d_ccount[1], n); mem.want(d_temp, temp_bytes);
    if (!mem.commit())
    { error = "device tree build: out of device memory"; return false; }

    cudaStream_t stream = nullptr;      // the build is synchronous: the legacy stream orders everything
    const int max_grid = prop.multiProcessorCount * 8;
    auto grid_for = [max_grid](const size_t items) { return static_cast<int>(std::min<size_t>((items + PlocThreads - 1) / PlocThreads, size_t(max_grid))); };

    if (cuda_failed(cudaMemcpyAsync(d_boxes, boxes, n * 24, cudaMemcpyHostToDevice, stream), "H2D boxes")) return false;
    ploc_keys_kernel<<<grid_for(n), PlocThreads, 0, stream>>>(d_boxes, count, origin, scale, d_keys, d_ids);
    if (cuda_failed(cub::DeviceRadixSort::SortPairs(d_temp, sort_bytes, d_keys, d_keys_sorted, d_ids, d_sorted_ids, static_cast<int>(count), 0, 63, stream),
                    "radix sort")) return false;
    ploc_init_kernel<<<grid_for(n), PlocThreads, 0, stream>>>(d_boxes, d_sorted_ids, count, d_cbox[0], d_cref[0], d_ccount[0]);

    // Rounds.  round_begin[k] = index of the first (lowest) node round k created.
    std::vector<uint32_t> round_begin, round_end;
    uint32_t clusters = count, next_node = count - 2;
    int cur = 0;
    // The rounds stop at n / PlocTopRatio clusters: the top of the tree over them is the sweep SAH's
    // (build_cluster_top, tree_builder.cpp), which is good exactly where agglomeration is poor.
    const uint32_t stop_at = std::max<uint32_t>(2u, count / PlocTopRatio);
    while (clusters > stop_at)
    {
        const int grid = grid_for(clusters);
        ploc_nearest_kernel<<<grid, PlocThreads, 0, stream>>>(d_cbox[cur], clusters, radius, d_nearest);
        ploc_fate_kernel<<<grid, PlocThreads, 0, stream>>>(d_nearest, clusters, d_fate);
        if (cuda_failed(cub::DeviceScan::ExclusiveSum(d_temp, scan_bytes, d_fate, d_offsets, static_cast<int>(clusters), stream), "scan")) return false;
        ploc_merge_kernel<<<grid, PlocThreads, 0, stream>>>(d_cbox[cur], d_cref[cur], d_ccount[cur], d_nearest, d_fate, d_offsets, clusters, next_node,
                                                            d_cbox[cur ^ 1], d_cref[cur ^ 1], d_ccount[cur ^ 1], d_left, d_right, d_leaves, d_node_boxes);
        unsigned long long tail[2];
        if (cuda_failed(cudaMemcpyAsync(&tail[0], d_offsets + (clusters - 1), 8, cudaMemcpyDeviceToHost, stream), "D2H round totals") ||
            cuda_failed(cudaMemcpyAsync(&tail[1], d_fate + (clusters - 1), 8, cudaMemcpyDeviceToHost, stream), "D2H round totals") ||
            cuda_failed(cudaStreamSynchronize(stream), "round"))
            return false;
        const unsigned long long total = tail[0] + tail[1];
        const uint32_t survivors = static_cast<uint32_t>(total), created = static_cast<uint32_t>(total >> 32);
        if (created == 0 || survivors + created != clusters) { error = "device tree build: a clustering round made no progress"; return false; }
        round_begin.push_back(next_node - (created - 1));
        round_end.push_back(next_node + 1);
        next_node -= created;
        clusters = survivors;
        cur ^= 1;
    }

    const bool timing = getenv("ASGPU_BUILD_TIMING") != nullptr;
    const auto t_rounds = std::chrono::steady_clock::now();
    if (timing) fprintf(stderr, "asgpu build:   clustering rounds %zu, %u clusters left   %.3f s\n", round_begin.size(), clusters,
                        std::chrono::duration<double>(t_rounds - t_begin).count());

    // Top of the tree on the host: nodes 0 .. clusters - 2, breadth first.
    ClusterTop top;
    {
        std::vector<float> h_cbox(size_t(clusters) * 6);
        std::vector<uint32_t> h_cref(clusters), h_ccount(clusters);
        if (cuda_failed(cudaMemcpyAsync(h_cbox.data(), d_cbox[cur], size_t(clusters) * 24, cudaMemcpyDeviceToHost, stream), "D2H clusters") ||
            cuda_failed(cudaMemcpyAsync(h_cref.data(), d_cref[cur], size_t(clusters) * 4, cudaMemcpyDeviceToHost, stream), "D2H clusters") ||
            cuda_failed(cudaMemcpyAsync(h_ccount.data(), d_ccount[cur], size_t(clusters) * 4, cudaMemcpyDeviceToHost, stream), "D2H clusters") ||
            cuda_failed(cudaStreamSynchronize(stream), "clusters"))
            return false;
        if (!build_cluster_top(h_cbox.data(), h_cref.data(), h_ccount.data(), clusters, static_cast<int>(std::max(1u, std::thread::hardware_concurrency())), top, error)) return false;
        const uint32_t top_nodes = clusters - 1;
        if (next_node != top_nodes - 1) { error = "device tree build: node numbering of the rounds and of the top do not meet"; return false; }
        if (cuda_failed(cudaMemcpyAsync(d_left, top.left.data(), size_t(top_nodes) * 4, cudaMemcpyHostToDevice, stream), "H2D top") ||
            cuda_failed(cudaMemcpyAsync(d_right, top.right.data(), size_t(top_nodes) * 4, cudaMemcpyHostToDevice, stream), "H2D top") ||
            cuda_failed(cudaMemcpyAsync(d_leaves, top.leaves.data(), size_t(top_nodes) * 4, cudaMemcpyHostToDevice, stream), "H2D top") ||
            cuda_failed(cudaMemcpyAsync(d_node_boxes, top.boxes.data(), size_t(top_nodes) * 24, cudaMemcpyHostToDevice, stream), "H2D top"))
            return false;
    }

    // Leaf ranges, root first (the two words stay alive until the final synchronisation below).
    const uint32_t root_range[2] = { 0u, count - 1 };
    if (cuda_failed(cudaMemcpyAsync(d_first, &root_range[0], 4, cudaMemcpyHostToDevice, stream), "H2D root range") ||
        cuda_failed(cudaMemcpyAsync(d_last, &root_range[1], 4, cudaMemcpyHostToDevice, stream), "H2D root range"))
        return false;
    if (timing) fprintf(stderr, "asgpu build:   sweep SAH over the clusters (host)          %.3f s\n",
                        std::chrono::duration<double>(std::chrono::steady_clock::now() - t_rounds).count());
    for (size_t l = 0; l + 1 < top.level_begin.size(); ++l)     // the top, level by level (parents first)
        ploc_ranges_kernel<<<grid_for(top.level_begin[l + 1] - top.level_begin[l]), PlocThreads, 0, stream>>>(top.level_begin[l], top.level_begin[l + 1], d_left, d_right,
                                                                                                              d_leaves, d_first, d_last, d_sorted_ids, d_order);
    for (size_t k = round_begin.size(); k-- > 0; )              // then the rounds, last round first
        ploc_ranges_kernel<<<grid_for(round_end[k] - round_begin[k]), PlocThreads, 0, stream>>>(round_begin[k], round_end[k], d_left, d_right, d_leaves,
                                                                                                d_first, d_last, d_sorted_ids, d_order);
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

}   // namespace asgpu
