//
// sort.cu -- optional coherence sort of a ray batch (ASGPU_TRACE_SORT, asgpu_sort_rays).
//
// Rays get a 24-bit key: 15 bits of Morton code of the origin (5 bits per axis inside the batch's
// own origin bounds) above 9 bits of Morton code of the direction (3 bits per axis inside the
// batch's direction bounds), so rays that start close together and point the same way end up
// next to each other (origin-major, after Aila & Laine 2009 / Garanzha & Loop 2010).  The trace kernels then pull
// rays through the resulting permutation (KernelArgs::order): a warp's 32 rays walk the same part
// of the tree, which raises SIMT efficiency and L1 hit rate at the price of gathering the rays
// and scattering the hit records.
//
// The sort is a least-significant-digit radix sort, 3 passes of 8 bits, written for this use:
// every warp owns a contiguous run of 2048 elements; pass = per-warp digit histogram ->
// exclusive scan over (digit, warp) -> stable scatter (ranks inside a 32-element round by
// __match_any_sync, rounds in order).  All of it is HBM-bandwidth bound streaming: 16 bytes read +
// 8 written per element and pass.
//

#include "kernels.h"

#include <cuda_runtime.h>

#include <cfloat>

namespace asgpu
{

namespace
{

const int SortThreads = 256;
const int WarpRun = 2048;           // elements per warp
const int Digits = 256;

// Monotone float <-> uint encoding for atomicMin / atomicMax.
__device__ __forceinline__ uint32_t float_key(const float f)
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ float key_float(const uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// bounds[0..5] = min of org xyz, dir xyz; bounds[6..11] = max (encoded).
__global__ void __launch_bounds__(SortThreads)
bounds_kernel(const asgpu_rays rays, const unsigned long long* n_dev, const unsigned long long n_host, uint32_t* bounds)
{
    const unsigned long long n = n_dev ? min(*n_dev, n_host) : n_host;
    float lo[6], hi[6];
    #pragma unroll
    for (int k = 0; k < 6; ++k) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; }
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    {
        #pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            const float o = static_cast<float>(rays.org[i * 3 + k]), d = static_cast<float>(rays.dir[i * 3 + k]);
            if (o == o) { lo[k] = fminf(lo[k], o); hi[k] = fmaxf(hi[k], o); }
            if (d == d) { lo[3 + k] = fminf(lo[3 + k], d); hi[3 + k] = fmaxf(hi[3 + k], d); }
        }
    }
    #pragma unroll
    for (int k = 0; k < 6; ++k)
    {
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0)
        {
            atomicMin(bounds + k, float_key(lo[k]));
            atomicMax(bounds + 6 + k, float_key(hi[k]));
        }
    }
}

__device__ __forceinline__ uint32_t spread3(uint32_t x)     // 0b...cba -> 0b..c..b..a (up to 10 bits)
{
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

__device__ __forceinline__ uint32_t quantise(const float v, const float lo, const float hi, const uint32_t cells)
{
    const float extent = hi - lo;
    if (!(extent > 0.0f) || !(v == v)) return 0;
    const float q = (v - lo) / extent * static_cast<float>(cells);
    return min(cells - 1u, static_cast<uint32_t>(fmaxf(q, 0.0f)));
}

__global__ void __launch_bounds__(SortThreads)
key_kernel(const asgpu_rays rays, const unsigned long long* n_dev, const unsigned long long n_host, const uint32_t* bounds, uint32_t* keys, uint32_t* index)
{
    const unsigned long long n = n_dev ? min(*n_dev, n_host) : n_host;
    float lo[6], hi[6];
    #pragma unroll
    for (int k = 0; k < 6; ++k) { lo[k] = key_float(bounds[k]); hi[k] = key_float(bounds[6 + k]); }
    // Elements past n (device-side counts) sort to the end and keep their own index.
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_host;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    {
        uint32_t key = 0x00FFFFFFu;
        if (i < n)
        {
            uint32_t o = 0, d = 0;
            #pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                o |= spread3(quantise(static_cast<float>(rays.org[i * 3 + k]), lo[k], hi[k], 32u)) << k;
                d |= spread3(quantise(static_cast<float>(rays.dir[i * 3 + k]), lo[3 + k], hi[3 + k], 8u)) << k;
            }
            key = (o << 9) | d;
        }
        keys[i] = key;
        index[i] = static_cast<uint32_t>(i);
    }
}

// Per-warp digit histogram: hist[digit * warps + warp].
__global__ void __launch_bounds__(SortThreads)
histogram_kernel(const uint32_t* __restrict__ keys, const unsigned long long n, const int shift, uint32_t* hist, const unsigned warps)
{
    __shared__ uint32_t counts[SortThreads / 32][Digits];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned long long warp = static_cast<unsigned long long>(blockIdx.x) * (SortThreads / 32) + w;
    for (int d = lane; d < Digits; d += 32) counts[w][d] = 0;
    __syncwarp();
    if (warp < warps)
    {
        const unsigned long long begin = warp * WarpRun, end = min(begin + WarpRun, n);
        for (unsigned long long i = begin + lane; i < end; i += 32)
            atomicAdd(&counts[w][(keys[i] >> shift) & 0xFFu], 1u);
        __syncwarp();
        for (int d = lane; d < Digits; d += 32) hist[static_cast<unsigned long long>(d) * warps + warp] = counts[w][d];
    }
}

// Exclusive scan of `count` values in place, one block: each thread scans a contiguous slice,
// then the slice totals are scanned in shared memory.
__global__ void __launch_bounds__(1024)
scan_kernel(uint32_t* data, const unsigned long long count)
{
    __shared__ uint32_t totals[1024];
    const unsigned long long per = (count + 1023ull) / 1024ull;
    const unsigned long long begin = min(count, per * threadIdx.x), end = min(count, begin + per);
    uint32_t sum = 0;
    for (unsigned long long i = begin; i < end; ++i) sum += data[i];
    totals[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1)
    {
        const uint32_t v = threadIdx.x >= o ? totals[threadIdx.x - o] : 0u;
        __syncthreads();
        totals[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = totals[threadIdx.x] - sum;
    for (unsigned long long i = begin; i < end; ++i) { const uint32_t v = data[i]; data[i] = run; run += v; }
}

__global__ void __launch_bounds__(SortThreads)
scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ index, const unsigned long long n, const int shift,
               const uint32_t* __restrict__ hist, const unsigned warps, uint32_t* keys_out, uint32_t* index_out)
{
    __shared__ uint32_t base[SortThreads / 32][Digits];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned long long warp = static_cast<unsigned long long>(blockIdx.x) * (SortThreads / 32) + w;
    if (warp >= warps) return;
    for (int d = lane; d < Digits; d += 32) base[w][d] = hist[static_cast<unsigned long long>(d) * warps + warp];
    __syncwarp();
    const unsigned long long begin = warp * WarpRun, end = min(begin + WarpRun, n);
    for (unsigned long long r = begin; r < end; r += 32)
    {
        const unsigned long long i = r + lane;
        const bool valid = i < end;
        const uint32_t key = valid ? keys[i] : 0u;
        const uint32_t digit = valid ? (key >> shift) & 0xFFu : 0x100u + lane;      // invalid lanes match nobody
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, digit);
        const unsigned rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t dst = 0;
        if (valid) dst = base[w][digit] + rank;
        __syncwarp();
        if (valid && rank == 0) base[w][digit] += __popc(peers);
        __syncwarp();
        if (valid) { keys_out[dst] = key; index_out[dst] = index[i]; }
    }
}

__global__ void init_bounds_kernel(uint32_t* bounds)
{
    if (threadIdx.x < 6) bounds[threadIdx.x] = 0xFFFFFFFFu;
    else if (threadIdx.x < 12) bounds[threadIdx.x] = 0u;
}

}   // anonymous namespace

size_t ray_sort_workspace_bytes(const size_t n)
{
    const size_t warps = (n + WarpRun - 1) / WarpRun;
    return 4 * n * sizeof(uint32_t) + static_cast<size_t>(Digits) * warps * sizeof(uint32_t) + 256;
}

int ray_sort_launch_count() { return 3 + 3 * 3; }

// Produces the permutation in `order` (n entries) and, when `keys_out` is set, the sorted keys.
// `workspace` holds ray_sort_workspace_bytes(n) bytes.  Returns a cudaError_t value.
int launch_ray_sort(const asgpu_rays& rays, const size_t n, const unsigned long long* n_dev, uint32_t* order, uint32_t* keys_out,
                    void* workspace, const int sm_count, void* stream_)
{
    if (n == 0) return 0;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const unsigned warps = static_cast<unsigned>((n + WarpRun - 1) / WarpRun);
    uint32_t* keys[2]; uint32_t* index[2];
    uint32_t* w = static_cast<uint32_t*>(workspace);
    keys[0] = w; keys[1] = w + n; index[0] = w + 2 * n; index[1] = w + 3 * n;
    uint32_t* hist = w + 4 * n;
    uint32_t* bounds = hist + static_cast<size_t>(Digits) * warps;

    const int grid = sm_count * 8;
    init_bounds_kernel<<<1, 32, 0, stream>>>(bounds);
    bounds_kernel<<<grid, SortThreads, 0, stream>>>(rays, n_dev, n, bounds);
    key_kernel<<<grid, SortThreads, 0, stream>>>(rays, n_dev, n, bounds, keys[0], index[0]);
    const unsigned blocks = (warps + SortThreads / 32 - 1) / (SortThreads / 32);
    int src = 0;
    for (int pass = 0; pass < 3; ++pass)
    {
        const int shift = pass * 8;
        histogram_kernel<<<blocks, SortThreads, 0, stream>>>(keys[src], n, shift, hist, warps);
        scan_kernel<<<1, 1024, 0, stream>>>(hist, static_cast<unsigned long long>(Digits) * warps);
        // The last pass writes the permutation where the caller wants it.
        uint32_t* index_dst = pass == 2 ? order : index[src ^ 1];
        uint32_t* keys_dst = (pass == 2 && keys_out) ? keys_out : keys[src ^ 1];
        scatter_kernel<<<blocks, SortThreads, 0, stream>>>(keys[src], index[src], n, shift, hist, warps, keys_dst, index_dst);
        src ^= 1;
    }
    return static_cast<int>(cudaGetLastError());
}

}   // namespace asgpu
