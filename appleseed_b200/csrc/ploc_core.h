//
// ploc_core.h -- per-element arithmetic of the agglomerative device tree builder (ploc.cu), shared
// with the TEST-ONLY host build (tests/hostsim) the same way lbvh_core.h is.
//
// SURVEY.md section 8(f) rank 4, second step.  The linear BVH of lbvh.cu is built in 0.5 ms per
// million triangles but traces 12-25 % slower than the reference's sweep-SAH tree
// (bvh_sahpartitioner.h:99-170): its splits follow the bits of the Morton code, not the surface
// area.  This builder keeps the Morton ORDER (locality) but chooses the topology by surface area,
// bottom-up: parallel locally-ordered clustering (Meister and Bittner 2018, "Parallel Locally-
// Ordered Clustering for Bounding Volume Hierarchy Construction").  Every round, each cluster looks
// at its `radius` neighbours on either side in the current order and picks the one whose union
// with it has the smallest surface area; pairs that picked each other merge into a new node that
// takes the left partner's place; the survivors are compacted and the next round starts.  Rounds
// shrink the cluster count by a third or so, all clusters in parallel.
//
// Everything here is a pure function of the round's input arrays, with explicit (uncontracted,
// round-to-nearest) float arithmetic and index tie-breaks, so the device kernels and the sequential
// host run build the SAME tree, and repeated builds are identical.
//
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
    #define PLOC_HD __host__ __device__ __forceinline__
#else
    #define PLOC_HD inline
#endif

namespace asgpu
{

const int PlocMaxRadius = 32;

PLOC_HD float ploc_mul(const float a, const float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}

PLOC_HD float ploc_add(const float a, const float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}

PLOC_HD float ploc_min(const float a, const float b) { return a < b ? a : b; }
PLOC_HD float ploc_max(const float a, const float b) { return a > b ? a : b; }

// Half the surface area of the union of two boxes (lo[3], hi[3] each).
PLOC_HD float ploc_union_half_area(const float* a, const float* b)
{
    const float dx = ploc_add(ploc_max(a[3], b[3]), -ploc_min(a[0], b[0]));
    const float dy = ploc_add(ploc_max(a[4], b[4]), -ploc_min(a[1], b[1]));
    const float dz = ploc_add(ploc_max(a[5], b[5]), -ploc_min(a[2], b[2]));
    const float area = ploc_add(ploc_add(ploc_mul(dx, dy), ploc_mul(dy, dz)), ploc_mul(dz, dx));
    return area == area ? area : 3.402823466e38f;       // a NaN (inf - inf, 0 * inf) orders like the largest area
}

// The neighbour cluster i merges best with: the j in [i - radius, i + radius], j != i, with the
// smallest union area under a strict total order on the PAIRS: (area, lower position, higher
// position).  The smallest pair of a round under that order always chooses each other, so every
// round merges at least one pair whatever the boxes are.
// `box_at(k)` returns a pointer to the six floats of cluster k.
template <typename BoxAt>
PLOC_HD uint32_t ploc_nearest(const BoxAt& box_at, const uint32_t count, const uint32_t i, const int radius)
{
    const float* mine = box_at(i);
    uint32_t best = i;
    float best_area = 0.0f;
    // Strict total order on the pairs: (area, lower position, higher position).
    auto better = [&](const float area, const uint32_t j) -> bool
    {
        if (best == i) return true;
        if (area < best_area) return true;
        if (!(area == best_area)) return false;
        const uint32_t lo_new = i < j ? i : j, hi_new = i < j ? j : i;
        const uint32_t lo_old = i < best ? i : best, hi_old = i < best ? best : i;
        return lo_new < lo_old || (lo_new == lo_old && hi_new < hi_old);
    };
    for (int d = 1; d <= radius; ++d)
    {
        if (i >= static_cast<uint32_t>(d))
        {
            const uint32_t j = i - static_cast<uint32_t>(d);
            const float area = ploc_union_half_area(mine, box_at(j));
            if (better(area, j)) { best = j; best_area = area; }
        }
        if (i + static_cast<uint32_t>(d) < count)
        {
            const uint32_t j = i + static_cast<uint32_t>(d);
            const float area = ploc_union_half_area(mine, box_at(j));
            if (better(area, j)) { best = j; best_area = area; }
        }
    }
    return best;
}

// What becomes of cluster i in this round: bit 0 = it survives (alone, or as the merged cluster in
// the left partner's place), bit 32 = it is the left partner of a merging pair (creates a node).
PLOC_HD uint64_t ploc_fate(const uint32_t* nearest, const uint32_t i)
{
    const uint32_t j = nearest[i];
    const bool mutual = j != i && nearest[j] == i;
    if (!mutual) return 1ull;
    return i < j ? ((1ull << 32) | 1ull) : 0ull;
}

}   // namespace asgpu
