//
// lbvh_core.h -- per-element arithmetic of the device tree builder (lbvh.cu), shared with the
// TEST-ONLY host build (tests/hostsim) the same way traverse_core.h is.
//
// the reference builds every triangle tree on one CPU thread with
// a sweep SAH (bvh_sahpartitioner.h:99-170, bvh_builder.h:163-229).  This is the alternative the
// north star's hardware wants: a linear BVH in Morton order of the triangle centroids, all levels
// built at once, one thread per node (Karras 2012, "Maximizing Parallelism in the Construction of
// BVHs, Octrees, and k-d Trees").  The binary tree it yields is emitted in the reference's node
// format (tree_builder.cpp: emit_lbvh), so everything downstream -- motion boxes, leaf payloads,
// the flattener's wide collapse, both kernels -- is unchanged.  The tree differs from the
// reference's; hit records do not (every valid BVH over the same triangles finds the same nearest
// hit, exact-t ties aside).
//
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
    #define LBVH_HD __host__ __device__ __forceinline__
#else
    #define LBVH_HD inline
#endif

namespace asgpu
{

const uint32_t LbvhLeafFlag = 0x80000000u;      // child reference: leaf position | flag, or interior node index

LBVH_HD int lbvh_clz64(const uint64_t v)
{
#if defined(__CUDA_ARCH__)
    return __clzll(static_cast<long long>(v));
#else
    return v == 0 ? 64 : __builtin_clzll(v);
#endif
}

// The low 21 bits of v, two zero bits after each.
LBVH_HD uint64_t lbvh_spread21(uint64_t v)
{
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x001F00000000FFFFull;
    v = (v | (v << 16)) & 0x001F0000FF0000FFull;
    v = (v | (v << 8))  & 0x100F00F00F00F00Full;
    v = (v | (v << 4))  & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2))  & 0x1249249249249249ull;
    return v;
}

// 63-bit Morton code of a box's centre inside the root box.  `box` = lo[3], hi[3]; the key uses
// lo + hi (twice the centre: the quantity the reference sorts by, bvh_bboxsortpredicate.h) and
// `origin` = 2 * root.lo, `scale` = 2^21 / (2 * root extent) (0 on a flat axis).
LBVH_HD uint64_t lbvh_morton(const float* box, const float origin[3], const float scale[3])
{
    uint64_t code = 0;
    for (int a = 0; a < 3; ++a)
    {
        const float f = ((box[a] + box[3 + a]) - origin[a]) * scale[a];
        uint32_t q = f > 0.0f ? (f < 2097151.0f ? static_cast<uint32_t>(f) : 2097151u) : 0u;     // NaN -> 0
        code |= lbvh_spread21(q) << (2 - a);
    }
    return code;
}

// Length of the common prefix of the sorted keys i and j; equal keys are told apart by their
// position (Karras 2012, section 4, "duplicate keys"); -1 outside the array.
LBVH_HD int lbvh_delta(const uint64_t* keys, const int64_t n, const int64_t i, const int64_t j)
{
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a != b) return lbvh_clz64(a ^ b);
    return 64 + lbvh_clz64(static_cast<uint64_t>(i) ^ static_cast<uint64_t>(j));
}

// Interior node i of n - 1 (n >= 2 sorted keys): the key range [first, last] it covers and its two
// children (Karras 2012, algorithm of figure 4).  Node 0 is the root.
LBVH_HD void lbvh_node(const uint64_t* keys, const int64_t n, const int64_t i, uint32_t& left, uint32_t& right, uint32_t& first, uint32_t& last)
{
    const int64_t d = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int64_t lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int64_t l = 0;
    for (int64_t t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int64_t j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int64_t s = 0, t = l;
    do
    {
        t = (t + 1) / 2;
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int64_t gamma = i + s * d + (d < 0 ? -1 : 0);
    const int64_t lo = i < j ? i : j, hi = i < j ? j : i;
    left = static_cast<uint32_t>(gamma) | (lo == gamma ? LbvhLeafFlag : 0u);
    right = static_cast<uint32_t>(gamma + 1) | (hi == gamma + 1 ? LbvhLeafFlag : 0u);
    first = static_cast<uint32_t>(lo);
    last = static_cast<uint32_t>(hi);
}

}   // namespace asgpu
