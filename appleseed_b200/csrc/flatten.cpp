//
// flatten.cpp -- host flattener: the reference's bvh::Tree arrays -> GPU blob.
//
// Input is exactly what appleseed's own classes hold (include/asgpu.h, *_view structs):
//   bvh::Tree::m_nodes / m_node_bboxes           foundation/math/bvh/bvh_tree.h:77-78
//   TriangleTree::m_triangle_keys / m_leaf_data  renderer/kernel/intersection/triangletree.h:116-125
//   AssemblyTree nodes + items                   renderer/kernel/intersection/assemblytree.h:98-134
// Leaf payloads are decoded as TriangleLeafVisitor reads them (triangletree.cpp:1363-1383,
// layout written by triangleencoder.cpp:72-103).
//
// Output: see gpu_layout.h.  The EXACT layout is a lossless re-packing.  The WIDE layout
// collapses each binary tree into 8-wide nodes (SAH-optimal: the dynamic programme of Ylitie, Karras, Laine 2017, section 4.1), orders children
// into octant slots, stores child boxes quantised to 8 bits per plane ROUNDED OUTWARD, and
// re-orders triangle records so that one node's leaves are contiguous.  Because every wide box
// contains the binary boxes it replaces, a traversal that tests wide boxes conservatively can only
// reach MORE leaves than the reference traversal, never fewer.
//

#include "flatten.h"
#include "as_format.h"
#include "parallel.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <mutex>

namespace asgpu
{

namespace
{

struct FBox
{
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = std::numeric_limits<float>::max(); hi[a] = -std::numeric_limits<float>::max(); } }
    void grow(const FBox& b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    void grow(const float p[3]) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    bool valid() const { return lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]; }
    float half_area() const
    {
        if (!valid()) return 0.0f;
        const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
        return x * y + x * z + y * z;
    }
};

inline float float_below(const double v)    // largest float <= v
{
    float f = static_cast<float>(v);
    if (static_cast<double>(f) > v) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
    return f;
}

inline float float_above(const double v)    // smallest float >= v
{
    float f = static_cast<float>(v);
    if (static_cast<double>(f) < v) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
    return f;
}

class BlobWriter
{
  public:
    explicit BlobWriter(HostBlob& bytes) : m_bytes(bytes) { m_bytes.clear(); }

    // Room for `size` bytes at the next SectionAlign boundary (the gap is zeroed, the room is not);
    // returns the offset.
    uint64_t reserve(const size_t size)
    {
        const size_t old = m_bytes.size();
        const uint64_t offset = (old + SectionAlign - 1) / SectionAlign * SectionAlign;
        m_bytes.grow_to(offset + size);
        if (offset > old) std::memset(m_bytes.data() + old, 0, offset - old);
        return offset;
    }

    // Appends `size` bytes at the next SectionAlign boundary; returns the offset.
    uint64_t append(const void* data, const size_t size)
    {
        const uint64_t offset = reserve(size);
        parallel_memcpy(m_bytes.data() + offset, data, size, host_threads());
        return offset;
    }

    template <typename T> uint64_t append(const std::vector<T>& v)
    {
        return append(v.empty() ? nullptr : v.data(), v.size() * sizeof(T));
    }

    template <typename T> uint64_t append(const RawVector<T>& v)
    {
        return append(v.empty() ? nullptr : v.data(), v.size() * sizeof(T));
    }

    uint8_t* at(const uint64_t offset) { return m_bytes.data() + offset; }

    void finish() { reserve(0); }

  private:
    HostBlob& m_bytes;
};

//
// EXACT layout of one triangle tree.
//

struct ExactTree
{
    RawVector<BNodeF>       bnodes;
    RawVector<MNode>        mnodes;
    RawVector<MBox>         mboxes;
    RawVector<TriRecord>    tris;
    RawVector<float>        poses;
    RawVector<HitKey>       keys;
    uint64_t                moving = 0;
};

// Shape of a binary tree in the reference's layout, whichever GPU layout it is flattened into:
// bvh::Builder appends the two children of a node after every node that exists so far
// (bvh_builder.h:197-205, bvh_spatialbuilder.h:209-221; so does the optional node reordering of
// foundation/math/treeoptimizer.h), so child indices grow towards the leaves -- which rules out cycles --
// and the exact traversal keeps the reference's 64-entry stack, one entry per level at most
// (intersectionsettings.h:95; the reference itself does not check).
int check_hierarchy(const AsNode* nodes, const uint64_t count, const char* what, std::string& error)
{
    std::vector<uint8_t> depth(count, 0);
    depth[0] = 1;
    for (uint64_t n = 0; n < count; ++n)
    {
        if (!nodes[n].interior() || depth[n] == 0) continue;       // depth 0: not reachable from the root
        const uint64_t child = nodes[n].index;
        if (child <= n || child + 1 >= count) { error = std::string(what) + " is not in parent-before-child order"; return ASGPU_E_INVALID; }
        if (depth[child] != 0 || depth[child + 1] != 0) { error = std::string(what) + " node has two parents"; return ASGPU_E_INVALID; }
        if (depth[n] >= 64) { error = std::string(what) + " deeper than the 64-entry traversal stack"; return ASGPU_E_UNSUPPORTED; }
        depth[child] = depth[child + 1] = static_cast<uint8_t>(depth[n] + 1);
    }
    return ASGPU_OK;
}

// ASGPU_BUILD_TIMING=1: seconds since the previous mark, on stderr.
struct PhaseTimer
{
    const bool enabled = std::getenv("ASGPU_BUILD_TIMING") != nullptr;
    std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    void mark(const char* what)
    {
        if (!enabled) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "asgpu flatten:   %-20s %.3f s\n", what, std::chrono::duration<double>(now - last).count());
        last = now;
    }
};

// First error of a parallel section (later ones are dropped).
class ErrorSink
{
  public:
    bool failed() const { return m_failed.load(std::memory_order_relaxed); }
    void fail(const int code, const char* message)
    {
        std::lock_guard<std::mutex> lock(m_mutex);
        if (!m_failed.load()) { m_code = code; m_message = message; m_failed.store(true); }
    }
    int finish(std::string& error) const { if (failed()) error = m_message; return failed() ? m_code : ASGPU_OK; }

  private:
    std::atomic<bool>   m_failed{false};
    std::mutex          m_mutex;
    int                 m_code = ASGPU_OK;
    std::string         m_message;
};

// Where the payload of a leaf lives: in the node when its first u32 is ~0, else at m_leaf_data[offset].
inline bool leaf_payload(const asgpu_triangle_tree_view& v, const AsNode& src, const uint8_t*& p, const uint8_t*& limit)
{
    const uint8_t* user = src.user_data();
    uint32_t offset;
    std::memcpy(&offset, user, 4);
    if (src.item_count == 0) { p = limit = user; }
    else if (offset == 0xFFFFFFFFu) { p = user + 4; limit = user + AsNodeUserDataSize; }
    else
    {
        if (offset >= v.leaf_data_size) return false;
        p = v.leaf_data + offset; limit = v.leaf_data + v.leaf_data_size;
    }
    return true;
}

// Nodes are decoded in parallel over contiguous node ranges; the pose pool keeps the reference's leaf
// order because every range knows where its poses start (prefix sum of a first, counting pass).
int decode_tree(const asgpu_triangle_tree_view& v, ExactTree& out, std::string& error)
{
    if (v.node_count == 0 || !v.nodes) { error = "triangle tree without nodes"; return ASGPU_E_INVALID; }
    if (v.node_count >= 0xFFFFFFFFull || v.triangle_key_count >= 0xFFFFFFFFull) { error = "triangle tree too large"; return ASGPU_E_UNSUPPORTED; }
    const AsNode* nodes = static_cast<const AsNode*>(v.nodes);
    const AsTriangleKey* keys = static_cast<const AsTriangleKey*>(v.triangle_keys);
    const bool motion = v.moving_triangle_count > 0;
    const int shape = check_hierarchy(nodes, v.node_count, "triangle tree", error);
    if (shape != ASGPU_OK) return shape;
    const int threads = host_threads();
    const size_t grain = 1 << 15;
    PhaseTimer phase;
    phase.mark("hierarchy check");

    out.bnodes.resize(v.node_count);
    if (motion) out.mnodes.resize(v.node_count);
    out.tris.resize(v.triangle_key_count);
    out.keys.resize(v.triangle_key_count);
    parallel_chunks(v.triangle_key_count, threads, grain, [&](int, size_t begin, size_t end)
    {
        for (size_t i = begin; i < end; ++i)
        {
            out.keys[i].object_instance_index = keys[i].object_instance_index;
            out.keys[i].triangle_index = keys[i].triangle_index;
        }
    });

    ErrorSink sink;
    if (motion)
    {
        out.mboxes.resize(v.node_bbox_count);
        parallel_chunks(v.node_bbox_count, threads, grain, [&](int, size_t begin, size_t end)
        {
            for (size_t i = begin; i < end; ++i)
                for (int k = 0; k < 6; ++k)
                {
                    const double d = v.node_bboxes[i * 6 + k];
                    const float f = static_cast<float>(d);
                    if (static_cast<double>(f) != d) { sink.fail(ASGPU_E_UNSUPPORTED, "motion box is not float-representable"); return; }
                    out.mboxes[i].v[k] = f;
                }
        });
        if (sink.failed()) return sink.finish(error);
    }

    phase.mark("alloc + keys");
    // Pass 1: pose floats of every node range (walks the leaf payload headers only).
    const int chunks = chunk_count(v.node_count, threads, grain);
    std::vector<uint64_t> pose_floats(chunks + 1, 0);
    parallel_chunks(v.node_count, threads, grain, [&](int chunk, size_t begin, size_t end)
    {
        uint64_t total = 0;
        for (size_t n = begin; n < end && !sink.failed(); ++n)
        {
            const AsNode& src = nodes[n];
            if (src.interior()) continue;
            const uint8_t* p; const uint8_t* limit;
            if (!leaf_payload(v, src, p, limit)) { sink.fail(ASGPU_E_INVALID, "leaf data offset out of range"); return; }
            for (uint32_t j = 0; j < src.item_count; ++j)
            {
                if (p + 8 > limit) { sink.fail(ASGPU_E_INVALID, "truncated leaf payload"); return; }
                uint32_t msc;
                std::memcpy(&msc, p + 4, 4); p += 8;
                const size_t bytes = msc == 0 ? size_t(AsTriangleBytes) : (size_t(msc) + 1) * AsPoseBytes;
                if (p + bytes > limit) { sink.fail(ASGPU_E_INVALID, "truncated leaf payload"); return; }
                p += bytes;
                if (msc != 0) total += bytes / 4;
            }
        }
        pose_floats[chunk + 1] = total;
    });
    if (sink.failed()) return sink.finish(error);
    for (int c = 0; c < chunks; ++c) pose_floats[c + 1] += pose_floats[c];
    if (pose_floats[chunks] >= 0xFFFFFFFEull) { error = "pose pool too large"; return ASGPU_E_UNSUPPORTED; }
    out.poses.resize(pose_floats[chunks]);

    phase.mark("pass 1");
    // Pass 2: the records.
    RawVector<uint8_t> seen;
    seen.assign(v.triangle_key_count, uint8_t(0), threads);
    std::vector<uint64_t> moving(chunks, 0);
    parallel_chunks(v.node_count, threads, grain, [&](int chunk, size_t begin, size_t end)
    {
        uint64_t pose_cursor = pose_floats[chunk];
        for (size_t n = begin; n < end && !sink.failed(); ++n)
        {
            const AsNode& src = nodes[n];
            BNodeF& dst = out.bnodes[n];
            std::memset(&dst, 0, sizeof(dst));
            dst.index = src.index;
            dst.item_count = src.item_count;

            if (src.interior())
            {
                if (uint64_t(src.index) + 1 >= v.node_count) { sink.fail(ASGPU_E_INVALID, "child node index out of range"); return; }
                for (int k = 0; k < 12; ++k)
                {
                    const float f = static_cast<float>(src.bbox[k]);
                    // The reference builds triangle-tree boxes in float and widens them
                    // (triangletree.cpp:543, bvh_builder.h:193-204).
                    if (static_cast<double>(f) != src.bbox[k]) { sink.fail(ASGPU_E_UNSUPPORTED, "triangle tree box is not float-representable"); return; }
                    dst.box[k] = f;
                }
                if (motion)
                {
                    MNode& m = out.mnodes[n];
                    m.left_index = src.left_bbox_index; m.left_count = src.left_bbox_count;
                    m.right_index = src.right_bbox_index; m.right_count = src.right_bbox_count;
                    if ((m.left_count > 1 && uint64_t(m.left_index) + m.left_count > v.node_bbox_count) ||
                        (m.right_count > 1 && uint64_t(m.right_index) + m.right_count > v.node_bbox_count) ||
                        m.left_count == 0 || m.right_count == 0)
                    { sink.fail(ASGPU_E_INVALID, "motion box range out of range"); return; }
                }
                continue;
            }

            if (uint64_t(src.index) + src.item_count > v.triangle_key_count) { sink.fail(ASGPU_E_INVALID, "leaf item range out of range"); return; }
            if (motion) std::memset(&out.mnodes[n], 0, sizeof(MNode));
            const uint8_t* p = nullptr; const uint8_t* limit = nullptr;
            leaf_payload(v, src, p, limit);
            for (uint32_t j = 0; j < src.item_count; ++j)
            {
                const uint32_t slot = src.index + j;
                if (__atomic_exchange_n(&seen[slot], uint8_t(1), __ATOMIC_RELAXED)) { sink.fail(ASGPU_E_INVALID, "triangle slot referenced twice"); return; }
                uint32_t vis, msc;
                std::memcpy(&vis, p, 4); std::memcpy(&msc, p + 4, 4); p += 8;
                TriRecord& tri = out.tris[slot];
                std::memset(&tri, 0, sizeof(tri));
                tri.vis_flags = vis;
                tri.ref_slot = slot;
                if (msc == 0)
                {
                    std::memcpy(tri.v0, p, AsTriangleBytes); p += AsTriangleBytes;
                    tri.motion = 0;
                }
                else
                {
                    const size_t bytes = (size_t(msc) + 1) * AsPoseBytes;
                    tri.motion = static_cast<uint32_t>(pose_cursor) + 1;
                    std::memcpy(&tri.v0[0], &msc, 4);
                    std::memcpy(out.poses.data() + pose_cursor, p, bytes); p += bytes;
                    pose_cursor += bytes / 4;
                    ++moving[chunk];
                }
            }
        }
    });
    if (sink.failed()) return sink.finish(error);
    phase.mark("pass 2");
    out.moving = 0;
    for (const uint64_t m : moving) out.moving += m;
    parallel_chunks(v.triangle_key_count, threads, grain, [&](int, size_t begin, size_t end)
    {
        for (size_t i = begin; i < end; ++i)
            if (!seen[i]) { sink.fail(ASGPU_E_INVALID, "triangle slot not referenced by any leaf"); return; }
    });
    if (sink.failed()) return sink.finish(error);
    if ((out.moving > 0) != motion) { error = "moving_triangle_count does not match the leaf payloads"; return ASGPU_E_INVALID; }
    return ASGPU_OK;
}

FBox triangle_box(const ExactTree& t, const uint32_t slot)
{
    FBox b; b.reset();
    const TriRecord& tri = t.tris[slot];
    if (tri.motion == 0)
    {
        const float v1[3] = { tri.v0[0] + tri.e0[0], tri.v0[1] + tri.e0[1], tri.v0[2] + tri.e0[2] };
        const float v2[3] = { tri.v0[0] + tri.e1[0], tri.v0[1] + tri.e1[1], tri.v0[2] + tri.e1[2] };
        b.grow(tri.v0); b.grow(v1); b.grow(v2);
        // v0 + e is rounded; pad by an ulp-scale margin so the box is a true bound.
        for (int a = 0; a < 3; ++a)
        {
            b.lo[a] = std::nextafterf(b.lo[a], -std::numeric_limits<float>::infinity());
            b.hi[a] = std::nextafterf(b.hi[a], std::numeric_limits<float>::infinity());
        }
    }
    else
    {
        uint32_t msc; std::memcpy(&msc, &tri.v0[0], 4);
        const float* p = t.poses.data() + (tri.motion - 1);
        for (uint32_t k = 0; k < (msc + 1) * 3; ++k) b.grow(p + k * 3);
        // Interpolated vertices are computed in float with two roundings (triangletree.cpp:1438-1445).
        for (int a = 0; a < 3; ++a)
        {
            const float m = std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a])) * 4.0e-7f + std::numeric_limits<float>::min();
            b.lo[a] -= m; b.hi[a] += m;
        }
    }
    return b;
}

//
// WIDE layout: generic collapse of a binary tree whose leaves hold item ranges.
//

struct BinaryView
{
    uint64_t                    node_count;
    RawVector<uint32_t>         child;          // first child or InteriorMark^... (leaf: InteriorMark)
    RawVector<uint32_t>         first, count;   // leaf item range
    RawVector<FBox>             box;            // conservative all-time box of each node
    // Trees with motion: box of node x over the ray-time interval [t0, t1] (empty = use `box`).
    std::function<FBox(uint32_t x, float t0, float t1)> timed_box;
};

struct WideOut
{
    std::vector<WSlice>     slices;         // nodes.size() * slice_count entries (trees with motion)
    std::vector<WNode>      nodes;
    std::vector<uint32_t>   leaf_order;     // wide item order -> original item (slot) index
    uint32_t                depth = 0;      // number of node levels (root = 1)
};

struct Element
{
    bool        is_node;        // subtree that becomes a wide node of its own, else an item range (leaf child)
    uint32_t    node;           // node of the augmented binary tree (is_node)
    uint32_t    src;            // node of the ORIGINAL binary tree whose box this is
    uint32_t    first, count;   // item range (!is_node)
    FBox        box;
};

struct Work { uint32_t node; int budget; };

int quantise_node(WNode& w, const Element* kids, const int n, const int* slot_of, std::string& error)
{
    FBox nb; nb.reset();
    for (int i = 0; i < n; ++i) nb.grow(kids[i].box);
    if (!nb.valid()) { for (int a = 0; a < 3; ++a) nb.lo[a] = nb.hi[a] = 0.0f; }

    for (int a = 0; a < 3; ++a)
    {
        w.origin[a] = nb.lo[a];
        const double extent = double(nb.hi[a]) - double(nb.lo[a]);
        int e = -126;
        if (extent > 0.0)
        {
            e = static_cast<int>(std::ceil(std::log2(extent / 255.0)));
            while (std::ceil(extent / std::ldexp(1.0, e)) > 255.0) ++e;
            while (e > -126 && std::ceil(extent / std::ldexp(1.0, e - 1)) <= 255.0) --e;
        }
        if (e < -126) e = -126;
        // The kernels form 2^(e + 15) by adding to the exponent field.
        if (e > 127 - 16) { error = "scene extent too large to quantise"; return ASGPU_E_UNSUPPORTED; }
        w.exp[a] = static_cast<uint8_t>(e + 127);
        const double scale = std::ldexp(1.0, e);
        for (int i = 0; i < n; ++i)
        {
            const int s = slot_of[i];
            if (!kids[i].box.valid()) { w.qlo[a][s] = 255; w.qhi[a][s] = 0; continue; }
            double lo = std::floor((double(kids[i].box.lo[a]) - double(nb.lo[a])) / scale);
            double hi = std::ceil((double(kids[i].box.hi[a]) - double(nb.lo[a])) / scale);
            lo = std::min(255.0, std::max(0.0, lo));
            hi = std::min(255.0, std::max(lo, hi));
            w.qlo[a][s] = static_cast<uint8_t>(lo);
            w.qhi[a][s] = static_cast<uint8_t>(hi);
            // Outward-rounding invariants.
            if (double(nb.lo[a]) + lo * scale > double(kids[i].box.lo[a]) || double(nb.lo[a]) + hi * scale < double(kids[i].box.hi[a]))
            { error = "internal error: quantised box does not contain its child"; return ASGPU_E_INVALID; }
        }
    }
    return ASGPU_OK;
}

// Child planes of one time slice in the frame quantise_node chose (rounded outward, clamped into
// the all-time planes, which contain them).
int quantise_slice(const WNode& w, WSlice& out, const FBox* boxes, const int n, const int* slot_of, std::string& error)
{
    for (int a = 0; a < 3; ++a)
    {
        for (int sl = 0; sl < 8; ++sl) { out.qlo[a][sl] = 255; out.qhi[a][sl] = 0; }
        const double scale = std::ldexp(1.0, int(w.exp[a]) - 127);
        for (int i = 0; i < n; ++i)
        {
            const int s = slot_of[i];
            if (!boxes[i].valid()) continue;
            double lo = std::floor((double(boxes[i].lo[a]) - double(w.origin[a])) / scale);
            double hi = std::ceil((double(boxes[i].hi[a]) - double(w.origin[a])) / scale);
            lo = std::min(double(w.qhi[a][s]), std::max(double(w.qlo[a][s]), lo));
            hi = std::min(double(w.qhi[a][s]), std::max(lo, hi));
            out.qlo[a][s] = static_cast<uint8_t>(lo);
            out.qhi[a][s] = static_cast<uint8_t>(hi);
            // A slice plane may only be clamped where the all-time plane already bounds the child.
            const double plo = double(w.origin[a]) + lo * scale, phi = double(w.origin[a]) + hi * scale;
            const double alo = double(w.origin[a]) + double(w.qlo[a][s]) * scale, ahi = double(w.origin[a]) + double(w.qhi[a][s]) * scale;
            if ((plo > double(boxes[i].lo[a]) && plo > alo) || (phi < double(boxes[i].hi[a]) && phi < ahi))
            { error = "internal error: time-slice box does not contain its child"; return ASGPU_E_INVALID; }
        }
    }
    return ASGPU_OK;
}

// Relative costs of the collapse's surface-area heuristic: testing one wide node / one leaf item.
// The defaults are the measured optimum for the wide kernels (profiles/README.md); the
// environment variable ASGPU_COLLAPSE_ITEM_COST overrides the item cost for experiments.
struct CollapseCosts { double node, item; };

// Time slices per wide node of a tree with moving triangles (ASGPU_TIME_SLICES overrides; 0 = none).
uint32_t time_slices()
{
    uint32_t t = 16;        // measured on C4: 8 -> 1784, 16 -> 1976, 32 -> 2050 Mrays/s (no slices: 567); 768 B per node at 16
    if (const char* e = std::getenv("ASGPU_TIME_SLICES")) { const int v = std::atoi(e); t = v < 0 ? 0u : static_cast<uint32_t>(std::min(v, 64)); }
    return t;
}

CollapseCosts collapse_costs()
{
    CollapseCosts c = { 1.0, 1.0 };
    if (const char* e = std::getenv("ASGPU_COLLAPSE_ITEM_COST")) { const double v = std::atof(e); if (v > 0.0) c.item = v; }
    return c;
}

// Collapses a binary tree into 8-wide nodes.  Which binary subtrees become wide nodes, which
// become leaf children (up to leaf_cap items) and how the 8 child slots of a wide node are spent
// is decided by the dynamic programme of Ylitie, Karras, Laine 2017 (section 4.1) minimising the
// expected SAH cost: C(n, i) = cheapest way to cover subtree n with at most i children of one
// wide node.  Every wide child box is the box of a binary subtree, so it contains all the binary
// boxes below it.
int collapse(const BinaryView& bv, const uint32_t leaf_cap, const CollapseCosts costs, WideOut& out, std::string& error, const uint32_t slice_count = 0)
{
    out.slices.clear();
    out.nodes.clear();
    out.leaf_order.clear();
    out.depth = 0;

    // ---- augmented binary tree: leaves larger than leaf_cap are halved into virtual nodes --------
    const uint32_t Leaf = InteriorMark;
    const int threads = host_threads();
    PhaseTimer phase;
    RawVector<uint32_t> lc, rc, first, count, src;                  // lc: left child (right = left + 1 for real nodes)
    RawVector<FBox> box;
    lc.copy_from(bv.child.data(), bv.node_count, threads);
    first.copy_from(bv.first.data(), bv.node_count, threads);
    count.copy_from(bv.count.data(), bv.node_count, threads);
    box.copy_from(bv.box.data(), bv.node_count, threads);
    rc.resize(bv.node_count); src.resize(bv.node_count);
    parallel_chunks(bv.node_count, threads, 1 << 16, [&](int, size_t begin, size_t end)
    {
        for (size_t i = begin; i < end; ++i) { src[i] = static_cast<uint32_t>(i); rc[i] = lc[i] == Leaf ? Leaf : lc[i] + 1; }
    });
    for (uint64_t i = 0; i < lc.size(); ++i)
    {
        if (lc[i] != Leaf || count[i] <= leaf_cap) continue;
        // An over-full leaf (max_leaf_size > cap, or a range the SAH refused to split): both halves keep its box.
        const uint32_t a = static_cast<uint32_t>(lc.size());
        const uint32_t half = count[i] / 2;
        lc.push_back(Leaf); rc.push_back(Leaf); first.push_back(first[i]); count.push_back(half); box.push_back(box[i]); src.push_back(src[i]);
        lc.push_back(Leaf); rc.push_back(Leaf); first.push_back(first[i] + half); count.push_back(count[i] - half); box.push_back(box[i]); src.push_back(src[i]);
        lc[i] = a; rc[i] = a + 1;
    }
    const uint64_t n = lc.size();
    if (n >= 0x7FFFFFFFull) { error = "binary tree too large to collapse"; return ASGPU_E_UNSUPPORTED; }

    // ---- bottom-up: item range of every subtree, cost tables --------------------------------------
    // Children have larger indices than their parent (bvh_builder.h:197-205; virtual nodes are
    // appended), so one ascending sweep gives every node its depth; the dynamic programme then runs
    // level by level from the deepest one, the nodes of a level in parallel.
    phase.mark("augment");
    const double Inf = std::numeric_limits<double>::infinity();
    RawVector<uint32_t> level_of;
    level_of.assign(n, 0u, threads);
    uint32_t max_level = 0;
    for (uint64_t x = 0; x < n; ++x)
    {
        if (lc[x] == Leaf) continue;
        const uint64_t l = lc[x], r = rc[x];
        if (l <= x || r <= x || l >= n || r >= n) { error = "binary tree is not in parent-before-child order"; return ASGPU_E_INVALID; }
        // (max: in a tree every node has one parent, but a malformed input that slipped a shared child
        // past check_hierarchy through an unreachable node must still be processed children first.)
        level_of[l] = std::max(level_of[l], level_of[x] + 1);
        level_of[r] = std::max(level_of[r], level_of[x] + 1);
        max_level = std::max(max_level, level_of[x] + 1);
    }
    std::vector<uint64_t> level_begin(size_t(max_level) + 2, 0);
    for (uint64_t x = 0; x < n; ++x) ++level_begin[level_of[x] + 1];
    for (size_t d = 0; d <= max_level; ++d) level_begin[d + 1] += level_begin[d];
    RawVector<uint32_t> by_level(n);
    {
        std::vector<uint64_t> cursor(level_begin.begin(), level_begin.end() - 1);
        for (uint64_t x = 0; x < n; ++x) by_level[cursor[level_of[x]]++] = static_cast<uint32_t>(x);
    }
    level_of.release();
    phase.mark("levels");

    // (Uninitialised: every entry that is read was written by cost_tables -- entry 0 of a row never is.)
    RawVector<uint32_t> sub_first(n), sub_count(n);
    RawVector<uint8_t> contiguous(n);
    RawVector<float> cost(n * 8);           // cost[x * 8 + i], i = 1..7
    RawVector<uint8_t> choice;              // i = 1: 0 leaf child, 1 wide node;  i > 1: 0 = same as i - 1, else children given to the left subtree
    choice.assign(n * 8, uint8_t(0), threads);
    RawVector<uint8_t> split8(n);           // children given to the left subtree when x becomes a wide node
    auto area_of = [&](const uint64_t x) -> double
    {
        if (!box[x].valid()) return 0.0;
        const double a = box[x].half_area();
        return a < 1.0e30 ? a : 1.0e30;
    };
    auto cost_tables = [&](const uint64_t x)
    {
        float* cx = &cost[x * 8];
        uint8_t* hx = &choice[x * 8];
        const double area = area_of(x);
        if (lc[x] == Leaf)
        {
            sub_first[x] = first[x]; sub_count[x] = count[x]; contiguous[x] = 1;
            for (int i = 1; i < 8; ++i) { cx[i] = static_cast<float>(area * count[x] * costs.item); hx[i] = 0; }
            return;
        }
        const uint64_t l = lc[x], r = rc[x];
        contiguous[x] = contiguous[l] && contiguous[r] && (sub_count[l] == 0 || sub_count[r] == 0 || sub_first[l] + sub_count[l] == sub_first[r]);
        sub_first[x] = sub_count[l] ? sub_first[l] : sub_first[r];
        sub_count[x] = sub_count[l] + sub_count[r];
        const float* cl = &cost[l * 8];
        const float* cr = &cost[r * 8];
        double dist[9]; int dist_k[9];
        for (int j = 2; j <= 8; ++j)
        {
            dist[j] = Inf; dist_k[j] = 1;
            for (int k = 1; k < j; ++k)
            {
                if (k > 7 || j - k > 7) continue;
                const double c = static_cast<double>(cl[k]) + static_cast<double>(cr[j - k]);
                if (c < dist[j]) { dist[j] = c; dist_k[j] = k; }
            }
        }
        const double c_leaf = (contiguous[x] && sub_count[x] <= leaf_cap) ? area * sub_count[x] * costs.item : Inf;
        const double c_node = dist[8] + area * costs.node;
        split8[x] = static_cast<uint8_t>(dist_k[8]);
        if (c_leaf <= c_node) { cx[1] = static_cast<float>(c_leaf); hx[1] = 0; }
        else { cx[1] = static_cast<float>(c_node); hx[1] = 1; }
        for (int i = 2; i < 8; ++i)
        {
            if (dist[i] < cx[i - 1]) { cx[i] = static_cast<float>(dist[i]); hx[i] = static_cast<uint8_t>(dist_k[i]); }
            else { cx[i] = cx[i - 1]; hx[i] = 0; }
        }
    };
    phase.mark("tables alloc");
    for (size_t d = size_t(max_level) + 1; d-- > 0; )
    {
        const uint32_t* level = by_level.data() + level_begin[d];
        parallel_chunks(level_begin[d + 1] - level_begin[d], threads, 1 << 13, [&](int, size_t begin, size_t end)
        {
            for (size_t i = begin; i < end; ++i) cost_tables(level[i]);
        });
    }
    by_level.release();
    cost.release();
    phase.mark("cost tables");

    // ---- top-down: emit wide nodes breadth first, one level at a time --------------------------------
    // The nodes of a level choose their children, slots and quantised planes in parallel (pass A);
    // prefix sums over the level give every node the index of its first internal child and of its first
    // leaf item -- the positions a sequential breadth-first walk would hand out -- and pass B writes.
    struct Pending { uint32_t wide_index; uint32_t node; };
    struct Kid { uint32_t is_node, node, src, first, count; };
    struct Temp
    {
        WNode       w;
        Kid         kids[8];
        int8_t      kid_in_slot[8];
        int8_t      slot_of[8];
        uint32_t    n_children, internal, items;
    };
    std::vector<Pending> level, next_level;
    {
        WNode blank; std::memset(&blank, 0, sizeof(blank));
        out.nodes.push_back(blank);
        Pending p; p.wide_index = 0; p.node = 0;
        level.push_back(p);
    }
    RawVector<Temp> temp;
    std::vector<uint64_t> node_offset, item_offset;
    ErrorSink sink;

    while (!level.empty())
    {
        ++out.depth;
        const size_t m = level.size();
        temp.resize(m);

        // Pass A.
        parallel_chunks(m, threads, 256, [&](int, size_t begin, size_t end)
        {
            std::vector<Work> work;
            for (size_t qi = begin; qi < end && !sink.failed(); ++qi)
            {
                const Pending cur = level[qi];
                Temp& t = temp[qi];
                Element kids[8];
                int n_kids = 0;
                bool overflow = false;

                auto emit_leaf = [&](const uint32_t x)
                {
                    Element e; e.is_node = false; e.node = 0; e.src = src[x]; e.first = sub_first[x]; e.count = sub_count[x]; e.box = box[x];
                    if (n_kids < 8) kids[n_kids++] = e; else overflow = true;
                };
                auto emit_node = [&](const uint32_t x)
                {
                    Element e; e.is_node = true; e.node = x; e.src = src[x]; e.first = e.count = 0; e.box = box[x];
                    if (n_kids < 8) kids[n_kids++] = e; else overflow = true;
                };

                if (lc[cur.node] == Leaf) { if (count[cur.node] > 0) emit_leaf(cur.node); }    // the whole tree is one small leaf
                else
                {
                    work.clear();
                    const int k8 = split8[cur.node];
                    Work wr; wr.node = rc[cur.node]; wr.budget = 8 - k8; work.push_back(wr);
                    Work wl; wl.node = lc[cur.node]; wl.budget = k8; work.push_back(wl);
                    while (!work.empty())
                    {
                        Work wk = work.back(); work.pop_back();
                        const uint32_t x = wk.node;
                        int i = wk.budget > 7 ? 7 : wk.budget;
                        if (lc[x] == Leaf) { if (count[x] > 0) emit_leaf(x); continue; }
                        while (i > 1 && choice[size_t(x) * 8 + i] == 0) --i;
                        if (i == 1)
                        {
                            if (choice[size_t(x) * 8 + 1] == 0) emit_leaf(x); else emit_node(x);
                            continue;
                        }
                        const int k = choice[size_t(x) * 8 + i];
                        Work b; b.node = rc[x]; b.budget = i - k; work.push_back(b);
                        Work a; a.node = lc[x]; a.budget = k; work.push_back(a);
                    }
                }
                if (overflow) { sink.fail(ASGPU_E_INVALID, "internal error: wide collapse produced more than 8 children"); return; }
                const int n_children = n_kids;

                // Octant slot assignment: slot s should hold the child that comes first for rays whose
                // direction signs are those of octant s (bit a set = negative along axis a).
                FBox nb; nb.reset();
                for (int i = 0; i < n_children; ++i) nb.grow(kids[i].box);
                int slot_of[8];
                bool slot_used[8] = { false, false, false, false, false, false, false, false };
                bool kid_done[8] = { false, false, false, false, false, false, false, false };
                float slot_cost[8][8];
                for (int i = 0; i < n_children; ++i)
                    for (int sl = 0; sl < 8; ++sl)
                    {
                        float c = 0.0f;
                        for (int a = 0; a < 3; ++a)
                        {
                            const float centre = kids[i].box.valid() ? 0.5f * (kids[i].box.lo[a] + kids[i].box.hi[a]) - 0.5f * (nb.lo[a] + nb.hi[a]) : 0.0f;
                            c += ((sl >> a) & 1) ? -centre : centre;
                        }
                        slot_cost[i][sl] = c;
                    }
                for (int round = 0; round < n_children; ++round)
                {
                    int bi = -1, bs = -1; float bc = std::numeric_limits<float>::max();
                    for (int i = 0; i < n_children; ++i)
                    {
                        if (kid_done[i]) continue;
                        for (int sl = 0; sl < 8; ++sl)
                            if (!slot_used[sl] && slot_cost[i][sl] < bc) { bc = slot_cost[i][sl]; bi = i; bs = sl; }
                    }
                    if (bi < 0)         // all remaining costs are +max/NaN: take any free pair
                    {
                        for (int i = 0; i < n_children && bi < 0; ++i) if (!kid_done[i]) bi = i;
                        for (int sl = 0; sl < 8 && bs < 0; ++sl) if (!slot_used[sl]) bs = sl;
                    }
                    slot_of[bi] = bs; slot_used[bs] = true; kid_done[bi] = true;
                }

                WNode& w = t.w;
                std::memset(&w, 0, sizeof(w));
                for (int a = 0; a < 3; ++a) for (int sl = 0; sl < 8; ++sl) { w.qlo[a][sl] = 255; w.qhi[a][sl] = 0; }
                std::string local_error;
                const int rc_q = quantise_node(w, kids, n_children, slot_of, local_error);
                if (rc_q != ASGPU_OK) { sink.fail(rc_q, local_error.c_str()); return; }

                t.n_children = static_cast<uint32_t>(n_children);
                t.internal = t.items = 0;
                for (int sl = 0; sl < 8; ++sl) t.kid_in_slot[sl] = -1;
                for (int i = 0; i < n_children; ++i)
                {
                    t.kid_in_slot[slot_of[i]] = static_cast<int8_t>(i);
                    t.slot_of[i] = static_cast<int8_t>(slot_of[i]);
                    Kid& k = t.kids[i];
                    k.is_node = kids[i].is_node ? 1u : 0u; k.node = kids[i].node; k.src = kids[i].src; k.first = kids[i].first; k.count = kids[i].count;
                    if (kids[i].is_node) ++t.internal; else t.items += kids[i].count;
                }
            }
        });
        if (sink.failed()) return sink.finish(error);

        // Positions: what a sequential breadth-first walk would hand out.
        node_offset.resize(m + 1); item_offset.resize(m + 1);
        node_offset[0] = out.nodes.size(); item_offset[0] = out.leaf_order.size();
        for (size_t qi = 0; qi < m; ++qi)
        {
            node_offset[qi + 1] = node_offset[qi] + temp[qi].internal;
            item_offset[qi + 1] = item_offset[qi] + temp[qi].items;
        }
        if (node_offset[m] >= 0xFFFFFFFFull) { error = "wide tree too large"; return ASGPU_E_UNSUPPORTED; }
        out.nodes.resize(node_offset[m]);
        out.leaf_order.resize(item_offset[m]);
        next_level.resize(node_offset[m] - node_offset[0]);
        const bool sliced = slice_count != 0 && static_cast<bool>(bv.timed_box);
        if (sliced) out.slices.resize(out.nodes.size() * slice_count);

        // Pass B.
        parallel_chunks(m, threads, 256, [&](int, size_t begin, size_t end)
        {
            for (size_t qi = begin; qi < end && !sink.failed(); ++qi)
            {
                const Pending cur = level[qi];
                Temp& t = temp[qi];
                WNode& w = t.w;
                // Internal children are allocated contiguously in slot order; leaves append their items
                // to the wide item order in slot order.
                w.child_base = static_cast<uint32_t>(node_offset[qi]);
                w.tri_base = static_cast<uint32_t>(item_offset[qi]);
                uint32_t leaf_offset = 0, child_cursor = 0;
                for (int sl = 0; sl < 8; ++sl)
                {
                    const int i = t.kid_in_slot[sl];
                    if (i < 0) continue;
                    const Kid& k = t.kids[i];
                    if (k.is_node)
                    {
                        w.imask |= uint8_t(1u << sl);
                        w.meta[sl] = uint8_t(0x20 | (24 + sl));
                        Pending p; p.wide_index = w.child_base + child_cursor; p.node = k.node;
                        next_level[node_offset[qi] - node_offset[0] + child_cursor] = p;
                        ++child_cursor;
                    }
                    else
                    {
                        if (k.count == 0) { w.meta[sl] = 0; continue; }
                        if (k.count > 3 || leaf_offset + k.count > 24) { sink.fail(ASGPU_E_INVALID, "internal error: wide leaf does not fit its node"); return; }
                        const uint32_t unary = (1u << k.count) - 1;     // 1 -> 001, 2 -> 011, 3 -> 111
                        w.meta[sl] = uint8_t((unary << 5) | leaf_offset);
                        for (uint32_t j = 0; j < k.count; ++j) out.leaf_order[item_offset[qi] + leaf_offset + j] = k.first + j;
                        leaf_offset += k.count;
                    }
                }
                out.nodes[cur.wide_index] = w;

                if (sliced)
                {
                    // Time slices: the same children in the same frame, bounded over 1 / T of the time axis
                    // each (a small overlap absorbs the rounding of the ray's slice index).
                    int slot_of[8];
                    for (uint32_t i = 0; i < t.n_children; ++i) slot_of[i] = t.slot_of[i];
                    for (uint32_t j = 0; j < slice_count; ++j)
                    {
                        const float t0 = std::max(0.0f, (float(j) - 1.0e-3f) / float(slice_count));
                        const float t1 = std::min(1.0f, (float(j) + 1.0f + 1.0e-3f) / float(slice_count));
                        FBox boxes[8];
                        for (uint32_t i = 0; i < t.n_children; ++i) boxes[i] = bv.timed_box(t.kids[i].src, t0, t1);
                        std::string local_error;
                        const int rs = quantise_slice(w, out.slices[size_t(cur.wide_index) * slice_count + j], boxes, static_cast<int>(t.n_children), slot_of, local_error);
                        if (rs != ASGPU_OK) { sink.fail(rs, local_error.c_str()); return; }
                    }
                }
            }
        });
        if (sink.failed()) return sink.finish(error);
        level.swap(next_level);
    }
    phase.mark("emit");
    return ASGPU_OK;
}

// Conservative all-time boxes of every binary node of a triangle tree.
void triangle_tree_boxes(const asgpu_triangle_tree_view& v, const ExactTree& t, BinaryView& bv)
{
    const AsNode* nodes = static_cast<const AsNode*>(v.nodes);
    const uint64_t n = v.node_count;
    bv.node_count = n;
    bv.child.assign(n, InteriorMark, host_threads());
    bv.first.assign(n, 0u, host_threads());
    bv.count.assign(n, 0u, host_threads());
    bv.box.resize(n);
    const bool motion = t.moving > 0;

    auto child_box = [&](const AsNode& node, const int side) -> FBox
    {
        FBox b;
        const uint32_t count = side == 0 ? node.left_bbox_count : node.right_bbox_count;
        const uint32_t index = side == 0 ? node.left_bbox_index : node.right_bbox_index;
        if (motion && count > 1)
        {
            // The traversal interpolates between consecutive entries (bvh_intersector.h:680-722);
            // their union bounds every interpolant.  a*w1 + b*w2 is rounded twice: pad slightly.
            b.reset();
            for (uint32_t k = 0; k < count; ++k)
            {
                const MBox& mb = t.mboxes[index + k];
                for (int a = 0; a < 3; ++a) { b.lo[a] = std::min(b.lo[a], mb.v[a * 2]); b.hi[a] = std::max(b.hi[a], mb.v[a * 2 + 1]); }
            }
            for (int a = 0; a < 3; ++a)
            {
                const float m = std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a])) * 2.0e-7f + std::numeric_limits<float>::min();
                b.lo[a] -= m; b.hi[a] += m;
            }
        }
        else
        {
            for (int a = 0; a < 3; ++a)
            {
                b.lo[a] = static_cast<float>(node.bbox[a * 4 + side]);
                b.hi[a] = static_cast<float>(node.bbox[a * 4 + 2 + side]);
            }
        }
        return b;
    };

    // Where the motion boxes of node x live (as a child of its parent): index / count into t.mboxes.
    std::vector<uint32_t> mindex, mcount;
    if (motion) { mindex.assign(n, 0); mcount.assign(n, 0); }

    // Children always have larger indices than their parent (bvh_builder.h:197-205), so one
    // forward pass assigns every box before it is needed.
    if (nodes[0].interior())
    {
        bv.box[0] = child_box(nodes[0], 0);
        bv.box[0].grow(child_box(nodes[0], 1));
    }
    else
    {
        bv.box[0].reset();
        for (uint32_t j = 0; j < nodes[0].item_count; ++j) bv.box[0].grow(triangle_box(t, nodes[0].index + j));
    }
    // (Every node has one parent -- check_hierarchy -- so the writes of different nodes never meet.)
    parallel_chunks(n, host_threads(), 1 << 15, [&](int, size_t begin, size_t end)
    {
        for (size_t i = begin; i < end; ++i)
        {
            const AsNode& node = nodes[i];
            if (node.interior())
            {
                bv.child[i] = node.index;
                bv.box[node.index] = child_box(node, 0);
                bv.box[node.index + 1] = child_box(node, 1);
                if (motion)
                {
                    mindex[node.index] = node.left_bbox_index; mcount[node.index] = node.left_bbox_count;
                    mindex[node.index + 1] = node.right_bbox_index; mcount[node.index + 1] = node.right_bbox_count;
                }
            }
            else
            {
                bv.first[i] = node.index;
                bv.count[i] = node.item_count;
            }
        }
    });

    if (motion)
    {
        // Box of node x over the ray times [t0, t1]: the reference interpolates consecutive motion
        // boxes linearly in the ray time (bvh_intersector.h:680-722), so the union over an interval
        // is the union of the interpolants at its ends and at the knots inside it.  Padded for the
        // float rounding of the interpolation and of the time products.
        // (bv.box and t.mboxes outlive the closure: both belong to the caller of collapse().)
        const FBox* all = bv.box.data();
        const MBox* mboxes = t.mboxes.data();
        bv.timed_box = [all, mindex, mcount, mboxes](const uint32_t x, const float t0, const float t1) -> FBox
        {
            if (x == 0 || mcount[x] <= 1) return all[x];
            const uint32_t segments = mcount[x] - 1;
            const MBox* mb = mboxes + mindex[x];
            FBox b; b.reset();
            double span[3] = { 0.0, 0.0, 0.0 };
            auto add_at = [&](const double time)
            {
                const double ts = std::min(double(segments), std::max(0.0, time * segments));
                uint32_t k = static_cast<uint32_t>(ts);
                if (k >= segments) k = segments - 1;
                const double f = ts - k;
                for (int a = 0; a < 3; ++a)
                {
                    const double lo = double(mb[k].v[a * 2]) * (1.0 - f) + double(mb[k + 1].v[a * 2]) * f;
                    const double hi = double(mb[k].v[a * 2 + 1]) * (1.0 - f) + double(mb[k + 1].v[a * 2 + 1]) * f;
                    b.lo[a] = std::min(b.lo[a], float_below(lo));
                    b.hi[a] = std::max(b.hi[a], float_above(hi));
                    span[a] = std::max(span[a], std::max(std::fabs(double(mb[k + 1].v[a * 2]) - double(mb[k].v[a * 2])),
                                                         std::fabs(double(mb[k + 1].v[a * 2 + 1]) - double(mb[k].v[a * 2 + 1]))));
                }
            };
            add_at(t0);
            add_at(t1);
            for (uint32_t k = 1; k < segments; ++k)
            {
                const double knot = double(k) / segments;
                if (knot > t0 && knot < t1) add_at(knot);
            }
            for (int a = 0; a < 3; ++a)
            {
                const float m = static_cast<float>(std::max(std::fabs(double(b.lo[a])), std::fabs(double(b.hi[a]))) * 4.0e-7 + span[a] * 1.0e-5)
                              + std::numeric_limits<float>::min();
                b.lo[a] -= m; b.hi[a] += m;
            }
            return b;
        };
    }
}

int assembly_tree_boxes(const asgpu_assembly_tree_view& top, BinaryView& bv, std::string& error)
{
    const AsNode* nodes = static_cast<const AsNode*>(top.nodes);
    const uint64_t n = top.node_count;
    bv.node_count = n;
    bv.child.assign(n, InteriorMark, 1);
    bv.first.assign(n, 0u, 1);
    bv.count.assign(n, 0u, 1);
    bv.box.resize(n);
    auto child_box = [&](const AsNode& node, const int side) -> FBox
    {
        FBox b;
        for (int a = 0; a < 3; ++a)
        {
            b.lo[a] = float_below(node.bbox[a * 4 + side]);
            b.hi[a] = float_above(node.bbox[a * 4 + 2 + side]);
        }
        return b;
    };
    if (nodes[0].interior())
    {
        bv.box[0] = child_box(nodes[0], 0);
        bv.box[0].grow(child_box(nodes[0], 1));
    }
    else
    {
        // A single-leaf assembly tree carries no box at all: use one that is unbounded for every
        // scene the quantiser accepts (extents beyond ~6e35 are rejected as unsupported anyway).
        for (int a = 0; a < 3; ++a) { bv.box[0].lo[a] = -1.0e30f; bv.box[0].hi[a] = 1.0e30f; }
    }
    for (uint64_t i = 0; i < n; ++i)
    {
        const AsNode& node = nodes[i];
        if (node.interior())
        {
            if (uint64_t(node.index) + 1 >= n) { error = "assembly tree child index out of range"; return ASGPU_E_INVALID; }
            bv.child[i] = node.index;
            bv.box[node.index] = child_box(node, 0);
            bv.box[node.index + 1] = child_box(node, 1);
        }
        else
        {
            if (uint64_t(node.index) + node.item_count > top.item_count) { error = "assembly tree item range out of range"; return ASGPU_E_INVALID; }
            bv.first[i] = node.index;
            bv.count[i] = node.item_count;
        }
    }
    return ASGPU_OK;
}

}   // anonymous namespace

int flatten_scene(
    const asgpu_triangle_tree_view* trees,
    uint32_t                        tree_count,
    const asgpu_assembly_tree_view& top,
    const asgpu_source_geometry*    sources,
    uint32_t                        flags,
    HostBlob&                       blob,
    std::string&                    error)
{
    if ((flags & (ASGPU_SCENE_EXACT | ASGPU_SCENE_WIDE)) == 0) { error = "scene flags select no layout"; return ASGPU_E_INVALID; }
    if (tree_count && !trees) { error = "null triangle tree array"; return ASGPU_E_INVALID; }
    if (!top.nodes || top.node_count == 0) { error = "assembly tree without nodes"; return ASGPU_E_INVALID; }
    if (top.item_count && !top.items) { error = "null assembly item array"; return ASGPU_E_INVALID; }
    if (top.node_count >= 0xFFFFFFFFull || top.item_count >= 0xFFFFFFFFull) { error = "assembly tree too large"; return ASGPU_E_UNSUPPORTED; }
    {
        // The exact layout uploads these nodes as they are: check them whatever the layout flags say.
        const AsNode* top_nodes = static_cast<const AsNode*>(top.nodes);
        const int shape = check_hierarchy(top_nodes, top.node_count, "assembly tree", error);
        if (shape != ASGPU_OK) return shape;
        for (uint64_t i = 0; i < top.node_count; ++i)
            if (!top_nodes[i].interior() && uint64_t(top_nodes[i].index) + top_nodes[i].item_count > top.item_count)
            { error = "assembly tree item range out of range"; return ASGPU_E_INVALID; }
    }

    const bool want_exact = (flags & ASGPU_SCENE_EXACT) != 0;
    const bool want_wide = (flags & ASGPU_SCENE_WIDE) != 0;

    BlobHeader header; std::memset(&header, 0, sizeof(header));
    header.magic = BlobMagic;
    header.version = BlobVersion;
    header.flags = flags & (ASGPU_SCENE_EXACT | ASGPU_SCENE_WIDE);
    header.tree_count = tree_count;
    header.item_count = static_cast<uint32_t>(top.item_count);

    BlobWriter writer(blob);
    writer.append(&header, sizeof(header));     // rewritten at the end

    const bool timing = std::getenv("ASGPU_BUILD_TIMING") != nullptr;
    auto stamp = [timing, last = std::chrono::steady_clock::now()](const char* what) mutable
    {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "asgpu flatten: %-22s %.3f s\n", what, std::chrono::duration<double>(now - last).count());
        last = now;
    };

    uint32_t max_bottom_depth = 0;
    std::vector<TreeDesc> descs(tree_count);
    for (uint32_t ti = 0; ti < tree_count; ++ti)
    {
        ExactTree et;
        int rc = decode_tree(trees[ti], et, error);
        if (rc != ASGPU_OK) return rc;
        stamp("decode tree");

        TreeDesc& d = descs[ti];
        std::memset(&d, 0, sizeof(d));
        d.slot_count = static_cast<uint32_t>(et.tris.size());
        d.moving = static_cast<uint32_t>(et.moving);
        d.keys = writer.append(et.keys);
        if (!et.poses.empty()) d.poses = writer.append(et.poses);
        header.triangle_count += et.tris.size();
        header.moving_triangle_count += et.moving;
        header.triangle_bytes += et.poses.size() * 4;

        if (want_exact)
        {
            d.bnodes = writer.append(et.bnodes);
            d.bnode_count = static_cast<uint32_t>(et.bnodes.size());
            if (et.moving > 0)
            {
                d.mnodes = writer.append(et.mnodes);
                d.mboxes = writer.append(et.mboxes);
                d.mbox_count = static_cast<uint32_t>(et.mboxes.size());
            }
            d.tris = writer.append(et.tris);
            header.binary_node_count += et.bnodes.size();
            header.binary_node_bytes += et.bnodes.size() * sizeof(BNodeF) + et.mnodes.size() * sizeof(MNode) + et.mboxes.size() * sizeof(MBox);
            header.triangle_bytes += et.tris.size() * sizeof(TriRecord);
        }

        stamp("exact layout");
        if (sources && sources[ti].object_count != 0)
        {
            // Source geometry for the device-side refine_and_offset: packed copies of every object
            // instance's vertices and vertex indices.
            const asgpu_source_geometry& sg = sources[ti];
            if (!sg.objects) { error = "null source object array"; return ASGPU_E_INVALID; }
            std::vector<SrcObject> objs(sg.object_count);
            for (uint32_t o = 0; o < sg.object_count; ++o)
            {
                const asgpu_source_object& so = sg.objects[o];
                SrcObject& dst = objs[o];
                std::memset(&dst, 0, sizeof(dst));
                if ((so.vertex_count && !so.vertices) || (so.triangle_count && !so.triangles) || (so.triangle_count && so.triangle_stride < 12) ||
                    (so.motion_segment_count && so.vertex_count && !so.vertex_poses))
                { error = "malformed source object"; return ASGPU_E_INVALID; }
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) dst.parent_to_local[r * 3 + c] = so.parent_to_local[r * 4 + c];
                dst.vertex_count = so.vertex_count;
                dst.triangle_count = so.triangle_count;
                dst.vertices = writer.append(so.vertices, size_t(so.vertex_count) * 12);
                if (so.motion_segment_count && so.vertex_count)
                {
                    dst.poses = writer.append(so.vertex_poses, size_t(so.vertex_count) * so.motion_segment_count * 12);
                    dst.motion_segment_count = so.motion_segment_count;
                }
                // Vertex indices packed straight into the blob (and checked), several threads.
                dst.triangles = writer.reserve(size_t(so.triangle_count) * 12);
                {
                    uint32_t* packed = reinterpret_cast<uint32_t*>(writer.at(dst.triangles));
                    const uint8_t* src = static_cast<const uint8_t*>(so.triangles);
                    std::atomic<bool> bad(false);
                    parallel_chunks(so.triangle_count, host_threads(), 1 << 16, [&](int, size_t begin, size_t end)
                    {
                        for (size_t t = begin; t < end; ++t)
                        {
                            std::memcpy(&packed[t * 3], src + t * so.triangle_stride, 12);
                            for (int k = 0; k < 3; ++k)
                                if (packed[t * 3 + k] >= so.vertex_count) bad.store(true, std::memory_order_relaxed);
                        }
                    });
                    if (bad.load()) { error = "source triangle references a missing vertex"; return ASGPU_E_INVALID; }
                }
            }
            {
                std::atomic<bool> bad(false);
                parallel_chunks(et.keys.size(), host_threads(), 1 << 16, [&](int, size_t begin, size_t end)
                {
                    for (size_t i = begin; i < end; ++i)
                    {
                        const HitKey& key = et.keys[i];
                        if (key.object_instance_index >= sg.object_count || key.triangle_index >= objs[key.object_instance_index].triangle_count)
                            bad.store(true, std::memory_order_relaxed);
                    }
                });
                if (bad.load()) { error = "triangle key outside the source geometry"; return ASGPU_E_INVALID; }
            }
            d.src_objects = writer.append(objs);
            d.src_object_count = sg.object_count;

            if (sg.filters)
            {
                // Intersection filters: masks and UVs copied as they are, plus the primitive-attribute
                // index of every leaf slot (the material mask is chosen by it).
                std::vector<FilterRecord> frs(sg.object_count);
                bool any = false;
                auto put_mask = [&](const asgpu_alpha_mask& m, MaskRecord& out)
                {
                    std::memset(&out, 0, sizeof(out));
                    if (!m.bits || m.width == 0 || m.height == 0) return;
                    out.bits = writer.append(m.bits, size_t((m.width + 7) / 8) * m.height);
                    out.width = m.width; out.height = m.height;
                };
                for (uint32_t o = 0; o < sg.object_count; ++o)
                {
                    const asgpu_intersection_filter& f = sg.filters[o];
                    FilterRecord& fr = frs[o];
                    std::memset(&fr, 0, sizeof(fr));
                    if (!f.uv) continue;
                    if (f.material_mask_count && !f.material_masks) { error = "null material mask array"; return ASGPU_E_INVALID; }
                    any = true;
                    fr.uv = writer.append(f.uv, size_t(sg.objects[o].triangle_count) * 6 * sizeof(float));
                    put_mask(f.object_mask, fr.object_mask);
                    std::vector<MaskRecord> mats(f.material_mask_count);
                    for (uint32_t k = 0; k < f.material_mask_count; ++k) put_mask(f.material_masks[k], mats[k]);
                    if (!mats.empty()) fr.material_masks = writer.append(mats);
                    fr.material_mask_count = f.material_mask_count;
                }
                if (any)
                {
                    const AsTriangleKey* keys = static_cast<const AsTriangleKey*>(trees[ti].triangle_keys);
                    std::vector<uint16_t> pa(et.keys.size());
                    for (size_t k = 0; k < pa.size(); ++k) pa[k] = keys[k].triangle_pa;
                    d.key_pa = writer.append(pa);
                    d.filters = writer.append(frs);
                    d.filter_count = sg.object_count;
                    header.flags |= BlobHasFilters;
                }
            }
        }

        stamp("source geometry");
        if (want_wide)
        {
            BinaryView bv;
            triangle_tree_boxes(trees[ti], et, bv);
            stamp("binary boxes");
            WideOut wo;
            const uint32_t slice_count = et.moving > 0 ? time_slices() : 0;
            rc = collapse(bv, 3, collapse_costs(), wo, error, slice_count);
            if (rc != ASGPU_OK) return rc;
            stamp("wide collapse");
            if (wo.leaf_order.size() != et.tris.size()) { error = "internal error: wide collapse lost triangles"; return ASGPU_E_INVALID; }
            max_bottom_depth = std::max(max_bottom_depth, wo.depth);
            d.wnodes = writer.append(wo.nodes);
            d.wnode_count = static_cast<uint32_t>(wo.nodes.size());
            // The triangle records in wide-node order, gathered straight into the blob.
            const size_t wtri_count = wo.leaf_order.size();
            d.wtris = writer.reserve(wtri_count * sizeof(TriRecord));
            {
                TriRecord* wtris = reinterpret_cast<TriRecord*>(writer.at(d.wtris));
                parallel_chunks(wtri_count, host_threads(), 1 << 15, [&](int, size_t begin, size_t end)
                {
                    for (size_t i = begin; i < end; ++i) wtris[i] = et.tris[wo.leaf_order[i]];
                });
            }
            if (slice_count != 0 && !wo.slices.empty())
            {
                d.wslices = writer.append(wo.slices);
                d.wslice_count = slice_count;
            }
            header.wide_node_count += wo.nodes.size();
            header.wide_node_bytes += wo.nodes.size() * sizeof(WNode) + wo.slices.size() * sizeof(WSlice);
            if (!want_exact) header.triangle_bytes += wtri_count * sizeof(TriRecord);
            stamp("wide triangles");
        }
    }
    header.trees = writer.append(descs);

    // Items.
    std::vector<ItemRecord> items(top.item_count);
    for (uint64_t i = 0; i < top.item_count; ++i)
    {
        const asgpu_assembly_item& src = top.items[i];
        // Transform::point_to_local divides by w only when w != 1 (transform.h:311-344); with an
        // affine last row w is exactly 1 for finite points, so only affine instances are accepted.
        if (src.parent_to_local[12] != 0.0 || src.parent_to_local[13] != 0.0 || src.parent_to_local[14] != 0.0 || src.parent_to_local[15] != 1.0)
        { error = "projective assembly-instance transforms are not supported"; return ASGPU_E_UNSUPPORTED; }
        if (src.triangle_tree != ASGPU_MISS && src.triangle_tree >= tree_count) { error = "assembly item references a missing triangle tree"; return ASGPU_E_INVALID; }
        ItemRecord& dst = items[i];
        std::memset(&dst, 0, sizeof(dst));
        std::memcpy(dst.m, src.parent_to_local, 12 * sizeof(double));
        dst.tree = src.triangle_tree;
        dst.vis_flags = src.vis_flags;
        dst.assembly_instance = src.assembly_instance;
        if (top.item_motion && top.item_motion[i].key_count >= 2)
        {
            // Animated instance: key times, key matrices (3 x 4 rows) and interpolator segments.
            const asgpu_item_motion& mo = top.item_motion[i];
            if (!mo.key_times || !mo.key_parent_to_local || !mo.segments) { error = "animated assembly instance misses an array"; return ASGPU_E_INVALID; }
            const uint32_t k = mo.key_count;
            const size_t time_doubles = (size_t(k) * 4 + 15) / 16 * 2;
            std::vector<double> block(time_doubles + size_t(k) * 12 + size_t(k - 1) * 20, 0.0);
            std::memcpy(block.data(), mo.key_times, size_t(k) * 4);
            for (uint32_t key = 0; key < k; ++key)
            {
                const double* m = mo.key_parent_to_local + size_t(key) * 16;
                if (key > 0 && !(mo.key_times[key] > mo.key_times[key - 1])) { error = "transform keys are not in ascending time order"; return ASGPU_E_INVALID; }
                if (m[12] != 0.0 || m[13] != 0.0 || m[14] != 0.0 || m[15] != 1.0) { error = "projective assembly-instance transforms are not supported"; return ASGPU_E_UNSUPPORTED; }
                std::memcpy(block.data() + time_doubles + size_t(key) * 12, m, 12 * sizeof(double));
            }
            std::memcpy(block.data() + time_doubles + size_t(k) * 12, mo.segments, size_t(k - 1) * sizeof(asgpu_transform_segment));
            dst.key_count = k;
            dst.motion = writer.append(block);
            header.flags |= BlobHasAnimatedInstances;
        }
    }
    header.items = writer.append(items);

    if (want_exact)
    {
        static_assert(sizeof(BNodeD) == sizeof(AsNode), "top-level nodes are uploaded as they are");
        header.top_nodes = writer.append(top.nodes, top.node_count * sizeof(AsNode));
        header.top_node_count = static_cast<uint32_t>(top.node_count);
        header.binary_node_count += top.node_count;
        header.binary_node_bytes += top.node_count * sizeof(AsNode);
    }

    if (want_wide)
    {
        BinaryView bv;
        int rc = assembly_tree_boxes(top, bv, error);
        if (rc != ASGPU_OK) return rc;
        WideOut wo;
        rc = collapse(bv, 1, collapse_costs(), wo, error);
        if (rc != ASGPU_OK) return rc;
        if (wo.leaf_order.size() != top.item_count) { error = "internal error: wide collapse lost instances"; return ASGPU_E_INVALID; }
        header.top_wnodes = writer.append(wo.nodes);
        header.top_wnode_count = static_cast<uint32_t>(wo.nodes.size());
        header.top_witems = writer.append(wo.leaf_order);
        // Deepest traversal stack of the wide kernels: one saved node group per level below the
        // root in either tree, plus node group, instance group and sentinel when entering an instance.
        header.wide_stack_need = (wo.depth > 0 ? wo.depth - 1 : 0) + 3 + (max_bottom_depth > 0 ? max_bottom_depth - 1 : 0);
        if (header.wide_stack_need > WideStackMax)
        { error = "wide trees too deep for the traversal stack"; return ASGPU_E_UNSUPPORTED; }
        header.wide_node_count += wo.nodes.size();
        header.wide_node_bytes += wo.nodes.size() * sizeof(WNode);
    }

    stamp("items + top level");
    writer.finish();
    header.total_bytes = blob.size();
    std::memcpy(blob.data(), &header, sizeof(header));
    return ASGPU_OK;
}

int validate_blob_tables(const BlobReader& read, const uint64_t size, std::string& error)
{
    if (size < sizeof(BlobHeader)) { error = "blob too small"; return ASGPU_E_INVALID; }
    BlobHeader h;
    if (!read(0, &h, sizeof(h))) { error = "blob header unreadable"; return ASGPU_E_INVALID; }
    if (h.magic != BlobMagic || h.version != BlobVersion) { error = "blob magic/version mismatch"; return ASGPU_E_INVALID; }
    if (h.total_bytes != size) { error = "blob size mismatch"; return ASGPU_E_INVALID; }
    auto inside = [&](uint64_t off, uint64_t bytes) { return off <= size && bytes <= size - off; };
    if (!inside(h.trees, uint64_t(h.tree_count) * sizeof(TreeDesc)) || !inside(h.items, uint64_t(h.item_count) * sizeof(ItemRecord)) ||
        !inside(h.top_nodes, uint64_t(h.top_node_count) * sizeof(BNodeD)) || !inside(h.top_wnodes, uint64_t(h.top_wnode_count) * sizeof(WNode)) ||
        !inside(h.top_witems, h.top_witems ? uint64_t(h.item_count) * 4 : 0))
    { error = "blob section out of range"; return ASGPU_E_INVALID; }
    for (uint32_t i = 0; i < h.tree_count; ++i)
    {
        TreeDesc d;
        if (!read(h.trees + uint64_t(i) * sizeof(TreeDesc), &d, sizeof(d))) { error = "blob tree table unreadable"; return ASGPU_E_INVALID; }
        if (!inside(d.bnodes, uint64_t(d.bnode_count) * sizeof(BNodeF)) || !inside(d.wnodes, uint64_t(d.wnode_count) * sizeof(WNode)) ||
            !inside(d.keys, uint64_t(d.slot_count) * sizeof(HitKey)) || !inside(d.tris, d.tris ? uint64_t(d.slot_count) * sizeof(TriRecord) : 0) ||
            !inside(d.wtris, d.wtris ? uint64_t(d.slot_count) * sizeof(TriRecord) : 0) ||
            !inside(d.mnodes, d.mnodes ? uint64_t(d.bnode_count) * sizeof(MNode) : 0) ||
            !inside(d.mboxes, uint64_t(d.mbox_count) * sizeof(MBox)) ||
            !inside(d.wslices, uint64_t(d.wnode_count) * d.wslice_count * sizeof(WSlice)) ||
            !inside(d.src_objects, uint64_t(d.src_object_count) * sizeof(SrcObject)) ||
            !inside(d.filters, uint64_t(d.filter_count) * sizeof(FilterRecord)) ||
            !inside(d.key_pa, d.key_pa ? uint64_t(d.slot_count) * 2 : 0))
        { error = "blob tree section out of range"; return ASGPU_E_INVALID; }
        for (uint32_t o = 0; o < d.src_object_count; ++o)
        {
            SrcObject so;
            if (!read(d.src_objects + uint64_t(o) * sizeof(SrcObject), &so, sizeof(so))) { error = "blob source table unreadable"; return ASGPU_E_INVALID; }
            if (!inside(so.vertices, uint64_t(so.vertex_count) * 12) || !inside(so.triangles, uint64_t(so.triangle_count) * 12) ||
                !inside(so.poses, so.poses ? uint64_t(so.vertex_count) * so.motion_segment_count * 12 : 0) || (so.motion_segment_count && so.vertex_count && !so.poses))
            { error = "blob source geometry out of range"; return ASGPU_E_INVALID; }
        }
        for (uint32_t o = 0; o < d.filter_count; ++o)
        {
            FilterRecord fr;
            if (!read(d.filters + uint64_t(o) * sizeof(FilterRecord), &fr, sizeof(fr))) { error = "blob filter table unreadable"; return ASGPU_E_INVALID; }
            auto mask_ok = [&](const MaskRecord& m) { return m.bits == 0 || inside(m.bits, uint64_t((m.width + 7) / 8) * m.height); };
            bool ok = mask_ok(fr.object_mask) && inside(fr.material_masks, uint64_t(fr.material_mask_count) * sizeof(MaskRecord));
            for (uint32_t k = 0; ok && k < fr.material_mask_count; ++k)
            {
                MaskRecord m;
                ok = read(fr.material_masks + uint64_t(k) * sizeof(MaskRecord), &m, sizeof(m)) && mask_ok(m);
            }
            if (!ok) { error = "blob intersection filter out of range"; return ASGPU_E_INVALID; }
        }
    }
    return ASGPU_OK;
}

int validate_blob(const uint8_t* blob, size_t size, std::string& error)
{
    if (!blob) { error = "blob too small"; return ASGPU_E_INVALID; }
    return validate_blob_tables(
        [blob, size](uint64_t offset, void* dst, size_t bytes) -> bool
        {
            if (offset > size || bytes > size - offset) return false;
            std::memcpy(dst, blob + offset, bytes);
            return true;
        }, size, error);
}

}   // namespace asgpu
