//
// tree_builder.cpp -- host builder for reference-format trees.
//
// Produces, for the same input, the same binary BVHs the reference builds, so that the GPU
// layouts derived from them (flatten.cpp) inherit the reference's leaf contents and visit order:
//
//   triangle trees   TriangleTree::build_bvh          renderer/kernel/intersection/triangletree.cpp:497-597
//                    collect_*_triangles              :105-383
//                    compute_motion_bboxes            :755-876
//                    store_triangles + TriangleEncoder :878-978, triangleencoder.cpp:48-103
//   assembly tree    AssemblyTree::rebuild_assembly_tree  assemblytree.cpp:111-245
//   split search     bvh::SAHPartitioner::partition   foundation/math/bvh/bvh_sahpartitioner.h:99-170
//   index upkeep     PartitionerBase::sort_indices    foundation/math/bvh/bvh_partitionerbase.h:135-198
//   node order       bvh::Builder::subdivide_recurse  foundation/math/bvh/bvh_builder.h:163-229
//
// Unlike the reference's single-threaded recursion, the work is organised as independent range
// tasks: the three initial centroid sorts run concurrently, the three per-axis sweeps of large
// nodes run concurrently, and subtrees below a grain size are built in parallel into private
// node arrays that are stitched into the reference's depth-first node order afterwards.  The
// arithmetic (float boxes and costs for triangle trees, double for the assembly tree, no FMA
// contraction -- this file is compiled with -ffp-contract=off) and every tie-break are the
// reference's, so the resulting arrays are identical.
//

#include "tree_builder.h"
#include "lbvh_core.h"
#include "motion_bounds.h"
#include "parallel.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <mutex>
#include <thread>

namespace asgpu
{

namespace
{

template <typename T>
struct Bounds
{
    T lo[3], hi[3];

    void reset()
    {
        for (int a = 0; a < 3; ++a)
        {
            lo[a] = std::numeric_limits<T>::max();
            hi[a] = -std::numeric_limits<T>::max();
        }
    }

    void grow(const T x, const T y, const T z)
    {
        const T p[3] = { x, y, z };
        for (int a = 0; a < 3; ++a)
        {
            if (lo[a] > p[a]) lo[a] = p[a];
            if (hi[a] < p[a]) hi[a] = p[a];
        }
    }

    void grow(const Bounds& b)
    {
        for (int a = 0; a < 3; ++a)
        {
            if (lo[a] > b.lo[a]) lo[a] = b.lo[a];
            if (hi[a] < b.hi[a]) hi[a] = b.hi[a];
        }
    }

    // Number of axes with non-zero extent (AABBBase::rank, aabb.h:422-433).
    int rank() const
    {
        return (lo[0] < hi[0] ? 1 : 0) + (lo[1] < hi[1] ? 1 : 0) + (lo[2] < hi[2] ? 1 : 0);
    }

    bool valid() const { return lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]; }

    // half_surface_area (aabb.h:723-730).
    T half_area() const
    {
        const T ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
        return ex * ey + ex * ez + ey * ez;
    }
};

typedef Bounds<float> BoundsF;
typedef Bounds<double> BoundsD;

struct P3 { float x, y, z; };

// Transform<double>::point_to_parent<float> (foundation/math/transform.h:346-375).
inline P3 to_parent(const double* m, const P3& p)
{
    P3 r;
    r.x = static_cast<float>(m[0] * double(p.x) + m[1] * double(p.y) + m[ 2] * double(p.z) + m[ 3]);
    r.y = static_cast<float>(m[4] * double(p.x) + m[5] * double(p.y) + m[ 6] * double(p.z) + m[ 7]);
    r.z = static_cast<float>(m[8] * double(p.x) + m[9] * double(p.y) + m[10] * double(p.z) + m[11]);
    const float w = static_cast<float>(m[12] * double(p.x) + m[13] * double(p.y) + m[14] * double(p.z) + m[15]);
    if (w != 1.0f) { r.x /= w; r.y /= w; r.z /= w; }
    return r;
}

// Transform::to_parent(AABB) (transform.h:528-546): corner order matters only through min/max,
// which is order-independent, but the point transform must be the float-returning one.
inline BoundsF box_to_parent(const double* m, const BoundsF& b)
{
    if (!b.valid()) return b;
    BoundsF r; r.reset();
    for (int k = 0; k < 8; ++k)
    {
        const P3 c = { (k & 4) ? b.hi[0] : b.lo[0], (k & 2) ? b.hi[1] : b.lo[1], (k & 1) ? b.hi[2] : b.lo[2] };
        const P3 q = to_parent(m, c);
        r.grow(q.x, q.y, q.z);
    }
    return r;
}

// square_area(v0, v1, v2) == 0 (foundation/math/area.h:57-66), evaluated in float.
inline bool degenerate(const P3& a, const P3& b, const P3& c)
{
    const float ux = b.x - a.x, uy = b.y - a.y, uz = b.z - a.z;
    const float vx = c.x - a.x, vy = c.y - a.y, vz = c.z - a.z;
    const float nx = uy * vz - vy * uz;
    const float ny = uz * vx - vz * ux;
    const float nz = ux * vy - vx * uy;
    float s = 0.0f;
    s += nx * nx; s += ny * ny; s += nz * nz;
    return 0.25f * s == 0.0f;
}

//
// Sweep-SAH range builder.
//

template <typename T>
class SweepBuilder
{
  public:
    typedef Bounds<T> Box;

    SweepBuilder(const std::vector<Box>& boxes, size_t max_leaf_size, T traversal_cost, T item_cost, int threads)
      : m_boxes(boxes)
      , m_max_leaf(max_leaf_size)
      , m_ct(traversal_cost)
      , m_ci(item_cost)
      , m_threads(threads < 1 ? 1 : threads)
    {
        if (const char* e = getenv("ASGPU_PARALLEL_SWEEP_MIN")) { const long v = atol(e); if (v >= 1024) ParallelSweepMin = static_cast<size_t>(v); }
    }

    // Final item ordering (PartitionerBase::get_item_ordering(0)).
    const std::vector<uint32_t>& ordering() const { return m_order[0]; }

    void build(AsNodeVector& nodes)
    {
        const size_t n = m_boxes.size();
        // ASGPU_BUILD_TIMING: the four phases of trees worth timing (stderr).
        const bool timing = n >= 100000 && getenv("ASGPU_BUILD_TIMING") != nullptr;
        auto t_phase = std::chrono::steady_clock::now();
        auto phase = [&](const char* what)
        {
            if (!timing) return;
            const auto now = std::chrono::steady_clock::now();
            fprintf(stderr, "asgpu build:   sweep SAH / %-18s %.3f s\n", what, std::chrono::duration<double>(now - t_phase).count());
            t_phase = now;
        };
        sort_centroids(n);
        phase("centroid sorts");
        m_scratch.resize(n);
        m_side.resize(n);
        for (int a = 0; a < 3; ++a) m_prefix[a].resize(n);

        m_grain = std::max<size_t>(2048, n / (static_cast<size_t>(m_threads) * 16));

        // Phase 1: expand the top of the tree serially (per-axis sweeps of each node run
        // concurrently) down to ranges of at most m_grain items.
        Box root; root.reset();
        for (size_t i = 0; i < n; ++i) root.grow(m_boxes[m_order[0][i]]);
        m_top.clear();
        m_jobs.clear();
        m_top.reserve(1024);
        m_fork_depth = 0;
        for (int t = m_threads; t > 1; t >>= 1) ++m_fork_depth;      // ~log2(threads) forking levels
        expand_top(0, n, widen(root));
        phase("top (forked)");

        // Phase 2: build the deferred subtrees in parallel, largest first.
        std::vector<size_t> by_size(m_jobs.size());
        for (size_t i = 0; i < by_size.size(); ++i) by_size[i] = i;
        std::sort(by_size.begin(), by_size.end(), [this](size_t a, size_t b)
        {
            return (m_jobs[a].end - m_jobs[a].begin) > (m_jobs[b].end - m_jobs[b].begin);
        });
        std::atomic<size_t> next(0);
        auto worker = [&]()
        {
            for (;;)
            {
                const size_t k = next.fetch_add(1);
                if (k >= by_size.size()) break;
                Job& job = m_jobs[by_size[k]];
                build_range(job);
            }
        };
        if (m_threads == 1 || m_jobs.size() < 2) worker();
        else
        {
            std::vector<std::thread> pool;
            for (int t = 0; t < m_threads; ++t) pool.emplace_back(worker);
            for (std::thread& th : pool) th.join();
        }

        phase("subtrees");

        // Phase 3: lay everything out in the reference's depth-first order: a node's two
        // children are adjacent and are followed by the whole left subtree, then the right one
        // (bvh_builder.h:197-228).
        size_t total = 1;
        for (const TopNode& t : m_top) if (t.kind == Interior) total += 2;
        for (const Job& j : m_jobs) total += j.nodes.size();
        // The array is not zero-filled first (1.3 GB for 10 M triangles, touched by one thread): the
        // walk over the top writes its own nodes and notes where every job's block goes; the blocks
        // -- every remaining node -- are then copied by all threads.
        nodes.clear();
        nodes.resize(total);
        m_placed.clear();
        size_t cursor = 1;
        place(0, 0, cursor, nodes);
        std::atomic<size_t> next_block(0);
        auto copier = [&]()
        {
            for (;;)
            {
                const size_t k = next_block.fetch_add(1);
                if (k >= m_placed.size()) break;
                const Placed& p = m_placed[k];
                const Job& job = m_jobs[p.job];
                nodes[p.slot] = job.root;
                if (nodes[p.slot].interior()) nodes[p.slot].index += static_cast<uint32_t>(p.base);
                for (size_t i = 0; i < job.nodes.size(); ++i)
                {
                    AsNode& dst = nodes[p.base + i];
                    dst = job.nodes[i];
                    if (dst.interior()) dst.index += static_cast<uint32_t>(p.base);
                }
            }
        };
        if (m_threads == 1 || m_placed.size() < 2) copier();
        else
        {
            std::vector<std::thread> pool;
            for (int t = 0; t < m_threads; ++t) pool.emplace_back(copier);
            for (std::thread& th : pool) th.join();
        }
        phase("placement");
    }

  private:
    enum Kind { Interior, Deferred };

    struct TopNode
    {
        Kind    kind;
        size_t  left, right;        // TopNode indices (Interior)
        size_t  job;                // Job index (Deferred)
        BoundsD left_box, right_box;
    };

    struct Job
    {
        size_t              begin, end;
        BoundsD             box;
        AsNode              root;       // the subtree's own root (lives in its parent's child pair)
        std::vector<AsNode> nodes;      // descendants, child indices relative to nodes[0]
    };

    const std::vector<Box>&     m_boxes;
    const size_t                m_max_leaf;
    const T                     m_ct, m_ci;
    const int                   m_threads;
    size_t                      m_grain;
    std::vector<uint32_t>       m_order[3];
    std::vector<uint32_t>       m_scratch;
    std::vector<uint8_t>        m_side;
    std::vector<T>              m_prefix[3];
    std::vector<TopNode>        m_top;
    std::vector<Job>            m_jobs;
    std::mutex                  m_top_mutex;
    int                         m_fork_depth = 0;
    struct Placed { size_t job, slot, base; };      // where a job's root and block go in the final array
    std::vector<Placed>         m_placed;

    static BoundsD widen(const Box& b)
    {
        BoundsD r;
        for (int a = 0; a < 3; ++a) { r.lo[a] = static_cast<double>(b.lo[a]); r.hi[a] = static_cast<double>(b.hi[a]); }
        return r;
    }

    static Box narrow(const BoundsD& b)
    {
        Box r;
        for (int a = 0; a < 3; ++a) { r.lo[a] = static_cast<T>(b.lo[a]); r.hi[a] = static_cast<T>(b.hi[a]); }
        return r;
    }

    // PartitionerBase constructor (bvh_partitionerbase.h:95-120): identity order, then an
    // unstable std::sort by (min + max) with a strict '<' (bvh_bboxsortpredicate.h:113-126).
    // Ties fall where libstdc++'s introsort puts them, as they do in the reference; the key
    // type must therefore also be the reference's (size_t), sorted here and narrowed after.
    void sort_centroids(const size_t n)
    {
        auto sort_axis = [this, n](const int a)
        {
            // The same std::sort over the same sequence with the same outcome of every comparison
            // as the reference's sort of item indices by looked-up (min + max) -- introsort is
            // driven by comparison results and n alone, so the permutation (ties included) is the
            // reference's -- but the keys travel with the indices instead of being fetched from
            // the box array at random on every comparison (3 - 4 x faster on 10 M items).
            struct Keyed { T key; size_t index; };
            std::vector<Keyed> idx(n);
            const std::vector<Box>& bb = m_boxes;
            for (size_t i = 0; i < n; ++i) { idx[i].key = bb[i].lo[a] + bb[i].hi[a]; idx[i].index = i; }
            std::sort(idx.begin(), idx.end(), [](const Keyed& l, const Keyed& r) { return l.key < r.key; });
            m_order[a].resize(n);
            for (size_t i = 0; i < n; ++i) m_order[a][i] = static_cast<uint32_t>(idx[i].index);
        };
        if (m_threads >= 3 && n > 50000)
        {
            std::thread t1(sort_axis, 1), t2(sort_axis, 2);
            sort_axis(0);
            t1.join(); t2.join();
        }
        else for (int a = 0; a < 3; ++a) sort_axis(a);
    }

    // Big ranges (the first levels of a multi-million-item tree) are swept by `inner` threads each:
    // unions of boxes are exact and order-independent (min / max), so chunked prefix / suffix boxes
    // -- and with them every area, every cost and, with ties resolved in the same scan order, the
    // pivot -- are the serial ones bit for bit.
    // (ASGPU_PARALLEL_SWEEP_MIN lowers the threshold so that tests cover the chunked paths on small meshes.)
    size_t ParallelSweepMin = 262144;

    Box range_bounds(const size_t begin, const size_t end, const int inner = 1) const
    {
        Box b; b.reset();
        if (inner > 1 && end - begin >= ParallelSweepMin)
        {
            std::vector<Box> part(static_cast<size_t>(chunk_count(end - begin, inner, 65536)));
            const uint32_t* order = m_order[0].data() + begin;
            parallel_chunks(end - begin, inner, 65536, [&](int c, size_t lo, size_t hi)
            {
                Box p; p.reset();
                for (size_t i = lo; i < hi; ++i) p.grow(m_boxes[order[i]]);
                part[static_cast<size_t>(c)] = p;
            });
            for (const Box& p : part) b.grow(p);
            return b;
        }
        for (size_t i = begin; i < end; ++i) b.grow(m_boxes[m_order[0][i]]);
        return b;
    }

    struct Candidate { T cost; size_t pivot; };

    // One axis of SAHPartitioner::partition (bvh_sahpartitioner.h:118-151): prefix areas left to
    // right, then a right-to-left sweep keeping the strictly cheapest split.
    Candidate sweep_axis(const int a, const size_t begin, const size_t end, const int inner = 1)
    {
        const size_t count = end - begin;
        const uint32_t* order = m_order[a].data() + begin;
        T* prefix = m_prefix[a].data() + begin;
        Box acc;

        if (inner > 1 && count >= ParallelSweepMin)
        {
            // Forward: prefix[i] = area of the union of items 0 .. i, for i in [0, count - 1).
            const size_t fwd = count - 1;
            const int chunks = chunk_count(fwd, inner, 65536);
            std::vector<Box> part(static_cast<size_t>(chunks));
            parallel_chunks(fwd, inner, 65536, [&](int c, size_t lo, size_t hi)
            {
                Box p; p.reset();
                for (size_t i = lo; i < hi; ++i) p.grow(m_boxes[order[i]]);
                part[static_cast<size_t>(c)] = p;
            });
            std::vector<Box> before(static_cast<size_t>(chunks));
            acc.reset();
            for (int c = 0; c < chunks; ++c) { before[static_cast<size_t>(c)] = acc; acc.grow(part[static_cast<size_t>(c)]); }
            parallel_chunks(fwd, inner, 65536, [&](int c, size_t lo, size_t hi)
            {
                Box run = before[static_cast<size_t>(c)];
                for (size_t i = lo; i < hi; ++i) { run.grow(m_boxes[order[i]]); prefix[i] = run.half_area(); }
            });

            // Backward: positions count - 1 down to 1, i.e. offsets k = i - 1 in [0, count - 1).
            parallel_chunks(fwd, inner, 65536, [&](int c, size_t lo, size_t hi)
            {
                Box p; p.reset();
                for (size_t k = lo; k < hi; ++k) p.grow(m_boxes[order[k + 1]]);
                part[static_cast<size_t>(c)] = p;
            });
            std::vector<Box> after(static_cast<size_t>(chunks));
            acc.reset();
            for (int c = chunks - 1; c >= 0; --c) { after[static_cast<size_t>(c)] = acc; acc.grow(part[static_cast<size_t>(c)]); }
            std::vector<Candidate> local(static_cast<size_t>(chunks));
            parallel_chunks(fwd, inner, 65536, [&](int c, size_t lo, size_t hi)
            {
                Candidate b = { std::numeric_limits<T>::max(), 0 };
                Box run = after[static_cast<size_t>(c)];
                for (size_t k = hi; k-- > lo; )
                {
                    const size_t i = k + 1;
                    run.grow(m_boxes[order[i]]);
                    const T left_cost = prefix[i - 1] * i;
                    const T right_cost = run.half_area() * (count - i);
                    const T cost = left_cost + right_cost;
                    if (b.cost > cost) { b.cost = cost; b.pivot = i; }
                }
                local[static_cast<size_t>(c)] = b;
            });
            // Same scan order as the serial sweep (high positions first), same strict '>'.
            Candidate best = { std::numeric_limits<T>::max(), 0 };
            for (int c = chunks - 1; c >= 0; --c)
                if (best.cost > local[static_cast<size_t>(c)].cost) best = local[static_cast<size_t>(c)];
            return best;
        }

        acc.reset();
        for (size_t i = 0; i + 1 < count; ++i)
        {
            acc.grow(m_boxes[order[i]]);
            prefix[i] = acc.half_area();
        }

        Candidate best = { std::numeric_limits<T>::max(), 0 };
        acc.reset();
        for (size_t i = count - 1; i > 0; --i)
        {
            acc.grow(m_boxes[order[i]]);
            const T left_cost = prefix[i - 1] * i;
            const T right_cost = acc.half_area() * (count - i);
            const T cost = left_cost + right_cost;
            if (best.cost > cost) { best.cost = cost; best.pivot = i; }
        }
        return best;
    }

    // Returns the pivot (absolute position) or `end` when the range becomes a leaf.
    size_t split(const size_t begin, const size_t end, const Box& box, const bool concurrent_axes, const int inner = 1)
    {
        if (box.rank() < 2) return end;                 // only degenerate items
        const size_t count = end - begin;
        if (count <= m_max_leaf) return end;

        Candidate cand[3];
        if (concurrent_axes)
        {
            std::future<Candidate> f1 = std::async(std::launch::async, [=]() { return sweep_axis(1, begin, end, inner); });
            std::future<Candidate> f2 = std::async(std::launch::async, [=]() { return sweep_axis(2, begin, end, inner); });
            cand[0] = sweep_axis(0, begin, end, inner);
            cand[1] = f1.get();
            cand[2] = f2.get();
        }
        else for (int a = 0; a < 3; ++a) cand[a] = sweep_axis(a, begin, end);

        // First axis / first pivot wins ties: strict '>' in axis order (bvh_sahpartitioner.h:144).
        T best_cost = std::numeric_limits<T>::max();
        int best_axis = 0;
        size_t best_pivot = 0;
        for (int a = 0; a < 3; ++a)
            if (best_cost > cand[a].cost) { best_cost = cand[a].cost; best_axis = a; best_pivot = cand[a].pivot; }

        // No candidate at all: every cost overflowed or is NaN (coordinates near FLT_MAX, non-finite
        // vertices).  The reference would split at `begin` and recurse for ever; make a leaf instead.
        if (best_pivot == 0) return end;

        const T split_cost = m_ct + best_cost / box.half_area() * m_ci;
        const T leaf_cost = count * m_ci;
        if (leaf_cost <= split_cost) return end;

        const size_t pivot = begin + best_pivot;
        repartition(best_axis, begin, end, pivot, concurrent_axes ? inner * 3 : inner);
        return pivot;
    }

    // PartitionerBase::sort_indices (bvh_partitionerbase.h:135-198): stable partition of the two
    // other axes' orders by membership in the left set of the split axis.
    void repartition(const int axis, const size_t begin, const size_t end, const size_t pivot, const int inner = 1)
    {
        const uint32_t* split_order = m_order[axis].data();
        if (inner > 1 && end - begin >= ParallelSweepMin)
        {
            // The same stable partition, chunk by chunk: the number of left items before each chunk
            // fixes where its items land.
            parallel_chunks(end - begin, inner, 65536, [&](int, size_t lo, size_t hi)
            {
                for (size_t i = begin + lo; i < begin + hi; ++i) m_side[split_order[i]] = i < pivot ? 0 : 1;
            });
            for (int a = 0; a < 3; ++a)
            {
                if (a == axis) continue;
                uint32_t* order = m_order[a].data();
                const int chunks = chunk_count(end - begin, inner, 65536);
                std::vector<size_t> lefts(static_cast<size_t>(chunks) + 1, 0);
                parallel_chunks(end - begin, inner, 65536, [&](int c, size_t lo, size_t hi)
                {
                    size_t n_left = 0;
                    for (size_t i = begin + lo; i < begin + hi; ++i) n_left += m_side[order[i]] == 0 ? 1 : 0;
                    lefts[static_cast<size_t>(c) + 1] = n_left;
                });
                for (int c = 0; c < chunks; ++c) lefts[static_cast<size_t>(c) + 1] += lefts[static_cast<size_t>(c)];
                parallel_chunks(end - begin, inner, 65536, [&](int c, size_t lo, size_t hi)
                {
                    size_t l = begin + lefts[static_cast<size_t>(c)];
                    size_t r = pivot + (lo - lefts[static_cast<size_t>(c)]);
                    for (size_t i = begin + lo; i < begin + hi; ++i)
                    {
                        const uint32_t item = order[i];
                        if (m_side[item] == 0) m_scratch[l++] = item;
                        else m_scratch[r++] = item;
                    }
                });
                parallel_chunks(end - begin, inner, 65536, [&](int, size_t lo, size_t hi)
                {
                    std::memcpy(order + begin + lo, m_scratch.data() + begin + lo, (hi - lo) * sizeof(uint32_t));
                });
            }
            return;
        }
        for (size_t i = begin; i < pivot; ++i) m_side[split_order[i]] = 0;
        for (size_t i = pivot; i < end; ++i) m_side[split_order[i]] = 1;
        for (int a = 0; a < 3; ++a)
        {
            if (a == axis) continue;
            uint32_t* order = m_order[a].data();
            size_t l = begin, r = pivot;
            for (size_t i = begin; i < end; ++i)
            {
                const uint32_t item = order[i];
                if (m_side[item] == 0) m_scratch[l++] = item;
                else m_scratch[r++] = item;
            }
            std::memcpy(order + begin, m_scratch.data() + begin, (end - begin) * sizeof(uint32_t));
        }
    }

    static void set_child_box(AsNode& node, const int side, const BoundsD& b)
    {
        for (int a = 0; a < 3; ++a)
        {
            node.bbox[a * 4 + side] = b.lo[a];
            node.bbox[a * 4 + 2 + side] = b.hi[a];
        }
    }

    // The two children of a top node cover disjoint ranges of every shared array (orders, prefix
    // areas, scratch: by position; side flags: by item), so they are expanded concurrently down
    // to `fork_depth` levels: the root still costs one pass over all items per axis, but the
    // levels below it cost one such pass TOGETHER instead of one each (the serial top was 57 % of
    // a 2 M-item build).  The tree does not depend on the order of execution: m_top / m_jobs are
    // only appended to (under the lock) and addressed by index.
    size_t new_top_node()
    {
        std::lock_guard<std::mutex> lock(m_top_mutex);
        m_top.push_back(TopNode());
        return m_top.size() - 1;
    }

    void make_deferred(const size_t self, const size_t begin, const size_t end, const BoundsD& box)
    {
        std::lock_guard<std::mutex> lock(m_top_mutex);
        m_top[self].kind = Deferred;
        m_top[self].job = m_jobs.size();
        Job job;
        std::memset(&job.root, 0, sizeof(job.root));
        job.begin = begin; job.end = end; job.box = box;
        m_jobs.push_back(std::move(job));
        m_jobs.back().nodes.clear();
    }

    size_t expand_top(const size_t begin, const size_t end, const BoundsD& box, const int depth = 0)
    {
        const size_t self = new_top_node();
        if (end - begin <= m_grain)
        {
            make_deferred(self, begin, end, box);
            return self;
        }
        // Threads of this node: the machine's, shared by the 2^depth nodes of its level and the 3 axes.
        const int inner = std::max(1, m_threads / (3 << std::min(depth, 20)));
        const size_t pivot = split(begin, end, narrow(box), m_threads >= 3, inner);
        if (pivot == end)
        {
            // A large range that refuses to split (all-degenerate boxes): a single-leaf job.
            make_deferred(self, begin, end, box);
            return self;
        }
        const BoundsD lb = widen(range_bounds(begin, pivot, inner * 3));
        const BoundsD rb = widen(range_bounds(pivot, end, inner * 3));
        size_t l, r;
        if (depth < m_fork_depth && end - begin > 4 * m_grain)
        {
            std::future<size_t> left = std::async(std::launch::async, [=]() { return expand_top(begin, pivot, lb, depth + 1); });
            r = expand_top(pivot, end, rb, depth + 1);
            l = left.get();
        }
        else
        {
            l = expand_top(begin, pivot, lb, depth + 1);
            r = expand_top(pivot, end, rb, depth + 1);
        }
        std::lock_guard<std::mutex> lock(m_top_mutex);
        TopNode& t = m_top[self];
        t.kind = Interior;
        t.left = l; t.right = r;
        t.left_box = lb; t.right_box = rb;
        return self;
    }

    // Builds one deferred range with an explicit stack; child indices are relative to job.nodes[0].
    void build_range(Job& job)
    {
        struct Task { size_t slot; size_t begin, end; BoundsD box; };   // slot: index in job.nodes, or ~0 for job.root
        AsNode blank; std::memset(&blank, 0, sizeof(blank));
        job.root = blank;
        job.nodes.clear();
        job.nodes.reserve(2 * (job.end - job.begin) / (m_max_leaf ? m_max_leaf : 1) + 2);

        std::vector<Task> stack;
        stack.push_back(Task{ ~size_t(0), job.begin, job.end, job.box });
        while (!stack.empty())
        {
            const Task t = stack.back();
            stack.pop_back();

            size_t pivot = t.end;
            if (t.end - t.begin > 1)
                pivot = split(t.begin, t.end, narrow(t.box), false);

            if (pivot == t.end)
            {
                AsNode& node = t.slot == ~size_t(0) ? job.root : job.nodes[t.slot];
                node.item_count = static_cast<uint32_t>(t.end - t.begin);
                node.index = static_cast<uint32_t>(t.begin);
                continue;
            }

            const BoundsD lb = widen(range_bounds(t.begin, pivot));
            const BoundsD rb = widen(range_bounds(pivot, t.end));
            const size_t pair = job.nodes.size();
            job.nodes.push_back(blank);
            job.nodes.push_back(blank);
            AsNode& node = t.slot == ~size_t(0) ? job.root : job.nodes[t.slot];
            node.item_count = 0xFFFFFFFFu;
            node.index = static_cast<uint32_t>(pair);
            set_child_box(node, 0, lb);
            set_child_box(node, 1, rb);
            // Left subtree first: push right, then left.
            stack.push_back(Task{ pair + 1, pivot, t.end, rb });
            stack.push_back(Task{ pair, t.begin, pivot, lb });
        }
    }

    void place(const size_t top_index, const size_t slot, size_t& cursor, AsNodeVector& nodes)
    {
        const TopNode& t = m_top[top_index];
        if (t.kind == Interior)
        {
            const size_t pair = cursor;
            cursor += 2;
            AsNode& node = nodes[slot];
            std::memset(&node, 0, sizeof(node));
            node.item_count = 0xFFFFFFFFu;
            node.index = static_cast<uint32_t>(pair);
            set_child_box(node, 0, t.left_box);
            set_child_box(node, 1, t.right_box);
            place(t.left, pair, cursor, nodes);
            place(t.right, pair + 1, cursor, nodes);
        }
        else
        {
            const Job& job = m_jobs[t.job];
            m_placed.push_back(Placed{ t.job, slot, cursor });
            cursor += job.nodes.size();
        }
    }
};

//
// Triangle collection (triangletree.cpp:105-383).
//

struct TriInfo { uint64_t first_vertex; uint32_t msc; uint32_t vis; };

struct Collected
{
    std::vector<AsTriangleKey>  keys;
    std::vector<TriInfo>        infos;
    std::vector<P3>             vertices;       // assembly space, float; (msc + 1) * 3 per triangle, pose-major
    std::vector<BoundsF>        boxes;          // build boxes (mid-time box for moving triangles)
};

inline P3 vertex_at(const asgpu_mesh& m, const size_t i)
{
    const P3 p = { m.vertices[i * 3], m.vertices[i * 3 + 1], m.vertices[i * 3 + 2] };
    return p;
}

inline P3 pose_at(const asgpu_mesh& m, const size_t v, const size_t seg)
{
    const float* q = m.vertex_poses + (v * m.motion_segment_count + seg) * 3;
    const P3 p = { q[0], q[1], q[2] };
    return p;
}

// First two stages of foundation::intersect(bbox, v0, v1, v2) (intersection/aabbtriangle.h):
// a vertex inside the box accepts, all vertices beyond one face rejects.  The tree box is the
// union of the object instances' boxes (assemblytree.cpp:401-405), so triangles of this assembly
// always take the first exit; anything else is kept (conservative).
inline bool touches(const BoundsF& b, const P3& v0, const P3& v1, const P3& v2)
{
    const P3* vs[3] = { &v0, &v1, &v2 };
    unsigned all = 0;
    for (int k = 0; k < 3; ++k)
    {
        const float p[3] = { vs[k]->x, vs[k]->y, vs[k]->z };
        unsigned m = 0;
        for (int a = 0; a < 3; ++a)
        {
            if (p[a] >= b.lo[a]) m |= 1u << a;
            if (p[a] <= b.hi[a]) m |= 8u << a;
        }
        if (m == 0x3F) return true;
        all |= m;
    }
    return all == 0x3F;
}

// Triangles [tri_begin, tri_end) of object instance `oi`, appended to `out` (first_vertex counts from
// the start of out.vertices).
void collect_range(const asgpu_scene_desc& desc, const asgpu_assembly& assembly, const BoundsF& tree_box, const uint32_t oi,
                   const uint32_t tri_begin, const uint32_t tri_end, Collected& out)
{
    const double time = assembly.time;
    uint64_t vertex_cursor = out.vertices.size();
    std::vector<BoundsF> pose_box;
    {
        const asgpu_object_instance& inst = assembly.object_instances[oi];
        const asgpu_mesh& mesh = desc.meshes[inst.mesh_index];
        const double* m = inst.local_to_parent;
        const uint32_t msc = mesh.motion_segment_count;
        pose_box.resize(msc + 1);

        for (uint32_t ti = tri_begin; ti < tri_end; ++ti)
        {
            const uint32_t* tv = mesh.triangles + size_t(ti) * 3;
            BoundsF build_box;
            const size_t vertex_mark = out.vertices.size();

            if (msc == 0)
            {
                const P3 a_os = vertex_at(mesh, tv[0]), b_os = vertex_at(mesh, tv[1]), c_os = vertex_at(mesh, tv[2]);
                if (degenerate(a_os, b_os, c_os)) continue;
                const P3 a = to_parent(m, a_os), b = to_parent(m, b_os), c = to_parent(m, c_os);
                if (degenerate(a, b, c)) continue;
                if (!touches(tree_box, a, b, c)) continue;
                build_box.reset();
                build_box.grow(a.x, a.y, a.z); build_box.grow(b.x, b.y, b.z); build_box.grow(c.x, c.y, c.z);
                out.vertices.push_back(a); out.vertices.push_back(b); out.vertices.push_back(c);
            }
            else
            {
                // Pose 0 is the base mesh, poses 1..msc the vertex poses (triangletree.cpp:238-256).
                for (uint32_t s = 0; s <= msc; ++s)
                {
                    pose_box[s].reset();
                    for (int k = 0; k < 3; ++k)
                    {
                        const P3 p = to_parent(m, s == 0 ? vertex_at(mesh, tv[k]) : pose_at(mesh, tv[k], s - 1));
                        pose_box[s].grow(p.x, p.y, p.z);
                        out.vertices.push_back(p);
                    }
                }
                BoundsF motion_box = pose_box[0];
                for (uint32_t s = 1; s <= msc; ++s) motion_box.grow(pose_box[s]);
                bool keep = motion_box.rank() >= 2;
                for (int a = 0; keep && a < 3; ++a)
                    if (tree_box.lo[a] > motion_box.hi[a] || tree_box.hi[a] < motion_box.lo[a]) keep = false;
                if (keep)
                {
                    // Box at acceleration_structure.time (renderer/utility/bbox.h:110-125):
                    // lerp(a, b, k) = (1 - k) * a + k * b in float.
                    const size_t prev = static_cast<size_t>(time * msc);
                    const float k = static_cast<float>(time * msc - prev);
                    const float w = 1.0f - k;
                    for (int a = 0; a < 3; ++a)
                    {
                        build_box.lo[a] = w * pose_box[prev].lo[a] + k * pose_box[prev + 1].lo[a];
                        build_box.hi[a] = w * pose_box[prev].hi[a] + k * pose_box[prev + 1].hi[a];
                    }
                    keep = build_box.rank() >= 2;
                }
                if (!keep) { out.vertices.resize(vertex_mark); continue; }
            }

            AsTriangleKey key; std::memset(&key, 0, sizeof(key));
            key.object_instance_index = oi;
            key.triangle_index = ti;
            key.triangle_pa = mesh.triangle_pa ? mesh.triangle_pa[ti] : 0;
            out.keys.push_back(key);
            const TriInfo info = { vertex_cursor, msc, inst.vis_flags };
            out.infos.push_back(info);
            out.boxes.push_back(build_box);
            vertex_cursor += uint64_t(msc + 1) * 3;
        }
    }
}

// Every triangle of the assembly, in the reference's order (object instances, then triangles:
// triangletree.cpp:105-383).  The triangles of an instance are collected by several threads over
// contiguous ranges and stitched in order, so the result does not depend on the number of threads.
void collect(const asgpu_scene_desc& desc, const asgpu_assembly& assembly, const BoundsF& tree_box, Collected& out, const int threads)
{
    for (uint32_t oi = 0; oi < assembly.object_instance_count; ++oi)
    {
        const asgpu_mesh& mesh = desc.meshes[assembly.object_instances[oi].mesh_index];
        const int chunks = chunk_count(mesh.triangle_count, threads, 1 << 14);
        if (chunks <= 1) { collect_range(desc, assembly, tree_box, oi, 0, mesh.triangle_count, out); continue; }
        std::vector<Collected> parts(chunks);
        parallel_chunks(mesh.triangle_count, threads, 1 << 14, [&](int c, size_t begin, size_t end)
        {
            collect_range(desc, assembly, tree_box, oi, static_cast<uint32_t>(begin), static_cast<uint32_t>(end), parts[c]);
        });
        std::vector<size_t> tri_at(chunks + 1), vertex_begin(chunks + 1);
        tri_at[0] = out.keys.size(); vertex_begin[0] = out.vertices.size();
        for (int c = 0; c < chunks; ++c) { tri_at[c + 1] = tri_at[c] + parts[c].keys.size(); vertex_begin[c + 1] = vertex_begin[c] + parts[c].vertices.size(); }
        out.keys.resize(tri_at[chunks]); out.infos.resize(tri_at[chunks]); out.boxes.resize(tri_at[chunks]);
        out.vertices.resize(vertex_begin[chunks]);
        parallel_chunks(static_cast<size_t>(chunks), threads, 1, [&](int, size_t begin, size_t end)
        {
            for (size_t c = begin; c < end; ++c)
            {
                const Collected& p = parts[c];
                if (p.keys.empty()) continue;
                std::memcpy(&out.keys[tri_at[c]], p.keys.data(), p.keys.size() * sizeof(p.keys[0]));
                std::memcpy(&out.boxes[tri_at[c]], p.boxes.data(), p.boxes.size() * sizeof(p.boxes[0]));
                std::memcpy(&out.vertices[vertex_begin[c]], p.vertices.data(), p.vertices.size() * sizeof(p.vertices[0]));
                for (size_t i = 0; i < p.infos.size(); ++i)
                {
                    TriInfo info = p.infos[i];
                    info.first_vertex += vertex_begin[c];
                    out.infos[tri_at[c] + i] = info;
                }
            }
        });
    }
}

//
// Motion boxes (triangletree.cpp:755-876), bottom-up with an explicit post-order walk.
//

typedef std::vector<BoundsF> BoxSeq;

BoxSeq leaf_motion_boxes(const AsNode& node, const std::vector<uint32_t>& order, const Collected& c)
{
    uint32_t max_msc = 0;
    BoundsF base; base.reset();
    for (uint32_t i = 0; i < node.item_count; ++i)
    {
        const TriInfo& info = c.infos[order[node.index + i]];
        max_msc = std::max(max_msc, info.msc);
        for (int k = 0; k < 3; ++k)
        {
            const P3& p = c.vertices[info.first_vertex + k];
            base.grow(p.x, p.y, p.z);
        }
    }
    BoxSeq seq(max_msc + 1);
    seq[0] = base;
    if (max_msc == 0) return seq;

    for (uint32_t s = 1; s < max_msc; ++s)
    {
        seq[s].reset();
        const double time = static_cast<double>(s) / max_msc;
        for (uint32_t i = 0; i < node.item_count; ++i)
        {
            const TriInfo& info = c.infos[order[node.index + i]];
            const size_t prev = static_cast<size_t>(time * info.msc);
            const float k = static_cast<float>(time * info.msc - prev);
            const float w = 1.0f - k;
            const P3* a = &c.vertices[info.first_vertex + prev * 3];
            for (int v = 0; v < 3; ++v)
                seq[s].grow(w * a[v].x + k * a[v + 3].x, w * a[v].y + k * a[v + 3].y, w * a[v].z + k * a[v + 3].z);
        }
    }
    seq[max_msc].reset();
    for (uint32_t i = 0; i < node.item_count; ++i)
    {
        const TriInfo& info = c.infos[order[node.index + i]];
        const P3* a = &c.vertices[info.first_vertex + size_t(info.msc) * 3];
        for (int v = 0; v < 3; ++v) seq[max_msc].grow(a[v].x, a[v].y, a[v].z);
    }
    return seq;
}

void append_swizzled(std::vector<double>& dst, const BoxSeq& seq)      // triangletree.cpp:725-738
{
    for (const BoundsF& b : seq)
        for (int a = 0; a < 3; ++a)
        {
            dst.push_back(static_cast<double>(b.lo[a]));
            dst.push_back(static_cast<double>(b.hi[a]));
        }
}

// The reference recursion numbers m_node_bboxes entries in post-order (left subtree, right
// subtree, then this node's left and right sequences); the explicit stack below reproduces it.
void propagate_motion_boxes(HostTriangleTree& tree, const std::vector<uint32_t>& order, const Collected& c)
{
    if (tree.moving_triangle_count == 0)
    {
        // Every sequence has length one: the walk would only set both counts to 1.
        for (AsNode& node : tree.nodes)
            if (node.interior()) node.left_bbox_count = node.right_bbox_count = 1;
        return;
    }

    struct Frame { uint32_t node; int stage; BoxSeq left; };
    std::vector<Frame> stack;
    std::vector<BoxSeq> results;        // return values travelling up
    stack.push_back(Frame{ 0, 0, BoxSeq() });
    while (!stack.empty())
    {
        Frame& f = stack.back();
        AsNode& node = tree.nodes[f.node];
        if (!node.interior())
        {
            results.push_back(leaf_motion_boxes(node, order, c));
            stack.pop_back();
            continue;
        }
        if (f.stage == 0)
        {
            f.stage = 1;
            const uint32_t child = node.index;
            stack.push_back(Frame{ child, 0, BoxSeq() });
            continue;
        }
        if (f.stage == 1)
        {
            f.left = std::move(results.back());
            results.pop_back();
            f.stage = 2;
            const uint32_t child = node.index + 1;
            stack.push_back(Frame{ child, 0, BoxSeq() });
            continue;
        }
        BoxSeq right = std::move(results.back());
        results.pop_back();
        const BoxSeq& left = f.left;

        node.left_bbox_count = static_cast<uint32_t>(left.size());
        node.right_bbox_count = static_cast<uint32_t>(right.size());
        if (left.size() > 1)
        {
            node.left_bbox_index = static_cast<uint32_t>(tree.node_bboxes.size() / 6);
            append_swizzled(tree.node_bboxes, left);
        }
        if (right.size() > 1)
        {
            node.right_bbox_index = static_cast<uint32_t>(tree.node_bboxes.size() / 6);
            append_swizzled(tree.node_bboxes, right);
        }
        const size_t count = std::max(left.size(), right.size());
        BoxSeq merged(count);
        for (size_t i = 0; i < count; ++i)
        {
            merged[i] = left[i * left.size() / count];
            merged[i].grow(right[i * right.size() / count]);
        }
        results.push_back(std::move(merged));
        stack.pop_back();
    }
}

//
// Leaf payloads (triangleencoder.cpp:48-103, triangletree.cpp:878-978).
//

size_t payload_size(const AsNode& node, const std::vector<uint32_t>& order, const Collected& c)
{
    size_t s = 0;
    for (uint32_t i = 0; i < node.item_count; ++i)
    {
        const uint32_t msc = c.infos[order[node.index + i]].msc;
        s += 8 + (msc == 0 ? AsTriangleBytes : (size_t(msc) + 1) * AsPoseBytes);
    }
    return s;
}

uint8_t* write_payload(uint8_t* out, const uint32_t first, const uint32_t count, const std::vector<uint32_t>& order, const Collected& c)
{
    for (uint32_t i = 0; i < count; ++i)
    {
        const TriInfo& info = c.infos[order[first + i]];
        std::memcpy(out, &info.vis, 4); out += 4;
        std::memcpy(out, &info.msc, 4); out += 4;
        const P3* v = &c.vertices[info.first_vertex];
        if (info.msc == 0)
        {
            // TriangleMT<float>(v0, v1, v2): edges in float (raytrianglemt.h:128-137).
            const float tri[9] = { v[0].x, v[0].y, v[0].z,
                                   v[1].x - v[0].x, v[1].y - v[0].y, v[1].z - v[0].z,
                                   v[2].x - v[0].x, v[2].y - v[0].y, v[2].z - v[0].z };
            std::memcpy(out, tri, AsTriangleBytes); out += AsTriangleBytes;
        }
        else
        {
            const size_t bytes = (size_t(info.msc) + 1) * AsPoseBytes;
            std::memcpy(out, v, bytes); out += bytes;
        }
    }
    return out;
}

// Leaf payloads (triangletree.cpp:878-978).  Keys and spilled payloads land where a sequential walk
// over the nodes would put them: every node range knows its first key and its first spill byte from a
// counting pass, then the ranges are written by several threads.
void store_leaves(HostTriangleTree& tree, const std::vector<uint32_t>& order, const Collected& c, const int threads)
{
    const size_t in_node_limit = AsNodeUserDataSize - sizeof(uint32_t);     // 92 bytes
    const size_t node_count = tree.nodes.size();
    const size_t grain = 1 << 14;
    const int chunks = chunk_count(node_count, threads, grain);
    std::vector<size_t> key_at(chunks + 1, 0), spill_at(chunks + 1, 0);
    parallel_chunks(node_count, threads, grain, [&](int chunk, size_t begin, size_t end)
    {
        size_t keys = 0, spill = 0;
        for (size_t n = begin; n < end; ++n)
        {
            const AsNode& node = tree.nodes[n];
            if (node.interior()) continue;
            keys += node.item_count;
            const size_t s = payload_size(node, order, c);
            if (s > in_node_limit) spill += s;
        }
        key_at[chunk + 1] = keys; spill_at[chunk + 1] = spill;
    });
    for (int k = 0; k < chunks; ++k) { key_at[k + 1] += key_at[k]; spill_at[k + 1] += spill_at[k]; }
    tree.leaf_data.resize(spill_at[chunks]);
    tree.keys.resize(key_at[chunks]);

    parallel_chunks(node_count, threads, grain, [&](int chunk, size_t begin, size_t end)
    {
        size_t key_cursor = key_at[chunk];
        uint8_t* spill_writer = tree.leaf_data.data() + spill_at[chunk];
        for (size_t n = begin; n < end; ++n)
        {
            AsNode& node = tree.nodes[n];
            if (node.interior()) continue;
            const uint32_t first = node.index, count = node.item_count;
            const size_t s = payload_size(node, order, c);
            node.index = static_cast<uint32_t>(key_cursor);
            for (uint32_t j = 0; j < count; ++j) tree.keys[key_cursor++] = c.keys[order[first + j]];
            uint8_t* user = node.user_data();
            if (s <= in_node_limit)
            {
                const uint32_t in_node = 0xFFFFFFFFu;
                std::memcpy(user, &in_node, 4);
                write_payload(user + 4, first, count, order, c);
            }
            else
            {
                const uint32_t offset = static_cast<uint32_t>(spill_writer - tree.leaf_data.data());
                std::memcpy(user, &offset, 4);
                spill_writer = write_payload(spill_writer, first, count, order, c);
            }
        }
    });
}

bool check_desc(const asgpu_scene_desc& d, std::string& error)
{
    if ((d.mesh_count && !d.meshes) || (d.assembly_count && !d.assemblies) || (d.assembly_instance_count && !d.assembly_instances))
    { error = "scene description has a null array"; return false; }
    for (uint32_t i = 0; i < d.mesh_count; ++i)
    {
        const asgpu_mesh& m = d.meshes[i];
        if ((m.vertex_count && !m.vertices) || (m.triangle_count && !m.triangles)) { error = "mesh with null vertex/triangle array"; return false; }
        if (m.motion_segment_count && !m.vertex_poses) { error = "moving mesh without vertex poses"; return false; }
        for (size_t k = 0; k < size_t(m.triangle_count) * 3; ++k)
            if (m.triangles[k] >= m.vertex_count) { error = "triangle vertex index out of range"; return false; }
    }
    for (uint32_t a = 0; a < d.assembly_count; ++a)
    {
        const asgpu_assembly& as = d.assemblies[a];
        if (as.object_instance_count && !as.object_instances) { error = "assembly with null object instance array"; return false; }
        if (as.max_leaf_size == 0) { error = "max_leaf_size must be positive"; return false; }
        if (!(as.time >= 0.0 && as.time < 1.0)) { error = "acceleration_structure.time must be in [0, 1)"; return false; }
        for (uint32_t o = 0; o < as.object_instance_count; ++o)
            if (as.object_instances[o].mesh_index >= d.mesh_count) { error = "object instance mesh index out of range"; return false; }
    }
    for (uint32_t i = 0; i < d.assembly_instance_count; ++i)
        if (d.assembly_instances[i].assembly_index >= d.assembly_count) { error = "assembly instance index out of range"; return false; }
    return true;
}

}   // anonymous namespace

//
// Top of a clustered device build (ploc.cu): sweep SAH over the clusters the rounds left.
//
// Agglomeration is good at the bottom of a tree and poor at its top (greedy merges of ever larger
// clusters: +19 % / +32 % wide-node visits against the sweep tree on the C2 grid, incoherent /
// coherent rays), the sweep SAH is the other way round (its cost is in the many small nodes).  So
// the device rounds stop when n / PlocTopRatio clusters are left and this builder -- the product's
// own sweep SAH, the reference's partitioner (bvh_sahpartitioner.h:99-170) -- builds the top over
// their boxes.  Nodes are numbered breadth first (root 0, parents before children, every level a
// contiguous index range): the numbering the clustering rounds continue downwards from.
//

bool build_cluster_top(const float* cbox, const uint32_t* cref, const uint32_t* ccount, const size_t count, const int threads,
                       ClusterTop& out, std::string& error)
{
    out.left.clear(); out.right.clear(); out.leaves.clear(); out.boxes.clear(); out.level_begin.clear();
    if (count < 2) { error = "cluster top: fewer than two clusters"; return false; }
    std::vector<BoundsF> boxes(count);
    for (size_t i = 0; i < count; ++i)
        for (int a = 0; a < 3; ++a) { boxes[i].lo[a] = cbox[i * 6 + a]; boxes[i].hi[a] = cbox[i * 6 + 3 + a]; }
    AsNodeVector tree;
    SweepBuilder<float> builder(boxes, 1, 0.0f, 1.0f, threads);
    builder.build(tree);
    const std::vector<uint32_t>& order = builder.ordering();

    // A subtree of the top is a node of the sweep tree, or (where the sweep made a leaf of several
    // clusters: degenerate boxes) a range of its ordering that is halved until single clusters remain.
    struct Task { uint32_t node; uint32_t begin, end; };        // node != None: sweep node; else the range [begin, end)
    const uint32_t NoNode = 0xFFFFFFFFu;
    auto task_of = [&tree, NoNode](const uint32_t node) -> Task
    {
        const AsNode& nd = tree[node];
        if (nd.item_count == 0xFFFFFFFFu) return Task{ node, 0u, 0u };
        return Task{ NoNode, nd.index, nd.index + nd.item_count };
    };
    auto is_cluster = [NoNode](const Task& t) { return t.node == NoNode && t.end - t.begin == 1; };

    // Pass 1, breadth first: indices and children.
    std::vector<Task> tasks;                    // interior nodes of the top, in index order
    std::vector<Task> kids;                     // two per interior node
    const Task root = task_of(0);
    if (is_cluster(root)) { error = "cluster top: the sweep left one cluster"; return false; }
    tasks.push_back(root);
    out.level_begin.push_back(0);
    size_t level_end = 1;
    for (size_t i = 0; i < tasks.size(); ++i)
    {
        if (i == level_end) { out.level_begin.push_back(static_cast<uint32_t>(i)); level_end = tasks.size(); }
        const Task t = tasks[i];
        Task child[2];
        if (t.node != NoNode)
        {
            child[0] = task_of(tree[t.node].index);
            child[1] = task_of(tree[t.node].index + 1);
        }
        else
        {
            const uint32_t mid = t.begin + (t.end - t.begin) / 2;
            child[0] = Task{ NoNode, t.begin, mid };
            child[1] = Task{ NoNode, mid, t.end };
        }
        for (int side = 0; side < 2; ++side)
        {
            kids.push_back(child[side]);
            if (!is_cluster(child[side])) tasks.push_back(child[side]);
        }
    }
    out.level_begin.push_back(static_cast<uint32_t>(tasks.size()));
    if (tasks.size() != count - 1) { error = "cluster top: node count does not match the cluster count"; return false; }

    // Child references: interior children were appended in the order they were met.
    const size_t nodes = tasks.size();
    out.left.resize(nodes); out.right.resize(nodes); out.leaves.assign(nodes, 0); out.boxes.assign(nodes * 6, 0.0f);
    uint32_t next_interior = 1;
    for (size_t i = 0; i < nodes; ++i)
        for (int side = 0; side < 2; ++side)
        {
            const Task& c = kids[i * 2 + side];
            const uint32_t ref = is_cluster(c) ? cref[order[c.begin]] : next_interior++;
            (side == 0 ? out.left : out.right)[i] = ref;
        }

    // Pass 2, children before parents (children have higher indices): leaf counts and boxes.
    for (size_t i = nodes; i-- > 0; )
    {
        float* box = &out.boxes[i * 6];
        uint32_t leaves = 0;
        for (int side = 0; side < 2; ++side)
        {
            const Task& c = kids[i * 2 + side];
            const float* src;
            if (is_cluster(c)) { const uint32_t k = order[c.begin]; src = cbox + size_t(k) * 6; leaves += ccount[k]; }
            else
            {
                const uint32_t child = side == 0 ? out.left[i] : out.right[i];
                src = &out.boxes[size_t(child) * 6];
                leaves += out.leaves[child];
            }
            for (int a = 0; a < 3; ++a)
            {
                box[a] = side == 0 ? src[a] : std::min(box[a], src[a]);
                box[3 + a] = side == 0 ? src[3 + a] : std::max(box[3 + a], src[3 + a]);
            }
        }
        out.leaves[i] = leaves;
    }
    return true;
}

//
// Linear BVH -> reference node format.
//
// The topology arrives from lbvh.cu (or its host simulation); this lays it out exactly like
// bvh::Builder does (bvh_builder.h:197-228: the two children of a node are adjacent, followed by
// the whole left subtree, then the right one), with child boxes widened from float
// (bvh_builder.h:193-204) and every subtree of at most max_leaf_size items folded into one leaf,
// so that motion boxes, leaf payloads and the flattener see a tree of the usual shape.
//

static bool emit_lbvh(const LbvhTopology& t, const std::vector<BoundsF>& boxes, const size_t max_leaf_size, AsNodeVector& nodes, std::string& error, const int threads)
{
    const size_t n = boxes.size();
    if (t.order.size() != n || t.left.size() != n - 1 || t.right.size() != n - 1 || t.first.size() != n - 1 || t.last.size() != n - 1 ||
        t.node_boxes.size() != (n - 1) * 6)
    { error = "device tree build returned arrays of the wrong size"; return false; }
    std::vector<uint8_t> seen(n, 0);
    for (const uint32_t o : t.order)
    {
        if (o >= n || seen[o]) { error = "device tree build returned an ordering that is not a permutation"; return false; }
        seen[o] = 1;
    }

    auto child_box = [&](const uint32_t ref) -> BoundsD
    {
        BoundsD b;
        if (ref & LbvhLeafFlag)
        {
            const BoundsF& f = boxes[t.order[ref & ~LbvhLeafFlag]];
            for (int a = 0; a < 3; ++a) { b.lo[a] = static_cast<double>(f.lo[a]); b.hi[a] = static_cast<double>(f.hi[a]); }
        }
        else
        {
            const float* f = t.node_boxes.data() + size_t(ref) * 6;
            for (int a = 0; a < 3; ++a) { b.lo[a] = static_cast<double>(f[a]); b.hi[a] = static_cast<double>(f[3 + a]); }
        }
        return b;
    };

    // Pass 1, one thread: the depth-first walk that gives every node its place (a node's children
    // follow it as a pair, then the whole left subtree, then the right one) and checks the hierarchy.
    // Pass 2, all threads: the nodes themselves (child boxes widened to double, 128 bytes written
    // each) into an array that is not zero-filled first.
    struct Task { uint32_t slot; uint32_t ref; uint32_t first, last; uint32_t depth; };
    struct Rec { uint32_t slot, ref, first, last, pair; };         // pair = 0xFFFFFFFF: a leaf
    const uint32_t NoPair = 0xFFFFFFFFu;
    if (2 * n + 2 >= NoPair) { error = "device tree build: too many nodes"; return false; }
    // The walk reads four fields of every node it meets, in no memory order: one 16-byte record per
    // node (gathered by all threads) costs it one cache miss instead of four.
    struct Packed { uint32_t left, right, first, last; };
    std::vector<Packed> packed(n - 1);
    {
        Packed* dst = packed.data();
        parallel_chunks(n - 1, threads, size_t(1) << 16, [&](int, size_t begin, size_t end)
        {
            for (size_t i = begin; i < end; ++i) dst[i] = Packed{ t.left[i], t.right[i], t.first[i], t.last[i] };
        });
    }
    std::vector<Rec> recs;
    recs.reserve(2 * n / (max_leaf_size ? max_leaf_size : 1) + 2);
    std::vector<Task> stack;
    stack.push_back(Task{ 0u, 0u, 0u, static_cast<uint32_t>(n - 1), 1u });
    size_t placed = 0;
    uint32_t total = 1;
    while (!stack.empty())
    {
        const Task k = stack.back();
        stack.pop_back();
        const uint32_t first = k.first, last = k.last;
        if (!(k.ref & LbvhLeafFlag))
        {
            // The range an interior node reports must be the one its parent handed down.
            if (k.ref >= n - 1 || packed[k.ref].first != first || packed[k.ref].last != last) { error = "device tree build returned an inconsistent hierarchy"; return false; }
        }
        const size_t count = size_t(last) - first + 1;
        if ((k.ref & LbvhLeafFlag) || count <= max_leaf_size)
        {
            recs.push_back(Rec{ k.slot, k.ref, first, last, NoPair });
            placed += count;
            continue;
        }
        // The exact traversal keeps the reference's 64-entry stack (intersectionsettings.h:95).
        if (k.depth >= 64) { error = "linear BVH deeper than the 64-entry traversal stack (too many coincident centroids)"; return false; }
        const uint32_t l = packed[k.ref].left;
        const uint32_t split = (l & LbvhLeafFlag) ? (l & ~LbvhLeafFlag) : packed[l < n - 1 ? l : 0].last;
        if ((!(l & LbvhLeafFlag) && l >= n - 1) || split < first || split >= last) { error = "device tree build returned an inconsistent split"; return false; }
        const uint32_t pair = total;
        total += 2;
        recs.push_back(Rec{ k.slot, k.ref, first, last, pair });
        // Left subtree first: push right, then left.
        stack.push_back(Task{ pair + 1, packed[k.ref].right, split + 1, last, k.depth + 1 });
        stack.push_back(Task{ pair, l, first, split, k.depth + 1 });
    }
    if (placed != n) { error = "device tree build lost triangles"; return false; }

    nodes.clear();
    nodes.resize(total);                    // default-initialised: first touched by the threads below
    AsNode* out_nodes = nodes.data();
    const Rec* all = recs.data();
    parallel_chunks(recs.size(), threads, size_t(1) << 14, [&](int, size_t begin, size_t end)
    {
        for (size_t i = begin; i < end; ++i)
        {
            const Rec& k = all[i];
            AsNode node; std::memset(&node, 0, sizeof(node));
            if (k.pair == NoPair)
            {
                node.item_count = k.last - k.first + 1;
                node.index = k.first;
            }
            else
            {
                node.item_count = 0xFFFFFFFFu;
                node.index = k.pair;
                const BoundsD lb = child_box(packed[k.ref].left), rb = child_box(packed[k.ref].right);
                for (int a = 0; a < 3; ++a)
                {
                    node.bbox[a * 4 + 0] = lb.lo[a]; node.bbox[a * 4 + 2] = lb.hi[a];
                    node.bbox[a * 4 + 1] = rb.lo[a]; node.bbox[a * 4 + 3] = rb.hi[a];
                }
            }
            out_nodes[k.slot] = node;
        }
    });
    return true;
}

bool build_host_trees(const asgpu_scene_desc& desc, int threads, HostTrees& out, std::string& error, LbvhTopologyFn lbvh, void* lbvh_context,
                      const asgpu_instance_keys* keys)
{
    if (!check_desc(desc, error)) return false;
    if (keys)
        for (uint32_t i = 0; i < desc.assembly_instance_count; ++i)
        {
            const asgpu_instance_keys& k = keys[i];
            if (k.key_count < 2) continue;
            if (!k.times || !k.local_to_parent || !k.parent_to_local) { error = "animated assembly instance without key arrays"; return false; }
            for (uint32_t j = 0; j < k.key_count; ++j)
            {
                if (!(k.times[j] == k.times[j]) || (j > 0 && !(k.times[j - 1] < k.times[j]))) { error = "key times of an animated assembly instance must ascend strictly"; return false; }
                for (int e = 0; e < 16; ++e)
                    if (!std::isfinite(k.local_to_parent[size_t(j) * 16 + e]) || !std::isfinite(k.parent_to_local[size_t(j) * 16 + e]))
                    { error = "non-finite key transform"; return false; }
            }
        }
    if (threads < 1) threads = std::max(1u, std::thread::hardware_concurrency());
    const auto t0 = std::chrono::steady_clock::now();

    // Assembly-space boxes: Assembly::compute_non_hierarchical_local_bbox
    // (renderer/modeling/scene/assembly.cpp:219-225) over ObjectInstance::compute_parent_bbox
    // (objectinstance.cpp:255-267) over StaticTessellation::compute_local_bbox
    // (statictessellation.h:462-479).
    std::vector<BoundsF> mesh_box(desc.mesh_count);
    for (uint32_t i = 0; i < desc.mesh_count; ++i)
    {
        const asgpu_mesh& m = desc.meshes[i];
        BoundsF b; b.reset();
        for (uint32_t v = 0; v < m.vertex_count; ++v)
        {
            b.grow(m.vertices[size_t(v) * 3], m.vertices[size_t(v) * 3 + 1], m.vertices[size_t(v) * 3 + 2]);
            for (uint32_t s = 0; s < m.motion_segment_count; ++s)
            {
                const P3 p = pose_at(m, v, s);
                b.grow(p.x, p.y, p.z);
            }
        }
        mesh_box[i] = b;
    }

    out.assembly_to_tree.assign(desc.assembly_count, -1);
    std::vector<BoundsF> assembly_box(desc.assembly_count);
    for (uint32_t a = 0; a < desc.assembly_count; ++a)
    {
        const asgpu_assembly& assembly = desc.assemblies[a];
        BoundsF ab; ab.reset();
        for (uint32_t o = 0; o < assembly.object_instance_count; ++o)
        {
            const asgpu_object_instance& oi = assembly.object_instances[o];
            ab.grow(box_to_parent(oi.local_to_parent, mesh_box[oi.mesh_index]));
        }
        assembly_box[a] = ab;
        if (assembly.object_instance_count == 0) continue;

        // One triangle tree per assembly with mesh instances (assemblytree.cpp:372-420).
        out.assembly_to_tree[a] = static_cast<int>(out.triangle_trees.size());
        out.triangle_trees.emplace_back(new HostTriangleTree());
        HostTriangleTree& tree = *out.triangle_trees.back();

        // Source geometry (for the device-side refine_and_offset): private copies of the meshes.
        if (out.mesh_vertices.empty()) { out.mesh_vertices.resize(desc.mesh_count); out.mesh_triangles.resize(desc.mesh_count); out.mesh_poses.resize(desc.mesh_count); }
        tree.source_objects.resize(assembly.object_instance_count);
        for (uint32_t o = 0; o < assembly.object_instance_count; ++o)
        {
            const asgpu_object_instance& oi = assembly.object_instances[o];
            const asgpu_mesh& mesh = desc.meshes[oi.mesh_index];
            std::vector<float>& mv = out.mesh_vertices[oi.mesh_index];
            std::vector<uint32_t>& mt = out.mesh_triangles[oi.mesh_index];
            if (mv.empty() && mesh.vertex_count) mv.assign(mesh.vertices, mesh.vertices + size_t(mesh.vertex_count) * 3);
            if (mt.empty() && mesh.triangle_count) mt.assign(mesh.triangles, mesh.triangles + size_t(mesh.triangle_count) * 3);
            std::vector<float>& mp = out.mesh_poses[oi.mesh_index];
            const bool deforming = mesh.motion_segment_count != 0 && mesh.vertex_poses != nullptr && mesh.vertex_count != 0;
            if (mp.empty() && deforming) mp.assign(mesh.vertex_poses, mesh.vertex_poses + size_t(mesh.vertex_count) * mesh.motion_segment_count * 3);
            asgpu_source_object& so = tree.source_objects[o];
            std::memset(&so, 0, sizeof(so));
            so.vertices = mv.data();
            so.triangles = mt.data();
            so.vertex_poses = deforming ? mp.data() : nullptr;
            so.motion_segment_count = deforming ? mesh.motion_segment_count : 0;
            so.vertex_count = mesh.vertex_count;
            so.triangle_count = mesh.triangle_count;
            so.triangle_stride = 12;
            std::memcpy(so.parent_to_local, oi.parent_to_local, sizeof(so.parent_to_local));
        }

        const bool timing = std::getenv("ASGPU_BUILD_TIMING") != nullptr;
        auto stamp = [timing, last = std::chrono::steady_clock::now()](const char* what) mutable
        {
            if (!timing) return;
            const auto now = std::chrono::steady_clock::now();
            std::fprintf(stderr, "asgpu build: %-24s %.3f s\n", what, std::chrono::duration<double>(now - last).count());
            last = now;
        };
        stamp("source geometry");
        Collected c;
        collect(desc, assembly, ab, c, threads);
        stamp("collect");
        if (c.keys.size() >= 0xFFFFFFFFull) { error = "too many triangles in one assembly"; return false; }
        for (const TriInfo& info : c.infos) (info.msc == 0 ? tree.static_triangle_count : tree.moving_triangle_count) += 1;

        if (lbvh && c.boxes.size() >= 2 && c.boxes.size() > assembly.max_leaf_size)
        {
            if (c.boxes.size() >= LbvhLeafFlag) { error = "too many triangles in one assembly for the device tree build"; return false; }
            BoundsF root; root.reset();
            for (const BoundsF& b : c.boxes) root.grow(b);
            LbvhTopology topology;
            static_assert(sizeof(BoundsF) == 24, "boxes are passed as lo[3], hi[3]");
            const auto l0 = std::chrono::steady_clock::now();
            if (!lbvh(&c.boxes[0].lo[0], c.boxes.size(), root.lo, root.hi, lbvh_context, topology, error)) return false;
            out.topology_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - l0).count();
            stamp("topology (device)");
            if (!emit_lbvh(topology, c.boxes, assembly.max_leaf_size, tree.nodes, error, threads)) return false;
            stamp("emit nodes");
            propagate_motion_boxes(tree, topology.order, c);
            stamp("motion boxes");
            store_leaves(tree, topology.order, c, threads);
            stamp("store leaves");
            continue;
        }

        SweepBuilder<float> builder(c.boxes, assembly.max_leaf_size, assembly.interior_node_traversal_cost,
                                    assembly.triangle_intersection_cost, threads);
        builder.build(tree.nodes);
        stamp("sweep SAH");
        propagate_motion_boxes(tree, builder.ordering(), c);
        stamp("motion boxes");
        store_leaves(tree, builder.ordering(), c, threads);
        stamp("store leaves");
    }

    // Top level (assemblytree.cpp:111-245): one item per assembly instance whose assembly has
    // object instances; box = to_parent(assembly box) widened to double and robust_grow(1e-15)
    // (aabb.h:621-641); SAH with leaf size 1 and costs (1, 10) (intersectionsettings.h:51-53).
    std::vector<asgpu_assembly_item> items;
    std::vector<BoundsD> item_box;
    std::vector<uint32_t> item_instance;        // assembly instance behind every item (for its keys)
    bool any_animated = false;
    for (uint32_t i = 0; i < desc.assembly_instance_count; ++i)
    {
        const asgpu_assembly_instance& inst = desc.assembly_instances[i];
        if (desc.assemblies[inst.assembly_index].object_instance_count == 0) continue;
        const bool animated = keys != nullptr && keys[i].key_count >= 2;
        any_animated = any_animated || animated;
        asgpu_assembly_item item; std::memset(&item, 0, sizeof(item));
        std::memcpy(item.parent_to_local, animated ? keys[i].parent_to_local : inst.parent_to_local, sizeof(item.parent_to_local));
        item.assembly_instance = i;
        item.triangle_tree = static_cast<uint32_t>(out.assembly_to_tree[inst.assembly_index]);
        item.vis_flags = inst.vis_flags;
        items.push_back(item);
        item_instance.push_back(i);

        // cumulated_transform_seq.to_parent(assembly box) (assemblytree.cpp:146-149): one key = the
        // box through that transform; several = the box over the whole motion (motion_bounds.cpp).
        BoundsF wb;
        if (animated)
        {
            const BoundsF& ab = assembly_box[inst.assembly_index];
            motion_bounds(keys[i].local_to_parent, keys[i].parent_to_local, keys[i].key_count, ab.lo, ab.hi, wb.lo, wb.hi);
        }
        else wb = box_to_parent(inst.local_to_parent, assembly_box[inst.assembly_index]);
        BoundsD b;
        for (int a = 0; a < 3; ++a)
        {
            b.lo[a] = static_cast<double>(wb.lo[a]);
            b.hi[a] = static_cast<double>(wb.hi[a]);
        }
        for (int a = 0; a < 3; ++a)
        {
            const double centre = 0.5 * (b.lo[a] + b.hi[a]);
            const double extent = b.hi[a] - b.lo[a];
            double dominant = centre < 0.0 ? -centre : centre;
            if (extent > dominant) dominant = extent;
            if (!(dominant > 1.0)) dominant = 1.0;
            const double delta = dominant * 1.0e-15;
            b.lo[a] -= delta;
            b.hi[a] += delta;
        }
        item_box.push_back(b);
    }

    SweepBuilder<double> top_builder(item_box, 1, 1.0, 10.0, threads);
    top_builder.build(out.assembly_tree.nodes);
    out.assembly_tree.items.resize(items.size());
    for (size_t i = 0; i < items.size(); ++i)
        out.assembly_tree.items[i] = items[top_builder.ordering()[i]];

    // Animated instances: keys and interpolator segments per item, in tree order.
    HostAssemblyTree& top = out.assembly_tree;
    top.item_motion.clear(); top.key_times.clear(); top.key_parent_to_local.clear(); top.segments.clear();
    if (any_animated)
    {
        top.item_motion.resize(items.size());
        top.key_times.resize(items.size()); top.key_parent_to_local.resize(items.size()); top.segments.resize(items.size());
        for (size_t i = 0; i < items.size(); ++i)
        {
            asgpu_item_motion& m = top.item_motion[i];
            std::memset(&m, 0, sizeof(m));
            const asgpu_instance_keys& k = keys[item_instance[top_builder.ordering()[i]]];
            if (k.key_count < 2) continue;
            top.key_times[i].assign(k.times, k.times + k.key_count);
            top.key_parent_to_local[i].assign(k.parent_to_local, k.parent_to_local + size_t(k.key_count) * 16);
            top.segments[i].resize(k.key_count - 1);
            for (uint32_t j = 0; j + 1 < k.key_count; ++j)
                make_transform_segment(k.local_to_parent + size_t(j) * 16, k.local_to_parent + size_t(j + 1) * 16, top.segments[i][j]);
            m.key_times = top.key_times[i].data();
            m.key_parent_to_local = top.key_parent_to_local[i].data();
            m.segments = top.segments[i].data();
            m.key_count = k.key_count;
        }
    }

    out.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return true;
}

}   // namespace asgpu
