//
// oracle.cpp -- self-contained CPU restatement of appleseed's Intersector::trace() /
// trace_probe() path.  No reference header is included; every block cites the
// reference file:line it follows (paths relative to src/appleseed/ of the reference).
//
// TEST INFRASTRUCTURE ONLY (see oracle_api.h): this is the checker the CUDA engine is
// compared against and the "port" CPU baseline; the product never links it.
//
// Pinning: tests/test_oracle_kat.py replays the reference's own known-answer tests
// (foundation/meta/tests/test_intersection_raytriangle.cpp, test_intersection_rayaabb.cpp,
// test_ray.cpp, renderer/meta/tests/test_tracer.cpp, test_intersector.cpp), and
// tests/test_oracle_vs_ref.py compares trees and hit records bit-for-bit with
// oracle/_ref (the reference's own headers) wherever that library exists; golden
// outputs of oracle/_ref are committed under tests/golden/.
//
// Arithmetic contract: float geometry, double rays/boxes/tests, no FMA contraction
// (compiled with -ffp-contract=off -msse2), same operation order as the reference.
//

#include "oracle_api.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

namespace
{

const float  FLT_BIG = std::numeric_limits<float>::max();
const double DBL_BIG = std::numeric_limits<double>::max();
const double DBL_INF = std::numeric_limits<double>::infinity();

// ---------------------------------------------------------------------------------------------
// Small vector / box helpers (foundation/math/vector.h, aabb.h).
// ---------------------------------------------------------------------------------------------

struct V3f { float x, y, z; };
struct V3d { double x, y, z; };

template <typename T> struct Box
{
    T mn[3], mx[3];

    // AABBBase::invalidate (aabb.h:357-364).
    void invalidate()
    {
        for (int i = 0; i < 3; ++i)
        {
            mn[i] = std::numeric_limits<T>::max();
            mx[i] = -std::numeric_limits<T>::max();
        }
    }

    // AABBBase::insert (aabb.h:367-388).
    void insert(const T p[3])
    {
        for (int i = 0; i < 3; ++i)
        {
            if (mn[i] > p[i]) mn[i] = p[i];
            if (mx[i] < p[i]) mx[i] = p[i];
        }
    }
    void insert(const Box& b)
    {
        for (int i = 0; i < 3; ++i)
        {
            if (mn[i] > b.mn[i]) mn[i] = b.mn[i];
            if (mx[i] < b.mx[i]) mx[i] = b.mx[i];
        }
    }

    // AABBBase::rank (aabb.h:422-433).
    int rank() const
    {
        int r = 0;
        for (int i = 0; i < 3; ++i)
            if (mn[i] < mx[i]) ++r;
        return r;
    }

    bool is_valid() const
    {
        for (int i = 0; i < 3; ++i)
            if (!(mn[i] <= mx[i])) return false;
        return true;
    }
};

typedef Box<float> Box3f;
typedef Box<double> Box3d;

// half_surface_area (aabb.h:723-730).
template <typename T> inline T half_area(const Box<T>& b)
{
    const T e0 = b.mx[0] - b.mn[0];
    const T e1 = b.mx[1] - b.mn[1];
    const T e2 = b.mx[2] - b.mn[2];
    return e0 * e1 + e0 * e2 + e1 * e2;
}

inline void insert_v(Box3f& b, const V3f& v) { const float p[3] = { v.x, v.y, v.z }; b.insert(p); }

// Transform<double>::point_to_parent<float> / point_to_local<double> (transform.h:311-375):
// row-major 4x4, products accumulated left to right in double, cast, then divide by w iff w != 1.
inline V3f xform_point_f(const double* m, const V3f& p)
{
    V3f r;
    r.x = static_cast<float>(m[ 0] * double(p.x) + m[ 1] * double(p.y) + m[ 2] * double(p.z) + m[ 3]);
    r.y = static_cast<float>(m[ 4] * double(p.x) + m[ 5] * double(p.y) + m[ 6] * double(p.z) + m[ 7]);
    r.z = static_cast<float>(m[ 8] * double(p.x) + m[ 9] * double(p.y) + m[10] * double(p.z) + m[11]);
    const float w =
          static_cast<float>(m[12] * double(p.x) + m[13] * double(p.y) + m[14] * double(p.z) + m[15]);
    if (w != 1.0f) { r.x /= w; r.y /= w; r.z /= w; }
    return r;
}

inline V3d xform_point_d(const double* m, const V3d& p)
{
    V3d r;
    r.x = m[ 0] * p.x + m[ 1] * p.y + m[ 2] * p.z + m[ 3];
    r.y = m[ 4] * p.x + m[ 5] * p.y + m[ 6] * p.z + m[ 7];
    r.z = m[ 8] * p.x + m[ 9] * p.y + m[10] * p.z + m[11];
    const double w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (w != 1.0) { r.x /= w; r.y /= w; r.z /= w; }
    return r;
}

// Transform<double>::vector_to_local (transform.h:381-400).
inline V3d xform_vector_d(const double* m, const V3d& v)
{
    V3d r;
    r.x = m[0] * v.x + m[1] * v.y + m[ 2] * v.z;
    r.y = m[4] * v.x + m[5] * v.y + m[ 6] * v.z;
    r.z = m[8] * v.x + m[9] * v.y + m[10] * v.z;
    return r;
}

// Transform::to_parent(AABB) (transform.h:528-546): the 8 corners in this exact order.
inline Box3f xform_box_f(const double* m, const Box3f& b)
{
    if (!b.is_valid()) return b;
    Box3f r; r.invalidate();
    const float* lo = b.mn; const float* hi = b.mx;
    const V3f c[8] = {
        { lo[0], lo[1], lo[2] }, { lo[0], lo[1], hi[2] }, { lo[0], hi[1], hi[2] }, { lo[0], hi[1], lo[2] },
        { hi[0], hi[1], lo[2] }, { hi[0], hi[1], hi[2] }, { hi[0], lo[1], hi[2] }, { hi[0], lo[1], lo[2] } };
    for (int i = 0; i < 8; ++i) insert_v(r, xform_point_f(m, c[i]));
    return r;
}

inline Box3d xform_box_d(const double* m, const Box3d& b)
{
    if (!b.is_valid()) return b;
    Box3d r; r.invalidate();
    const double* lo = b.mn; const double* hi = b.mx;
    const V3d c[8] = {
        { lo[0], lo[1], lo[2] }, { lo[0], lo[1], hi[2] }, { lo[0], hi[1], hi[2] }, { lo[0], hi[1], lo[2] },
        { hi[0], hi[1], lo[2] }, { hi[0], hi[1], hi[2] }, { hi[0], lo[1], hi[2] }, { hi[0], lo[1], lo[2] } };
    for (int i = 0; i < 8; ++i)
    {
        const V3d p = xform_point_d(m, c[i]);
        const double q[3] = { p.x, p.y, p.z };
        r.insert(q);
    }
    return r;
}

// square_area (area.h:57-66) == 0 test, in float: 0.25 * |cross(e0, e1)|^2.
inline bool zero_area(const V3f& v0, const V3f& v1, const V3f& v2)
{
    const float e0x = v1.x - v0.x, e0y = v1.y - v0.y, e0z = v1.z - v0.z;
    const float e1x = v2.x - v0.x, e1y = v2.y - v0.y, e1z = v2.z - v0.z;
    const float cx = e0y * e1z - e1y * e0z;
    const float cy = e0z * e1x - e1z * e0x;
    const float cz = e0x * e1y - e1x * e0y;
    float n = 0.0f;
    n += cx * cx; n += cy * cy; n += cz * cz;       // square_norm = dot(v, v) (vector.h:745-753)
    return 0.25f * n == 0.0f;
}

// ---------------------------------------------------------------------------------------------
// BVH node, identical in layout to bvh::Node<AABB3d> (bvh_node.h:100-107): 6 x u32 header,
// 8 bytes of padding, then 12 doubles [minL minR maxL maxR] x (x, y, z) (bvh_node.h:141-162);
// a leaf reuses the 96-byte box area as user data.
// ---------------------------------------------------------------------------------------------

struct alignas(64) Node
{
    uint32_t    item_count;         // ~0 = interior
    uint32_t    index;              // first child (interior) or first item (leaf)
    uint32_t    left_bbox_index;
    uint32_t    left_bbox_count;
    uint32_t    right_bbox_index;
    uint32_t    right_bbox_count;
    uint32_t    pad[2];
    double      bbox[12];
};
static_assert(sizeof(Node) == 128, "node layout");

inline bool is_interior(const Node& n) { return n.item_count == ~uint32_t(0); }

template <typename T>
inline void set_child_bbox(Node& n, const int side, const Box<T>& b)
{
    for (int i = 0; i < 3; ++i)
    {
        n.bbox[i * 4 + side]     = static_cast<double>(b.mn[i]);
        n.bbox[i * 4 + 2 + side] = static_cast<double>(b.mx[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// Sweep-SAH partitioner + recursive builder
// (bvh_partitionerbase.h:95-198, bvh_bboxsortpredicate.h:113-126, bvh_sahpartitioner.h:99-170,
//  bvh_builder.h:113-229).  T is float for triangle trees (vector<GAABB3>, triangletree.cpp:543)
// and double for the assembly tree (assemblytree.cpp:176-186).
// ---------------------------------------------------------------------------------------------

template <typename T>
struct SahBuilder
{
    const std::vector<Box<T>>&  boxes;
    const size_t                max_leaf_size;
    const T                     traversal_cost;
    const T                     intersection_cost;
    std::vector<size_t>         indices[3];
    std::vector<size_t>         tmp;
    std::vector<uint8_t>        tags;
    std::vector<T>              left_areas;
    std::vector<Node>&          nodes;

    SahBuilder(const std::vector<Box<T>>& b, size_t mls, T ct, T ci, std::vector<Node>& out)
      : boxes(b), max_leaf_size(mls), traversal_cost(ct), intersection_cost(ci), nodes(out)
    {
        const size_t size = boxes.size();
        for (int d = 0; d < 3; ++d)
        {
            indices[d].resize(size);
            for (size_t i = 0; i < size; ++i) indices[d][i] = i;
            // Unstable std::sort on (min + max) with a strict '<': tie order is whatever this
            // libstdc++ produces -- the same library the reference would be built with here.
            const std::vector<Box<T>>& bb = boxes;
            std::sort(indices[d].begin(), indices[d].end(), [&bb, d](const size_t lhs, const size_t rhs)
            {
                const T lc = bb[lhs].mn[d] + bb[lhs].mx[d];
                const T rc = bb[rhs].mn[d] + bb[rhs].mx[d];
                return lc < rc;
            });
        }
        tmp.resize(size);
        tags.resize(size);
        left_areas.resize(size > 1 ? size - 1 : 0);
    }

    Box<T> compute_bbox(const size_t begin, const size_t end) const
    {
        Box<T> b; b.invalidate();
        for (size_t i = begin; i < end; ++i) b.insert(boxes[indices[0][i]]);
        return b;
    }

    void sort_indices(const int dimension, const size_t begin, const size_t end, const size_t pivot)
    {
        const std::vector<size_t>& split = indices[dimension];
        for (size_t i = begin; i < pivot; ++i) tags[split[i]] = 0;
        for (size_t i = pivot; i < end; ++i) tags[split[i]] = 1;

        for (int d = 0; d < 3; ++d)
        {
            if (d == dimension) continue;
            std::vector<size_t>& ind = indices[d];
            size_t left = begin, right = pivot;
            for (size_t i = begin; i < end; ++i)
            {
                const size_t index = ind[i];
                if (tags[index] == 0) tmp[left++] = index;
                else tmp[right++] = index;
            }
            // The reference swaps whole vectors for large ranges; the visible result is the same.
            for (size_t i = begin; i < end; ++i) ind[i] = tmp[i];
        }
    }

    size_t partition(const size_t begin, const size_t end, const Box<T>& bbox)
    {
        if (bbox.rank() < 2) return end;

        const size_t count = end - begin;
        if (count <= max_leaf_size) return end;

        T best_cost = std::numeric_limits<T>::max();
        int best_dim = 0;
        size_t best_pivot = 0;

        for (int d = 0; d < 3; ++d)
        {
            const std::vector<size_t>& ind = indices[d];
            Box<T> acc;

            acc.invalidate();
            for (size_t i = 0; i < count - 1; ++i)
            {
                acc.insert(boxes[ind[begin + i]]);
                left_areas[i] = half_area(acc);
            }

            acc.invalidate();
            for (size_t i = count - 1; i > 0; --i)
            {
                acc.insert(boxes[ind[begin + i]]);
                const T left_cost = left_areas[i - 1] * i;          // size_t -> T conversion
                const T right_cost = half_area(acc) * (count - i);
                const T split_cost = left_cost + right_cost;
                if (best_cost > split_cost)
                {
                    best_cost = split_cost;
                    best_dim = d;
                    best_pivot = i;
                }
            }
        }

        const T split_cost = traversal_cost + best_cost / half_area(bbox) * intersection_cost;
        const T leaf_cost = count * intersection_cost;
        if (leaf_cost <= split_cost) return end;

        const size_t pivot = begin + best_pivot;
        sort_indices(best_dim, begin, end, pivot);
        return pivot;
    }

    // Builder::subdivide_recurse (bvh_builder.h:163-229).  The box handed to partition() is the
    // node's (double) box converted back to T.
    void subdivide(const size_t node_index, const size_t begin, const size_t end, const Box<double>& bbox)
    {
        size_t pivot = end;
        if (end - begin > 1)
        {
            Box<T> pb;
            for (int i = 0; i < 3; ++i) { pb.mn[i] = static_cast<T>(bbox.mn[i]); pb.mx[i] = static_cast<T>(bbox.mx[i]); }
            pivot = partition(begin, end, pb);
        }

        if (pivot == end)
        {
            Node& node = nodes[node_index];
            if (node.item_count == ~uint32_t(0)) node.item_count = 0;
            node.index = static_cast<uint32_t>(begin);
            node.item_count = static_cast<uint32_t>(end - begin);
        }
        else
        {
            const Box<T> lb = compute_bbox(begin, pivot);
            const Box<T> rb = compute_bbox(pivot, end);
            Box<double> lbd, rbd;
            for (int i = 0; i < 3; ++i)
            {
                lbd.mn[i] = static_cast<double>(lb.mn[i]); lbd.mx[i] = static_cast<double>(lb.mx[i]);
                rbd.mn[i] = static_cast<double>(rb.mn[i]); rbd.mx[i] = static_cast<double>(rb.mx[i]);
            }

            const size_t left_index = nodes.size();
            {
                Node& node = nodes[node_index];
                node.item_count = ~uint32_t(0);
                set_child_bbox(node, 0, lb);
                set_child_bbox(node, 1, rb);
                node.index = static_cast<uint32_t>(left_index);
            }
            Node blank; std::memset(&blank, 0, sizeof(blank));
            nodes.push_back(blank);
            nodes.push_back(blank);

            subdivide(left_index, begin, pivot, lbd);
            subdivide(left_index + 1, pivot, end, rbd);
        }
    }

    void build()
    {
        nodes.clear();
        Node blank; std::memset(&blank, 0, sizeof(blank));
        nodes.push_back(blank);
        const Box<T> root = compute_bbox(0, boxes.size());
        Box<double> rd;
        for (int i = 0; i < 3; ++i) { rd.mn[i] = static_cast<double>(root.mn[i]); rd.mx[i] = static_cast<double>(root.mx[i]); }
        subdivide(0, 0, boxes.size(), rd);
    }
};

// ---------------------------------------------------------------------------------------------
// TriangleTree build (renderer/kernel/intersection/triangletree.cpp).
// ---------------------------------------------------------------------------------------------

struct Key { uint32_t object_instance_index; uint16_t pa; uint16_t hole; uint32_t triangle_index; };
static_assert(sizeof(Key) == 12, "TriangleKey layout (trianglekey.h:63-65)");

struct VertexInfo { size_t vertex_index; size_t msc; uint32_t vis; };     // trianglevertexinfo.h:46-48

inline V3f mesh_vertex(const orc_mesh& m, const size_t i)
{
    const V3f v = { m.vertices[i * 3], m.vertices[i * 3 + 1], m.vertices[i * 3 + 2] };
    return v;
}

inline V3f mesh_pose(const orc_mesh& m, const size_t v, const size_t seg)
{
    const float* p = m.vertex_poses + (v * m.motion_segment_count + seg) * 3;
    const V3f r = { p[0], p[1], p[2] };
    return r;
}

// lerp(a, b, k) = (1 - k) * a + k * b (scalar.h:934-938), per component in float.
inline V3f lerp_v(const V3f& a, const V3f& b, const float k)
{
    const float w = 1.0f - k;
    const V3f r = { w * a.x + k * b.x, w * a.y + k * b.y, w * a.z + k * b.z };
    return r;
}

// foundation::intersect(bbox, v0, v1, v2) (intersection/aabbtriangle.h:290-...): only the two
// first stages are restated.  The tree box handed in is the union of every object instance's
// parent box (assemblytree.cpp:401-405), so a triangle of this assembly always has a vertex
// inside it; `undecided` counts triangles that would need the edge/face stages (kept, flagged).
inline bool box_overlaps_triangle(const Box3f& b, const V3f& v0, const V3f& v1, const V3f& v2, size_t& undecided)
{
    const V3f* vs[3] = { &v0, &v1, &v2 };
    uint8_t masks[3];
    for (int k = 0; k < 3; ++k)
    {
        const float p[3] = { vs[k]->x, vs[k]->y, vs[k]->z };
        uint8_t m = 0;
        for (int i = 0; i < 3; ++i)
        {
            if (p[i] >= b.mn[i]) m |= uint8_t(1u << i);
            if (p[i] <= b.mx[i]) m |= uint8_t(1u << (3 + i));
        }
        if (m == 0x3F) return true;
        masks[k] = m;
    }
    if ((masks[0] | masks[1] | masks[2]) != 0x3F) return false;
    ++undecided;
    return true;
}

// IntersectionFilter (intersectionfilter.h:84-128, 169-205).
struct AlphaMask
{
    std::vector<uint8_t>    bits;
    uint32_t                width = 0, height = 0;
    bool present() const { return width != 0 && height != 0; }
    bool is_opaque(const float ux, const float uy) const
    {
        const float max_x = static_cast<float>(width) - 1.0f, max_y = static_cast<float>(height) - 1.0f;
        float fx = ux * static_cast<float>(width), fy = uy * static_cast<float>(height);
        fx = fx < 0.0f ? 0.0f : fx > max_x ? max_x : fx;
        fy = fy < 0.0f ? 0.0f : fy > max_y ? max_y : fy;
        const size_t ix = static_cast<size_t>(fx), iy = static_cast<size_t>(fy);
        return (bits[iy * ((width + 7) / 8) + ix / 8] & (1u << (ix & 7))) != 0;
    }
};

struct Filter
{
    AlphaMask               object_mask;
    std::vector<AlphaMask>  material_masks;
    std::vector<float>      uv;

    bool accept(const uint32_t triangle_index, const uint32_t pa, const double u, const double v) const
    {
        if (u != u || v != v) return true;
        const AlphaMask* mtl = pa < material_masks.size() && material_masks[pa].present() ? &material_masks[pa] : nullptr;
        if (object_mask.present() || mtl)
        {
            const float fu = static_cast<float>(u), fv = static_cast<float>(v);
            const float w = 1.0f - fu - fv;
            const float* t = uv.data() + size_t(triangle_index) * 6;
            float ux = t[0] * w, uy = t[1] * w;
            ux += t[2] * fu; uy += t[3] * fu;
            ux += t[4] * fv; uy += t[5] * fv;
            if (object_mask.present() && !object_mask.is_opaque(ux, uy)) return false;
            if (mtl) return mtl->is_opaque(ux, uy);
        }
        return true;
    }
};

struct TriTree
{
    std::vector<std::unique_ptr<Filter>> filters;   // per object instance, or empty
    std::vector<Node>       nodes;
    std::vector<double>     node_bboxes;        // 6 doubles each: minx maxx miny maxy minz maxz
    std::vector<uint8_t>    leaf_data;
    std::vector<Key>        keys;
    std::vector<float>      slot_tri;           // 9 floats per leaf slot: TriangleMT<float> of a static triangle (zeros if moving)
    std::vector<uint32_t>   slot_msc;           // per leaf slot: motion segment count of the triangle
    std::vector<size_t>     slot_pose;          // per leaf slot: index of the first vertex of pose 0 in slot_poses (moving triangles)
    std::vector<V3f>        slot_poses;         // (msc + 1) * 3 assembly-space vertices per moving triangle, as the leaf stores them
    size_t                  static_count = 0;
    size_t                  moving_count = 0;
    size_t                  undecided = 0;
};

struct TriCollect
{
    std::vector<Key>        keys;
    std::vector<VertexInfo> infos;
    std::vector<V3f>        vertices;
    std::vector<Box3f>      bboxes;
};

// collect_static_triangles / collect_moving_triangles / collect_triangles (triangletree.cpp:105-383).
// The reference runs two passes (boxes first, vertices later); one pass yields the same arrays.
void collect_triangles(const orc_scene_desc& desc, const orc_assembly& assembly, const Box3f& tree_bbox,
                       const double time, TriCollect& out, size_t& undecided)
{
    size_t vertex_count = 0;
    for (size_t oi = 0; oi < assembly.object_instance_count; ++oi)
    {
        const orc_object_instance& inst = assembly.object_instances[oi];
        const orc_mesh& mesh = desc.meshes[inst.mesh_index];
        const double* m = inst.local_to_parent;
        const size_t msc = mesh.motion_segment_count;
        std::vector<Box3f> pose_boxes(msc + 1);

        for (size_t i = 0; i < mesh.triangle_count; ++i)
        {
            const size_t i0 = mesh.triangles[i * 3], i1 = mesh.triangles[i * 3 + 1], i2 = mesh.triangles[i * 3 + 2];
            const V3f v0_os = mesh_vertex(mesh, i0), v1_os = mesh_vertex(mesh, i1), v2_os = mesh_vertex(mesh, i2);
            Box3f tri_box;

            if (msc == 0)
            {
                if (zero_area(v0_os, v1_os, v2_os)) continue;
                const V3f v0 = xform_point_f(m, v0_os), v1 = xform_point_f(m, v1_os), v2 = xform_point_f(m, v2_os);
                if (zero_area(v0, v1, v2)) continue;
                tri_box.invalidate();
                insert_v(tri_box, v0); insert_v(tri_box, v1); insert_v(tri_box, v2);
                if (!box_overlaps_triangle(tree_bbox, v0, v1, v2, undecided)) continue;
                out.vertices.push_back(v0); out.vertices.push_back(v1); out.vertices.push_back(v2);
            }
            else
            {
                const V3f v0 = xform_point_f(m, v0_os), v1 = xform_point_f(m, v1_os), v2 = xform_point_f(m, v2_os);
                pose_boxes[0].invalidate();
                insert_v(pose_boxes[0], v0); insert_v(pose_boxes[0], v1); insert_v(pose_boxes[0], v2);
                for (size_t s = 0; s < msc; ++s)
                {
                    pose_boxes[s + 1].invalidate();
                    insert_v(pose_boxes[s + 1], xform_point_f(m, mesh_pose(mesh, i0, s)));
                    insert_v(pose_boxes[s + 1], xform_point_f(m, mesh_pose(mesh, i1, s)));
                    insert_v(pose_boxes[s + 1], xform_point_f(m, mesh_pose(mesh, i2, s)));
                }
                Box3f motion_box = pose_boxes[0];
                for (size_t s = 1; s <= msc; ++s) motion_box.insert(pose_boxes[s]);
                if (motion_box.rank() < 2) continue;
                // AABB::overlap (aabb.h): boxes overlap unless separated along one axis.
                bool overlap = true;
                for (int a = 0; a < 3; ++a)
                    if (tree_bbox.mn[a] > motion_box.mx[a] || tree_bbox.mx[a] < motion_box.mn[a]) overlap = false;
                if (!overlap) continue;
                // interpolate<GAABB3> (renderer/utility/bbox.h:110-125).
                const size_t prev = static_cast<size_t>(time * msc);
                const float k = static_cast<float>(time * msc - prev);
                const float w = 1.0f - k;
                for (int a = 0; a < 3; ++a)
                {
                    tri_box.mn[a] = w * pose_boxes[prev].mn[a] + k * pose_boxes[prev + 1].mn[a];
                    tri_box.mx[a] = w * pose_boxes[prev].mx[a] + k * pose_boxes[prev + 1].mx[a];
                }
                if (tri_box.rank() < 2) continue;
                out.vertices.push_back(v0); out.vertices.push_back(v1); out.vertices.push_back(v2);
                for (size_t s = 0; s < msc; ++s)
                {
                    out.vertices.push_back(xform_point_f(m, mesh_pose(mesh, i0, s)));
                    out.vertices.push_back(xform_point_f(m, mesh_pose(mesh, i1, s)));
                    out.vertices.push_back(xform_point_f(m, mesh_pose(mesh, i2, s)));
                }
            }

            Key key; std::memset(&key, 0, sizeof(key));
            key.object_instance_index = static_cast<uint32_t>(oi);
            key.triangle_index = static_cast<uint32_t>(i);
            key.pa = mesh.triangle_pa ? mesh.triangle_pa[i] : 0;
            out.keys.push_back(key);
            const VertexInfo info = { vertex_count, msc, inst.vis_flags };
            out.infos.push_back(info);
            out.bboxes.push_back(tri_box);
            vertex_count += (msc + 1) * 3;
        }
    }
}

inline void push_swizzled(std::vector<double>& dst, const Box3f& b)     // triangletree.cpp:725-738
{
    for (int a = 0; a < 3; ++a)
    {
        dst.push_back(static_cast<double>(b.mn[a]));
        dst.push_back(static_cast<double>(b.mx[a]));
    }
}

// compute_motion_bboxes (triangletree.cpp:755-876).
std::vector<Box3f> compute_motion_bboxes(TriTree& tree, const std::vector<size_t>& order, const TriCollect& c, const size_t node_index)
{
    if (is_interior(tree.nodes[node_index]))
    {
        const size_t child = tree.nodes[node_index].index;
        const std::vector<Box3f> left = compute_motion_bboxes(tree, order, c, child);
        const std::vector<Box3f> right = compute_motion_bboxes(tree, order, c, child + 1);

        Node& node = tree.nodes[node_index];
        node.left_bbox_count = static_cast<uint32_t>(left.size());
        node.right_bbox_count = static_cast<uint32_t>(right.size());
        if (left.size() > 1)
        {
            node.left_bbox_index = static_cast<uint32_t>(tree.node_bboxes.size() / 6);
            for (const Box3f& b : left) push_swizzled(tree.node_bboxes, b);
        }
        if (right.size() > 1)
        {
            node.right_bbox_index = static_cast<uint32_t>(tree.node_bboxes.size() / 6);
            for (const Box3f& b : right) push_swizzled(tree.node_bboxes, b);
        }

        const size_t count = std::max(left.size(), right.size());
        std::vector<Box3f> boxes(count);
        for (size_t i = 0; i < count; ++i)
        {
            boxes[i] = left[i * left.size() / count];
            boxes[i].insert(right[i * right.size() / count]);
        }
        return boxes;
    }

    const Node& node = tree.nodes[node_index];
    const size_t item_begin = node.index, item_count = node.item_count;
    size_t max_msc = 0;
    Box3f base; base.invalidate();
    for (size_t i = 0; i < item_count; ++i)
    {
        const VertexInfo& info = c.infos[order[item_begin + i]];
        if (max_msc < info.msc) max_msc = info.msc;
        insert_v(base, c.vertices[info.vertex_index + 0]);
        insert_v(base, c.vertices[info.vertex_index + 1]);
        insert_v(base, c.vertices[info.vertex_index + 2]);
    }

    std::vector<Box3f> boxes(max_msc + 1);
    boxes[0] = base;
    if (max_msc > 0)
    {
        for (size_t m = 0; m < max_msc - 1; ++m)
        {
            boxes[m + 1].invalidate();
            const double time = static_cast<double>(m + 1) / max_msc;
            for (size_t i = 0; i < item_count; ++i)
            {
                const VertexInfo& info = c.infos[order[item_begin + i]];
                const size_t prev = static_cast<size_t>(time * info.msc);
                const size_t bv = info.vertex_index + prev * 3;
                const float k = static_cast<float>(time * info.msc - prev);
                insert_v(boxes[m + 1], lerp_v(c.vertices[bv + 0], c.vertices[bv + 3], k));
                insert_v(boxes[m + 1], lerp_v(c.vertices[bv + 1], c.vertices[bv + 4], k));
                insert_v(boxes[m + 1], lerp_v(c.vertices[bv + 2], c.vertices[bv + 5], k));
            }
        }
        boxes[max_msc].invalidate();
        for (size_t i = 0; i < item_count; ++i)
        {
            const VertexInfo& info = c.infos[order[item_begin + i]];
            const size_t bv = info.vertex_index + info.msc * 3;
            insert_v(boxes[max_msc], c.vertices[bv + 0]);
            insert_v(boxes[max_msc], c.vertices[bv + 1]);
            insert_v(boxes[max_msc], c.vertices[bv + 2]);
        }
    }
    return boxes;
}

// TriangleEncoder (triangleencoder.cpp:48-103).
size_t encoded_size(const TriCollect& c, const std::vector<size_t>& order, size_t begin, size_t count)
{
    size_t size = 0;
    for (size_t i = 0; i < count; ++i)
    {
        const VertexInfo& info = c.infos[order[begin + i]];
        size += 8;
        size += info.msc == 0 ? 36 : (info.msc + 1) * 36;
    }
    return size;
}

uint8_t* encode(const TriCollect& c, const std::vector<size_t>& order, size_t begin, size_t count, uint8_t* out)
{
    for (size_t i = 0; i < count; ++i)
    {
        const VertexInfo& info = c.infos[order[begin + i]];
        const uint32_t vis = info.vis, msc = static_cast<uint32_t>(info.msc);
        std::memcpy(out, &vis, 4); out += 4;
        std::memcpy(out, &msc, 4); out += 4;
        if (msc == 0)
        {
            // TriangleMT<float>(v0, v1, v2): edges computed in float (raytrianglemt.h:128-137).
            const V3f& v0 = c.vertices[info.vertex_index], &v1 = c.vertices[info.vertex_index + 1], &v2 = c.vertices[info.vertex_index + 2];
            const float t[9] = { v0.x, v0.y, v0.z, v1.x - v0.x, v1.y - v0.y, v1.z - v0.z, v2.x - v0.x, v2.y - v0.y, v2.z - v0.z };
            std::memcpy(out, t, 36); out += 36;
        }
        else
        {
            const size_t bytes = (msc + 1) * 36;
            std::memcpy(out, &c.vertices[info.vertex_index], bytes); out += bytes;
        }
    }
    return out;
}

// store_triangles (triangletree.cpp:878-978).
void store_triangles(TriTree& tree, const std::vector<size_t>& order, const TriCollect& c)
{
    size_t spill = 0;
    for (const Node& node : tree.nodes)
        if (!is_interior(node))
        {
            const size_t s = encoded_size(c, order, node.index, node.item_count);
            if (s > 92) spill += s;
        }
    tree.keys.reserve(order.size());
    tree.leaf_data.resize(spill);
    uint8_t* writer = tree.leaf_data.empty() ? nullptr : tree.leaf_data.data();

    for (Node& node : tree.nodes)
    {
        if (is_interior(node)) continue;
        const size_t begin = node.index, count = node.item_count;
        node.index = static_cast<uint32_t>(tree.keys.size());
        for (size_t j = 0; j < count; ++j)
        {
            tree.keys.push_back(c.keys[order[begin + j]]);
            const VertexInfo& info = c.infos[order[begin + j]];
            float t[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
            if (info.msc == 0)
            {
                const V3f& v0 = c.vertices[info.vertex_index], &v1 = c.vertices[info.vertex_index + 1], &v2 = c.vertices[info.vertex_index + 2];
                const float u[9] = { v0.x, v0.y, v0.z, v1.x - v0.x, v1.y - v0.y, v1.z - v0.z, v2.x - v0.x, v2.y - v0.y, v2.z - v0.z };
                std::memcpy(t, u, 36);
            }
            tree.slot_tri.insert(tree.slot_tri.end(), t, t + 9);
            tree.slot_msc.push_back(static_cast<uint32_t>(info.msc));
            tree.slot_pose.push_back(tree.slot_poses.size());
            if (info.msc != 0)
                tree.slot_poses.insert(tree.slot_poses.end(), c.vertices.begin() + info.vertex_index, c.vertices.begin() + info.vertex_index + (info.msc + 1) * 3);
        }
        const size_t s = encoded_size(c, order, begin, count);
        uint8_t* user = reinterpret_cast<uint8_t*>(node.bbox);
        if (s <= 92)
        {
            const uint32_t marker = ~uint32_t(0);
            std::memcpy(user, &marker, 4);
            encode(c, order, begin, count, user + 4);
        }
        else
        {
            const uint32_t offset = static_cast<uint32_t>(writer - tree.leaf_data.data());
            std::memcpy(user, &offset, 4);
            writer = encode(c, order, begin, count, writer);
        }
    }
}

// TriangleTree::build_bvh (triangletree.cpp:497-597).
void build_triangle_tree(const orc_scene_desc& desc, const orc_assembly& assembly, const Box3f& bbox, TriTree& tree)
{
    TriCollect c;
    collect_triangles(desc, assembly, bbox, assembly.time, c, tree.undecided);
    for (const VertexInfo& info : c.infos)
        if (info.msc == 0) ++tree.static_count;
    tree.moving_count = c.infos.size() - tree.static_count;

    SahBuilder<float> builder(c.bboxes, assembly.max_leaf_size, assembly.interior_node_traversal_cost,
                              assembly.triangle_intersection_cost, tree.nodes);
    builder.build();
    compute_motion_bboxes(tree, builder.indices[0], c, 0);
    store_triangles(tree, builder.indices[0], c);
}

// ---------------------------------------------------------------------------------------------
// Rays and primitive tests.
// ---------------------------------------------------------------------------------------------

struct Ray
{
    V3d         org, dir;
    double      tmin, tmax;
    float       time_absolute, time_normalized;
    uint32_t    flags;
};

// RayInfo (ray.h:313-321): rcp_dir = 1 / dir, sgn = rcp_dir >= 0.
struct RayInfo
{
    double  rcp[3];
    int     sgn[3];
    explicit RayInfo(const Ray& r)
    {
        const double d[3] = { r.dir.x, r.dir.y, r.dir.z };
        for (int i = 0; i < 3; ++i)
        {
            rcp[i] = 1.0 / d[i];
            sgn[i] = rcp[i] >= 0.0 ? 1 : 0;
        }
    }
};

// minmax.h:171-180 -- the same selection _mm_min_pd / _mm_max_pd make (second operand on NaN).
inline double ssemin(const double a, const double b) { return a < b ? a : b; }
inline double ssemax(const double a, const double b) { return a > b ? a : b; }

// Slab test of both children of an interior node (bvh_intersector.h:519-536; also the generic
// 4-argument test rayaabb.h:228-252).  box[] is [minL minR maxL maxR] per axis.
// Returns bit 0 = left hit, bit 1 = right hit; tmin_out[] = clamped entry distances.
inline int slab2(const double* box, const double org[3], const RayInfo& info, const double ray_tmin, const double ray_tmax, double tmin_out[2])
{
    int hits = 0;
    for (int side = 0; side < 2; ++side)
    {
        double l1[3], l2[3];
        for (int a = 0; a < 3; ++a)
        {
            const double near_plane = box[a * 4 + 2 * (1 - info.sgn[a]) + side];
            const double far_plane  = box[a * 4 + 2 * (    info.sgn[a]) + side];
            l1[a] = info.rcp[a] * (near_plane - org[a]);
            l2[a] = info.rcp[a] * (far_plane - org[a]);
        }
        const double tmin = ssemax(l1[2], ssemax(l1[1], ssemax(l1[0], ray_tmin)));
        const double tmax = ssemin(l2[2], ssemin(l2[1], ssemin(l2[0], ray_tmax)));
        tmin_out[side] = tmin;
        if (!(tmin > tmax || tmax < ray_tmin || tmin >= ray_tmax))
            hits |= 1 << side;
    }
    return hits;
}

// TriangleMT<double>::intersect (raytrianglemt.h:148-212) on a stored float triangle widened to
// double (raytrianglemt.h:139-146).  cross: vector.h:1239-1246; dot accumulates from 0 left to
// right: vector.h:745-753.
struct MTd { double v0[3], e0[3], e1[3]; };

inline void cross3(const double a[3], const double b[3], double r[3])
{
    r[0] = a[1] * b[2] - b[1] * a[2];
    r[1] = a[2] * b[0] - b[2] * a[0];
    r[2] = a[0] * b[1] - b[0] * a[1];
}

inline double dot3(const double a[3], const double b[3])
{
    double r = 0.0;
    r += a[0] * b[0]; r += a[1] * b[1]; r += a[2] * b[2];
    return r;
}

inline bool mt_intersect(const MTd& tri, const Ray& ray, double& t, double& u, double& v)
{
    const double dir[3] = { ray.dir.x, ray.dir.y, ray.dir.z };
    double pvec[3]; cross3(dir, tri.e1, pvec);
    const double det = dot3(tri.e0, pvec);
    const double tvec[3] = { ray.org.x - tri.v0[0], ray.org.y - tri.v0[1], ray.org.z - tri.v0[2] };
    double qvec[3];
    if (det > 0.0)
    {
        u = dot3(tvec, pvec);
        if (u < 0.0 || u > det) return false;
        cross3(tvec, tri.e0, qvec);
        v = dot3(dir, qvec);
        if (v < 0.0 || u + v > det) return false;
        t = dot3(tri.e1, qvec);
        if (t >= ray.tmax * det || t < ray.tmin * det) return false;
    }
    else
    {
        u = dot3(tvec, pvec);
        if (u > 0.0 || u < det) return false;
        cross3(tvec, tri.e0, qvec);
        v = dot3(dir, qvec);
        if (v > 0.0 || u + v < det) return false;
        t = dot3(tri.e1, qvec);
        if (t <= ray.tmax * det || t > ray.tmin * det) return false;
    }
    const double rcp_det = 1.0 / det;
    t *= rcp_det; u *= rcp_det; v *= rcp_det;
    return true;
}

// Boolean variant (raytrianglemt.h:214-268): identical inequalities, no division.
inline bool mt_intersect_bool(const MTd& tri, const Ray& ray)
{
    const double dir[3] = { ray.dir.x, ray.dir.y, ray.dir.z };
    double pvec[3]; cross3(dir, tri.e1, pvec);
    const double det = dot3(tri.e0, pvec);
    const double tvec[3] = { ray.org.x - tri.v0[0], ray.org.y - tri.v0[1], ray.org.z - tri.v0[2] };
    double qvec[3];
    if (det > 0.0)
    {
        const double u = dot3(tvec, pvec);
        if (u < 0.0 || u > det) return false;
        cross3(tvec, tri.e0, qvec);
        const double v = dot3(dir, qvec);
        if (v < 0.0 || u + v > det) return false;
        const double t = dot3(tri.e1, qvec);
        if (t >= ray.tmax * det || t < ray.tmin * det) return false;
    }
    else
    {
        const double u = dot3(tvec, pvec);
        if (u > 0.0 || u < det) return false;
        cross3(tvec, tri.e0, qvec);
        const double v = dot3(dir, qvec);
        if (v > 0.0 || u + v < det) return false;
        const double t = dot3(tri.e1, qvec);
        if (t <= ray.tmax * det || t > ray.tmin * det) return false;
    }
    return true;
}

// Reads one encoded triangle (triangleencoder.cpp:72-103) at `p` for a ray time and returns the
// double-precision MT triangle.  `base_time` follows the caller's arithmetic (float product for
// closest hit, double product for probes -- triangletree.cpp:1433 vs :1570).
inline const uint8_t* read_triangle(const uint8_t* p, const uint32_t msc, const double base_time, MTd& tri, uint32_t& base_index_out)
{
    float f[9];
    if (msc == 0)
    {
        std::memcpy(f, p, 36);
        p += 36;
        base_index_out = 0;
    }
    else
    {
        const size_t base_index = static_cast<size_t>(base_time);
        const float frac = static_cast<float>(base_time - base_index);
        const float omf = 1.0f - frac;
        float a[9], b[9];
        std::memcpy(a, p + base_index * 36, 36);
        std::memcpy(b, p + (base_index + 1) * 36, 36);
        float v[9];
        for (int i = 0; i < 9; ++i) { v[i] = a[i] * omf; v[i] += b[i] * frac; }
        // TriangleMT<float>(v0, v1, v2) (raytrianglemt.h:128-137).
        f[0] = v[0]; f[1] = v[1]; f[2] = v[2];
        f[3] = v[3] - v[0]; f[4] = v[4] - v[1]; f[5] = v[5] - v[2];
        f[6] = v[6] - v[0]; f[7] = v[7] - v[1]; f[8] = v[8] - v[2];
        p += (msc + 1) * 36;
        base_index_out = static_cast<uint32_t>(base_index);
    }
    for (int i = 0; i < 3; ++i)
    {
        tri.v0[i] = static_cast<double>(f[i]);
        tri.e0[i] = static_cast<double>(f[3 + i]);
        tri.e1[i] = static_cast<double>(f[6 + i]);
    }
    return p;
}

inline const uint8_t* leaf_payload(const TriTree& tree, const Node& node)
{
    const uint8_t* user = reinterpret_cast<const uint8_t*>(node.bbox);
    uint32_t offset;
    std::memcpy(&offset, user, 4);
    return offset == ~uint32_t(0) ? user + 4 : tree.leaf_data.data() + offset;
}

// ---------------------------------------------------------------------------------------------
// Bottom level: bvh::Intersector<TriangleTree, Visitor, Ray3d, 64, 3> (bvh_intersector.h:472-616
// static, :623-890 motion) with TriangleLeafVisitor / TriangleLeafProbeVisitor
// (triangletree.cpp:1352-1499, 1506-1603) inlined at the leaf.
// ---------------------------------------------------------------------------------------------

enum Mode { ClosestHit, AnyHit, TwoNearest };

struct LocalHit
{
    bool        hit = false;
    float       u = 0.0f, v = 0.0f;
    uint32_t    slot = 0, motion_segment = 0;
};

struct TwoBest
{
    double t1 = DBL_INF, t2 = DBL_INF;
    void add(const double t)
    {
        if (t < t1) { t2 = t1; t1 = t; }
        else if (t < t2) t2 = t;
    }
};

// Child boxes of an interior node at ray_time (bvh_intersector.h:675-836): a child whose
// bbox_count > 1 is lerp(m_node_bboxes[i + prev], m_node_bboxes[i + prev + 1]) computed as
// a * w1 + b * w2 with w2 = t - trunc(t), w1 = 1 - w2.
inline void motion_boxes(const TriTree& tree, const Node& node, const double ray_time, double out[12])
{
    for (int side = 0; side < 2; ++side)
    {
        const uint32_t count = side == 0 ? node.left_bbox_count : node.right_bbox_count;
        const uint32_t index = side == 0 ? node.left_bbox_index : node.right_bbox_index;
        const size_t segments = static_cast<size_t>(count) - 1;
        if (segments > 0)
        {
            const double t = ray_time * static_cast<double>(segments);
            const int prev = static_cast<int>(t);
            const double w2 = t - static_cast<double>(prev);
            const double w1 = 1.0 - w2;
            const double* b = tree.node_bboxes.data() + (static_cast<size_t>(index) + prev) * 6;
            for (int a = 0; a < 3; ++a)
            {
                out[a * 4 + side]     = b[a * 2 + 0] * w1 + b[6 + a * 2 + 0] * w2;
                out[a * 4 + 2 + side] = b[a * 2 + 1] * w1 + b[6 + a * 2 + 1] * w2;
            }
        }
        else
        {
            for (int a = 0; a < 3; ++a)
            {
                out[a * 4 + side]     = node.bbox[a * 4 + side];
                out[a * 4 + 2 + side] = node.bbox[a * 4 + 2 + side];
            }
        }
    }
}

template <Mode M>
bool traverse_triangle_tree(const TriTree& tree, Ray& ray, const bool motion, LocalHit& hit, TwoBest* two, orc_counters& cnt)
{
    const RayInfo info(ray);
    const double org[3] = { ray.org.x, ray.org.y, ray.org.z };
    const double ray_time = static_cast<double>(ray.time_normalized);

    const Node* stack[64];
    const Node** sp = stack;
    const Node* node = tree.nodes.data();
    double rtmax = ray.tmax;

    while (true)
    {
        ++cnt.triangle_nodes_visited;
        if (is_interior(*node))
        {
            double tmin[2];
            int hits;
            if (motion)
            {
                double boxes[12];
                motion_boxes(tree, *node, ray_time, boxes);
                hits = slab2(boxes, org, info, ray.tmin, rtmax, tmin);
            }
            else hits = slab2(node->bbox, org, info, ray.tmin, rtmax, tmin);

            const int hit_left = hits & 1, hit_right = hits >> 1;
            node = tree.nodes.data() + node->index;
            node += hit_right;

            if (hit_left ^ hit_right) continue;
            if (hits)
            {
                const int far_index = tmin[0] < tmin[1] ? 1 : 0;
                *sp++ = node + far_index - 1;
                node -= far_index;
                continue;
            }
            if (sp == stack) break;
            node = *--sp;
            continue;
        }

        // Leaf.
        const uint8_t* p = leaf_payload(tree, *node);
        uint32_t slot = node->index;
        for (uint32_t k = node->item_count; k--; ++slot)
        {
            ++cnt.triangles_tested;
            uint32_t vis, msc;
            std::memcpy(&vis, p, 4); std::memcpy(&msc, p + 4, 4); p += 8;
            if (!(vis & ray.flags))
            {
                p += msc == 0 ? 36 : (msc + 1) * 36;
                continue;
            }
            double base_time = 0.0;
            if (msc != 0)
            {
                if (M == ClosestHit || M == TwoNearest)
                    base_time = static_cast<double>(ray.time_normalized * static_cast<float>(msc));     // float product (:1433)
                else base_time = ray_time * static_cast<double>(msc);                                 // double product (:1570)
            }
            MTd tri; uint32_t base_index;
            p = read_triangle(p, msc, base_time, tri, base_index);

            if (M == ClosestHit)
            {
                double t, u, v;
                if (mt_intersect(tri, ray, t, u, v))
                {
                    if (!tree.filters.empty())
                    {
                        const Key& key = tree.keys[slot];
                        const Filter* f = key.object_instance_index < tree.filters.size() ? tree.filters[key.object_instance_index].get() : nullptr;
                        if (f && !f->accept(key.triangle_index, key.pa, u, v)) continue;
                    }
                    hit.hit = true;
                    hit.slot = slot;
                    hit.motion_segment = base_index;
                    hit.u = static_cast<float>(u);
                    hit.v = static_cast<float>(v);
                    ray.tmax = t;
                }
            }
            else if (M == AnyHit)
            {
                if (mt_intersect_bool(tri, ray)) return true;
            }
            else
            {
                double t, u, v;
                if (mt_intersect(tri, ray, t, u, v))
                {
                    const Key& key = tree.keys[slot];
                    const Filter* f = key.object_instance_index < tree.filters.size() ? tree.filters[key.object_instance_index].get() : nullptr;
                    if (!f || f->accept(key.triangle_index, key.pa, u, v)) two->add(t);
                }
            }
        }

        // distance = ray.tmax for both visitors (:1479, :1601); rtmax only ever shrinks.
        if (rtmax > ray.tmax) rtmax = ray.tmax;
        if (sp == stack) break;
        node = *--sp;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// Top level: AssemblyTree (assemblytree.cpp) traversed by the generic scalar intersector
// (bvh_intersector.h:139-258) with AssemblyLeafVisitor / AssemblyLeafProbeVisitor.
// ---------------------------------------------------------------------------------------------

struct Item
{
    uint32_t    assembly_instance;
    uint32_t    tree;
    uint32_t    vis_flags;
    double      parent_to_local[16];
};

struct Scene
{
    orc_scene_desc                          desc;
    std::vector<std::unique_ptr<TriTree>>   trees;
    std::vector<int>                        assembly_tree;
    std::vector<Node>                       nodes;
    std::vector<Item>                       items;
    std::vector<uint32_t>                   item_assembly_instance, item_tree;
};

void build_scene(Scene& s)
{
    const orc_scene_desc& desc = s.desc;

    // Assembly boxes: ObjectInstance::compute_parent_bbox (objectinstance.cpp:255-267) over
    // StaticTessellation::compute_local_bbox (statictessellation.h:462-479); one triangle tree
    // per assembly with object instances (assemblytree.cpp:372-420).
    s.assembly_tree.assign(desc.assembly_count, -1);
    std::vector<Box3f> assembly_boxes(desc.assembly_count);
    for (size_t a = 0; a < desc.assembly_count; ++a)
    {
        const orc_assembly& assembly = desc.assemblies[a];
        Box3f ab; ab.invalidate();
        for (size_t o = 0; o < assembly.object_instance_count; ++o)
        {
            const orc_object_instance& oi = assembly.object_instances[o];
            const orc_mesh& mesh = desc.meshes[oi.mesh_index];
            Box3f lb; lb.invalidate();
            for (size_t i = 0; i < mesh.vertex_count; ++i)
            {
                insert_v(lb, mesh_vertex(mesh, i));
                for (size_t j = 0; j < mesh.motion_segment_count; ++j) insert_v(lb, mesh_pose(mesh, i, j));
            }
            ab.insert(xform_box_f(oi.local_to_parent, lb));
        }
        assembly_boxes[a] = ab;
        if (assembly.object_instance_count > 0)
        {
            s.assembly_tree[a] = static_cast<int>(s.trees.size());
            s.trees.emplace_back(new TriTree());
            build_triangle_tree(desc, assembly, ab, *s.trees.back());
        }
    }

    // collect_assembly_instances (assemblytree.cpp:111-152).
    std::vector<Box3d> inst_boxes;
    std::vector<Item> items;
    for (size_t i = 0; i < desc.assembly_instance_count; ++i)
    {
        const orc_assembly_instance& inst = desc.assembly_instances[i];
        if (desc.assemblies[inst.assembly_index].object_instance_count == 0) continue;
        Item item;
        item.assembly_instance = static_cast<uint32_t>(i);
        item.tree = static_cast<uint32_t>(s.assembly_tree[inst.assembly_index]);
        item.vis_flags = inst.vis_flags;
        std::memcpy(item.parent_to_local, inst.parent_to_local, sizeof(item.parent_to_local));
        items.push_back(item);

        // TransformSequence::to_parent(GAABB3) on the float assembly box, then AABB3d, then
        // robust_grow(1e-15) (aabb.h:621-641).
        const Box3f wb = xform_box_f(inst.local_to_parent, assembly_boxes[inst.assembly_index]);
        Box3d b;
        for (int a = 0; a < 3; ++a) { b.mn[a] = static_cast<double>(wb.mn[a]); b.mx[a] = static_cast<double>(wb.mx[a]); }
        for (int a = 0; a < 3; ++a)
        {
            const double c = 0.5 * (b.mn[a] + b.mx[a]);
            const double e = b.mx[a] - b.mn[a];
            const double ac = c < 0.0 ? -c : c;
            double dominant = ac > e ? ac : e;
            if (!(dominant > 1.0)) dominant = 1.0;
            const double delta = dominant * 1.0e-15;
            b.mn[a] -= delta;
            b.mx[a] += delta;
        }
        inst_boxes.push_back(b);
    }

    // rebuild_assembly_tree (assemblytree.cpp:154-212): leaf size 1, costs (1, 10).
    SahBuilder<double> builder(inst_boxes, 1, 1.0, 10.0, s.nodes);
    builder.build();
    s.items.resize(items.size());
    for (size_t i = 0; i < items.size(); ++i) s.items[i] = items[builder.indices[0][i]];
    for (const Item& it : s.items)
    {
        s.item_assembly_instance.push_back(it.assembly_instance);
        s.item_tree.push_back(it.tree);
    }
}

// rayaabb.h:228-252 for one child (generic 4-argument test); box[] laid out as in the node.
inline bool slab1(const double* box, const int side, const Ray& ray, const RayInfo& info, double& tmin_out)
{
    const double org[3] = { ray.org.x, ray.org.y, ray.org.z };
    double l1[3], l2[3];
    for (int a = 0; a < 3; ++a)
    {
        l1[a] = info.rcp[a] * (box[a * 4 + 2 * (1 - info.sgn[a]) + side] - org[a]);
        l2[a] = info.rcp[a] * (box[a * 4 + 2 * (    info.sgn[a]) + side] - org[a]);
    }
    const double tmin = ssemax(l1[2], ssemax(l1[1], ssemax(l1[0], ray.tmin)));
    const double tmax = ssemin(l2[2], ssemin(l2[1], ssemin(l2[0], ray.tmax)));
    if (tmin > tmax || tmax < ray.tmin || tmin >= ray.tmax) return false;
    tmin_out = ssemax(ray.tmin, tmin);
    return true;
}

struct WorldHit
{
    bool        hit = false;
    float       u = 0.0f, v = 0.0f;
    uint32_t    assembly_instance = ~uint32_t(0), tree = 0, slot = 0, motion_segment = 0;
};

template <Mode M>
bool traverse_scene(const Scene& s, Ray& ray, WorldHit& wh, TwoBest* two, orc_counters& cnt, const orc_parent* parent = nullptr)
{
    const RayInfo info(ray);
    const Node* stack[64];
    const Node** sp = stack;
    const Node* node = s.nodes.data();
    double ray_tmax = ray.tmax;

    while (true)
    {
        ++cnt.assembly_nodes_visited;
        if (is_interior(*node))
        {
            double tmin[2] = { 0.0, 0.0 };
            // The 4-argument test reads the LIVE ray.tmax (which the closest-hit visitor shrinks)
            // and the loop additionally requires tmin < ray_tmax (bvh_intersector.h:177-186).
            const int hit_left = slab1(node->bbox, 0, ray, info, tmin[0]) && tmin[0] < ray_tmax ? 1 : 0;
            const int hit_right = slab1(node->bbox, 1, ray, info, tmin[1]) && tmin[1] < ray_tmax ? 1 : 0;
            node = s.nodes.data() + node->index;
            node += hit_right;
            if (hit_left ^ hit_right) continue;
            if (hit_left | hit_right)
            {
                const int far_index = tmin[0] < tmin[1] ? 1 : 0;
                *sp++ = node + far_index - 1;
                node -= far_index;
                continue;
            }
            if (sp == stack) break;
            node = *--sp;
            continue;
        }

        // AssemblyLeafVisitor::visit / AssemblyLeafProbeVisitor::visit (assemblytree.cpp:604-838, 845-1054).
        for (uint32_t i = 0; i < node->item_count; ++i)
        {
            const Item& item = s.items[node->index + i];
            if (!(item.vis_flags & ray.flags)) continue;
            ++cnt.instances_visited;

            // compute_assembly_instance_ray (assemblytree.cpp:556-596): inside the assembly instance
            // that holds the parent's hit the origin is the parent's offset point
            // (ShadingPoint::get_offset_point, shadingpoint.h:604-613), elsewhere the transformed origin.
            Ray local;
            local.dir = xform_vector_d(item.parent_to_local, ray.dir);
            if (parent && parent->assembly_instance == item.assembly_instance)
            {
                const double d = 0.0 + parent->geo_normal[0] * local.dir.x + parent->geo_normal[1] * local.dir.y + parent->geo_normal[2] * local.dir.z;
                const double* p = d > 0.0 ? parent->front : parent->back;
                local.org.x = p[0]; local.org.y = p[1]; local.org.z = p[2];
            }
            else local.org = xform_point_d(item.parent_to_local, ray.org);
            local.tmin = ray.tmin;
            local.tmax = ray.tmax;
            local.time_absolute = ray.time_absolute;
            local.time_normalized = ray.time_normalized;
            local.flags = ray.flags;

            if (item.tree == ~uint32_t(0)) continue;
            const TriTree& tree = *s.trees[item.tree];
            LocalHit lh;
            const bool found = traverse_triangle_tree<M>(tree, local, tree.moving_count > 0, lh, two, cnt);

            if (M == AnyHit)
            {
                if (found) return true;
            }
            else if (M == ClosestHit)
            {
                if (lh.hit && local.tmax < ray.tmax)
                {
                    ray.tmax = local.tmax;
                    wh.hit = true;
                    wh.u = lh.u; wh.v = lh.v;
                    wh.assembly_instance = item.assembly_instance;
                    wh.tree = item.tree;
                    wh.slot = lh.slot;
                    wh.motion_segment = lh.motion_segment;
                }
            }
        }

        // distance = m_shading_point.m_ray.m_tmax (closest hit) or ray.m_tmax (probe).
        if (ray_tmax > ray.tmax) ray_tmax = ray.tmax;
        if (sp == stack) break;
        node = *--sp;
    }
    return false;
}

inline void load_ray(const orc_rays& rays, const size_t i, Ray& r)
{
    r.org.x = rays.org[i * 3]; r.org.y = rays.org[i * 3 + 1]; r.org.z = rays.org[i * 3 + 2];
    r.dir.x = rays.dir[i * 3]; r.dir.y = rays.dir[i * 3 + 1]; r.dir.z = rays.dir[i * 3 + 2];
    r.tmin = rays.tmin[i];
    r.tmax = rays.tmax[i];
    r.time_absolute = rays.time_absolute ? rays.time_absolute[i] : 0.0f;
    r.time_normalized = rays.time_normalized ? rays.time_normalized[i] : 0.0f;
    r.flags = rays.flags ? rays.flags[i] : ~uint32_t(0);
}

template <typename F>
void parallel_ranges(const size_t n, int threads, F f)
{
    if (threads < 1) threads = 1;
    if (static_cast<size_t>(threads) > n) threads = n > 0 ? static_cast<int>(n) : 1;
    if (threads == 1) { f(0, size_t(0), n); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back(f, t, n * t / threads, n * (t + 1) / threads);
    for (std::thread& th : pool) th.join();
}

void accumulate(orc_counters* total, const std::vector<orc_counters>& parts)
{
    if (!total) return;
    std::memset(total, 0, sizeof(*total));
    for (const orc_counters& c : parts)
    {
        total->rays += c.rays;
        total->assembly_nodes_visited += c.assembly_nodes_visited;
        total->instances_visited += c.instances_visited;
        total->triangle_nodes_visited += c.triangle_nodes_visited;
        total->triangles_tested += c.triangles_tested;
        total->hits += c.hits;
    }
}

}   // anonymous namespace

extern "C" {

void* orc_scene_create(const orc_scene_desc* desc)
{
    Scene* s = new Scene();
    s->desc = *desc;
    build_scene(*s);
    return s;
}

void orc_scene_destroy(void* scene) { delete static_cast<Scene*>(scene); }

int orc_tree_count(const void* scene) { return static_cast<int>(static_cast<const Scene*>(scene)->trees.size()); }

int orc_assembly_tree_index(const void* scene, uint32_t assembly) { return static_cast<const Scene*>(scene)->assembly_tree[assembly]; }

void orc_get_triangle_tree(const void* scene, int tree, orc_triangle_tree_view* out)
{
    const TriTree& t = *static_cast<const Scene*>(scene)->trees[tree];
    out->nodes = t.nodes.data();
    out->node_count = t.nodes.size();
    out->node_bboxes = t.node_bboxes.empty() ? nullptr : t.node_bboxes.data();
    out->node_bbox_count = t.node_bboxes.size() / 6;
    out->leaf_data = t.leaf_data.empty() ? nullptr : t.leaf_data.data();
    out->leaf_data_size = t.leaf_data.size();
    out->triangle_keys = t.keys.empty() ? nullptr : t.keys.data();
    out->triangle_key_count = t.keys.size();
    out->static_triangle_count = t.static_count;
    out->moving_triangle_count = t.moving_count;
}

void orc_get_assembly_tree(const void* scene, orc_assembly_tree_view* out)
{
    const Scene& s = *static_cast<const Scene*>(scene);
    out->nodes = s.nodes.data();
    out->node_count = s.nodes.size();
    out->item_assembly_instance = s.item_assembly_instance.data();
    out->item_tree = s.item_tree.data();
    out->item_count = s.items.size();
}

// Intersector::trace (intersector.cpp:124-189), parent_shading_point == nullptr.
void orc_trace(const void* scene, const orc_rays* rays, size_t n, orc_hit* out, int threads, orc_counters* counters)
{
    const Scene& s = *static_cast<const Scene*>(scene);
    std::vector<orc_counters> parts(threads < 1 ? 1 : threads);
    std::memset(parts.data(), 0, parts.size() * sizeof(orc_counters));
    parallel_ranges(n, threads, [&](int tid, size_t begin, size_t end)
    {
        orc_counters local; std::memset(&local, 0, sizeof(local));
        for (size_t i = begin; i < end; ++i)
        {
            Ray ray; load_ray(*rays, i, ray);
            WorldHit wh;
            traverse_scene<ClosestHit>(s, ray, wh, nullptr, local);
            orc_hit& h = out[i];
            std::memset(&h, 0, sizeof(h));
            h.t = ray.tmax;
            h.assembly_instance = ~uint32_t(0);
            if (wh.hit)
            {
                // read_hit_triangle_data (triangletree.cpp:1483-1499).
                const Key& key = s.trees[wh.tree]->keys[wh.slot];
                h.u = wh.u; h.v = wh.v;
                h.assembly_instance = wh.assembly_instance;
                h.object_instance_index = key.object_instance_index;
                h.primitive_index = key.triangle_index;
                h.tri_slot = wh.slot;
                h.motion_segment = wh.motion_segment;
                h.prim_type = 2;
                ++local.hits;
            }
            ++local.rays;
        }
        parts[tid] = local;
    });
    accumulate(counters, parts);
}

// Intersector::trace_probe (intersector.cpp:191-238), parent_shading_point == nullptr.
void orc_trace_probe(const void* scene, const orc_rays* rays, size_t n, uint8_t* out, int threads, orc_counters* counters)
{
    const Scene& s = *static_cast<const Scene*>(scene);
    std::vector<orc_counters> parts(threads < 1 ? 1 : threads);
    std::memset(parts.data(), 0, parts.size() * sizeof(orc_counters));
    parallel_ranges(n, threads, [&](int tid, size_t begin, size_t end)
    {
        orc_counters local; std::memset(&local, 0, sizeof(local));
        for (size_t i = begin; i < end; ++i)
        {
            Ray ray; load_ray(*rays, i, ray);
            WorldHit wh;
            const bool hit = traverse_scene<AnyHit>(s, ray, wh, nullptr, local);
            out[i] = hit ? 1 : 0;
            if (hit) ++local.hits;
            ++local.rays;
        }
        parts[tid] = local;
    });
    accumulate(counters, parts);
}

void orc_set_filter(void* scene, uint32_t assembly, uint32_t object_instance, const orc_intersection_filter* filter)
{
    Scene& s = *static_cast<Scene*>(scene);
    const int ti = s.assembly_tree[assembly];
    if (ti < 0 || !filter) return;
    TriTree& tree = *s.trees[ti];
    const orc_assembly& a = s.desc.assemblies[assembly];
    if (tree.filters.empty()) tree.filters.resize(a.object_instance_count);
    const uint32_t triangle_count = s.desc.meshes[a.object_instances[object_instance].mesh_index].triangle_count;
    auto copy_mask = [](const orc_alpha_mask& m, AlphaMask& out)
    {
        if (!m.bits || m.width == 0 || m.height == 0) return;
        out.width = m.width; out.height = m.height;
        out.bits.assign(m.bits, m.bits + size_t((m.width + 7) / 8) * m.height);
    };
    std::unique_ptr<Filter> f(new Filter());
    copy_mask(filter->object_mask, f->object_mask);
    f->material_masks.resize(filter->material_mask_count);
    for (uint32_t i = 0; i < filter->material_mask_count; ++i) copy_mask(filter->material_masks[i], f->material_masks[i]);
    if (filter->uv) f->uv.assign(filter->uv, filter->uv + size_t(triangle_count) * 6);
    else f->uv.assign(size_t(triangle_count) * 6, 0.0f);
    tree.filters[object_instance] = std::move(f);
}

void orc_two_nearest(const void* scene, const orc_rays* rays, size_t n, double* t1, double* t2, int threads)
{
    const Scene& s = *static_cast<const Scene*>(scene);
    parallel_ranges(n, threads, [&](int, size_t begin, size_t end)
    {
        orc_counters local; std::memset(&local, 0, sizeof(local));
        for (size_t i = begin; i < end; ++i)
        {
            Ray ray; load_ray(*rays, i, ray);
            WorldHit wh; TwoBest two;
            traverse_scene<TwoNearest>(s, ray, wh, &two, local);
            t1[i] = two.t1; t2[i] = two.t2;
        }
    });
}

}   // extern "C" (reopened below)

// ---------------------------------------------------------------------------------------------
// ShadingPoint::refine_and_offset (shadingpoint.cpp:362-466), triangle branch with
// RENDERER_ADAPTIVE_OFFSET (intersectionsettings.h:150); refine / adaptive_offset of
// renderer/kernel/intersection/refining.h:97-221; TriangleMTSupportPlane::intersect
// (raytrianglemt.h:300-309).  Vector / scalar is a multiplication by the reciprocal
// (vector.h:638-642).
// ---------------------------------------------------------------------------------------------

namespace
{

inline double plane_intersect(const MTd& tri, const double org[3], const double dir[3])
{
    const double tvec[3] = { org[0] - tri.v0[0], org[1] - tri.v0[1], org[2] - tri.v0[2] };
    double qvec[3], pvec[3];
    cross3(tvec, tri.e0, qvec);
    cross3(dir, tri.e1, pvec);
    return dot3(tri.e1, qvec) / dot3(tri.e0, pvec);
}

inline void offset_step(const double p[3], const double n[3], const int64_t mag, double out[3])
{
    const double Threshold = 1.0e-25;
    const int64_t eps_lut[2] = { mag, -mag };
    for (int i = 0; i < 3; ++i)
    {
        if (std::fabs(p[i]) < Threshold) out[i] = p[i] + n[i] * Threshold;
        else
        {
            uint64_t pi, ni;
            std::memcpy(&pi, &p[i], 8); std::memcpy(&ni, &n[i], 8);
            const uint64_t r = pi + static_cast<uint64_t>(eps_lut[(pi ^ ni) >> 63]);
            std::memcpy(&out[i], &r, 8);
        }
    }
}

inline void offset_point(const MTd& tri, const double p[3], const double n[3], double out[3])
{
    int64_t mag = 8;
    double result[3] = { p[0], p[1], p[2] };
    for (int i = 0; i < 64; ++i)
    {
        double next[3];
        offset_step(result, n, mag, next);
        result[0] = next[0]; result[1] = next[1]; result[2] = next[2];
        if (plane_intersect(tri, result, n) < 0.0) break;
        mag *= 2;
    }
    out[0] = result[0]; out[1] = result[1]; out[2] = result[2];
}

// The triangle the closest-hit leaf visitor leaves behind for a hit (m_hit_triangle,
// triangletree.cpp:1413, 1468-1469), from which read_hit_triangle_data (:1483-1499) makes the
// ShadingPoint's support plane: the stored triangle, or for a moving triangle the one interpolated at
// the ray's normalized time (float product time * msc, float lerp), widened to double.
void hit_triangle(const TriTree& tree, const uint32_t slot, const float time_normalized, MTd& tri)
{
    const uint32_t msc = tree.slot_msc[slot];
    if (msc == 0)
    {
        const float* f = tree.slot_tri.data() + size_t(slot) * 9;
        for (int i = 0; i < 3; ++i) { tri.v0[i] = f[i]; tri.e0[i] = f[3 + i]; tri.e1[i] = f[6 + i]; }
        return;
    }
    const double base_time = time_normalized * msc;         // float * uint32 -> float product, widened
    const size_t base_index = static_cast<size_t>(base_time);
    const float frac = static_cast<float>(base_time - base_index);
    const float omf = 1.0f - frac;
    const V3f* p = tree.slot_poses.data() + tree.slot_pose[slot] + base_index * 3;
    float v[3][3];
    for (int c = 0; c < 3; ++c)
    {
        v[c][0] = p[c].x * omf + p[3 + c].x * frac;
        v[c][1] = p[c].y * omf + p[3 + c].y * frac;
        v[c][2] = p[c].z * omf + p[3 + c].z * frac;
    }
    // TriangleMT<float>(v0, v1, v2): edges in float (raytrianglemt.h:128-137).
    for (int i = 0; i < 3; ++i)
    {
        tri.v0[i] = v[0][i];
        tri.e0[i] = v[1][i] - v[0][i];
        tri.e1[i] = v[2][i] - v[0][i];
    }
}

// Source vertex `index` of a mesh at the ray time (fetch_triangle_source_geometry,
// shadingpoint.cpp:186-256): previous pose * (1 - frac) + next pose * frac in float.
V3f source_vertex(const orc_mesh& mesh, const uint32_t index, const float time_normalized)
{
    if (mesh.motion_segment_count == 0 || mesh.vertex_poses == nullptr) return mesh_vertex(mesh, index);
    const size_t msc = mesh.motion_segment_count;
    const double base_time = time_normalized * msc;         // float * size_t -> float product, widened
    const size_t base_index = static_cast<size_t>(base_time);
    const float frac = static_cast<float>(base_time - base_index);
    const float omf = 1.0f - frac;
    const V3f prev = base_index == 0 ? mesh_vertex(mesh, index) : mesh_pose(mesh, index, base_index - 1);
    const V3f next = mesh_pose(mesh, index, base_index);
    V3f r;
    r.x = prev.x * omf + next.x * frac;
    r.y = prev.y * omf + next.y * frac;
    r.z = prev.z * omf + next.z * frac;
    return r;
}

void refine_offset_one(const Scene& s, const Ray& ray, const orc_hit& h, orc_parent& out)
{
    std::memset(&out, 0, sizeof(out));
    out.assembly_instance = ~uint32_t(0);
    if (h.prim_type != 2) return;
    const orc_assembly_instance& ai = s.desc.assembly_instances[h.assembly_instance];
    const orc_assembly& assembly = s.desc.assemblies[ai.assembly_index];
    const TriTree& tree = *s.trees[s.assembly_tree[ai.assembly_index]];
    out.assembly_instance = h.assembly_instance;

    // refine_space_ray = m_assembly_instance_transform.to_local(m_ray); m_ray.m_tmax is the hit distance.
    const V3d lo = xform_point_d(ai.parent_to_local, ray.org), ld = xform_vector_d(ai.parent_to_local, ray.dir);
    const double dir[3] = { ld.x, ld.y, ld.z };
    double p[3] = { lo.x + dir[0] * h.t, lo.y + dir[1] * h.t, lo.z + dir[2] * h.t };

    MTd tri;
    hit_triangle(tree, h.tri_slot, ray.time_normalized, tri);

    for (int step = 0; step < 2; ++step)
    {
        const double t = plane_intersect(tri, p, dir);
        p[0] += dir[0] * t; p[1] += dir[1] * t; p[2] += dir[2] * t;
    }

    // Geometric normal from the source vertices (object space, float), to assembly space, facing the ray.
    const orc_object_instance& oi = assembly.object_instances[h.object_instance_index];
    const orc_mesh& mesh = s.desc.meshes[oi.mesh_index];
    const uint32_t* t3 = mesh.triangles + size_t(h.primitive_index) * 3;
    const V3f v0 = source_vertex(mesh, t3[0], ray.time_normalized), v1 = source_vertex(mesh, t3[1], ray.time_normalized), v2 = source_vertex(mesh, t3[2], ray.time_normalized);
    const float a[3] = { v1.x - v0.x, v1.y - v0.y, v1.z - v0.z }, b[3] = { v2.x - v0.x, v2.y - v0.y, v2.z - v0.z };
    const float nf[3] = { a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1] };
    const double nd[3] = { nf[0], nf[1], nf[2] };
    const double* m = oi.parent_to_local;
    double n[3] = { m[0] * nd[0] + m[4] * nd[1] + m[8] * nd[2], m[1] * nd[0] + m[5] * nd[1] + m[9] * nd[2], m[2] * nd[0] + m[6] * nd[1] + m[10] * nd[2] };
    if (!(dot3(n, dir) < 0.0)) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
    out.geo_normal[0] = n[0]; out.geo_normal[1] = n[1]; out.geo_normal[2] = n[2];

    // adaptive_offset: n = normalize(n) = n * (1 / norm).
    const double rcp = 1.0 / std::sqrt(dot3(n, n));
    const double un[3] = { n[0] * rcp, n[1] * rcp, n[2] * rcp }, mn[3] = { -un[0], -un[1], -un[2] };
    offset_point(tri, p, un, out.front);
    offset_point(tri, p, mn, out.back);
}

}   // namespace

extern "C" {

void orc_refine_offset(const void* scene, const orc_rays* rays, const orc_hit* hits, size_t n, orc_parent* out, int threads)
{
    const Scene& s = *static_cast<const Scene*>(scene);
    parallel_ranges(n, threads, [&](int, size_t begin, size_t end)
    {
        for (size_t i = begin; i < end; ++i)
        {
            Ray ray; load_ray(*rays, i, ray);
            refine_offset_one(s, ray, hits[i], out[i]);
        }
    });
}

// ShadingPoint::m_triangle_support_plane of every hit: v0, e0, e1 as doubles (zeros for a miss).
void orc_support_planes(const void* scene, const orc_rays* rays, const orc_hit* hits, size_t n, double* planes, int threads)
{
    const Scene& s = *static_cast<const Scene*>(scene);
    parallel_ranges(n, threads, [&](int, size_t begin, size_t end)
    {
        for (size_t i = begin; i < end; ++i)
        {
            double* dst = planes + i * 9;
            for (int k = 0; k < 9; ++k) dst[k] = 0.0;
            const orc_hit& h = hits[i];
            if (h.prim_type != 2) continue;
            const orc_assembly_instance& ai = s.desc.assembly_instances[h.assembly_instance];
            const TriTree& tree = *s.trees[s.assembly_tree[ai.assembly_index]];
            MTd tri;
            hit_triangle(tree, h.tri_slot, rays->time_normalized ? rays->time_normalized[i] : 0.0f, tri);
            for (int k = 0; k < 3; ++k) { dst[k] = tri.v0[k]; dst[3 + k] = tri.e0[k]; dst[6 + k] = tri.e1[k]; }
        }
    });
}

void orc_trace_parents(const void* scene, const orc_rays* rays, const orc_parent* parents, size_t n, orc_hit* out, int threads)
{
    const Scene& s = *static_cast<const Scene*>(scene);
    parallel_ranges(n, threads, [&](int, size_t begin, size_t end)
    {
        orc_counters local; std::memset(&local, 0, sizeof(local));
        for (size_t i = begin; i < end; ++i)
        {
            Ray ray; load_ray(*rays, i, ray);
            WorldHit wh;
            const orc_parent* parent = parents[i].assembly_instance != ~uint32_t(0) ? &parents[i] : nullptr;
            traverse_scene<ClosestHit>(s, ray, wh, nullptr, local, parent);
            orc_hit& h = out[i];
            std::memset(&h, 0, sizeof(h));
            h.t = ray.tmax;
            h.assembly_instance = ~uint32_t(0);
            if (wh.hit)
            {
                const Key& key = s.trees[wh.tree]->keys[wh.slot];
                h.u = wh.u; h.v = wh.v;
                h.assembly_instance = wh.assembly_instance;
                h.object_instance_index = key.object_instance_index;
                h.primitive_index = key.triangle_index;
                h.tri_slot = wh.slot;
                h.motion_segment = wh.motion_segment;
                h.prim_type = 2;
            }
        }
    });
}

void orc_trace_probe_parents(const void* scene, const orc_rays* rays, const orc_parent* parents, size_t n, uint8_t* out, int threads)
{
    const Scene& s = *static_cast<const Scene*>(scene);
    parallel_ranges(n, threads, [&](int, size_t begin, size_t end)
    {
        orc_counters local; std::memset(&local, 0, sizeof(local));
        for (size_t i = begin; i < end; ++i)
        {
            Ray ray; load_ray(*rays, i, ray);
            WorldHit wh;
            const orc_parent* parent = parents[i].assembly_instance != ~uint32_t(0) ? &parents[i] : nullptr;
            out[i] = traverse_scene<AnyHit>(s, ray, wh, nullptr, local, parent) ? 1 : 0;
        }
    });
}

}   // extern "C"

extern "C" {

static void make_mt(const double v0[3], const double v1[3], const double v2[3], MTd& tri)
{
    for (int i = 0; i < 3; ++i) { tri.v0[i] = v0[i]; tri.e0[i] = v1[i] - v0[i]; tri.e1[i] = v2[i] - v0[i]; }
}

static void make_ray(const double org[3], const double dir[3], double tmin, double tmax, Ray& r)
{
    r.org.x = org[0]; r.org.y = org[1]; r.org.z = org[2];
    r.dir.x = dir[0]; r.dir.y = dir[1]; r.dir.z = dir[2];
    r.tmin = tmin; r.tmax = tmax;
    r.time_absolute = r.time_normalized = 0.0f; r.flags = ~uint32_t(0);
}

int orc_kat_ray_triangle(const double v0[3], const double v1[3], const double v2[3], const double org[3], const double dir[3], double tmin, double tmax, double tuv[3])
{
    MTd tri; make_mt(v0, v1, v2, tri);
    Ray r; make_ray(org, dir, tmin, tmax, r);
    double t = 0.0, u = 0.0, v = 0.0;
    const bool hit = mt_intersect(tri, r, t, u, v);
    tuv[0] = t; tuv[1] = u; tuv[2] = v;
    return hit ? 1 : 0;
}

int orc_kat_ray_triangle_bool(const double v0[3], const double v1[3], const double v2[3], const double org[3], const double dir[3], double tmin, double tmax)
{
    MTd tri; make_mt(v0, v1, v2, tri);
    Ray r; make_ray(org, dir, tmin, tmax, r);
    return mt_intersect_bool(tri, r) ? 1 : 0;
}

int orc_kat_ray_aabb(const double bmin[3], const double bmax[3], const double org[3], const double dir[3], double tmin, double tmax, double* tmin_out)
{
    double box[12];
    for (int a = 0; a < 3; ++a) { box[a * 4] = bmin[a]; box[a * 4 + 1] = bmin[a]; box[a * 4 + 2] = bmax[a]; box[a * 4 + 3] = bmax[a]; }
    Ray r; make_ray(org, dir, tmin, tmax, r);
    const RayInfo info(r);
    double t = 0.0;
    const bool hit = slab1(box, 0, r, info, t);
    if (tmin_out) *tmin_out = t;
    return hit ? 1 : 0;
}

// The other entry points of rayaabb.h the reference tests: the 3-argument intersect (the SSE2
// specialisation :192-228 computes the same products and comparisons as the generic code, two lanes
// at a time), the 4-argument one with its "distance unchanged on a miss" contract (:230-252) and
// clip (:310-336).  mode 0: io untouched; mode 1: io[0] = distance in / out; mode 2: io[0], io[1] =
// ray tmin, tmax in / out.
int orc_kat_ray_aabb_ex(int mode, const double bmin[3], const double bmax[3], const double org[3], const double dir[3], double tmin, double tmax, double* io)
{
    Ray r; make_ray(org, dir, tmin, tmax, r);
    const RayInfo info(r);
    const double o[3] = { r.org.x, r.org.y, r.org.z };
    double l1[3], l2[3];
    for (int a = 0; a < 3; ++a)
    {
        const double near_plane = info.sgn[a] ? bmin[a] : bmax[a];      // bbox[1 - sgn]
        const double far_plane = info.sgn[a] ? bmax[a] : bmin[a];       // bbox[sgn]
        l1[a] = info.rcp[a] * (near_plane - o[a]);
        l2[a] = info.rcp[a] * (far_plane - o[a]);
    }
    const double t0 = ssemax(l1[2], ssemax(l1[1], ssemax(l1[0], r.tmin)));
    const double t1 = ssemin(l2[2], ssemin(l2[1], ssemin(l2[0], r.tmax)));
    if (t0 > t1 || t1 < r.tmin || t0 >= r.tmax) return 0;
    if (mode == 1) io[0] = ssemax(r.tmin, t0);
    if (mode == 2) { io[0] = ssemax(r.tmin, t0); io[1] = ssemin(r.tmax, t1); }
    return 1;
}

void orc_kat_ray_info(const double dir[3], double rcp[3], uint32_t sgn[3])
{
    Ray r; const double o[3] = { 0.0, 0.0, 0.0 };
    make_ray(o, dir, 0.0, DBL_BIG, r);
    const RayInfo info(r);
    for (int i = 0; i < 3; ++i) { rcp[i] = info.rcp[i]; sgn[i] = static_cast<uint32_t>(info.sgn[i]); }
}

}   // extern "C"
