//
// ref_driver.cpp -- oracle/_ref: the reference's OWN foundation headers, compiled from
// /root/reference where they lie, driven through the C interface of oracle_api.h.
//
// TEST INFRASTRUCTURE ONLY (see oracle_api.h).  Output goes to oracle/_ref/libasref.so
// (git-ignored); no reference source is copied into this repository.
//
// What is the reference's own code here (by #include, unmodified):
//   foundation/math/bvh/*            Node, Tree, Builder, SAHPartitioner, Intersector (generic + SSE2)
//   foundation/math/intersection/*   rayaabb.h, raytrianglemt.h, aabbtriangle.h
//   foundation/math/{ray,aabb,vector,matrix,transform,area,scalar}.h
// What is restated here because libappleseed cannot be built in this image (Boost, OSL, OIIO,
// OpenEXR, Xerces absent -- SURVEY.md section 8(c)); each block cites the lines it follows:
//   renderer/kernel/intersection/triangletree.cpp      collect / build_bvh / motion boxes / store / leaf visitors
//   renderer/kernel/intersection/triangleencoder.cpp   leaf payload
//   renderer/kernel/intersection/assemblytree.cpp      instance flatten, top tree, leaf visitors, ray transform
//   renderer/kernel/intersection/intersector.cpp       trace / trace_probe
//

#include "foundation/containers/alignedvector.h"
#include "foundation/math/aabb.h"
#include "foundation/math/area.h"
#include "foundation/math/bvh.h"
#include "foundation/math/intersection/aabbtriangle.h"
#include "foundation/math/intersection/rayaabb.h"
#include "foundation/math/intersection/raytrianglemt.h"
#include "foundation/math/matrix.h"
#include "foundation/math/ray.h"
#include "foundation/math/scalar.h"
#include "foundation/math/transform.h"
#include "foundation/math/vector.h"
#include "foundation/memory/alignedallocator.h"
#include "foundation/utility/bitmask.h"
#include "foundation/utility/casts.h"
#include "renderer/kernel/intersection/refining.h"
#include "renderer/utility/transformsequence.h"
#include "renderer/utility/triangle.h"

#include "oracle_api.h"

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

// The product's C++ adaptor (include/asgpu_adaptor.hpp) reads the trees through one friend
// declaration per class, exactly as an in-tree integration would add to renderer::TriangleTree and
// renderer::AssemblyTree; tests/adaptor compiles it against these types.
namespace asgpu_adaptor { class GpuSceneFlattener; }

using namespace foundation;

// Allocation logging hooks declared in main/allocator.h; foundation/memory/memory.cpp calls them.
void log_allocation(const void*, const void*, const size_t) {}
void log_allocation_failure(const size_t) {}
void log_deallocation(const void*, const void*) {}

namespace
{

// renderer/global/globaltypes.h:55-60
typedef float GScalar;
typedef Vector<GScalar, 3> GVector3;
typedef AABB<GScalar, 3> GAABB3;

// renderer/kernel/intersection/intersectionsettings.h:65-68
typedef TriangleMT<GScalar> GTriangleType;
typedef TriangleMT<double> TriangleType;

// A do-nothing timer for bvh::Builder::build<Timer> (foundation/utility/stopwatch.h:129,163,187).
struct NullTimer
{
    std::uint64_t frequency() { return 1; }
    std::uint64_t read_start() { return 0; }
    std::uint64_t read_end() { return 0; }
    std::uint64_t read() { return 0; }
};

// renderer/kernel/intersection/trianglekey.h:63-65
struct TriangleKey
{
    std::uint32_t   m_object_instance_index;
    std::uint16_t   m_triangle_pa;
    std::uint32_t   m_triangle_index;
};
static_assert(sizeof(TriangleKey) == 12, "TriangleKey must be 12 bytes");

// renderer/kernel/intersection/trianglevertexinfo.h:46-48
struct TriangleVertexInfo
{
    size_t          m_vertex_index;
    size_t          m_motion_segment_count;
    std::uint32_t   m_vis_flags;
};

typedef bvh::Node<AABB3d> NodeType;
typedef AlignedVector<NodeType> NodeVector;
static_assert(sizeof(NodeType) == 128, "bvh::Node<AABB3d> must be 128 bytes");

inline Transformd make_transform(const double* l2p, const double* p2l)
{
    return Transformd(Matrix4d::from_array(l2p), Matrix4d::from_array(p2l));
}

struct TriLeafVisitor;
struct TriLeafProbeVisitor;

//
// TriangleTree (renderer/kernel/intersection/triangletree.h:72-125).
//

// IntersectionFilter (renderer/kernel/intersection/intersectionfilter.h:84-128, 169-205) over the
// reference's own BitMask2, Vector2f, clamp and truncate.  (The class itself needs the texture
// system to BUILD its masks and cannot be compiled here; masks arrive ready-made.)
struct RefAlphaMask
{
    const float     m_max_x, m_max_y;
    BitMask2        m_bitmask;

    RefAlphaMask(const orc_alpha_mask& src)
      : m_max_x(static_cast<float>(src.width) - 1.0f)
      , m_max_y(static_cast<float>(src.height) - 1.0f)
      , m_bitmask(src.width, src.height)
    {
        const size_t block_width = (src.width + 7) / 8;
        for (size_t y = 0; y < src.height; ++y)
            for (size_t x = 0; x < src.width; ++x)
                m_bitmask.set(x, y, (src.bits[y * block_width + x / 8] >> (x & 7)) & 1);
    }

    bool is_opaque(const Vector2f& uv) const
    {
        const float fx = clamp(uv[0] * m_bitmask.get_width(), 0.0f, m_max_x);
        const float fy = clamp(uv[1] * m_bitmask.get_height(), 0.0f, m_max_y);
        const size_t ix = truncate<size_t>(fx);
        const size_t iy = truncate<size_t>(fy);
        return m_bitmask.is_set(ix, iy);
    }

    bool is_transparent(const Vector2f& uv) const { return !is_opaque(uv); }
};

struct RefIntersectionFilter
{
    std::unique_ptr<RefAlphaMask>               m_obj_alpha_mask;
    std::vector<std::unique_ptr<RefAlphaMask>>  m_material_alpha_masks;
    std::vector<Vector2f>                       m_uv;

    bool accept(const TriangleKey& triangle_key, const double u, const double v) const
    {
        if (u != u || v != v)
            return true;

        const RefAlphaMask* mtl_alpha_mask =
            triangle_key.m_triangle_pa < m_material_alpha_masks.size() ? m_material_alpha_masks[triangle_key.m_triangle_pa].get() : nullptr;

        if (m_obj_alpha_mask || mtl_alpha_mask)
        {
            const size_t triangle_index = triangle_key.m_triangle_index;
            const float fu = static_cast<float>(u);
            const float fv = static_cast<float>(v);
            const Vector2f uv =
                  m_uv[triangle_index * 3 + 0] * (1.0f - fu - fv)
                + m_uv[triangle_index * 3 + 1] * fu
                + m_uv[triangle_index * 3 + 2] * fv;

            if (m_obj_alpha_mask && m_obj_alpha_mask->is_transparent(uv))
                return false;

            if (mtl_alpha_mask)
                return mtl_alpha_mask->is_opaque(uv);
        }

        return true;
    }
};

class RefTriangleTree
  : public bvh::Tree<NodeVector>
{
  public:
    RefTriangleTree(const orc_scene_desc& desc, const orc_assembly& assembly, const GAABB3& bbox)
      : bvh::Tree<NodeVector>(AlignedAllocator<void>(64))
      , m_desc(desc)
      , m_assembly(assembly)
      , m_bbox(bbox)
    {
        build_bvh();
    }

    friend class asgpu_adaptor::GpuSceneFlattener;

    std::vector<TriangleKey>    m_triangle_keys;
    std::vector<std::uint8_t>   m_leaf_data;
    std::vector<GTriangleType>  m_slot_triangles;       // per leaf slot: the static triangle the leaf stores (checker convenience)
    std::vector<std::uint32_t>  m_slot_msc;             // per leaf slot: the triangle's motion segment count
    std::vector<size_t>         m_slot_pose;            // per leaf slot: first vertex of pose 0 in m_slot_poses (moving triangles)
    std::vector<GVector3>       m_slot_poses;           // (msc + 1) * 3 vertices per moving triangle, as encoded in the leaf
    std::vector<std::unique_ptr<RefIntersectionFilter>> m_intersection_filters;    // per object instance, or empty
    size_t                      m_static_triangle_count = 0;
    size_t                      m_moving_triangle_count = 0;

    const NodeVector& nodes() const { return m_nodes; }
    const std::vector<AABB3d>& node_bboxes() const { return m_node_bboxes; }

  private:
    const orc_scene_desc&   m_desc;
    const orc_assembly&     m_assembly;
    const GAABB3            m_bbox;

    static GVector3 vertex(const orc_mesh& mesh, const size_t i)
    {
        return GVector3(mesh.vertices[i * 3 + 0], mesh.vertices[i * 3 + 1], mesh.vertices[i * 3 + 2]);
    }

    // StaticTessellation::get_vertex_pose (statictessellation.h:320-337).
    static GVector3 vertex_pose(const orc_mesh& mesh, const size_t v, const size_t m)
    {
        const float* p = mesh.vertex_poses + (v * mesh.motion_segment_count + m) * 3;
        return GVector3(p[0], p[1], p[2]);
    }

    // triangletree.cpp:105-196.
    void collect_static_triangles(
        const orc_object_instance&          object_instance,
        const size_t                        object_instance_index,
        const orc_mesh&                     tess,
        std::vector<TriangleKey>*           triangle_keys,
        std::vector<TriangleVertexInfo>*    triangle_vertex_infos,
        std::vector<GVector3>*              triangle_vertices,
        std::vector<GAABB3>*                triangle_bboxes,
        size_t&                             triangle_vertex_count) const
    {
        const Transformd transform =
            make_transform(object_instance.local_to_parent, object_instance.parent_to_local);
        const size_t triangle_count = tess.triangle_count;

        for (size_t i = 0; i < triangle_count; ++i)
        {
            const GVector3 v0_os = vertex(tess, tess.triangles[i * 3 + 0]);
            const GVector3 v1_os = vertex(tess, tess.triangles[i * 3 + 1]);
            const GVector3 v2_os = vertex(tess, tess.triangles[i * 3 + 2]);

            if (square_area(v0_os, v1_os, v2_os) == GScalar(0.0))
                continue;

            const GVector3 v0 = transform.point_to_parent(v0_os);
            const GVector3 v1 = transform.point_to_parent(v1_os);
            const GVector3 v2 = transform.point_to_parent(v2_os);

            if (square_area(v0, v1, v2) == GScalar(0.0))
                continue;

            GAABB3 triangle_bbox;
            triangle_bbox.invalidate();
            triangle_bbox.insert(v0);
            triangle_bbox.insert(v1);
            triangle_bbox.insert(v2);

            if (!intersect(m_bbox, v0, v1, v2))
                continue;

            if (triangle_keys)
            {
                TriangleKey key;
                std::memset(&key, 0, sizeof(key));
                key.m_object_instance_index = static_cast<std::uint32_t>(object_instance_index);
                key.m_triangle_index = static_cast<std::uint32_t>(i);
                key.m_triangle_pa = tess.triangle_pa ? tess.triangle_pa[i] : 0;
                triangle_keys->push_back(key);
            }

            if (triangle_vertex_infos)
            {
                TriangleVertexInfo info;
                info.m_vertex_index = triangle_vertex_count;
                info.m_motion_segment_count = 0;
                info.m_vis_flags = object_instance.vis_flags;
                triangle_vertex_infos->push_back(info);
            }

            if (triangle_vertices)
            {
                triangle_vertices->push_back(v0);
                triangle_vertices->push_back(v1);
                triangle_vertices->push_back(v2);
            }
            triangle_vertex_count += 3;

            if (triangle_bboxes)
                triangle_bboxes->push_back(triangle_bbox);
        }
    }

    // triangletree.cpp:198-314; bbox helpers renderer/utility/bbox.h:96-125.
    void collect_moving_triangles(
        const orc_object_instance&          object_instance,
        const size_t                        object_instance_index,
        const orc_mesh&                     tess,
        const double                        time,
        std::vector<TriangleKey>*           triangle_keys,
        std::vector<TriangleVertexInfo>*    triangle_vertex_infos,
        std::vector<GVector3>*              triangle_vertices,
        std::vector<GAABB3>*                triangle_bboxes,
        size_t&                             triangle_vertex_count) const
    {
        const Transformd transform =
            make_transform(object_instance.local_to_parent, object_instance.parent_to_local);
        const size_t motion_segment_count = tess.motion_segment_count;
        const size_t triangle_count = tess.triangle_count;

        std::vector<GAABB3> tri_pose_bboxes(motion_segment_count + 1);

        for (size_t i = 0; i < triangle_count; ++i)
        {
            const size_t i0 = tess.triangles[i * 3 + 0];
            const size_t i1 = tess.triangles[i * 3 + 1];
            const size_t i2 = tess.triangles[i * 3 + 2];

            const GVector3 v0 = transform.point_to_parent(vertex(tess, i0));
            const GVector3 v1 = transform.point_to_parent(vertex(tess, i1));
            const GVector3 v2 = transform.point_to_parent(vertex(tess, i2));

            tri_pose_bboxes[0].invalidate();
            tri_pose_bboxes[0].insert(v0);
            tri_pose_bboxes[0].insert(v1);
            tri_pose_bboxes[0].insert(v2);
            for (size_t m = 0; m < motion_segment_count; ++m)
            {
                tri_pose_bboxes[m + 1].invalidate();
                tri_pose_bboxes[m + 1].insert(transform.point_to_parent(vertex_pose(tess, i0, m)));
                tri_pose_bboxes[m + 1].insert(transform.point_to_parent(vertex_pose(tess, i1, m)));
                tri_pose_bboxes[m + 1].insert(transform.point_to_parent(vertex_pose(tess, i2, m)));
            }

            // compute_union (bbox.h:96-108).
            GAABB3 triangle_motion_bbox = tri_pose_bboxes[0];
            for (size_t m = 1; m <= motion_segment_count; ++m)
                triangle_motion_bbox.insert(tri_pose_bboxes[m]);

            if (triangle_motion_bbox.rank() < 2)
                continue;

            if (!GAABB3::overlap(m_bbox, triangle_motion_bbox))
                continue;

            // interpolate (bbox.h:110-125).
            const size_t prev_index = truncate<size_t>(time * motion_segment_count);
            const GScalar k = static_cast<GScalar>(time * motion_segment_count - prev_index);
            const GAABB3 triangle_midtime_bbox =
                lerp(tri_pose_bboxes[prev_index], tri_pose_bboxes[prev_index + 1], k);

            if (triangle_midtime_bbox.rank() < 2)
                continue;

            if (triangle_keys)
            {
                TriangleKey key;
                std::memset(&key, 0, sizeof(key));
                key.m_object_instance_index = static_cast<std::uint32_t>(object_instance_index);
                key.m_triangle_index = static_cast<std::uint32_t>(i);
                key.m_triangle_pa = tess.triangle_pa ? tess.triangle_pa[i] : 0;
                triangle_keys->push_back(key);
            }

            if (triangle_vertex_infos)
            {
                TriangleVertexInfo info;
                info.m_vertex_index = triangle_vertex_count;
                info.m_motion_segment_count = motion_segment_count;
                info.m_vis_flags = object_instance.vis_flags;
                triangle_vertex_infos->push_back(info);
            }

            if (triangle_vertices)
            {
                triangle_vertices->push_back(v0);
                triangle_vertices->push_back(v1);
                triangle_vertices->push_back(v2);
                for (size_t m = 0; m < motion_segment_count; ++m)
                {
                    triangle_vertices->push_back(transform.point_to_parent(vertex_pose(tess, i0, m)));
                    triangle_vertices->push_back(transform.point_to_parent(vertex_pose(tess, i1, m)));
                    triangle_vertices->push_back(transform.point_to_parent(vertex_pose(tess, i2, m)));
                }
            }
            triangle_vertex_count += (motion_segment_count + 1) * 3;

            if (triangle_bboxes)
                triangle_bboxes->push_back(triangle_midtime_bbox);
        }
    }

    // triangletree.cpp:316-383.
    void collect_triangles(
        const double                        time,
        std::vector<TriangleKey>*           triangle_keys,
        std::vector<TriangleVertexInfo>*    triangle_vertex_infos,
        std::vector<GVector3>*              triangle_vertices,
        std::vector<GAABB3>*                triangle_bboxes) const
    {
        size_t triangle_vertex_count = 0;

        for (size_t i = 0; i < m_assembly.object_instance_count; ++i)
        {
            const orc_object_instance& object_instance = m_assembly.object_instances[i];
            const orc_mesh& tess = m_desc.meshes[object_instance.mesh_index];

            if (tess.motion_segment_count > 0)
            {
                collect_moving_triangles(
                    object_instance, i, tess, time,
                    triangle_keys, triangle_vertex_infos, triangle_vertices, triangle_bboxes,
                    triangle_vertex_count);
            }
            else
            {
                collect_static_triangles(
                    object_instance, i, tess,
                    triangle_keys, triangle_vertex_infos, triangle_vertices, triangle_bboxes,
                    triangle_vertex_count);
            }
        }
    }

    // triangletree.cpp:497-597.
    void build_bvh()
    {
        const double time = m_assembly.time;

        std::vector<TriangleKey> triangle_keys;
        std::vector<TriangleVertexInfo> triangle_vertex_infos;
        std::vector<GAABB3> triangle_bboxes;
        collect_triangles(time, &triangle_keys, &triangle_vertex_infos, nullptr, &triangle_bboxes);

        for (size_t i = 0; i < triangle_vertex_infos.size(); ++i)
        {
            if (triangle_vertex_infos[i].m_motion_segment_count == 0)
                ++m_static_triangle_count;
        }
        m_moving_triangle_count = triangle_vertex_infos.size() - m_static_triangle_count;

        typedef bvh::SAHPartitioner<std::vector<GAABB3>> Partitioner;
        Partitioner partitioner(
            triangle_bboxes,
            m_assembly.max_leaf_size,
            m_assembly.interior_node_traversal_cost,
            m_assembly.triangle_intersection_cost);

        typedef bvh::Builder<RefTriangleTree, Partitioner> Builder;
        Builder builder;
        builder.build<NullTimer>(*this, partitioner, triangle_keys.size(), m_assembly.max_leaf_size);

        std::vector<GVector3> triangle_vertices;
        collect_triangles(time, nullptr, nullptr, &triangle_vertices, nullptr);

        compute_motion_bboxes(
            partitioner.get_item_ordering(), triangle_vertex_infos, triangle_vertices, 0);

        store_triangles(
            partitioner.get_item_ordering(), triangle_vertex_infos, triangle_vertices, triangle_keys);
    }

    // triangletree.cpp:725-738 (APPLESEED_USE_SSE variant).
    static AABB3d swizzle(const AABB3d& bbox)
    {
        AABB3d result;
        double* flat_result = &result[0][0];
        for (size_t i = 0; i < 3; ++i)
        {
            flat_result[i * 2 + 0] = bbox[0][i];
            flat_result[i * 2 + 1] = bbox[1][i];
        }
        return result;
    }

    // triangletree.cpp:755-876.
    std::vector<GAABB3> compute_motion_bboxes(
        const std::vector<size_t>&              triangle_indices,
        const std::vector<TriangleVertexInfo>&  triangle_vertex_infos,
        const std::vector<GVector3>&            triangle_vertices,
        const size_t                            node_index)
    {
        if (m_nodes[node_index].is_interior())
        {
            const size_t child = m_nodes[node_index].get_child_node_index();

            const std::vector<GAABB3> left_bboxes =
                compute_motion_bboxes(triangle_indices, triangle_vertex_infos, triangle_vertices, child + 0);
            const std::vector<GAABB3> right_bboxes =
                compute_motion_bboxes(triangle_indices, triangle_vertex_infos, triangle_vertices, child + 1);

            NodeType& node = m_nodes[node_index];
            node.set_left_bbox_count(left_bboxes.size());
            node.set_right_bbox_count(right_bboxes.size());

            if (left_bboxes.size() > 1)
            {
                node.set_left_bbox_index(m_node_bboxes.size());
                for (const GAABB3& b : left_bboxes)
                    m_node_bboxes.push_back(swizzle(AABB3d(b)));
            }

            if (right_bboxes.size() > 1)
            {
                node.set_right_bbox_index(m_node_bboxes.size());
                for (const GAABB3& b : right_bboxes)
                    m_node_bboxes.push_back(swizzle(AABB3d(b)));
            }

            const size_t bbox_count = std::max(left_bboxes.size(), right_bboxes.size());
            std::vector<GAABB3> bboxes(bbox_count);

            for (size_t i = 0; i < bbox_count; ++i)
            {
                bboxes[i] = left_bboxes[i * left_bboxes.size() / bbox_count];
                bboxes[i].insert(right_bboxes[i * right_bboxes.size() / bbox_count]);
            }

            return bboxes;
        }
        else
        {
            const NodeType& node = m_nodes[node_index];
            const size_t item_begin = node.get_item_index();
            const size_t item_count = node.get_item_count();

            size_t max_motion_segment_count = 0;

            GAABB3 base_pose_bbox;
            base_pose_bbox.invalidate();

            for (size_t i = 0; i < item_count; ++i)
            {
                const size_t triangle_index = triangle_indices[item_begin + i];
                const TriangleVertexInfo& vertex_info = triangle_vertex_infos[triangle_index];

                if (max_motion_segment_count < vertex_info.m_motion_segment_count)
                    max_motion_segment_count = vertex_info.m_motion_segment_count;

                base_pose_bbox.insert(triangle_vertices[vertex_info.m_vertex_index + 0]);
                base_pose_bbox.insert(triangle_vertices[vertex_info.m_vertex_index + 1]);
                base_pose_bbox.insert(triangle_vertices[vertex_info.m_vertex_index + 2]);
            }

            std::vector<GAABB3> bboxes(max_motion_segment_count + 1);
            bboxes[0] = base_pose_bbox;

            if (max_motion_segment_count > 0)
            {
                for (size_t m = 0; m < max_motion_segment_count - 1; ++m)
                {
                    bboxes[m + 1].invalidate();

                    const double time = static_cast<double>(m + 1) / max_motion_segment_count;

                    for (size_t i = 0; i < item_count; ++i)
                    {
                        const size_t triangle_index = triangle_indices[item_begin + i];
                        const TriangleVertexInfo& vertex_info = triangle_vertex_infos[triangle_index];

                        const size_t prev_pose_index = truncate<size_t>(time * vertex_info.m_motion_segment_count);
                        const size_t base_vertex_index = vertex_info.m_vertex_index + prev_pose_index * 3;
                        const GScalar k = static_cast<GScalar>(time * vertex_info.m_motion_segment_count - prev_pose_index);

                        bboxes[m + 1].insert(lerp(triangle_vertices[base_vertex_index + 0], triangle_vertices[base_vertex_index + 3], k));
                        bboxes[m + 1].insert(lerp(triangle_vertices[base_vertex_index + 1], triangle_vertices[base_vertex_index + 4], k));
                        bboxes[m + 1].insert(lerp(triangle_vertices[base_vertex_index + 2], triangle_vertices[base_vertex_index + 5], k));
                    }
                }

                bboxes[max_motion_segment_count].invalidate();

                for (size_t i = 0; i < item_count; ++i)
                {
                    const size_t triangle_index = triangle_indices[item_begin + i];
                    const TriangleVertexInfo& vertex_info = triangle_vertex_infos[triangle_index];
                    const size_t base_vertex_index = vertex_info.m_vertex_index + vertex_info.m_motion_segment_count * 3;

                    bboxes[max_motion_segment_count].insert(triangle_vertices[base_vertex_index + 0]);
                    bboxes[max_motion_segment_count].insert(triangle_vertices[base_vertex_index + 1]);
                    bboxes[max_motion_segment_count].insert(triangle_vertices[base_vertex_index + 2]);
                }
            }

            return bboxes;
        }
    }

    // triangleencoder.cpp:48-70.
    static size_t encoded_size(
        const std::vector<TriangleVertexInfo>&  triangle_vertex_infos,
        const std::vector<size_t>&              triangle_indices,
        const size_t                            item_begin,
        const size_t                            item_count)
    {
        size_t size = 0;
        for (size_t i = 0; i < item_count; ++i)
        {
            const TriangleVertexInfo& vertex_info = triangle_vertex_infos[triangle_indices[item_begin + i]];
            size += 2 * sizeof(std::uint32_t);
            if (vertex_info.m_motion_segment_count == 0)
                size += sizeof(GTriangleType);
            else size += (vertex_info.m_motion_segment_count + 1) * 3 * sizeof(GVector3);
        }
        return size;
    }

    // triangleencoder.cpp:72-103.
    static std::uint8_t* encode(
        const std::vector<TriangleVertexInfo>&  triangle_vertex_infos,
        const std::vector<GVector3>&            triangle_vertices,
        const std::vector<size_t>&              triangle_indices,
        const size_t                            item_begin,
        const size_t                            item_count,
        std::uint8_t*                           out)
    {
        for (size_t i = 0; i < item_count; ++i)
        {
            const TriangleVertexInfo& vertex_info = triangle_vertex_infos[triangle_indices[item_begin + i]];

            const std::uint32_t vis = vertex_info.m_vis_flags;
            const std::uint32_t msc = static_cast<std::uint32_t>(vertex_info.m_motion_segment_count);
            std::memcpy(out, &vis, 4); out += 4;
            std::memcpy(out, &msc, 4); out += 4;

            if (msc == 0)
            {
                const GTriangleType triangle(
                    triangle_vertices[vertex_info.m_vertex_index + 0],
                    triangle_vertices[vertex_info.m_vertex_index + 1],
                    triangle_vertices[vertex_info.m_vertex_index + 2]);
                std::memcpy(out, &triangle, sizeof(triangle)); out += sizeof(triangle);
            }
            else
            {
                const size_t bytes = (msc + 1) * 3 * sizeof(GVector3);
                std::memcpy(out, &triangle_vertices[vertex_info.m_vertex_index], bytes); out += bytes;
            }
        }
        return out;
    }

    // triangletree.cpp:878-978.
    void store_triangles(
        const std::vector<size_t>&              triangle_indices,
        const std::vector<TriangleVertexInfo>&  triangle_vertex_infos,
        const std::vector<GVector3>&            triangle_vertices,
        const std::vector<TriangleKey>&         triangle_keys)
    {
        const size_t node_count = m_nodes.size();
        const size_t MaxUserDataSize = NodeType::MaxUserDataSize;

        size_t leaf_data_size = 0;
        for (size_t i = 0; i < node_count; ++i)
        {
            const NodeType& node = m_nodes[i];
            if (node.is_leaf())
            {
                const size_t leaf_size =
                    encoded_size(triangle_vertex_infos, triangle_indices, node.get_item_index(), node.get_item_count());
                // The reference sizes m_leaf_data with '< MaxUserDataSize' (:911) but spills with
                // '<= MaxUserDataSize - 4' (:950); sizes in (92, 96) cannot occur (all sizes are 8 + 36k).
                if (!(leaf_size <= MaxUserDataSize - sizeof(std::uint32_t)))
                    leaf_data_size += leaf_size;
            }
        }

        m_triangle_keys.reserve(triangle_indices.size());
        m_leaf_data.resize(leaf_data_size);
        std::uint8_t* leaf_writer = m_leaf_data.empty() ? nullptr : &m_leaf_data[0];

        for (size_t i = 0; i < node_count; ++i)
        {
            NodeType& node = m_nodes[i];
            if (node.is_leaf())
            {
                const size_t item_begin = node.get_item_index();
                const size_t item_count = node.get_item_count();

                node.set_item_index(m_triangle_keys.size());

                for (size_t j = 0; j < item_count; ++j)
                {
                    m_triangle_keys.push_back(triangle_keys[triangle_indices[item_begin + j]]);
                    const TriangleVertexInfo& info = triangle_vertex_infos[triangle_indices[item_begin + j]];
                    GTriangleType triangle;
                    triangle.m_v0 = triangle.m_e0 = triangle.m_e1 = GVector3(0.0f);
                    if (info.m_motion_segment_count == 0)
                        triangle = GTriangleType(triangle_vertices[info.m_vertex_index], triangle_vertices[info.m_vertex_index + 1], triangle_vertices[info.m_vertex_index + 2]);
                    m_slot_triangles.push_back(triangle);
                    m_slot_msc.push_back(static_cast<std::uint32_t>(info.m_motion_segment_count));
                    m_slot_pose.push_back(m_slot_poses.size());
                    if (info.m_motion_segment_count != 0)
                        m_slot_poses.insert(m_slot_poses.end(), triangle_vertices.begin() + info.m_vertex_index,
                                            triangle_vertices.begin() + info.m_vertex_index + (info.m_motion_segment_count + 1) * 3);
                }

                const size_t leaf_size =
                    encoded_size(triangle_vertex_infos, triangle_indices, item_begin, item_count);

                std::uint8_t* user_data = &node.get_user_data<std::uint8_t>();

                if (leaf_size <= MaxUserDataSize - sizeof(std::uint32_t))
                {
                    const std::uint32_t marker = ~std::uint32_t(0);
                    std::memcpy(user_data, &marker, 4);
                    encode(triangle_vertex_infos, triangle_vertices, triangle_indices, item_begin, item_count, user_data + 4);
                }
                else
                {
                    const std::uint32_t offset = static_cast<std::uint32_t>(leaf_writer - &m_leaf_data[0]);
                    std::memcpy(user_data, &offset, 4);
                    leaf_writer = encode(triangle_vertex_infos, triangle_vertices, triangle_indices, item_begin, item_count, leaf_writer);
                }
            }
        }
    }
};

//
// The slice of ShadingRay / ShadingPoint the path touches
// (shadingray.h:99-109, shadingpoint.h:289-302).
//

struct RefShadingRay
  : public Ray3d
{
    float           m_time_absolute;
    float           m_time_normalized;
    std::uint32_t   m_flags;
};

struct RefShadingPoint
{
    RefShadingRay   m_ray;
    bool            m_hit = false;          // m_primitive_type == PrimitiveTriangle
    float           m_bary[2];
    std::uint32_t   m_assembly_instance = ~std::uint32_t(0);
    std::uint32_t   m_object_instance_index = 0;
    std::uint32_t   m_primitive_index = 0;
    std::uint32_t   m_tri_slot = 0;
    std::uint32_t   m_motion_segment = 0;
    TriangleMTSupportPlane<double> m_triangle_support_plane;    // shadingpoint.h:300, written at triangletree.cpp:1496-1497
    Transformd      m_assembly_instance_transform;              // shadingpoint.h:297, written at assemblytree.cpp:738
};

//
// TriangleLeafVisitor (triangletree.cpp:1352-1499).
//

struct TriLeafVisitor
{
    const RefTriangleTree&  m_tree;
    RefShadingPoint&        m_shading_point;
    bool                    m_has_hit = false;
    size_t                  m_hit_triangle_index = 0;
    std::uint32_t           m_hit_motion_segment = 0;
    // m_hit_triangle (triangletree.h:233) points into the leaf data for a static triangle and at the
    // member m_interpolated_triangle (triangletree.h:232, assigned at triangletree.cpp:1468-1469) for
    // a moving one; this driver copies leaf triangles out of unaligned storage, so both cases keep a copy.
    GTriangleType           m_interpolated_triangle;
    orc_counters*           m_counters;

    TriLeafVisitor(const RefTriangleTree& tree, RefShadingPoint& sp, orc_counters* counters)
      : m_tree(tree), m_shading_point(sp), m_counters(counters) {}

    bool visit(
        const NodeType&     node,
        const Ray3d&        ray,
        const RayInfo3d&    /*ray_info*/,
        double&             distance)
    {
        const std::uint8_t* user_data = &node.get_user_data<std::uint8_t>();
        std::uint32_t leaf_data_index;
        std::memcpy(&leaf_data_index, user_data, 4);
        const std::uint8_t* reader =
            leaf_data_index == ~std::uint32_t(0)
                ? user_data + sizeof(std::uint32_t)
                : &m_tree.m_leaf_data[leaf_data_index];

        for (size_t triangle_index = node.get_item_index(),
                    triangle_count = node.get_item_count();
                    triangle_count--;
                    triangle_index++)
        {
            ++m_counters->triangles_tested;

            std::uint32_t vis_flags, motion_segment_count;
            std::memcpy(&vis_flags, reader, 4); reader += 4;
            std::memcpy(&motion_segment_count, reader, 4); reader += 4;

            if (motion_segment_count == 0)
            {
                if (!(vis_flags & m_shading_point.m_ray.m_flags))
                {
                    reader += sizeof(GTriangleType);
                    continue;
                }

                GTriangleType gtriangle;
                std::memcpy(&gtriangle, reader, sizeof(GTriangleType)); reader += sizeof(GTriangleType);
                const TriangleType triangle(gtriangle);

                double t, u, v;
                if (triangle.intersect(ray, t, u, v))
                {
                    // Optionally filter intersections (triangletree.cpp:1404-1411).
                    if (!m_tree.m_intersection_filters.empty())
                    {
                        const TriangleKey& triangle_key = m_tree.m_triangle_keys[triangle_index];
                        const RefIntersectionFilter* filter = m_tree.m_intersection_filters[triangle_key.m_object_instance_index].get();
                        if (filter && !filter->accept(triangle_key, u, v))
                            continue;
                    }
                    m_has_hit = true;
                    m_interpolated_triangle = gtriangle;
                    m_hit_triangle_index = triangle_index;
                    m_hit_motion_segment = 0;
                    m_shading_point.m_ray.m_tmax = t;
                    m_shading_point.m_bary[0] = static_cast<float>(u);
                    m_shading_point.m_bary[1] = static_cast<float>(v);
                }
            }
            else
            {
                const size_t TriangleSize = 3 * sizeof(GVector3);

                if (!(vis_flags & m_shading_point.m_ray.m_flags))
                {
                    reader += (motion_segment_count + 1) * TriangleSize;
                    continue;
                }

                // float * uint32 -> float product, widened (triangletree.cpp:1433).
                const double base_time = m_shading_point.m_ray.m_time_normalized * motion_segment_count;
                const size_t base_index = truncate<size_t>(base_time);
                reader += base_index * TriangleSize;

                const GScalar frac = static_cast<GScalar>(base_time - base_index);
                const GScalar one_minus_frac = GScalar(1.0) - frac;
                GVector3 p[6];
                std::memcpy(p, reader, 2 * TriangleSize); reader += 2 * TriangleSize;
                GVector3 v0 = p[0] * one_minus_frac;
                GVector3 v1 = p[1] * one_minus_frac;
                GVector3 v2 = p[2] * one_minus_frac;
                v0 += p[3] * frac;
                v1 += p[4] * frac;
                v2 += p[5] * frac;

                reader += (motion_segment_count - base_index - 1) * TriangleSize;

                const GTriangleType gtriangle(v0, v1, v2);
                const TriangleType triangle(gtriangle);

                double t, u, v;
                if (triangle.intersect(ray, t, u, v))
                {
                    // Optionally filter intersections (triangletree.cpp:1455-1462).
                    if (!m_tree.m_intersection_filters.empty())
                    {
                        const TriangleKey& triangle_key = m_tree.m_triangle_keys[triangle_index];
                        const RefIntersectionFilter* filter = m_tree.m_intersection_filters[triangle_key.m_object_instance_index].get();
                        if (filter && !filter->accept(triangle_key, u, v))
                            continue;
                    }
                    m_has_hit = true;
                    m_interpolated_triangle = gtriangle;
                    m_hit_triangle_index = triangle_index;
                    m_hit_motion_segment = static_cast<std::uint32_t>(base_index);
                    m_shading_point.m_ray.m_tmax = t;
                    m_shading_point.m_bary[0] = static_cast<float>(u);
                    m_shading_point.m_bary[1] = static_cast<float>(v);
                }
            }
        }

        distance = m_shading_point.m_ray.m_tmax;
        return true;
    }

    // triangletree.cpp:1483-1499.
    void read_hit_triangle_data() const
    {
        if (m_has_hit)
        {
            m_shading_point.m_hit = true;
            const TriangleKey& key = m_tree.m_triangle_keys[m_hit_triangle_index];
            m_shading_point.m_object_instance_index = key.m_object_instance_index;
            m_shading_point.m_primitive_index = key.m_triangle_index;
            m_shading_point.m_tri_slot = static_cast<std::uint32_t>(m_hit_triangle_index);
            m_shading_point.m_motion_segment = m_hit_motion_segment;
            // Compute and store the support plane of the hit triangle (triangletree.cpp:1495-1497).
            m_shading_point.m_triangle_support_plane.initialize(TriangleType(m_interpolated_triangle));
        }
    }
};

//
// TriangleLeafProbeVisitor (triangletree.cpp:1506-1603).
//

struct TriLeafProbeVisitor
{
    const RefTriangleTree&  m_tree;
    const double            m_ray_time;
    const std::uint32_t     m_ray_flags;
    bool                    m_hit = false;
    orc_counters*           m_counters;

    TriLeafProbeVisitor(const RefTriangleTree& tree, const double ray_time, const std::uint32_t ray_flags, orc_counters* counters)
      : m_tree(tree), m_ray_time(ray_time), m_ray_flags(ray_flags), m_counters(counters) {}

    bool visit(
        const NodeType&     node,
        const Ray3d&        ray,
        const RayInfo3d&    /*ray_info*/,
        double&             distance)
    {
        const std::uint8_t* user_data = &node.get_user_data<std::uint8_t>();
        std::uint32_t leaf_data_index;
        std::memcpy(&leaf_data_index, user_data, 4);
        const std::uint8_t* reader =
            leaf_data_index == ~std::uint32_t(0)
                ? user_data + sizeof(std::uint32_t)
                : &m_tree.m_leaf_data[leaf_data_index];

        for (size_t triangle_count = node.get_item_count(); triangle_count--; )
        {
            ++m_counters->triangles_tested;

            std::uint32_t vis_flags, motion_segment_count;
            std::memcpy(&vis_flags, reader, 4); reader += 4;
            std::memcpy(&motion_segment_count, reader, 4); reader += 4;

            if (motion_segment_count == 0)
            {
                if (!(vis_flags & m_ray_flags))
                {
                    reader += sizeof(GTriangleType);
                    continue;
                }

                GTriangleType gtriangle;
                std::memcpy(&gtriangle, reader, sizeof(GTriangleType)); reader += sizeof(GTriangleType);
                const TriangleType triangle(gtriangle);

                if (triangle.intersect(ray))
                {
                    m_hit = true;
                    return false;
                }
            }
            else
            {
                const size_t TriangleSize = 3 * sizeof(GVector3);

                if (!(vis_flags & m_ray_flags))
                {
                    reader += (motion_segment_count + 1) * TriangleSize;
                    continue;
                }

                // double * uint32 -> double product (triangletree.cpp:1570).
                const double base_time = m_ray_time * motion_segment_count;
                const size_t base_index = truncate<size_t>(base_time);
                reader += base_index * TriangleSize;

                const GScalar frac = static_cast<GScalar>(base_time - base_index);
                const GScalar one_minus_frac = GScalar(1.0) - frac;
                GVector3 p[6];
                std::memcpy(p, reader, 2 * TriangleSize); reader += 2 * TriangleSize;
                GVector3 v0 = p[0] * one_minus_frac;
                GVector3 v1 = p[1] * one_minus_frac;
                GVector3 v2 = p[2] * one_minus_frac;
                v0 += p[3] * frac;
                v1 += p[4] * frac;
                v2 += p[5] * frac;

                const GTriangleType gtriangle(v0, v1, v2);
                const TriangleType triangle(gtriangle);

                if (triangle.intersect(ray))
                {
                    m_hit = true;
                    return false;
                }

                reader += (motion_segment_count - base_index - 1) * TriangleSize;
            }
        }

        distance = ray.m_tmax;
        return true;
    }
};

//
// AssemblyTree (assemblytree.h:72-134, assemblytree.cpp:111-245, 372-420).
//

struct RefItem
{
    std::uint32_t   m_assembly_instance;    // index into desc.assembly_instances
    std::uint32_t   m_tree;                 // triangle tree index or ~0
    std::uint32_t   m_vis_flags;
    Transformd      m_transform;
    // Animated instances: the reference's own TransformSequence (several keys), plus what an
    // in-tree flattener would read out of it for the GPU engine.
    bool                                m_animated = false;
    renderer::TransformSequence         m_transform_sequence;
    std::vector<float>                  m_key_times;
    std::vector<double>                 m_key_parent_to_local;
    std::vector<orc_transform_segment>  m_segments;
};

class RefAssemblyTree
  : public bvh::Tree<NodeVector>
{
  public:
    RefAssemblyTree()
      : bvh::Tree<NodeVector>(AlignedAllocator<void>(64)) {}

    friend class asgpu_adaptor::GpuSceneFlattener;

    std::vector<RefItem>                            m_items;
    std::vector<std::unique_ptr<RefTriangleTree>>   m_triangle_trees;       // one per assembly with geometry
    std::vector<int>                                m_assembly_tree_index;  // assembly -> tree or -1
    std::vector<std::uint32_t>                      m_item_assembly_instance;
    std::vector<std::uint32_t>                      m_item_tree;

    const NodeVector& nodes() const { return m_nodes; }

    // ObjectInstance::compute_parent_bbox (objectinstance.cpp:255-267) over
    // StaticTessellation::compute_local_bbox (statictessellation.h:462-479).
    static GAABB3 object_instance_parent_bbox(const orc_scene_desc& desc, const orc_object_instance& oi)
    {
        const orc_mesh& mesh = desc.meshes[oi.mesh_index];
        GAABB3 bbox;
        bbox.invalidate();
        for (size_t i = 0; i < mesh.vertex_count; ++i)
        {
            bbox.insert(GVector3(mesh.vertices[i * 3], mesh.vertices[i * 3 + 1], mesh.vertices[i * 3 + 2]));
            for (size_t j = 0; j < mesh.motion_segment_count; ++j)
            {
                const float* p = mesh.vertex_poses + (i * mesh.motion_segment_count + j) * 3;
                bbox.insert(GVector3(p[0], p[1], p[2]));
            }
        }
        return make_transform(oi.local_to_parent, oi.parent_to_local).to_parent(bbox);
    }

    // Assembly::compute_non_hierarchical_local_bbox (assembly.cpp:219-225).
    static GAABB3 assembly_bbox(const orc_scene_desc& desc, const orc_assembly& assembly)
    {
        GAABB3 bbox;
        bbox.invalidate();
        for (size_t i = 0; i < assembly.object_instance_count; ++i)
            bbox.insert(object_instance_parent_bbox(desc, assembly.object_instances[i]));
        return bbox;
    }

    void build(const orc_scene_desc& desc, const orc_instance_keys* keys = nullptr)
    {
        // update_tree_hierarchy / create_triangle_tree (assemblytree.cpp:247-298, 394-420):
        // one triangle tree per assembly that has mesh object instances.
        m_assembly_tree_index.assign(desc.assembly_count, -1);
        std::vector<GAABB3> assembly_bboxes(desc.assembly_count);
        for (size_t a = 0; a < desc.assembly_count; ++a)
        {
            const orc_assembly& assembly = desc.assemblies[a];
            assembly_bboxes[a] = assembly_bbox(desc, assembly);
            if (assembly.object_instance_count > 0)
            {
                m_assembly_tree_index[a] = static_cast<int>(m_triangle_trees.size());
                m_triangle_trees.emplace_back(new RefTriangleTree(desc, assembly, assembly_bboxes[a]));
            }
        }

        // collect_assembly_instances (assemblytree.cpp:111-152); the description is already flat.
        std::vector<AABB3d> assembly_instance_bboxes;
        for (size_t i = 0; i < desc.assembly_instance_count; ++i)
        {
            const orc_assembly_instance& inst = desc.assembly_instances[i];
            const orc_assembly& assembly = desc.assemblies[inst.assembly_index];

            if (assembly.object_instance_count == 0)
                continue;

            RefItem item;
            item.m_assembly_instance = static_cast<std::uint32_t>(i);
            item.m_tree = static_cast<std::uint32_t>(m_assembly_tree_index[inst.assembly_index]);
            item.m_vis_flags = inst.vis_flags;
            item.m_transform = make_transform(inst.local_to_parent, inst.parent_to_local);
            if (keys && keys[i].key_count >= 2)
            {
                // cumulated_transform_seq of an animated instance (assemblytree.cpp:124-127).
                item.m_animated = true;
                for (std::uint32_t k = 0; k < keys[i].key_count; ++k)
                {
                    const Transformd xf = make_transform(keys[i].local_to_parent + k * 16, keys[i].parent_to_local + k * 16);
                    item.m_transform_sequence.set_transform(keys[i].times[k], xf);
                    item.m_key_times.push_back(keys[i].times[k]);
                    for (int e = 0; e < 16; ++e) item.m_key_parent_to_local.push_back(xf.get_parent_to_local()[e]);
                }
                item.m_transform_sequence.prepare();
                for (std::uint32_t k = 0; k + 1 < keys[i].key_count; ++k)
                {
                    // TransformSequence::prepare builds exactly these (transformsequence.cpp:218-226).
                    const TransformInterpolatord interp(
                        make_transform(keys[i].local_to_parent + k * 16, keys[i].parent_to_local + k * 16),
                        make_transform(keys[i].local_to_parent + (k + 1) * 16, keys[i].parent_to_local + (k + 1) * 16));
                    orc_transform_segment seg;
                    for (int a = 0; a < 3; ++a)
                    {
                        seg.s0[a] = interp.get_s0()[a]; seg.s1[a] = interp.get_s1()[a];
                        seg.t0[a] = interp.get_t0()[a]; seg.t1[a] = interp.get_t1()[a];
                        seg.q0[1 + a] = interp.get_q0().v[a]; seg.q1[1 + a] = interp.get_q1().v[a];
                    }
                    seg.q0[0] = interp.get_q0().s; seg.q1[0] = interp.get_q1().s;
                    item.m_segments.push_back(seg);
                }
            }
            m_items.push_back(item);

            // cumulated_transform_seq.to_parent(...) (assemblytree.cpp:146-149): for an animated
            // instance the reference's motion bounding box (transformsequence.h:215-241).
            AABB3d assembly_instance_bbox(
                item.m_animated ? item.m_transform_sequence.to_parent(assembly_bboxes[inst.assembly_index])
                                : item.m_transform.to_parent(assembly_bboxes[inst.assembly_index]));
            assembly_instance_bbox.robust_grow(1.0e-15);
            assembly_instance_bboxes.push_back(assembly_instance_bbox);
        }

        // rebuild_assembly_tree (assemblytree.cpp:154-212); intersectionsettings.h:51-53.
        typedef bvh::SAHPartitioner<std::vector<AABB3d>> Partitioner;
        Partitioner partitioner(assembly_instance_bboxes, 1, 1.0, 10.0);

        typedef bvh::Builder<RefAssemblyTree, Partitioner> Builder;
        Builder builder;
        builder.build<NullTimer>(*this, partitioner, m_items.size(), 1);

        if (!m_items.empty())
        {
            const std::vector<size_t>& ordering = partitioner.get_item_ordering();
            std::vector<RefItem> reordered(ordering.size());
            for (size_t i = 0; i < ordering.size(); ++i)
                reordered[i] = m_items[ordering[i]];
            m_items.swap(reordered);
        }

        for (const RefItem& item : m_items)
        {
            m_item_assembly_instance.push_back(item.m_assembly_instance);
            m_item_tree.push_back(item.m_tree);
        }
    }
};

// compute_assembly_instance_ray (assemblytree.cpp:556-596).  `parent` stands for parent_sp: its
// assembly instance uid, and get_offset_point (shadingpoint.h:604-613) over the refined points.
inline void compute_assembly_instance_ray(
    const Transformd&       transform,
    const std::uint32_t     assembly_instance,
    const orc_parent*       parent,
    const RefShadingRay&    input_ray,
    RefShadingRay&          output_ray)
{
    output_ray.m_dir = transform.vector_to_local(input_ray.m_dir);
    if (parent && parent->assembly_instance == assembly_instance)
    {
        const Vector3d geo_normal(parent->geo_normal[0], parent->geo_normal[1], parent->geo_normal[2]);
        const double* p = dot(geo_normal, output_ray.m_dir) > 0.0 ? parent->front : parent->back;
        output_ray.m_org = Vector3d(p[0], p[1], p[2]);
    }
    else output_ray.m_org = transform.point_to_local(input_ray.m_org);
    output_ray.m_tmin = input_ray.m_tmin;
    output_ray.m_tmax = input_ray.m_tmax;
    output_ray.m_time_absolute = input_ray.m_time_absolute;
    output_ray.m_time_normalized = input_ray.m_time_normalized;
    output_ray.m_flags = input_ray.m_flags;
}

typedef bvh::Intersector<RefTriangleTree, TriLeafVisitor, Ray3d, 64> TriangleTreeIntersector;
typedef bvh::Intersector<RefTriangleTree, TriLeafProbeVisitor, Ray3d, 64> TriangleTreeProbeIntersector;

// AssemblyLeafVisitor (assemblytree.cpp:604-838), triangle branch.
struct AsmLeafVisitor
{
    RefShadingPoint&        m_shading_point;
    const RefAssemblyTree&  m_tree;
    orc_counters*           m_counters;
    const orc_parent*       m_parent = nullptr;

    bool visit(
        const NodeType&         node,
        const RefShadingRay&    ray,
        const RayInfo3d&        /*ray_info*/,
        double&                 distance)
    {
        const size_t assembly_instance_index = node.get_item_index();
        const size_t assembly_instance_count = node.get_item_count();
        const RefItem* items = m_tree.m_items.data() + assembly_instance_index;

        for (size_t i = 0; i < assembly_instance_count; ++i)
        {
            const RefItem& item = items[i];

            if (!(item.m_vis_flags & ray.m_flags))
                continue;

            ++m_counters->instances_visited;

            RefShadingPoint asm_inst_shading_point;
            // Evaluate the transformation of the assembly instance at ray time (assemblytree.cpp:635-639).
            Transformd scratch;
            const Transformd& assembly_instance_transform =
                item.m_animated ? item.m_transform_sequence.evaluate(ray.m_time_absolute, scratch) : item.m_transform;
            compute_assembly_instance_ray(assembly_instance_transform, item.m_assembly_instance, m_parent, ray, asm_inst_shading_point.m_ray);
            const RayInfo3d asm_inst_ray_info(asm_inst_shading_point.m_ray);

            if (item.m_tree != ~std::uint32_t(0))
            {
                const RefTriangleTree& triangle_tree = *m_tree.m_triangle_trees[item.m_tree];
                TriangleTreeIntersector intersector;
                TriLeafVisitor visitor(triangle_tree, asm_inst_shading_point, m_counters);
                if (triangle_tree.m_moving_triangle_count > 0)
                {
                    intersector.intersect_motion(
                        triangle_tree,
                        asm_inst_shading_point.m_ray,
                        asm_inst_ray_info,
                        asm_inst_shading_point.m_ray.m_time_normalized,
                        visitor);
                }
                else
                {
                    intersector.intersect_no_motion(
                        triangle_tree,
                        asm_inst_shading_point.m_ray,
                        asm_inst_ray_info,
                        visitor);
                }
                visitor.read_hit_triangle_data();
            }

            if (asm_inst_shading_point.m_hit && asm_inst_shading_point.m_ray.m_tmax < m_shading_point.m_ray.m_tmax)
            {
                m_shading_point.m_ray.m_tmax = asm_inst_shading_point.m_ray.m_tmax;
                m_shading_point.m_hit = true;
                m_shading_point.m_bary[0] = asm_inst_shading_point.m_bary[0];
                m_shading_point.m_bary[1] = asm_inst_shading_point.m_bary[1];
                m_shading_point.m_assembly_instance = item.m_assembly_instance;
                m_shading_point.m_object_instance_index = asm_inst_shading_point.m_object_instance_index;
                m_shading_point.m_primitive_index = asm_inst_shading_point.m_primitive_index;
                m_shading_point.m_tri_slot = asm_inst_shading_point.m_tri_slot;
                m_shading_point.m_motion_segment = asm_inst_shading_point.m_motion_segment;
                m_shading_point.m_triangle_support_plane = asm_inst_shading_point.m_triangle_support_plane;     // assemblytree.cpp:743
                m_shading_point.m_assembly_instance_transform = assembly_instance_transform;                     // assemblytree.cpp:738
            }
        }

        distance = m_shading_point.m_ray.m_tmax;
        return true;
    }
};

// AssemblyLeafProbeVisitor (assemblytree.cpp:845-1054), triangle branch.
struct AsmLeafProbeVisitor
{
    const RefAssemblyTree&  m_tree;
    orc_counters*           m_counters;
    bool                    m_hit = false;
    const orc_parent*       m_parent = nullptr;

    bool visit(
        const NodeType&         node,
        const RefShadingRay&    ray,
        const RayInfo3d&        /*ray_info*/,
        double&                 distance)
    {
        const size_t assembly_instance_count = node.get_item_count();
        const RefItem* items = m_tree.m_items.data() + node.get_item_index();

        for (size_t i = 0; i < assembly_instance_count; ++i)
        {
            const RefItem& item = items[i];

            if (!(item.m_vis_flags & ray.m_flags))
                continue;

            ++m_counters->instances_visited;

            RefShadingRay asm_inst_ray;
            Transformd scratch;
            const Transformd& assembly_instance_transform =
                item.m_animated ? item.m_transform_sequence.evaluate(ray.m_time_absolute, scratch) : item.m_transform;
            compute_assembly_instance_ray(assembly_instance_transform, item.m_assembly_instance, m_parent, ray, asm_inst_ray);
            const RayInfo3d asm_inst_ray_info(asm_inst_ray);

            if (item.m_tree != ~std::uint32_t(0))
            {
                const RefTriangleTree& triangle_tree = *m_tree.m_triangle_trees[item.m_tree];
                TriangleTreeProbeIntersector intersector;
                TriLeafProbeVisitor visitor(triangle_tree, asm_inst_ray.m_time_normalized, asm_inst_ray.m_flags, m_counters);
                if (triangle_tree.m_moving_triangle_count > 0)
                {
                    intersector.intersect_motion(
                        triangle_tree,
                        asm_inst_ray,
                        asm_inst_ray_info,
                        asm_inst_ray.m_time_normalized,
                        visitor);
                }
                else
                {
                    intersector.intersect_no_motion(
                        triangle_tree,
                        asm_inst_ray,
                        asm_inst_ray_info,
                        visitor);
                }

                if (visitor.m_hit)
                {
                    m_hit = true;
                    return false;
                }
            }
        }

        distance = ray.m_tmax;
        return true;
    }
};

// The assembly tree is traversed by the GENERIC scalar intersector because its Ray template
// argument is ShadingRay, not Ray3d (assemblytree.h:284-294, bvh_intersector.h:429-434).
typedef bvh::Intersector<RefAssemblyTree, AsmLeafVisitor, RefShadingRay, 64> AssemblyTreeIntersector;
typedef bvh::Intersector<RefAssemblyTree, AsmLeafProbeVisitor, RefShadingRay, 64> AssemblyTreeProbeIntersector;

struct RefScene
{
    orc_scene_desc      m_desc;
    RefAssemblyTree     m_assembly_tree;
};

inline void load_ray(const orc_rays& rays, const size_t i, RefShadingRay& ray)
{
    ray.m_org = Vector3d(rays.org[i * 3], rays.org[i * 3 + 1], rays.org[i * 3 + 2]);
    ray.m_dir = Vector3d(rays.dir[i * 3], rays.dir[i * 3 + 1], rays.dir[i * 3 + 2]);
    ray.m_tmin = rays.tmin[i];
    ray.m_tmax = rays.tmax[i];
    ray.m_time_absolute = rays.time_absolute ? rays.time_absolute[i] : 0.0f;
    ray.m_time_normalized = rays.time_normalized ? rays.time_normalized[i] : 0.0f;
    ray.m_flags = rays.flags ? rays.flags[i] : ~std::uint32_t(0);
}

template <typename F>
void parallel_ranges(const size_t n, int threads, F f)
{
    if (threads < 1) threads = 1;
    if (static_cast<size_t>(threads) > n) threads = n > 0 ? static_cast<int>(n) : 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
    {
        const size_t begin = n * t / threads;
        const size_t end = n * (t + 1) / threads;
        if (threads == 1) f(t, begin, end);
        else pool.emplace_back(f, t, begin, end);
    }
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

// The hit triangle of a leaf slot at a ray time, recomputed the way TriangleLeafVisitor::visit makes
// it (triangletree.cpp:1432-1451); asref_trace_planes returns the one the traversal itself left in
// the ShadingPoint, and the tests demand that the two agree.
static GTriangleType hit_triangle_of_slot(const RefTriangleTree& tree, const std::uint32_t slot, const float time_normalized)
{
    const std::uint32_t motion_segment_count = tree.m_slot_msc[slot];
    if (motion_segment_count == 0) return tree.m_slot_triangles[slot];
    const double base_time = time_normalized * motion_segment_count;
    const size_t base_index = truncate<size_t>(base_time);
    const GVector3* p = tree.m_slot_poses.data() + tree.m_slot_pose[slot] + base_index * 3;
    const GScalar frac = static_cast<GScalar>(base_time - base_index);
    const GScalar one_minus_frac = GScalar(1.0) - frac;
    GVector3 v0 = p[0] * one_minus_frac;
    GVector3 v1 = p[1] * one_minus_frac;
    GVector3 v2 = p[2] * one_minus_frac;
    v0 += p[3] * frac;
    v1 += p[4] * frac;
    v2 += p[5] * frac;
    return GTriangleType(v0, v1, v2);
}

// ShadingPoint::fetch_triangle_source_geometry (shadingpoint.cpp:186-256), vertices only.
static void fetch_source_vertices(const orc_mesh& mesh, const std::uint32_t* t3, const float time_normalized, GVector3& m_v0, GVector3& m_v1, GVector3& m_v2)
{
    const auto vertex = [&mesh](const std::uint32_t i) { return GVector3(mesh.vertices[i * 3], mesh.vertices[i * 3 + 1], mesh.vertices[i * 3 + 2]); };
    const auto get_vertex_pose = [&mesh](const std::uint32_t i, const size_t m)
    {
        const float* p = mesh.vertex_poses + (size_t(i) * mesh.motion_segment_count + m) * 3;
        return GVector3(p[0], p[1], p[2]);
    };
    const size_t motion_segment_count = mesh.vertex_poses ? mesh.motion_segment_count : 0;
    const double base_time = time_normalized * motion_segment_count;
    const size_t base_index = truncate<size_t>(base_time);
    const GScalar frac = static_cast<GScalar>(base_time - base_index);
    const GScalar one_minus_frac = GScalar(1.0) - frac;
    if (motion_segment_count > 0)
    {
        if (base_index == 0)
        {
            m_v0 = vertex(t3[0]);
            m_v1 = vertex(t3[1]);
            m_v2 = vertex(t3[2]);
        }
        else
        {
            m_v0 = get_vertex_pose(t3[0], base_index - 1);
            m_v1 = get_vertex_pose(t3[1], base_index - 1);
            m_v2 = get_vertex_pose(t3[2], base_index - 1);
        }
        m_v0 *= one_minus_frac;
        m_v1 *= one_minus_frac;
        m_v2 *= one_minus_frac;
        m_v0 += get_vertex_pose(t3[0], base_index) * frac;
        m_v1 += get_vertex_pose(t3[1], base_index) * frac;
        m_v2 += get_vertex_pose(t3[2], base_index) * frac;
    }
    else
    {
        m_v0 = vertex(t3[0]);
        m_v1 = vertex(t3[1]);
        m_v2 = vertex(t3[2]);
    }
}

extern "C" {

void* asref_scene_create(const orc_scene_desc* desc)
{
    RefScene* scene = new RefScene();
    scene->m_desc = *desc;
    scene->m_assembly_tree.build(scene->m_desc);
    return scene;
}

void asref_scene_destroy(void* scene)
{
    delete static_cast<RefScene*>(scene);
}

int asref_tree_count(const void* scene)
{
    return static_cast<int>(static_cast<const RefScene*>(scene)->m_assembly_tree.m_triangle_trees.size());
}

int asref_assembly_tree_index(const void* scene, uint32_t assembly)
{
    return static_cast<const RefScene*>(scene)->m_assembly_tree.m_assembly_tree_index[assembly];
}

void asref_get_triangle_tree(const void* scene, int tree, orc_triangle_tree_view* out)
{
    const RefTriangleTree& t = *static_cast<const RefScene*>(scene)->m_assembly_tree.m_triangle_trees[tree];
    out->nodes = t.nodes().empty() ? nullptr : &t.nodes()[0];
    out->node_count = t.nodes().size();
    out->node_bboxes = t.node_bboxes().empty() ? nullptr : &t.node_bboxes()[0][0][0];
    out->node_bbox_count = t.node_bboxes().size();
    out->leaf_data = t.m_leaf_data.empty() ? nullptr : t.m_leaf_data.data();
    out->leaf_data_size = t.m_leaf_data.size();
    out->triangle_keys = t.m_triangle_keys.empty() ? nullptr : t.m_triangle_keys.data();
    out->triangle_key_count = t.m_triangle_keys.size();
    out->static_triangle_count = t.m_static_triangle_count;
    out->moving_triangle_count = t.m_moving_triangle_count;
}

void asref_get_assembly_tree(const void* scene, orc_assembly_tree_view* out)
{
    const RefAssemblyTree& t = static_cast<const RefScene*>(scene)->m_assembly_tree;
    out->nodes = t.nodes().empty() ? nullptr : &t.nodes()[0];
    out->node_count = t.nodes().size();
    out->item_assembly_instance = t.m_item_assembly_instance.data();
    out->item_tree = t.m_item_tree.data();
    out->item_count = t.m_items.size();
}

// Intersector::trace (intersector.cpp:124-189), parent_shading_point == nullptr.
void asref_trace(const void* scene_, const orc_rays* rays, size_t n, orc_hit* out, int threads, orc_counters* counters)
{
    const RefScene& scene = *static_cast<const RefScene*>(scene_);
    std::vector<orc_counters> parts(threads < 1 ? 1 : threads);
    std::memset(parts.data(), 0, parts.size() * sizeof(orc_counters));

    parallel_ranges(n, threads, [&](int tid, size_t begin, size_t end)
    {
        orc_counters local;
        std::memset(&local, 0, sizeof(local));
        for (size_t i = begin; i < end; ++i)
        {
            RefShadingPoint shading_point;
            load_ray(*rays, i, shading_point.m_ray);

            const RayInfo3d ray_info(shading_point.m_ray);

            AssemblyTreeIntersector intersector;
            AsmLeafVisitor visitor{shading_point, scene.m_assembly_tree, &local};
            intersector.intersect_no_motion(scene.m_assembly_tree, shading_point.m_ray, ray_info, visitor);

            orc_hit& hit = out[i];
            hit.t = shading_point.m_ray.m_tmax;
            hit.u = shading_point.m_hit ? shading_point.m_bary[0] : 0.0f;
            hit.v = shading_point.m_hit ? shading_point.m_bary[1] : 0.0f;
            hit.assembly_instance = shading_point.m_hit ? shading_point.m_assembly_instance : ~std::uint32_t(0);
            hit.object_instance_index = shading_point.m_hit ? shading_point.m_object_instance_index : 0;
            hit.primitive_index = shading_point.m_hit ? shading_point.m_primitive_index : 0;
            hit.tri_slot = shading_point.m_hit ? shading_point.m_tri_slot : 0;
            hit.motion_segment = shading_point.m_hit ? shading_point.m_motion_segment : 0;
            hit.prim_type = shading_point.m_hit ? 2 : 0;
            ++local.rays;
            if (shading_point.m_hit) ++local.hits;
        }
        parts[tid] = local;
    });

    accumulate(counters, parts);
}

// Intersector::trace with the support plane the traversal stored in the ShadingPoint
// (m_triangle_support_plane: v0, e0, e1 as doubles; zeros for a miss).
void asref_trace_planes(const void* scene_, const orc_rays* rays, size_t n, orc_hit* out, double* planes, int threads)
{
    const RefScene& scene = *static_cast<const RefScene*>(scene_);
    parallel_ranges(n, threads, [&](int, size_t begin, size_t end)
    {
        orc_counters local;
        std::memset(&local, 0, sizeof(local));
        for (size_t i = begin; i < end; ++i)
        {
            RefShadingPoint shading_point;
            load_ray(*rays, i, shading_point.m_ray);
            const RayInfo3d ray_info(shading_point.m_ray);
            AssemblyTreeIntersector intersector;
            AsmLeafVisitor visitor{shading_point, scene.m_assembly_tree, &local};
            intersector.intersect_no_motion(scene.m_assembly_tree, shading_point.m_ray, ray_info, visitor);
            orc_hit& hit = out[i];
            hit.t = shading_point.m_ray.m_tmax;
            hit.u = shading_point.m_hit ? shading_point.m_bary[0] : 0.0f;
            hit.v = shading_point.m_hit ? shading_point.m_bary[1] : 0.0f;
            hit.assembly_instance = shading_point.m_hit ? shading_point.m_assembly_instance : ~std::uint32_t(0);
            hit.object_instance_index = shading_point.m_hit ? shading_point.m_object_instance_index : 0;
            hit.primitive_index = shading_point.m_hit ? shading_point.m_primitive_index : 0;
            hit.tri_slot = shading_point.m_hit ? shading_point.m_tri_slot : 0;
            hit.motion_segment = shading_point.m_hit ? shading_point.m_motion_segment : 0;
            hit.prim_type = shading_point.m_hit ? 2 : 0;
            double* dst = planes + i * 9;
            for (int k = 0; k < 9; ++k) dst[k] = 0.0;
            if (shading_point.m_hit)
            {
                const TriangleMTSupportPlane<double>& sp = shading_point.m_triangle_support_plane;
                for (int k = 0; k < 3; ++k) { dst[k] = sp.m_v0[k]; dst[3 + k] = sp.m_e0[k]; dst[6 + k] = sp.m_e1[k]; }
            }
        }
    });
}

// The same planes recomputed from the hit records (leaf slot + ray time), as the GPU ABI call
// asgpu_get_support_planes does.
void asref_support_planes(const void* scene_, const orc_rays* rays, const orc_hit* hits, size_t n, double* planes, int threads)
{
    const RefScene& scene = *static_cast<const RefScene*>(scene_);
    const orc_scene_desc& desc = scene.m_desc;
    parallel_ranges(n, threads, [&](int, size_t begin, size_t end)
    {
        for (size_t i = begin; i < end; ++i)
        {
            double* dst = planes + i * 9;
            for (int k = 0; k < 9; ++k) dst[k] = 0.0;
            const orc_hit& hit = hits[i];
            if (hit.prim_type != 2) continue;
            const orc_assembly_instance& inst = desc.assembly_instances[hit.assembly_instance];
            const RefTriangleTree& tree = *scene.m_assembly_tree.m_triangle_trees[scene.m_assembly_tree.m_assembly_tree_index[inst.assembly_index]];
            const TriangleMTSupportPlane<double> sp{TriangleType(hit_triangle_of_slot(tree, hit.tri_slot, rays->time_normalized ? rays->time_normalized[i] : 0.0f))};
            for (int k = 0; k < 3; ++k) { dst[k] = sp.m_v0[k]; dst[3 + k] = sp.m_e0[k]; dst[6 + k] = sp.m_e1[k]; }
        }
    });
}

// Intersector::trace_probe (intersector.cpp:191-238), parent_shading_point == nullptr.
void asref_trace_probe(const void* scene_, const orc_rays* rays, size_t n, uint8_t* out, int threads, orc_counters* counters)
{
    const RefScene& scene = *static_cast<const RefScene*>(scene_);
    std::vector<orc_counters> parts(threads < 1 ? 1 : threads);
    std::memset(parts.data(), 0, parts.size() * sizeof(orc_counters));

    parallel_ranges(n, threads, [&](int tid, size_t begin, size_t end)
    {
        orc_counters local;
        std::memset(&local, 0, sizeof(local));
        for (size_t i = begin; i < end; ++i)
        {
            RefShadingRay ray;
            load_ray(*rays, i, ray);

            const RayInfo3d ray_info(ray);

            AssemblyTreeProbeIntersector intersector;
            AsmLeafProbeVisitor visitor{scene.m_assembly_tree, &local};
            intersector.intersect_no_motion(scene.m_assembly_tree, ray, ray_info, visitor);

            out[i] = visitor.m_hit ? 1 : 0;
            ++local.rays;
            if (visitor.m_hit) ++local.hits;
        }
        parts[tid] = local;
    });

    accumulate(counters, parts);
}

// ShadingPoint::refine_and_offset (shadingpoint.cpp:362-425), triangle branch, calling the
// reference's own refine() / adaptive_offset() (renderer/kernel/intersection/refining.h),
// TriangleMTSupportPlane, compute_triangle_normal, Transform::normal_to_parent and faceforward.
void asref_refine_offset(const void* scene_, const orc_rays* rays, const orc_hit* hits, size_t n, orc_parent* out, int threads)
{
    const RefScene& scene = *static_cast<const RefScene*>(scene_);
    const orc_scene_desc& desc = scene.m_desc;
    parallel_ranges(n, threads, [&](int, size_t begin, size_t end)
    {
        for (size_t i = begin; i < end; ++i)
        {
            orc_parent& p = out[i];
            std::memset(&p, 0, sizeof(p));
            p.assembly_instance = ~std::uint32_t(0);
            const orc_hit& hit = hits[i];
            if (hit.prim_type != 2) continue;
            p.assembly_instance = hit.assembly_instance;

            RefShadingRay m_ray;
            load_ray(*rays, i, m_ray);
            m_ray.m_tmax = hit.t;

            const orc_assembly_instance& inst = desc.assembly_instances[hit.assembly_instance];
            const orc_assembly& assembly = desc.assemblies[inst.assembly_index];
            // m_assembly_instance_transform as the traversal stored it (assemblytree.cpp:738-739): the
            // transform sequence evaluated at the ray's absolute time for an animated instance.
            Transformd assembly_instance_transform = make_transform(inst.local_to_parent, inst.parent_to_local);
            for (const RefItem& item : scene.m_assembly_tree.m_items)
            {
                if (item.m_assembly_instance == hit.assembly_instance && item.m_animated)
                {
                    Transformd scratch;
                    assembly_instance_transform = item.m_transform_sequence.evaluate(m_ray.m_time_absolute, scratch);
                }
            }
            const RefTriangleTree& tree = *scene.m_assembly_tree.m_triangle_trees[scene.m_assembly_tree.m_assembly_tree_index[inst.assembly_index]];

            // m_triangle_support_plane.initialize(TriangleType(*m_hit_triangle)) (triangletree.cpp:1483-1499).
            const TriangleMTSupportPlane<double> support_plane{TriangleType(hit_triangle_of_slot(tree, hit.tri_slot, m_ray.m_time_normalized))};

            // Source geometry (fetch_triangle_source_geometry, shadingpoint.cpp:186-256).
            const orc_object_instance& oi = assembly.object_instances[hit.object_instance_index];
            const orc_mesh& mesh = desc.meshes[oi.mesh_index];
            const std::uint32_t* t3 = mesh.triangles + size_t(hit.primitive_index) * 3;
            GVector3 v0, v1, v2;
            fetch_source_vertices(mesh, t3, m_ray.m_time_normalized, v0, v1, v2);
            const Transformd object_instance_transform = make_transform(oi.local_to_parent, oi.parent_to_local);

            Ray3d refine_space_ray = assembly_instance_transform.to_local(static_cast<const Ray3d&>(m_ray));
            refine_space_ray.m_org += refine_space_ray.m_tmax * refine_space_ray.m_dir;

            const auto intersection_handling = [&support_plane](const Vector3d& pt, const Vector3d& nn)
            {
                return support_plane.intersect(pt, nn);
            };

            refine_space_ray.m_org = renderer::refine(refine_space_ray.m_org, refine_space_ray.m_dir, intersection_handling);

            Vector3d geo_normal = Vector3d(renderer::compute_triangle_normal(v0, v1, v2));
            geo_normal = object_instance_transform.normal_to_parent(geo_normal);
            geo_normal = faceforward(geo_normal, refine_space_ray.m_dir);

            Vector3d front, back;
            renderer::adaptive_offset(refine_space_ray.m_org, geo_normal, front, back, intersection_handling);

            for (int k = 0; k < 3; ++k) { p.front[k] = front[k]; p.back[k] = back[k]; p.geo_normal[k] = geo_normal[k]; }
        }
    });
}

void asref_trace_parents(const void* scene_, const orc_rays* rays, const orc_parent* parents, size_t n, orc_hit* out, int threads)
{
    const RefScene& scene = *static_cast<const RefScene*>(scene_);
    parallel_ranges(n, threads, [&](int, size_t begin, size_t end)
    {
        orc_counters local;
        std::memset(&local, 0, sizeof(local));
        for (size_t i = begin; i < end; ++i)
        {
            RefShadingPoint shading_point;
            load_ray(*rays, i, shading_point.m_ray);
            const RayInfo3d ray_info(shading_point.m_ray);
            AssemblyTreeIntersector intersector;
            AsmLeafVisitor visitor{shading_point, scene.m_assembly_tree, &local,
                                   parents[i].assembly_instance != ~std::uint32_t(0) ? &parents[i] : nullptr};
            intersector.intersect_no_motion(scene.m_assembly_tree, shading_point.m_ray, ray_info, visitor);
            orc_hit& hit = out[i];
            hit.t = shading_point.m_ray.m_tmax;
            hit.u = shading_point.m_hit ? shading_point.m_bary[0] : 0.0f;
            hit.v = shading_point.m_hit ? shading_point.m_bary[1] : 0.0f;
            hit.assembly_instance = shading_point.m_hit ? shading_point.m_assembly_instance : ~std::uint32_t(0);
            hit.object_instance_index = shading_point.m_hit ? shading_point.m_object_instance_index : 0;
            hit.primitive_index = shading_point.m_hit ? shading_point.m_primitive_index : 0;
            hit.tri_slot = shading_point.m_hit ? shading_point.m_tri_slot : 0;
            hit.motion_segment = shading_point.m_hit ? shading_point.m_motion_segment : 0;
            hit.prim_type = shading_point.m_hit ? 2 : 0;
        }
    });
}

void asref_trace_probe_parents(const void* scene_, const orc_rays* rays, const orc_parent* parents, size_t n, uint8_t* out, int threads)
{
    const RefScene& scene = *static_cast<const RefScene*>(scene_);
    parallel_ranges(n, threads, [&](int, size_t begin, size_t end)
    {
        orc_counters local;
        std::memset(&local, 0, sizeof(local));
        for (size_t i = begin; i < end; ++i)
        {
            RefShadingRay ray;
            load_ray(*rays, i, ray);
            const RayInfo3d ray_info(ray);
            AssemblyTreeProbeIntersector intersector;
            AsmLeafProbeVisitor visitor{scene.m_assembly_tree, &local, false,
                                        parents[i].assembly_instance != ~std::uint32_t(0) ? &parents[i] : nullptr};
            intersector.intersect_no_motion(scene.m_assembly_tree, ray, ray_info, visitor);
            out[i] = visitor.m_hit ? 1 : 0;
        }
    });
}

void* asref_scene_create_animated(const orc_scene_desc* desc, const orc_instance_keys* keys)
{
    RefScene* scene = new RefScene();
    scene->m_desc = *desc;
    scene->m_assembly_tree.build(scene->m_desc, keys);
    return scene;
}

void asref_get_item_motion(const void* scene_, uint32_t item_index, orc_item_motion* out)
{
    const RefItem& item = static_cast<const RefScene*>(scene_)->m_assembly_tree.m_items[item_index];
    out->key_times = item.m_key_times.empty() ? nullptr : item.m_key_times.data();
    out->key_parent_to_local = item.m_key_parent_to_local.empty() ? nullptr : item.m_key_parent_to_local.data();
    out->segments = item.m_segments.empty() ? nullptr : item.m_segments.data();
    out->key_count = static_cast<uint32_t>(item.m_key_times.size());
    out->reserved = 0;
}

void asref_get_item_parent_to_local(const void* scene_, uint32_t item_index, double out[16])
{
    const RefItem& item = static_cast<const RefScene*>(scene_)->m_assembly_tree.m_items[item_index];
    for (int e = 0; e < 16; ++e) out[e] = item.m_transform.get_parent_to_local()[e];
}

void asref_set_filter(void* scene_, uint32_t assembly, uint32_t object_instance, const orc_intersection_filter* filter)
{
    RefScene& scene = *static_cast<RefScene*>(scene_);
    const int ti = scene.m_assembly_tree.m_assembly_tree_index[assembly];
    if (ti < 0 || !filter) return;
    RefTriangleTree& tree = *scene.m_assembly_tree.m_triangle_trees[ti];
    const orc_assembly& a = scene.m_desc.assemblies[assembly];
    if (tree.m_intersection_filters.empty()) tree.m_intersection_filters.resize(a.object_instance_count);
    const size_t triangle_count = scene.m_desc.meshes[a.object_instances[object_instance].mesh_index].triangle_count;
    std::unique_ptr<RefIntersectionFilter> f(new RefIntersectionFilter());
    if (filter->object_mask.bits && filter->object_mask.width && filter->object_mask.height)
        f->m_obj_alpha_mask.reset(new RefAlphaMask(filter->object_mask));
    f->m_material_alpha_masks.resize(filter->material_mask_count);
    for (uint32_t i = 0; i < filter->material_mask_count; ++i)
        if (filter->material_masks[i].bits && filter->material_masks[i].width && filter->material_masks[i].height)
            f->m_material_alpha_masks[i].reset(new RefAlphaMask(filter->material_masks[i]));
    f->m_uv.resize(triangle_count * 3, Vector2f(0.0f));
    if (filter->uv)
        for (size_t i = 0; i < triangle_count * 3; ++i) f->m_uv[i] = Vector2f(filter->uv[i * 2], filter->uv[i * 2 + 1]);
    tree.m_intersection_filters[object_instance] = std::move(f);
}

int asref_kat_ray_triangle(
    const double v0[3], const double v1[3], const double v2[3],
    const double org[3], const double dir[3], double tmin, double tmax, double tuv[3])
{
    const TriangleMT<double> triangle(
        Vector3d(v0[0], v0[1], v0[2]), Vector3d(v1[0], v1[1], v1[2]), Vector3d(v2[0], v2[1], v2[2]));
    const Ray3d ray(Vector3d(org[0], org[1], org[2]), Vector3d(dir[0], dir[1], dir[2]), tmin, tmax);
    double t = 0.0, u = 0.0, v = 0.0;
    const bool hit = triangle.intersect(ray, t, u, v);
    tuv[0] = t; tuv[1] = u; tuv[2] = v;
    return hit ? 1 : 0;
}

int asref_kat_ray_triangle_bool(
    const double v0[3], const double v1[3], const double v2[3],
    const double org[3], const double dir[3], double tmin, double tmax)
{
    const TriangleMT<double> triangle(
        Vector3d(v0[0], v0[1], v0[2]), Vector3d(v1[0], v1[1], v1[2]), Vector3d(v2[0], v2[1], v2[2]));
    const Ray3d ray(Vector3d(org[0], org[1], org[2]), Vector3d(dir[0], dir[1], dir[2]), tmin, tmax);
    return triangle.intersect(ray) ? 1 : 0;
}

int asref_kat_ray_aabb(
    const double bmin[3], const double bmax[3], const double org[3], const double dir[3],
    double tmin, double tmax, double* tmin_out)
{
    const AABB3d bbox(Vector3d(bmin[0], bmin[1], bmin[2]), Vector3d(bmax[0], bmax[1], bmax[2]));
    const Ray3d ray(Vector3d(org[0], org[1], org[2]), Vector3d(dir[0], dir[1], dir[2]), tmin, tmax);
    const RayInfo3d ray_info(ray);
    double t = 0.0;
    const bool hit = intersect(ray, ray_info, bbox, t);
    if (tmin_out) *tmin_out = t;
    return hit ? 1 : 0;
}

// The reference's own 3-argument intersect (SSE2 specialisation), 4-argument intersect and clip
// (rayaabb.h), as test_intersection_rayaabb.cpp calls them.  mode / io: see oracle.cpp.
int asref_kat_ray_aabb_ex(int mode, const double bmin[3], const double bmax[3], const double org[3], const double dir[3], double tmin, double tmax, double* io)
{
    const AABB3d bbox(Vector3d(bmin[0], bmin[1], bmin[2]), Vector3d(bmax[0], bmax[1], bmax[2]));
    Ray3d ray(Vector3d(org[0], org[1], org[2]), Vector3d(dir[0], dir[1], dir[2]), tmin, tmax);
    const RayInfo3d ray_info(ray);
    if (mode == 0) return intersect(ray, ray_info, bbox) ? 1 : 0;
    if (mode == 1) return intersect(ray, ray_info, bbox, io[0]) ? 1 : 0;
    ray.m_tmin = io[0]; ray.m_tmax = io[1];
    const bool hit = clip(ray, ray_info, bbox);
    io[0] = ray.m_tmin; io[1] = ray.m_tmax;
    return hit ? 1 : 0;
}

void asref_kat_ray_info(const double dir[3], double rcp[3], uint32_t sgn[3])
{
    const Ray3d ray(Vector3d(0.0), Vector3d(dir[0], dir[1], dir[2]));
    const RayInfo3d ray_info(ray);
    for (int i = 0; i < 3; ++i)
    {
        rcp[i] = ray_info.m_rcp_dir[i];
        sgn[i] = static_cast<uint32_t>(ray_info.m_sgn_dir[i]);
    }
}

// foundation/meta/tests/test_bvh.cpp:62-75 through the reference's own bvh::Node<AABB3d>: stores
// the two child boxes, reads them back, and hands out the node's 128 raw bytes -- the layout the
// product's as_format.h (AsNode) and flattener rely on.
void asref_kat_node_pack(const double left[6], const double right[6], uint32_t child_index, double back[12], unsigned char raw[128])
{
    static_assert(sizeof(bvh::Node<AABB3d>) == 128, "reference node size");
    bvh::Node<AABB3d> node;
    std::memset(&node, 0, sizeof(node));
    node.make_interior();
    node.set_child_node_index(child_index);
    node.set_left_bbox(AABB3d(Vector3d(left[0], left[1], left[2]), Vector3d(left[3], left[4], left[5])));
    node.set_right_bbox(AABB3d(Vector3d(right[0], right[1], right[2]), Vector3d(right[3], right[4], right[5])));
    const AABB3d l = node.get_left_bbox(), r = node.get_right_bbox();
    for (int k = 0; k < 3; ++k)
    {
        back[k] = l.min[k]; back[3 + k] = l.max[k];
        back[6 + k] = r.min[k]; back[9 + k] = r.max[k];
    }
    std::memcpy(raw, &node, 128);
}

// foundation/meta/tests/test_bitmask.cpp:101-128 (StressTest) on the reference's own BitMask2: applies
// `count` set(x, y, value) calls, then reports get(x, y) for every pixel and the object's raw
// storage (m_bits, reached through the object layout: four size_t, then the pointer) -- the bytes an
// in-tree integration hands to asgpu_alpha_mask::bits.
void asref_kat_bitmask(uint32_t width, uint32_t height, const uint32_t* xs, const uint32_t* ys, const unsigned char* values, uint32_t count,
                       unsigned char* got /* width * height */, unsigned char* storage /* ((width + 7) / 8) * height */)
{
    static_assert(sizeof(BitMask2) == 4 * sizeof(size_t) + sizeof(void*), "BitMask2 layout");
    BitMask2 mask(width, height);
    mask.clear();
    for (uint32_t i = 0; i < count; ++i) mask.set(xs[i], ys[i], values[i] != 0);
    for (uint32_t y = 0; y < height; ++y)
        for (uint32_t x = 0; x < width; ++x) got[y * width + x] = mask.get(x, y) ? 1 : 0;
    const std::uint8_t* bits = *reinterpret_cast<const std::uint8_t* const*>(reinterpret_cast<const char*>(&mask) + 4 * sizeof(size_t));
    std::memcpy(storage, bits, mask.get_memory_size() > 0 ? ((width + 7) / 8) * size_t(height) : 0);
}

}   // extern "C"
