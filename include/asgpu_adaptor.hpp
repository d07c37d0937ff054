//
// asgpu_adaptor.hpp -- the C++ glue between appleseed's own classes and the C ABI of asgpu.h.
//
// Header only, and written against the MEMBER NAMES of the reference's classes rather than against
// the classes themselves, so that the very same code compiles
//   * inside libappleseed, instantiated with renderer::TriangleTree, renderer::AssemblyTree,
//     renderer::ShadingPoint, renderer::ShadingRay (INTEGRATION.md shows the two friend
//     declarations it needs), and
//   * in this repository's test build (tests/adaptor/), instantiated with the tree types of
//     oracle/ref_driver.cpp -- which derive from the reference's own foundation::bvh::Tree and are
//     filled by the reference's own builder -- and with a ShadingPoint stand-in that has the
//     reference's primary block (shading/shadingpoint.h:289-302), member for member.
// Nothing here includes a renderer header; foundation types arrive as template parameters.
//
// Reference interfaces mirrored (paths relative to src/appleseed/):
//   GpuSceneFlattener::triangle_tree_view   foundation/math/bvh/bvh_tree.h:61-78 (m_nodes, m_node_bboxes),
//                                           renderer/kernel/intersection/triangletree.h:116-125
//                                           (m_triangle_keys, m_leaf_data, triangle counts)
//   GpuSceneFlattener::assembly_tree_view   renderer/kernel/intersection/assemblytree.h:98-134 (m_items)
//   GpuSceneFlattener::item_motion          renderer/utility/transformsequence.h:96-99, 185-210,
//                                           transformsequence.cpp:205-244 (what prepare() keeps)
//   make_triangle_shading_point             renderer/kernel/intersection/intersector.cpp:240-271
//   to_shading_points                       the result half of Intersector::trace, intersector.cpp:124-189
//
#ifndef ASGPU_ADAPTOR_HPP
#define ASGPU_ADAPTOR_HPP

#include "asgpu.h"

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

namespace asgpu_adaptor
{

// Animated assembly instance: the arrays asgpu_item_motion points into.
struct ItemMotionStorage
{
    std::vector<float>                      key_times;
    std::vector<double>                     key_parent_to_local;
    std::vector<asgpu_transform_segment>    segments;
};

// Befriended by the tree classes (one `friend class asgpu_adaptor::GpuSceneFlattener;` each): reads
// the protected / private arrays and hands out views, no copy.
class GpuSceneFlattener
{
  public:
    // TriangleTreeT: m_nodes (aligned vector of bvh::Node<AABB3d>), m_node_bboxes (vector of AABB3d),
    // m_leaf_data, m_triangle_keys, m_static_triangle_count, m_moving_triangle_count.
    template <typename TriangleTreeT>
    static asgpu_triangle_tree_view triangle_tree_view(const TriangleTreeT& tree)
    {
        asgpu_triangle_tree_view v;
        std::memset(&v, 0, sizeof(v));
        v.nodes = tree.m_nodes.empty() ? nullptr : &tree.m_nodes[0];
        v.node_count = tree.m_nodes.size();
        // AABB3d = two Vector3d (min, max): 6 doubles, the layout asgpu_triangle_tree_view documents
        // is Tree::m_node_bboxes as stored (the motion boxes were swizzled by the builder,
        // triangletree.cpp:725-738).
        v.node_bboxes = tree.m_node_bboxes.empty() ? nullptr : reinterpret_cast<const double*>(&tree.m_node_bboxes[0]);
        v.node_bbox_count = tree.m_node_bboxes.size();
        v.leaf_data = tree.m_leaf_data.empty() ? nullptr : &tree.m_leaf_data[0];
        v.leaf_data_size = tree.m_leaf_data.size();
        v.triangle_keys = tree.m_triangle_keys.empty() ? nullptr : &tree.m_triangle_keys[0];
        v.triangle_key_count = tree.m_triangle_keys.size();
        v.static_triangle_count = tree.m_static_triangle_count;
        v.moving_triangle_count = tree.m_moving_triangle_count;
        return v;
    }

    // AssemblyTreeT: m_nodes, m_items.  `describe(item, index, out)` fills one asgpu_assembly_item
    // from one AssemblyTree::Item (transform sequence evaluated for a single key, visibility flags,
    // triangle tree index) -- the only part that touches renderer:: entities, so the caller supplies it.
    template <typename AssemblyTreeT, typename DescribeItem>
    static asgpu_assembly_tree_view assembly_tree_view(
        const AssemblyTreeT&                tree,
        std::vector<asgpu_assembly_item>&   items,
        DescribeItem                        describe)
    {
        items.resize(tree.m_items.size());
        for (size_t i = 0; i < items.size(); ++i)
        {
            std::memset(&items[i], 0, sizeof(items[i]));
            describe(tree.m_items[i], i, items[i]);
        }
        asgpu_assembly_tree_view v;
        std::memset(&v, 0, sizeof(v));
        v.nodes = tree.m_nodes.empty() ? nullptr : &tree.m_nodes[0];
        v.node_count = tree.m_nodes.size();
        v.items = items.empty() ? nullptr : &items[0];
        v.item_count = items.size();
        return v;
    }

    // Keys and interpolator segments of a prepare()d TransformSequence with two or more keys.
    // InterpolatorT = foundation::TransformInterpolator<double>.
    template <typename InterpolatorT, typename TransformSequenceT, typename TransformT>
    static asgpu_item_motion item_motion(const TransformSequenceT& sequence, ItemMotionStorage& storage)
    {
        asgpu_item_motion m;
        std::memset(&m, 0, sizeof(m));
        storage.key_times.clear(); storage.key_parent_to_local.clear(); storage.segments.clear();
        const size_t key_count = sequence.size();
        if (key_count < 2) return m;
        TransformT previous;
        for (size_t k = 0; k < key_count; ++k)
        {
            float time;
            TransformT transform;
            sequence.get_transform(k, time, transform);
            storage.key_times.push_back(time);
            const double* p2l = &transform.get_parent_to_local()[0];
            storage.key_parent_to_local.insert(storage.key_parent_to_local.end(), p2l, p2l + 16);
            if (k > 0)
            {
                const InterpolatorT interpolator(previous, transform);
                asgpu_transform_segment sg;
                for (int a = 0; a < 3; ++a)
                {
                    sg.s0[a] = interpolator.get_s0()[a]; sg.s1[a] = interpolator.get_s1()[a];
                    sg.t0[a] = interpolator.get_t0()[a]; sg.t1[a] = interpolator.get_t1()[a];
                    sg.q0[1 + a] = interpolator.get_q0().v[a]; sg.q1[1 + a] = interpolator.get_q1().v[a];
                }
                sg.q0[0] = interpolator.get_q0().s; sg.q1[0] = interpolator.get_q1().s;
                storage.segments.push_back(sg);
            }
            previous = transform;
        }
        m.key_times = &storage.key_times[0];
        m.key_parent_to_local = &storage.key_parent_to_local[0];
        m.segments = &storage.segments[0];
        m.key_count = static_cast<uint32_t>(key_count);
        return m;
    }
};

// Intersector::make_triangle_shading_point (intersector.cpp:240-271), the primary block only: the
// context pointers (m_texture_cache, m_scene) stay with the caller's Intersector.  PrimitiveTriangle
// = ShadingPoint::PrimitiveTriangle (2), the value asgpu_hit::prim_type carries.
template <typename ShadingPointT, typename ShadingRayT, typename Vector2fT, typename AssemblyInstanceT, typename TransformT, typename SupportPlaneT>
inline void make_triangle_shading_point(
    ShadingPointT&              shading_point,
    const ShadingRayT&          shading_ray,
    const Vector2fT&            bary,
    const AssemblyInstanceT*    assembly_instance,
    const TransformT&           assembly_instance_transform,
    const size_t                object_instance_index,
    const size_t                primitive_index,
    const SupportPlaneT&        triangle_support_plane)
{
    shading_point.m_ray = shading_ray;

    // Primary intersection results.
    shading_point.m_primitive_type = ShadingPointT::PrimitiveTriangle;
    shading_point.m_bary = bary;
    shading_point.m_assembly_instance = assembly_instance;
    shading_point.m_assembly_instance_transform = assembly_instance_transform;
    shading_point.m_assembly_instance_transform_seq = &assembly_instance->transform_sequence();
    shading_point.m_object_instance_index = object_instance_index;
    shading_point.m_primitive_index = primitive_index;
    shading_point.m_triangle_support_plane = triangle_support_plane;

    // Available on-demand results: none.
    shading_point.m_members = 0;
}

// One batch of results back into ShadingPoints: what Intersector::trace leaves in its output
// argument.  hits / planes: host copies of asgpu_trace's records and of asgpu_get_support_planes'
// (9 doubles per hit).  `lookup(assembly_instance_id, ray, out_instance, out_transform)` resolves the
// caller's instance id (asgpu_assembly_item::assembly_instance) to the AssemblyInstance and to the
// transform the traversal would have stored (assemblytree.cpp:635-639, 738-739: the sequence
// evaluated at ray.m_time.m_absolute).  A miss leaves shading_points[i] a miss: m_ray = the ray,
// m_primitive_type = PrimitiveNone.  Returns the number of hits.
template <typename ShadingPointT, typename ShadingRayT, typename Vector2fT, typename Vector3dT, typename SupportPlaneT,
          typename AssemblyInstanceT, typename TransformT, typename Lookup>
inline size_t to_shading_points(
    const ShadingRayT*      rays,
    const asgpu_hit*        hits,
    const double*           planes,
    const size_t            count,
    ShadingPointT*          shading_points,
    Lookup                  lookup)
{
    size_t hit_count = 0;
    for (size_t i = 0; i < count; ++i)
    {
        const asgpu_hit& h = hits[i];
        ShadingPointT& sp = shading_points[i];
        if (h.prim_type != static_cast<uint32_t>(ShadingPointT::PrimitiveTriangle))
        {
            sp.m_ray = rays[i];
            sp.m_primitive_type = ShadingPointT::PrimitiveNone;
            sp.m_members = 0;
            continue;
        }
        const AssemblyInstanceT* assembly_instance = nullptr;
        TransformT assembly_instance_transform;
        lookup(h.assembly_instance, rays[i], assembly_instance, assembly_instance_transform);
        const double* p = planes + i * 9;
        SupportPlaneT plane;
        plane.m_v0 = Vector3dT(p[0], p[1], p[2]);
        plane.m_e0 = Vector3dT(p[3], p[4], p[5]);
        plane.m_e1 = Vector3dT(p[6], p[7], p[8]);
        ShadingRayT hit_ray(rays[i]);
        hit_ray.m_tmax = h.t;                   // m_shading_point.m_ray.m_tmax = t (triangletree.cpp:1415)
        make_triangle_shading_point(sp, hit_ray, Vector2fT(h.u, h.v), assembly_instance, assembly_instance_transform,
                                    h.object_instance_index, h.primitive_index, plane);
        ++hit_count;
    }
    return hit_count;
}

}   // namespace asgpu_adaptor

#endif  // ASGPU_ADAPTOR_HPP
