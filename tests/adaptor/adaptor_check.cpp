//
// adaptor_check.cpp -- TEST-ONLY translation unit: compiles the product's C++ adaptor
// (include/asgpu_adaptor.hpp) against the reference's own foundation headers and the tree types of
// oracle/ref_driver.cpp (which derive from foundation::bvh::Tree and are filled by the reference's
// own builder), then checks, ray by ray, that a batch of asgpu_hit records + support planes turned
// back into ShadingPoints by the adaptor equals what the reference traversal leaves in its
// ShadingPoint (primary block, shading/shadingpoint.h:289-302).
//
// The ShadingPoint / AssemblyInstance below are stand-ins with the reference's member names; the
// rest (Transformd, TransformSequence, TriangleMTSupportPlane, Vector2f, bvh::Node) is the
// reference's code.  Built by tests/adaptor/Makefile only where /root/reference exists; the prebuilt
// library travels to the GPU box like oracle/_ref.
//

#include "../../oracle/ref_driver.cpp"          // test infrastructure: _ref tree types, traversal, load_ray
#include "../../include/asgpu_adaptor.hpp"

#include <cstdio>
#include <map>
#include <string>

namespace
{

// renderer::AssemblyInstance, as far as the path goes (assemblyinstance.h: transform_sequence()).
struct TestAssemblyInstance
{
    std::uint32_t                   m_id;
    renderer::TransformSequence     m_transform_sequence;
    const renderer::TransformSequence& transform_sequence() const { return m_transform_sequence; }
};

// renderer::ShadingPoint's primary block, member for member (shadingpoint.h:84-96, 289-302).
struct TestShadingPoint
{
    enum PrimitiveType
    {
        PrimitiveNone               = 0,
        PrimitiveTriangle           = 1UL << 1,
        PrimitiveProceduralSurface  = 1UL << 2
    };

    mutable RefShadingRay                   m_ray;
    PrimitiveType                           m_primitive_type;
    Vector2f                                m_bary;
    const TestAssemblyInstance*             m_assembly_instance;
    Transformd                              m_assembly_instance_transform;
    const renderer::TransformSequence*      m_assembly_instance_transform_seq;
    size_t                                  m_object_instance_index;
    size_t                                  m_primitive_index;
    TriangleMTSupportPlane<double>          m_triangle_support_plane;
    std::uint32_t                           m_members;
};

struct AdaptorViews
{
    std::vector<asgpu_triangle_tree_view>               trees;
    std::vector<asgpu_assembly_item>                    items;
    std::vector<asgpu_adaptor::ItemMotionStorage>       motion_storage;
    std::vector<asgpu_item_motion>                      motions;
    asgpu_assembly_tree_view                            top;
};

static_assert(sizeof(TriangleKey) == 12, "asgpu_triangle_tree_view::triangle_keys are 12-byte TriangleKeys");
static_assert(sizeof(NodeType) == 128, "bvh::Node<AABB3d>");
static_assert(sizeof(AABB3d) == 48, "Tree::m_node_bboxes entries are 6 doubles");

bool same_matrix(const Matrix4d& a, const Matrix4d& b)
{
    return std::memcmp(&a[0], &b[0], 16 * sizeof(double)) == 0;
}

}   // namespace

extern "C" {

// Views of a _ref scene made by the product's GpuSceneFlattener (friend of the tree classes).
void* adaptor_views_create(const void* ref_scene)
{
    const RefScene& scene = *static_cast<const RefScene*>(ref_scene);
    const RefAssemblyTree& tree = scene.m_assembly_tree;
    AdaptorViews* v = new AdaptorViews();
    for (const auto& t : tree.m_triangle_trees)
        v->trees.push_back(asgpu_adaptor::GpuSceneFlattener::triangle_tree_view(*t));
    v->top = asgpu_adaptor::GpuSceneFlattener::assembly_tree_view(tree, v->items,
        [](const RefItem& item, const size_t, asgpu_assembly_item& out)
        {
            // Single-key sequences use this matrix (TransformSequence::evaluate, transformsequence.h:185-210).
            std::memcpy(out.parent_to_local, &item.m_transform.get_parent_to_local()[0], 16 * sizeof(double));
            out.assembly_instance = item.m_assembly_instance;
            out.triangle_tree = item.m_tree;
            out.vis_flags = item.m_vis_flags;
        });
    bool animated = false;
    v->motion_storage.resize(tree.m_items.size());
    v->motions.resize(tree.m_items.size());
    for (size_t i = 0; i < tree.m_items.size(); ++i)
    {
        std::memset(&v->motions[i], 0, sizeof(asgpu_item_motion));
        if (!tree.m_items[i].m_animated) continue;
        v->motions[i] = asgpu_adaptor::GpuSceneFlattener::item_motion<TransformInterpolatord, renderer::TransformSequence, Transformd>(
            tree.m_items[i].m_transform_sequence, v->motion_storage[i]);
        animated = true;
    }
    v->top.item_motion = animated ? v->motions.data() : nullptr;
    return v;
}

void adaptor_views_destroy(void* views) { delete static_cast<AdaptorViews*>(views); }
const asgpu_triangle_tree_view* adaptor_views_trees(const void* views, std::uint32_t* count)
{
    const AdaptorViews* v = static_cast<const AdaptorViews*>(views);
    *count = static_cast<std::uint32_t>(v->trees.size());
    return v->trees.data();
}
const asgpu_assembly_tree_view* adaptor_views_top(const void* views) { return &static_cast<const AdaptorViews*>(views)->top; }

// hits / planes: what asgpu_trace and asgpu_get_support_planes returned for `rays` (host copies).
// Turns them into ShadingPoints with the product's adaptor, runs the reference traversal on the same
// rays, compares the primary blocks.  Returns the number of rays that differ (0 = identical) and
// describes the first difference in `message`.  `exact_identity`: 0 = rays whose hit distance and
// barycentrics agree but whose hit triangle differs are not counted (exact-t ties of the throughput
// kernels); the number of such rays is returned through `ties`.
long long adaptor_check_shading_points(
    const void* ref_scene, const orc_rays* rays, size_t n, const asgpu_hit* hits, const double* planes,
    int exact_identity, long long* ties, char* message, size_t message_size)
{
    const RefScene& scene = *static_cast<const RefScene*>(ref_scene);
    const RefAssemblyTree& tree = scene.m_assembly_tree;

    // AssemblyInstance stand-ins, by the caller's instance id.
    std::map<std::uint32_t, TestAssemblyInstance> instances;
    for (const RefItem& item : tree.m_items)
    {
        TestAssemblyInstance& inst = instances[item.m_assembly_instance];
        inst.m_id = item.m_assembly_instance;
        if (item.m_animated) inst.m_transform_sequence = item.m_transform_sequence;
        else
        {
            inst.m_transform_sequence.set_transform(0.0f, item.m_transform);
            inst.m_transform_sequence.prepare();
        }
    }

    std::vector<RefShadingRay> shading_rays(n);
    for (size_t i = 0; i < n; ++i) load_ray(*rays, i, shading_rays[i]);
    std::vector<TestShadingPoint> points(n);
    asgpu_adaptor::to_shading_points<TestShadingPoint, RefShadingRay, Vector2f, Vector3d, TriangleMTSupportPlane<double>, TestAssemblyInstance, Transformd>(
        shading_rays.data(), hits, planes, n, points.data(),
        [&instances](const std::uint32_t id, const RefShadingRay& ray, const TestAssemblyInstance*& instance, Transformd& transform)
        {
            instance = &instances.at(id);
            Transformd scratch;
            transform = instance->transform_sequence().evaluate(ray.m_time_absolute, scratch);      // assemblytree.cpp:635-639
        });

    long long differing = 0, tie_count = 0;
    if (message && message_size) message[0] = 0;
    for (size_t i = 0; i < n; ++i)
    {
        RefShadingPoint ref;
        load_ray(*rays, i, ref.m_ray);
        const RayInfo3d ray_info(ref.m_ray);
        orc_counters local;
        std::memset(&local, 0, sizeof(local));
        AssemblyTreeIntersector intersector;
        AsmLeafVisitor visitor{ref, tree, &local};
        intersector.intersect_no_motion(tree, ref.m_ray, ray_info, visitor);

        const TestShadingPoint& sp = points[i];
        const char* what = nullptr;
        const bool hit = sp.m_primitive_type == TestShadingPoint::PrimitiveTriangle;
        if (hit != ref.m_hit) what = "hit / miss";
        else if (sp.m_ray.m_tmax != ref.m_ray.m_tmax) what = "m_ray.m_tmax";
        else if (hit)
        {
            const bool same_triangle =
                sp.m_assembly_instance->m_id == ref.m_assembly_instance &&
                sp.m_object_instance_index == ref.m_object_instance_index && sp.m_primitive_index == ref.m_primitive_index;
            if (sp.m_bary[0] != ref.m_bary[0] || sp.m_bary[1] != ref.m_bary[1]) what = "m_bary";
            else if (!same_triangle)
            {
                if (exact_identity) what = "hit identity";
                else { ++tie_count; continue; }
            }
            else if (std::memcmp(&sp.m_triangle_support_plane, &ref.m_triangle_support_plane, sizeof(ref.m_triangle_support_plane)) != 0) what = "m_triangle_support_plane";
            else if (!same_matrix(sp.m_assembly_instance_transform.get_parent_to_local(), ref.m_assembly_instance_transform.get_parent_to_local()) ||
                     !same_matrix(sp.m_assembly_instance_transform.get_local_to_parent(), ref.m_assembly_instance_transform.get_local_to_parent()))
                what = "m_assembly_instance_transform";
            else if (sp.m_assembly_instance_transform_seq != &sp.m_assembly_instance->transform_sequence()) what = "m_assembly_instance_transform_seq";
            else if (sp.m_members != 0) what = "m_members";
        }
        if (what)
        {
            if (differing == 0 && message && message_size)
                std::snprintf(message, message_size, "ray %zu: %s differs (adaptor t = %.17g, reference t = %.17g)", i, what, sp.m_ray.m_tmax, ref.m_ray.m_tmax);
            ++differing;
        }
    }
    if (ties) *ties = tie_count;
    return differing;
}

}   // extern "C"
