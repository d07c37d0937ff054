/*
 * oracle_api.h -- C interface shared by the two CPU checkers in this directory.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under appleseed_b200/ may include, link or
 * call anything in oracle/.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / the
 * CPU baseline.
 *
 * Two implementations export this same interface with different prefixes:
 *   orc_*    oracle.cpp      self-contained restatement of the reference algorithm
 *   asref_*  ref_driver.cpp  the reference's own foundation headers (compiled from
 *                            /root/reference where they lie) + the restated
 *                            renderer/kernel/intersection glue  -> oracle/_ref/
 *
 * The scene description mirrors appleseed's entities on the path:
 *   mesh              <-> StaticTriangleTess        (renderer/kernel/tessellation/statictessellation.h)
 *   object instance   <-> ObjectInstance            (renderer/modeling/scene/objectinstance.h)
 *   assembly          <-> Assembly                  (renderer/modeling/scene/assembly.h)
 *   assembly instance <-> AssemblyInstance with a single-key TransformSequence
 *                         (renderer/modeling/scene/assemblyinstance.h, renderer/utility/transformsequence.h:185-210)
 * Field layout is identical to include/asgpu.h's asgpu_* structs so the same
 * ctypes objects feed the GPU library and the checkers.
 */
#ifndef ORACLE_API_H
#define ORACLE_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_mesh {
    const float*    vertices;               /* vertex_count * 3, object space */
    const uint32_t* triangles;              /* triangle_count * 3 vertex indices */
    const uint16_t* triangle_pa;            /* triangle_count primitive-attribute indices, or NULL */
    const float*    vertex_poses;           /* vertex_count * motion_segment_count * 3, [v * msc + m], or NULL */
    uint32_t        vertex_count;
    uint32_t        triangle_count;
    uint32_t        motion_segment_count;   /* 0 = static mesh */
    uint32_t        reserved;
} orc_mesh;

typedef struct orc_object_instance {
    double          local_to_parent[16];    /* row-major 4x4, object -> assembly space */
    double          parent_to_local[16];
    uint32_t        mesh_index;
    uint32_t        vis_flags;              /* renderer/modeling/scene/visibilityflags.h:50-64 */
} orc_object_instance;

typedef struct orc_assembly {
    const orc_object_instance* object_instances;
    uint32_t        object_instance_count;
    uint32_t        max_leaf_size;                  /* acceleration_structure.max_leaf_size, default 2 */
    float           interior_node_traversal_cost;   /* default 1 */
    float           triangle_intersection_cost;     /* default 1 */
    double          time;                           /* acceleration_structure.time, default 0.5 */
} orc_assembly;

typedef struct orc_assembly_instance {
    double          local_to_parent[16];    /* cumulated assembly-instance transform, assembly -> world */
    double          parent_to_local[16];
    uint32_t        assembly_index;
    uint32_t        vis_flags;
} orc_assembly_instance;

typedef struct orc_scene_desc {
    const orc_mesh*              meshes;
    const orc_assembly*          assemblies;
    const orc_assembly_instance* assembly_instances;
    uint32_t        mesh_count;
    uint32_t        assembly_count;
    uint32_t        assembly_instance_count;
    uint32_t        reserved;
} orc_scene_desc;

/* Ray batch: the ShadingRay fields the path reads (renderer/kernel/shading/shadingray.h:99-109). */
typedef struct orc_rays {
    const double*   org;                /* n * 3 */
    const double*   dir;                /* n * 3 */
    const double*   tmin;               /* n */
    const double*   tmax;               /* n */
    const float*    time_absolute;      /* n, or NULL (0) */
    const float*    time_normalized;    /* n, or NULL (0) */
    const uint32_t* flags;              /* n, or NULL (AllRays) */
} orc_rays;

/* Hit record: ShadingPoint primary block (renderer/kernel/shading/shadingpoint.h:289-302). */
typedef struct orc_hit {
    double          t;                      /* m_ray.m_tmax after the trace */
    float           u, v;                   /* m_bary */
    uint32_t        assembly_instance;      /* index into assembly_instances; 0xFFFFFFFF = miss */
    uint32_t        object_instance_index;  /* m_object_instance_index */
    uint32_t        primitive_index;        /* m_primitive_index (triangle index in its mesh) */
    uint32_t        tri_slot;               /* leaf-order slot in the triangle tree (key / support plane) */
    uint32_t        motion_segment;         /* base pose index used for a moving triangle, else 0 */
    uint32_t        prim_type;              /* 0 = none, 2 = triangle (ShadingPoint::PrimitiveType) */
} orc_hit;

/* Read-only views of the reference-format trees (what an in-tree flattener would see). */
typedef struct orc_triangle_tree_view {
    const void*     nodes;              /* bvh::Node<AABB3d>, 128 B each (foundation/math/bvh/bvh_node.h:100-107) */
    const double*   node_bboxes;        /* Tree::m_node_bboxes, 6 doubles each, swizzled minx maxx miny maxy minz maxz */
    const uint8_t*  leaf_data;          /* TriangleTree::m_leaf_data */
    const void*     triangle_keys;      /* TriangleKey, 12 B each */
    uint64_t        node_count;
    uint64_t        node_bbox_count;
    uint64_t        leaf_data_size;
    uint64_t        triangle_key_count;
    uint64_t        static_triangle_count;
    uint64_t        moving_triangle_count;
} orc_triangle_tree_view;

typedef struct orc_assembly_tree_view {
    const void*     nodes;              /* bvh::Node<AABB3d>; leaves: item index/count in the header */
    const uint32_t* item_assembly_instance;  /* per item (tree order): index into assembly_instances */
    const uint32_t* item_tree;               /* per item: triangle tree index, 0xFFFFFFFF = none */
    uint64_t        node_count;
    uint64_t        item_count;
} orc_assembly_tree_view;

/* Per-batch traversal counters, mirroring bvh::TraversalStatistics (bvh_statistics.h:84-93). */
typedef struct orc_counters {
    uint64_t        rays;
    uint64_t        assembly_nodes_visited;
    uint64_t        instances_visited;
    uint64_t        triangle_nodes_visited;
    uint64_t        triangles_tested;
    uint64_t        hits;
} orc_counters;

/* What Intersector::trace reads of the PARENT ShadingPoint (intersector.cpp:145-149,
 * assemblytree.cpp:556-596): the assembly instance that holds the previous hit and the refined,
 * offset points of ShadingPoint::refine_and_offset (shadingpoint.cpp:362-466, triangle branch,
 * RENDERER_ADAPTIVE_OFFSET) in that instance's space. */
typedef struct orc_parent {
    uint32_t        assembly_instance;      /* 0xFFFFFFFF = no parent (parent_shading_point == nullptr) */
    uint32_t        reserved;
    double          front[3];               /* m_refine_space_front_point */
    double          back[3];                /* m_refine_space_back_point */
    double          geo_normal[3];          /* m_refine_space_geo_normal (face-forwarded, not unit length) */
} orc_parent;

/* renderer::IntersectionFilter of one object instance (renderer/kernel/intersection/
 * intersectionfilter.h): alpha masks are foundation::BitMask2 images
 * (bits[y * ((width + 7) / 8) + x / 8] >> (x & 7)), uv = m_uv (three Vector2f per triangle).
 * Applied by the closest-hit leaf visitor only (triangletree.cpp:1404-1411, 1455-1462). */
typedef struct orc_alpha_mask {
    const uint8_t*  bits;                   /* NULL = no mask */
    uint32_t        width, height;
} orc_alpha_mask;

typedef struct orc_intersection_filter {
    orc_alpha_mask          object_mask;            /* m_obj_alpha_mask */
    const orc_alpha_mask*   material_masks;         /* m_material_alpha_masks, indexed by TriangleKey::get_triangle_pa() */
    uint32_t                material_mask_count;
    uint32_t                reserved;
    const float*            uv;                     /* triangle_count * 6 floats */
} orc_intersection_filter;

/* Animated assembly instances (asref only: the reference's own renderer/utility/
 * transformsequence.cpp is compiled into oracle/_ref).  keys: TransformSequence keys in ascending
 * time order; key_count < 2 = not animated. */
typedef struct orc_instance_keys {
    const float*    times;              /* key_count */
    const double*   local_to_parent;    /* key_count * 16 */
    const double*   parent_to_local;    /* key_count * 16 */
    uint32_t        key_count;
    uint32_t        reserved;
} orc_instance_keys;

/* What TransformSequence::prepare() keeps per segment in its TransformInterpolator
 * (foundation/math/transform.h:640-655): scale, rotation (s, v.x, v.y, v.z), translation at both ends. */
typedef struct orc_transform_segment {
    double          s0[3], q0[4], t0[3], s1[3], q1[4], t1[3];
} orc_transform_segment;

typedef struct orc_item_motion {
    const float*                    key_times;
    const double*                   key_parent_to_local;    /* key_count * 16 */
    const orc_transform_segment*    segments;               /* key_count - 1 */
    uint32_t                        key_count;
    uint32_t                        reserved;
} orc_item_motion;

/* bvh::Node<AABB3d> of the reference: store two child boxes, read them back, copy the raw node. */
void    asref_kat_node_pack(const double left[6], const double right[6], uint32_t child_index, double back[12], unsigned char raw[128]);
/* foundation::BitMask2 of the reference after `count` set(x, y, value) calls: get() of every pixel and the raw storage. */
void    asref_kat_bitmask(uint32_t width, uint32_t height, const uint32_t* xs, const uint32_t* ys, const unsigned char* values, uint32_t count,
                          unsigned char* got, unsigned char* storage);
void*   asref_scene_create_animated(const orc_scene_desc* desc, const orc_instance_keys* keys /* per assembly instance */);
void    asref_get_item_motion(const void* scene, uint32_t item /* tree order */, orc_item_motion* out);
void    asref_get_item_parent_to_local(const void* scene, uint32_t item, double out[16]);
/* Intersector::trace returning, next to the hit records, ShadingPoint::m_triangle_support_plane as
 * the traversal stored it (triangletree.cpp:1483-1499; nine doubles v0, e0, e1; zeros for a miss). */
void    asref_trace_planes(const void* scene, const orc_rays* rays, size_t n, orc_hit* out, double* planes, int threads);

#define ORC_DECLARE(prefix)                                                                         \
    void*   prefix##_scene_create(const orc_scene_desc* desc);                                      \
    void    prefix##_scene_destroy(void* scene);                                                    \
    int     prefix##_tree_count(const void* scene);                                                 \
    int     prefix##_assembly_tree_index(const void* scene, uint32_t assembly);                     \
    void    prefix##_get_triangle_tree(const void* scene, int tree, orc_triangle_tree_view* out);   \
    void    prefix##_get_assembly_tree(const void* scene, orc_assembly_tree_view* out);             \
    void    prefix##_trace(const void* scene, const orc_rays* rays, size_t n, orc_hit* out,         \
                           int threads, orc_counters* counters);                                    \
    void    prefix##_trace_probe(const void* scene, const orc_rays* rays, size_t n, uint8_t* out,   \
                                 int threads, orc_counters* counters);                              \
    /* rays = the rays that produced `hits` (world space); static triangles only. */                \
    void    prefix##_refine_offset(const void* scene, const orc_rays* rays, const orc_hit* hits,    \
                                   size_t n, orc_parent* out, int threads);                         \
    /* trace / trace_probe with a parent shading point per ray. */                                  \
    void    prefix##_support_planes(const void* scene, const orc_rays* rays, const orc_hit* hits,   \
                                    size_t n, double* planes, int threads);                         \
    void    prefix##_trace_parents(const void* scene, const orc_rays* rays,                         \
                                   const orc_parent* parents, size_t n, orc_hit* out, int threads); \
    void    prefix##_trace_probe_parents(const void* scene, const orc_rays* rays,                   \
                                   const orc_parent* parents, size_t n, uint8_t* out, int threads); \
    /* Attach (a copy of) an intersection filter to one object instance of one assembly. */        \
    void    prefix##_set_filter(void* scene, uint32_t assembly, uint32_t object_instance,           \
                                const orc_intersection_filter* filter);

ORC_DECLARE(orc)
ORC_DECLARE(asref)

/* Checker helper (orc only): the two smallest candidate distances over ALL triangles the ray
 * intersects in [tmin, tmax) -- used for the north-star tie rule (|t1 - t2| <= 1e-6 * t1 exempts
 * identity comparison).  t2 = +inf when there is at most one candidate. */
void orc_two_nearest(const void* scene, const orc_rays* rays, size_t n, double* t1, double* t2, int threads);

/* Known-answer-test entry points for the primitive tests of the reference
 * (foundation/meta/tests/test_intersection_raytriangle.cpp, test_intersection_rayaabb.cpp, test_ray.cpp). */
#define ORC_DECLARE_KAT(prefix)                                                                     \
    int     prefix##_kat_ray_triangle(const double v0[3], const double v1[3], const double v2[3],   \
                                      const double org[3], const double dir[3], double tmin,        \
                                      double tmax, double tuv[3]);                                  \
    int     prefix##_kat_ray_triangle_bool(const double v0[3], const double v1[3],                  \
                                      const double v2[3], const double org[3],                      \
                                      const double dir[3], double tmin, double tmax);               \
    int     prefix##_kat_ray_aabb_ex(int mode, const double bmin[3], const double bmax[3],          \
                                     const double org[3], const double dir[3], double tmin,         \
                                     double tmax, double* io);                                      \
    int     prefix##_kat_ray_aabb(const double bmin[3], const double bmax[3], const double org[3],  \
                                  const double dir[3], double tmin, double tmax, double* tmin_out); \
    void    prefix##_kat_ray_info(const double dir[3], double rcp[3], uint32_t sgn[3]);

ORC_DECLARE_KAT(orc)
ORC_DECLARE_KAT(asref)

#ifdef __cplusplus
}
#endif

#endif /* ORACLE_API_H */
