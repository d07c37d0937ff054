/*
 * asgpu.h -- C ABI of the B200-native ray-intersection engine that stands in for
 * appleseed's renderer::Intersector::trace() / trace_probe() path.
 *
 * Everything is plain C: opaque handles, plain pointers and sizes, no torch / CUDA types in
 * the signatures (streams are passed as void*, i.e. a cudaStream_t).  Every function returns
 * 0 on success or a negative ASGPU_E_* code (handles: NULL on failure); asgpu_last_error()
 * returns a thread-local description.  Nothing throws across this boundary.
 *
 * Reference interfaces replaced (paths relative to src/appleseed/ of appleseedhq/appleseed):
 *
 *   asgpu_trees_build            TraceContext::update -> AssemblyTree::update
 *                                  renderer/kernel/intersection/tracecontext.cpp:84-87,
 *                                  assemblytree.cpp:95-245, triangletree.cpp:397-597
 *   asgpu_scene_create           (new) flattens the reference's bvh::Tree arrays
 *                                  foundation/math/bvh/bvh_tree.h:77-78, bvh_node.h:100-107,
 *                                  triangletree.h:116-125, assemblytree.h:98-134
 *   asgpu_trace                  Intersector::trace        intersector.cpp:124-189
 *   asgpu_trace_probe            Intersector::trace_probe  intersector.cpp:191-238
 *   asgpu_hit                    ShadingPoint primary block shading/shadingpoint.h:289-302,
 *                                  written at assemblytree.cpp:733-744
 *   asgpu_rays                   ShadingRay fields read by the path shading/shadingray.h:99-109
 *   asgpu_get_counters           bvh::TraversalStatistics  foundation/math/bvh/bvh_statistics.h:84-93,
 *                                  Intersector::get_statistics intersector.cpp:396-434
 *   asgpu_scene_export_blob /
 *   asgpu_scene_import_blob      (new) one contiguous device blob = the payload of the single
 *                                  NCCL broadcast that replicates the scene to every GPU
 *   asgpu_get_support_planes     ShadingPoint::m_triangle_support_plane, written by
 *                                  TriangleLeafVisitor::read_hit_triangle_data, triangletree.cpp:1483-1499
 *                                  (TriangleMTSupportPlane<double>, foundation/math/intersection/raytrianglemt.h:84-108);
 *                                  input of Intersector::make_triangle_shading_point, intersector.cpp:240-271
 *   asgpu_refine_and_offset      ShadingPoint::refine_and_offset   shading/shadingpoint.cpp:362-425
 *                                  with fetch_triangle_source_geometry :186-256 (static and deforming meshes)
 */
#ifndef ASGPU_H
#define ASGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASGPU_VERSION 2

/* Error codes. */
#define ASGPU_OK              0
#define ASGPU_E_INVALID      -1     /* bad argument / malformed input */
#define ASGPU_E_CUDA         -2     /* CUDA runtime error (no device, OOM, launch failure) */
#define ASGPU_E_UNSUPPORTED  -3     /* input uses a feature outside the path */
#define ASGPU_E_NOMEM        -4

/* Ray visibility flags: renderer/modeling/scene/visibilityflags.h:50-64. */
#define ASGPU_VIS_CAMERA        (1u << 0)
#define ASGPU_VIS_LIGHT         (1u << 1)
#define ASGPU_VIS_SHADOW        (1u << 2)
#define ASGPU_VIS_TRANSPARENCY  (1u << 3)
#define ASGPU_VIS_PROBE         (1u << 4)
#define ASGPU_VIS_DIFFUSE       (1u << 5)
#define ASGPU_VIS_GLOSSY        (1u << 6)
#define ASGPU_VIS_SPECULAR      (1u << 7)
#define ASGPU_VIS_SUBSURFACE    (1u << 8)
#define ASGPU_VIS_NPR           (1u << 9)
#define ASGPU_VIS_ALL           0xFFFFFFFFu

#define ASGPU_MISS              0xFFFFFFFFu

/* ------------------------------------------------------------------------------------------
 * Scene description (input of the host builder).  Mirrors StaticTriangleTess, ObjectInstance,
 * Assembly (+ its acceleration_structure parameters) and AssemblyInstance with a single-key
 * TransformSequence.  Matrices are row-major 4x4 doubles, both directions supplied as
 * foundation::Transformd stores them (foundation/math/transform.h:135-136).
 * ------------------------------------------------------------------------------------------ */

typedef struct asgpu_mesh {
    const float*    vertices;               /* vertex_count * 3, object space */
    const uint32_t* triangles;              /* triangle_count * 3 vertex indices */
    const uint16_t* triangle_pa;            /* triangle_count primitive-attribute indices, or NULL */
    const float*    vertex_poses;           /* vertex_count * motion_segment_count * 3, [v * msc + m], or NULL */
    uint32_t        vertex_count;
    uint32_t        triangle_count;
    uint32_t        motion_segment_count;   /* 0 = static mesh */
    uint32_t        reserved;
} asgpu_mesh;

typedef struct asgpu_object_instance {
    double          local_to_parent[16];
    double          parent_to_local[16];
    uint32_t        mesh_index;
    uint32_t        vis_flags;
} asgpu_object_instance;

typedef struct asgpu_assembly {
    const asgpu_object_instance* object_instances;
    uint32_t        object_instance_count;
    uint32_t        max_leaf_size;                  /* acceleration_structure.max_leaf_size (default 2) */
    float           interior_node_traversal_cost;   /* default 1 */
    float           triangle_intersection_cost;     /* default 1 */
    double          time;                           /* acceleration_structure.time (default 0.5) */
} asgpu_assembly;

typedef struct asgpu_assembly_instance {
    double          local_to_parent[16];            /* cumulated: assembly space -> world */
    double          parent_to_local[16];
    uint32_t        assembly_index;
    uint32_t        vis_flags;
} asgpu_assembly_instance;

typedef struct asgpu_scene_desc {
    const asgpu_mesh*               meshes;
    const asgpu_assembly*           assemblies;
    const asgpu_assembly_instance*  assembly_instances;
    uint32_t        mesh_count;
    uint32_t        assembly_count;
    uint32_t        assembly_instance_count;
    uint32_t        reserved;
} asgpu_scene_desc;

/* ------------------------------------------------------------------------------------------
 * Reference-format trees: exactly the arrays appleseed's own classes hold, so an in-tree
 * integration passes pointers to its live data (INTEGRATION.md shows the friend accessors).
 * ------------------------------------------------------------------------------------------ */

typedef struct asgpu_triangle_tree_view {
    const void*     nodes;              /* bvh::Node<AABB3d>[node_count], 128 B each, 64 B aligned */
    const double*   node_bboxes;        /* Tree::m_node_bboxes: 6 doubles each, swizzled minx maxx miny maxy minz maxz */
    const uint8_t*  leaf_data;          /* TriangleTree::m_leaf_data */
    const void*     triangle_keys;      /* TriangleKey[triangle_key_count], 12 B each */
    uint64_t        node_count;
    uint64_t        node_bbox_count;
    uint64_t        leaf_data_size;
    uint64_t        triangle_key_count;
    uint64_t        static_triangle_count;
    uint64_t        moving_triangle_count;
} asgpu_triangle_tree_view;

/* One AssemblyTree::Item (assemblytree.h:103-122) with its TransformSequence already evaluated
 * (single key: transformsequence.h:185-210). */
typedef struct asgpu_assembly_item {
    double          parent_to_local[16];    /* world -> assembly-instance space */
    uint32_t        assembly_instance;      /* caller's id, reported back in asgpu_hit */
    uint32_t        triangle_tree;          /* index into the triangle tree array, ASGPU_MISS = none */
    uint32_t        vis_flags;              /* AssemblyInstance::get_vis_flags() */
    uint32_t        reserved;
} asgpu_assembly_item;

/* Animated assembly instance: a TransformSequence with two or more keys
 * (renderer/utility/transformsequence.h).  The traversal evaluates the instance transform at the
 * ray's absolute time (assemblytree.cpp:635-639 -> TransformSequence::evaluate,
 * transformsequence.h:185-210 -> interpolate, transformsequence.cpp:331-356 ->
 * TransformInterpolator::evaluate, foundation/math/transform.h:694-789).  A segment is what
 * TransformSequence::prepare() keeps in its TransformInterpolator (get_s0() ... get_t1()):
 * scale, rotation quaternion (s, v.x, v.y, v.z) and translation at both ends. */
typedef struct asgpu_transform_segment {
    double          s0[3], q0[4], t0[3], s1[3], q1[4], t1[3];
} asgpu_transform_segment;

typedef struct asgpu_item_motion {
    const float*                    key_times;              /* key_count, ascending (m_keys[k].m_time) */
    const double*                   key_parent_to_local;    /* key_count * 16 (m_keys[k].m_transform), used outside the key range */
    const asgpu_transform_segment*  segments;               /* key_count - 1 */
    uint32_t                        key_count;              /* < 2: not animated, asgpu_assembly_item::parent_to_local is used */
    uint32_t                        reserved;
} asgpu_item_motion;

typedef struct asgpu_assembly_tree_view {
    const void*                 nodes;      /* bvh::Node<AABB3d>[node_count]; leaves address items by index/count */
    const asgpu_assembly_item*  items;      /* tree order (AssemblyTree::m_items after reordering) */
    uint64_t        node_count;
    uint64_t        item_count;
    const asgpu_item_motion*    item_motion;    /* item_count entries or NULL: animated instances (the node boxes already bound the motion) */
} asgpu_assembly_tree_view;

/* Optional source geometry, one asgpu_source_geometry per triangle tree: what
 * ShadingPoint::fetch_triangle_source_geometry (renderer/kernel/shading/shadingpoint.cpp:186-256)
 * reads -- StaticTriangleTess::m_vertices / m_primitives, the vertex poses of a deforming mesh
 * (get_vertex_pose, statictessellation.h) and the ObjectInstance transform.  Only needed for
 * asgpu_refine_and_offset. */
typedef struct asgpu_source_object {
    const float*    vertices;           /* vertex_count * 3, object space */
    const void*     triangles;          /* triangle i: three uint32_t vertex indices at triangles + i * triangle_stride */
    uint32_t        vertex_count;
    uint32_t        triangle_count;
    uint32_t        triangle_stride;    /* bytes: 12 for a packed index array, sizeof(renderer::Triangle) for m_primitives */
    uint32_t        motion_segment_count;   /* StaticTriangleTess::get_motion_segment_count(); 0 = static mesh */
    double          parent_to_local[16];    /* ObjectInstance::get_transform().get_parent_to_local() */
    const float*    vertex_poses;       /* vertex_count * motion_segment_count * 3, [v * msc + m] as in asgpu_mesh; NULL for a static mesh */
} asgpu_source_object;

/* renderer::IntersectionFilter of one object instance (renderer/kernel/intersection/
 * intersectionfilter.h:84-128, 169-205): cut-out geometry.  A closest-hit candidate is dropped when
 * its interpolated UV falls on a transparent texel of the object's or its material's alpha mask
 * (triangletree.cpp:1404-1411, 1455-1462); shadow probes ignore filters, as in the reference.
 * Masks are foundation::BitMask2 images: bits[y * ((width + 7) / 8) + x / 8] >> (x & 7). */
typedef struct asgpu_alpha_mask {
    const uint8_t*  bits;               /* AlphaMask::m_bitmask storage, NULL = no mask */
    uint32_t        width, height;
} asgpu_alpha_mask;

typedef struct asgpu_intersection_filter {
    asgpu_alpha_mask        object_mask;            /* m_obj_alpha_mask */
    const asgpu_alpha_mask* material_masks;         /* m_material_alpha_masks, indexed by TriangleKey::get_triangle_pa() */
    uint32_t                material_mask_count;
    uint32_t                reserved;
    const float*            uv;                     /* m_uv: three (u, v) pairs per triangle; NULL = this object has no filter */
} asgpu_intersection_filter;

typedef struct asgpu_source_geometry {
    const asgpu_source_object* objects; /* indexed by object_instance_index (TriangleKey) */
    uint32_t        object_count;
    uint32_t        reserved;
    const asgpu_intersection_filter* filters;   /* TriangleTree::m_intersection_filters: object_count entries, or NULL */
} asgpu_source_geometry;

/* ------------------------------------------------------------------------------------------
 * Host builder: the CPU side that defines the data (sweep-SAH binary BVHs identical to the ones
 * the reference builds).
 * ------------------------------------------------------------------------------------------ */

typedef struct asgpu_trees asgpu_trees;

asgpu_trees*    asgpu_trees_build(const asgpu_scene_desc* desc, int threads);

/* The keys of an ANIMATED assembly instance: its cumulated TransformSequence
 * (assemblytree.cpp:124-127), times strictly ascending as TransformSequence::prepare() leaves them. */
typedef struct asgpu_instance_keys {
    const float*    times;              /* key_count (TransformSequence::m_keys[k].m_time) */
    const double*   local_to_parent;    /* key_count * 16 (m_keys[k].m_transform) */
    const double*   parent_to_local;    /* key_count * 16 */
    uint32_t        key_count;          /* < 2: the instance is not animated, asgpu_assembly_instance's matrices are used */
    uint32_t        reserved;
} asgpu_instance_keys;

/* asgpu_trees_build for scenes with animated assembly instances: `keys` has one entry per
 * assembly instance of the description (NULL = asgpu_trees_build).  The assembly tree is built
 * over the reference's motion bounding boxes (TransformSequence::to_parent,
 * renderer/utility/transformsequence.h:212-236, transformsequence.cpp:509-616) and the items come
 * with the interpolator segments the traversal evaluates (asgpu_assembly_tree_view::item_motion). */
asgpu_trees*    asgpu_trees_build_animated(const asgpu_scene_desc* desc, const asgpu_instance_keys* keys, int threads);

/* Same trees-from-description entry point with the triangle-tree topology built on CUDA device
 * `device`: a linear BVH in Morton order of the triangle centroids
 * (Karras 2012), every node in parallel, instead of the reference's single-threaded sweep SAH
 * (foundation/math/bvh/bvh_sahpartitioner.h:99-170, bvh_builder.h:163-229).  Same node format, same
 * leaf payloads, same motion boxes; the TREE differs from the reference's (so traversal counters
 * and exact-t tie order do), hit records do not.  No host fallback: fails without a device. */
asgpu_trees*    asgpu_trees_build_on_device(const asgpu_scene_desc* desc, int threads, int device);
void            asgpu_trees_destroy(asgpu_trees* trees);
int             asgpu_trees_triangle_tree_count(const asgpu_trees* trees);
int             asgpu_trees_get_triangle_tree(const asgpu_trees* trees, int index, asgpu_triangle_tree_view* out);
int             asgpu_trees_get_assembly_tree(const asgpu_trees* trees, asgpu_assembly_tree_view* out);
double          asgpu_trees_build_seconds(const asgpu_trees* trees);
/* Part of it spent in the device topology build (upload, kernels, download); 0 for asgpu_trees_build. */
double          asgpu_trees_device_seconds(const asgpu_trees* trees);
/* Source geometry of triangle tree `index` (pointers into the handle's own copies). */
int             asgpu_trees_get_source_geometry(const asgpu_trees* trees, int index, asgpu_source_geometry* out);

/* ------------------------------------------------------------------------------------------
 * GPU scene.
 * ------------------------------------------------------------------------------------------ */

typedef struct asgpu_scene asgpu_scene;

/* Scene creation flags. */
#define ASGPU_SCENE_EXACT   (1u << 0)   /* keep the 1:1 binary fp64 layout (bit-exact arbiter kernels) */
#define ASGPU_SCENE_WIDE    (1u << 1)   /* build the wide quantised-box layout (throughput kernels) */
#define ASGPU_SCENE_DEFAULT (ASGPU_SCENE_EXACT | ASGPU_SCENE_WIDE)

/* Flatten reference-format trees into the GPU layouts and upload them to `device`. */
asgpu_scene*    asgpu_scene_create(
                    const asgpu_triangle_tree_view* triangle_trees,
                    uint32_t                        triangle_tree_count,
                    const asgpu_assembly_tree_view* assembly_tree,
                    uint32_t                        flags,
                    int                             device);

/* Same with source geometry (`sources`: triangle_tree_count entries, or NULL = asgpu_scene_create). */
asgpu_scene*    asgpu_scene_create_ex(
                    const asgpu_triangle_tree_view* triangle_trees,
                    uint32_t                        triangle_tree_count,
                    const asgpu_assembly_tree_view* assembly_tree,
                    const asgpu_source_geometry*    sources,
                    uint32_t                        flags,
                    int                             device);

/* Convenience: asgpu_trees_build + asgpu_scene_create_ex (source geometry included). */
asgpu_scene*    asgpu_scene_create_from_desc(const asgpu_scene_desc* desc, uint32_t flags, int device, int threads);

void            asgpu_scene_destroy(asgpu_scene* scene);

/* The whole flattened scene is one contiguous device allocation ("blob") addressed by offsets,
 * so replication to other GPUs is one broadcast of blob_size bytes followed by import_blob on
 * the receiving side (which adopts `blob`, a device pointer on `device`, without copying when
 * `adopt` is non-zero; the caller then keeps the memory alive until asgpu_scene_destroy). */
size_t          asgpu_scene_blob_size(const asgpu_scene* scene);
const void*     asgpu_scene_blob_device_ptr(const asgpu_scene* scene);
int             asgpu_scene_export_blob(const asgpu_scene* scene, void* device_dst, size_t capacity, void* stream);
asgpu_scene*    asgpu_scene_import_blob(const void* device_blob, size_t size, int device, int adopt);

typedef struct asgpu_scene_info {
    uint64_t        blob_bytes;
    uint64_t        triangle_tree_count;
    uint64_t        instance_count;
    uint64_t        triangle_count;             /* leaf slots over all trees */
    uint64_t        moving_triangle_count;
    uint64_t        binary_node_count;          /* exact layout */
    uint64_t        wide_node_count;            /* wide layout */
    uint64_t        binary_node_bytes;
    uint64_t        wide_node_bytes;
    uint64_t        triangle_bytes;             /* per-slot records + pose data */
    uint32_t        flags;
    uint32_t        wide_stack_depth;           /* traversal stack entries the wide layout can need */
} asgpu_scene_info;

int             asgpu_scene_get_info(const asgpu_scene* scene, asgpu_scene_info* out);

/* ------------------------------------------------------------------------------------------
 * Rays and hit records.
 * ------------------------------------------------------------------------------------------ */

/* One array per ShadingRay field the path consumes.  NULL optional arrays mean: time 0,
 * flags = all rays. */
typedef struct asgpu_rays {
    const double*   org;                /* n * 3  (Ray3d::m_org) */
    const double*   dir;                /* n * 3  (Ray3d::m_dir, not necessarily unit length) */
    const double*   tmin;               /* n      (inclusive) */
    const double*   tmax;               /* n      (exclusive) */
    const float*    time_absolute;      /* n or NULL (ShadingRay::Time::m_absolute) */
    const float*    time_normalized;    /* n or NULL (ShadingRay::Time::m_normalized, in [0, 1)) */
    const uint32_t* flags;              /* n or NULL (ShadingRay::m_flags) */
} asgpu_rays;

typedef struct asgpu_hit {              /* 40 bytes */
    double          t;                      /* m_ray.m_tmax after the trace (unchanged on a miss) */
    float           u, v;                   /* m_bary */
    uint32_t        assembly_instance;      /* asgpu_assembly_item::assembly_instance, ASGPU_MISS on a miss */
    uint32_t        object_instance_index;  /* m_object_instance_index */
    uint32_t        primitive_index;        /* m_primitive_index: triangle index in its mesh */
    uint32_t        tri_slot;               /* leaf-order slot in the triangle tree -> TriangleKey, support plane */
    uint32_t        motion_segment;         /* pose interval used for a moving triangle, else 0 */
    uint32_t        prim_type;              /* ShadingPoint::PrimitiveType: 0 none, 2 triangle */
} asgpu_hit;

/* Trace flags. */
#define ASGPU_TRACE_EXACT       (1u << 0)   /* run the 1:1 binary fp64 kernels (reference visit order) */
#define ASGPU_TRACE_COUNTERS    (1u << 1)   /* accumulate traversal counters (slower) */
#define ASGPU_TRACE_SORT        (1u << 2)   /* reorder rays by origin/direction Morton key first */

/* Closest hit for n rays.  All pointers are DEVICE pointers on the scene's device. */
int             asgpu_trace(asgpu_scene* scene, const asgpu_rays* rays, size_t n, asgpu_hit* hits,
                            uint32_t flags, void* stream);

/* Any hit in [tmin, tmax) for n rays: occluded[i] = 1 or 0.  DEVICE pointers. */
int             asgpu_trace_probe(asgpu_scene* scene, const asgpu_rays* rays, size_t n, uint8_t* occluded,
                                  uint32_t flags, void* stream);

/* Same with HOST buffers: the batch is cut into chunks (1 Mi rays) that go through device staging
 * buffers on three streams, so that the H2D copy of one chunk, the kernel of the previous one and
 * the D2H copy of the one before overlap; returns when the results are in `hits` / `occluded`.
 * The copies are issued straight from / into the caller's arrays: they are asynchronous (and the
 * overlap real) when those arrays are page-locked -- asgpu_pin_host below, or any pinned allocation
 * of the caller's; from pageable memory the CUDA driver stages every copy synchronously. */
int             asgpu_trace_host(asgpu_scene* scene, const asgpu_rays* rays, size_t n, asgpu_hit* hits, uint32_t flags);
int             asgpu_trace_probe_host(asgpu_scene* scene, const asgpu_rays* rays, size_t n, uint8_t* occluded, uint32_t flags);

/* Page-locks / unlocks a host range for the calls above (cudaHostRegister / cudaHostUnregister), so
 * that a caller without CUDA headers can pin its ray and hit arenas once. */
int             asgpu_pin_host(void* ptr, size_t bytes);
int             asgpu_unpin_host(void* ptr);

/* The support plane of every hit: planes[i * 9 ...] = v0, e0, e1 of the hit triangle as doubles,
 * exactly what TriangleLeafVisitor::read_hit_triangle_data stores in
 * ShadingPoint::m_triangle_support_plane (triangletree.cpp:1483-1499): the float triangle of the
 * leaf -- for a moving triangle the one interpolated at rays->time_normalized[i] by the closest-hit
 * visitor (:1432-1469) -- widened to double.  Zeros for a miss.  rays = the rays that produced
 * `hits` (only the time is read).  DEVICE pointers.  Needs the exact layout. */
int             asgpu_get_support_planes(asgpu_scene* scene, const asgpu_rays* rays, const asgpu_hit* hits, size_t n,
                                         double* planes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Parent shading points.  Intersector::trace(ray, shading_point,
 * parent) refines and offsets the parent's hit point once (ShadingPoint::refine_and_offset,
 * shadingpoint.cpp:362-466: refine() and adaptive_offset() of renderer/kernel/intersection/
 * refining.h:97-221 against the triangle's support plane) and, inside the assembly instance that
 * holds the parent's hit, starts the child ray from the front or back point
 * (compute_assembly_instance_ray, assemblytree.cpp:556-596; get_offset_point, shadingpoint.h:604-613).
 * Here both halves run on the device, so that bounce and shadow rays never go back to the host.
 * ------------------------------------------------------------------------------------------ */

typedef struct asgpu_parent {           /* 80 bytes */
    uint32_t        assembly_instance;      /* asgpu_hit::assembly_instance of the parent, ASGPU_MISS = no parent */
    uint32_t        reserved;
    double          front[3];               /* m_refine_space_front_point */
    double          back[3];                /* m_refine_space_back_point */
    double          geo_normal[3];          /* m_refine_space_geo_normal (faces the parent ray, not unit length) */
} asgpu_parent;

/* For every hit of a closest-hit trace: the parent record its ShadingPoint would hold.  rays =
 * the rays that produced `hits`.  DEVICE pointers.  Needs source geometry (asgpu_scene_create_ex)
 * and the exact layout.  Moving triangles: the support plane is the triangle interpolated at
 * rays->time_normalized[i] and the geometric normal comes from the source vertices interpolated
 * between the two poses around that time (shadingpoint.cpp:186-256).  Animated assembly instances:
 * the refine space is the instance transform at rays->time_absolute[i]. */
int             asgpu_refine_and_offset(asgpu_scene* scene, const asgpu_rays* rays, const asgpu_hit* hits, size_t n,
                                        asgpu_parent* parents, void* stream);

/* asgpu_trace / asgpu_trace_probe with a parent per ray (parents[i].assembly_instance ==
 * ASGPU_MISS: none).  DEVICE pointers. */
int             asgpu_trace_with_parents(asgpu_scene* scene, const asgpu_rays* rays, const asgpu_parent* parents, size_t n,
                                         asgpu_hit* hits, uint32_t flags, void* stream);
int             asgpu_trace_probe_with_parents(asgpu_scene* scene, const asgpu_rays* rays, const asgpu_parent* parents, size_t n,
                                               uint8_t* occluded, uint32_t flags, void* stream);

/* The coherence sort behind ASGPU_TRACE_SORT on its own: order[i] = index of the ray to process
 * at position i (a permutation of 0..n-1, device array) by ascending 24-bit origin / direction
 * Morton key; keys (optional, device) receives the sorted keys.  The workspace is a stream-ordered
 * allocation of the call (as for ASGPU_TRACE_SORT): any number of sorts may be in flight. */
int             asgpu_sort_rays(asgpu_scene* scene, const asgpu_rays* rays, size_t n, uint32_t* order, uint32_t* keys, void* stream);

typedef struct asgpu_counters {
    uint64_t        rays;
    uint64_t        assembly_nodes_visited;
    uint64_t        instances_visited;
    uint64_t        triangle_nodes_visited;
    uint64_t        triangles_tested;
    uint64_t        hits;
    uint64_t        kernel_launches;        /* launches of this library's kernels since the last reset */
    uint64_t        reserved;
} asgpu_counters;

int             asgpu_get_counters(asgpu_scene* scene, asgpu_counters* out, int reset);
/* The same counters kept apart for closest-hit (asgpu_trace*) and any-hit (asgpu_trace_probe*)
 * launches; either pointer may be NULL.  asgpu_get_counters returns their sum. */
int             asgpu_get_counters_by_kind(asgpu_scene* scene, asgpu_counters* closest, asgpu_counters* probe, int reset);

/* Diagnostic next to the counters (accumulated by ASGPU_TRACE_COUNTERS launches of the throughput
 * kernels, cleared with them): where the lane slots of the warps' node-test rounds go -- one round
 * offers 32 slots; `testing` of them run a node test, the others idle for the reason named.  This
 * is the kernel's SIMD efficiency seen from inside (ncu reports the same thing per instruction). */
typedef struct asgpu_lane_profile {
    uint64_t        rounds;             /* node-test rounds executed by all warps */
    uint64_t        testing;            /* lane slots that tested a node */
    uint64_t        no_ray;             /* lane had no ray (waiting for the next refill) */
    uint64_t        traversed;          /* ray done walking, waits for queued triangle candidates */
    uint64_t        held;               /* must not enter the next instance before its candidates are tested */
    uint64_t        want_enter;         /* waits for the warp's batched instance entry */
    uint64_t        found_leaf;         /* found leaf triangles earlier in this iteration */
    uint64_t        nothing_to_fetch;   /* choosing (pop / back to world space) took this round */
    uint64_t        iterations;         /* passes of the warps' outer loops */
    uint64_t        batched_entries, lanes_entered;
    uint64_t        refills, lanes_refilled;
    uint64_t        reserved[3];
} asgpu_lane_profile;
int             asgpu_get_lane_profile(asgpu_scene* scene, asgpu_lane_profile* closest, asgpu_lane_profile* probe);

/* ------------------------------------------------------------------------------------------
 * Wavefront ray queues: the renderer's recursive per-sample trace
 * loop -- GenericSampleRenderer::render_sample (generic/genericsamplerenderer.cpp:164-299) ->
 * PathTracer::trace (lighting/pathtracer.h:218-...) -> Tracer::trace_between (lighting/tracer.h:
 * 252-259) -- restructured into stages over device-resident queues:
 *
 *     generate -> [closest-hit queue] -> asgpu_trace -> shade / enqueue -> [closest-hit queue']
 *                                                                  \-> [shadow-probe queue] -> asgpu_trace_probe -> accumulate
 *
 * A queue is a ray batch in HBM (the same SoA arrays as asgpu_rays) plus a path id per ray and a
 * device-side ray count, so that a stage can be enqueued without the host knowing how many rays
 * the previous stage produced: no host round trip between wavefronts.
 * ------------------------------------------------------------------------------------------ */

typedef struct asgpu_ray_queue asgpu_ray_queue;

asgpu_ray_queue* asgpu_queue_create(asgpu_scene* scene, size_t capacity);
void            asgpu_queue_destroy(asgpu_ray_queue* queue);
size_t          asgpu_queue_capacity(const asgpu_ray_queue* queue);
/* Device pointers of the queue's arrays (for a caller's own generate / shade kernels):
 * rays->org ... flags as in asgpu_rays, the two time arrays included; path_ids =
 * uint32_t[capacity]; count = uint64_t in device memory.  A producer that counts past `capacity`
 * only loses the rays beyond it: every consumer clamps the count. */
int             asgpu_queue_device_arrays(asgpu_ray_queue* queue, asgpu_rays* rays, uint32_t** path_ids, uint64_t** count);
int             asgpu_queue_reset(asgpu_ray_queue* queue, void* stream);                    /* count = 0 */
int             asgpu_queue_count(asgpu_ray_queue* queue, void* stream, uint64_t* count);   /* synchronises `stream` */
/* Appends n rays from HOST arrays (path_ids may be NULL: ids = position; NULL time arrays: time 0). */
int             asgpu_queue_push_host(asgpu_ray_queue* queue, const asgpu_rays* rays, const uint32_t* path_ids, size_t n, void* stream);
/* Trace everything in the queue (the ray count is read on the device).  hits / occluded: DEVICE
 * arrays of at least `capacity` entries.  ASGPU_TRACE_SORT is refused (ASGPU_E_UNSUPPORTED): the
 * sort needs the ray count on the host. */
int             asgpu_trace_queue(asgpu_scene* scene, asgpu_ray_queue* queue, asgpu_hit* hits, uint32_t flags, void* stream);
int             asgpu_trace_probe_queue(asgpu_scene* scene, asgpu_ray_queue* queue, uint8_t* occluded, uint32_t flags, void* stream);

/* The synthetic path stream of BASELINE.json configs[4]: per pixel
 * sample one pinhole camera ray (PinholeCamera::spawn_ray, renderer/modeling/camera/
 * pinholecamera.cpp:159-195), then up to max_bounces cosine-weighted bounces
 * (sample_hemisphere_cosine, foundation/math/sampling/mappings.h:299-314) about the geometric
 * normal, and at every path vertex one shadow probe to one of the point lights with
 * tmax = distance * (1 - 1e-6) (Tracer::trace_between, renderer/kernel/lighting/tracer.h:252-259).
 * Random numbers are a counter-based hash of (seed, pixel, sample, depth): the result does not
 * depend on batching, queue order or the number of GPUs. */
/* Child rays start AT the hit point and carry their parent shading point, refined and offset on the
 * device (asgpu_refine_and_offset + the parent origin rule): what the reference's path tracer does
 * by passing parent_shading_point to Intersector::trace.  Without it they start offset_eps along
 * the geometric normal (the parent == nullptr convention).  Needs source geometry. */
#define ASGPU_STREAM_PARENTS    (1u << 0)

typedef struct asgpu_path_stream_desc {
    uint32_t        width, height;          /* image resolution */
    uint32_t        spp;                    /* camera paths per pixel */
    uint32_t        max_bounces;            /* bounces after the primary ray (3 for C5) */
    uint32_t        tile_size;              /* 32: Frame's default tile size (frame.cpp:1331) */
    uint32_t        light_count;            /* 1..8 */
    uint32_t        trace_flags;            /* ASGPU_TRACE_* used for the stream's trace launches */
    uint32_t        stream_flags;           /* ASGPU_STREAM_* */
    uint64_t        seed;
    double          camera_to_world[12];    /* 3 x 4 row-major: rotation | translation */
    double          film_width, film_height, focal_length;
    double          lights[8][3];           /* point light positions, world space */
    double          offset_eps;             /* next-ray origin offset along the geometric normal (parent == nullptr convention) */
    /* Ray time: every camera path draws one normalized time in [0, 1) (a float) and gets
     * ShadingRay::Time::create_with_normalized_time(t, shutter_open, shutter_close)
     * (shading/shadingray.h:230-239: absolute = lerp(open, close, t)); bounce and shadow rays inherit
     * their path's time (pathtracer.h:764).  open == close == 0: time 0 for every ray. */
    float           shutter_open, shutter_close;
} asgpu_path_stream_desc;

typedef struct asgpu_path_stream_stats {
    uint64_t        camera_rays, bounce_rays, probe_rays;   /* rays traced, by kind */
    uint64_t        surface_hits, escaped, unoccluded;
    uint64_t        wavefronts;             /* closest-hit wavefronts processed */
    uint64_t        kernel_launches;        /* launches of this library's kernels (generate, trace, shade, accumulate) */
} asgpu_path_stream_stats;

typedef struct asgpu_path_stream asgpu_path_stream;

/* queue_capacity: rays per wavefront (rounded down to whole tiles' worth of paths). */
asgpu_path_stream* asgpu_path_stream_create(asgpu_scene* scene, const asgpu_path_stream_desc* desc, size_t queue_capacity);
void            asgpu_path_stream_destroy(asgpu_path_stream* stream);
uint32_t        asgpu_path_stream_tile_count(const asgpu_path_stream* stream);
/* Renders the given tiles (HOST array of tile indices, row-major tile grid) in as many batches as
 * the queue capacity requires; asynchronous on `stream` except for the upload of the tile list. */
int             asgpu_path_stream_render(asgpu_path_stream* stream, const uint32_t* tiles, size_t tile_count, void* cuda_stream);
/* Per pixel accumulators, HOST array of width * height * 4 uint32_t:
 * [0] surface hits, [1] unoccluded shadow probes, [2] escaped rays, [3] sum of hit-identity hashes.  Synchronises. */
int             asgpu_path_stream_read_image(asgpu_path_stream* stream, uint32_t* accum);
/* The same accumulators for the listed tiles only, HOST array of tile_count * tile_size * tile_size * 4
 * uint32_t: tile after tile in list order, row-major inside a tile, zero where an edge tile sticks
 * out of the image.  With the tiles of a frame dealt to several GPUs every process reads back its
 * own tiles (1 / N of the image) instead of the whole frame.  Synchronises. */
int             asgpu_path_stream_read_tiles(asgpu_path_stream* stream, const uint32_t* tiles, size_t tile_count, uint32_t* accum);
int             asgpu_path_stream_clear(asgpu_path_stream* stream);
int             asgpu_path_stream_get_stats(asgpu_path_stream* stream, asgpu_path_stream_stats* out);
/* Test hook: keep a copy of every wavefront's rays and results of the NEXT render call (HOST side,
 * bounded by max_rays); asgpu_path_stream_capture_get returns wavefront k: kind 0 = closest hit
 * (results = asgpu_hit[n]), 1 = shadow probe (results = uint8_t[n]); parents = the parent records
 * the rays carried (ASGPU_STREAM_PARENTS; assembly_instance = ASGPU_MISS otherwise).  Returns the
 * number of rays or a negative error; arrays may be NULL to query the size. */
int             asgpu_path_stream_capture(asgpu_path_stream* stream, size_t max_rays);
int             asgpu_path_stream_capture_count(const asgpu_path_stream* stream);
long long       asgpu_path_stream_capture_get(const asgpu_path_stream* stream, int k, int* kind, uint32_t* depth,
                                              double* org, double* dir, double* tmin, double* tmax, uint32_t* flags,
                                              uint32_t* path_ids, void* results, asgpu_parent* parents);
/* The ray times of captured wavefront k (either array may be NULL). */
long long       asgpu_path_stream_capture_get_times(const asgpu_path_stream* stream, int k, float* time_absolute, float* time_normalized);
/* Device time of the trace launches of the render calls since the last clear, measured with CUDA
 * events on the stream's own launch stream when profiling is switched on (events are recorded
 * around every trace launch; off by default).  closest_ms / probe_ms: sums over the launches;
 * closest_rays / probe_rays: rays those launches traced (read back after the frame). */
typedef struct asgpu_path_stream_profile {
    double          closest_ms, probe_ms, refine_ms, stage_ms;  /* stage = generate + shade + accumulate */
    uint64_t        closest_launches, probe_launches;
} asgpu_path_stream_profile;
int             asgpu_path_stream_set_profiling(asgpu_path_stream* stream, int enabled);
int             asgpu_path_stream_get_profile(asgpu_path_stream* stream, asgpu_path_stream_profile* out);     /* synchronises */

/* Forgets the cached ASGPU_* scheduling knobs (environment variables read once per process):
 * the next launch reads the environment again.  Tuning experiments only. */
void            asgpu_reload_tuning(void);

const char*     asgpu_last_error(void);
int             asgpu_version(void);

#ifdef __cplusplus
}
#endif

#endif /* ASGPU_H */
