//
// gpu_layout.h -- the flattened scene as it lives in HBM.  Shared by the host flattener and the
// CUDA kernels.
//
// The whole scene is ONE contiguous allocation (the "blob"): a BlobHeader at offset 0, then
// 256-byte aligned sections addressed by byte offsets, so that a scene is position independent
// and can be replicated to other GPUs with a single broadcast (asgpu_scene_export/import_blob).
//
// Two layouts of the same trees coexist in the blob:
//
//  * EXACT  -- a 1:1 image of the reference's binary BVHs.  Triangle-tree child boxes are stored
//              as the floats they were built from (the reference widens float boxes to double,
//              bvh_builder.h:193-204, so this is lossless) in 64-byte nodes; the assembly tree
//              keeps the reference's own 128-byte double nodes.  Kernels on this layout repeat the
//              reference's arithmetic and visit order operation for operation.
//  * WIDE   -- 8-wide nodes with 8-bit quantised child boxes (80 bytes) over the same leaves,
//              for the throughput kernels.
//
// Both share the per-slot triangle records (48 bytes: the 36-byte float Moeller-Trumbore triangle
// the reference stores in its leaves + visibility flags + motion info), the pose pool of moving
// triangles, and the 8-byte hit keys.
//
#pragma once

#include <cstdint>

namespace asgpu
{

const uint32_t BlobMagic = 0x42534131u;     // "1ASB"
const uint32_t BlobVersion = 9;
const uint32_t WideStackMax = 64;           // deepest traversal stack any wide kernel variant offers
const uint64_t SectionAlign = 256;

const uint32_t InteriorMark = 0xFFFFFFFFu;
const uint32_t BlobHasFilters = 1u << 8;
const uint32_t BlobHasAnimatedInstances = 1u << 9;     // BlobHeader::flags: some tree carries intersection filters

// EXACT: binary node of a triangle tree, 64 bytes.  box[] = [minL minR maxL maxR] x (x, y, z),
// the order of bvh::Node::m_bbox_data (bvh_node.h:141-162).
struct BNodeF
{
    uint32_t    index;          // interior: first child node; leaf: first triangle slot
    uint32_t    item_count;     // InteriorMark for interior nodes
    float       box[12];
    uint32_t    pad[2];
};
static_assert(sizeof(BNodeF) == 64, "BNodeF");

// EXACT: motion information of an interior node (only for trees with moving triangles).
struct MNode
{
    uint32_t    left_index, left_count, right_index, right_count;   // into the motion box pool
};
static_assert(sizeof(MNode) == 16, "MNode");

// One entry of Tree::m_node_bboxes, kept in the reference's swizzled order
// minx maxx miny maxy minz maxz (triangletree.cpp:725-738) as floats (lossless, see above).
struct MBox { float v[6]; };

// Triangle record, 48 bytes = three 16-byte loads.  The EXACT layout keeps one per slot in the
// reference's leaf order (ref_slot == its own index); the WIDE layout keeps a second array
// ordered so that the leaves of one wide node are contiguous, ref_slot pointing back.
// Static triangle (motion == 0): v0/e0/e1 = TriangleMT<float> exactly as TriangleEncoder stores it.
// Moving triangle (motion != 0): motion - 1 = index of the first float of pose 0 in the pose pool
// ((msc + 1) poses of 9 floats, pose-major) and the bits of v0[0] hold msc.
struct TriRecord
{
    float       v0[3], e0[3], e1[3];
    uint32_t    vis_flags;
    uint32_t    ref_slot;
    uint32_t    motion;
};
static_assert(sizeof(TriRecord) == 48, "TriRecord");

// Hit key (TriangleKey without the primitive-attribute index, which stays on the host).
struct HitKey { uint32_t object_instance_index, triangle_index; };

// WIDE: 8-wide node, 80 bytes.  Child k's box is
//   lo = origin + qlo[axis][k] * 2^exp[axis],  hi = origin + qhi[axis][k] * 2^exp[axis]
// rounded outward.  meta[k]: 0 = empty; internal child: 0x20 | (24 + slot);  leaf child:
// (triangle count as a unary mask << 5) | first-slot offset from tri_base (compressed-wide-BVH
// encoding of Ylitie et al. 2017).
struct WNode
{
    float       origin[3];
    uint8_t     exp[3];
    uint8_t     imask;          // bit k set: child k is an internal node
    uint32_t    child_base;     // index of the first internal child node
    uint32_t    tri_base;       // first triangle slot referenced by this node's leaves
    uint8_t     meta[8];
    uint8_t     qlo[3][8];
    uint8_t     qhi[3][8];
};
static_assert(sizeof(WNode) == 80, "WNode");

// Per triangle tree.
struct TreeDesc
{
    uint64_t    bnodes;         // BNodeF[]        (EXACT)
    uint64_t    mnodes;         // MNode[] or 0    (EXACT, trees with motion)
    uint64_t    mboxes;         // MBox[] or 0
    uint64_t    tris;           // TriRecord[slot_count], leaf order       (EXACT)
    uint64_t    poses;          // float[] or 0
    uint64_t    keys;           // HitKey[slot_count], leaf order
    uint64_t    wnodes;         // WNode[] or 0                            (WIDE)
    uint64_t    wtris;          // TriRecord[slot_count], wide-node order  (WIDE)
    uint32_t    bnode_count;
    uint32_t    wnode_count;
    uint32_t    slot_count;
    uint32_t    moving;         // number of moving triangles
    uint32_t    mbox_count;
    uint32_t    src_object_count;
    uint64_t    src_objects;    // SrcObject[src_object_count] or 0: source geometry for refine_and_offset
    uint64_t    wslices;        // WSlice[wnode_count * wslice_count] or 0: time-sliced child boxes (trees with motion)
    uint32_t    wslice_count;   // T: slice j bounds the children over ray times [j / T, (j + 1) / T]
    uint32_t    filter_count;   // entries of `filters` (0 = the tree has no intersection filters)
    uint64_t    filters;        // FilterRecord[filter_count], indexed by object instance
    uint64_t    key_pa;         // uint16_t[slot_count]: TriangleKey::m_triangle_pa per leaf slot (trees with filters)
};
static_assert(sizeof(TreeDesc) == 128, "TreeDesc");

// Intersection filter of one object instance (intersectionfilter.h): uv == 0 means "no filter".
struct MaskRecord
{
    uint64_t    bits;           // BitMask2 storage, 0 = no mask
    uint32_t    width, height;
};
static_assert(sizeof(MaskRecord) == 16, "MaskRecord");

struct FilterRecord
{
    uint64_t    uv;             // float[triangle_count * 6] or 0
    MaskRecord  object_mask;
    uint64_t    material_masks; // MaskRecord[material_mask_count]
    uint32_t    material_mask_count;
    uint32_t    pad;
};
static_assert(sizeof(FilterRecord) == 40, "FilterRecord");

// WIDE, trees with moving triangles: the quantised child planes of one wide node for one slice of
// the ray-time axis, same frame (origin, exp) and same slots as the node's own qlo / qhi (which
// bound the whole motion).  The reference interpolates motion boxes at the ray time
// (bvh_intersector.h:675-836); a time slice is the cheap equivalent for a quantised wide node:
// no per-plane arithmetic, only a different 48-byte block per ray.
struct WSlice
{
    uint8_t     qlo[3][8];
    uint8_t     qhi[3][8];
};
static_assert(sizeof(WSlice) == 48, "WSlice");

// Source geometry of one object instance of an assembly (what ShadingPoint::
// fetch_triangle_source_geometry reads, shadingpoint.cpp:186-256): object-space vertices, vertex
// indices, the vertex poses of a deforming mesh (StaticTriangleTess::get_vertex_pose) and the rows
// of ObjectInstance's parent_to_local that Transform::normal_to_parent uses (transform.h:446-463).
struct SrcObject
{
    double      parent_to_local[9];     // m[0] m[1] m[2] / m[4] m[5] m[6] / m[8] m[9] m[10]
    uint64_t    vertices;               // float[vertex_count * 3]
    uint64_t    triangles;              // uint32_t[triangle_count * 3]
    uint32_t    vertex_count, triangle_count;
    uint64_t    poses;                  // float[vertex_count * motion_segment_count * 3], [v * msc + m], or 0
    uint32_t    motion_segment_count;   // 0 = static mesh
    uint32_t    pad;
};
static_assert(sizeof(SrcObject) == 112, "SrcObject");

// One assembly-tree item (tree order): world -> instance rows of parent_to_local.
struct ItemRecord
{
    double      m[12];          // 3 x 4, row-major
    uint32_t    tree;           // TreeDesc index or 0xFFFFFFFF
    uint32_t    vis_flags;
    uint32_t    assembly_instance;
    uint32_t    key_count;      // >= 2: animated instance, `motion` holds its keys; else m[] is the transform
    uint64_t    motion;         // animated: float times[key_count] (padded to 16 bytes), double parent_to_local[key_count][12],
                                //           double segments[key_count - 1][20] (asgpu_transform_segment)
    uint32_t    pad[2];
};
static_assert(sizeof(ItemRecord) == 128, "ItemRecord");

// EXACT top level: the reference's own node (bvh_node.h:100-107), double boxes.
struct BNodeD
{
    uint32_t    item_count;
    uint32_t    index;
    uint32_t    unused[6];
    double      box[12];
};
static_assert(sizeof(BNodeD) == 128, "BNodeD");

struct BlobHeader
{
    uint32_t    magic, version;
    uint32_t    flags;              // ASGPU_SCENE_*
    uint32_t    tree_count;
    uint32_t    item_count;
    uint32_t    top_node_count;
    uint32_t    top_wnode_count;
    uint32_t    wide_stack_need;    // entries of traversal stack the wide layout can require
    uint64_t    total_bytes;
    uint64_t    trees;              // TreeDesc[tree_count]
    uint64_t    items;              // ItemRecord[item_count]
    uint64_t    top_nodes;          // BNodeD[top_node_count]
    uint64_t    top_wnodes;         // WNode[] over the items or 0
    uint64_t    top_witems;         // uint32_t[item_count]: wide leaf order -> ItemRecord index
    // statistics
    uint64_t    triangle_count, moving_triangle_count;
    uint64_t    binary_node_count, wide_node_count;
    uint64_t    binary_node_bytes, wide_node_bytes, triangle_bytes;
    uint64_t    pad1[1];
};
static_assert(sizeof(BlobHeader) % 16 == 0, "BlobHeader alignment");

}   // namespace asgpu
