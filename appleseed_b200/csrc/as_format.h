//
// as_format.h -- binary layouts of the reference's tree arrays, as consumed by the flattener
// and produced by the host builder.  Layout facts (not code) from the reference:
//
//   bvh::Node<AABB3d>   foundation/math/bvh/bvh_node.h:100-107 -- 6 x u32, pad to 32, 12 doubles
//                       [minL minR maxL maxR] x (x, y, z) (:141-162); a leaf overlays the 96-byte
//                       box area with user data (:256-259).
//   leaf payload        renderer/kernel/intersection/triangleencoder.cpp:72-103, located by
//                       triangletree.cpp:1363-1368 (first u32 == ~0 -> payload follows in-node).
//   TriangleKey         renderer/kernel/intersection/trianglekey.h:63-65 (12 bytes, 2-byte hole).
//
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>

namespace asgpu
{

struct alignas(64) AsNode
{
    uint32_t    item_count;         // 0xFFFFFFFF = interior node
    uint32_t    index;              // interior: first child; leaf: first item
    uint32_t    left_bbox_index;
    uint32_t    left_bbox_count;
    uint32_t    right_bbox_index;
    uint32_t    right_bbox_count;
    uint32_t    pad[2];
    double      bbox[12];           // interior: child boxes; leaf: user data

    bool interior() const { return item_count == 0xFFFFFFFFu; }
    const uint8_t* user_data() const { return reinterpret_cast<const uint8_t*>(bbox); }
    uint8_t* user_data() { return reinterpret_cast<uint8_t*>(bbox); }
};
static_assert(sizeof(AsNode) == 128, "reference node is 128 bytes");
static_assert(offsetof(AsNode, bbox) == 32, "box area starts at byte 32");

const size_t AsNodeUserDataSize = 96;
const size_t AsTriangleBytes = 36;          // TriangleMT<float>: v0, e0, e1
const size_t AsPoseBytes = 36;              // one pose: 3 x GVector3

struct AsTriangleKey
{
    uint32_t    object_instance_index;
    uint16_t    triangle_pa;
    uint16_t    hole;
    uint32_t    triangle_index;
};
static_assert(sizeof(AsTriangleKey) == 12, "TriangleKey is 12 bytes");

}   // namespace asgpu
