//
// tree_builder.h -- host builder producing reference-format trees (see as_format.h) from an
// asgpu_scene_desc.  Product code: the CPU side of the path that "defines the data"
// (SURVEY.md section 8, row a16).
//
#pragma once

#include "../../include/asgpu.h"
#include "as_format.h"

#include <cstdlib>
#include <memory>
#include <new>
#include <string>
#include <utility>
#include <vector>

namespace asgpu
{

// 64-byte aligned storage for AsNode arrays (the reference aligns its node vector to the L1
// line size, triangletree.cpp:398).
template <typename T>
struct Aligned64Allocator
{
    typedef T value_type;
    Aligned64Allocator() {}
    template <typename U> Aligned64Allocator(const Aligned64Allocator<U>&) {}
    T* allocate(size_t n)
    {
        void* p = nullptr;
        const size_t bytes = n > 0 ? n * sizeof(T) : size_t(64);
        if (posix_memalign(&p, 64, bytes) != 0) throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, size_t) { free(p); }
    // resize(n) default-initialises (no zero fill, no first touch of the pages by one thread): the
    // builder writes every node itself, from several threads.  Construction from arguments is the usual one.
    template <typename U> void construct(U* p) { ::new (static_cast<void*>(p)) U; }
    template <typename U, typename A, typename... Args> void construct(U* p, A&& a, Args&&... args)
    {
        ::new (static_cast<void*>(p)) U(std::forward<A>(a), std::forward<Args>(args)...);
    }
    template <typename U> bool operator==(const Aligned64Allocator<U>&) const { return true; }
    template <typename U> bool operator!=(const Aligned64Allocator<U>&) const { return false; }
};

typedef std::vector<AsNode, Aligned64Allocator<AsNode>> AsNodeVector;

struct HostTriangleTree
{
    AsNodeVector                nodes;
    std::vector<double>         node_bboxes;        // 6 doubles per entry, swizzled
    std::vector<uint8_t>        leaf_data;
    std::vector<AsTriangleKey>  keys;
    uint64_t                    static_triangle_count = 0;
    uint64_t                    moving_triangle_count = 0;
    // Source geometry of the assembly's object instances (views into HostTrees::mesh_*).
    std::vector<asgpu_source_object> source_objects;
};

struct HostAssemblyTree
{
    AsNodeVector                        nodes;
    std::vector<asgpu_assembly_item>    items;      // tree order
    // Animated instances (asgpu_trees_build_animated): per item, in tree order, what the flattener
    // takes as asgpu_item_motion; the pointers of `item_motion` look into the three pools.  Empty
    // when no instance is animated.
    std::vector<asgpu_item_motion>          item_motion;
    std::vector<std::vector<float>>         key_times;
    std::vector<std::vector<double>>        key_parent_to_local;
    std::vector<std::vector<asgpu_transform_segment>> segments;
};

struct HostTrees
{
    std::vector<std::unique_ptr<HostTriangleTree>>  triangle_trees;
    std::vector<int>                                assembly_to_tree;   // -1 = assembly without geometry
    HostAssemblyTree                                assembly_tree;
    std::vector<std::vector<float>>                 mesh_vertices;      // copies of the meshes (source geometry)
    std::vector<std::vector<float>>                 mesh_poses;         // ... and of the vertex poses of deforming meshes
    std::vector<std::vector<uint32_t>>              mesh_triangles;
    double                                          build_seconds = 0.0;
    double                                          topology_seconds = 0.0;    // inside the LbvhTopologyFn calls (device build only)
};

// Topology of a linear BVH over n >= 2 boxes (lbvh_core.h): interior node i of n - 1 covers the
// positions [first[i], last[i]] of `order` (box indices in Morton order); a child reference is an
// interior node index, or a position | LbvhLeafFlag; node_boxes = lo[3], hi[3] per interior node.
struct LbvhTopology
{
    std::vector<uint32_t>   order, left, right, first, last;
    std::vector<float>      node_boxes;
};

// Computes the topology for `n` boxes (lo[3], hi[3] each) inside the root box: lbvh.cu on the
// device in the product, a sequential host run of the same lbvh_core.h in tests/hostsim.
typedef bool (*LbvhTopologyFn)(const float* boxes, size_t n, const float root_lo[3], const float root_hi[3], void* context,
                               LbvhTopology& out, std::string& error);

// The product's LbvhTopologyFn (lbvh.cu): `context` points at an int, the CUDA device ordinal.
bool lbvh_topology_device(const float* boxes, size_t n, const float root_lo[3], const float root_hi[3], void* context,
                          LbvhTopology& out, std::string& error);

// Top of a clustered build: a sweep-SAH tree over the `count` clusters the clustering rounds left
// (boxes lo[3] hi[3], the reference each cluster's subtree is known by, its leaf count).  Nodes are
// numbered breadth first: node i of count - 1, root 0, level l = [level_begin[l], level_begin[l + 1]);
// a child reference is a node of the top (an index below count - 1) or a cluster's own reference.
struct ClusterTop
{
    std::vector<uint32_t>   left, right, leaves, level_begin;
    std::vector<float>      boxes;
};
const uint32_t PlocTopRatio = 8;       // the rounds stop at n / PlocTopRatio clusters (at least 2)
bool build_cluster_top(const float* cbox, const uint32_t* cref, const uint32_t* ccount, size_t count, int threads,
                       ClusterTop& out, std::string& error);

// The product's default LbvhTopologyFn (ploc.cu): parallel locally-ordered clustering over the same
// Morton order -- surface-area driven topology, same output arrays.  ASGPU_PLOC_RADIUS (1..32,
// default 16) sets the search radius.
bool ploc_topology_device(const float* boxes, size_t n, const float root_lo[3], const float root_hi[3], void* context,
                          LbvhTopology& out, std::string& error);
int ploc_radius();

// Returns false and fills `error` on malformed input.  `lbvh` = null: the reference's sweep SAH for
// every tree (result-identical to the reference); else the triangle trees take their topology from
// `lbvh(..., lbvh_context, ...)` (the small assembly tree always uses the sweep).
// `keys`: null, or one entry per assembly instance (animated instances: motion boxes + segments).
bool build_host_trees(const asgpu_scene_desc& desc, int threads, HostTrees& out, std::string& error,
                      LbvhTopologyFn lbvh = nullptr, void* lbvh_context = nullptr, const asgpu_instance_keys* keys = nullptr);

}   // namespace asgpu
