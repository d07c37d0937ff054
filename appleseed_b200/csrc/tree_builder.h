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
};

struct HostTrees
{
    std::vector<std::unique_ptr<HostTriangleTree>>  triangle_trees;
    std::vector<int>                                assembly_to_tree;   // -1 = assembly without geometry
    HostAssemblyTree                                assembly_tree;
    std::vector<std::vector<float>>                 mesh_vertices;      // copies of the static meshes (source geometry)
    std::vector<std::vector<uint32_t>>              mesh_triangles;
    double                                          build_seconds = 0.0;
};

// Returns false and fills `error` on malformed input.
bool build_host_trees(const asgpu_scene_desc& desc, int threads, HostTrees& out, std::string& error);

}   // namespace asgpu
