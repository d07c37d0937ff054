//
// hostsim.cpp -- TEST-ONLY host build of the engine's per-ray traversal code.
//
// Compiles appleseed_b200/csrc/traverse_core.h with g++ (directed-rounding intrinsics emulated
// through <cfenv>) together with the real host builder and flattener, so that the CPU-only test
// tier (-m "not gpu") can check the traversal LOGIC of both GPU layouts against the oracle.
// This library is never loaded by the product (appleseed_b200/); the CUDA kernels remain the only
// execution path of the engine.
//
#include "../../appleseed_b200/csrc/flatten.h"
#include "../../appleseed_b200/csrc/refine_core.h"
#include "../../appleseed_b200/csrc/lbvh_core.h"
#include "../../appleseed_b200/csrc/ploc_core.h"
#include "../../appleseed_b200/csrc/traverse_core.h"
#include "../../appleseed_b200/csrc/tree_builder.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

using namespace asgpu;

namespace
{
    struct SimScene
    {
        HostBlob                blob;
        SceneView               view;
    };

    std::string g_error;

    template <bool ANY, bool WIDE>
    void run(const SimScene& s, const asgpu_rays& rays, size_t n, asgpu_hit* hits, uint8_t* occluded, uint64_t* counters,
             const asgpu_parent* parents = nullptr)
    {
        std::vector<uint2> stack(WideStackSize);
        Stats stats; std::memset(&stats, 0, sizeof(stats));
        uint64_t found_count = 0;
        for (size_t i = 0; i < n; ++i)
        {
            Ray ray; load_ray(rays, i, ray);
            Hit hit; bool found;
            const uint8_t* parent = parents ? reinterpret_cast<const uint8_t*>(parents + i) : nullptr;
            if (WIDE) found = wide_trace<ANY, true>(s.view, rays, i, ray, hit, stats, stack.data(), 1, parent);
            else found = exact_trace<ANY, true>(s.view, ray, hit, stats, parent);
            found_count += found ? 1 : 0;
            if (ANY) { occluded[i] = found ? 1 : 0; continue; }
            asgpu_hit& h = hits[i];
            std::memset(&h, 0, sizeof(h));
            h.t = ray.tmax;
            h.assembly_instance = ASGPU_MISS;
            if (found)
            {
                const ItemRecord* item = reinterpret_cast<const ItemRecord*>(s.blob.data() + s.view.items) + hit.item;
                const TreeDesc* td = reinterpret_cast<const TreeDesc*>(s.blob.data() + s.view.trees) + item->tree;
                const HitKey* key = reinterpret_cast<const HitKey*>(s.blob.data() + td->keys) + hit.slot;
                h.u = hit.u; h.v = hit.v;
                h.assembly_instance = item->assembly_instance;
                h.object_instance_index = key->object_instance_index;
                h.primitive_index = key->triangle_index;
                h.tri_slot = hit.slot;
                h.motion_segment = hit.segment;
                h.prim_type = 2;
            }
        }
        if (counters)
        {
            counters[0] = n; counters[1] = stats.top_nodes; counters[2] = stats.instances;
            counters[3] = stats.nodes; counters[4] = stats.triangles; counters[5] = found_count;
        }
    }
}

// Sequential host run of lbvh_core.h -- what lbvh.cu computes with one thread per element (keys,
// stable sort by key as the radix sort is, every interior node, boxes bottom-up).
static bool lbvh_topology_host(const float* boxes, size_t n, const float root_lo[3], const float root_hi[3], void*, LbvhTopology& out, std::string& error)
{
    float origin[3], scale[3];
    for (int a = 0; a < 3; ++a)
    {
        const float extent = 2.0f * (root_hi[a] - root_lo[a]);
        origin[a] = 2.0f * root_lo[a];
        scale[a] = extent > 0.0f ? 2097152.0f / extent : 0.0f;
    }
    std::vector<uint64_t> keys(n), sorted(n);
    out.order.resize(n);
    for (size_t i = 0; i < n; ++i) { keys[i] = lbvh_morton(boxes + i * 6, origin, scale); out.order[i] = static_cast<uint32_t>(i); }
    std::stable_sort(out.order.begin(), out.order.end(), [&keys](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    for (size_t i = 0; i < n; ++i) sorted[i] = keys[out.order[i]];
    out.left.resize(n - 1); out.right.resize(n - 1); out.first.resize(n - 1); out.last.resize(n - 1);
    out.node_boxes.assign((n - 1) * 6, 0.0f);
    for (size_t i = 0; i + 1 < n; ++i)
        lbvh_node(sorted.data(), static_cast<int64_t>(n), static_cast<int64_t>(i), out.left[i], out.right[i], out.first[i], out.last[i]);
    // Bottom-up boxes: post-order over the interior nodes.
    std::vector<uint32_t> stack(1, 0u), post;
    while (!stack.empty())
    {
        const uint32_t x = stack.back(); stack.pop_back();
        post.push_back(x);
        if (post.size() > n) { error = "hierarchy has a cycle"; return false; }
        if (!(out.left[x] & LbvhLeafFlag)) stack.push_back(out.left[x]);
        if (!(out.right[x] & LbvhLeafFlag)) stack.push_back(out.right[x]);
    }
    for (size_t k = post.size(); k-- > 0; )
    {
        const uint32_t x = post[k];
        const uint32_t c[2] = { out.left[x], out.right[x] };
        float* dst = out.node_boxes.data() + size_t(x) * 6;
        for (int side = 0; side < 2; ++side)
        {
            const float* src = (c[side] & LbvhLeafFlag) ? boxes + size_t(out.order[c[side] & ~LbvhLeafFlag]) * 6 : out.node_boxes.data() + size_t(c[side]) * 6;
            for (int a = 0; a < 3; ++a)
            {
                dst[a] = side == 0 ? src[a] : std::min(dst[a], src[a]);
                dst[3 + a] = side == 0 ? src[3 + a] : std::max(dst[3 + a], src[3 + a]);
            }
        }
    }
    return true;
}

// Sequential host run of ploc_core.h -- what ploc.cu computes with one thread per cluster: the same
// rounds (nearest / fate / prefix sum / merge) down to n / PlocTopRatio clusters, the same sweep-SAH top over
// them (build_cluster_top, the product's own function), the same node numbering (the rounds from n - 2
// downwards in cluster order, the top breadth first from 0), the same hand-down of the leaf ranges.
static int g_ploc_radius = 16;

static bool ploc_topology_host(const float* boxes, size_t n, const float root_lo[3], const float root_hi[3], void*, LbvhTopology& out, std::string& error)
{
    float origin[3], scale[3];
    for (int a = 0; a < 3; ++a)
    {
        const float extent = 2.0f * (root_hi[a] - root_lo[a]);
        origin[a] = 2.0f * root_lo[a];
        scale[a] = extent > 0.0f ? 2097152.0f / extent : 0.0f;
    }
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> sorted_ids(n);
    for (size_t i = 0; i < n; ++i) { keys[i] = lbvh_morton(boxes + i * 6, origin, scale); sorted_ids[i] = static_cast<uint32_t>(i); }
    std::stable_sort(sorted_ids.begin(), sorted_ids.end(), [&keys](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });

    out.order.assign(n, 0); out.left.assign(n - 1, 0); out.right.assign(n - 1, 0); out.first.assign(n - 1, 0); out.last.assign(n - 1, 0);
    out.node_boxes.assign((n - 1) * 6, 0.0f);
    std::vector<uint32_t> leaves(n - 1, 0);
    std::vector<float> cbox(n * 6), obox;
    std::vector<uint32_t> cref(n), ccount(n, 1), oref, ocount, nearest;
    for (size_t i = 0; i < n; ++i)
    {
        for (int k = 0; k < 6; ++k) cbox[i * 6 + k] = boxes[size_t(sorted_ids[i]) * 6 + k];
        cref[i] = static_cast<uint32_t>(i) | LbvhLeafFlag;
    }
    std::vector<std::pair<uint32_t, uint32_t>> rounds;
    uint32_t clusters = static_cast<uint32_t>(n), next_node = static_cast<uint32_t>(n) - 2;
    // The rounds stop at n / PlocTopRatio clusters; the top over them is the sweep SAH's (build_cluster_top).
    const uint32_t stop_at = std::max<uint32_t>(2u, static_cast<uint32_t>(n / PlocTopRatio));
    while (clusters > stop_at)
    {
        nearest.resize(clusters);
        const float* base = cbox.data();
        auto box_at = [base](const uint32_t k) -> const float* { return base + size_t(k) * 6; };
        for (uint32_t i = 0; i < clusters; ++i) nearest[i] = ploc_nearest(box_at, clusters, i, g_ploc_radius);
        obox.clear(); oref.clear(); ocount.clear();
        uint32_t created = 0;
        for (uint32_t i = 0; i < clusters; ++i)
        {
            const uint64_t f = ploc_fate(nearest.data(), i);
            if (!(f & 1ull)) continue;
            float box[6];
            for (int k = 0; k < 6; ++k) box[k] = cbox[size_t(i) * 6 + k];
            uint32_t ref = cref[i], count = ccount[i];
            if (f >> 32)
            {
                const uint32_t j = nearest[i], node = next_node - created;
                ++created;
                for (int k = 0; k < 3; ++k)
                {
                    box[k] = ploc_min(box[k], cbox[size_t(j) * 6 + k]);
                    box[3 + k] = ploc_max(box[3 + k], cbox[size_t(j) * 6 + 3 + k]);
                }
                out.left[node] = ref; out.right[node] = cref[j];
                count += ccount[j];
                leaves[node] = count;
                for (int k = 0; k < 6; ++k) out.node_boxes[size_t(node) * 6 + k] = box[k];
                ref = node;
            }
            for (int k = 0; k < 6; ++k) obox.push_back(box[k]);
            oref.push_back(ref); ocount.push_back(count);
        }
        if (created == 0) { error = "a clustering round made no progress"; return false; }
        rounds.push_back(std::make_pair(next_node - (created - 1), next_node + 1));
        next_node -= created;
        clusters = static_cast<uint32_t>(oref.size());
        cbox.swap(obox); cref.swap(oref); ccount.swap(ocount);
    }
    // Top of the tree: nodes 0 .. clusters - 2 (the rounds used n - 2 down to clusters - 1).
    uint32_t top_nodes = 0;
    if (clusters > 1)
    {
        ClusterTop top;
        if (!build_cluster_top(cbox.data(), cref.data(), ccount.data(), clusters, 2, top, error)) return false;
        top_nodes = clusters - 1;
        if (next_node != top_nodes - 1) { error = "node numbering of the rounds and of the top do not meet"; return false; }
        for (uint32_t i = 0; i < top_nodes; ++i)
        {
            out.left[i] = top.left[i]; out.right[i] = top.right[i]; leaves[i] = top.leaves[i];
            for (int k = 0; k < 6; ++k) out.node_boxes[size_t(i) * 6 + k] = top.boxes[size_t(i) * 6 + k];
        }
    }
    out.first[0] = 0; out.last[0] = static_cast<uint32_t>(n) - 1;
    for (uint32_t node = 0; node < top_nodes; ++node)
    {
        const uint32_t f = out.first[node], e = out.last[node], l = out.left[node], r = out.right[node];
        const uint32_t left_leaves = (l & LbvhLeafFlag) ? 1u : leaves[l];
        if (l & LbvhLeafFlag) { out.order[f] = sorted_ids[l & ~LbvhLeafFlag]; out.left[node] = f | LbvhLeafFlag; }
        else { out.first[l] = f; out.last[l] = f + left_leaves - 1; }
        if (r & LbvhLeafFlag) { out.order[e] = sorted_ids[r & ~LbvhLeafFlag]; out.right[node] = e | LbvhLeafFlag; }
        else { out.first[r] = f + left_leaves; out.last[r] = e; }
    }
    for (size_t k = rounds.size(); k-- > 0; )
        for (uint32_t node = rounds[k].first; node < rounds[k].second; ++node)
        {
            const uint32_t f = out.first[node], e = out.last[node], l = out.left[node], r = out.right[node];
            const uint32_t left_leaves = (l & LbvhLeafFlag) ? 1u : leaves[l];
            if (l & LbvhLeafFlag) { out.order[f] = sorted_ids[l & ~LbvhLeafFlag]; out.left[node] = f | LbvhLeafFlag; }
            else { out.first[l] = f; out.last[l] = f + left_leaves - 1; }
            if (r & LbvhLeafFlag) { out.order[e] = sorted_ids[r & ~LbvhLeafFlag]; out.right[node] = e | LbvhLeafFlag; }
            else { out.first[r] = f + left_leaves; out.last[r] = e; }
        }
    return true;
}

static bool g_use_lbvh = false;
static bool g_use_ploc = false;

extern "C" {

const char* hostsim_last_error() { return g_error.c_str(); }

void* hostsim_scene_create_filtered(const asgpu_scene_desc* desc, uint32_t flags, int threads, uint32_t filter_count,
                                    const uint32_t* filter_tree, const uint32_t* filter_object, const asgpu_intersection_filter* filters);

void* hostsim_scene_create(const asgpu_scene_desc* desc, uint32_t flags, int threads)
{
    return hostsim_scene_create_filtered(desc, flags, threads, 0, nullptr, nullptr, nullptr);
}

// Same with intersection filters: filters[k] belongs to object instance filter_object[k] of triangle tree filter_tree[k].
void* hostsim_scene_create_filtered(const asgpu_scene_desc* desc, uint32_t flags, int threads, uint32_t filter_count,
                                    const uint32_t* filter_tree, const uint32_t* filter_object, const asgpu_intersection_filter* filters)
{
    HostTrees trees;
    if (!build_host_trees(*desc, threads, trees, g_error, g_use_ploc ? ploc_topology_host : (g_use_lbvh ? lbvh_topology_host : nullptr))) return nullptr;
    std::vector<asgpu_triangle_tree_view> views(trees.triangle_trees.size());
    for (size_t i = 0; i < views.size(); ++i)
    {
        const HostTriangleTree& t = *trees.triangle_trees[i];
        asgpu_triangle_tree_view& v = views[i];
        v.nodes = t.nodes.data();
        v.node_bboxes = t.node_bboxes.empty() ? nullptr : t.node_bboxes.data();
        v.leaf_data = t.leaf_data.empty() ? nullptr : t.leaf_data.data();
        v.triangle_keys = t.keys.empty() ? nullptr : t.keys.data();
        v.node_count = t.nodes.size();
        v.node_bbox_count = t.node_bboxes.size() / 6;
        v.leaf_data_size = t.leaf_data.size();
        v.triangle_key_count = t.keys.size();
        v.static_triangle_count = t.static_triangle_count;
        v.moving_triangle_count = t.moving_triangle_count;
    }
    asgpu_assembly_tree_view top;
    top.nodes = trees.assembly_tree.nodes.data();
    top.items = trees.assembly_tree.items.empty() ? nullptr : trees.assembly_tree.items.data();
    top.node_count = trees.assembly_tree.nodes.size();
    top.item_count = trees.assembly_tree.items.size();
    top.item_motion = nullptr;

    // Source geometry (always) and the filters of the trees that have some.
    std::vector<asgpu_source_geometry> sources(views.size());
    std::vector<std::vector<asgpu_intersection_filter>> per_tree(views.size());
    for (size_t i = 0; i < views.size(); ++i)
    {
        const HostTriangleTree& t = *trees.triangle_trees[i];
        sources[i].objects = t.source_objects.empty() ? nullptr : t.source_objects.data();
        sources[i].object_count = static_cast<uint32_t>(t.source_objects.size());
        sources[i].reserved = 0;
        sources[i].filters = nullptr;
    }
    for (uint32_t k = 0; k < filter_count; ++k)
    {
        std::vector<asgpu_intersection_filter>& v = per_tree[filter_tree[k]];
        if (v.empty())
        {
            asgpu_intersection_filter none; std::memset(&none, 0, sizeof(none));
            v.assign(sources[filter_tree[k]].object_count, none);
        }
        v[filter_object[k]] = filters[k];
        sources[filter_tree[k]].filters = v.data();
    }

    SimScene* s = new SimScene();
    const int rc = flatten_scene(views.empty() ? nullptr : views.data(), static_cast<uint32_t>(views.size()), top,
                                 sources.empty() ? nullptr : sources.data(), flags, s->blob, g_error);
    if (rc != ASGPU_OK) { delete s; return nullptr; }
    if (validate_blob(s->blob.data(), s->blob.size(), g_error) != ASGPU_OK) { delete s; return nullptr; }
    BlobHeader h; std::memcpy(&h, s->blob.data(), sizeof(h));
    s->view.blob = s->blob.data();
    s->view.trees = h.trees; s->view.items = h.items; s->view.top_nodes = h.top_nodes;
    s->view.top_wnodes = h.top_wnodes; s->view.top_witems = h.top_witems;
    s->view.tree_count = h.tree_count; s->view.item_count = h.item_count;
    s->view.top_node_count = h.top_node_count; s->view.top_wnode_count = h.top_wnode_count;
    s->view.wide_stack_need = h.wide_stack_need;
    return s;
}

// Triangle trees as asgpu_trees_build_on_device makes them, with the topology from the host run above.
void* hostsim_scene_create_lbvh(const asgpu_scene_desc* desc, uint32_t flags, int threads)
{
    g_use_lbvh = true;
    void* s = hostsim_scene_create_filtered(desc, flags, threads, 0, nullptr, nullptr, nullptr);
    g_use_lbvh = false;
    return s;
}

// Same with the topology of ploc.cu (search radius as ASGPU_PLOC_RADIUS sets it for the product).
void* hostsim_scene_create_ploc(const asgpu_scene_desc* desc, uint32_t flags, int threads, int radius)
{
    g_use_ploc = true;
    g_ploc_radius = radius < 1 ? 1 : (radius > PlocMaxRadius ? PlocMaxRadius : radius);
    void* s = hostsim_scene_create_filtered(desc, flags, threads, 0, nullptr, nullptr, nullptr);
    g_use_ploc = false;
    return s;
}

// The product's flattener on reference-format trees supplied by the caller (what asgpu_scene_create takes).
// `sources`: view_count entries (what asgpu_scene_create_ex takes) or null.
void* hostsim_scene_create_views(const asgpu_triangle_tree_view* views, uint32_t view_count, const asgpu_assembly_tree_view* top,
                                 const asgpu_source_geometry* sources, uint32_t flags)
{
    SimScene* s = new SimScene();
    const int rc = flatten_scene(views, view_count, *top, sources, flags, s->blob, g_error);
    if (rc != ASGPU_OK) { delete s; return nullptr; }
    if (validate_blob(s->blob.data(), s->blob.size(), g_error) != ASGPU_OK) { delete s; return nullptr; }
    BlobHeader h; std::memcpy(&h, s->blob.data(), sizeof(h));
    s->view.blob = s->blob.data();
    s->view.trees = h.trees; s->view.items = h.items; s->view.top_nodes = h.top_nodes;
    s->view.top_wnodes = h.top_wnodes; s->view.top_witems = h.top_witems;
    s->view.tree_count = h.tree_count; s->view.item_count = h.item_count;
    s->view.top_node_count = h.top_node_count; s->view.top_wnode_count = h.top_wnode_count;
    s->view.wide_stack_need = h.wide_stack_need;
    return s;
}

void hostsim_scene_destroy(void* scene) { delete static_cast<SimScene*>(scene); }

size_t hostsim_blob_size(void* scene) { return static_cast<SimScene*>(scene)->blob.size(); }
const uint8_t* hostsim_blob_data(void* scene) { return static_cast<SimScene*>(scene)->blob.data(); }

void hostsim_trace(void* scene, const asgpu_rays* rays, size_t n, asgpu_hit* hits, int wide, uint64_t* counters)
{
    if (wide) run<false, true>(*static_cast<SimScene*>(scene), *rays, n, hits, nullptr, counters);
    else run<false, false>(*static_cast<SimScene*>(scene), *rays, n, hits, nullptr, counters);
}

// asgpu_trace_with_parents / asgpu_trace_probe_with_parents on the host build.
void hostsim_trace_parents(void* scene, const asgpu_rays* rays, const asgpu_parent* parents, size_t n, asgpu_hit* hits, int wide)
{
    if (wide) run<false, true>(*static_cast<SimScene*>(scene), *rays, n, hits, nullptr, nullptr, parents);
    else run<false, false>(*static_cast<SimScene*>(scene), *rays, n, hits, nullptr, nullptr, parents);
}

void hostsim_trace_probe_parents(void* scene, const asgpu_rays* rays, const asgpu_parent* parents, size_t n, uint8_t* occluded, int wide)
{
    if (wide) run<true, true>(*static_cast<SimScene*>(scene), *rays, n, nullptr, occluded, nullptr, parents);
    else run<true, false>(*static_cast<SimScene*>(scene), *rays, n, nullptr, occluded, nullptr, parents);
}

// asgpu_refine_and_offset on the host build (refine_core.h).
void hostsim_refine_offset(void* scene, const asgpu_rays* rays, const asgpu_hit* hits, size_t n, asgpu_parent* out)
{
    const SimScene& s = *static_cast<SimScene*>(scene);
    const ItemRecord* items = reinterpret_cast<const ItemRecord*>(s.blob.data() + s.view.items);
    for (size_t i = 0; i < n; ++i)
    {
        double* dst = reinterpret_cast<double*>(out + i);
        std::memset(dst, 0, sizeof(asgpu_parent));
        out[i].assembly_instance = ASGPU_MISS;
        if (hits[i].prim_type != 2) continue;
        uint32_t item = ASGPU_MISS;
        for (uint32_t k = 0; k < s.view.item_count; ++k) if (items[k].assembly_instance == hits[i].assembly_instance) { item = k; break; }
        if (item == ASGPU_MISS) continue;
        const double org[3] = { rays->org[i * 3], rays->org[i * 3 + 1], rays->org[i * 3 + 2] };
        const double dir[3] = { rays->dir[i * 3], rays->dir[i * 3 + 1], rays->dir[i * 3 + 2] };
        const float time_absolute = rays->time_absolute ? rays->time_absolute[i] : 0.0f;
        const float time_normalized = rays->time_normalized ? rays->time_normalized[i] : 0.0f;
        refine_offset_one(s.view, org, dir, time_absolute, time_normalized, hits[i].t, item, hits[i].object_instance_index, hits[i].primitive_index, hits[i].tri_slot, dst);
    }
}

// asgpu_get_support_planes on the host build (hit_triangle of traverse_core.h).
void hostsim_support_planes(void* scene, const asgpu_rays* rays, const asgpu_hit* hits, size_t n, double* planes)
{
    const SimScene& s = *static_cast<SimScene*>(scene);
    const ItemRecord* items = reinterpret_cast<const ItemRecord*>(s.blob.data() + s.view.items);
    for (size_t i = 0; i < n; ++i)
    {
        double* dst = planes + i * 9;
        for (int k = 0; k < 9; ++k) dst[k] = 0.0;
        if (hits[i].prim_type != 2) continue;
        uint32_t item = ASGPU_MISS;
        for (uint32_t k = 0; k < s.view.item_count; ++k) if (items[k].assembly_instance == hits[i].assembly_instance) { item = k; break; }
        if (item == ASGPU_MISS) continue;
        TreeDesc td; load_tree_desc(s.view, items[item].tree, td);
        TriD tri;
        hit_triangle(s.blob.data() + td.tris + static_cast<uint64_t>(hits[i].tri_slot) * sizeof(TriRecord), s.blob.data() + td.poses,
                     rays->time_normalized ? rays->time_normalized[i] : 0.0f, tri);
        for (int k = 0; k < 3; ++k) { dst[k] = tri.v0[k]; dst[3 + k] = tri.e0[k]; dst[6 + k] = tri.e1[k]; }
    }
}

void hostsim_trace_probe(void* scene, const asgpu_rays* rays, size_t n, uint8_t* occluded, int wide, uint64_t* counters)
{
    if (wide) run<true, true>(*static_cast<SimScene*>(scene), *rays, n, nullptr, occluded, counters);
    else run<true, false>(*static_cast<SimScene*>(scene), *rays, n, nullptr, occluded, counters);
}

}   // extern "C"
