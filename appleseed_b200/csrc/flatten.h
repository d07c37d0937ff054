//
// flatten.h -- host flattener: reference-format trees -> GPU blob (gpu_layout.h).
//
#pragma once

#include "../../include/asgpu.h"
#include "gpu_layout.h"

#include <functional>
#include <string>
#include <vector>

namespace asgpu
{

// Builds the blob image in host memory.  Returns ASGPU_OK or a negative error code with `error`
// filled.  `flags` is a combination of ASGPU_SCENE_EXACT / ASGPU_SCENE_WIDE.
int flatten_scene(
    const asgpu_triangle_tree_view* trees,
    uint32_t                        tree_count,
    const asgpu_assembly_tree_view& top,
    const asgpu_source_geometry*    sources,    // tree_count entries or nullptr
    uint32_t                        flags,
    std::vector<uint8_t>&           blob,
    std::string&                    error);

// Structural validation of a blob (offsets and counts of every table stay inside it).  The
// flattener's own output goes through validate_blob; a blob received from elsewhere (the import
// path, where it lives in device memory) through validate_blob_tables with a reader that copies the
// header and the small tables to the host.
typedef std::function<bool(uint64_t offset, void* dst, size_t bytes)> BlobReader;
int validate_blob_tables(const BlobReader& read, uint64_t size, std::string& error);
int validate_blob(const uint8_t* blob, size_t size, std::string& error);

}   // namespace asgpu
