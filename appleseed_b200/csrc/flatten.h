//
// flatten.h -- host flattener: reference-format trees -> GPU blob (gpu_layout.h).
//
#pragma once

#include "../../include/asgpu.h"
#include "gpu_layout.h"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <new>
#include <string>
#include <vector>

namespace asgpu
{

// Growable host buffer of the blob image.  Unlike std::vector<uint8_t> it does not zero-fill what
// it grows into (the sections are about to be overwritten, by several threads), only the alignment
// gaps between sections, so that equal scenes give equal bytes.
class HostBlob
{
  public:
    HostBlob() {}
    ~HostBlob() { std::free(m_data); }
    HostBlob(const HostBlob&) = delete;
    HostBlob& operator=(const HostBlob&) = delete;

    uint8_t* data() { return m_data; }
    const uint8_t* data() const { return m_data; }
    size_t size() const { return m_size; }
    void clear() { m_size = 0; }
    void reserve(const size_t capacity)
    {
        if (capacity <= m_capacity) return;
        void* p = std::realloc(m_data, capacity);
        if (!p) throw std::bad_alloc();
        m_data = static_cast<uint8_t*>(p);
        m_capacity = capacity;
    }
    // New size; bytes in [old size, new size) are NOT initialised.
    void grow_to(const size_t size)
    {
        if (size > m_capacity) reserve(std::max(size, m_capacity + m_capacity / 2));
        m_size = size;
    }

  private:
    uint8_t*    m_data = nullptr;
    size_t      m_size = 0, m_capacity = 0;
};

// Builds the blob image in host memory.  Returns ASGPU_OK or a negative error code with `error`
// filled.  `flags` is a combination of ASGPU_SCENE_EXACT / ASGPU_SCENE_WIDE.
int flatten_scene(
    const asgpu_triangle_tree_view* trees,
    uint32_t                        tree_count,
    const asgpu_assembly_tree_view& top,
    const asgpu_source_geometry*    sources,    // tree_count entries or nullptr
    uint32_t                        flags,
    HostBlob&                       blob,
    std::string&                    error);

// Structural validation of a blob (offsets and counts of every table stay inside it).  The
// flattener's own output goes through validate_blob; a blob received from elsewhere (the import
// path, where it lives in device memory) through validate_blob_tables with a reader that copies the
// header and the small tables to the host.
typedef std::function<bool(uint64_t offset, void* dst, size_t bytes)> BlobReader;
int validate_blob_tables(const BlobReader& read, uint64_t size, std::string& error);
int validate_blob(const uint8_t* blob, size_t size, std::string& error);

}   // namespace asgpu
