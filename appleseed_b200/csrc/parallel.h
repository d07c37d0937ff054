//
// parallel.h -- the little threading the host side needs (tree_builder.cpp, flatten.cpp): a chunked
// parallel-for over index ranges on std::thread.  Results never depend on the number of threads:
// callers split work into chunks whose outputs land at positions fixed by prefix sums.
//
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <utility>
#include <vector>

namespace asgpu
{

// Threads for host-side set-up work: ASGPU_HOST_THREADS, else the hardware's.
inline int host_threads()
{
    if (const char* e = std::getenv("ASGPU_HOST_THREADS")) { const int v = std::atoi(e); if (v > 0) return std::min(v, 256); }
    const unsigned hw = std::thread::hardware_concurrency();
    return hw == 0 ? 1 : static_cast<int>(std::min(hw, 256u));
}

// Number of chunks parallel_chunks(n, ...) will make: at most `threads`, at least `grain` items each.
inline int chunk_count(const size_t n, const int threads, const size_t grain)
{
    if (n == 0) return 0;
    const size_t by_grain = (n + grain - 1) / grain;
    return static_cast<int>(std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(std::max(1, threads)), by_grain)));
}

inline void chunk_range(const size_t n, const int chunks, const int chunk, size_t& begin, size_t& end)
{
    begin = n * static_cast<size_t>(chunk) / static_cast<size_t>(chunks);
    end = n * static_cast<size_t>(chunk + 1) / static_cast<size_t>(chunks);
}

// f(chunk, begin, end) for every chunk of [0, n), each on its own thread (the caller's for the last).
template <typename F>
inline void parallel_chunks(const size_t n, const int threads, const size_t grain, F f)
{
    const int chunks = chunk_count(n, threads, grain);
    if (chunks == 0) return;
    if (chunks == 1) { f(0, size_t(0), n); return; }
    std::vector<std::thread> pool;
    pool.reserve(chunks - 1);
    for (int c = 0; c + 1 < chunks; ++c)
    {
        size_t b, e; chunk_range(n, chunks, c, b, e);
        pool.emplace_back([=, &f]() { f(c, b, e); });
    }
    size_t b, e; chunk_range(n, chunks, chunks - 1, b, e);
    f(chunks - 1, b, e);
    for (std::thread& t : pool) t.join();
}

// memcpy of a large block on several threads (page-fault and bandwidth bound work both scale).
#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wstringop-overflow"     // gcc 13 mis-sizes the chunk of the single-chunk path
#pragma GCC diagnostic ignored "-Wrestrict"
#endif
inline void parallel_memcpy(void* dst, const void* src, const size_t bytes, const int threads)
{
    if (bytes < (size_t(32) << 20) || threads <= 1) { if (bytes) std::memcpy(dst, src, bytes); return; }
    parallel_chunks(bytes, threads, size_t(16) << 20, [=](int, size_t b, size_t e)
    {
        if (e > b) std::memcpy(static_cast<unsigned char*>(dst) + b, static_cast<const unsigned char*>(src) + b, e - b);
    });
}
#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC diagnostic pop
#endif

// std::vector for trivially copyable element types, minus the value initialisation: resize() leaves
// new elements uninitialised, so that a multi-gigabyte array is first touched (page-faulted) by the
// parallel loop that fills it instead of by one thread zeroing it.
template <typename T>
class RawVector
{
  public:
    RawVector() {}
    explicit RawVector(const size_t n) { resize(n); }
    ~RawVector() { std::free(m_data); }
    RawVector(const RawVector&) = delete;
    RawVector& operator=(const RawVector&) = delete;

    T* data() { return m_data; }
    const T* data() const { return m_data; }
    size_t size() const { return m_size; }
    bool empty() const { return m_size == 0; }
    T& operator[](const size_t i) { return m_data[i]; }
    const T& operator[](const size_t i) const { return m_data[i]; }
    T* begin() { return m_data; }
    T* end() { return m_data + m_size; }
    const T* begin() const { return m_data; }
    const T* end() const { return m_data + m_size; }

    void reserve(const size_t n)
    {
        if (n <= m_capacity) return;
        void* p = std::realloc(m_data, n * sizeof(T));
        if (!p) throw std::bad_alloc();
        m_data = static_cast<T*>(p);
        m_capacity = n;
    }
    void resize(const size_t n) { reserve(n); m_size = n; }         // new elements are NOT initialised
    void push_back(const T& v)
    {
        const T copy = v;           // v may refer to an element of this vector
        if (m_size == m_capacity) reserve(std::max<size_t>(16, m_capacity + m_capacity / 2));
        m_data[m_size++] = copy;
    }
    void clear() { m_size = 0; }
    void release() { std::free(m_data); m_data = nullptr; m_size = m_capacity = 0; }
    void swap(RawVector& other) { std::swap(m_data, other.m_data); std::swap(m_size, other.m_size); std::swap(m_capacity, other.m_capacity); }
    // n copies of `value`, written by several threads.
    void assign(const size_t n, const T& value, const int threads)
    {
        resize(n);
        T* d = m_data;
        parallel_chunks(n, threads, size_t(1) << 16, [=](int, size_t b, size_t e) { for (size_t i = b; i < e; ++i) d[i] = value; });
    }
    // A copy of [src, src + n), made by several threads.
    void copy_from(const T* src, const size_t n, const int threads)
    {
        resize(n);
        T* d = m_data;
        parallel_chunks(n, threads, size_t(1) << 16, [=](int, size_t b, size_t e) { std::memcpy(d + b, src + b, (e - b) * sizeof(T)); });
    }

  private:
    T*      m_data = nullptr;
    size_t  m_size = 0, m_capacity = 0;
};

}   // namespace asgpu
