// scg_par.hpp -- a plain host thread pool substitute: split [0, n) into one chunk per hardware thread.
#pragma once

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <functional>
#include <thread>
#include <vector>

namespace scg
{

inline unsigned host_threads()
{
    unsigned nt = std::thread::hardware_concurrency();
    return nt > 32 ? 32 : (nt < 1 ? 1 : nt);
}

inline void host_parallel_for(int64_t n, std::function<void(int64_t, int64_t)> const & fn, int64_t serial_below = 200000)
{
    unsigned const nt = host_threads();
    if (nt < 2 || n < serial_below)
    {
        fn(0, n);
        return;
    }
    std::vector<std::thread> th;
    int64_t const chunk = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t)
    {
        int64_t const b = t * chunk, e = std::min<int64_t>(n, b + chunk);
        if (b >= e) { break; }
        th.emplace_back([=, &fn]() { fn(b, e); });
    }
    for (auto & t : th) { t.join(); }
}

// Host array allocated UNINITIALISED (a std::vector would value-initialise on one thread) and filled by the caller,
// normally through host_parallel_for.
template <typename T>
struct HostBuf
{
    T * p = nullptr;
    size_t n = 0;
    HostBuf() = default;
    HostBuf(HostBuf const &) = delete;
    HostBuf & operator=(HostBuf const &) = delete;
    ~HostBuf() { std::free(p); }
    void alloc(size_t count)
    {
        std::free(p);
        p = count ? static_cast<T *>(std::malloc(count * sizeof(T))) : nullptr;
        n = count;
    }
    void copy_from(T const * src, size_t count)
    {
        alloc(count);
        T * const dst = p;
        host_parallel_for(static_cast<int64_t>(count), [=](int64_t b, int64_t e) { std::copy(src + b, src + e, dst + b); });
    }
    T * data() { return p; }
    T const * data() const { return p; }
    size_t size() const { return n; }
    T & operator[](size_t k) { return p[k]; }
    T const & operator[](size_t k) const { return p[k]; }
};

// SCG_HMESH_TIMING=1: wall time of the setup stages on stderr
struct HostStageTimer
{
    bool on;
    char const * tag;
    std::chrono::steady_clock::time_point t;
    explicit HostStageTimer(char const * tag_)
        : on(std::getenv("SCG_HMESH_TIMING") && std::getenv("SCG_HMESH_TIMING")[0] == '1'), tag(tag_), t(std::chrono::steady_clock::now())
    {
    }
    void lap(char const * what)
    {
        if (!on) { return; }
        auto const n = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[%s] %-28s %8.3f s\n", tag, what, std::chrono::duration<double>(n - t).count());
        t = n;
    }
};

} // namespace scg
