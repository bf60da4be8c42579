// scg_par.hpp -- a plain host thread pool substitute: split [0, n) into one chunk per hardware thread.
#pragma once

#include <algorithm>
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

} // namespace scg
