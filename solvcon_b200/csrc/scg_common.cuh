// scg_common.cuh -- shared plumbing of libscg_cuda.so (error state, stream, launch helpers).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/scg.h"

namespace scg
{

// ---- error state ------------------------------------------------------------------------
void set_error(char const * fmt, ...);
int fail(char const * fmt, ...); // sets the error and returns -1

#define SCG_CUDA(expr)                                                                          \
    do                                                                                          \
    {                                                                                           \
        cudaError_t const e__ = (expr);                                                         \
        if (cudaSuccess != e__)                                                                 \
        {                                                                                       \
            return ::scg::fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                       \
    } while (0)

#define SCG_REQUIRE(cond, ...)               \
    do                                       \
    {                                        \
        if (!(cond))                         \
        {                                    \
            return ::scg::fail(__VA_ARGS__); \
        }                                    \
    } while (0)

// ---- runtime ------------------------------------------------------------------------------
cudaStream_t stream();   // the library's work stream (valid after scg_init)
int ensure_init();       // lazily binds device 0 if scg_init was never called
void count_launch(int n = 1);

// Per-kernel device timing for diagnostics (SCG_TIMING=1 in the environment): a TimeScope around a launch
// records two events on the work stream; scg_timing_dump() prints launches / average / total per key.
bool timing_on();
void timing_begin(char const * key);
void timing_end();
struct TimeScope
{
    bool on;
    explicit TimeScope(char const * key) : on(timing_on()) { if (on) { timing_begin(key); } }
    ~TimeScope() { if (on) { timing_end(); } }
};
int sm_count();

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- communicator (scg_comm.cu): NCCL send/recv on the library's communication stream ---------------------------
bool comm_ready();
int comm_rank();
int comm_nranks();
cudaStream_t comm_stream();
int comm_group_start();
int comm_group_end();
int comm_send(double const * buf, int64_t count, int peer);
int comm_recv(double * buf, int64_t count, int peer);

// device buffer with RAII; move-only
template <typename T>
struct DevBuf
{
    T * p = nullptr;
    int64_t n = 0;
    DevBuf() = default;
    DevBuf(DevBuf const &) = delete;
    DevBuf & operator=(DevBuf const &) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) { cudaFree(p); }
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(int64_t count)
    {
        release();
        n = count;
        if (0 == count) { return cudaSuccess; }
        return cudaMalloc(reinterpret_cast<void **>(&p), sizeof(T) * static_cast<size_t>(count));
    }
    cudaError_t alloc_zero(int64_t count)
    {
        cudaError_t e = alloc(count);
        if (cudaSuccess != e || 0 == count) { return e; }
        return cudaMemsetAsync(p, 0, sizeof(T) * static_cast<size_t>(count), stream());
    }
};

} // namespace scg
