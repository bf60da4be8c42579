// scg_core.cu -- runtime of libscg_cuda.so: device binding, work stream, error state,
// CUDA-event timer, L2 flush, pinned host memory.
#include "scg_common.cuh"

#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include <atomic>
#include <cstring>

namespace scg
{

static thread_local std::string g_error;
static cudaStream_t g_stream = nullptr;
static int g_device = -1;
static int g_sm_count = 0;
static std::atomic<int64_t> g_launches{0};
static cudaEvent_t g_ev0 = nullptr;
static cudaEvent_t g_ev1 = nullptr;
static void * g_flush_buf = nullptr;
static int64_t g_flush_bytes = 0;

void set_error(char const * fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

int fail(char const * fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return -1;
}

cudaStream_t stream() { return g_stream; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

namespace
{
struct TimingRec
{
    std::string key;
    cudaEvent_t a, b;
};
std::vector<TimingRec> g_timing;
int g_timing_state = -1;
} // namespace

bool timing_on()
{
    if (g_timing_state < 0)
    {
        char const * e = std::getenv("SCG_TIMING");
        g_timing_state = (e && e[0] == '1') ? 1 : 0;
    }
    return 1 == g_timing_state;
}

void timing_begin(char const * key)
{
    TimingRec r;
    r.key = key;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, g_stream);
    if (g_timing.size() >= (size_t(1) << 16)) { scg_timing_dump(); } // bounded: flush the records of a long run
    g_timing.push_back(r);
}

void timing_end() { cudaEventRecord(g_timing.back().b, g_stream); }
int sm_count() { return g_sm_count; }

static int init_device(int device)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (cudaSuccess != e || 0 == ndev)
    {
        return fail("libscg_cuda: no CUDA device available (%s); there is no CPU fallback",
                    cudaSuccess != e ? cudaGetErrorString(e) : "device count is 0");
    }
    SCG_REQUIRE(device >= 0 && device < ndev, "scg_init: device %d out of range [0, %d)", device, ndev);
    if (g_device == device && g_stream) { return 0; }
    SCG_REQUIRE(g_device < 0, "scg_init: already bound to device %d", g_device);
    SCG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SCG_CUDA(cudaGetDeviceProperties(&prop, device));
    g_sm_count = prop.multiProcessorCount;
    SCG_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    SCG_CUDA(cudaEventCreate(&g_ev0));
    SCG_CUDA(cudaEventCreate(&g_ev1));
    g_device = device;
    return 0;
}

int ensure_init()
{
    if (g_stream)
    {
        // a ctypes caller may hop threads; keep the device current on this one
        cudaSetDevice(g_device);
        return 0;
    }
    return init_device(0);
}

} // namespace scg

using namespace scg;

extern "C" {

const char * scg_last_error(void) { return g_error.c_str(); }

int scg_build_is_strict(void)
{
#ifdef SCG_STRICT_FP
    return 1;
#else
    return 0;
#endif
}

int scg_device_count(int * count)
{
    int n = 0;
    cudaError_t const e = cudaGetDeviceCount(&n);
    if (cudaSuccess != e)
    {
        *count = 0;
        return fail("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    *count = n;
    return 0;
}

int scg_init(int device) { return init_device(device); }

int scg_shutdown(void)
{
    if (g_stream)
    {
        cudaStreamSynchronize(g_stream);
        if (g_flush_buf) { cudaFree(g_flush_buf); }
        g_flush_buf = nullptr;
        g_flush_bytes = 0;
        cudaEventDestroy(g_ev0);
        cudaEventDestroy(g_ev1);
        cudaStreamDestroy(g_stream);
        g_stream = nullptr;
        g_device = -1;
    }
    return 0;
}

int scg_sync(void)
{
    if (int rc = ensure_init()) { return rc; }
    SCG_CUDA(cudaStreamSynchronize(g_stream));
    SCG_CUDA(cudaGetLastError());
    return 0;
}

int scg_mem_info(int64_t * free_bytes, int64_t * total_bytes)
{
    if (int rc = ensure_init()) { return rc; }
    size_t f = 0, t = 0;
    SCG_CUDA(cudaMemGetInfo(&f, &t));
    *free_bytes = static_cast<int64_t>(f);
    *total_bytes = static_cast<int64_t>(t);
    return 0;
}

int scg_timing_dump(void)
{
    if (g_timing.empty()) { return 0; }
    cudaStreamSynchronize(g_stream);
    std::map<std::string, std::pair<int, double>> agg;
    double total = 0.0;
    for (auto & r : g_timing)
    {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto & x = agg[r.key];
        x.first += 1;
        x.second += ms;
        total += ms;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_timing.clear();
    for (auto const & kv : agg)
    {
        std::fprintf(stderr, "[scg timing] %-28s %6d x %10.1f us  %5.1f%%\n", kv.first.c_str(), kv.second.first,
                     1e3 * kv.second.second / kv.second.first, 100.0 * kv.second.second / total);
    }
    return 0;
}

void * scg_stream_handle(void) { return static_cast<void *>(g_stream); }

int64_t scg_launch_count(int reset)
{
    int64_t const v = g_launches.load();
    if (reset) { g_launches.store(0); }
    return v;
}

int scg_timer_start(void)
{
    if (int rc = ensure_init()) { return rc; }
    SCG_CUDA(cudaEventRecord(g_ev0, g_stream));
    return 0;
}

int scg_timer_stop(double * elapsed_ms)
{
    if (int rc = ensure_init()) { return rc; }
    SCG_CUDA(cudaEventRecord(g_ev1, g_stream));
    SCG_CUDA(cudaEventSynchronize(g_ev1));
    float ms = 0.f;
    SCG_CUDA(cudaEventElapsedTime(&ms, g_ev0, g_ev1));
    *elapsed_ms = static_cast<double>(ms);
    return 0;
}

int scg_flush_l2(int64_t bytes)
{
    if (int rc = ensure_init()) { return rc; }
    if (bytes > g_flush_bytes)
    {
        if (g_flush_buf) { cudaFree(g_flush_buf); }
        g_flush_buf = nullptr;
        g_flush_bytes = 0;
        SCG_CUDA(cudaMalloc(&g_flush_buf, static_cast<size_t>(bytes)));
        g_flush_bytes = bytes;
    }
    static int toggle = 0;
    toggle ^= 1;
    SCG_CUDA(cudaMemsetAsync(g_flush_buf, toggle, static_cast<size_t>(bytes), g_stream));
    return 0;
}

// page-locked host memory for the API-visible arrays: host<->device copies of pageable numpy memory run at
// ~12 GB/s through the driver's bounce buffer, of pinned memory at the PCIe rate (~50 GB/s on Gen5)
int scg_host_alloc(void ** ptr, int64_t bytes)
{
    if (int rc = ensure_init()) { return rc; }
    SCG_CUDA(cudaHostAlloc(ptr, static_cast<size_t>(bytes > 0 ? bytes : 1), cudaHostAllocPortable));
    return 0;
}

int scg_dev_alloc(void ** ptr, int64_t bytes)
{
    if (int rc = ensure_init()) { return rc; }
    SCG_CUDA(cudaMalloc(ptr, static_cast<size_t>(bytes > 0 ? bytes : 1)));
    return 0;
}

int scg_dev_free(void * ptr)
{
    if (ptr) { SCG_CUDA(cudaFree(ptr)); }
    return 0;
}

int scg_dev_copy(void * dst, const void * src, int64_t bytes)
{
    if (int rc = ensure_init()) { return rc; }
    SCG_CUDA(cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), cudaMemcpyDefault, g_stream));
    return 0;
}

int scg_host_free(void * ptr)
{
    if (ptr) { SCG_CUDA(cudaFreeHost(ptr)); }
    return 0;
}

} // extern "C"
