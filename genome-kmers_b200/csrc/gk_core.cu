// gk_core.cu -- library plumbing: status strings, thread-local error text, launch counter,
// memory pool configuration, device info.
#include <stdarg.h>
#include <stdlib.h>
#include <time.h>

#include "gk_common.cuh"

namespace gk {

static thread_local char g_err[512] = "";
static thread_local uint64_t g_launches = 0;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches += (uint64_t)n; }

int ensure_pool_configured()
{
    static thread_local int configured_device = -1;
    int dev = 0;
    GK_CUDA(cudaGetDevice(&dev));
    if (configured_device == dev) return GK_OK;
    cudaMemPool_t pool;
    GK_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t threshold = UINT64_MAX;
    GK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    configured_device = dev;
    return GK_OK;
}

// ---- GK_TRACE=1: where does host time go? ----------------------------------------------------------------
static thread_local double g_alloc_ms = 0.0;
static thread_local uint64_t g_alloc_calls = 0, g_alloc_bytes = 0;
static int trace_enabled()
{
    static int on = -1;
    if (on < 0) { const char *e = getenv("GK_TRACE"); on = (e && *e && *e != '0') ? 1 : 0; }
    return on;
}
double trace_now_ms()
{
    if (!trace_enabled()) return 0.0;
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}
void trace_alloc(double ms, size_t bytes)
{
    if (!trace_enabled()) return;
    g_alloc_ms += ms;
    g_alloc_calls += 1;
    g_alloc_bytes += bytes;
}
void trace_report(const char *what)
{
    if (!trace_enabled()) return;
    fprintf(stderr, "[gk trace] %s: %llu pool allocations, %.1f MB, %.3f ms in cudaMallocAsync\n", what,
            (unsigned long long)g_alloc_calls, (double)g_alloc_bytes / 1e6, g_alloc_ms);
    g_alloc_ms = 0.0;
    g_alloc_calls = g_alloc_bytes = 0;
}

int sm_count()
{
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace gk

extern "C" {

int gk_version(void) { return GKB200_VERSION; }

const char *gk_status_string(int status)
{
    switch (status) {
    case GK_OK: return "ok";
    case GK_ERR_CUDA: return "CUDA error";
    case GK_ERR_ARG: return "invalid argument";
    case GK_ERR_UNSUPPORTED: return "unsupported on the GPU path";
    case GK_ERR_INTERNAL: return "device-side consistency check failed";
    case GK_ERR_INVALID_KMERS: return "k-mers shorter than min_kmer_len";
    case GK_ERR_STATE: return "invalid call order";
    default: return "unknown status";
    }
}

const char *gk_last_error(void) { return gk::g_err; }

uint64_t gk_launch_count(int reset)
{
    uint64_t v = gk::g_launches;
    if (reset) gk::g_launches = 0;
    return v;
}

// ---- peer-visible device memory (multi-GPU exchange buffers) -----------------------------------------
// cudaMalloc memory (pool memory cannot be exported) + CUDA IPC handles; one process per GPU on one box.
int gk_peer_alloc(uint64_t bytes, void **d_ptr_out)
{
    if (!d_ptr_out) return GK_ERR_ARG;
    *d_ptr_out = nullptr;
    GK_CUDA(cudaMalloc(d_ptr_out, bytes ? bytes : 16));
    return GK_OK;
}

int gk_peer_free(void *d_ptr)
{
    if (d_ptr) GK_CUDA(cudaFree(d_ptr));
    return GK_OK;
}

int gk_peer_export(void *d_ptr, uint8_t *handle64_out)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!d_ptr || !handle64_out) return GK_ERR_ARG;
    cudaIpcMemHandle_t h;
    GK_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle64_out, &h, 64);
    return GK_OK;
}

int gk_peer_open(const uint8_t *handle64, void **d_ptr_out)
{
    if (!handle64 || !d_ptr_out) return GK_ERR_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    GK_CUDA(cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return GK_OK;
}

int gk_peer_close(void *d_ptr)
{
    if (d_ptr) GK_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return GK_OK;
}

int gk_device_info(int *sm_count, int *cc_major, int *cc_minor, uint64_t *total_mem_bytes)
{
    int dev = 0;
    GK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    GK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (total_mem_bytes) *total_mem_bytes = (uint64_t)prop.totalGlobalMem;
    return GK_OK;
}

}  // extern "C"
