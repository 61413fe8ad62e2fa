// gk_core.cu -- library plumbing: status strings, thread-local error text, launch counter,
// memory pool configuration, device info.
#include <stdarg.h>
#include <stdlib.h>
#include <time.h>

#include <mutex>
#include <vector>

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

// ---- block cache in front of the pool ------------------------------------------------------------------------
// The stream-ordered pool serves a request from whatever free virtual ranges it has and re-maps physical memory
// when no range is large enough; for the multi-GB key and start buffers of a sort that re-mapping costs 0.6-2 ms
// (45 ms once) in two of every four consecutive sorts of the same size (profiles/README.md R2.5), and the GPU
// idles behind it.  Big blocks are therefore kept here when they are released and handed back, as they are, to
// the next request of the same size on the same stream: the buffers of one sort are the buffers of the next.
// Reuse on the same stream needs no synchronisation (the new user's work is ordered behind the old user's).
// The cache holds at most kCacheEntries blocks and a quarter of the device's memory; the least recently released
// block goes back to the pool first, and everything does when the pool runs out of memory.  GK_BLOCK_CACHE=0
// switches it off.
namespace {
struct CachedBlock {
    void *ptr;
    size_t bytes;
    cudaStream_t stream;
    int device;
};
constexpr size_t kCacheMinBytes = 1u << 20;
constexpr size_t kCacheEntries = 64;
std::mutex g_cache_mu;
std::vector<CachedBlock> g_cache;   // oldest first
size_t g_cache_bytes = 0;

bool cache_enabled()
{
    static int on = -1;
    if (on < 0) { const char *e = getenv("GK_BLOCK_CACHE"); on = (e && *e == '0') ? 0 : 1; }
    return on == 1;
}
size_t cache_byte_limit()
{
    static size_t limit = 0;
    if (!limit) {
        size_t free_b = 0, total_b = 0;
        limit = (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b) ? total_b / 4 : ((size_t)32 << 30);
    }
    return limit;
}
void cache_flush_locked()
{
    for (const CachedBlock &b : g_cache) cudaFreeAsync(b.ptr, b.stream);
    g_cache.clear();
    g_cache_bytes = 0;
}
}  // namespace

int pool_alloc(void **out, size_t n, cudaStream_t s)
{
    *out = nullptr;
    if (n == 0) return GK_OK;
    GK_TRY(ensure_pool_configured());
    int dev = 0;
    if (n >= kCacheMinBytes && cache_enabled() && cudaGetDevice(&dev) == cudaSuccess) {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        for (size_t i = g_cache.size(); i-- > 0;) {
            const CachedBlock &b = g_cache[i];
            if (b.bytes == n && b.stream == s && b.device == dev) {
                *out = b.ptr;
                g_cache_bytes -= b.bytes;
                g_cache.erase(g_cache.begin() + (long)i);
                return GK_OK;
            }
        }
    }
    cudaError_t err = cudaMallocAsync(out, n, s);
    if (err == cudaErrorMemoryAllocation) {   // give the cached blocks back and try once more
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> lock(g_cache_mu);
            cache_flush_locked();
        }
        cudaDeviceSynchronize();
        err = cudaMallocAsync(out, n, s);
    }
    GK_CUDA(err);
    return GK_OK;
}

void pool_free(void *ptr, size_t n, cudaStream_t s)
{
    if (!ptr) return;
    int dev = 0;
    if (n >= kCacheMinBytes && cache_enabled() && cudaGetDevice(&dev) == cudaSuccess) {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        g_cache.push_back({ptr, n, s, dev});
        g_cache_bytes += n;
        while (g_cache.size() > kCacheEntries || (g_cache_bytes > cache_byte_limit() && g_cache.size() > 1)) {
            cudaFreeAsync(g_cache.front().ptr, g_cache.front().stream);
            g_cache_bytes -= g_cache.front().bytes;
            g_cache.erase(g_cache.begin());
        }
        return;
    }
    cudaFreeAsync(ptr, s);
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
// host-side wall clock at named points of a call (printed by trace_report as offsets from the first point)
static thread_local int g_n_points = 0;
static thread_local const char *g_point_name[32];
static thread_local double g_point_ms[32];
void trace_point(const char *name)
{
    if (!trace_enabled() || g_n_points >= 32) return;
    g_point_name[g_n_points] = name;
    g_point_ms[g_n_points++] = trace_now_ms();
}
void trace_report(const char *what)
{
    if (!trace_enabled()) return;
    if (g_n_points) {
        fprintf(stderr, "[gk trace] %s host ms:", what);
        for (int i = 0; i < g_n_points; ++i) fprintf(stderr, " %s %.3f", g_point_name[i], g_point_ms[i] - g_point_ms[0]);
        fprintf(stderr, "\n");
        g_n_points = 0;
    }
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
