// gk_common.cuh -- shared helpers for the libgkb200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gkb200.h"

namespace gk {

constexpr uint8_t kSep = 36;  // '$', sequence_collection.py:689-691

// per-slot flag bits of a sorted order (one byte per k-mer)
constexpr uint8_t kFlagHead = 1;   // first k-mer of a group of equal k-mers
constexpr uint8_t kFlagAmb = 2;    // slot holds a non-ACGT window (key class bit 0)
constexpr uint8_t kFlagPass = 4;   // element passes a filter / validity test (scratch arrays)
constexpr uint8_t kFlagMulti = 8;  // member of a group with more than one element (scratch arrays)
constexpr uint8_t kFlagLong = 16;  // member of a prefix run too long for the in-place tie repair

constexpr int kBigBucketCap = 15;  // out-of-order buckets too long for one CTA, listed for the host
constexpr int kDescentCap = 4096;  // out-of-order positions of long prefix runs listed for the bucket-wise repair

// ---- error plumbing ------------------------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define GK_CUDA(expr)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            gk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                          __LINE__);                                                         \
            return GK_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#define GK_TRY(expr)                  \
    do {                              \
        int _s = (expr);              \
        if (_s != GK_OK) return _s;   \
    } while (0)

#define GK_LAUNCH_CHECK()                                                                    \
    do {                                                                                     \
        gk::count_launch();                                                                  \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) {                                                             \
            gk::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),        \
                          __FILE__, __LINE__);                                               \
            return GK_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

// ---- stream-ordered temporary device memory ------------------------------------------------
// All scratch comes from the device's default mempool (cudaMallocAsync); the release threshold
// is raised once so repeated sorts reuse the same pages instead of going back to the driver.
int ensure_pool_configured();
int pool_alloc(void **out, size_t n, cudaStream_t s);   // block cache in front of cudaMallocAsync (gk_core.cu)
void pool_free(void *ptr, size_t n, cudaStream_t s);
void trace_alloc(double ms, size_t bytes);  // GK_TRACE=1: allocation time accounting (gk_core.cu)

double trace_now_ms();

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    ~DeviceBuffer() { release(); }
    int alloc(size_t n, cudaStream_t s)
    {
        release();
        stream = s;
        bytes = n;
        if (n == 0) return GK_OK;
        const double t0 = trace_now_ms();
        GK_TRY(pool_alloc(&ptr, n, s));
        trace_alloc(trace_now_ms() - t0, n);
        return GK_OK;
    }
    void release()
    {
        if (ptr) pool_free(ptr, bytes, stream);
        ptr = nullptr;
        bytes = 0;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(ptr); }
};

// device time of one radix sort, split the way bench.py reports it.  A sort that runs without a
// synchronise (deferred error word) leaves its events pending; the caller resolves them after its own
// final synchronise.
struct SortTiming {
    float hist_ms = 0.f;    // digit histogram + scan (one read of the keys)
    float passes_ms = 0.f;  // all onesweep passes (and their status memsets)
    int passes = 0;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    bool pending = false;
    SortTiming() = default;
    SortTiming(const SortTiming &) = delete;
    SortTiming &operator=(const SortTiming &) = delete;
    ~SortTiming() { drop(); }
    void drop()
    {
        for (auto &e : ev) {
            if (e) cudaEventDestroy(e);
            e = nullptr;
        }
        pending = false;
    }
    void resolve()  // after a synchronise of the stream the events were recorded on
    {
        if (pending) {
            cudaEventElapsedTime(&hist_ms, ev[0], ev[1]);
            cudaEventElapsedTime(&passes_ms, ev[1], ev[2]);
        }
        drop();
    }
};

// Ambiguous-window fragments: the pack kernel lists every block of consecutive identical
// non-ACGT windows (all windows inside an N run are one block per pack tile) as
// (radix key, 4-bit rank words, first start, number of windows).  Unordered (slots come from an atomic
// counter); entries beyond `capacity` are counted but not written.
struct FragOut {
    uint64_t *key = nullptr, *w0 = nullptr, *w1 = nullptr, *start = nullptr;
    uint32_t *count = nullptr;
    unsigned long long *counter = nullptr;
    uint64_t capacity = 0;
};

// Sorted fragment list (device): what frag_sort_device leaves for frag_expand_device (gk_frag.cu).
struct FragSorted {
    DeviceBuffer skey_a, skey_b, perm_a, perm_b, sstart, off, whead, slot0;
    const uint64_t *skey = nullptr;
    uint64_t F = 0, cap = 0;
    // Room for up to `f` fragments.  Called on the MAIN stream before the side stream forks, so that the pool
    // never has to hand memory from one stream to the other (that costs a fresh device allocation).
    int reserve(uint64_t f, cudaStream_t st)
    {
        if (f <= cap) return GK_OK;
        const size_t pad = (size_t)((f + 1) & ~1ull);  // (16-byte aligned key buffers)
        GK_TRY(skey_a.alloc(pad * 8, st));
        GK_TRY(skey_b.alloc(pad * 8, st));
        GK_TRY(perm_a.alloc(pad * 4, st));
        GK_TRY(perm_b.alloc(pad * 4, st));
        GK_TRY(sstart.alloc((size_t)f * 8, st));
        GK_TRY(off.alloc((size_t)(f + 1) * 8, st));
        GK_TRY(whead.alloc((size_t)f, st));
        GK_TRY(slot0.alloc((size_t)f * 8, st));
        cap = f;
        return GK_OK;
    }
    // buffers that were (re)allocated on a side stream are last used on the main stream: release them there
    void rebind(cudaStream_t st)
    {
        for (DeviceBuffer *b : {&skey_a, &skey_b, &perm_a, &perm_b, &sstart, &off, &whead, &slot0}) b->stream = st;
    }
};

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
int sm_count();

// ---- device helpers ------------------------------------------------------------------------
// 2-bit code with A<C<G<T: A=65,C=67,G=71,T=84 -> bits 2:1 are 00,01,11,10.
__device__ __forceinline__ uint32_t code2(uint32_t b)
{
    uint32_t c = (b >> 1) & 3u;
    return c ^ (c >> 1);
}
__device__ __forceinline__ bool is_acgt(uint32_t b)
{
    // letters live in 64..95; bit (b & 31) of the mask marks A(1) C(3) G(7) T(20)
    constexpr uint32_t mask = (1u << 1) | (1u << 3) | (1u << 7) | (1u << 20);
    return ((b & 0xE0u) == 0x40u) && ((mask >> (b & 31u)) & 1u);
}
// number of A/C/G/T that sort below byte b in raw ASCII order (kmers.py:381-388); '$' -> 0
__device__ __forceinline__ uint32_t acgt_below(uint32_t b)
{
    return (b > 65u) + (b > 67u) + (b > 71u) + (b > 84u);
}
// 4-bit rank of an allowed symbol: '$'/unknown = 0, then A B C D G H K M N R S T V W Y = 1..15
__device__ __forceinline__ uint32_t rank4(uint32_t b)
{
    constexpr uint64_t lo = (1ull << 4) | (2ull << 8) | (3ull << 12) | (4ull << 16) |
                            (5ull << 28) | (6ull << 32) | (7ull << 44) | (8ull << 52) |
                            (9ull << 56);
    constexpr uint64_t hi = (10ull << 8) | (11ull << 12) | (12ull << 16) | (13ull << 24) |
                            (14ull << 28) | (15ull << 36);
    if ((b & 0xE0u) != 0x40u) return 0u;
    uint32_t i = b & 31u;
    uint64_t t = (i & 16u) ? hi : lo;
    return (uint32_t)(t >> (4u * (i & 15u))) & 15u;
}
__device__ __forceinline__ uint32_t complement_byte(uint32_t b)
{
    // sequence_collection.py:410-427; bytes the reference's table does not know map to 0
    switch (b) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'R': return 'Y'; case 'Y': return 'R'; case 'S': return 'S'; case 'W': return 'W';
    case 'K': return 'M'; case 'M': return 'K'; case 'B': return 'V'; case 'D': return 'H';
    case 'H': return 'D'; case 'V': return 'B'; case 'N': return 'N'; case '$': return '$';
    default: return 0;
    }
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// index of the last entry of a sorted table that is <= x (table[0] <= x assumed)
__device__ __forceinline__ uint32_t upper_seg(const uint64_t *__restrict__ starts, uint32_t n,
                                              uint64_t x)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (x < starts[mid]) hi = mid; else lo = mid + 1;
    }
    return lo - 1;
}

}  // namespace gk
