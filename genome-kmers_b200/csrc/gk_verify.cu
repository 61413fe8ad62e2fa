// gk_verify.cu -- self-check of a sorted index, straight from the sequence bytes (gk_index_verify).
//
// Independent of how the order was produced: every pair of neighbours is compared with the reference's
// '$'-terminated byte comparator (kmers.py:306-397), ties must be in ascending start order (the reference's
// break_ties=True order, kmers.py:1710-1711), every start must begin a k-mer of min_kmer_len symbols inside
// one record (kmers.py:814-826) and occur once.  With the k-mer count these properties pin the result: a
// sorted permutation of the init set.  Used by bench.py's verification leg and by the full-size tests, where
// no CPU oracle can run.
#include "gk_common.cuh"

namespace gk {

__device__ __forceinline__ int compare_windows(const uint8_t *__restrict__ sba, uint64_t len, uint64_t a,
                                               uint64_t b, uint32_t kmer_len)
{
    for (uint32_t j = 0;; ++j) {
        const uint32_t ca = (a + j < len) ? sba[a + j] : kSep;
        const uint32_t cb = (b + j < len) ? sba[b + j] : kSep;
        const bool a_out = ca == kSep, b_out = cb == kSep;
        if (a_out || b_out) return (a_out && b_out) ? 0 : (a_out ? -1 : 1);  // the shorter k-mer sorts first
        if (ca != cb) return ca < cb ? -1 : 1;
        if (kmer_len && j == kmer_len - 1) return 0;
    }
}

// report: [1] order violations, [2] tie-order violations, [3] invalid starts, [4] duplicate starts,
//         [5] neighbours that differ (groups - 1), [6] head flags that disagree with the bytes
template <typename IdxT>
__global__ void __launch_bounds__(256)
verify_order_kernel(const uint8_t *__restrict__ sba, uint64_t sba_len, const IdxT *__restrict__ idx, uint64_t n,
                    uint32_t kmer_len, const uint64_t *__restrict__ seg_starts, uint32_t n_seg, uint32_t min_len,
                    const uint8_t *__restrict__ flags, uint32_t *__restrict__ seen,
                    unsigned long long *__restrict__ report)
{
    uint32_t bad_order = 0, bad_tie = 0, bad_start = 0, dup = 0, differ = 0, bad_flag = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        const uint64_t s = (uint64_t)idx[r];
        bool ok = s < sba_len;
        if (ok) {
            const uint32_t seg = upper_seg(seg_starts, n_seg, s);
            const uint64_t seg_end = (seg + 1 < n_seg) ? seg_starts[seg + 1] - 1 : sba_len;
            ok = s + min_len <= seg_end;
        }
        if (!ok) { ++bad_start; continue; }
        const uint32_t bit = 1u << (s & 31u);
        if (atomicOr(&seen[s >> 5], bit) & bit) ++dup;
        bool head = true;
        if (r > 0) {
            const uint64_t p = (uint64_t)idx[r - 1];
            if (p < sba_len) {
                const int c = compare_windows(sba, sba_len, p, s, kmer_len);
                if (c > 0) ++bad_order;
                if (c == 0 && p >= s) ++bad_tie;
                if (c != 0) ++differ;
                head = c != 0;
            }
        }
        if (flags && ((flags[r] & kFlagHead) != 0) != head) ++bad_flag;
    }
    const uint32_t v[6] = {bad_order, bad_tie, bad_start, dup, differ, bad_flag};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const uint32_t c = warp_sum(v[i]);
        if (lane_id() == 0 && c) atomicAdd(&report[1 + i], (unsigned long long)c);
    }
}

int verify_order_device(const uint8_t *d_sba, uint64_t sba_len, const void *d_idx, int idx_bytes, uint64_t n,
                        uint32_t kmer_len, const uint64_t *d_seg_starts, uint32_t n_seg, uint32_t min_len,
                        const uint8_t *d_flags, uint64_t *h_report8, uint32_t *d_seen_out, cudaStream_t st)
{
    for (int i = 0; i < 8; ++i) h_report8[i] = 0;
    h_report8[0] = n;
    if (n == 0) return GK_OK;
    DeviceBuffer seen, report;
    const size_t words = (size_t)(sba_len / 32 + 1);
    uint32_t *d_seen = d_seen_out;   // caller's bitmap (already zero) or a scratch one
    if (!d_seen) {
        GK_TRY(seen.alloc(words * 4, st));
        GK_CUDA(cudaMemsetAsync(seen.ptr, 0, words * 4, st));
        d_seen = seen.as<uint32_t>();
    }
    GK_TRY(report.alloc(64, st));
    GK_CUDA(cudaMemsetAsync(report.ptr, 0, 64, st));
    uint64_t blocks = (n + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    if (idx_bytes == 4)
        verify_order_kernel<uint32_t><<<(unsigned)blocks, 256, 0, st>>>(
            d_sba, sba_len, (const uint32_t *)d_idx, n, kmer_len, d_seg_starts, n_seg, min_len, d_flags,
            d_seen, report.as<unsigned long long>());
    else
        verify_order_kernel<uint64_t><<<(unsigned)blocks, 256, 0, st>>>(
            d_sba, sba_len, (const uint64_t *)d_idx, n, kmer_len, d_seg_starts, n_seg, min_len, d_flags,
            d_seen, report.as<unsigned long long>());
    GK_LAUNCH_CHECK();
    unsigned long long h[8];
    GK_CUDA(cudaMemcpyAsync(h, report.ptr, 64, cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    for (int i = 1; i < 7; ++i) h_report8[i] = h[i];
    h_report8[5] = h[5] + 1;          // groups = differing neighbours + 1
    h_report8[7] = d_flags ? 1 : 0;   // were cached head flags compared?
    return GK_OK;
}

__global__ void __launch_bounds__(256)
popcount_words_kernel(const uint32_t *__restrict__ words, uint64_t n_words, unsigned long long *__restrict__ out)
{
    unsigned long long c = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) c += __popc(words[i]);
    c = warp_sum(c);
    if (lane_id() == 0 && c) atomicAdd(out, c);
}

}  // namespace gk

using namespace gk;

extern "C" int gk_popcount_words(const uint32_t *d_words, uint64_t n_words, uint64_t *h_count_out, void *stream)
{
    if (!h_count_out || (n_words && !d_words)) return GK_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    *h_count_out = 0;
    if (n_words == 0) return GK_OK;
    DeviceBuffer out;
    GK_TRY(out.alloc(8, st));
    GK_CUDA(cudaMemsetAsync(out.ptr, 0, 8, st));
    uint64_t blocks = (n_words + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    popcount_words_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_words, n_words, out.as<unsigned long long>());
    GK_LAUNCH_CHECK();
    GK_CUDA(cudaMemcpyAsync(h_count_out, out.ptr, 8, cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    return GK_OK;
}
