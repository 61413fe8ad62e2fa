// gk_sba.cu -- sequence byte array helpers: alphabet scan, reverse
// complement, the forward||'$'||revcomp layout, and k-mer start index initialisation.
// All of them are single streaming passes over byte arrays (HBM-bound, no reuse).
#include "gk_common.cuh"

namespace gk {

// ---- alphabet scan (sequence_collection.py:441-459, :693-697) ------------------------------
__global__ void __launch_bounds__(256) scan_alphabet_kernel(const uint8_t *__restrict__ sba,
                                                            uint64_t len,
                                                            unsigned long long *__restrict__ counts)
{
    uint32_t bad = 0, sep = 0, amb = 0;
    const uint64_t n_vec = len / 16;
    const uint4 *v = reinterpret_cast<const uint4 *>(sba);
    const bool aligned = (reinterpret_cast<uintptr_t>(sba) & 15u) == 0;
    auto classify = [&](uint32_t b) {
        if (b == kSep) ++sep;
        else if (is_acgt(b)) {}
        else if (rank4(b) != 0) ++amb;
        else ++bad;
    };
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (aligned) {
        for (uint64_t c = tid; c < n_vec; c += stride) {
            uint4 q = v[c];
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) classify((w[i] >> (8 * j)) & 0xFFu);
        }
        for (uint64_t i = n_vec * 16 + tid; i < len; i += stride) classify(sba[i]);
    } else {
        for (uint64_t i = tid; i < len; i += stride) classify(sba[i]);
    }
    bad = warp_sum(bad);
    sep = warp_sum(sep);
    amb = warp_sum(amb);
    if (lane_id() == 0) {
        if (bad) atomicAdd(&counts[0], (unsigned long long)bad);
        if (sep) atomicAdd(&counts[1], (unsigned long long)sep);
        if (amb) atomicAdd(&counts[2], (unsigned long long)amb);
    }
}

// ---- reverse complement (sequence_collection.py:42-73) -------------------------------------
// One thread produces 16 consecutive output bytes with one aligned 16-byte store.  Their 16 source bytes are
// contiguous too (the mirror image), but not aligned: two aligned 16-byte loads and funnel shifts cut them
// out, PRMT reverses each word, and the complement table sits in shared memory.
__global__ void __launch_bounds__(256) revcomp_kernel(const uint8_t *__restrict__ in,
                                                      uint64_t len, uint8_t *__restrict__ out)
{
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = (uint8_t)complement_byte(threadIdx.x);
    __syncthreads();
    const uint64_t head = (16 - (reinterpret_cast<uintptr_t>(out) & 15u)) & 15u;
    const uint64_t n_chunks = 1 + (len > head ? (len - head + 15) / 16 : 0);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uintptr_t in_addr = reinterpret_cast<uintptr_t>(in);
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += stride) {
        uint64_t o0 = (c == 0) ? 0 : head + (c - 1) * 16;
        uint64_t o1 = (c == 0) ? (head < len ? head : len) : (o0 + 16 < len ? o0 + 16 : len);
        if (c != 0 && o1 - o0 == 16) {
            // source bytes [lo, lo + 16), lo = len - 16 - o0; out[o0 + i] = comp(in[lo + 15 - i])
            const uint64_t lo = len - 16 - o0;
            const uint32_t off = (uint32_t)((in_addr + lo) & 15u);
            uint32_t r[4];
            // the second load may reach past the last source byte but stays inside the array unless the
            // source window ends in the array's last 16-byte line: those few chunks go byte by byte
            if (off == 0) {
                const uint4 q = *reinterpret_cast<const uint4 *>(in + lo);
                r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w;
            } else if (lo >= off && lo - off + 32 <= len) {
                const uint8_t *base = in + lo - off;              // 16-byte aligned
                const uint4 q0 = *reinterpret_cast<const uint4 *>(base);
                const uint4 q1 = *reinterpret_cast<const uint4 *>(base + 16);
                const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                const uint32_t ws = off >> 2, bs = (off & 3u) * 8u;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t a = 0, b2 = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {   // (register-only selection: no local-memory indexing)
                        if ((uint32_t)j == ws + i) a = w[j];
                        if ((uint32_t)j == ws + i + 1) b2 = w[j];
                    }
                    r[i] = __funnelshift_r(a, b2, bs);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t x = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) x |= (uint32_t)in[lo + 4 * i + j] << (8 * j);
                    r[i] = x;
                }
            }
            uint32_t w_out[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t x = __byte_perm(r[3 - i], 0, 0x0123);   // reverse the four bytes
                w_out[i] = (uint32_t)lut[x & 0xFFu] | ((uint32_t)lut[(x >> 8) & 0xFFu] << 8) |
                           ((uint32_t)lut[(x >> 16) & 0xFFu] << 16) | ((uint32_t)lut[x >> 24] << 24);
            }
            *reinterpret_cast<uint4 *>(out + o0) = make_uint4(w_out[0], w_out[1], w_out[2], w_out[3]);
        } else {
            for (uint64_t o = o0; o < o1; ++o) out[o] = lut[in[len - 1 - o]];
        }
    }
}

// ---- k-mer start index initialisation (kmers.py:814-826) ------------------------------------
// Position p of the init array belongs to the last segment s with seg_start[s] - s*k <= p and
// holds start = p + s*k (every earlier record contributes k positions that start no k-mer:
// its last k-1 bases and the '$').
template <typename IdxT>
__global__ void __launch_bounds__(256) init_indices_kernel(const uint64_t *__restrict__ seg_starts,
                                                           uint32_t n_seg, uint32_t k, uint64_t n,
                                                           IdxT *__restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        uint32_t lo = 0, hi = n_seg;
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (p < seg_starts[mid] - (uint64_t)mid * k) hi = mid; else lo = mid + 1;
        }
        out[p] = (IdxT)(p + (uint64_t)(lo - 1) * k);
    }
}

static int grid_for(uint64_t work_items, int block)
{
    uint64_t blocks = (work_items + block - 1) / block;
    uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// alphabet counters left on the device (zeroed here), no synchronise: for callers that read them back later
int scan_alphabet_async(const uint8_t *d_sba, uint64_t len, unsigned long long *d_counts3, cudaStream_t st)
{
    GK_CUDA(cudaMemsetAsync(d_counts3, 0, 3 * sizeof(unsigned long long), st));
    if (len) {
        scan_alphabet_kernel<<<grid_for(len / 16 + 1, 256), 256, 0, st>>>(d_sba, len, d_counts3);
        GK_LAUNCH_CHECK();
    }
    return GK_OK;
}

int kmer_count_host(const uint64_t *h_seg_starts, uint32_t n_seg, uint64_t sba_len, uint32_t k,
                    uint64_t *n_out)
{
    // kmers.py:837-861
    if (!h_seg_starts || n_seg == 0 || k == 0) {
        set_error("kmer_count: empty segment table or kmer_len 0");
        return GK_ERR_ARG;
    }
    uint64_t n = 0;
    for (uint32_t s = 0; s < n_seg; ++s) {
        uint64_t a = h_seg_starts[s];
        uint64_t e_excl = (s + 1 < n_seg) ? h_seg_starts[s + 1] - 1 : sba_len;
        if (e_excl < a + k) {
            set_error("kmer_len (%u) exceeds the length of record %u (%llu)", k, s,
                      (unsigned long long)(e_excl - a));
            return GK_ERR_ARG;
        }
        n += (e_excl - a) - k + 1;
    }
    *n_out = n;
    return GK_OK;
}

int init_indices_device(const uint64_t *d_seg_starts, uint32_t n_seg, uint32_t k, uint64_t n,
                        int idx_bytes, void *d_out, cudaStream_t st)
{
    if (n == 0) return GK_OK;
    int grid = grid_for(n, 256);
    if (idx_bytes == 4)
        init_indices_kernel<uint32_t><<<grid, 256, 0, st>>>(d_seg_starts, n_seg, k, n,
                                                            (uint32_t *)d_out);
    else
        init_indices_kernel<uint64_t><<<grid, 256, 0, st>>>(d_seg_starts, n_seg, k, n,
                                                            (uint64_t *)d_out);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

}  // namespace gk

using namespace gk;

extern "C" {

int gk_sba_scan_alphabet(const uint8_t *d_sba, uint64_t len, uint64_t *h_counts3, void *stream)
{
    if (!h_counts3 || (!d_sba && len)) {
        set_error("gk_sba_scan_alphabet: null pointer");
        return GK_ERR_ARG;
    }
    cudaStream_t st = as_stream(stream);
    DeviceBuffer counts;
    GK_TRY(counts.alloc(3 * sizeof(unsigned long long), st));
    GK_CUDA(cudaMemsetAsync(counts.ptr, 0, counts.bytes, st));
    if (len) {
        scan_alphabet_kernel<<<grid_for(len / 16 + 1, 256), 256, 0, st>>>(
            d_sba, len, counts.as<unsigned long long>());
        GK_LAUNCH_CHECK();
    }
    GK_CUDA(cudaMemcpyAsync(h_counts3, counts.ptr, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    return GK_OK;
}

int gk_sba_scan_alphabet_async(const uint8_t *d_sba, uint64_t len, uint64_t *d_counts3, void *stream)
{
    if (!d_counts3 || (!d_sba && len)) {
        set_error("gk_sba_scan_alphabet_async: null pointer");
        return GK_ERR_ARG;
    }
    return scan_alphabet_async(d_sba, len, reinterpret_cast<unsigned long long *>(d_counts3), as_stream(stream));
}

int gk_sba_revcomp(const uint8_t *d_in, uint64_t len, uint8_t *d_out, void *stream)
{
    if ((!d_in || !d_out) && len) {
        set_error("gk_sba_revcomp: null pointer");
        return GK_ERR_ARG;
    }
    if (d_in == d_out && len) {
        set_error("gk_sba_revcomp: in-place operation is not supported");
        return GK_ERR_ARG;
    }
    if (len == 0) return GK_OK;
    revcomp_kernel<<<grid_for(len / 16 + 2, 256), 256, 0, as_stream(stream)>>>(d_in, len, d_out);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int gk_sba_both_strands(const uint8_t *d_in, uint64_t len, uint8_t *d_out, void *stream)
{
    if ((!d_in || !d_out) && len) {
        set_error("gk_sba_both_strands: null pointer");
        return GK_ERR_ARG;
    }
    cudaStream_t st = as_stream(stream);
    if (len) GK_CUDA(cudaMemcpyAsync(d_out, d_in, len, cudaMemcpyDeviceToDevice, st));
    GK_CUDA(cudaMemsetAsync(d_out + len, kSep, 1, st));
    return gk_sba_revcomp(d_in, len, d_out + len + 1, stream);
}

/* Bytes [begin, end) of the both-strand layout only (a GPU's slice of the start positions first, the rest
 * later on another stream).  The reverse-complement part of the range is the reverse complement of the
 * mirrored part of the input, so this is gk_sba_revcomp on a sub-array. */
int gk_sba_both_strands_range(const uint8_t *d_in, uint64_t len, uint8_t *d_out, uint64_t begin, uint64_t end,
                              void *stream)
{
    if ((!d_in || !d_out) && len) {
        set_error("gk_sba_both_strands_range: null pointer");
        return GK_ERR_ARG;
    }
    const uint64_t total = 2 * len + 1;
    if (end > total) end = total;
    if (begin >= end) return GK_OK;
    cudaStream_t st = as_stream(stream);
    if (begin < len) {
        const uint64_t e = end < len ? end : len;
        GK_CUDA(cudaMemcpyAsync(d_out + begin, d_in + begin, e - begin, cudaMemcpyDeviceToDevice, st));
    }
    if (begin <= len && len < end) GK_CUDA(cudaMemsetAsync(d_out + len, kSep, 1, st));
    if (end > len + 1) {
        const uint64_t rb = begin > len + 1 ? begin - (len + 1) : 0, re = end - (len + 1);   // range inside revcomp
        return gk_sba_revcomp(d_in + (len - re), re - rb, d_out + len + 1 + rb, stream);
    }
    return GK_OK;
}

int gk_kmer_count(const uint64_t *h_seg_starts, uint32_t n_seg, uint64_t sba_len,
                  uint32_t kmer_len, uint64_t *n_out)
{
    if (!n_out) return GK_ERR_ARG;
    return kmer_count_host(h_seg_starts, n_seg, sba_len, kmer_len, n_out);
}

int gk_kmer_init_indices(const uint64_t *h_seg_starts, uint32_t n_seg, uint64_t sba_len,
                         uint32_t kmer_len, int idx_bytes, void *d_idx_out, void *stream)
{
    if (idx_bytes != 4 && idx_bytes != 8) {
        set_error("idx_bytes must be 4 or 8");
        return GK_ERR_ARG;
    }
    uint64_t n = 0;
    GK_TRY(kmer_count_host(h_seg_starts, n_seg, sba_len, kmer_len, &n));
    cudaStream_t st = as_stream(stream);
    DeviceBuffer segs;
    GK_TRY(segs.alloc((size_t)n_seg * 8, st));
    GK_CUDA(cudaMemcpyAsync(segs.ptr, h_seg_starts, (size_t)n_seg * 8, cudaMemcpyHostToDevice, st));
    GK_TRY(init_indices_device(segs.as<uint64_t>(), n_seg, kmer_len, n, idx_bytes, d_idx_out, st));
    GK_CUDA(cudaStreamSynchronize(st));  // h_seg_starts may be pageable: do not outlive the call
    return GK_OK;
}

}  // extern "C"
