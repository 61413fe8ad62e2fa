// gk_frag.cu -- ambiguous windows as run-length blocks ("fragments"), from pack time to the final order
// (DESIGN.md 4.2b).
//
// A window that holds a non-ACGT symbol gets a radix key that orders it exactly against every pure window
// (gk_pack.cu), so the main sort puts the ambiguous windows into the right SLOTS as a set.  What is left is
// their order among themselves, by the reference's raw byte comparison (kmers.py:381-388).  Real genomes and
// the bench workload hold almost all of them in N runs -- millions of identical windows -- so they are never
// handled as elements here:
//   pack kernel        lists every block of consecutive identical ambiguous windows as one fragment
//                      (key, 4-bit rank words w0/w1, first start, count): ~25 k fragments for 8.4 M windows
//   frag_sort_device   orders the fragments by (key, w0, w1, start) -- a few single-CTA sorts, independent of
//                      the main sort, so it runs beside it on a second stream
//   frag_expand_device finds the slot range of every fragment (lower bound of its key in the sorted keys +
//                      windows of earlier fragments with the same key) and writes starts and head flags
// Nothing here reads or writes per-k-mer data except the final writes into the ambiguous slots.
#include "gk_common.cuh"

namespace gk {

int radix_sort_pairs_device(uint64_t *, uint64_t *, void *, void *, int, uint64_t, int, int, int *, cudaStream_t,
                            SortTiming *, const unsigned long long *d_pre_hist = nullptr, void *d_vals_final = nullptr);

static int frag_grid(uint64_t items)
{
    uint64_t blocks = (items + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

__global__ void __launch_bounds__(256)
frag_iota_kernel(uint32_t *__restrict__ perm, uint64_t F)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < F; q += stride) perm[q] = (uint32_t)q;
}

__global__ void __launch_bounds__(256)
frag_gather_kernel(const uint64_t *__restrict__ word, const uint32_t *__restrict__ perm, uint64_t F,
                   uint64_t *__restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < F; q += stride) out[q] = word[perm[q]];
}

// sorted order -> dense arrays: first start, count, "first fragment of a group of equal k-mers"
__global__ void __launch_bounds__(256)
frag_finish_kernel(const uint64_t *__restrict__ skey, const uint64_t *__restrict__ w0,
                   const uint64_t *__restrict__ w1, const uint64_t *__restrict__ start,
                   const uint32_t *__restrict__ count, const uint32_t *__restrict__ perm, uint64_t F,
                   uint64_t *__restrict__ sstart, unsigned long long *__restrict__ off, uint8_t *__restrict__ whead)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < F; q += stride) {
        const uint32_t f = perm[q];
        sstart[q] = start[f];
        off[q] = count[f];
        bool head = true;
        if (q > 0) {
            const uint32_t g = perm[q - 1];
            head = skey[q] != skey[q - 1] || w0[f] != w0[g] || w1[f] != w1[g];
        }
        whead[q] = head ? 1 : 0;
    }
}

// one CTA: in-place exclusive scan of F counts, total appended at [F]
__global__ void __launch_bounds__(1024)
frag_scan_kernel(unsigned long long *__restrict__ data, uint64_t F)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint64_t base = 0; base < F; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long c = (i < F) ? data[i] : 0;
        unsigned long long inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += v;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        unsigned long long pre = s_carry;
        for (uint32_t w = 0; w < warp; ++w) pre += s_warp[w];
        if (i < F) data[i] = pre + inc - c;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = pre + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) data[F] = s_carry;
}

__device__ __forceinline__ uint64_t lower_bound_u64(const uint64_t *__restrict__ a, uint64_t n, uint64_t x)
{
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// First slot of every fragment in the sorted order.  All windows with an ambiguous key K sit in one run of
// slots that starts at the lower bound of K in the sorted keys (no pure window shares a class-0 key);
// inside the run the fragments follow each other in (w0, w1, start) order.
// err value 2: fragment windows != ambiguous windows counted by the pack kernel; 4: key not found.
__global__ void __launch_bounds__(256)
frag_slots_kernel(const uint64_t *__restrict__ skey, const unsigned long long *__restrict__ off, uint64_t F,
                  const uint64_t *__restrict__ keys_sorted, uint64_t n, const unsigned int *__restrict__ descent,
                  const unsigned long long *__restrict__ n_amb_expected, unsigned long long *__restrict__ slot0,
                  int *__restrict__ err)
{
    if (*descent) return;   // the keys are not fully sorted yet: the caller repairs them and calls again
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid == 0 && n_amb_expected && off[F] != *n_amb_expected) atomicOr(err, 2);
    for (uint64_t q = tid; q < F; q += stride) {
        const uint64_t kf = skey[q];
        const uint64_t g = lower_bound_u64(skey, F, kf);
        const uint64_t lb = lower_bound_u64(keys_sorted, n, kf);
        if (lb >= n || keys_sorted[lb] != kf) atomicOr(err, 4);
        slot0[q] = lb + (off[q] - off[g]);
    }
}

template <typename IdxT>
__global__ void __launch_bounds__(256)
frag_expand_kernel(const uint64_t *__restrict__ sstart, const unsigned long long *__restrict__ off,
                   const unsigned long long *__restrict__ slot0, const uint8_t *__restrict__ whead, uint64_t F,
                   uint64_t n, const unsigned int *__restrict__ descent, const int *__restrict__ err,
                   IdxT *__restrict__ d_idx, uint8_t *__restrict__ d_flags)
{
    if (*descent || (*err & 6)) return;
    const uint64_t m = off[F];
    // a warp writes 256 consecutive windows, lane-strided (coalesced stores): one binary search for the chunk's
    // first window, then every lane walks forward on its own (almost all windows lie in fragments of thousands
    // -- the pieces of an N run -- so the walk rarely moves)
    constexpr int kPer = 8;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp_id = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t o0 = warp_id * (32 * kPer); o0 < m; o0 += n_warps * (32 * kPer)) {
        uint64_t lo = 0, hi = F;  // last fragment whose offset is <= o0
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (o0 < off[mid]) hi = mid; else lo = mid + 1;
        }
        uint64_t q = lo - 1;
        uint64_t q_off = off[q], q_end = off[q + 1], q_slot = slot0[q], q_start = sstart[q];
        uint8_t q_head = whead[q];
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const uint64_t o = o0 + (uint64_t)u * 32 + lane;
            if (o >= m) break;
            if (o >= q_end) {   // (fragments are never empty)
                do {
                    ++q;
                    q_off = q_end;
                    q_end = off[q + 1];
                } while (o >= q_end);
                q_slot = slot0[q]; q_start = sstart[q]; q_head = whead[q];
            }
            const uint64_t i = o - q_off;
            const uint64_t slot = q_slot + i;
            if (slot < n) {
                d_idx[slot] = (IdxT)(q_start + i);
                d_flags[slot] = kFlagAmb | ((i == 0 && q_head) ? kFlagHead : 0);
            }
        }
    }
}

// Order F fragments by (key, w0, w1, start); stable LSD, least significant word first.  Only small sorts:
// no synchronise (the look-back error word of larger lists is the caller's deferred one).
// with_key = false: by (w0, w1, start) only.  That is the same order -- the words hold the window's symbols in
// raw byte order and the radix key is a monotone function of the window -- and saves the key's digit passes.
static int frag_sort_perm(const FragOut &frag, uint64_t F, uint32_t key_len, int key_bits, int start_bits,
                          bool with_key, FragSorted &out, cudaStream_t st, uint32_t **perm_out,
                          const uint64_t **last_word_out)
{
    GK_TRY(out.reserve(F, st));
    const int grid = frag_grid(F);
    frag_iota_kernel<<<grid, 256, 0, st>>>(out.perm_a.as<uint32_t>(), F);
    GK_LAUNCH_CHECK();
    uint32_t *cur = out.perm_a.as<uint32_t>(), *alt = out.perm_b.as<uint32_t>();
    uint64_t *ka = out.skey_a.as<uint64_t>(), *kb = out.skey_b.as<uint64_t>();
    const uint64_t *sorted_word = ka;
    auto sort_word = [&](const uint64_t *word, int begin_bit, int end_bit) -> int {
        frag_gather_kernel<<<grid, 256, 0, st>>>(word, cur, F, ka);
        GK_LAUNCH_CHECK();
        int alt_has = 0;
        if (F > 1)
            GK_TRY(radix_sort_pairs_device(ka, kb, cur, alt, 4, F, begin_bit, end_bit, &alt_has, st, nullptr));
        if (alt_has) { uint32_t *t = cur; cur = alt; alt = t; }
        sorted_word = alt_has ? kb : ka;
        return GK_OK;
    };
    GK_TRY(sort_word(frag.start, 0, start_bits));
    if (key_len > 16) GK_TRY(sort_word(frag.w1, 64 - 4 * ((int)key_len - 16), 64));
    GK_TRY(sort_word(frag.w0, 64 - 4 * ((int)key_len < 16 ? (int)key_len : 16), 64));
    if (with_key) GK_TRY(sort_word(frag.key, 0, key_bits));
    *perm_out = cur;
    *last_word_out = sorted_word;
    return GK_OK;
}

int frag_sort_device(const FragOut &frag, uint64_t F, uint32_t key_len, int key_bits, int start_bits,
                     FragSorted &out, cudaStream_t st)
{
    out.F = F;
    if (F == 0) return GK_OK;
    uint32_t *perm = nullptr;
    const uint64_t *sorted_key = nullptr;
    GK_TRY(frag_sort_perm(frag, F, key_len, key_bits, start_bits, true, out, st, &perm, &sorted_key));
    out.skey = sorted_key;
    frag_finish_kernel<<<frag_grid(F), 256, 0, st>>>(out.skey, frag.w0, frag.w1, frag.start, frag.count, perm, F,
                                                     out.sstart.as<uint64_t>(), out.off.as<unsigned long long>(),
                                                     out.whead.as<uint8_t>());
    GK_LAUNCH_CHECK();
    frag_scan_kernel<<<1, 1024, 0, st>>>(out.off.as<unsigned long long>(), F);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

// ---- multi-GPU: fragments are sorted where they are listed, and merged where they are used -----------------
// A rank sorts the fragment list of its slice (a few ten thousand entries, beside the exchange); the owner of a
// key range then MERGES the sorted lists of all ranks instead of sorting their union: the position of a
// fragment is its index in its own list plus, for every other list, the number of fragments that go before it
// (one binary search each).  Lists of lower ranks hold lower starts, so equal windows order by rank.
__global__ void __launch_bounds__(256)
frag_permute_kernel(FragOut in, const uint32_t *__restrict__ perm, uint64_t F, FragOut out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < F; q += stride) {
        const uint32_t f = perm[q];
        out.key[q] = in.key[f];
        out.w0[q] = in.w0[f];
        out.w1[q] = in.w1[f];
        out.start[q] = in.start[f];
        out.count[q] = in.count[f];
    }
}

int frag_sort_copy_device(const FragOut &in, uint64_t F, uint32_t key_len, int start_bits, const FragOut &out,
                          cudaStream_t st)
{
    if (F == 0) return GK_OK;
    FragSorted scratch;
    uint32_t *perm = nullptr;
    const uint64_t *unused = nullptr;
    GK_TRY(frag_sort_perm(in, F, key_len, 0, start_bits, false, scratch, st, &perm, &unused));
    frag_permute_kernel<<<frag_grid(F), 256, 0, st>>>(in, perm, F, out);
    GK_LAUNCH_CHECK();
    return GK_OK;   // (scratch is released in stream order)
}

struct FragList {   // one rank's list inside the gathered buffer
    const uint64_t *key, *w0, *w1, *start;
    const uint32_t *count;
};
__device__ __forceinline__ FragList frag_list_of(const unsigned char *gathered, uint32_t src, uint64_t cap)
{
    const uint64_t *b = reinterpret_cast<const uint64_t *>(gathered + (size_t)src * cap * 36);
    FragList l;
    l.key = b; l.w0 = b + cap; l.w1 = b + 2 * cap; l.start = b + 3 * cap;
    l.count = reinterpret_cast<const uint32_t *>(b + 4 * cap);
    return l;
}

// range[2 s], range[2 s + 1]: the fragments of list s whose key lies in [key_lo, key_hi) -- contiguous, because
// the key is monotone along a sorted list; *n_frag = their number over all lists.  One warp.
__global__ void frag_ranges_kernel(const unsigned char *__restrict__ gathered,
                                   const unsigned long long *__restrict__ counts, uint32_t world, uint64_t cap,
                                   uint64_t key_lo, uint64_t key_hi, unsigned long long *__restrict__ range,
                                   unsigned long long *__restrict__ n_frag)
{
    const uint32_t s = threadIdx.x;
    unsigned long long len = 0;
    if (s < world) {
        const FragList l = frag_list_of(gathered, s, cap);
        const uint64_t n_s = counts[s] < cap ? counts[s] : cap;
        const uint64_t a = lower_bound_u64(l.key, n_s, key_lo);
        const uint64_t b = key_hi ? lower_bound_u64(l.key, n_s, key_hi) : n_s;
        range[2 * s] = a;
        range[2 * s + 1] = b;
        len = b - a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(0xffffffffu, len, o);
    if (s == 0) *n_frag = len;
}

// fragments of [lo, hi) of list l that go before the words (x0, x1); with `or_equal` equal words count too
__device__ __forceinline__ uint64_t frag_words_before(const FragList &l, uint64_t lo, uint64_t hi, uint64_t x0,
                                                      uint64_t x1, bool or_equal)
{
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        const uint64_t m0 = l.w0[mid];
        bool before = m0 < x0;
        if (m0 == x0) {
            const uint64_t m1 = l.w1[mid];
            before = or_equal ? (m1 <= x1) : (m1 < x1);
        }
        if (before) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
frag_merge_kernel(const unsigned char *__restrict__ gathered, uint32_t world, uint64_t cap,
                  const unsigned long long *__restrict__ range, uint64_t key_lo, uint64_t out_cap,
                  uint64_t *__restrict__ skey, uint64_t *__restrict__ w0m, uint64_t *__restrict__ w1m,
                  uint64_t *__restrict__ sstart, uint32_t *__restrict__ count)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t total = (uint64_t)world * cap;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const uint32_t src = (uint32_t)(t / cap);
        const uint64_t j = t - (uint64_t)src * cap;
        if (j < range[2 * src] || j >= range[2 * src + 1]) continue;
        const FragList mine = frag_list_of(gathered, src, cap);
        const uint64_t x0 = mine.w0[j], x1 = mine.w1[j];
        uint64_t pos = j - range[2 * src];
        for (uint32_t o = 0; o < world; ++o) {
            if (o == src) continue;
            const uint64_t a = range[2 * o], b = range[2 * o + 1];
            if (a == b) continue;
            pos += frag_words_before(frag_list_of(gathered, o, cap), a, b, x0, x1, o < src) - a;
        }
        if (pos < out_cap) {
            skey[pos] = mine.key[j] - key_lo;
            w0m[pos] = x0;
            w1m[pos] = x1;
            sstart[pos] = mine.start[j];
            count[pos] = mine.count[j];
        }
    }
}

__global__ void __launch_bounds__(256)
frag_word_heads_kernel(const uint64_t *__restrict__ w0m, const uint64_t *__restrict__ w1m,
                       const unsigned long long *__restrict__ n_frag, uint64_t cap, uint8_t *__restrict__ whead)
{
    const uint64_t F = *n_frag < cap ? *n_frag : cap;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < F; q += stride)
        whead[q] = (q == 0 || w0m[q] != w0m[q - 1] || w1m[q] != w1m[q - 1]) ? 1 : 0;
}

// Write the fragments into the ambiguous slots of the sorted order.  keys_sorted: the fully sorted radix keys
// (the kernels do nothing when *d_descent != 0: the caller then repairs the order element-wise).
// d_err (device int): bits 2/4 are set on inconsistency and make the caller fall back as well.
int frag_expand_device(FragSorted &fs, const uint64_t *keys_sorted, uint64_t n, int idx_bytes, void *d_idx,
                       uint8_t *d_flags, const unsigned int *d_descent, const unsigned long long *d_n_amb_expected,
                       uint64_t n_amb_host, int *d_err, cudaStream_t st)
{
    if (fs.F == 0) return GK_OK;
    frag_slots_kernel<<<frag_grid(fs.F), 256, 0, st>>>(fs.skey, fs.off.as<unsigned long long>(), fs.F, keys_sorted,
                                                       n, d_descent, d_n_amb_expected,
                                                       fs.slot0.as<unsigned long long>(), d_err);
    GK_LAUNCH_CHECK();
    const int grid = frag_grid(n_amb_host ? (n_amb_host + 7) / 8 : 1);
    if (idx_bytes == 4)
        frag_expand_kernel<uint32_t><<<grid, 256, 0, st>>>(fs.sstart.as<uint64_t>(), fs.off.as<unsigned long long>(),
                                                           fs.slot0.as<unsigned long long>(), fs.whead.as<uint8_t>(),
                                                           fs.F, n, d_descent, d_err, (uint32_t *)d_idx, d_flags);
    else
        frag_expand_kernel<uint64_t><<<grid, 256, 0, st>>>(fs.sstart.as<uint64_t>(), fs.off.as<unsigned long long>(),
                                                           fs.slot0.as<unsigned long long>(), fs.whead.as<uint8_t>(),
                                                           fs.F, n, d_descent, d_err, (uint64_t *)d_idx, d_flags);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

// ---- multi-GPU: the fragments of a shard ---------------------------------------------------------------------
// Every rank lists the fragments of its slice of the byte array (gk_pack_slice) and the lists are all-gathered;
// a rank keeps those whose key falls into its key range [key_lo, key_hi) (key_hi == 0: no upper bound), with
// the key made relative to key_lo like the pairs it received.  Order does not matter (the list is sorted later).
__global__ void __launch_bounds__(256)
frag_filter_kernel(const unsigned char *__restrict__ gathered, const unsigned long long *__restrict__ counts,
                   uint32_t world, uint64_t cap, uint64_t key_lo, uint64_t key_hi, FragOut out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t total = (uint64_t)world * cap;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const uint32_t src = (uint32_t)(t / cap);
        const uint64_t j = t - (uint64_t)src * cap;
        if (j >= counts[src]) continue;
        const uint64_t *b = reinterpret_cast<const uint64_t *>(gathered + (size_t)src * cap * 36);
        const uint64_t key = b[j];
        if (key < key_lo || (key_hi && key >= key_hi)) continue;
        const unsigned long long slot = atomicAdd(out.counter, 1ull);
        if (slot < out.capacity) {
            out.key[slot] = key - key_lo;
            out.w0[slot] = b[cap + j];
            out.w1[slot] = b[2 * cap + j];
            out.start[slot] = b[3 * cap + j];
            out.count[slot] = reinterpret_cast<const uint32_t *>(b + 4 * cap)[j];
        }
    }
}

// one CTA: off[q] = sum of count[0 .. q), off[F] = total, F read from the device
__global__ void __launch_bounds__(1024)
frag_scan_counts_kernel(const uint32_t *__restrict__ count, const unsigned long long *__restrict__ n_frag,
                        uint64_t capacity, unsigned long long *__restrict__ off)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const uint64_t F = *n_frag < capacity ? *n_frag : capacity;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint64_t base = 0; base < F; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long c = (i < F) ? count[i] : 0;
        unsigned long long inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += v;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        unsigned long long pre = s_carry;
        for (uint32_t w = 0; w < warp; ++w) pre += s_warp[w];
        if (i < F) off[i] = pre + inc - c;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = pre + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[F] = s_carry;
}

// The ambiguous windows of the shard as (key, start) pairs behind the n_pure received ones: they go through the
// local sort like any other pair and so mark the slots that frag_expand_device fills in afterwards.
// err |= 8 when the fragments do not add up to the n_amb windows the ranks counted for this key range.
template <typename IdxT>
__global__ void __launch_bounds__(256)
frag_placeholders_kernel(FragOut frag, const unsigned long long *__restrict__ off, uint64_t n_pure, uint64_t n_amb,
                         uint64_t *__restrict__ keys, IdxT *__restrict__ idx, int *__restrict__ err)
{
    const uint64_t F = *frag.counter < frag.capacity ? *frag.counter : frag.capacity;
    const uint64_t m = off[F];
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid == 0 && (m != n_amb || *frag.counter > frag.capacity)) atomicOr(err, 8);
    const uint64_t lim = m < n_amb ? m : n_amb;
    // as frag_expand_kernel: a warp writes 256 consecutive pairs, lane-strided, after one binary search
    constexpr int kPer = 8;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp_id = tid >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t o0 = warp_id * (32 * kPer); o0 < lim; o0 += n_warps * (32 * kPer)) {
        uint64_t lo = 0, hi = F;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (o0 < off[mid]) hi = mid; else lo = mid + 1;
        }
        uint64_t q = lo - 1;
        uint64_t q_off = off[q], q_end = off[q + 1], q_key = frag.key[q], q_start = frag.start[q];
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const uint64_t o = o0 + (uint64_t)u * 32 + lane;
            if (o >= lim) break;
            if (o >= q_end) {
                do {
                    ++q;
                    q_off = q_end;
                    q_end = off[q + 1];
                } while (o >= q_end);
                q_key = frag.key[q]; q_start = frag.start[q];
            }
            keys[n_pure + o] = q_key;
            idx[n_pure + o] = (IdxT)(q_start + (o - q_off));
        }
    }
}

int frag_filter_device(const void *d_gathered, const unsigned long long *d_counts, uint32_t world, uint64_t cap,
                       uint64_t key_lo, uint64_t key_hi, const FragOut &out, cudaStream_t st)
{
    if (world == 0 || cap == 0) return GK_OK;
    frag_filter_kernel<<<frag_grid((uint64_t)world * cap), 256, 0, st>>>((const unsigned char *)d_gathered, d_counts,
                                                                         world, cap, key_lo, key_hi, out);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int frag_placeholders_device(const FragOut &frag, unsigned long long *d_off, uint64_t n_pure, uint64_t n_amb,
                             uint64_t *d_keys, void *d_idx, int idx_bytes, int *d_err, cudaStream_t st)
{
    frag_scan_counts_kernel<<<1, 1024, 0, st>>>(frag.count, frag.counter, frag.capacity, d_off);
    GK_LAUNCH_CHECK();
    const int grid = frag_grid(n_amb ? (n_amb + 7) / 8 : 1);
    if (idx_bytes == 4)
        frag_placeholders_kernel<uint32_t><<<grid, 256, 0, st>>>(frag, d_off, n_pure, n_amb, d_keys, (uint32_t *)d_idx,
                                                                 d_err);
    else
        frag_placeholders_kernel<uint64_t><<<grid, 256, 0, st>>>(frag, d_off, n_pure, n_amb, d_keys, (uint64_t *)d_idx,
                                                                 d_err);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

// The sorted fragment list of a key range from the gathered, individually sorted lists of all ranks.  Leaves
// `fs` as frag_sort_device would (fs.F is set by the caller once the count is known on the host), the number of
// fragments in *d_n_frag, and a view of the merged list (key, start, count) for the placeholder pairs.
int frag_merge_device(const void *d_gathered, const unsigned long long *d_counts, uint32_t world, uint64_t cap,
                      uint64_t key_lo, uint64_t key_hi, uint64_t out_cap, unsigned long long *d_n_frag,
                      FragSorted &fs, FragOut *view, cudaStream_t st)
{
    if (world == 0 || world > 32 || cap == 0) return GK_ERR_ARG;
    GK_TRY(fs.reserve(out_cap, st));
    DeviceBuffer range, w1m;
    GK_TRY(range.alloc((size_t)world * 16, st));
    GK_TRY(w1m.alloc((size_t)out_cap * 8, st));
    frag_ranges_kernel<<<1, 32, 0, st>>>((const unsigned char *)d_gathered, d_counts, world, cap, key_lo, key_hi,
                                         range.as<unsigned long long>(), d_n_frag);
    GK_LAUNCH_CHECK();
    uint64_t *skey = fs.skey_a.as<uint64_t>(), *w0m = fs.skey_b.as<uint64_t>();
    uint32_t *count = fs.perm_a.as<uint32_t>();
    frag_merge_kernel<<<frag_grid((uint64_t)world * cap), 256, 0, st>>>(
        (const unsigned char *)d_gathered, world, cap, range.as<unsigned long long>(), key_lo, out_cap, skey, w0m,
        w1m.as<uint64_t>(), fs.sstart.as<uint64_t>(), count);
    GK_LAUNCH_CHECK();
    frag_word_heads_kernel<<<frag_grid(out_cap), 256, 0, st>>>(w0m, w1m.as<uint64_t>(), d_n_frag, out_cap,
                                                               fs.whead.as<uint8_t>());
    GK_LAUNCH_CHECK();
    frag_scan_counts_kernel<<<1, 1024, 0, st>>>(count, d_n_frag, out_cap, fs.off.as<unsigned long long>());
    GK_LAUNCH_CHECK();
    fs.skey = skey;
    if (view) {
        view->key = skey;
        view->w0 = w0m;
        view->w1 = nullptr;
        view->start = fs.sstart.as<uint64_t>();
        view->count = count;
        view->counter = d_n_frag;
        view->capacity = out_cap;
    }
    return GK_OK;
}

}  // namespace gk
