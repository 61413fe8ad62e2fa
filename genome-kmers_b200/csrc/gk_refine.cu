// gk_refine.cu -- kernels of the two refinement stages that follow the main sort (SURVEY.md 8a A4,
// north_star subsystem 2 "extended with prefix-doubling refinement"); orchestration in gk_index.cu.
//
// 1. Prefix doubling (k-mers longer than one key word, variable-length mode).  After the windows are
//    sorted by their first h symbols (head flags mark the h-groups), the rank of a start is the sorted
//    position of its group's first member.  A window of h2 <= 2h symbols is then ordered by the pair
//    (rank_h[start], rank_h[start + h2 - h]); a window that ends at its record's '$' before start + h2 - h
//    has an empty second half, which sorts first.  Only members of groups with more than one element can
//    move, so the rounds run on that shrinking list and the rest keeps its slot.
//      head_positions     group id (= position of the group's head) for every sorted position, and
//                         rank_of_start[idx[p]] = that id.  Tile max-scan, 3 kernels.
//      valid_flags        which windows fit their record at a given length (final drop of short windows)
//      gid_flags          head flags from group ids + "group has more than one member" marks
//      pair_keys_var      key2 = (group << 32) | (rank_of_start[start + delta] + 1, or 0 when empty)
//      key2_scatter       re-sorted list -> slots, with fresh head flags
//      subset_rank_update ranks of the list members after a round
// 2. Run-length compressed stable sort of a selected subset (ambiguous windows, out-of-order long prefix
//    runs): subset_rep_flags, block offsets (scan), subset_expand -- see below.
// Indices are 32-bit in stage 1: the multi-level path is limited to byte arrays below 2^32 positions.
#include "gk_common.cuh"

namespace gk {


constexpr int kHpThreads = 256;
constexpr int kHpPerThread = 16;
constexpr int kHpTile = kHpThreads * kHpPerThread;

// last head position inside each tile (kNoHead if none)
__global__ void __launch_bounds__(kHpThreads)
tile_last_head_kernel(const uint8_t *__restrict__ flags, uint64_t n, uint32_t *__restrict__ tile_last)
{
    __shared__ uint32_t s_max[kHpThreads / 32];
    const uint64_t p0 = (uint64_t)blockIdx.x * kHpTile + (uint64_t)threadIdx.x * kHpPerThread;
    uint32_t last = 0;  // store position+1 so that 0 means "none"
    for (int i = 0; i < kHpPerThread; ++i) {
        const uint64_t p = p0 + i;
        if (p < n && (flags[p] & kFlagHead)) last = (uint32_t)p + 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t v = __shfl_xor_sync(0xffffffffu, last, o);
        last = v > last ? v : last;
    }
    if (lane_id() == 0) s_max[threadIdx.x >> 5] = last;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t m = 0;
        for (int w = 0; w < kHpThreads / 32; ++w) m = s_max[w] > m ? s_max[w] : m;
        tile_last[blockIdx.x] = m;
    }
}

// one CTA: carry[t] = last head (+1) in tiles < t (exclusive running max)
__global__ void __launch_bounds__(1024)
tile_carry_kernel(const uint32_t *__restrict__ tile_last, uint64_t n_tiles, uint32_t *__restrict__ carry)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint64_t base = 0; base < n_tiles; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint32_t v = (i < n_tiles) ? tile_last[i] : 0;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc = u > inc ? u : inc;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t pre = s_carry;
        for (uint32_t w = 0; w < warp; ++w) pre = s_warp[w] > pre ? s_warp[w] : pre;
        // exclusive: max over everything strictly before i
        uint32_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0;
        excl = excl > pre ? excl : pre;
        if (i < n_tiles) carry[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = inc > pre ? inc : pre;
        __syncthreads();
    }
}

// gid[p] = position of the head of p's group; optionally rank_of_start[idx[p]] = gid[p]
__global__ void __launch_bounds__(kHpThreads)
head_positions_kernel(const uint8_t *__restrict__ flags, const uint32_t *__restrict__ idx, uint64_t n,
                      const uint32_t *__restrict__ carry, uint32_t *__restrict__ gid,
                      uint32_t *__restrict__ rank_of_start)
{
    __shared__ uint32_t s_warp[kHpThreads / 32];
    const uint64_t p0 = (uint64_t)blockIdx.x * kHpTile + (uint64_t)threadIdx.x * kHpPerThread;
    uint32_t local[kHpPerThread];
    uint32_t run = 0;  // position+1 of the latest head seen by this thread
#pragma unroll
    for (int i = 0; i < kHpPerThread; ++i) {
        const uint64_t p = p0 + i;
        if (p < n && (flags[p] & kFlagHead)) run = (uint32_t)p + 1;
        local[i] = run;
    }
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc = u > inc ? u : inc;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t pre = carry[blockIdx.x];
    for (uint32_t w = 0; w < warp; ++w) pre = s_warp[w] > pre ? s_warp[w] : pre;
    uint32_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = 0;
    excl = excl > pre ? excl : pre;
#pragma unroll
    for (int i = 0; i < kHpPerThread; ++i) {
        const uint64_t p = p0 + i;
        if (p < n) {
            const uint32_t h = (local[i] > excl ? local[i] : excl) - 1;  // position 0 is always a head
            gid[p] = h;
            if (rank_of_start) rank_of_start[idx[p]] = h;
        }
    }
}

// Segment tables up to this size are staged in shared memory by the per-element kernels below (a binary
// search per element through global memory is what made them slow); larger tables are searched in place.
constexpr uint32_t kSegSmem = 1024;

__device__ __forceinline__ const uint64_t *stage_segments(const uint64_t *__restrict__ seg_starts, uint32_t n_seg,
                                                          uint64_t *s_seg)
{
    if (n_seg > kSegSmem) return seg_starts;
    for (uint32_t i = threadIdx.x; i < n_seg; i += blockDim.x) s_seg[i] = seg_starts[i];
    __syncthreads();
    return s_seg;
}

// flags[p] = kFlagPass iff the window of `len` symbols at idx[p] stays inside its record.
// Elements [first, n), one per thread (the tail of the vector kernel, or everything).
__global__ void __launch_bounds__(256)
valid_flags_kernel(const uint32_t *__restrict__ idx, uint64_t first, uint64_t n,
                   const uint64_t *__restrict__ seg_starts, uint32_t n_seg, uint64_t sba_len, uint32_t len,
                   uint8_t *__restrict__ flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const uint64_t s = idx[p];
        const uint32_t seg = upper_seg(seg_starts, n_seg, s);
        const uint64_t seg_end = (seg + 1 < n_seg) ? seg_starts[seg + 1] - 1 : sba_len;
        flags[p] = (s + len <= seg_end) ? kFlagPass : 0;
    }
}

// the same for groups of four starts: one 16-byte load, one 4-byte store per thread and group
__global__ void __launch_bounds__(256)
valid_flags_vec_kernel(const uint32_t *__restrict__ idx, uint64_t groups, const uint64_t *__restrict__ seg_starts,
                       uint32_t n_seg, uint64_t sba_len, uint32_t len, uint8_t *__restrict__ flags)
{
    __shared__ uint64_t s_seg[kSegSmem];
    const uint64_t *segs = stage_segments(seg_starts, n_seg, s_seg);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        const uint4 q = reinterpret_cast<const uint4 *>(idx)[g];
        const uint32_t v[4] = {q.x, q.y, q.z, q.w};
        uint32_t out = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint64_t s = v[i];
            const uint32_t seg = upper_seg(segs, n_seg, s);
            const uint64_t seg_end = (seg + 1 < n_seg) ? segs[seg + 1] - 1 : sba_len;
            out |= ((s + len <= seg_end) ? (uint32_t)kFlagPass : 0u) << (8 * i);
        }
        reinterpret_cast<uint32_t *>(flags)[g] = out;
    }
}

// head flag where the group id changes; kFlagMulti on every member of a group with > 1 element
__global__ void __launch_bounds__(256)
gid_flags_kernel(const uint32_t *__restrict__ gid, uint64_t n, uint8_t *__restrict__ flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const uint32_t g = gid[p];
        const bool head = (p == 0) || gid[p - 1] != g;
        const bool next_head = (p + 1 == n) || gid[p + 1] != g;
        flags[p] = (head ? kFlagHead : 0) | ((head && next_head) ? 0 : kFlagMulti);
    }
}

// Variable-length mode (kmers.py:360-378): a window ends at its record's '$'.  When start + delta is at or
// past that '$' the second half is empty and sorts first (0); real ranks are shifted up by one.
__global__ void __launch_bounds__(256)
pair_keys_var_kernel(const uint32_t *__restrict__ sub_idx, const uint32_t *__restrict__ sub_gid, uint64_t m,
                     const uint32_t *__restrict__ rank_of_start, uint32_t delta,
                     const uint64_t *__restrict__ seg_starts, uint32_t n_seg, uint64_t sba_len,
                     uint64_t *__restrict__ keys)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) {
        const uint64_t s = sub_idx[r];
        const uint32_t seg = upper_seg(seg_starts, n_seg, s);
        const uint64_t seg_end = (seg + 1 < n_seg) ? seg_starts[seg + 1] - 1 : sba_len;  // position of '$'
        const uint64_t q = s + delta;
        const uint32_t second = (q < seg_end) ? rank_of_start[q] + 1u : 0u;
        keys[r] = ((uint64_t)sub_gid[r] << 32) | second;
    }
}

// Word rounds (multi-GPU shards: the ranks of other starts live on other GPUs): the second half of the pair
// is read from the bytes instead -- the 4-bit ranks of the `span` <= 8 symbols from start + h on, most
// significant first; a '$' / the end of the array ends the k-mer (all later symbols 0, kmers.py:360-378).
template <typename IdxT>
__global__ void __launch_bounds__(256)
pair_keys_words_kernel(const IdxT *__restrict__ sub_idx, const uint32_t *__restrict__ sub_gid, uint64_t m,
                       const uint8_t *__restrict__ sba, uint64_t sba_len, uint64_t h, uint32_t span,
                       uint64_t *__restrict__ keys)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) {
        const uint64_t q = (uint64_t)sub_idx[r] + h;
        uint32_t second = 0;
        for (uint32_t j = 0; j < span; ++j) {
            const uint32_t code = (q + j < sba_len) ? rank4(sba[q + j]) : 0u;
            if (code == 0) break;
            second |= code << (4u * (7u - j));
        }
        keys[r] = ((uint64_t)sub_gid[r] << 32) | second;
    }
}

// the re-sorted subset goes back to its slots with fresh head flags
template <typename IdxT>
__global__ void __launch_bounds__(256)
key2_scatter_kernel(const uint64_t *__restrict__ keys_sorted, const IdxT *__restrict__ sub_idx_sorted,
                    const uint32_t *__restrict__ slots, uint64_t m, IdxT *__restrict__ idx,
                    uint8_t *__restrict__ flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) {
        const bool head = (r == 0) || keys_sorted[r] != keys_sorted[r - 1];
        const uint32_t slot = slots[r];
        idx[slot] = sub_idx_sorted[r];
        flags[slot] = head ? kFlagHead : 0;
    }
}

// ---- run-length compressed stable sort of a selected subset ---------------------------------------------
// After the main sort, two kinds of slots may still hold the wrong element: members of long prefix runs
// that are out of order, and ambiguous windows (ordered by `value` only, not yet by their real symbols).
// Both sets are dominated by huge blocks of identical elements (every window inside an N run, every copy
// of an exact repeat), so the subset is sorted in compressed form:
//   subset_rep_flags   r starts a block iff its (key, w0, w1) differs from element r-1 of the subset
//   (select)           block representatives: first subset position of every block
//   (LSD sorts)        the representatives only, stable, least significant word first
//   block offsets      exclusive scan of the block lengths in sorted order
//   subset_expand      every output position finds its block by binary search and copies the start index
//                      into the slot; the head flag is set at the first element of a block whose words
//                      differ from the previous block's
// Cost is proportional to the number of blocks, not to the number of elements.

__global__ void __launch_bounds__(256)
subset_rep_flags_kernel(const uint64_t *__restrict__ key, const uint64_t *__restrict__ w0,
                        const uint64_t *__restrict__ w1, uint64_t m, uint8_t *__restrict__ rh)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) {
        bool head = (r == 0) || key[r] != key[r - 1];
        if (!head && w0) head = w0[r] != w0[r - 1];
        if (!head && w1) head = w1[r] != w1[r - 1];
        rh[r] = head ? kFlagHead : 0;
    }
}

// out[q] = src[at[perm ? perm[q] : q]]
template <typename T>
__global__ void __launch_bounds__(256)
gather2_u64_kernel(const uint64_t *__restrict__ src, const T *__restrict__ at, const T *__restrict__ perm,
                   uint64_t count, uint64_t *__restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < count; q += stride) {
        const uint64_t j = perm ? (uint64_t)perm[q] : q;
        out[q] = src[(uint64_t)at[j]];
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
iota_kernel(T *__restrict__ out, uint64_t count)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < count; q += stride) out[q] = (T)q;
}

constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 8;
constexpr int kScanTile = kScanThreads * kScanPerThread;

template <typename T>
__device__ __forceinline__ unsigned long long block_len(const T *__restrict__ rep_start, uint64_t R, uint64_t m,
                                                        const T *__restrict__ perm, uint64_t q)
{
    const uint64_t j = (uint64_t)perm[q];
    const uint64_t a = (uint64_t)rep_start[j];
    const uint64_t b = (j + 1 < R) ? (uint64_t)rep_start[j + 1] : m;
    return b - a;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads)
block_len_tile_sums_kernel(const T *__restrict__ rep_start, uint64_t R, uint64_t m, const T *__restrict__ perm,
                           unsigned long long *__restrict__ tile_sums)
{
    __shared__ unsigned long long s_warp[kScanThreads / 32];
    const uint64_t q0 = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanPerThread;
    unsigned long long sum = 0;
    for (int i = 0; i < kScanPerThread; ++i)
        if (q0 + i < R) sum += block_len(rep_start, R, m, perm, q0 + i);
    sum = warp_sum(sum);
    if (lane_id() == 0) s_warp[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += s_warp[w];
        tile_sums[blockIdx.x] = t;
    }
}

// one CTA: exclusive scan of 64-bit tile sums in place
__global__ void __launch_bounds__(1024)
scan_u64_kernel(unsigned long long *__restrict__ data, uint64_t count)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint64_t base = 0; base < count; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long c = (i < count) ? data[i] : 0;
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
        if (i < count) data[i] = pre + inc - c;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = pre + inc;
        __syncthreads();
    }
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads)
block_offsets_kernel(const T *__restrict__ rep_start, uint64_t R, uint64_t m, const T *__restrict__ perm,
                     const unsigned long long *__restrict__ tile_offsets,
                     unsigned long long *__restrict__ out_off)
{
    __shared__ unsigned long long s_warp[kScanThreads / 32];
    const uint64_t q0 = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanPerThread;
    unsigned long long len[kScanPerThread];
    unsigned long long sum = 0;
#pragma unroll
    for (int i = 0; i < kScanPerThread; ++i) {
        len[i] = (q0 + i < R) ? block_len(rep_start, R, m, perm, q0 + i) : 0;
        sum += len[i];
    }
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    unsigned long long inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += v;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned long long run = tile_offsets[blockIdx.x] + inc - sum;
    for (uint32_t w = 0; w < warp; ++w) run += s_warp[w];
#pragma unroll
    for (int i = 0; i < kScanPerThread; ++i) {
        if (q0 + i < R) out_off[q0 + i] = run;
        run += len[i];
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
subset_expand_kernel(const T *__restrict__ pos, const T *__restrict__ idx_sel, const uint64_t *__restrict__ key,
                     const uint64_t *__restrict__ w0, const uint64_t *__restrict__ w1,
                     const T *__restrict__ rep_start, const T *__restrict__ perm,
                     const unsigned long long *__restrict__ out_off, uint64_t R, uint64_t m, int class_bit,
                     T *__restrict__ d_idx, uint8_t *__restrict__ d_flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < m; o += stride) {
        uint64_t lo = 0, hi = R;  // last block whose offset is <= o
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (o < out_off[mid]) hi = mid; else lo = mid + 1;
        }
        const uint64_t q = lo - 1;
        const uint64_t src0 = (uint64_t)rep_start[(uint64_t)perm[q]];
        const uint64_t src = src0 + (o - out_off[q]);
        const uint64_t slot = (uint64_t)pos[o];
        bool head = (o == 0) || ((uint64_t)pos[o - 1] + 1 != slot);  // the slot before is not in the subset
        if (!head && o == out_off[q]) {
            const uint64_t prv = (uint64_t)rep_start[(uint64_t)perm[q - 1]];
            head = key[src0] != key[prv] || (w0 && w0[src0] != w0[prv]) || (w1 && w1[src0] != w1[prv]);
        }
        const bool amb = class_bit && !(key[src] & 1ull);
        d_idx[slot] = idx_sel[src];
        d_flags[slot] = (amb ? kFlagAmb : 0) | (head ? kFlagHead : 0);
    }
}

// Long prefix runs re-sorted by their full keys (gk_index.cu: repair_long_runs): the sorted members go back to
// the slots of the set, in order, with fresh flags; the sorted keys are written back too, so that the fragment
// expansion can look its keys up afterwards.
template <typename T>
__global__ void __launch_bounds__(256)
scatter_sorted_subset_kernel(const T *__restrict__ pos, const uint64_t *__restrict__ keys_sorted,
                             const T *__restrict__ idx_sorted, uint64_t m, int class_bit,
                             uint64_t *__restrict__ keys, T *__restrict__ idx, uint8_t *__restrict__ flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) {
        const uint64_t k = keys_sorted[r];
        const bool head = (r == 0) || keys_sorted[r - 1] != k;
        const bool amb = class_bit && !(k & 1ull);
        const uint64_t p = (uint64_t)pos[r];
        keys[p] = k;
        idx[p] = idx_sorted[r];
        flags[p] = amb ? kFlagAmb : (head ? kFlagHead : 0);
    }
}

// subset bookkeeping of the doubling rounds: global slot of every member's group head, and the new rank of
// every member's start (rank = sorted position of the first member of its group)
__global__ void __launch_bounds__(256)
subset_rank_update_kernel(const uint32_t *__restrict__ slots, const uint32_t *__restrict__ gid_sub,
                          const uint32_t *__restrict__ idx_sorted, uint64_t m, uint32_t *__restrict__ gid_slot,
                          uint32_t *__restrict__ rank_of_start)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) {
        const uint32_t g = slots[gid_sub[r]];
        gid_slot[r] = g;
        if (rank_of_start) rank_of_start[idx_sorted[r]] = g;
    }
}

// kFlagMulti on every member of a group with more than one element, straight from the head flags.
// Elements [first, n), one per thread (the tail of the vector kernel, or everything).
__global__ void __launch_bounds__(256)
multi_flags_kernel(const uint8_t *__restrict__ flags, uint64_t first, uint64_t n, uint8_t *__restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const bool head = flags[p] & kFlagHead;
        const bool next_head = (p + 1 == n) || (flags[p + 1] & kFlagHead);
        out[p] = (head ? kFlagHead : 0) | ((head && next_head) ? 0 : kFlagMulti);
    }
}

// the same for groups of sixteen flags: one 16-byte load (+ the next group's first flag), one 16-byte store
__global__ void __launch_bounds__(256)
multi_flags_vec_kernel(const uint8_t *__restrict__ flags, uint64_t groups, uint64_t n, uint8_t *__restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        const uint4 q = reinterpret_cast<const uint4 *>(flags)[g];
        const uint64_t p0 = g * 16;
        const uint32_t next = (p0 + 16 < n) ? flags[p0 + 16] : (uint32_t)kFlagHead;
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint32_t heads = (next & kFlagHead) << 16;  // bit i: flag i of the group (bit 16: the next group's first)
#pragma unroll
        for (int i = 0; i < 16; ++i) heads |= ((w[i >> 2] >> (8 * (i & 3))) & (uint32_t)kFlagHead) << i;
        uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const bool head = (heads >> i) & 1u, next_head = (heads >> (i + 1)) & 1u;
            const uint32_t f = (head ? (uint32_t)kFlagHead : 0u) | ((head && next_head) ? 0u : (uint32_t)kFlagMulti);
            o[i >> 2] |= f << (8 * (i & 3));
        }
        reinterpret_cast<uint4 *>(out)[g] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// A window that reaches its record's '$' within the first h symbols is fully compared: its group holds
// identical k-mers and can never split again.  Clear kFlagMulti on those members.
// Elements [first, n), one per thread (the tail of the vector kernel, or everything).
__global__ void __launch_bounds__(256)
clear_finished_multi_kernel(const uint32_t *__restrict__ idx, uint64_t first, uint64_t n,
                            const uint64_t *__restrict__ seg_starts, uint32_t n_seg, uint64_t sba_len, uint64_t h,
                            uint8_t *__restrict__ flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const uint8_t f = flags[p];
        if (!(f & kFlagMulti)) continue;
        const uint64_t s = idx[p];
        const uint32_t seg = upper_seg(seg_starts, n_seg, s);
        const uint64_t seg_end = (seg + 1 < n_seg) ? seg_starts[seg + 1] - 1 : sba_len;  // position of '$'
        if (s + h > seg_end) flags[p] = f & (uint8_t)~kFlagMulti;
    }
}

// the same for groups of sixteen flags: a group without a marked member costs one 16-byte load
__global__ void __launch_bounds__(256)
clear_finished_multi_vec_kernel(const uint32_t *__restrict__ idx, uint64_t groups,
                                const uint64_t *__restrict__ seg_starts, uint32_t n_seg, uint64_t sba_len,
                                uint64_t h, uint8_t *__restrict__ flags)
{
    __shared__ uint64_t s_seg[kSegSmem];
    const uint64_t *segs = stage_segments(seg_starts, n_seg, s_seg);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        const uint4 q = reinterpret_cast<const uint4 *>(flags)[g];
        const uint32_t multi = (uint32_t)kFlagMulti * 0x01010101u;
        if (((q.x | q.y | q.z | q.w) & multi) == 0) continue;
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        const uint64_t p0 = g * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint8_t f = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
            if (!(f & kFlagMulti)) continue;
            const uint64_t s = idx[p0 + i];
            const uint32_t seg = upper_seg(segs, n_seg, s);
            const uint64_t seg_end = (seg + 1 < n_seg) ? segs[seg + 1] - 1 : sba_len;
            if (s + h > seg_end) flags[p0 + i] = f & (uint8_t)~kFlagMulti;  // only this thread touches the group
        }
    }
}

__global__ void __launch_bounds__(256)
gather_u32_kernel(const uint32_t *__restrict__ src, const uint32_t *__restrict__ at, uint64_t count,
                  uint32_t *__restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < count; r += stride) out[r] = src[at[r]];
}

static int grid_for(uint64_t items)
{
    uint64_t blocks = (items + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int head_positions_device(const uint8_t *d_flags, const uint32_t *d_idx, uint64_t n, uint32_t *d_gid,
                          uint32_t *d_rank_of_start, cudaStream_t st)
{
    if (n == 0) return GK_OK;
    const uint64_t tiles = (n + kHpTile - 1) / kHpTile;
    DeviceBuffer temp;
    GK_TRY(temp.alloc((size_t)tiles * 8, st));
    uint32_t *d_last = temp.as<uint32_t>();
    uint32_t *d_carry = d_last + tiles;
    tile_last_head_kernel<<<(unsigned)tiles, kHpThreads, 0, st>>>(d_flags, n, d_last);
    GK_LAUNCH_CHECK();
    tile_carry_kernel<<<1, 1024, 0, st>>>(d_last, tiles, d_carry);
    GK_LAUNCH_CHECK();
    head_positions_kernel<<<(unsigned)tiles, kHpThreads, 0, st>>>(d_flags, d_idx, n, d_carry, d_gid,
                                                                  d_rank_of_start);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int valid_flags_device(const uint32_t *d_idx, uint64_t n, const uint64_t *d_seg_starts, uint32_t n_seg,
                       uint64_t sba_len, uint32_t len, uint8_t *d_flags, cudaStream_t st)
{
    if (n == 0) return GK_OK;
    // vector kernel over the 16-byte aligned bulk, scalar kernel over the tail (or over everything)
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_idx) & 15u) | (reinterpret_cast<uintptr_t>(d_flags) & 3u)) == 0;
    const uint64_t groups = aligned ? n / 4 : 0;
    if (groups) {
        valid_flags_vec_kernel<<<grid_for(groups), 256, 0, st>>>(d_idx, groups, d_seg_starts, n_seg, sba_len, len,
                                                                  d_flags);
        GK_LAUNCH_CHECK();
    }
    if (groups * 4 < n) {
        valid_flags_kernel<<<grid_for(n - groups * 4), 256, 0, st>>>(d_idx, groups * 4, n, d_seg_starts, n_seg,
                                                                     sba_len, len, d_flags);
        GK_LAUNCH_CHECK();
    }
    return GK_OK;
}

int gid_flags_device(const uint32_t *d_gid, uint64_t n, uint8_t *d_flags, cudaStream_t st)
{
    if (n == 0) return GK_OK;
    gid_flags_kernel<<<grid_for(n), 256, 0, st>>>(d_gid, n, d_flags);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int pair_keys_var_device(const uint32_t *d_sub_idx, const uint32_t *d_sub_gid, uint64_t m,
                         const uint32_t *d_rank_of_start, uint32_t delta, const uint64_t *d_seg_starts,
                         uint32_t n_seg, uint64_t sba_len, uint64_t *d_keys, cudaStream_t st)
{
    if (m == 0) return GK_OK;
    pair_keys_var_kernel<<<grid_for(m), 256, 0, st>>>(d_sub_idx, d_sub_gid, m, d_rank_of_start, delta,
                                                      d_seg_starts, n_seg, sba_len, d_keys);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int pair_keys_words_device(const void *d_sub_idx, int idx_bytes, const uint32_t *d_sub_gid, uint64_t m,
                           const uint8_t *d_sba, uint64_t sba_len, uint64_t h, uint32_t span, uint64_t *d_keys,
                           cudaStream_t st)
{
    if (m == 0) return GK_OK;
    if (idx_bytes == 4)
        pair_keys_words_kernel<uint32_t><<<grid_for(m), 256, 0, st>>>((const uint32_t *)d_sub_idx, d_sub_gid, m, d_sba,
                                                                      sba_len, h, span, d_keys);
    else
        pair_keys_words_kernel<uint64_t><<<grid_for(m), 256, 0, st>>>((const uint64_t *)d_sub_idx, d_sub_gid, m, d_sba,
                                                                      sba_len, h, span, d_keys);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

// key2_scatter for either width of start index (the word rounds run on 64-bit starts too)
int key2_scatter_any_device(const uint64_t *d_keys_sorted, const void *d_sub_idx_sorted, const uint32_t *d_slots,
                            uint64_t m, int idx_bytes, void *d_idx, uint8_t *d_flags, cudaStream_t st)
{
    if (m == 0) return GK_OK;
    if (idx_bytes == 4)
        key2_scatter_kernel<uint32_t><<<grid_for(m), 256, 0, st>>>(d_keys_sorted, (const uint32_t *)d_sub_idx_sorted,
                                                                   d_slots, m, (uint32_t *)d_idx, d_flags);
    else
        key2_scatter_kernel<uint64_t><<<grid_for(m), 256, 0, st>>>(d_keys_sorted, (const uint64_t *)d_sub_idx_sorted,
                                                                   d_slots, m, (uint64_t *)d_idx, d_flags);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int scatter_sorted_subset_device(const void *d_pos, const uint64_t *d_keys_sorted, const void *d_idx_sorted, uint64_t m,
                                 int class_bit, int t_bytes, uint64_t *d_keys, void *d_idx, uint8_t *d_flags,
                                 cudaStream_t st)
{
    if (m == 0) return GK_OK;
    if (t_bytes == 4)
        scatter_sorted_subset_kernel<uint32_t><<<grid_for(m), 256, 0, st>>>(
            (const uint32_t *)d_pos, d_keys_sorted, (const uint32_t *)d_idx_sorted, m, class_bit, d_keys,
            (uint32_t *)d_idx, d_flags);
    else
        scatter_sorted_subset_kernel<uint64_t><<<grid_for(m), 256, 0, st>>>(
            (const uint64_t *)d_pos, d_keys_sorted, (const uint64_t *)d_idx_sorted, m, class_bit, d_keys,
            (uint64_t *)d_idx, d_flags);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int subset_rank_update_device(const uint32_t *d_slots, const uint32_t *d_gid_sub, const uint32_t *d_idx_sorted,
                              uint64_t m, uint32_t *d_gid_slot, uint32_t *d_rank_of_start, cudaStream_t st)
{
    if (m == 0) return GK_OK;
    subset_rank_update_kernel<<<grid_for(m), 256, 0, st>>>(d_slots, d_gid_sub, d_idx_sorted, m, d_gid_slot,
                                                           d_rank_of_start);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int multi_flags_device(const uint8_t *d_flags, uint64_t n, uint8_t *d_out, cudaStream_t st)
{
    if (n == 0) return GK_OK;
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_flags) | reinterpret_cast<uintptr_t>(d_out)) & 15u) == 0;
    const uint64_t groups = aligned ? n / 16 : 0;
    if (groups) {
        multi_flags_vec_kernel<<<grid_for(groups), 256, 0, st>>>(d_flags, groups, n, d_out);
        GK_LAUNCH_CHECK();
    }
    if (groups * 16 < n) {
        multi_flags_kernel<<<grid_for(n - groups * 16), 256, 0, st>>>(d_flags, groups * 16, n, d_out);
        GK_LAUNCH_CHECK();
    }
    return GK_OK;
}

int clear_finished_multi_device(const uint32_t *d_idx, uint64_t n, const uint64_t *d_seg_starts, uint32_t n_seg,
                                uint64_t sba_len, uint64_t h, uint8_t *d_flags, cudaStream_t st)
{
    if (n == 0) return GK_OK;
    const bool aligned = (reinterpret_cast<uintptr_t>(d_flags) & 15u) == 0;
    const uint64_t groups = aligned ? n / 16 : 0;
    if (groups) {
        clear_finished_multi_vec_kernel<<<grid_for(groups), 256, 0, st>>>(d_idx, groups, d_seg_starts, n_seg, sba_len,
                                                                           h, d_flags);
        GK_LAUNCH_CHECK();
    }
    if (groups * 16 < n) {
        clear_finished_multi_kernel<<<grid_for(n - groups * 16), 256, 0, st>>>(d_idx, groups * 16, n, d_seg_starts,
                                                                               n_seg, sba_len, h, d_flags);
        GK_LAUNCH_CHECK();
    }
    return GK_OK;
}

int gather_u32_device(const uint32_t *d_src, const uint32_t *d_at, uint64_t count, uint32_t *d_out, cudaStream_t st)
{
    if (count == 0) return GK_OK;
    gather_u32_kernel<<<grid_for(count), 256, 0, st>>>(d_src, d_at, count, d_out);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int key2_scatter_device(const uint64_t *d_keys_sorted, const uint32_t *d_sub_idx_sorted,
                        const uint32_t *d_slots, uint64_t m, uint32_t *d_idx, uint8_t *d_flags,
                        cudaStream_t st)
{
    if (m == 0) return GK_OK;
    key2_scatter_kernel<uint32_t><<<grid_for(m), 256, 0, st>>>(d_keys_sorted, d_sub_idx_sorted, d_slots, m, d_idx,
                                                     d_flags);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int subset_rep_flags_device(const uint64_t *d_key, const uint64_t *d_w0, const uint64_t *d_w1, uint64_t m,
                            uint8_t *d_rh, cudaStream_t st)
{
    if (m == 0) return GK_OK;
    subset_rep_flags_kernel<<<grid_for(m), 256, 0, st>>>(d_key, d_w0, d_w1, m, d_rh);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int gather2_u64_device(const uint64_t *d_src, const void *d_at, const void *d_perm, uint64_t count, int t_bytes,
                       uint64_t *d_out, cudaStream_t st)
{
    if (count == 0) return GK_OK;
    if (t_bytes == 4)
        gather2_u64_kernel<uint32_t><<<grid_for(count), 256, 0, st>>>(d_src, (const uint32_t *)d_at,
                                                                      (const uint32_t *)d_perm, count, d_out);
    else
        gather2_u64_kernel<uint64_t><<<grid_for(count), 256, 0, st>>>(d_src, (const uint64_t *)d_at,
                                                                      (const uint64_t *)d_perm, count, d_out);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int iota_device(void *d_out, uint64_t count, int t_bytes, cudaStream_t st)
{
    if (count == 0) return GK_OK;
    if (t_bytes == 4) iota_kernel<uint32_t><<<grid_for(count), 256, 0, st>>>((uint32_t *)d_out, count);
    else iota_kernel<uint64_t><<<grid_for(count), 256, 0, st>>>((uint64_t *)d_out, count);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

// d_out_off[q] = sum of the lengths of the blocks perm[0..q)
template <typename T>
static int block_offsets_impl(const T *d_rep_start, uint64_t R, uint64_t m, const T *d_perm,
                              unsigned long long *d_out_off, cudaStream_t st)
{
    const uint64_t tiles = (R + kScanTile - 1) / kScanTile;
    DeviceBuffer sums;
    GK_TRY(sums.alloc((size_t)tiles * 8, st));
    block_len_tile_sums_kernel<T><<<(unsigned)tiles, kScanThreads, 0, st>>>(d_rep_start, R, m, d_perm,
                                                                            sums.as<unsigned long long>());
    GK_LAUNCH_CHECK();
    scan_u64_kernel<<<1, 1024, 0, st>>>(sums.as<unsigned long long>(), tiles);
    GK_LAUNCH_CHECK();
    block_offsets_kernel<T><<<(unsigned)tiles, kScanThreads, 0, st>>>(d_rep_start, R, m, d_perm,
                                                                      sums.as<unsigned long long>(), d_out_off);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int block_offsets_device(const void *d_rep_start, uint64_t R, uint64_t m, const void *d_perm, int t_bytes,
                         unsigned long long *d_out_off, cudaStream_t st)
{
    if (R == 0) return GK_OK;
    if (t_bytes == 4)
        return block_offsets_impl<uint32_t>((const uint32_t *)d_rep_start, R, m, (const uint32_t *)d_perm,
                                            d_out_off, st);
    return block_offsets_impl<uint64_t>((const uint64_t *)d_rep_start, R, m, (const uint64_t *)d_perm, d_out_off,
                                        st);
}

int subset_expand_device(const void *d_pos, const void *d_idx_sel, const uint64_t *d_key, const uint64_t *d_w0,
                         const uint64_t *d_w1, const void *d_rep_start, const void *d_perm,
                         const unsigned long long *d_out_off, uint64_t R, uint64_t m, int class_bit, int t_bytes,
                         void *d_idx, uint8_t *d_flags, cudaStream_t st)
{
    if (m == 0) return GK_OK;
    if (t_bytes == 4)
        subset_expand_kernel<uint32_t><<<grid_for(m), 256, 0, st>>>(
            (const uint32_t *)d_pos, (const uint32_t *)d_idx_sel, d_key, d_w0, d_w1, (const uint32_t *)d_rep_start,
            (const uint32_t *)d_perm, d_out_off, R, m, class_bit, (uint32_t *)d_idx, d_flags);
    else
        subset_expand_kernel<uint64_t><<<grid_for(m), 256, 0, st>>>(
            (const uint64_t *)d_pos, (const uint64_t *)d_idx_sel, d_key, d_w0, d_w1, (const uint64_t *)d_rep_start,
            (const uint64_t *)d_perm, d_out_off, R, m, class_bit, (uint64_t *)d_idx, d_flags);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

}  // namespace gk
