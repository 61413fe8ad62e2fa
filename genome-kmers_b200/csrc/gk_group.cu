// gk_group.cu -- segmented run-length pass over the sorted k-mers (north_star subsystem 3;
// replaces the linear group walk kmers.py:523-648 and the histogram kmers.py:454-520).
//
//   head flags   flag[p] = k-mer at sorted position p differs from the one at p-1
//                  * from sorted radix keys (8 B/k-mer, the fast path), or
//                  * from the sequence bytes with the reference's '$'-terminated comparator
//                    (kmers.py:306-397) for anything the keys cannot answer: ambiguous windows,
//                    a kmer_len different from the sort length, indices loaded from disk.
//   select       ordered stream compaction of flagged positions -> group offsets (one entry per
//                distinct k-mer).  Three kernels: per-tile counts, one-CTA scan of the counts,
//                ordered write.
//   histogram    counts_by_group_size[min(size, max_bin)] += 1, total += size over groups with
//                min_group <= size <= max_group (kmers.py:514-518, :612-614); small sizes are
//                privatised in shared memory because a random genome puts almost every group in
//                bin 1.
#include <stdlib.h>

#include <vector>

#include "gk_common.cuh"

namespace gk {


// ---- head flags from sorted keys ---------------------------------------------------------------
__global__ void __launch_bounds__(256)
key_flags_kernel(const uint64_t *__restrict__ keys, uint64_t n, int class_bit,
                 uint8_t *__restrict__ flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n8 = (n + 7) / 8;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n8; c += stride) {
        const uint64_t p0 = c * 8;
        uint64_t prev = (p0 == 0) ? 0 : keys[p0 - 1];
        uint64_t packed = 0;
        if (p0 + 8 <= n) {
            const ulonglong2 *v = reinterpret_cast<const ulonglong2 *>(keys + p0);
            uint64_t k[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ulonglong2 q = v[i];
                k[2 * i] = q.x;
                k[2 * i + 1] = q.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool amb = class_bit && !(k[i] & 1ull);
                const bool head = (p0 + i == 0) || (k[i] != prev);
                const uint64_t f = amb ? kFlagAmb : (head ? kFlagHead : 0);
                packed |= f << (8 * i);
                prev = k[i];
            }
            *reinterpret_cast<uint64_t *>(flags + p0) = packed;
        } else {
            for (uint64_t p = p0; p < n; ++p) {
                const uint64_t k = keys[p];
                const bool amb = class_bit && !(k & 1ull);
                const bool head = (p == 0) || (k != prev);
                flags[p] = amb ? kFlagAmb : (head ? kFlagHead : 0);
                prev = k;
            }
        }
    }
}

// ---- head flags + tie repair after a PREFIX sort ------------------------------------------------------
// gk_index.cu sorts only the top bits of the keys (enough that almost every k-mer is alone in its
// prefix bucket) and leaves the rest to this pass, which reads the keys once -- the read the flags
// pass needs anyway.  For every run of equal prefixes:
//   * one element: nothing to do;
//   * up to kTieMaxRun elements: the thread of the run's first element loads the run, ranks it
//     stably by the full key in registers and writes keys, values and flags back in place.  Other
//     threads only ever compare PREFIXES of these slots, and the prefix of a slot never changes;
//   * longer runs (giant groups: N runs, exact repeats, low complexity) are left untouched and
//     marked kFlagLong.  Their members write their own flags and report a descent (key smaller than
//     its predecessor) through *descent; only then does the host sort those runs with full LSD
//     passes.  A long run without a descent is already in final order.
constexpr int kTieMaxRun = 8;
constexpr int kTieThreads = 256;
constexpr int kTieDefaultPer = 8;   // slots per thread (GK_TIE_PER overrides)
constexpr int kTieHalo = kTieMaxRun;                   // keys staged either side of the tile

// A CTA stages the PREFIXES of its tile of keys (plus a halo) in shared memory, so that every neighbour
// look-up is a shared-memory read of 4 bytes (PreT = uint32_t whenever the prefix fits).  Nineteen slots
// in twenty are alone in their prefix bucket, so the work is split into three dense passes instead of one
// divergent one (a warp pays for every path one of its lanes takes):
//   1  every slot: compare prefixes with both neighbours; alone -> write the flag, else queue the slot
//   2  queued slots: bounded search for the run's ends; long run -> own flag (+ descent report),
//      first slot of a short run -> queue the run
//   3  queued runs: rank the run by the full key in registers, write keys, values and flags in place
template <typename ValT, typename PreT, int kTiePerThread>
// (five CTAs per SM with 32-bit start indices: 48 registers instead of 54, no spills; 7.48 -> 7.42 ms per bench
// step.  Six spill, and so do five with 64-bit starts: four there.)
__global__ void __launch_bounds__(kTieThreads, sizeof(ValT) == 8 ? 4 : 5)
tie_fix_flags_kernel(uint64_t *__restrict__ keys, ValT *__restrict__ vals, uint64_t n, int lo_bits,
                     int class_bit, uint8_t *__restrict__ flags, unsigned int *__restrict__ descent,
                     unsigned long long *__restrict__ n_amb_out /* nullable: count of ambiguous slots */,
                     unsigned long long *__restrict__ descent_list /* nullable: first kDescentCap positions */)
{
    constexpr int kTieTile = kTieThreads * kTiePerThread;  // slots per CTA
    constexpr int kRunShift = kTiePerThread > 8 ? 12 : 11;  // s_run entry: slot | (length - 1) << kRunShift
    static_assert(kTieTile <= (1 << kRunShift) && kRunShift + 3 <= 16, "run entries are 16 bits");
    // bit (i + 32) of s_cont: slot i (tile-relative, -8 <= i < kTieTile + 8) has the same prefix as slot i - 1
    constexpr int kContWords = kTieTile / 32 + 2;
    __shared__ PreT s_pre[kTieTile + 2 * kTieHalo];
    __shared__ uint32_t s_cont[kContWords + 1];
    __shared__ uint16_t s_member[kTieTile];
    __shared__ uint16_t s_run[kTieTile / 2];
    __shared__ uint32_t s_n_member, s_n_run;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint64_t tile0 = (uint64_t)blockIdx.x * kTieTile;
    const int n_in_tile = (n - tile0 < (uint64_t)kTieTile) ? (int)(n - tile0) : kTieTile;
    if (t == 0) { s_n_member = 0; s_n_run = 0; s_cont[kContWords] = 0; }
    // slot i lives at s_pre[i + kTieHalo]; slots outside [0, n) are never looked at
    auto valid = [&](int i) -> bool { return (i >= 0 || tile0 >= (uint64_t)(-i)) && tile0 + (int64_t)i < n; };
    // count bits [first, first + count) of the continuation mask, count <= 8, first >= -32
    auto cont_bits = [&](int first, int count) -> uint32_t {
        const uint32_t u = (uint32_t)(first + 32);
        const uint64_t w = ((uint64_t)s_cont[(u >> 5) + 1] << 32) | s_cont[u >> 5];
        return (uint32_t)(w >> (u & 31u)) & ((1u << count) - 1u);
    };
    uint8_t *tile_flags = flags + tile0;
    uint32_t tied_mask = 0;
    // nearly every tile lies inside the array with both halos: no bounds logic at all there
    const bool interior = tile0 >= (uint64_t)kTieHalo && tile0 + kTieTile + kTieHalo <= n;
    if (interior) {
        const uint64_t *tile_keys = keys + tile0;
        PreT my_pre[kTiePerThread];
        uint32_t my_amb = 0;
#pragma unroll
        for (int j = 0; j < kTiePerThread; ++j) {
            const uint64_t k = tile_keys[j * kTieThreads + (int)t];
            my_pre[j] = (PreT)(k >> lo_bits);
            my_amb |= (uint32_t)(~k & 1ull) << j;
            s_pre[kTieHalo + j * kTieThreads + (int)t] = my_pre[j];
        }
        if (!class_bit) my_amb = 0;
        if (n_amb_out && class_bit) {
            const uint32_t c = warp_sum((uint32_t)__popc(my_amb));
            if (lane == 0 && c) atomicAdd(n_amb_out, (unsigned long long)c);
        }
        if (t < 2 * kTieHalo) {
            const int i = t < (uint32_t)kTieHalo ? (int)t - kTieHalo : kTieTile + (int)t - kTieHalo;
            s_pre[i + kTieHalo] = (PreT)(tile_keys[i] >> lo_bits);
        }
        __syncthreads();
        // ---- 1 (interior) ---------------------------------------------------------------------------------
#pragma unroll
        for (int j = 0; j < kTiePerThread; ++j) {
            const int i = j * kTieThreads + (int)t;
            const PreT pre = my_pre[j];
            const bool ph = s_pre[i - 1 + kTieHalo] != pre;
            const bool nh = s_pre[i + 1 + kTieHalo] != pre;
            if (ph & nh) tile_flags[i] = ((my_amb >> j) & 1u) ? kFlagAmb : kFlagHead;
            else tied_mask |= 1u << j;
            const uint32_t word = __ballot_sync(0xffffffffu, !ph);
            if (lane == 0) s_cont[j * (kTieThreads / 32) + warp + 1] = word;
        }
        if (warp == 0) {  // the halo words: eight slots before the tile, eight after it
            const int ib = -32 + (int)lane, ia = kTieTile + (int)lane;
            const bool cb = ib >= -(kTieHalo - 1) && s_pre[ib - 1 + kTieHalo] == s_pre[ib + kTieHalo];
            const bool ca = lane < (uint32_t)kTieHalo && s_pre[ia - 1 + kTieHalo] == s_pre[ia + kTieHalo];
            const uint32_t wb = __ballot_sync(0xffffffffu, cb), wa = __ballot_sync(0xffffffffu, ca);
            if (lane == 0) { s_cont[0] = wb; s_cont[kContWords - 1] = wa; }
        }
    } else {
        PreT my_pre[kTiePerThread];
        uint32_t my_amb = 0;
#pragma unroll
        for (int j = 0; j < kTiePerThread; ++j) {
            const int i = j * kTieThreads + (int)t;
            const uint64_t k = (i < n_in_tile) ? keys[tile0 + i] : 0ull;
            my_pre[j] = (PreT)(k >> lo_bits);
            my_amb |= ((i < n_in_tile && class_bit && !(k & 1ull)) ? 1u : 0u) << j;
            s_pre[kTieHalo + i] = my_pre[j];
        }
        if (n_amb_out && class_bit) {
            const uint32_t c = warp_sum((uint32_t)__popc(my_amb));
            if (lane == 0 && c) atomicAdd(n_amb_out, (unsigned long long)c);
        }
        if (t < 2 * kTieHalo) {
            const bool before = t < kTieHalo;
            const uint64_t p = before ? tile0 - kTieHalo + t : tile0 + kTieTile + (t - kTieHalo);
            const bool ok = before ? (tile0 >= (uint64_t)kTieHalo - t) : (p < n);
            if (ok) s_pre[before ? t : kTieHalo + kTieTile + (t - kTieHalo)] = (PreT)(keys[p] >> lo_bits);
        }
        __syncthreads();
        // ---- 1 (array ends) -------------------------------------------------------------------------------
        const int first_i = (tile0 == 0) ? 0 : -1;                                  // slot without predecessor
        const int last_i = (n - tile0 <= (uint64_t)kTieTile) ? n_in_tile - 1 : -1;  // slot without successor
#pragma unroll
        for (int j = 0; j < kTiePerThread; ++j) {
            const int i = j * kTieThreads + (int)t;
            bool cont = false;
            if (i < n_in_tile) {
                const PreT pre = my_pre[j];
                const bool ph = (i == first_i) | (s_pre[i - 1 + kTieHalo] != pre);
                const bool nh = (i == last_i) | (s_pre[i + 1 + kTieHalo] != pre);
                cont = !ph;
                if (ph & nh) tile_flags[i] = ((my_amb >> j) & 1u) ? kFlagAmb : kFlagHead;
                else tied_mask |= 1u << j;
            }
            const uint32_t word = __ballot_sync(0xffffffffu, cont);
            if (lane == 0) s_cont[j * (kTieThreads / 32) + warp + 1] = word;
        }
        if (warp == 0) {
            const int ib = -32 + (int)lane, ia = kTieTile + (int)lane;
            const bool cb = ib >= -(kTieHalo - 1) && valid(ib - 1) && valid(ib) &&
                            s_pre[ib - 1 + kTieHalo] == s_pre[ib + kTieHalo];
            const bool ca = lane < (uint32_t)kTieHalo && valid(ia) &&
                            s_pre[ia - 1 + kTieHalo] == s_pre[ia + kTieHalo];
            const uint32_t wb = __ballot_sync(0xffffffffu, cb), wa = __ballot_sync(0xffffffffu, ca);
            if (lane == 0) { s_cont[0] = wb; s_cont[kContWords - 1] = wa; }
        }
    }
    {   // one queue reservation per warp for all eight rounds
        const uint32_t cnt = __popc(tied_mask);
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += v;
        }
        uint32_t base = 0;
        if (lane == 31 && inc) base = atomicAdd(&s_n_member, inc);
        base = __shfl_sync(0xffffffffu, base, 31) + inc - cnt;
        while (tied_mask) {
            const int j = __ffs(tied_mask) - 1;
            tied_mask &= tied_mask - 1;
            s_member[base++] = (uint16_t)(j * kTieThreads + (int)t);
        }
    }
    __syncthreads();

    // ---- 2 ---------------------------------------------------------------------------------------------
    // run geometry from the continuation bits: d slots back to the run's first slot, f slots on to its last
    const uint32_t n_member = s_n_member;
    for (uint32_t r = t; r < n_member; r += kTieThreads) {
        const int i = s_member[r];
        const bool ph = !cont_bits(i, 1);
        bool is_long = false;
        int d = 0;
        if (!ph) {
            const uint32_t back = cont_bits(i - (kTieMaxRun - 1), kTieMaxRun - 1);  // slots i-7 .. i-1
            const int ones = __clz((int)~(back << (32 - (kTieMaxRun - 1))));      // leading ones from slot i-1
            if (ones >= kTieMaxRun - 1) is_long = true; else d = ones + 1;
        }
        const uint32_t fwd = cont_bits(i + 1, kTieMaxRun);                         // slots i+1 .. i+8
        const int f = __ffs((int)~fwd) - 1;                                        // trailing ones
        if (d + f + 1 > kTieMaxRun) is_long = true;
        if (is_long) {  // nobody rewrites the keys of a long run: read them where they are
            const uint64_t p = tile0 + (uint64_t)i;
            const uint64_t k = keys[p];
            const uint64_t kp = ph ? 0 : keys[p - 1];
            const bool head = ph || kp != k;
            const bool amb = class_bit && !(k & 1ull);
            if (!ph && k < kp) {   // *descent counts them; the first kDescentCap positions are listed
                const unsigned int s_ = atomicAdd(descent, 1u);
                if (descent_list && s_ < (unsigned int)kDescentCap) descent_list[s_] = p;
            }
            flags[p] = (amb ? kFlagAmb : (head ? kFlagHead : 0)) | kFlagLong;
        } else if (ph) {  // a short run that starts in this tile: slot and length - 1
            s_run[atomicAdd(&s_n_run, 1u)] = (uint16_t)(i | (f << kRunShift));
        }
    }
    __syncthreads();

    // ---- 3 ---------------------------------------------------------------------------------------------
    const uint32_t n_runs = s_n_run;
    for (uint32_t r = t; r < n_runs; r += kTieThreads) {
        const int h = s_run[r] & ((1 << kRunShift) - 1);
        const int len = (s_run[r] >> kRunShift) + 1;
        const uint64_t g = tile0 + (uint64_t)h;
        if (len == 2) {  // nearly every run: one compare, at most one swap
            const uint64_t k0 = keys[g], k1 = keys[g + 1];
            const bool a0 = class_bit && !(k0 & 1ull), a1 = class_bit && !(k1 & 1ull);
            if (k1 < k0) {
                const ValT v0 = vals[g], v1 = vals[g + 1];
                keys[g] = k1; keys[g + 1] = k0;
                vals[g] = v1; vals[g + 1] = v0;
                flags[g] = a1 ? kFlagAmb : kFlagHead;
                flags[g + 1] = a0 ? kFlagAmb : kFlagHead;
            } else {
                flags[g] = a0 ? kFlagAmb : kFlagHead;
                flags[g + 1] = a1 ? kFlagAmb : (k1 != k0 ? kFlagHead : 0);
            }
            continue;
        }
        uint64_t kk[kTieMaxRun];
#pragma unroll
        for (int i = 0; i < kTieMaxRun; ++i) kk[i] = (i < len) ? keys[g + i] : ~0ull;
        bool sorted = true;
#pragma unroll
        for (int i = 1; i < kTieMaxRun; ++i) sorted = sorted && (i >= len || kk[i - 1] <= kk[i]);
        if (sorted) {  // only the flags are missing
#pragma unroll
            for (int i = 0; i < kTieMaxRun; ++i) {
                if (i < len) {
                    const bool a = class_bit && !(kk[i] & 1ull);
                    const bool first = (i == 0) || kk[i - 1] != kk[i];
                    flags[g + i] = a ? kFlagAmb : (first ? kFlagHead : 0);
                }
            }
            continue;
        }
        ValT vv[kTieMaxRun];
#pragma unroll
        for (int i = 0; i < kTieMaxRun; ++i) vv[i] = (i < len) ? vals[g + i] : (ValT)0;
#pragma unroll
        for (int i = 0; i < kTieMaxRun; ++i) {
            int rank = 0;
            bool first = true;
#pragma unroll
            for (int j = 0; j < kTieMaxRun; ++j) {
                if (j < i) {
                    rank += (kk[j] <= kk[i]) ? 1 : 0;
                    first = first && (kk[j] != kk[i]);
                } else if (j > i) {
                    rank += (kk[j] < kk[i]) ? 1 : 0;
                }
            }
            if (i < len) {
                const bool a = class_bit && !(kk[i] & 1ull);
                flags[g + rank] = a ? kFlagAmb : (first ? kFlagHead : 0);
                keys[g + rank] = kk[i];  // in place is safe: every source value is in registers
                vals[g + rank] = vv[i];
            }
        }
    }
}

// ---- flags of re-sorted prefix buckets (the bucket-wise repair lives in gk_sort.cu / gk_index.cu) ------------
// head / ambiguous flags of the fully sorted slots [lo, hi) of each range (a bucket start is always a head)
__global__ void __launch_bounds__(256)
range_key_flags_kernel(const uint64_t *__restrict__ keys, const unsigned long long *__restrict__ ranges,
                       int class_bit, uint8_t *__restrict__ flags)
{
    const uint64_t lo = ranges[2 * blockIdx.y], hi = ranges[2 * blockIdx.y + 1];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < hi; p += stride) {
        const uint64_t k = keys[p];
        const bool amb = class_bit && !(k & 1ull);
        const bool head = (p == lo) || keys[p - 1] != k;
        flags[p] = amb ? kFlagAmb : (head ? kFlagHead : 0);
    }
}

int range_key_flags_device(const uint64_t *d_keys, const unsigned long long *d_ranges, uint32_t n_ranges,
                           uint64_t longest, int class_bit, uint8_t *d_flags, cudaStream_t st)
{
    if (n_ranges == 0) return GK_OK;
    uint64_t bx = (longest + 255) / 256;
    if (bx > 1024) bx = 1024;
    if (bx < 1) bx = 1;
    range_key_flags_kernel<<<dim3((unsigned)bx, n_ranges), 256, 0, st>>>(d_keys, d_ranges, class_bit, d_flags);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

// ---- head flags from the sequence bytes (reference comparator) ------------------------------------
// equal <=> compare_sba_kmers_lexicographically(...) == 0: same bytes up to kmer_len, or both
// terminated ('$' / end of array) at the same offset.  kmer_len == 0 means None.
__device__ __forceinline__ bool windows_equal(const uint8_t *__restrict__ sba, uint64_t len,
                                              uint64_t a, uint64_t b, uint32_t kmer_len)
{
    if (a == b) return true;
    for (uint32_t j = 0;; ++j) {
        const uint32_t ca = (a + j < len) ? sba[a + j] : kSep;
        const uint32_t cb = (b + j < len) ? sba[b + j] : kSep;
        const bool a_out = ca == kSep, b_out = cb == kSep;
        if (a_out || b_out) return a_out && b_out;
        if (ca != cb) return false;
        if (kmer_len && j == kmer_len - 1) return true;
    }
}

// dst == nullptr: flags[r] = head bit.  dst != nullptr: flags[dst[r]] = extra | head bit.
template <typename IdxT>
__global__ void __launch_bounds__(256)
sba_flags_kernel(const uint8_t *__restrict__ sba, uint64_t len, const IdxT *__restrict__ idx,
                 uint64_t n, uint32_t kmer_len, const IdxT *__restrict__ dst, uint8_t extra,
                 uint8_t *__restrict__ flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        bool head = true;
        if (r > 0) head = !windows_equal(sba, len, (uint64_t)idx[r - 1], (uint64_t)idx[r], kmer_len);
        const uint8_t f = extra | (head ? kFlagHead : 0);
        if (dst) flags[(uint64_t)dst[r]] = f; else flags[r] = f;
    }
}

// ---- ordered select of flagged positions ----------------------------------------------------------
constexpr int kSelThreads = 256;
constexpr int kSelPerThread = 16;
constexpr int kSelTile = kSelThreads * kSelPerThread;

__device__ __forceinline__ uint32_t load_flag_bits(const uint8_t *__restrict__ flags, uint64_t n,
                                                   uint64_t p0, uint8_t mask)
{
    // bit i of the result = (flags[p0+i] & mask) != 0, for i < 16
    uint32_t bits = 0;
    if (p0 + 16 <= n) {
        const uint4 q = *reinterpret_cast<const uint4 *>(flags + p0);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 16; ++i)
            bits |= (((w[i >> 2] >> (8 * (i & 3))) & mask) ? 1u : 0u) << i;
    } else {
        for (int i = 0; i < 16; ++i)
            if (p0 + i < n && (flags[p0 + i] & mask)) bits |= 1u << i;
    }
    return bits;
}

__global__ void __launch_bounds__(kSelThreads)
select_count_kernel(const uint8_t *__restrict__ flags, uint64_t n, uint8_t mask,
                    uint32_t *__restrict__ tile_counts)
{
    __shared__ uint32_t s_warp[kSelThreads / 32];
    const uint64_t p0 = (uint64_t)blockIdx.x * kSelTile + (uint64_t)threadIdx.x * kSelPerThread;
    uint32_t c = (p0 < n) ? __popc(load_flag_bits(flags, n, p0, mask)) : 0;
    c = warp_sum(c);
    if (lane_id() == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < kSelThreads / 32; ++w) s += s_warp[w];
        tile_counts[blockIdx.x] = s;
    }
}

// one CTA: exclusive scan of tile counts into 64-bit offsets; total appended at [n_tiles]
__global__ void __launch_bounds__(1024)
select_scan_kernel(const uint32_t *__restrict__ tile_counts, uint64_t n_tiles,
                   unsigned long long *__restrict__ tile_offsets)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint64_t base = 0; base < n_tiles; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long c = (i < n_tiles) ? tile_counts[i] : 0;
        unsigned long long inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += v;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        unsigned long long pre = s_carry;
        for (uint32_t w = 0; w < warp; ++w) pre += s_warp[w];
        if (i < n_tiles) tile_offsets[i] = pre + inc - c;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = pre + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_offsets[n_tiles] = s_carry;
}

template <typename T>
__global__ void __launch_bounds__(kSelThreads)
select_write_kernel(const uint8_t *__restrict__ flags, uint64_t n, uint8_t mask,
                    const unsigned long long *__restrict__ tile_offsets, T *__restrict__ pos_out,
                    const T *__restrict__ pay_in, T *__restrict__ pay_out,
                    const T *__restrict__ pay2_in, T *__restrict__ pay2_out,
                    const uint8_t *__restrict__ pay8_in, uint8_t *__restrict__ pay8_out, uint8_t pay8_mask)
{
    __shared__ uint32_t s_warp[kSelThreads / 32];
    {   // a tile that keeps all of its elements (dropping a handful of windows from a whole index): a shifted,
        // fully coalesced copy
        const uint64_t tile0 = (uint64_t)blockIdx.x * kSelTile;
        const uint64_t t_off = tile_offsets[blockIdx.x];
        const uint64_t t_cnt = tile_offsets[blockIdx.x + 1] - t_off;
        const uint64_t t_len = (n - tile0 < (uint64_t)kSelTile) ? n - tile0 : (uint64_t)kSelTile;
        if (t_cnt == t_len) {
            for (uint32_t i = threadIdx.x; i < (uint32_t)t_len; i += kSelThreads) {
                if (pos_out) pos_out[t_off + i] = (T)(tile0 + i);
                if (pay_out) pay_out[t_off + i] = pay_in[tile0 + i];
                if (pay2_out) pay2_out[t_off + i] = pay2_in[tile0 + i];
                if (pay8_out) pay8_out[t_off + i] = pay8_in[tile0 + i] & pay8_mask;
            }
            return;
        }
    }
    const uint64_t p0 = (uint64_t)blockIdx.x * kSelTile + (uint64_t)threadIdx.x * kSelPerThread;
    const uint32_t bits = (p0 < n) ? load_flag_bits(flags, n, p0, mask) : 0;
    const uint32_t c = __popc(bits);
    uint32_t inc = c;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += v;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t pre = 0;
    for (uint32_t w = 0; w < warp; ++w) pre += s_warp[w];
    uint64_t out = tile_offsets[blockIdx.x] + pre + inc - c;
    uint32_t b = bits;
    while (b) {
        const uint32_t i = __ffs(b) - 1;
        b &= b - 1;
        if (pos_out) pos_out[out] = (T)(p0 + i);
        if (pay_out) pay_out[out] = pay_in[p0 + i];
        if (pay2_out) pay2_out[out] = pay2_in[p0 + i];
        if (pay8_out) pay8_out[out] = pay8_in[p0 + i] & pay8_mask;
        ++out;
    }
}

// Ordered selection of the positions p with (flags[p] & mask) != 0: optionally the positions
// themselves, up to two payload arrays and one byte array (masked with pay8_mask) gathered at those positions.
template <typename T>
int select_flagged_device(const uint8_t *d_flags, uint64_t n, uint8_t mask, T *d_pos_out,
                          const T *d_pay_in, T *d_pay_out, const T *d_pay2_in, T *d_pay2_out,
                          uint64_t *h_count, cudaStream_t st, const uint8_t *d_pay8_in = nullptr,
                          uint8_t *d_pay8_out = nullptr, uint8_t pay8_mask = 0xff)
{
    if (h_count) *h_count = 0;
    if (n == 0) return GK_OK;
    const uint64_t tiles = (n + kSelTile - 1) / kSelTile;
    DeviceBuffer temp;
    const size_t counts_bytes = ((size_t)tiles * 4 + 15) & ~(size_t)15;
    GK_TRY(temp.alloc(counts_bytes + (size_t)(tiles + 1) * 8, st));
    uint32_t *d_counts = temp.as<uint32_t>();
    unsigned long long *d_offsets =
        reinterpret_cast<unsigned long long *>(temp.as<unsigned char>() + counts_bytes);
    select_count_kernel<<<(unsigned)tiles, kSelThreads, 0, st>>>(d_flags, n, mask, d_counts);
    GK_LAUNCH_CHECK();
    select_scan_kernel<<<1, 1024, 0, st>>>(d_counts, tiles, d_offsets);
    GK_LAUNCH_CHECK();
    select_write_kernel<T><<<(unsigned)tiles, kSelThreads, 0, st>>>(
        d_flags, n, mask, d_offsets, d_pos_out, d_pay_in, d_pay_out, d_pay2_in, d_pay2_out, d_pay8_in, d_pay8_out,
        pay8_mask);
    GK_LAUNCH_CHECK();
    if (h_count) {
        GK_CUDA(cudaMemcpyAsync(h_count, d_offsets + tiles, 8, cudaMemcpyDeviceToHost, st));
        GK_CUDA(cudaStreamSynchronize(st));
    }
    return GK_OK;
}

template int select_flagged_device<uint32_t>(const uint8_t *, uint64_t, uint8_t, uint32_t *, const uint32_t *,
                                             uint32_t *, const uint32_t *, uint32_t *, uint64_t *,
                                             cudaStream_t, const uint8_t *, uint8_t *, uint8_t);
template int select_flagged_device<uint64_t>(const uint8_t *, uint64_t, uint8_t, uint64_t *, const uint64_t *,
                                             uint64_t *, const uint64_t *, uint64_t *, uint64_t *,
                                             cudaStream_t, const uint8_t *, uint8_t *, uint8_t);

// ordered selection of (key, value) pairs at flagged positions, with the positions themselves
template <typename T>
__global__ void __launch_bounds__(kSelThreads)
select_pairs_write_kernel(const uint8_t *__restrict__ flags, uint64_t n, uint8_t mask,
                          const unsigned long long *__restrict__ tile_offsets, T *__restrict__ pos_out,
                          const uint64_t *__restrict__ keys_in, uint64_t *__restrict__ keys_out,
                          const T *__restrict__ vals_in, T *__restrict__ vals_out)
{
    __shared__ uint32_t s_warp[kSelThreads / 32];
    const uint64_t p0 = (uint64_t)blockIdx.x * kSelTile + (uint64_t)threadIdx.x * kSelPerThread;
    const uint32_t bits = (p0 < n) ? load_flag_bits(flags, n, p0, mask) : 0;
    const uint32_t c = __popc(bits);
    uint32_t inc = c;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += v;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t pre = 0;
    for (uint32_t w = 0; w < warp; ++w) pre += s_warp[w];
    uint64_t out = tile_offsets[blockIdx.x] + pre + inc - c;
    uint32_t b = bits;
    while (b) {
        const uint32_t i = __ffs(b) - 1;
        b &= b - 1;
        pos_out[out] = (T)(p0 + i);
        keys_out[out] = keys_in[p0 + i];
        vals_out[out] = vals_in[p0 + i];
        ++out;
    }
}

// Two-step ordered selection of (key, value) pairs: select_pairs_count leaves the tile offsets in
// `temp` and returns the number of flagged positions (synchronises), so that the caller can size the
// outputs; select_pairs_write then fills them.
int select_pairs_count(const uint8_t *d_flags, uint64_t n, uint8_t mask, DeviceBuffer &temp, uint64_t *h_count,
                       cudaStream_t st)
{
    *h_count = 0;
    if (n == 0) return GK_OK;
    const uint64_t tiles = (n + kSelTile - 1) / kSelTile;
    const size_t counts_bytes = ((size_t)tiles * 4 + 15) & ~(size_t)15;
    GK_TRY(temp.alloc(counts_bytes + (size_t)(tiles + 1) * 8, st));
    uint32_t *d_counts = temp.as<uint32_t>();
    unsigned long long *d_offsets =
        reinterpret_cast<unsigned long long *>(temp.as<unsigned char>() + counts_bytes);
    select_count_kernel<<<(unsigned)tiles, kSelThreads, 0, st>>>(d_flags, n, mask, d_counts);
    GK_LAUNCH_CHECK();
    select_scan_kernel<<<1, 1024, 0, st>>>(d_counts, tiles, d_offsets);
    GK_LAUNCH_CHECK();
    GK_CUDA(cudaMemcpyAsync(h_count, d_offsets + tiles, 8, cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    return GK_OK;
}

int select_pairs_write(const uint8_t *d_flags, uint64_t n, uint8_t mask, const DeviceBuffer &temp, int val_bytes,
                       void *d_pos_out, const uint64_t *d_keys_in, uint64_t *d_keys_out, const void *d_vals_in,
                       void *d_vals_out, cudaStream_t st)
{
    if (n == 0) return GK_OK;
    const uint64_t tiles = (n + kSelTile - 1) / kSelTile;
    const size_t counts_bytes = ((size_t)tiles * 4 + 15) & ~(size_t)15;
    const unsigned long long *d_offsets =
        reinterpret_cast<const unsigned long long *>(temp.as<unsigned char>() + counts_bytes);
    if (val_bytes == 4)
        select_pairs_write_kernel<uint32_t><<<(unsigned)tiles, kSelThreads, 0, st>>>(
            d_flags, n, mask, d_offsets, (uint32_t *)d_pos_out, d_keys_in, d_keys_out,
            (const uint32_t *)d_vals_in, (uint32_t *)d_vals_out);
    else
        select_pairs_write_kernel<uint64_t><<<(unsigned)tiles, kSelThreads, 0, st>>>(
            d_flags, n, mask, d_offsets, (uint64_t *)d_pos_out, d_keys_in, d_keys_out,
            (const uint64_t *)d_vals_in, (uint64_t *)d_vals_out);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

// type-erased front end (elem_bytes 4 or 8)
int select_flagged(const uint8_t *d_flags, uint64_t n, uint8_t mask, int elem_bytes, void *d_pos_out,
                   const void *d_pay_in, void *d_pay_out, const void *d_pay2_in, void *d_pay2_out,
                   uint64_t *h_count, cudaStream_t st)
{
    if (elem_bytes == 4)
        return select_flagged_device<uint32_t>(d_flags, n, mask, (uint32_t *)d_pos_out,
                                               (const uint32_t *)d_pay_in, (uint32_t *)d_pay_out,
                                               (const uint32_t *)d_pay2_in, (uint32_t *)d_pay2_out, h_count, st);
    return select_flagged_device<uint64_t>(d_flags, n, mask, (uint64_t *)d_pos_out,
                                           (const uint64_t *)d_pay_in, (uint64_t *)d_pay_out,
                                           (const uint64_t *)d_pay2_in, (uint64_t *)d_pay2_out, h_count, st);
}

// ordered selection of one payload array plus the (masked) bytes of a byte array, e.g. start indices and
// their head flags; no position list
int select_flagged_with_bytes(const uint8_t *d_flags, uint64_t n, uint8_t mask, int elem_bytes,
                              const void *d_pay_in, void *d_pay_out, const uint8_t *d_bytes_in,
                              uint8_t *d_bytes_out, uint8_t bytes_mask, uint64_t *h_count, cudaStream_t st)
{
    if (elem_bytes == 4)
        return select_flagged_device<uint32_t>(d_flags, n, mask, nullptr, (const uint32_t *)d_pay_in,
                                               (uint32_t *)d_pay_out, nullptr, nullptr, h_count, st, d_bytes_in,
                                               d_bytes_out, bytes_mask);
    return select_flagged_device<uint64_t>(d_flags, n, mask, nullptr, (const uint64_t *)d_pay_in,
                                           (uint64_t *)d_pay_out, nullptr, nullptr, h_count, st, d_bytes_in,
                                           d_bytes_out, bytes_mask);
}

// ---- group-size histogram -----------------------------------------------------------------------------
constexpr int kHistSmallBins = 2048;

template <typename PosT>
__global__ void __launch_bounds__(256)
group_hist_kernel(const PosT *__restrict__ offsets, uint64_t n_groups, uint64_t n,
                  const uint8_t *__restrict__ flags, uint8_t skip_mask, uint64_t min_group,
                  uint64_t max_group, uint64_t max_bin, unsigned long long *__restrict__ hist,
                  unsigned long long *__restrict__ totals)
{
    __shared__ uint32_t s_small[kHistSmallBins];
    __shared__ unsigned long long s_total[3];
    for (int i = threadIdx.x; i < kHistSmallBins; i += blockDim.x) s_small[i] = 0;
    if (threadIdx.x < 3) s_total[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long total = 0, counted = 0, top_bin = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        const uint64_t a = (uint64_t)offsets[g];
        const uint64_t b = (g + 1 < n_groups) ? (uint64_t)offsets[g + 1] : n;
        const uint64_t size = b - a;
        if (skip_mask && (flags[a] & skip_mask)) continue;  // e.g. groups of ambiguous k-mers
        if (size >= min_group && (max_group == 0 || size <= max_group)) {
            total += size;
            ++counted;
            const uint64_t bin = size < max_bin ? size : max_bin;
            if (bin > top_bin) top_bin = bin;
            if (bin < kHistSmallBins) atomicAdd(&s_small[bin], 1u);
            else atomicAdd(&hist[bin], 1ull);
        }
    }
    total = warp_sum(total);
    counted = warp_sum(counted);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, top_bin, o);
        top_bin = other > top_bin ? other : top_bin;
    }
    if (lane_id() == 0) {
        atomicAdd(&s_total[0], total);
        atomicAdd(&s_total[1], counted);
        atomicMax(&s_total[2], top_bin);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kHistSmallBins; i += blockDim.x) {
        const uint32_t c = s_small[i];
        if (c && (uint64_t)i <= max_bin) atomicAdd(&hist[i], (unsigned long long)c);
    }
    if (threadIdx.x < 2 && s_total[threadIdx.x]) atomicAdd(&totals[threadIdx.x], s_total[threadIdx.x]);
    if (threadIdx.x == 2 && s_total[2]) atomicMax(&totals[2], s_total[2]);
}

static int launch_grid(uint64_t items, int block)
{
    uint64_t blocks = (items + block - 1) / block;
    const uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int key_flags_device(const uint64_t *d_keys, uint64_t n, int class_bit, uint8_t *d_flags,
                     cudaStream_t st)
{
    if (n == 0) return GK_OK;
    key_flags_kernel<<<launch_grid((n + 7) / 8, 256), 256, 0, st>>>(d_keys, n, class_bit, d_flags);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

// head/ambiguous flags of keys sorted on bits [lo_bits, 64) only, repairing short prefix runs in place;
// *d_descent (zeroed by the caller) becomes 1 when a kFlagLong run is out of order
template <int PER>
static int tie_fix_launch(uint64_t *d_keys, void *d_vals, int val_bytes, uint64_t n, int lo_bits, int class_bit,
                          uint8_t *d_flags, unsigned int *d_descent, unsigned long long *d_n_amb,
                          unsigned long long *d_descent_list, cudaStream_t st)
{
    constexpr int kTile = kTieThreads * PER;
    const unsigned grid = (unsigned)((n + kTile - 1) / kTile);
    const bool narrow = (64 - lo_bits) <= 32;  // the prefix fits 32 bits
    if (val_bytes == 4 && narrow)
        tie_fix_flags_kernel<uint32_t, uint32_t, PER><<<grid, kTieThreads, 0, st>>>(
            d_keys, (uint32_t *)d_vals, n, lo_bits, class_bit, d_flags, d_descent, d_n_amb, d_descent_list);
    else if (val_bytes == 4)
        tie_fix_flags_kernel<uint32_t, uint64_t, PER><<<grid, kTieThreads, 0, st>>>(
            d_keys, (uint32_t *)d_vals, n, lo_bits, class_bit, d_flags, d_descent, d_n_amb, d_descent_list);
    else if (narrow)
        tie_fix_flags_kernel<uint64_t, uint32_t, PER><<<grid, kTieThreads, 0, st>>>(
            d_keys, (uint64_t *)d_vals, n, lo_bits, class_bit, d_flags, d_descent, d_n_amb, d_descent_list);
    else
        tie_fix_flags_kernel<uint64_t, uint64_t, PER><<<grid, kTieThreads, 0, st>>>(
            d_keys, (uint64_t *)d_vals, n, lo_bits, class_bit, d_flags, d_descent, d_n_amb, d_descent_list);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

// head/ambiguous flags of keys sorted on bits [lo_bits, 64) only, repairing short prefix runs in place;
// *d_descent (zeroed by the caller) counts the slots of kFlagLong runs that are out of order.
// GK_TIE_PER=8|16 selects the slots per thread (tuning runs).
int tie_fix_flags_device(uint64_t *d_keys, void *d_vals, int val_bytes, uint64_t n, int lo_bits,
                         int class_bit, uint8_t *d_flags, unsigned int *d_descent,
                         unsigned long long *d_n_amb, cudaStream_t st, unsigned long long *d_descent_list)
{
    if (n == 0) return GK_OK;
    const char *e = getenv("GK_TIE_PER");
    const int per = (e && *e) ? atoi(e) : kTieDefaultPer;
    if (per == 16)
        return tie_fix_launch<16>(d_keys, d_vals, val_bytes, n, lo_bits, class_bit, d_flags, d_descent, d_n_amb,
                                  d_descent_list, st);
    return tie_fix_launch<8>(d_keys, d_vals, val_bytes, n, lo_bits, class_bit, d_flags, d_descent, d_n_amb,
                             d_descent_list, st);
}

int sba_flags_device(const uint8_t *d_sba, uint64_t sba_len, const void *d_idx, int idx_bytes,
                     uint64_t n, uint32_t kmer_len, const void *d_dst, uint8_t extra,
                     uint8_t *d_flags, cudaStream_t st)
{
    if (n == 0) return GK_OK;
    const int grid = launch_grid(n, 256);
    if (idx_bytes == 4)
        sba_flags_kernel<uint32_t><<<grid, 256, 0, st>>>(d_sba, sba_len, (const uint32_t *)d_idx, n,
                                                         kmer_len, (const uint32_t *)d_dst, extra,
                                                         d_flags);
    else
        sba_flags_kernel<uint64_t><<<grid, 256, 0, st>>>(d_sba, sba_len, (const uint64_t *)d_idx, n,
                                                         kmer_len, (const uint64_t *)d_dst, extra,
                                                         d_flags);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

// The device histogram is a dense table of max_bin + 1 counters (8 MB at the reference's default
// max_counts_bin), but only a handful of bins are ever occupied, and one giant group (an N run) puts
// a count into the LAST bin.  So the table never crosses PCIe: a compaction kernel lists the non-empty
// bins as (bin, count) pairs, and the host scatters them into the caller's zeroed table.
static thread_local std::vector<unsigned long long> g_last_pairs;  // (bin, count) of the last call
static thread_local uint64_t g_last_top_bin = 0;
uint64_t last_hist_top_bin() { return g_last_top_bin; }
const std::vector<unsigned long long> &last_hist_pairs() { return g_last_pairs; }
void set_last_hist_pairs(const std::vector<unsigned long long> &pairs, uint64_t top_bin)
{
    g_last_pairs = pairs;
    g_last_top_bin = top_bin;
}
void set_last_hist_single(uint64_t bin, uint64_t count)
{
    g_last_pairs.clear();
    if (count) { g_last_pairs.push_back(bin); g_last_pairs.push_back(count); }
    g_last_top_bin = count ? bin : 0;
}

constexpr uint64_t kPairsInline = 250;    // pairs that come back with the totals in one copy
constexpr uint64_t kPairsCapacity = 65536;

__global__ void __launch_bounds__(256)
hist_compact_kernel(const unsigned long long *__restrict__ hist, uint64_t n_bins,
                    unsigned long long *__restrict__ n_pairs, unsigned long long *__restrict__ pairs,
                    uint64_t capacity)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_bins; i += stride) {
        const unsigned long long c = hist[i];
        if (c) {
            const unsigned long long slot = atomicAdd(n_pairs, 1ull);
            if (slot < capacity) {
                pairs[2 * slot] = i;
                pairs[2 * slot + 1] = c;
            }
        }
    }
}

// device layout used by both histogram drivers: [hist (max_bin+1)][totals x3][n_pairs][pairs 2*capacity]
static size_t hist_buffer_bytes(uint64_t max_bin) { return (size_t)(max_bin + 1) * 8 + 32 + kPairsCapacity * 16; }

// After the histogram kernel: compact, fetch totals + pairs, scatter into h_hist (already zero, may be NULL).
static int collect_hist(unsigned long long *d_hist, uint64_t max_bin, int64_t *h_hist, int64_t *h_total,
                        int64_t *h_counted, cudaStream_t st)
{
    unsigned long long *d_totals = d_hist + (max_bin + 1);
    unsigned long long *d_n_pairs = d_totals + 3;
    unsigned long long *d_pairs = d_n_pairs + 1;
    const uint64_t n_bins = max_bin + 1;
    uint64_t blocks = (n_bins + 255) / 256;
    const uint64_t cap_blocks = (uint64_t)sm_count() * 8;
    if (blocks > cap_blocks) blocks = cap_blocks;
    hist_compact_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_hist, n_bins, d_n_pairs, d_pairs, kPairsCapacity);
    GK_LAUNCH_CHECK();
    unsigned long long head[4 + 2 * kPairsInline];
    GK_CUDA(cudaMemcpyAsync(head, d_totals, sizeof(head), cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    const uint64_t n_pairs = head[3];
    g_last_pairs.assign(head + 4, head + 4 + 2 * (n_pairs < kPairsInline ? n_pairs : kPairsInline));
    if (n_pairs > kPairsInline) {
        DeviceBuffer big;
        const unsigned long long *src = d_pairs;
        if (n_pairs > kPairsCapacity) {  // more distinct group sizes than the list holds: redo with room
            GK_TRY(big.alloc((size_t)(n_pairs + 1) * 16, st));
            GK_CUDA(cudaMemsetAsync(big.ptr, 0, 8, st));
            hist_compact_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_hist, n_bins, big.as<unsigned long long>(),
                                                                   big.as<unsigned long long>() + 2, n_pairs);
            GK_LAUNCH_CHECK();
            src = big.as<unsigned long long>() + 2;
        }
        g_last_pairs.resize((size_t)2 * n_pairs);
        GK_CUDA(cudaMemcpyAsync(g_last_pairs.data(), src, (size_t)n_pairs * 16, cudaMemcpyDeviceToHost, st));
        GK_CUDA(cudaStreamSynchronize(st));
    }
    g_last_top_bin = head[2];
    if (h_hist)
        for (size_t i = 0; i + 1 < g_last_pairs.size(); i += 2) h_hist[g_last_pairs[i]] = (int64_t)g_last_pairs[i + 1];
    if (h_total) *h_total = (int64_t)head[0];
    if (h_counted) *h_counted = (int64_t)head[1];
    return GK_OK;
}

// hist (host, max_bin+1 int64, ALREADY ZERO, may be NULL) and totals[0] = sum of sizes, totals[1] = groups counted.
// Groups whose first slot has (flags & skip_mask) != 0 are left out (skip_mask 0: none).
static int group_hist_impl(const void *d_offsets, int pos_bytes, uint64_t n_groups, uint64_t n,
                           const uint8_t *d_flags, uint8_t skip_mask, uint64_t min_group,
                           uint64_t max_group, uint64_t max_bin, int64_t *h_hist, int64_t *h_total,
                           int64_t *h_counted, cudaStream_t st)
{
    if (min_group < 1) {
        set_error("min_group_size (%llu) must be >= 1", (unsigned long long)min_group);
        return GK_ERR_ARG;
    }
    if (max_group != 0 && max_group < min_group) {
        set_error("max_group_size must be >= min_group_size");
        return GK_ERR_ARG;
    }
    if (max_bin < 1) {
        set_error("max_counts_bin (%llu) must be >= 1", (unsigned long long)max_bin);
        return GK_ERR_ARG;
    }
    DeviceBuffer buf;
    const size_t hist_bytes = (size_t)(max_bin + 1) * 8;
    GK_TRY(buf.alloc(hist_buffer_bytes(max_bin), st));
    GK_CUDA(cudaMemsetAsync(buf.ptr, 0, hist_bytes + 32, st));
    unsigned long long *d_hist = buf.as<unsigned long long>();
    unsigned long long *d_totals = d_hist + (max_bin + 1);
    if (n_groups) {
        const int grid = launch_grid(n_groups, 256);
        if (pos_bytes == 4)
            group_hist_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t *)d_offsets, n_groups, n,
                                                              d_flags, skip_mask, min_group,
                                                              max_group, max_bin, d_hist, d_totals);
        else
            group_hist_kernel<uint64_t><<<grid, 256, 0, st>>>((const uint64_t *)d_offsets, n_groups, n,
                                                              d_flags, skip_mask, min_group,
                                                              max_group, max_bin, d_hist, d_totals);
        GK_LAUNCH_CHECK();
    }
    return collect_hist(d_hist, max_bin, h_hist, h_total, h_counted, st);
}

int group_hist_device(const void *d_offsets, int pos_bytes, uint64_t n_groups, uint64_t n,
                      uint64_t min_group, uint64_t max_group, uint64_t max_bin, int64_t *h_hist,
                      int64_t *h_total, int64_t *h_counted, cudaStream_t st)
{
    return group_hist_impl(d_offsets, pos_bytes, n_groups, n, nullptr, 0, min_group, max_group, max_bin,
                           h_hist, h_total, h_counted, st);
}

// ---- group-size histogram straight from the head flags (no offsets array) ---------------------------
// The cached-flags fast path of gk_index_group_counts.  A group is closed when the NEXT head is met:
// its size is the distance between the two heads, so every thread only needs the position of the
// latest head before its 16 flags.  Three launches: last head per tile, one-CTA running max over
// the tiles, then the histogram pass.  Reads the flags twice (2 B per k-mer) and writes nothing per
// k-mer, instead of materialising 4-8 B per distinct k-mer and reading them back.
constexpr int kGfThreads = 256;
constexpr int kGfPerThread = 64;
constexpr int kGfTile = kGfThreads * kGfPerThread;

// bit i of the result = (flags[p0+i] & (1 << bit)) != 0 for i < 64; positions >= n read as 0
__device__ __forceinline__ uint64_t load_flag_bits64(const uint8_t *__restrict__ flags, uint64_t n, uint64_t p0,
                                                     int bit)
{
    uint64_t out = 0;
    if (p0 + 64 <= n) {
        const uint4 *v = reinterpret_cast<const uint4 *>(flags + p0);
        uint4 q[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = v[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t w[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // four flag bytes -> four bits: the products land on bits 24..27 without carries
                const uint32_t nib = ((((w[j] >> bit) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu;
                out |= (uint64_t)nib << (16 * i + 4 * j);
            }
        }
    } else {
        for (int i = 0; i < 64; ++i)
            if (p0 + i < n && ((flags[p0 + i] >> bit) & 1u)) out |= 1ull << i;
    }
    return out;
}

// last head position + 1 inside each tile (0 = none)
__global__ void __launch_bounds__(kGfThreads)
flag_tile_last_head_kernel(const uint8_t *__restrict__ flags, uint64_t n,
                           unsigned long long *__restrict__ tile_last)
{
    __shared__ unsigned long long s_max[kGfThreads / 32];
    const uint64_t p0 = (uint64_t)blockIdx.x * kGfTile + (uint64_t)threadIdx.x * kGfPerThread;
    const uint64_t bits = (p0 < n) ? load_flag_bits64(flags, n, p0, 0) : 0ull;
    unsigned long long last = bits ? p0 + (63 - __clzll((long long)bits)) + 1 : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long v = __shfl_xor_sync(0xffffffffu, last, o);
        last = v > last ? v : last;
    }
    if (lane_id() == 0) s_max[threadIdx.x >> 5] = last;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long m = 0;
        for (int w = 0; w < kGfThreads / 32; ++w) m = s_max[w] > m ? s_max[w] : m;
        tile_last[blockIdx.x] = m;
    }
}

// one CTA: carry[t] = max of tile_last over tiles < t (exclusive running max)
__global__ void __launch_bounds__(1024)
flag_tile_carry_kernel(const unsigned long long *__restrict__ tile_last, uint64_t n_tiles,
                       unsigned long long *__restrict__ carry)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint64_t base = 0; base < n_tiles; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long v = (i < n_tiles) ? tile_last[i] : 0;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc = u > inc ? u : inc;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        unsigned long long pre = s_carry;
        for (uint32_t w = 0; w < warp; ++w) pre = s_warp[w] > pre ? s_warp[w] : pre;
        unsigned long long excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0;
        excl = excl > pre ? excl : pre;
        if (i < n_tiles) carry[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = inc > pre ? inc : pre;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kGfThreads)
flag_group_hist_kernel(const uint8_t *__restrict__ flags, uint64_t n, uint64_t n_tiles,
                       const unsigned long long *__restrict__ carry, uint8_t skip_mask,
                       uint64_t min_group, uint64_t max_group, uint64_t max_bin,
                       unsigned long long *__restrict__ hist, unsigned long long *__restrict__ totals,
                       unsigned long long *__restrict__ big_list /* nullable: [0] count, then sizes >= 2048 */,
                       uint64_t big_capacity)
{
    __shared__ uint32_t s_small[kHistSmallBins];
    __shared__ unsigned long long s_total[3];
    __shared__ unsigned long long s_warp[kGfThreads / 32];
    for (int i = threadIdx.x; i < kHistSmallBins; i += kGfThreads) s_small[i] = 0;
    if (threadIdx.x < 3) s_total[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long total = 0, counted = 0, top_bin = 0;
    uint32_t ones = 0, twos = 0;  // groups of size 1 and 2 (a random genome has little else): no atomics
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const int skip_bit = skip_mask ? (__ffs((int)skip_mask) - 1) : 0;

    auto close_group = [&](uint64_t head, uint64_t size) {
        if (skip_mask && (flags[head] & skip_mask)) return;  // e.g. groups of ambiguous k-mers
        if (size >= min_group && (max_group == 0 || size <= max_group)) {
            total += size;
            ++counted;
            const uint64_t bin = size < max_bin ? size : max_bin;
            if (bin > top_bin) top_bin = bin;
            if (bin == 1) ++ones;
            else if (bin == 2) ++twos;
            else if (bin < kHistSmallBins) atomicAdd(&s_small[bin], 1u);
            else if (big_list) {   // spectrum mode: the exact sizes of the few large groups are listed
                const unsigned long long slot = atomicAdd(&big_list[0], 1ull);
                if (slot < big_capacity) big_list[1 + slot] = size;
            } else atomicAdd(&hist[bin], 1ull);
        }
    };

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t p0 = tile * kGfTile + (uint64_t)threadIdx.x * kGfPerThread;
        const uint64_t bits = (p0 < n) ? load_flag_bits64(flags, n, p0, 0) : 0ull;
        const unsigned long long last = bits ? p0 + (63 - __clzll((long long)bits)) + 1 : 0ull;
        unsigned long long inc = last;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc = u > inc ? u : inc;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        unsigned long long pre = carry[tile];
        for (uint32_t w = 0; w < warp; ++w) pre = s_warp[w] > pre ? s_warp[w] : pre;
        unsigned long long incoming = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) incoming = 0;
        incoming = incoming > pre ? incoming : pre;  // position + 1 of the latest head before this thread
        // groups of one k-mer whose head and successor head are both in this thread's flags: bit tricks
        const uint64_t pairs = bits & (bits << 1);    // bit i: heads at i-1 and at i
        uint64_t fast = pairs;
        if (skip_mask) fast &= ~(load_flag_bits64(flags, n, p0, skip_bit) << 1);  // skipped groups: dropped
        if (min_group <= 1) {
            const uint32_t c = (uint32_t)__popcll(fast);
            ones += c;
            total += c;
            counted += c;
            if (c && top_bin < 1) top_bin = 1;
        }
        uint64_t slow = bits & ~pairs;                // heads that close a longer or a cross-thread group
        while (slow) {
            const int i = __ffsll((long long)slow) - 1;
            slow &= slow - 1;
            const uint64_t lower = bits & ((1ull << i) - 1ull);
            const unsigned long long prev = lower ? p0 + (63 - __clzll((long long)lower)) + 1 : incoming;
            if (prev) close_group(prev - 1, (p0 + i) - (prev - 1));
        }
        // the thread that owns the last position also closes the last group
        if (p0 < n && n - 1 < p0 + kGfPerThread) {
            const unsigned long long prev = bits ? last : incoming;
            if (prev) close_group(prev - 1, n - (prev - 1));
        }
        __syncthreads();
    }

    total = warp_sum(total);
    counted = warp_sum(counted);
    ones = warp_sum(ones);
    twos = warp_sum(twos);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, top_bin, o);
        top_bin = other > top_bin ? other : top_bin;
    }
    if (lane == 0) {
        if (ones) atomicAdd(&s_small[1], ones);
        if (twos) atomicAdd(&s_small[2], twos);
        atomicAdd(&s_total[0], total);
        atomicAdd(&s_total[1], counted);
        atomicMax(&s_total[2], top_bin);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kHistSmallBins; i += kGfThreads) {
        const uint32_t c = s_small[i];
        if (c && (uint64_t)i <= max_bin) atomicAdd(&hist[i], (unsigned long long)c);
    }
    if (threadIdx.x < 2 && s_total[threadIdx.x]) atomicAdd(&totals[threadIdx.x], s_total[threadIdx.x]);
    if (threadIdx.x == 2 && s_total[2]) atomicMax(&totals[2], s_total[2]);
}

// hist (host, max_bin+1 int64, may be NULL), total and number of groups counted, straight from flags
int flag_group_hist_device(const uint8_t *d_flags, uint64_t n, uint8_t skip_mask, uint64_t min_group,
                           uint64_t max_group, uint64_t max_bin, int64_t *h_hist, int64_t *h_total,
                           int64_t *h_counted, cudaStream_t st)
{
    if (min_group < 1) {
        set_error("min_group_size (%llu) must be >= 1", (unsigned long long)min_group);
        return GK_ERR_ARG;
    }
    if (max_group != 0 && max_group < min_group) {
        set_error("max_group_size must be >= min_group_size");
        return GK_ERR_ARG;
    }
    if (max_bin < 1) {
        set_error("max_counts_bin (%llu) must be >= 1", (unsigned long long)max_bin);
        return GK_ERR_ARG;
    }
    const size_t hist_bytes = (size_t)(max_bin + 1) * 8;
    set_last_hist_single(0, 0);
    if (h_total) *h_total = 0;
    if (h_counted) *h_counted = 0;
    if (n == 0) return GK_OK;
    const uint64_t tiles = (n + kGfTile - 1) / kGfTile;
    DeviceBuffer buf, scan;
    GK_TRY(buf.alloc(hist_buffer_bytes(max_bin), st));
    GK_TRY(scan.alloc((size_t)tiles * 16, st));
    GK_CUDA(cudaMemsetAsync(buf.ptr, 0, hist_bytes + 32, st));
    unsigned long long *d_hist = buf.as<unsigned long long>();
    unsigned long long *d_totals = d_hist + (max_bin + 1);
    unsigned long long *d_last = scan.as<unsigned long long>();
    unsigned long long *d_carry = d_last + tiles;
    flag_tile_last_head_kernel<<<(unsigned)tiles, kGfThreads, 0, st>>>(d_flags, n, d_last);
    GK_LAUNCH_CHECK();
    flag_tile_carry_kernel<<<1, 1024, 0, st>>>(d_last, tiles, d_carry);
    GK_LAUNCH_CHECK();
    uint64_t grid = (uint64_t)sm_count() * 8;
    if (grid > tiles) grid = tiles;
    flag_group_hist_kernel<<<(unsigned)grid, kGfThreads, 0, st>>>(d_flags, n, tiles, d_carry, skip_mask,
                                                                   min_group, max_group, max_bin, d_hist,
                                                                   d_totals, nullptr, 0);
    GK_LAUNCH_CHECK();
    return collect_hist(d_hist, max_bin, h_hist, h_total, h_counted, st);
}

// The whole group-size spectrum of a sorted order in one go, left on the device (no synchronise): d_out holds
// kHistSmallBins counts for the sizes below 2048, then the number of larger groups, then up to big_capacity of
// their exact sizes.  gk_index_sort computes it behind the sort so that every later group-count query with the
// sort length and no filter is answered on the host (any min/max group size, any max_counts_bin).
int flag_group_spectrum_device(const uint8_t *d_flags, uint64_t n, unsigned long long *d_out, uint64_t big_capacity,
                               cudaStream_t st)
{
    GK_CUDA(cudaMemsetAsync(d_out, 0, (size_t)(kHistSmallBins + 1) * 8, st));
    if (n == 0) return GK_OK;
    const uint64_t tiles = (n + kGfTile - 1) / kGfTile;
    DeviceBuffer scan, totals;
    GK_TRY(scan.alloc((size_t)tiles * 16, st));
    GK_TRY(totals.alloc(32, st));
    GK_CUDA(cudaMemsetAsync(totals.ptr, 0, 32, st));
    unsigned long long *d_last = scan.as<unsigned long long>();
    unsigned long long *d_carry = d_last + tiles;
    flag_tile_last_head_kernel<<<(unsigned)tiles, kGfThreads, 0, st>>>(d_flags, n, d_last);
    GK_LAUNCH_CHECK();
    flag_tile_carry_kernel<<<1, 1024, 0, st>>>(d_last, tiles, d_carry);
    GK_LAUNCH_CHECK();
    uint64_t grid = (uint64_t)sm_count() * 8;
    if (grid > tiles) grid = tiles;
    flag_group_hist_kernel<<<(unsigned)grid, kGfThreads, 0, st>>>(d_flags, n, tiles, d_carry, 0, 1, 0, ~0ull, d_out,
                                                                   totals.as<unsigned long long>(),
                                                                   d_out + kHistSmallBins, big_capacity);
    GK_LAUNCH_CHECK();
    return GK_OK;
}
int spectrum_small_bins() { return kHistSmallBins; }

// ---- k-mer filters as device predicates (kmers.py:14-259) ------------------------------------------
// returns 1 pass, 0 fail, -1 where the reference raises ValueError
__device__ __forceinline__ int eval_filter(const uint8_t *__restrict__ sba, uint64_t len, uint64_t s,
                                           int id, int64_t p0, int64_t p1, int64_t p2)
{
    switch (id) {
    case GK_FILTER_KEEP_ALL:
        return 1;
    case GK_FILTER_NO_AMBIGUOUS:  // kmers.py:209-227
        if (s + (uint64_t)p0 > len) return -1;
        for (int64_t i = 0; i < p0; ++i) {
            const uint32_t b = sba[s + i];
            if (b == kSep) return -1;
            if (!is_acgt(b)) return 0;
        }
        return 1;
    case GK_FILTER_MIN_LENGTH:  // kmers.py:30-32 -> :262-282
        for (int64_t i = 0; i < p0; ++i)
            if (s + i >= len || sba[s + i] == kSep) return 0;
        return 1;
    case GK_FILTER_HOMOPOLYMER: {  // kmers.py:63-98
        if (s + (uint64_t)p1 - 1 >= len) return -1;
        if (p1 < p0) return 1;
        int64_t run = 1;
        for (int64_t i = 1; i < p1; ++i) {
            const uint32_t b = sba[s + i];
            if (b == kSep) return -1;
            if (b == sba[s + i - 1]) {
                if (++run > p0) return 0;
            } else {
                run = 1;
            }
        }
        return 1;
    }
    case GK_FILTER_GC_COUNT: {  // kmers.py:150-190
        if (p1 < p0) return 0;
        int64_t gc = 0;
        for (int64_t i = 0; i < p2; ++i) {
            if (s + i >= len) return -1;
            const uint32_t b = sba[s + i];
            if (b == kSep) return -1;
            if (b == 'G' || b == 'C')
                if (++gc > p1) return 0;
        }
        return (p0 <= gc && gc <= p1) ? 1 : 0;
    }
    case GK_FILTER_NGG_PAM:  // kmers.py:232-259
        if (s + 23 > len) return -1;
        return (sba[s + 21] == 'G' && sba[s + 22] == 'G') ? 1 : 0;
    default:
        return -1;
    }
}

template <typename IdxT>
__global__ void __launch_bounds__(256)
filter_flags_kernel(const uint8_t *__restrict__ sba, uint64_t len, const IdxT *__restrict__ idx,
                    uint64_t n, int id, int64_t p0, int64_t p1, int64_t p2,
                    uint8_t *__restrict__ flags, int *__restrict__ err)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        const int v = eval_filter(sba, len, (uint64_t)idx[r], id, p0, p1, p2);
        if (v < 0) atomicExch(err, 1);
        flags[r] = v > 0 ? 4 : 0;  // kFlagPass
    }
}

int filter_flags_device(const uint8_t *d_sba, uint64_t sba_len, const void *d_idx, int idx_bytes,
                        uint64_t n, const gk_filter &f, uint8_t *d_flags, cudaStream_t st)
{
    if (n == 0) return GK_OK;
    if (f.id < GK_FILTER_KEEP_ALL || f.id > GK_FILTER_NGG_PAM) {
        set_error("unknown filter id %d", f.id);
        return GK_ERR_ARG;
    }
    DeviceBuffer err;
    GK_TRY(err.alloc(4, st));
    GK_CUDA(cudaMemsetAsync(err.ptr, 0, 4, st));
    const int grid = launch_grid(n, 256);
    if (idx_bytes == 4)
        filter_flags_kernel<uint32_t><<<grid, 256, 0, st>>>(d_sba, sba_len, (const uint32_t *)d_idx, n,
                                                            f.id, f.p0, f.p1, f.p2, d_flags,
                                                            err.as<int>());
    else
        filter_flags_kernel<uint64_t><<<grid, 256, 0, st>>>(d_sba, sba_len, (const uint64_t *)d_idx, n,
                                                            f.id, f.p0, f.p1, f.p2, d_flags,
                                                            err.as<int>());
    GK_LAUNCH_CHECK();
    int h_err = 0;
    GK_CUDA(cudaMemcpyAsync(&h_err, err.ptr, 4, cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    if (h_err) {
        set_error("k-mer filter raised: a k-mer of the requested length runs past the end of its "
                  "record or of the sequence byte array");
        return GK_ERR_ARG;
    }
    return GK_OK;
}

}  // namespace gk

using namespace gk;

extern "C" {

int gk_rle_keys(const uint64_t *d_keys_sorted, uint64_t n, uint64_t *d_offsets_out,
                uint64_t *h_n_groups, void *stream)
{
    if (n && (!d_keys_sorted || !d_offsets_out)) {
        set_error("gk_rle_keys: null buffer");
        return GK_ERR_ARG;
    }
    cudaStream_t st = as_stream(stream);
    if (h_n_groups) *h_n_groups = 0;
    if (n == 0) return GK_OK;
    DeviceBuffer flags;
    GK_TRY(flags.alloc((size_t)((n + 15) & ~15ull), st));
    GK_TRY(key_flags_device(d_keys_sorted, n, 0, flags.as<uint8_t>(), st));
    uint64_t count = 0;
    GK_TRY(select_flagged(flags.as<uint8_t>(), n, kFlagHead, 8, d_offsets_out, nullptr, nullptr, nullptr,
                          nullptr, &count, st));
    if (h_n_groups) *h_n_groups = count;
    return GK_OK;
}

int gk_group_size_hist(const uint64_t *d_offsets, uint64_t n_groups, uint64_t n, uint64_t min_group,
                       uint64_t max_group, uint64_t max_bin, int64_t *h_hist_out,
                       int64_t *h_total_out, void *stream)
{
    if (h_hist_out) memset(h_hist_out, 0, (size_t)(max_bin + 1) * 8);
    return group_hist_device(d_offsets, 8, n_groups, n, min_group, max_group, max_bin, h_hist_out,
                             h_total_out, nullptr, as_stream(stream));
}

}  // extern "C"
