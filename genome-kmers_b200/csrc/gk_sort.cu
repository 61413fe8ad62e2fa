// gk_sort.cu -- stable LSD "onesweep" radix sort of (64-bit key, 32/64-bit value) pairs
// (north_star subsystem 2; replaces numba.misc.quicksort behind kmers.py:1644-1648).
//
// Structure (after Adinets & Merrill, "Onesweep", 2022 -- restated, not copied):
//   1. digit_histogram_kernel: ONE read of the keys builds the 256-bin histogram of every digit
//      position (shared-memory privatised, flushed with 64-bit global atomics).
//   2. scan_histogram_kernel: exclusive scan per digit position -> global bin bases.
//   3. onesweep_kernel, once per 8-bit digit, least significant first.  A CTA claims the next
//      tile with an atomic ticket (so look-back never waits on an unscheduled tile), ranks its
//      keys with eight warp ballots per digit (stable: items are visited in memory order), publishes
//      its per-bin counts, resolves its global bin offsets by decoupled look-back over the
//      preceding tiles' status words, reorders the tile in shared memory and streams it out so
//      consecutive threads write consecutive addresses inside each bin.
// Every pass reads 12 (16) bytes and writes 12 (16) bytes per pair: the 2*W*N bytes of
// SURVEY.md 8d.  Ties leave in input order, so starting from ascending start indices the
// result is the reference's break_ties=True order (kmers.py:1710-1711).
//
// Memory layout: keys and values are separate arrays (SoA) so each warp-wide access is one
// fully used 256-byte / 128-byte segment; tile status is [tile][256] words so a look-back step
// of the 256 bin-threads is one coalesced 1 KB read.
#include <stdlib.h>

#include "gk_common.cuh"

namespace gk {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kMaxPasses = 8;

// A "digit" is normally 8 key bits; the multi-GPU partition step reuses the same kernels with
// digit = destination rank = number of splitters <= key (splitters sorted, at most 255).

__device__ __forceinline__ uint32_t splitter_digit(const uint64_t *s_split, uint32_t n_split, uint64_t key)
{
    uint32_t lo = 0, hi = n_split;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (key < s_split[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// ---- 1. histogram ------------------------------------------------------------------------------
constexpr int kHistThreads = 512;

#ifndef GK_HIST_MINB
#define GK_HIST_MINB 3   // 38 registers, no spills, one wave: 0.45 -> 0.34 ms for 2e8 keys (0: 40 registers + spills, 4 CTAs asked, 3 fit)
#endif
#if GK_HIST_MINB
#define GK_HIST_BOUNDS __launch_bounds__(kHistThreads, GK_HIST_MINB)
#define GK_HIST_GRID_PER_SM GK_HIST_MINB
#else
#define GK_HIST_BOUNDS __launch_bounds__(kHistThreads)
#define GK_HIST_GRID_PER_SM 4
#endif
template <typename KeyT>
__global__ void GK_HIST_BOUNDS
digit_histogram_kernel(const KeyT *__restrict__ keys, uint64_t n, int begin_bit, int end_bit,
                       const uint64_t *__restrict__ splitters, uint32_t n_split,
                       unsigned long long *__restrict__ g_hist /* [passes][256] */, int split_amb = 0)
{
    __shared__ uint32_t s_hist[kMaxPasses][kRadix];
    __shared__ uint64_t s_split[kRadix];
    const int passes = splitters ? 1 : (end_bit - begin_bit + kRadixBits - 1) / kRadixBits;
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += kHistThreads)
        (&s_hist[0][0])[i] = 0;
    if (splitters && threadIdx.x < n_split) s_split[threadIdx.x] = splitters[threadIdx.x];
    __syncthreads();

    const uint64_t stride = (uint64_t)gridDim.x * kHistThreads;
    auto add = [&](uint64_t key) {
        if (splitters) {
            // split_amb: keys of class 0 (ambiguous windows) are counted per destination in bins n_parts ...
            const uint32_t d = splitter_digit(s_split, n_split, key);
            atomicAdd(&s_hist[0][d + ((split_amb && !(key & 1ull)) ? n_split + 1 : 0u)], 1u);
            return;
        }
#pragma unroll
        for (int p = 0; p < kMaxPasses; ++p) {
            if (p < passes) {
                const int lo = begin_bit + p * kRadixBits;
                const int bits = (end_bit - lo < kRadixBits) ? end_bit - lo : kRadixBits;
                const uint32_t d = (uint32_t)(key >> lo) & ((1u << bits) - 1u);
                atomicAdd(&s_hist[p][d], 1u);
            }
        }
    };
    // several 16-byte loads in flight per thread: with one, 2 x 512 threads per SM keep ~16 KB in flight, half of
    // what the HBM latency-bandwidth product asks for
    constexpr int kLoads = 4;
    if constexpr (sizeof(KeyT) == 8) {  // two keys per 16-byte load
        const uint64_t n2 = n / 2;
        const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(keys);
        uint64_t i = (uint64_t)blockIdx.x * kHistThreads + threadIdx.x;
        for (; i + (kLoads - 1) * stride < n2; i += kLoads * stride) {
            ulonglong2 v[kLoads];
#pragma unroll
            for (int u = 0; u < kLoads; ++u) v[u] = k2[i + u * stride];
#pragma unroll
            for (int u = 0; u < kLoads; ++u) {
                add(v[u].x);
                add(v[u].y);
            }
        }
        for (; i < n2; i += stride) {
            ulonglong2 v = k2[i];
            add(v.x);
            add(v.y);
        }
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) add(keys[n - 1]);
    } else {                            // four keys per 16-byte load
        const uint64_t n4 = n / 4;
        const uint4 *k4 = reinterpret_cast<const uint4 *>(keys);
        uint64_t i = (uint64_t)blockIdx.x * kHistThreads + threadIdx.x;
        for (; i + (kLoads - 1) * stride < n4; i += kLoads * stride) {
            uint4 v[kLoads];
#pragma unroll
            for (int u = 0; u < kLoads; ++u) v[u] = k4[i + u * stride];
#pragma unroll
            for (int u = 0; u < kLoads; ++u) { add(v[u].x); add(v[u].y); add(v[u].z); add(v[u].w); }
        }
        for (; i < n4; i += stride) {
            const uint4 v = k4[i];
            add(v.x); add(v.y); add(v.z); add(v.w);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0)
            for (uint64_t j = n4 * 4; j < n; ++j) add((uint64_t)keys[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += kHistThreads) {
        uint32_t c = (&s_hist[0][0])[i];
        if (c) atomicAdd(&g_hist[i], (unsigned long long)c);
    }
}

// ---- 2. exclusive scan per digit position ------------------------------------------------------
__global__ void __launch_bounds__(kRadix)
scan_histogram_kernel(const unsigned long long *__restrict__ g_hist,
                      unsigned long long *__restrict__ g_base)
{
    __shared__ unsigned long long s_warp[kRadix / 32];
    const int p = blockIdx.x, t = threadIdx.x;
    const unsigned long long c = g_hist[p * kRadix + t];
    unsigned long long inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
        if ((t & 31) >= o) inc += v;
    }
    if ((t & 31) == 31) s_warp[t >> 5] = inc;
    __syncthreads();
    unsigned long long pre = 0;
    for (int w = 0; w < (t >> 5); ++w) pre += s_warp[w];
    g_base[p * kRadix + t] = pre + inc - c;
}

// ---- 3. onesweep pass --------------------------------------------------------------------------
// Status word: top 2 bits = state (0 empty, 1 tile-local count, 2 inclusive prefix), rest = value.
template <typename StatusT>
struct StatusTraits;
template <>
struct StatusTraits<uint32_t> {
    static constexpr int kShift = 30;
    static constexpr uint32_t kMask = (1u << 30) - 1u;
};
template <>
struct StatusTraits<uint64_t> {
    static constexpr int kShift = 62;
    static constexpr uint64_t kMask = (1ull << 62) - 1ull;
};

constexpr uint32_t kLookbackSpinLimit = 1u << 24;  // then flag an error instead of hanging the GPU
constexpr int kLookbackBatch = 4;                  // predecessor status words loaded per round trip

// Multi-GPU partition straight into the destination ranks' receive buffers (peer memory over NVLink):
// digit d's pairs go to keys[d] / vals[d]; bin_base[d] holds the offset of this rank's segment there.
// key_base[d] is subtracted from every key that goes to rank d (the rank's key range starts there, so its local
// sort sees keys relative to its own range); with skip_amb, keys of class 0 (ambiguous windows, which travel as
// run-length fragments instead) are counted in digit n_dest and written nowhere.
constexpr int kMaxPeers = 16;
struct PeerTable {
    uint64_t *keys[kMaxPeers];
    void *vals[kMaxPeers];
    uint64_t key_base[kMaxPeers];
    uint32_t n_dest;
    uint32_t skip_amb;
};

template <typename ValT, int THREADS, int IPT, typename KeyT = uint64_t>
struct OnesweepSmem {
    static constexpr int kTile = THREADS * IPT;
    static constexpr int kWarps = THREADS / 32;
    KeyT keys[kTile];
    ValT vals[kTile];
    uint32_t warp_cnt[kWarps][kRadix];  // per-warp digit counts, then running tile positions
    unsigned long long global_off[kRadix];  // bin's global start minus its start inside the tile
    uint32_t global_off32[kRadix];          // the same modulo 2^32 (enough when n < 2^32)
    uint64_t splitters[kRadix];
    uint64_t *peer_keys[kMaxPeers];
    ValT *peer_vals[kMaxPeers];
    uint64_t peer_key_base[kMaxPeers];
    uint32_t peer_n_dest, peer_skip_amb;
    uint32_t bin_excl[kRadix];
    uint32_t warp_sums[kRadix / 32];
    uint32_t tile;
};

template <typename StatusT>
__device__ __forceinline__ StatusT load_status(const StatusT *p)
{
    return *reinterpret_cast<const volatile StatusT *>(p);
}

// peers = lanes of this warp whose 8-bit digit equals mine, from eight ballots (about 3 SASS
// instructions per bit).  No shared memory is touched: that pipe is the scarce resource here.
__device__ __forceinline__ uint32_t digit_peers(uint32_t d)
{
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < kRadixBits; ++b) {
        uint32_t m;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b32 t;\n\t"
            "and.b32 t, %1, %2;\n\t"
            "setp.ne.u32 p, t, 0;\n\t"
            "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
            "@!p not.b32 %0, %0;\n\t"
            "}"
            : "=r"(m)
            : "r"(d), "r"(1u << b));
        peers &= m;
    }
    return peers;
}

#ifdef GK_PROBE
// Tuning aid (tools/probe_onesweep.py builds a second library with -DGK_PROBE): cycles thread 0 of every tile
// spends in each phase, and how far bin 0's look-back walks / how often it finds a status word unpublished.
__device__ unsigned long long g_probe[16];
#define GK_PROBE_MARK(i)                                                            \
    if (t == 0) {                                                                   \
        const long long probe_now = clock64();                                      \
        atomicAdd(&g_probe[i], (unsigned long long)(probe_now - probe_t0));         \
        probe_t0 = probe_now;                                                       \
    }
#else
#define GK_PROBE_MARK(i)
#endif

// One pass.  Phases (profiles/r01_onesweep.md explains the choices):
//   A  claim a tile, load its pairs, count digits per warp with shared-memory atomics
//   B  bin threads: per-warp exclusive offsets, tile totals -> publish the tile-local counts EARLY
//   C  stable ranking: the lanes of a warp that hold the same digit find each other with eight ballots
//      (MATCH.ANY issues once per ~180 cycles on sm_100a; a shared-memory atomicOr mask costs three more
//      bank-conflicted shared accesses per item, and the shared-memory pipe is what bounds this kernel),
//      read the warp's running position for the digit, and drop their pair straight into its tile-sorted
//      slot in shared memory
//   D  bin threads: decoupled look-back, several predecessor status words per round trip
//   E  stream the tile out; consecutive threads write consecutive addresses inside each bin
// KeyT = uint32_t: the same pass for 32-bit keys (8-byte pairs with 32-bit values); never with PARTITION.
template <typename ValT, typename StatusT, int THREADS, int IPT, int MINB, bool PARTITION, typename KeyT = uint64_t>
__global__ void __launch_bounds__(THREADS, MINB)
onesweep_kernel(const KeyT *__restrict__ keys_in, KeyT *__restrict__ keys_out,
                const ValT *__restrict__ vals_in, ValT *__restrict__ vals_out, uint64_t n,
                int shift, uint32_t digit_mask, const uint64_t *__restrict__ splitters, uint32_t n_split,
                const unsigned long long *__restrict__ bin_base, uint32_t *__restrict__ tile_counter,
                StatusT *__restrict__ status, int *__restrict__ err, const PeerTable *__restrict__ peer)
{
    using Smem = OnesweepSmem<ValT, THREADS, IPT, KeyT>;
    using ST = StatusTraits<StatusT>;
    constexpr int kTile = Smem::kTile;
    constexpr int kWarps = Smem::kWarps;
    static_assert(THREADS >= kRadix, "one thread per bin is needed for the look-back");
    static_assert(!PARTITION || sizeof(KeyT) == 8, "splitters are 64-bit keys");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &s = *reinterpret_cast<Smem *>(smem_raw);

    const uint32_t t = threadIdx.x;
    const uint32_t lane = t & 31u, warp = t >> 5;

    // ---- A ------------------------------------------------------------------------------------------
#ifdef GK_PROBE
    long long probe_t0 = clock64();
#endif
    if (t == 0) s.tile = atomicAdd(tile_counter, 1u);
    for (int i = t; i < kWarps * kRadix; i += THREADS) (&s.warp_cnt[0][0])[i] = 0;
    if (PARTITION && t < n_split) s.splitters[t] = splitters[t];
    if (PARTITION && t < kMaxPeers) {
        s.peer_keys[t] = peer ? peer->keys[t] : nullptr;
        s.peer_vals[t] = peer ? reinterpret_cast<ValT *>(peer->vals[t]) : nullptr;
        s.peer_key_base[t] = peer ? peer->key_base[t] : 0ull;
        if (t == 0) {
            s.peer_n_dest = peer ? peer->n_dest : 0u;
            s.peer_skip_amb = peer ? peer->skip_amb : 0u;
        }
    }
    __syncthreads();
    const uint32_t skip_amb_digit = (PARTITION && s.peer_skip_amb) ? s.peer_n_dest : 0xffffffffu;
    auto digit_of = [&](KeyT k) -> uint32_t {
        if (PARTITION) {
            if (skip_amb_digit != 0xffffffffu && !((uint64_t)k & 1ull)) return skip_amb_digit;
            return splitter_digit(s.splitters, n_split, (uint64_t)k);
        }
        return (uint32_t)(k >> shift) & digit_mask;
    };
    const uint64_t tile = s.tile;
    const uint64_t tile_base = tile * (uint64_t)kTile;
    const bool full_tile = (n - tile_base) >= (uint64_t)kTile;
    const uint32_t tile_valid = full_tile ? (uint32_t)kTile : (uint32_t)(n - tile_base);

    // warp-striped inside a warp-contiguous chunk, so memory order == (j, lane)
    KeyT key[IPT];
    ValT val[IPT];
    const uint64_t warp_base = tile_base + (uint64_t)warp * (32 * IPT);
    if (full_tile) {
#pragma unroll
        for (int j = 0; j < IPT; ++j) key[j] = __ldcs(keys_in + warp_base + j * 32 + lane);
#pragma unroll
        for (int j = 0; j < IPT; ++j) val[j] = __ldcs(vals_in + warp_base + j * 32 + lane);
    } else {
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const uint64_t g = warp_base + j * 32 + lane;
            const bool ok = g < n;
            // pads must rank behind every real pair of the tile: all ones has the largest digit in a sorting
            // pass and the last destination in a partition pass; when ambiguous pairs are dropped (digit n_dest,
            // beyond the last destination) the pad is made even so that it lands in that discard digit too
            key[j] = ok ? keys_in[g] : (KeyT)(~(KeyT)0 - (KeyT)(skip_amb_digit != 0xffffffffu ? 1 : 0));
            val[j] = ok ? vals_in[g] : (ValT)0;
        }
    }
    uint32_t *my_cnt = s.warp_cnt[warp];
#pragma unroll
    for (int j = 0; j < IPT; ++j) atomicAdd(&my_cnt[digit_of(key[j])], 1u);
    __syncthreads();
    GK_PROBE_MARK(0)

    // ---- B ------------------------------------------------------------------------------------------
    uint32_t bin_count = 0;
    if (t < kRadix) {
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = s.warp_cnt[w][t];
            s.warp_cnt[w][t] = sum;
            sum += c;
        }
        bin_count = sum;
        // (volatile: the two status stores of this thread must both reach memory, in order)
        *const_cast<volatile StatusT *>(status + tile * kRadix + t) =
            ((StatusT)1 << ST::kShift) | (StatusT)bin_count;
    }
    uint32_t inc = bin_count;  // exclusive scan of the tile totals over the 256 bins
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += v;
    }
    if (t < kRadix && lane == 31) s.warp_sums[warp] = inc;
    __syncthreads();
    uint32_t excl_in_tile = 0;
    if (t < kRadix) {
        uint32_t pre = 0;
        for (uint32_t w = 0; w < warp; ++w) pre += s.warp_sums[w];
        excl_in_tile = pre + inc - bin_count;
        s.bin_excl[t] = excl_in_tile;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s.warp_cnt[w][t] += excl_in_tile;  // -> running tile positions
    }
    __syncthreads();
    GK_PROBE_MARK(1)

    // ---- C ------------------------------------------------------------------------------------------
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const uint32_t d = digit_of(key[j]);
        const uint32_t peers = digit_peers(d);
        const uint32_t before = peers & lt_mask;
        const uint32_t base = my_cnt[d];                       // peers read the same word (broadcast)
        __syncwarp();
        if (before == 0) my_cnt[d] = base + __popc(peers);     // lowest peer lane advances the position
        __syncwarp();
        const uint32_t pos = base + __popc(before);
        s.keys[pos] = key[j];
        s.vals[pos] = val[j];
    }
    GK_PROBE_MARK(2)

    // ---- D ------------------------------------------------------------------------------------------
#ifdef GK_PROBE
    uint32_t probe_walk = 0, probe_spin = 0;
#endif
    if (t < kRadix) {
        // a tile completes chip-wide every few dozen cycles while one L2 round trip takes hundreds, so
        // a one-at-a-time walk never catches up with the inclusive frontier: batch the status loads
        uint64_t excl = 0;
        bool failed = false, done = (tile == 0);
        int64_t tp = (int64_t)tile - 1;
        while (!done) {
            StatusT sv[kLookbackBatch];
#pragma unroll
            for (int u = 0; u < kLookbackBatch; ++u)
                sv[u] = (tp - u >= 0) ? load_status(status + (uint64_t)(tp - u) * kRadix + t)
                                      : ((StatusT)2 << ST::kShift);  // before tile 0: inclusive 0
#pragma unroll
            for (int u = 0; u < kLookbackBatch; ++u) {
                if (done) break;
                uint32_t spins = 0;
                while ((sv[u] >> ST::kShift) == 0) {
                    if (++spins > kLookbackSpinLimit) { failed = true; break; }
#ifdef GK_PROBE
                    ++probe_spin;
#endif
                    __nanosleep(64);
                    sv[u] = load_status(status + (uint64_t)(tp - u) * kRadix + t);
                }
                if (failed) { done = true; break; }
                excl += (uint64_t)(sv[u] & ST::kMask);
#ifdef GK_PROBE
                ++probe_walk;
#endif
                if ((sv[u] >> ST::kShift) == 2) done = true;
            }
            tp -= kLookbackBatch;
        }
        if (failed) atomicExch(err, 1);
        *const_cast<volatile StatusT *>(status + tile * kRadix + t) =
            ((StatusT)2 << ST::kShift) | (StatusT)(excl + bin_count);
        const unsigned long long off = bin_base[t] + excl - (unsigned long long)excl_in_tile;
        s.global_off[t] = off;
        s.global_off32[t] = (uint32_t)off;
    }
    GK_PROBE_MARK(3)
    __syncthreads();
    GK_PROBE_MARK(4)
#ifdef GK_PROBE
    if (t == 0) {
        atomicAdd(&g_probe[8], (unsigned long long)probe_walk);
        atomicAdd(&g_probe[9], (unsigned long long)probe_spin);
        atomicAdd(&g_probe[10], 1ull);
        atomicMax(&g_probe[11], (unsigned long long)probe_walk);
    }
#endif

    // ---- E ------------------------------------------------------------------------------------------
    // 32-bit offsets when every destination index fits (StatusT is 32-bit exactly when n < 2^30)
    constexpr bool kNarrow = sizeof(StatusT) == 4;
    if (PARTITION && peer) {
        // every pair goes to its destination rank's buffer: long contiguous runs per destination
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const uint32_t p = t + j * THREADS;
            if (p < tile_valid) {
                const KeyT k = s.keys[p];
                const uint32_t d = digit_of(k);
                if (d < s.peer_n_dest) {
                    const uint64_t dst = (uint64_t)(s.global_off[d] + p);
                    s.peer_keys[d][dst] = (uint64_t)k - s.peer_key_base[d];
                    s.peer_vals[d][dst] = s.vals[p];
                }
            }
        }
        return;
    }
    if (full_tile) {
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const uint32_t p = t + j * THREADS;
            const KeyT k = s.keys[p];
            const uint32_t d = digit_of(k);
            const uint64_t dst = kNarrow ? (uint64_t)(uint32_t)(s.global_off32[d] + p)
                                         : (uint64_t)(s.global_off[d] + p);
            __stcs(keys_out + dst, k);
            __stcs(vals_out + dst, s.vals[p]);
        }
        GK_PROBE_MARK(5)
    } else {
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const uint32_t p = t + j * THREADS;
            if (p < tile_valid) {
                const KeyT k = s.keys[p];
                const uint64_t dst = (uint64_t)(s.global_off[digit_of(k)] + p);
                keys_out[dst] = k;
                vals_out[dst] = s.vals[p];
            }
        }
    }
}


// ---- small inputs: the whole sort in ONE launch ------------------------------------------------------
// The refinement stages sort lists of a few ten thousand elements by up to four 64-bit words.  Through
// onesweep_kernel that is one launch (plus a status memset) per digit, each far too small to fill the GPU:
// the time is launch latency.  One CTA does every digit pass of such a sort here: each warp owns a contiguous
// segment (so memory order == (warp, row, lane)), counts its digits into a private row of shared counters,
// the rows are scanned across warps and bins, and every warp then ranks its rows with the same ballot
// matching as the big kernel.  A digit that is the same for all elements costs a copy instead of a scatter,
// so the ping-pong parity stays a host-side constant.  Global data is read with ld.cg: the buffers are
// rewritten by this CTA between passes and must not be served from a stale L1 line.
constexpr int kSmallThreads = 1024;
constexpr int kSmallWarps = kSmallThreads / 32;
constexpr uint64_t kSmallSortMax = 1ull << 16;
constexpr int kSmallBatch = 8;

struct SmallSortSmem {
    uint32_t cnt[kSmallWarps][kRadix];
    uint32_t wsum[kRadix / 32];
    int skip;
};

// Body shared by the kernels below: sort n <= kSmallSortMax pairs on key bits [begin_bit, end_bit) by
// ping-ponging between the a and b arrays; with land_in_a the result is brought back to the a arrays when the
// number of passes is odd.  All threads of the CTA call it.
template <typename ValT>
__device__ __forceinline__ void small_sort_body(SmallSortSmem &sm, uint64_t *keys_a, uint64_t *keys_b, ValT *vals_a,
                                                ValT *vals_b, uint32_t n, int begin_bit, int end_bit, bool land_in_a)
{
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t rows = (n + kSmallThreads - 1) / kSmallThreads;  // rows of 32 elements per warp
    const uint32_t seg0 = warp * rows * 32u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint64_t *kin = keys_a, *kout = keys_b;
    ValT *vin = vals_a, *vout = vals_b;
    for (int lo = begin_bit; lo < end_bit; lo += kRadixBits) {
        const int bits = (end_bit - lo < kRadixBits) ? end_bit - lo : kRadixBits;
        const uint32_t mask = (1u << bits) - 1u;
        for (int i = t; i < kSmallWarps * kRadix; i += kSmallThreads) (&sm.cnt[0][0])[i] = 0;
        if (t == 0) sm.skip = 0;
        __syncthreads();
        for (uint32_t r0 = 0; r0 < rows; r0 += kSmallBatch) {   // several rows in flight per L2 round trip
            uint64_t kk[kSmallBatch];
#pragma unroll
            for (int u = 0; u < kSmallBatch; ++u) {
                const uint32_t i = seg0 + (r0 + u) * 32u + lane;
                kk[u] = (r0 + u < rows && i < n) ? __ldcg(kin + i) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < kSmallBatch; ++u) {
                const uint32_t i = seg0 + (r0 + u) * 32u + lane;
                if (r0 + u < rows && i < n) atomicAdd(&sm.cnt[warp][(uint32_t)(kk[u] >> lo) & mask], 1u);
            }
        }
        __syncthreads();
        uint32_t total = 0;
        if (t < kRadix) {
            uint32_t sum = 0;
#pragma unroll 8
            for (int w = 0; w < kSmallWarps; ++w) {
                const uint32_t c = sm.cnt[w][t];
                sm.cnt[w][t] = sum;
                sum += c;
            }
            total = sum;
            if (total == n) sm.skip = 1;  // every element has this digit: the pass is the identity
        }
        uint32_t inc = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += v;
        }
        if (t < kRadix && lane == 31) sm.wsum[warp] = inc;
        __syncthreads();
        if (t < kRadix) {
            uint32_t pre = 0;
            for (uint32_t w = 0; w < warp; ++w) pre += sm.wsum[w];
            const uint32_t excl = pre + inc - total;
#pragma unroll 8
            for (int w = 0; w < kSmallWarps; ++w) sm.cnt[w][t] += excl;
        }
        __syncthreads();
        if (sm.skip) {
            for (uint32_t i = t; i < n; i += kSmallThreads) {
                kout[i] = __ldcg(kin + i);
                vout[i] = __ldcg(vin + i);
            }
        } else {
            uint32_t *my_cnt = sm.cnt[warp];
            for (uint32_t r0 = 0; r0 < rows; r0 += kSmallBatch) {
                uint64_t kk[kSmallBatch];
                ValT vv[kSmallBatch];
#pragma unroll
                for (int u = 0; u < kSmallBatch; ++u) {
                    const uint32_t i = seg0 + (r0 + u) * 32u + lane;
                    const bool ok = r0 + u < rows && i < n;
                    kk[u] = ok ? __ldcg(kin + i) : 0ull;
                    vv[u] = ok ? __ldcg(vin + i) : (ValT)0;
                }
#pragma unroll
                for (int u = 0; u < kSmallBatch; ++u) {
                    const uint32_t i = seg0 + (r0 + u) * 32u + lane;
                    const bool ok = r0 + u < rows && i < n;
                    const uint32_t live = __ballot_sync(0xffffffffu, ok);
                    if (live == 0) break;  // (uniform: rows past the end of the array)
                    const uint32_t d = (uint32_t)(kk[u] >> lo) & mask;
                    const uint32_t peers = digit_peers(d) & live;
                    const uint32_t before = peers & lt_mask;
                    const uint32_t base = ok ? my_cnt[d] : 0u;
                    __syncwarp();
                    if (ok && before == 0) my_cnt[d] = base + __popc(peers);
                    __syncwarp();
                    if (ok) {
                        const uint32_t pos = base + __popc(before);
                        kout[pos] = kk[u];
                        vout[pos] = vv[u];
                    }
                }
            }
        }
        __syncthreads();  // orders this pass's global writes before the next pass's reads (same CTA)
        uint64_t *tk = kin; kin = kout; kout = tk;
        ValT *tv = vin; vin = vout; vout = tv;
    }
    if (land_in_a && kin != keys_a) {   // odd number of passes: bring the result home
        for (uint32_t i = t; i < n; i += kSmallThreads) {
            keys_a[i] = __ldcg(kin + i);
            vals_a[i] = __ldcg(vin + i);
        }
        __syncthreads();
    }
}

// ranges == nullptr: one list of n pairs.  Otherwise one CTA per (lo, hi) slot range of the arrays, sorted in
// place (ranges longer than kSmallSortMax are left to the host driver).
template <typename ValT>
__global__ void __launch_bounds__(kSmallThreads, 1)
small_sort_kernel(uint64_t *keys_a, uint64_t *keys_b, ValT *vals_a, ValT *vals_b, uint32_t n, int begin_bit,
                  int end_bit, const unsigned long long *__restrict__ ranges)
{
    __shared__ SmallSortSmem sm;
    if (ranges) {
        const unsigned long long lo_ = ranges[2 * blockIdx.x], hi_ = ranges[2 * blockIdx.x + 1];
        if (hi_ - lo_ > kSmallSortMax || hi_ - lo_ < 2) return;
        n = (uint32_t)(hi_ - lo_);
        keys_a += lo_; keys_b += lo_; vals_a += lo_; vals_b += lo_;
    }
    small_sort_body<ValT>(sm, keys_a, keys_b, vals_a, vals_b, n, begin_bit, end_bit, ranges != nullptr);
}

// Device-driven repair of the long prefix runs that the flags pass found out of order (gk_group.cu:
// tie_fix_flags_kernel lists the positions).  CTA b takes listed position b: finds the slot range of its
// prefix bucket (the array is sorted by prefix), drops out if an earlier list entry lies in the same bucket,
// sorts the bucket by its low key bits in place and rewrites its head / ambiguous flags.  Nothing comes back
// to the host unless *status ends up non-zero: bit 1 a bucket is longer than kSmallSortMax, bit 2 more
// positions than the list holds.
template <typename ValT>
__global__ void __launch_bounds__(kSmallThreads, 1)
repair_buckets_kernel(uint64_t *keys, uint64_t *keys_tmp, ValT *vals, ValT *vals_tmp, uint64_t n, int lo_bits,
                      int class_bit, uint8_t *__restrict__ flags, const unsigned int *__restrict__ count,
                      const unsigned long long *__restrict__ list, int *__restrict__ status,
                      unsigned long long *__restrict__ big /* [0] count, then up to kBigBucketCap (lo, hi) pairs */)
{
    __shared__ SmallSortSmem sm;
    __shared__ unsigned long long s_range[2];
    const unsigned int total = *count;
    if (total > (unsigned int)kDescentCap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(status, 2);
        return;
    }
    for (unsigned int b = blockIdx.x; b < total; b += gridDim.x) {
        if (threadIdx.x == 0) {
            const uint64_t pre = keys[list[b]] >> lo_bits;
            uint64_t lo = 0, hi = n;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                if ((__ldcg(keys + mid) >> lo_bits) < pre) lo = mid + 1; else hi = mid;
            }
            s_range[0] = lo;
            hi = n;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                if ((__ldcg(keys + mid) >> lo_bits) <= pre) lo = mid + 1; else hi = mid;
            }
            s_range[1] = lo;
        }
        __syncthreads();
        const uint64_t lo = s_range[0], hi = s_range[1];
        int dup = 0;
        for (unsigned int j = threadIdx.x; j < b; j += kSmallThreads) dup |= (list[j] >= lo && list[j] < hi) ? 1 : 0;
        dup = __syncthreads_or(dup);   // (also orders the reads of s_range before the next iteration's write)
        if (dup) continue;
        if (hi - lo > kSmallSortMax) {   // too long for one CTA: listed for the host, which runs the radix passes
            if (threadIdx.x == 0) {
                const unsigned long long slot = atomicAdd(&big[0], 1ull);
                if (slot < (unsigned long long)kBigBucketCap) {
                    big[1 + 2 * slot] = lo;
                    big[2 + 2 * slot] = hi;
                    atomicOr(status, 1);
                } else {
                    atomicOr(status, 2);
                }
            }
            continue;
        }
        const uint32_t len = (uint32_t)(hi - lo);
        small_sort_body<ValT>(sm, keys + lo, keys_tmp + lo, vals + lo, vals_tmp + lo, len, 0, lo_bits, true);
        for (uint32_t i = threadIdx.x; i < len; i += kSmallThreads) {
            const uint64_t k = __ldcg(keys + lo + i);
            const bool amb = class_bit && !(k & 1ull);
            const bool head = (i == 0) || __ldcg(keys + lo + i - 1) != k;
            flags[lo + i] = amb ? kFlagAmb : (head ? kFlagHead : 0);
        }
        __syncthreads();
    }
}

static bool small_sort_enabled()
{
    const char *e = getenv("GK_SMALL_SORT");
    return !(e && e[0] == '0');
}

// ---- host driver ---------------------------------------------------------------------------------
struct PassArgs {
    const void *kin; void *kout; const void *vin; void *vout; uint64_t n;
    int shift, bits; const uint64_t *splitters; uint32_t n_split;
    const unsigned long long *bin_base; uint32_t *tile_counter; void *status; int *err;
    const PeerTable *peer;
};

template <typename ValT, typename StatusT, int THREADS, int IPT, int MINB, bool PARTITION,
          typename KeyT = uint64_t>
static int launch_pass_impl(const PassArgs &a, cudaStream_t st)
{
    using Smem = OnesweepSmem<ValT, THREADS, IPT, KeyT>;
    auto kernel = onesweep_kernel<ValT, StatusT, THREADS, IPT, MINB, PARTITION, KeyT>;
    GK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    GK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 (int)cudaSharedmemCarveoutMaxShared));
    const uint64_t tiles = (a.n + Smem::kTile - 1) / Smem::kTile;
    kernel<<<(unsigned)tiles, THREADS, sizeof(Smem), st>>>(
        (const KeyT *)a.kin, (KeyT *)a.kout, (const ValT *)a.vin, (ValT *)a.vout, a.n, a.shift,
        (1u << a.bits) - 1u, a.splitters, a.n_split, a.bin_base, a.tile_counter, (StatusT *)a.status, a.err, a.peer);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

template <typename ValT, typename StatusT, int THREADS, int IPT, int MINB>
static int launch_pass(const PassArgs &a, cudaStream_t st)
{
    if (a.splitters) return launch_pass_impl<ValT, StatusT, THREADS, IPT, MINB, true>(a, st);
    return launch_pass_impl<ValT, StatusT, THREADS, IPT, MINB, false>(a, st);
}

// Tile shapes (threads, pairs per thread, CTAs per SM the register budget is compiled for).
// GK_SORT_CFG overrides the default for tuning runs.
struct SortConfig { int threads, ipt, minb; };
constexpr SortConfig kSortConfigs[] = {
    {256, 16, 3},  // 0: 4096-pair tiles, 61 KB smem, <=80 regs
    {512, 8, 2},   // 1: 4096-pair tiles, 69 KB smem, <=64 regs
    {384, 12, 2},  // 2: 4608-pair tiles
    {256, 24, 2},  // 3: 6144-pair tiles
    {512, 12, 1},  // 4: 6144-pair tiles
    {512, 16, 1},  // 5: 8192-pair tiles
    {256, 28, 2},  // 6: 7168-pair tiles
    {256, 32, 2},  // 7: 8192-pair tiles, 112 KB smem, 128 regs: the largest tile two CTAs per SM allow
};
constexpr int kNumSortConfigs = (int)(sizeof(kSortConfigs) / sizeof(kSortConfigs[0]));
constexpr int kDefaultSortConfig = 7;

static int sort_config_id_raw()
{
    // read on every call (a getenv per sort is nothing next to the launches): one tuning process can compare
    // tile shapes, tools/bench_sort.py --env GK_SORT_CFG
    const char *e = getenv("GK_SORT_CFG");
    if (e && *e) {
        const int v = atoi(e);
        if (v >= 0 && v < kNumSortConfigs) return v;
    }
    return kDefaultSortConfig;
}

// 64-bit values make a pair 16 bytes: the two largest tiles would no longer fit two CTAs per SM
static int sort_config_id(int val_bytes)
{
    const int id = sort_config_id_raw();
    return (val_bytes == 8 && (id == 6 || id == 7)) ? 3 : id;
}

template <typename ValT, typename StatusT>
static int dispatch_pass(int cfg, const PassArgs &a, cudaStream_t st)
{
    switch (cfg) {
    case 1: return launch_pass<ValT, StatusT, 512, 8, 2>(a, st);
    case 2: return launch_pass<ValT, StatusT, 384, 12, 2>(a, st);
    case 3: return launch_pass<ValT, StatusT, 256, 24, 2>(a, st);
    case 4: return launch_pass<ValT, StatusT, 512, 12, 1>(a, st);
    case 5: return launch_pass<ValT, StatusT, 512, 16, 1>(a, st);
    case 6: return launch_pass<ValT, StatusT, 256, 28, 2>(a, st);
    case 7: return launch_pass<ValT, StatusT, 256, 32, 2>(a, st);
    default: return launch_pass<ValT, StatusT, 256, 16, 3>(a, st);
    }
}

// A caller that runs many small sorts back to back (gk_index_sort: refinement, doubling rounds) lends a
// device error word for the look-back guard; such sorts return without synchronising the stream and the
// caller checks the word once at its own final synchronise.
static thread_local int *g_deferred_err = nullptr;
void set_deferred_sort_error_word(int *d_err) { g_deferred_err = d_err; }

// Shared driver: histogram(s) + scan + `passes` onesweep launches.  splitters != nullptr selects
// the single partition pass (digit = destination rank); h_bin_counts (256 entries, optional)
// receives the first pass's histogram.  d_vals_final (optional): a third value buffer that receives the
// sorted values whatever the number of passes (the outputs alternate final / alt counting back from the
// last pass, so the input buffer is only read by the first pass).
static int run_onesweep(uint64_t *d_keys, uint64_t *d_keys_alt, void *d_vals, void *d_vals_alt,
                        int val_bytes, uint64_t n, int begin_bit, int end_bit,
                        const uint64_t *d_splitters, uint32_t n_split, int *result_in_alt,
                        unsigned long long *h_bin_counts, cudaStream_t st, SortTiming *timing,
                        const unsigned long long *d_pre_hist = nullptr, void *d_vals_final = nullptr)
{
    if (timing) { timing->drop(); timing->hist_ms = 0.f; timing->passes_ms = 0.f; timing->passes = 0; }
    if (val_bytes != 4 && val_bytes != 8) {
        set_error("radix_sort_pairs: val_bytes must be 4 or 8");
        return GK_ERR_ARG;
    }
    if (begin_bit < 0 || end_bit > 64 || begin_bit > end_bit) {
        set_error("radix_sort_pairs: bad bit range [%d, %d)", begin_bit, end_bit);
        return GK_ERR_ARG;
    }
    if (result_in_alt) *result_in_alt = 0;
    if (h_bin_counts) memset(h_bin_counts, 0, kRadix * sizeof(unsigned long long));
    const int passes = d_splitters ? 1 : (end_bit - begin_bit + kRadixBits - 1) / kRadixBits;
    if (n == 0 || passes == 0 || (n == 1 && !d_splitters)) {
        if (d_vals_final && n) GK_CUDA(cudaMemcpyAsync(d_vals_final, d_vals, (size_t)n * val_bytes,
                                                       cudaMemcpyDeviceToDevice, st));
        return GK_OK;
    }
    if ((reinterpret_cast<uintptr_t>(d_keys) & 15u) || (reinterpret_cast<uintptr_t>(d_keys_alt) & 15u)) {
        set_error("radix_sort_pairs: key buffers must be 16-byte aligned");
        return GK_ERR_ARG;
    }
    const bool deferred = g_deferred_err && !h_bin_counts;
    if (timing) {
        for (auto &e : timing->ev) GK_CUDA(cudaEventCreate(&e));
        timing->pending = true;
        timing->passes = passes;
    }

    // ---- a list this small is sorted by one CTA in one launch ----------------------------------------------
    if (!d_splitters && !h_bin_counts && !d_pre_hist && !d_vals_final && n <= kSmallSortMax && small_sort_enabled()) {
        if (timing) { GK_CUDA(cudaEventRecord(timing->ev[0], st)); GK_CUDA(cudaEventRecord(timing->ev[1], st)); }
        if (val_bytes == 4)
            small_sort_kernel<uint32_t><<<1, kSmallThreads, 0, st>>>(d_keys, d_keys_alt, (uint32_t *)d_vals,
                                                                     (uint32_t *)d_vals_alt, (uint32_t)n, begin_bit,
                                                                     end_bit, nullptr);
        else
            small_sort_kernel<uint64_t><<<1, kSmallThreads, 0, st>>>(d_keys, d_keys_alt, (uint64_t *)d_vals,
                                                                     (uint64_t *)d_vals_alt, (uint32_t)n, begin_bit,
                                                                     end_bit, nullptr);
        GK_LAUNCH_CHECK();
        if (result_in_alt) *result_in_alt = passes & 1;
        if (timing) {
            GK_CUDA(cudaEventRecord(timing->ev[2], st));
            if (!deferred) { GK_CUDA(cudaStreamSynchronize(st)); timing->resolve(); }
        }
        return GK_OK;
    }

    const int cfg = sort_config_id(val_bytes);
    const int tile = kSortConfigs[cfg].threads * kSortConfigs[cfg].ipt;
    const uint64_t tiles = (n + tile - 1) / tile;
    const bool wide = n >= (1ull << 30);
    const size_t status_bytes = (size_t)tiles * kRadix * (wide ? 8 : 4);
    const size_t hist_bytes = (size_t)kMaxPasses * kRadix * sizeof(unsigned long long);
    // temp layout: [hist][base][counters (kMaxPasses u32) + err (int)][status]
    DeviceBuffer temp;
    const size_t ctr_bytes = 64;
    GK_TRY(temp.alloc(2 * hist_bytes + ctr_bytes + status_bytes, st));
    unsigned long long *d_hist = temp.as<unsigned long long>();
    unsigned long long *d_base = d_hist + kMaxPasses * kRadix;
    uint32_t *d_ctr = reinterpret_cast<uint32_t *>(d_base + kMaxPasses * kRadix);
    int *d_err = deferred ? g_deferred_err : reinterpret_cast<int *>(d_ctr + kMaxPasses);
    void *d_status = reinterpret_cast<unsigned char *>(d_ctr) + ctr_bytes;
    GK_CUDA(cudaMemsetAsync(temp.ptr, 0, 2 * hist_bytes + ctr_bytes, st));

    if (timing) GK_CUDA(cudaEventRecord(timing->ev[0], st));
    int hist_grid = sm_count() * GK_HIST_GRID_PER_SM;
    {
        uint64_t need = (n / 2 + kHistThreads - 1) / kHistThreads;
        if (need < 1) need = 1;
        if ((uint64_t)hist_grid > need) hist_grid = (int)need;
    }
    if (!d_pre_hist) {  // the producer of the keys may have counted the digits already (gk_pack.cu)
        digit_histogram_kernel<<<hist_grid, kHistThreads, 0, st>>>(d_keys, n, begin_bit, end_bit, d_splitters,
                                                                  n_split, d_hist);
        GK_LAUNCH_CHECK();
    }
    scan_histogram_kernel<<<passes, kRadix, 0, st>>>(d_pre_hist ? d_pre_hist : d_hist, d_base);
    GK_LAUNCH_CHECK();
    if (timing) GK_CUDA(cudaEventRecord(timing->ev[1], st));

    uint64_t *kin = d_keys, *kout = d_keys_alt;
    void *vin = d_vals;
    for (int p = 0; p < passes; ++p) {
        const int lo = begin_bit + p * kRadixBits;
        const int bits = d_splitters ? kRadixBits : ((end_bit - lo < kRadixBits) ? end_bit - lo : kRadixBits);
        void *vout;
        if (d_vals_final) vout = ((passes - 1 - p) & 1) ? d_vals_alt : d_vals_final;
        else vout = (vin == d_vals) ? d_vals_alt : d_vals;
        GK_CUDA(cudaMemsetAsync(d_status, 0, status_bytes, st));
        PassArgs pa = {kin, kout, vin, vout, n, lo, bits, d_splitters, n_split, d_base + p * kRadix,
                       d_ctr + p, d_status, d_err, nullptr};
        int rc;
        if (val_bytes == 4)
            rc = wide ? dispatch_pass<uint32_t, uint64_t>(cfg, pa, st) : dispatch_pass<uint32_t, uint32_t>(cfg, pa, st);
        else
            rc = wide ? dispatch_pass<uint64_t, uint64_t>(cfg, pa, st) : dispatch_pass<uint64_t, uint32_t>(cfg, pa, st);
        GK_TRY(rc);
        uint64_t *tk = kin; kin = kout; kout = tk;
        vin = vout;
    }
    if (result_in_alt) *result_in_alt = passes & 1;
    if (timing) GK_CUDA(cudaEventRecord(timing->ev[2], st));
    if (deferred) return GK_OK;  // scratch is released in stream order; the caller checks the error word

    int h_err = 0;
    GK_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (h_bin_counts)
        GK_CUDA(cudaMemcpyAsync(h_bin_counts, d_hist, kRadix * sizeof(unsigned long long),
                                cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    if (timing) timing->resolve();
    if (h_err) {
        set_error("radix_sort_pairs: decoupled look-back timed out");
        return GK_ERR_INTERNAL;
    }
    return GK_OK;
}

// ---- 32-bit keys ---------------------------------------------------------------------------------
// The same stable LSD onesweep for (u32 key, u32/u64 value) pairs: 8 bytes per pair with 32-bit values, so a
// pass moves 16 B per pair instead of 24 and a tile of the same shared-memory size holds 1.5x the pairs.
// Tile shapes for this pair size; GK_SORT32_CFG selects one (read per call, tools/bench_sort.py --env).
constexpr SortConfig kSort32Configs[] = {
    {256, 32, 2},  // 0:  8192-pair tiles,  64 KB of pairs
    {256, 40, 2},  // 1: 10240-pair tiles,  80 KB
    {256, 48, 2},  // 2: 12288-pair tiles,  96 KB: as much shared memory as the 64-bit default
    {256, 24, 3},  // 3:  6144-pair tiles,  48 KB, three CTAs per SM
};
constexpr int kNumSort32Configs = (int)(sizeof(kSort32Configs) / sizeof(kSort32Configs[0]));
constexpr int kDefaultSort32Config = 0;

static int sort32_config_id(int val_bytes)
{
    int id = kDefaultSort32Config;
    const char *e = getenv("GK_SORT32_CFG");
    if (e && *e) {
        const int v = atoi(e);
        if (v >= 0 && v < kNumSort32Configs) id = v;
    }
    // 64-bit values: 12 bytes per pair, the two largest tiles no longer fit two CTAs per SM
    return (val_bytes == 8 && (id == 1 || id == 2)) ? 0 : id;
}

template <typename ValT, typename StatusT>
static int dispatch_pass32(int cfg, const PassArgs &a, cudaStream_t st)
{
    if constexpr (sizeof(ValT) == 4) {  // (sort32_config_id never selects these for 64-bit values)
        if (cfg == 1) return launch_pass_impl<ValT, StatusT, 256, 40, 2, false, uint32_t>(a, st);
        if (cfg == 2) return launch_pass_impl<ValT, StatusT, 256, 48, 2, false, uint32_t>(a, st);
    }
    if (cfg == 3) return launch_pass_impl<ValT, StatusT, 256, 24, 3, false, uint32_t>(a, st);
    return launch_pass_impl<ValT, StatusT, 256, 32, 2, false, uint32_t>(a, st);
}

int radix_sort_pairs32_device(uint32_t *d_keys, uint32_t *d_keys_alt, void *d_vals, void *d_vals_alt, int val_bytes,
                              uint64_t n, int begin_bit, int end_bit, int *result_in_alt, cudaStream_t st)
{
    if (val_bytes != 4 && val_bytes != 8) {
        set_error("radix_sort_pairs32: val_bytes must be 4 or 8");
        return GK_ERR_ARG;
    }
    if (begin_bit < 0 || end_bit > 32 || begin_bit > end_bit) {
        set_error("radix_sort_pairs32: bad bit range [%d, %d)", begin_bit, end_bit);
        return GK_ERR_ARG;
    }
    if (result_in_alt) *result_in_alt = 0;
    const int passes = (end_bit - begin_bit + kRadixBits - 1) / kRadixBits;
    if (n < 2 || passes == 0) return GK_OK;
    if ((reinterpret_cast<uintptr_t>(d_keys) & 15u) || (reinterpret_cast<uintptr_t>(d_keys_alt) & 15u)) {
        set_error("radix_sort_pairs32: key buffers must be 16-byte aligned");
        return GK_ERR_ARG;
    }
    const int cfg = sort32_config_id(val_bytes);
    const int tile = kSort32Configs[cfg].threads * kSort32Configs[cfg].ipt;
    const uint64_t tiles = (n + tile - 1) / tile;
    const bool wide = n >= (1ull << 30);
    const size_t status_bytes = (size_t)tiles * kRadix * (wide ? 8 : 4);
    const size_t hist_bytes = (size_t)kMaxPasses * kRadix * sizeof(unsigned long long);
    // temp layout as in run_onesweep: [hist][base][counters (kMaxPasses u32) + err (int)][status]
    DeviceBuffer temp;
    const size_t ctr_bytes = 64;
    GK_TRY(temp.alloc(2 * hist_bytes + ctr_bytes + status_bytes, st));
    unsigned long long *d_hist = temp.as<unsigned long long>();
    unsigned long long *d_base = d_hist + kMaxPasses * kRadix;
    uint32_t *d_ctr = reinterpret_cast<uint32_t *>(d_base + kMaxPasses * kRadix);
    int *d_err = reinterpret_cast<int *>(d_ctr + kMaxPasses);
    void *d_status = reinterpret_cast<unsigned char *>(d_ctr) + ctr_bytes;
    GK_CUDA(cudaMemsetAsync(temp.ptr, 0, 2 * hist_bytes + ctr_bytes, st));
    int hist_grid = sm_count() * GK_HIST_GRID_PER_SM;
    {
        uint64_t need = (n / 4 + kHistThreads - 1) / kHistThreads;
        if (need < 1) need = 1;
        if ((uint64_t)hist_grid > need) hist_grid = (int)need;
    }
    digit_histogram_kernel<uint32_t><<<hist_grid, kHistThreads, 0, st>>>(d_keys, n, begin_bit, end_bit, nullptr, 0,
                                                                         d_hist);
    GK_LAUNCH_CHECK();
    scan_histogram_kernel<<<passes, kRadix, 0, st>>>(d_hist, d_base);
    GK_LAUNCH_CHECK();
    uint32_t *kin = d_keys, *kout = d_keys_alt;
    void *vin = d_vals, *vout = d_vals_alt;
    for (int p = 0; p < passes; ++p) {
        const int lo = begin_bit + p * kRadixBits;
        const int bits = (end_bit - lo < kRadixBits) ? end_bit - lo : kRadixBits;
        GK_CUDA(cudaMemsetAsync(d_status, 0, status_bytes, st));
        PassArgs pa = {kin, kout, vin, vout, n, lo, bits, nullptr, 0, d_base + p * kRadix,
                       d_ctr + p, d_status, d_err, nullptr};
        int rc;
        if (val_bytes == 4)
            rc = wide ? dispatch_pass32<uint32_t, uint64_t>(cfg, pa, st) : dispatch_pass32<uint32_t, uint32_t>(cfg, pa, st);
        else
            rc = wide ? dispatch_pass32<uint64_t, uint64_t>(cfg, pa, st) : dispatch_pass32<uint64_t, uint32_t>(cfg, pa, st);
        GK_TRY(rc);
        uint32_t *tk = kin; kin = kout; kout = tk;
        void *tv = vin; vin = vout; vout = tv;
    }
    if (result_in_alt) *result_in_alt = passes & 1;
    int h_err = 0;
    GK_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    if (h_err) {
        set_error("radix_sort_pairs32: decoupled look-back timed out");
        return GK_ERR_INTERNAL;
    }
    return GK_OK;
}

int radix_sort_pairs_device(uint64_t *d_keys, uint64_t *d_keys_alt, void *d_vals, void *d_vals_alt,
                            int val_bytes, uint64_t n, int begin_bit, int end_bit,
                            int *result_in_alt, cudaStream_t st, SortTiming *timing,
                            const unsigned long long *d_pre_hist, void *d_vals_final)
{
    return run_onesweep(d_keys, d_keys_alt, d_vals, d_vals_alt, val_bytes, n, begin_bit, end_bit, nullptr,
                        0, result_in_alt, nullptr, st, timing, d_pre_hist, d_vals_final);
}

// Stable sort of disjoint slot ranges of (keys, vals) on key bits [0, end_bit), in place (the *_tmp arrays of
// the same size are scratch).  h_ranges: n_ranges (lo, hi) pairs, also on the device (d_ranges).  Ranges of up
// to kSmallSortMax slots share one launch (one CTA each); larger ones go through the onesweep passes.
int sort_segments_device(uint64_t *d_keys, uint64_t *d_keys_tmp, void *d_vals, void *d_vals_tmp, int val_bytes,
                         const unsigned long long *d_ranges, const unsigned long long *h_ranges, uint32_t n_ranges,
                         int end_bit, cudaStream_t st)
{
    if (n_ranges == 0) return GK_OK;
    bool any_small = false;
    for (uint32_t i = 0; i < n_ranges; ++i) {
        const uint64_t lo = h_ranges[2 * i], hi = h_ranges[2 * i + 1];
        if (hi - lo <= kSmallSortMax) { any_small = true; continue; }
        // 16-byte aligned key pointers are required: an odd start takes the slot before it along, whose larger
        // prefix bits keep it first when all 64 bits are sorted
        const uint64_t a = lo & ~1ull;
        const int eb = (a == lo) ? end_bit : 64;
        int in_alt = 0;
        unsigned char *v = (unsigned char *)d_vals, *vt = (unsigned char *)d_vals_tmp;
        GK_TRY(run_onesweep(d_keys + a, d_keys_tmp + a, v + a * val_bytes, vt + a * val_bytes, val_bytes, hi - a, 0,
                            eb, nullptr, 0, &in_alt, nullptr, st, nullptr));
        if (in_alt) {
            GK_CUDA(cudaMemcpyAsync(d_keys + a, d_keys_tmp + a, (size_t)(hi - a) * 8, cudaMemcpyDeviceToDevice, st));
            GK_CUDA(cudaMemcpyAsync(v + a * val_bytes, vt + a * val_bytes, (size_t)(hi - a) * val_bytes,
                                    cudaMemcpyDeviceToDevice, st));
        }
    }
    if (any_small) {
        if (val_bytes == 4)
            small_sort_kernel<uint32_t><<<n_ranges, kSmallThreads, 0, st>>>(
                d_keys, d_keys_tmp, (uint32_t *)d_vals, (uint32_t *)d_vals_tmp, 0, 0, end_bit, d_ranges);
        else
            small_sort_kernel<uint64_t><<<n_ranges, kSmallThreads, 0, st>>>(
                d_keys, d_keys_tmp, (uint64_t *)d_vals, (uint64_t *)d_vals_tmp, 0, 0, end_bit, d_ranges);
        GK_LAUNCH_CHECK();
    }
    return GK_OK;
}

// ---- big out-of-order buckets that are one ambiguous key plus a few strangers ---------------------------------
// The bucket of the all-N windows holds millions of equal (class 0) keys; on a genome of any size a handful of
// other k-mers share its 32-bit prefix (a pure "TAAAAAAAAAAAAAAA..." 31-mer lands among them) and leave the bucket
// out of order.  Re-sorting millions of equal keys for that is wasteful: the strangers are collected (every
// element whose key differs from the bucket's majority key M), sorted, and put where they belong -- those below
// M at the front of the bucket, those above at its end; every other slot holds M.  The start indices of the M
// slots are not moved: M is an ambiguous key, and frag_expand_device rewrites all ambiguous slots from the
// fragment list anyway.  Buckets that do not fit this pattern (M pure, or more than kRareCap strangers) keep
// their entry in the list for the host.
constexpr int kRareCap = 8192;
struct RareList {
    unsigned long long count[kBigBucketCap];
    unsigned long long majority[kBigBucketCap];
    uint64_t key[kBigBucketCap][kRareCap];
    uint64_t pos[kBigBucketCap][kRareCap];
};

// the key that most of seven evenly spaced probes of the bucket agree on (a single probe can hit a stranger)
__device__ __forceinline__ uint64_t majority_probe(const uint64_t *__restrict__ keys, uint64_t lo, uint64_t hi)
{
    uint64_t probe[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) probe[i] = keys[lo + ((hi - lo) * (uint64_t)(2 * i + 1)) / 14];
    uint64_t best = probe[0];
    int best_votes = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        int votes = 0;
#pragma unroll
        for (int j = 0; j < 7; ++j) votes += (probe[j] == probe[i]) ? 1 : 0;
        if (votes > best_votes) { best_votes = votes; best = probe[i]; }
    }
    return best;
}

template <typename ValT>
__global__ void __launch_bounds__(256)
collect_rare_kernel(const uint64_t *__restrict__ keys, const unsigned long long *__restrict__ big /* [0] count */,
                    int class_bit, RareList *__restrict__ rare)
{
    const unsigned int n_big = (unsigned int)(big[0] < (unsigned long long)kBigBucketCap ? big[0] : kBigBucketCap);
    for (unsigned int b = 0; b < n_big; ++b) {
        const uint64_t lo = big[1 + 2 * b], hi = big[2 + 2 * b];
        const uint64_t m = majority_probe(keys, lo, hi);
        if (blockIdx.x == 0 && threadIdx.x == 0) rare->majority[b] = m;
        if (!class_bit || (m & 1ull)) continue;   // a pure majority key: its start indices would have to move
        const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
        for (uint64_t p = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < hi; p += stride) {
            const uint64_t k = keys[p];
            if (k != m) {
                const unsigned long long s_ = atomicAdd(&rare->count[b], 1ull);
                if (s_ < (unsigned long long)kRareCap) {
                    rare->key[b][s_] = k;
                    rare->pos[b][s_] = p;
                }
            }
        }
    }
}

// one CTA per big bucket: order the strangers by (key, old position), place them, restore M everywhere else
template <typename ValT>
__global__ void __launch_bounds__(kSmallThreads, 1)
place_rare_kernel(uint64_t *keys, ValT *vals, int class_bit, uint8_t *__restrict__ flags,
                  unsigned long long *__restrict__ big, RareList *__restrict__ rare, uint64_t *__restrict__ scratch,
                  int *__restrict__ status, int *__restrict__ handled_all)
{
    __shared__ SmallSortSmem sm;
    __shared__ uint32_t s_below;
    const unsigned int b = blockIdx.x;
    const unsigned int n_big = (unsigned int)(big[0] < (unsigned long long)kBigBucketCap ? big[0] : kBigBucketCap);
    if (b >= n_big) return;
    const uint64_t lo = big[1 + 2 * b], hi = big[2 + 2 * b];
    const uint64_t m = rare->majority[b];
    const unsigned long long cnt = rare->count[b];
    if (!class_bit || (m & 1ull) || cnt > (unsigned long long)kRareCap || cnt == 0) {   // not this kernel's case
        if (threadIdx.x == 0) atomicExch(handled_all, 0);
        return;
    }
    const uint32_t r = (uint32_t)cnt;
    // scratch rows of this bucket: [0] old positions, [1] ping-pong, [2] the strangers' start indices, [3] ping-pong
    uint64_t *pos_a = scratch + (size_t)b * 4 * kRareCap, *pos_b = pos_a + kRareCap;
    uint64_t *val_a = pos_b + kRareCap, *val_b = val_a + kRareCap;
    uint64_t *key_a = rare->key[b], *key_b = rare->pos[b];   // (pos is copied out first, then reused as ping-pong)
    for (uint32_t i = threadIdx.x; i < r; i += kSmallThreads) pos_a[i] = rare->pos[b][i];
    __syncthreads();
    // 1. by old position (restores the stable order the atomics lost), the key rides along as the value
    small_sort_body<uint64_t>(sm, pos_a, pos_b, key_a, key_b, r, 0, 40, true);
    // 2. stable by key, the old position rides along
    small_sort_body<uint64_t>(sm, key_a, key_b, pos_a, pos_b, r, 0, 64, true);
    // start indices of the strangers, read before anything is overwritten
    for (uint32_t i = threadIdx.x; i < r; i += kSmallThreads) val_a[i] = (uint64_t)vals[__ldcg(pos_a + i)];
    if (threadIdx.x == 0) s_below = 0;
    __syncthreads();
    uint32_t below = 0;
    for (uint32_t i = threadIdx.x; i < r; i += kSmallThreads) below += (__ldcg(key_a + i) < m) ? 1u : 0u;
    below = warp_sum(below);
    if ((threadIdx.x & 31u) == 0 && below) atomicAdd(&s_below, below);
    __syncthreads();
    const uint32_t n_lo = s_below, n_hi = r - n_lo;
    // old slots of the strangers go back to M (those inside the two end zones are overwritten right after)
    for (uint32_t i = threadIdx.x; i < r; i += kSmallThreads) {
        const uint64_t p = __ldcg(pos_a + i);
        keys[p] = m;
        flags[p] = kFlagAmb;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < r; i += kSmallThreads) {
        const uint64_t k = __ldcg(key_a + i);
        const uint64_t slot = (i < n_lo) ? lo + i : hi - n_hi + (i - n_lo);
        const bool amb = class_bit && !(k & 1ull);
        const bool head = (i == 0 && n_lo > 0) || (i == n_lo) || __ldcg(key_a + i - 1) != k;
        keys[slot] = k;
        vals[slot] = (ValT)__ldcg(val_a + i);
        flags[slot] = amb ? kFlagAmb : (head ? kFlagHead : 0);
    }
    (void)val_b;
    __syncthreads();
    if (threadIdx.x == 0) {   // this bucket is done: take it off the host's list
        big[1 + 2 * b] = 0;
        big[2 + 2 * b] = 0;
    }
}

// after place_rare_kernel: when every listed big bucket was handled (and nothing else is pending), the order
// is final and the fragment kernels may run
__global__ void clear_status_kernel(int *status, const int *handled_all, unsigned long long *big)
{
    if (*handled_all && (*status & ~1) == 0 && (*status & 1)) {
        *status = 0;
        big[0] = 0;
    }
}

int repair_big_ambiguous_buckets_on_device(uint64_t *d_keys, void *d_vals, int val_bytes, int class_bit, uint8_t *d_flags,
                                           unsigned long long *d_big, int *d_status, cudaStream_t st)
{
    DeviceBuffer rare, scratch, handled;
    GK_TRY(rare.alloc(sizeof(RareList), st));
    GK_TRY(scratch.alloc((size_t)kBigBucketCap * 4 * kRareCap * 8, st));
    GK_TRY(handled.alloc(4, st));
    GK_CUDA(cudaMemsetAsync(rare.ptr, 0, 2 * kBigBucketCap * sizeof(unsigned long long), st));
    GK_CUDA(cudaMemsetAsync(handled.ptr, 0xff, 4, st));
    const int grid = sm_count() * 8;
    if (val_bytes == 4) {
        collect_rare_kernel<uint32_t><<<grid, 256, 0, st>>>(d_keys, d_big, class_bit, rare.as<RareList>());
        GK_LAUNCH_CHECK();
        place_rare_kernel<uint32_t><<<kBigBucketCap, kSmallThreads, 0, st>>>(
            d_keys, (uint32_t *)d_vals, class_bit, d_flags, d_big, rare.as<RareList>(), scratch.as<uint64_t>(), d_status,
            handled.as<int>());
    } else {
        collect_rare_kernel<uint64_t><<<grid, 256, 0, st>>>(d_keys, d_big, class_bit, rare.as<RareList>());
        GK_LAUNCH_CHECK();
        place_rare_kernel<uint64_t><<<kBigBucketCap, kSmallThreads, 0, st>>>(
            d_keys, (uint64_t *)d_vals, class_bit, d_flags, d_big, rare.as<RareList>(), scratch.as<uint64_t>(), d_status,
            handled.as<int>());
    }
    GK_LAUNCH_CHECK();
    clear_status_kernel<<<1, 1, 0, st>>>(d_status, handled.as<int>(), d_big);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int repair_buckets_on_device(uint64_t *d_keys, uint64_t *d_keys_tmp, void *d_vals, void *d_vals_tmp, int val_bytes,
                             uint64_t n, int lo_bits, int class_bit, uint8_t *d_flags, const unsigned int *d_count,
                             const unsigned long long *d_list, int *d_status, unsigned long long *d_big,
                             cudaStream_t st)
{
    if (n == 0) return GK_OK;
    const int grid = 64;
    if (val_bytes == 4)
        repair_buckets_kernel<uint32_t><<<grid, kSmallThreads, 0, st>>>(d_keys, d_keys_tmp, (uint32_t *)d_vals,
                                                                        (uint32_t *)d_vals_tmp, n, lo_bits, class_bit,
                                                                        d_flags, d_count, d_list, d_status, d_big);
    else
        repair_buckets_kernel<uint64_t><<<grid, kSmallThreads, 0, st>>>(d_keys, d_keys_tmp, (uint64_t *)d_vals,
                                                                        (uint64_t *)d_vals_tmp, n, lo_bits, class_bit,
                                                                        d_flags, d_count, d_list, d_status, d_big);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

// Stable partition of the pairs by destination = number of splitters <= key.  Output always lands
// in the *_out buffers; h_counts[d] = pairs sent to destination d (n_parts entries).
int partition_pairs_device(uint64_t *d_keys, uint64_t *d_keys_out, void *d_vals, void *d_vals_out,
                           int val_bytes, uint64_t n, const uint64_t *d_splitters, uint32_t n_parts,
                           uint64_t *h_counts, cudaStream_t st)
{
    if (n_parts < 1 || n_parts > (uint32_t)kRadix) {
        set_error("partition_pairs: n_parts must be in [1, 256]");
        return GK_ERR_ARG;
    }
    unsigned long long bins[kRadix];
    int in_alt = 0;
    GK_TRY(run_onesweep(d_keys, d_keys_out, d_vals, d_vals_out, val_bytes, n, 0, 8, d_splitters, n_parts - 1,
                        &in_alt, bins, st, nullptr));
    for (uint32_t d = 0; d < n_parts; ++d) h_counts[d] = (n == 0) ? 0 : (uint64_t)bins[d];
    return GK_OK;
}

// Destination counts only (one read of the keys): the ranks exchange them to size and place the segments
// before any pair moves.
int partition_count_device(const uint64_t *d_keys, uint64_t n, const uint64_t *d_splitters, uint32_t n_parts,
                           uint64_t *h_counts, cudaStream_t st)
{
    if (n_parts < 1 || n_parts > (uint32_t)kRadix) {
        set_error("partition_count: n_parts must be in [1, 256]");
        return GK_ERR_ARG;
    }
    for (uint32_t d = 0; d < n_parts; ++d) h_counts[d] = 0;
    if (n == 0) return GK_OK;
    DeviceBuffer hist;
    GK_TRY(hist.alloc(kRadix * sizeof(unsigned long long), st));
    GK_CUDA(cudaMemsetAsync(hist.ptr, 0, hist.bytes, st));
    int grid = sm_count() * GK_HIST_GRID_PER_SM;
    uint64_t need = (n / 2 + kHistThreads - 1) / kHistThreads;
    if (need < 1) need = 1;
    if ((uint64_t)grid > need) grid = (int)need;
    digit_histogram_kernel<<<grid, kHistThreads, 0, st>>>(d_keys, n, 0, 8, d_splitters, n_parts - 1,
                                                          hist.as<unsigned long long>());
    GK_LAUNCH_CHECK();
    unsigned long long bins[kRadix];
    GK_CUDA(cudaMemcpyAsync(bins, hist.ptr, sizeof(bins), cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    for (uint32_t d = 0; d < n_parts; ++d) h_counts[d] = bins[d];
    return GK_OK;
}

// Destination counts left on the device, no synchronise: d_counts_out[d] = pure pairs for destination d,
// d_counts_out[n_parts + d] = ambiguous ones (class bit 0; only split off when class_bit is set).
int partition_count_split_device(const uint64_t *d_keys, uint64_t n, const uint64_t *d_splitters, uint32_t n_parts,
                                 int class_bit, unsigned long long *d_counts_out, cudaStream_t st)
{
    if (n_parts < 1 || n_parts > (uint32_t)kMaxPeers) {
        set_error("partition_count_split: n_parts must be in [1, %d]", kMaxPeers);
        return GK_ERR_ARG;
    }
    GK_CUDA(cudaMemsetAsync(d_counts_out, 0, (size_t)2 * n_parts * 8, st));
    if (n == 0) return GK_OK;
    DeviceBuffer hist;
    GK_TRY(hist.alloc(kRadix * sizeof(unsigned long long), st));
    GK_CUDA(cudaMemsetAsync(hist.ptr, 0, hist.bytes, st));
    int grid = sm_count() * GK_HIST_GRID_PER_SM;
    uint64_t need = (n / 2 + kHistThreads - 1) / kHistThreads;
    if (need < 1) need = 1;
    if ((uint64_t)grid > need) grid = (int)need;
    digit_histogram_kernel<<<grid, kHistThreads, 0, st>>>(d_keys, n, 0, 8, d_splitters ? d_splitters : d_keys,
                                                          n_parts - 1, hist.as<unsigned long long>(), class_bit);
    GK_LAUNCH_CHECK();
    GK_CUDA(cudaMemcpyAsync(d_counts_out, hist.ptr, (size_t)2 * n_parts * 8, cudaMemcpyDeviceToDevice, st));
    return GK_OK;
}

// One stable partition pass whose output lands in the destination ranks' buffers.  h_dst_keys / h_dst_vals:
// n_parts device pointers (local or peer-mapped); h_dst_offsets[d]: first element of this rank's segment
// in destination d; h_key_base (optional): subtracted from the keys sent to d; skip_amb: pairs of class 0 are
// dropped.  d_err (optional device word): receives the look-back guard's verdict and the call does not
// synchronise; without it the call synchronises and reports the failure itself.
int partition_pairs_peer_device(const uint64_t *d_keys, const void *d_vals, int val_bytes, uint64_t n,
                                const uint64_t *d_splitters, uint32_t n_parts, uint64_t *const *h_dst_keys,
                                void *const *h_dst_vals, const uint64_t *h_dst_offsets, const uint64_t *h_key_base,
                                int skip_amb, int *d_err_out, cudaStream_t st)
{
    if (n_parts < 1 || n_parts > (uint32_t)kMaxPeers) {
        set_error("partition_pairs_peer: n_parts must be in [1, %d]", kMaxPeers);
        return GK_ERR_ARG;
    }
    if (val_bytes != 4 && val_bytes != 8) {
        set_error("partition_pairs_peer: val_bytes must be 4 or 8");
        return GK_ERR_ARG;
    }
    if (n == 0) return GK_OK;
    const int cfg = sort_config_id(val_bytes);
    const int tile = kSortConfigs[cfg].threads * kSortConfigs[cfg].ipt;
    const uint64_t tiles = (n + tile - 1) / tile;
    const bool wide = n >= (1ull << 30);
    const size_t status_bytes = (size_t)tiles * kRadix * (wide ? 8 : 4);
    // temp layout: [bin_base 256 x u64][counter + err, 64 B][peer table][status]
    const size_t table_bytes = (sizeof(PeerTable) + 15) & ~(size_t)15;
    const size_t head_bytes = kRadix * 8 + 64 + table_bytes;
    DeviceBuffer temp;
    GK_TRY(temp.alloc(head_bytes + status_bytes, st));
    unsigned char *base = temp.as<unsigned char>();
    unsigned long long h_base[kRadix];
    memset(h_base, 0, sizeof(h_base));
    for (uint32_t d = 0; d < n_parts; ++d) h_base[d] = h_dst_offsets[d];
    PeerTable h_peer;
    memset(&h_peer, 0, sizeof(h_peer));
    for (uint32_t d = 0; d < n_parts; ++d) {
        h_peer.keys[d] = h_dst_keys[d];
        h_peer.vals[d] = h_dst_vals[d];
        h_peer.key_base[d] = h_key_base ? h_key_base[d] : 0ull;
    }
    h_peer.n_dest = n_parts;
    h_peer.skip_amb = skip_amb ? 1u : 0u;
    GK_CUDA(cudaMemsetAsync(base + kRadix * 8, 0, 64 + status_bytes + table_bytes, st));
    GK_CUDA(cudaMemcpyAsync(base, h_base, sizeof(h_base), cudaMemcpyHostToDevice, st));
    GK_CUDA(cudaMemcpyAsync(base + kRadix * 8 + 64, &h_peer, sizeof(h_peer), cudaMemcpyHostToDevice, st));
    // the two small uploads read pageable host memory: they have completed when the calls return
    uint32_t *d_ctr = reinterpret_cast<uint32_t *>(base + kRadix * 8);
    int *d_err = d_err_out ? d_err_out : reinterpret_cast<int *>(d_ctr + 8);
    PassArgs pa = {d_keys, nullptr, d_vals, nullptr, n, 0, kRadixBits, d_splitters ? d_splitters : d_keys,
                   n_parts - 1, reinterpret_cast<const unsigned long long *>(base), d_ctr,
                   base + head_bytes, d_err, reinterpret_cast<const PeerTable *>(base + kRadix * 8 + 64)};
    int rc;
    if (val_bytes == 4)
        rc = wide ? dispatch_pass<uint32_t, uint64_t>(cfg, pa, st) : dispatch_pass<uint32_t, uint32_t>(cfg, pa, st);
    else
        rc = wide ? dispatch_pass<uint64_t, uint64_t>(cfg, pa, st) : dispatch_pass<uint64_t, uint32_t>(cfg, pa, st);
    GK_TRY(rc);
    if (d_err_out) return GK_OK;   // the scratch is released in stream order; the caller reads the error word
    int h_err = 0;
    GK_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    if (h_err) {
        set_error("partition_pairs_peer: decoupled look-back timed out");
        return GK_ERR_INTERNAL;
    }
    return GK_OK;
}

}  // namespace gk

using namespace gk;

extern "C" int gk_radix_sort_pairs(uint64_t *d_keys, uint64_t *d_keys_alt, void *d_vals,
                                   void *d_vals_alt, int val_bytes, uint64_t n, int begin_bit,
                                   int end_bit, int *result_in_alt, void *stream)
{
    if (n && (!d_keys || !d_keys_alt || !d_vals || !d_vals_alt)) {
        set_error("gk_radix_sort_pairs: null buffer");
        return GK_ERR_ARG;
    }
    return radix_sort_pairs_device(d_keys, d_keys_alt, d_vals, d_vals_alt, val_bytes, n, begin_bit,
                                   end_bit, result_in_alt, as_stream(stream), nullptr, nullptr, nullptr);
}

extern "C" int gk_radix_sort_pairs32(uint32_t *d_keys, uint32_t *d_keys_alt, void *d_vals, void *d_vals_alt,
                                     int val_bytes, uint64_t n, int begin_bit, int end_bit, int *result_in_alt,
                                     void *stream)
{
    if (n && (!d_keys || !d_keys_alt || !d_vals || !d_vals_alt)) {
        set_error("gk_radix_sort_pairs32: null buffer");
        return GK_ERR_ARG;
    }
    return radix_sort_pairs32_device(d_keys, d_keys_alt, d_vals, d_vals_alt, val_bytes, n, begin_bit, end_bit,
                                     result_in_alt, as_stream(stream));
}

extern "C" int gk_partition_pairs(uint64_t *d_keys, uint64_t *d_keys_out, void *d_vals, void *d_vals_out,
                                  int val_bytes, uint64_t n, const uint64_t *d_splitters, uint32_t n_parts,
                                  uint64_t *h_counts_out, void *stream)
{
    if (!h_counts_out || (n && (!d_keys || !d_keys_out || !d_vals || !d_vals_out)) ||
        (n_parts > 1 && !d_splitters)) {
        set_error("gk_partition_pairs: null buffer");
        return GK_ERR_ARG;
    }
    // with a single destination the splitter list is empty: pass a non-null dummy to select the mode
    return partition_pairs_device(d_keys, d_keys_out, d_vals, d_vals_out, val_bytes, n,
                                  d_splitters ? d_splitters : d_keys, n_parts, h_counts_out,
                                  as_stream(stream));
}

extern "C" int gk_partition_count(const uint64_t *d_keys, uint64_t n, const uint64_t *d_splitters, uint32_t n_parts,
                                  uint64_t *h_counts_out, void *stream)
{
    if (!h_counts_out || (n && !d_keys) || (n_parts > 1 && !d_splitters)) {
        set_error("gk_partition_count: null buffer");
        return GK_ERR_ARG;
    }
    return partition_count_device(d_keys, n, d_splitters ? d_splitters : d_keys, n_parts, h_counts_out,
                                  as_stream(stream));
}

extern "C" int gk_partition_pairs_peer(const uint64_t *d_keys, const void *d_vals, int val_bytes, uint64_t n,
                                       const uint64_t *d_splitters, uint32_t n_parts,
                                       uint64_t *const *h_dst_keys, void *const *h_dst_vals,
                                       const uint64_t *h_dst_offsets, const uint64_t *h_key_base, int skip_ambiguous,
                                       int *d_err, void *stream)
{
    if (!h_dst_keys || !h_dst_vals || !h_dst_offsets || (n && (!d_keys || !d_vals)) ||
        (n_parts > 1 && !d_splitters)) {
        set_error("gk_partition_pairs_peer: null buffer");
        return GK_ERR_ARG;
    }
    return partition_pairs_peer_device(d_keys, d_vals, val_bytes, n, d_splitters, n_parts, h_dst_keys,
                                       h_dst_vals, h_dst_offsets, h_key_base, skip_ambiguous, d_err,
                                       as_stream(stream));
}

extern "C" int gk_partition_count_split(const uint64_t *d_keys, uint64_t n, const uint64_t *d_splitters,
                                        uint32_t n_parts, int class_bit, uint64_t *d_counts_out, void *stream)
{
    if (!d_counts_out || (n && !d_keys) || (n_parts > 1 && !d_splitters)) {
        set_error("gk_partition_count_split: null buffer");
        return GK_ERR_ARG;
    }
    return partition_count_split_device(d_keys, n, d_splitters, n_parts, class_bit,
                                        reinterpret_cast<unsigned long long *>(d_counts_out), as_stream(stream));
}

#ifdef GK_PROBE
// tools/probe_onesweep.py: read (and clear) the phase counters of onesweep_kernel
extern "C" int gk_probe_read(unsigned long long *h16)
{
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    if (cudaMemcpyFromSymbol(h16, gk::g_probe, sizeof(unsigned long long) * 16) != cudaSuccess) return 1;
    unsigned long long zero[16] = {};
    return cudaMemcpyToSymbol(gk::g_probe, zero, sizeof(zero)) == cudaSuccess ? 0 : 1;
}
#endif
