// gk_pack.cu -- k-mer key packing (north_star subsystem 1; SURVEY.md 8a rows A3 + A4).
//
// pack_keys_kernel: one streaming pass over the sequence byte array.  A CTA stages a tile of
// bases in shared memory with 128-bit loads, converts it once into a packed 2-bit stream plus
// per-position "not A/C/G/T" and "'$'" bit masks, and then every thread cuts its windows out
// of those streams with funnel shifts.  Output position = start - valid_len * segment, which
// is exactly the slot the reference's init loop gives that start (kmers.py:814-826), so the
// (key, start) pairs leave in ascending start order with no compaction pass.
//
// Key of a window w over its first key_len symbols (compare order = raw ASCII, kmers.py:381-388):
//   pure (all A/C/G/T):  value(w) = 2-bit code, A<C<G<T, most significant symbol first
//   otherwise:           value(w) = number of pure key_len-mers that sort below w
//                                 = prefix * 4^(key_len-j) + below(w[j]) * 4^(key_len-j-1)
//                        with j the first non-ACGT position, below() the count of A/C/G/T
//                        smaller than that byte ('$' -> 0: a terminated k-mer sorts first,
//                        kmers.py:372-375).
//   key = class_bit ? (value << 1) | is_pure : value
// A non-pure window never equals a pure one, and value() is monotone in the true order, so one
// stable integer sort places every window correctly relative to all pure windows; only runs
// of non-pure windows with equal value need the 4-bit refinement in gk_index.cu.
//
// pack4_words_kernel / rank4_stream_kernel: terminator-aware 4-bit rank words for an arbitrary list of
// starts (the refinement keys).
#include <stdlib.h>

#include "gk_common.cuh"

namespace gk {

constexpr int kPackThreads = 256;
constexpr int kPackPerThread = 16;
constexpr int kPackTile = kPackThreads * kPackPerThread;  // window starts per CTA
constexpr int kPackChunks = kPackTile / 16 + 2;           // 16-byte chunks staged (tile + 32 B halo)
constexpr int kPackHistPasses = 5;                        // digit positions counted on the fly (see pack_hist_passes_max)
constexpr int kPackKeyRow = kPackThreads + 1;             // padded row of the key staging buffer (bank-conflict free)

// ---- TMA bulk copy of the byte tile (cp.async.bulk + mbarrier, sm_90+/sm_100a) --------------------------------
// A persistent CTA knows its next tile, so one thread asks the copy engine for the next 4128 bytes while the
// CTA still works on the current tile; the bytes land in the other half of a double buffer and an mbarrier
// counts them in.  The threads then read their 16-byte chunk from shared memory instead of global memory.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_global, uint32_t bytes, uint64_t *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_global), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

template <typename IdxT, int HP>
__global__ void __launch_bounds__(kPackThreads)
pack_keys_kernel(const uint8_t *__restrict__ sba, uint64_t sba_len,
                 const uint64_t *__restrict__ seg_starts, uint32_t n_seg, uint32_t valid_len,
                 uint32_t key_len, int class_bit, uint64_t first_start, uint64_t end_start,
                 uint64_t out_base, uint64_t *__restrict__ keys_out, IdxT *__restrict__ idx_out,
                 unsigned long long *__restrict__ n_amb_out, uint64_t n_tiles, int hist_begin_bit,
                 int hist_end_bit, unsigned long long *__restrict__ g_hist /* [passes][256] or null */,
                 const FragOut frag /* frag.key == nullptr: no fragment list */, int use_bulk)
{
    __shared__ __align__(128) uint8_t s_bytes2[2][kPackChunks * 16];   // double buffer of the staged bytes
    __shared__ __align__(8) uint64_t s_mbar[2];
    __shared__ uint32_t s_codes[kPackChunks];
    __shared__ __align__(4) uint16_t s_amb[kPackChunks + 2];
    __shared__ __align__(4) uint16_t s_sep[kPackChunks + 2];
    __shared__ __align__(4) uint16_t s_ne[kPackChunks + 2];  // bit i: byte i differs from byte i + 1
    __shared__ uint32_t s_sep_pre[kPackChunks / 2 + 2];
    __shared__ uint32_t s_valid[kPackTile / 32];  // bit b of word w: a window may start at 32w + b
    __shared__ uint32_t s_seg0;
    // digit histograms of the keys this CTA emits, for the radix passes that follow (the sort would
    // otherwise read all keys once more just to count digits); flushed once per CTA
    __shared__ uint32_t s_hist[kPackHistPasses][256];
    // keys of a whole tile, written by the thread that cuts 16 CONSECUTIVE windows out of three code words and
    // read back position by position for coalesced stores: s_keys[j * kPackKeyRow + t] = key of position 16 t + j
    extern __shared__ __align__(16) unsigned char s_dyn[];
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(s_dyn);   // kPackPerThread * kPackKeyRow entries

    const uint32_t t = threadIdx.x;
    const int hist_passes = g_hist ? (hist_end_bit - hist_begin_bit + 7) / 8 : 0;
    if (g_hist)
        for (int i = t; i < kPackHistPasses * 256; i += kPackThreads) (&s_hist[0][0])[i] = 0;
    uint32_t n_amb = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(sba) & 15u) == 0;
    const uint64_t tile_first = first_start / kPackTile;
    // a tile whose staged range lies inside the array can come through the copy engine
    auto bulk_ok = [&](uint64_t tile_) -> bool {
        return use_bulk && aligned && tile_ < n_tiles &&
               (tile_first + tile_) * (uint64_t)kPackTile + (uint64_t)kPackChunks * 16 <= sba_len;
    };
    if (use_bulk) {
        if (t == 0) {
            mbar_init(&s_mbar[0], 1);
            mbar_init(&s_mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (t == 0 && bulk_ok(blockIdx.x))
            bulk_load(s_bytes2[0], sba + (tile_first + blockIdx.x) * (uint64_t)kPackTile, kPackChunks * 16, &s_mbar[0]);
    }
    uint32_t iter = 0, phase_bits = 0;   // bit b of phase_bits: parity the next wait on buffer b expects
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
    // tiles are aligned to the byte array, not to first_start, so 128-bit loads stay aligned
    const uint64_t tile0 = (tile_first + tile) * (uint64_t)kPackTile;
    uint32_t any_amb = 0;
    const uint32_t buf = iter & 1u;
    uint8_t *s_bytes = s_bytes2[buf];
    const bool from_bulk = bulk_ok(tile);
    // the other buffer was last read one iteration ago (every thread has passed that iteration's barriers)
    if (t == 0 && bulk_ok(tile + gridDim.x))
        bulk_load(s_bytes2[buf ^ 1u], sba + (tile0 + (uint64_t)gridDim.x * kPackTile), kPackChunks * 16,
                  &s_mbar[buf ^ 1u]);
    if (from_bulk) {
        mbar_wait(&s_mbar[buf], (phase_bits >> buf) & 1u);
        phase_bits ^= 1u << buf;
    }

    // ---- stage bytes, convert to streams ----------------------------------------------------
    for (uint32_t c = t; c < kPackChunks; c += kPackThreads) {
        const uint64_t g = tile0 + 16ull * c;
        uint32_t w[4];
        if (from_bulk) {
            const uint4 q = *reinterpret_cast<const uint4 *>(s_bytes + 16 * c);
            w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
        } else if (aligned && g + 16 <= sba_len) {
            uint4 q = *reinterpret_cast<const uint4 *>(sba + g);
            w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t x = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint64_t p = g + 4 * i + j;
                    uint32_t b = (p < sba_len) ? sba[p] : kSep;  // past the end == terminator
                    x |= b << (8 * j);
                }
                w[i] = x;
            }
        }
        if (!from_bulk) *reinterpret_cast<uint4 *>(s_bytes + 16 * c) = make_uint4(w[0], w[1], w[2], w[3]);
        // four bytes at a time (the kernel is bound by integer issue, not by memory)
        uint32_t codes = 0, amb = 0, sep = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t x = w[i];
            uint32_t c = (x >> 1) & 0x03030303u;              // 2-bit code of every byte (A<C<G<T) ...
            c ^= (c >> 1) & 0x01010101u;
            // ... the byte each code stands for, to tell A/C/G/T from everything else
            const uint32_t sel = (c & 0x3u) | ((c >> 4) & 0x30u) | ((c >> 8) & 0x300u) | ((c >> 12) & 0x3000u);
            const uint32_t canon = __byte_perm(0x54474341u, 0u, sel);
            const uint32_t d = x ^ canon;                     // non-zero byte: not A/C/G/T
            const uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
            const uint32_t e = x ^ 0x24242424u;               // zero byte: '$'
            const uint32_t zs = ~(((e & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | e) & 0x80808080u;
            const uint32_t nzb = nz >> 7, zsb = zs >> 7;      // one bit per byte at 0, 8, 16, 24 -> a nibble
            amb |= ((nzb | (nzb >> 7) | (nzb >> 14) | (nzb >> 21)) & 0xFu) << (4 * i);
            sep |= ((zsb | (zsb >> 7) | (zsb >> 14) | (zsb >> 21)) & 0xFu) << (4 * i);
            const uint32_t r = __byte_perm(c, 0u, 0x0123);    // first byte on top: most significant symbol first
            codes |= ((r | (r >> 6) | (r >> 12) | (r >> 18)) & 0xFFu) << (24 - 8 * i);
        }
        s_codes[c] = codes;
        s_amb[c] = (uint16_t)amb;
        s_sep[c] = (uint16_t)sep;
        any_amb |= amb;   // ('$' included: a terminated window is a non-ACGT window too)
    }
    if (t < 2) { s_amb[kPackChunks + t] = 0xFFFFu; s_sep[kPackChunks + t] = 0xFFFFu; s_ne[kPackChunks + t] = 0xFFFFu; }
    if (t == 0) s_seg0 = upper_seg(seg_starts, n_seg, tile0 < sba_len ? tile0 : sba_len - 1);
    // (barrier + vote) does the staged range hold anything but A/C/G/T?  Nearly all tiles do not.
    const bool dirty = __syncthreads_or((int)any_amb) != 0;
    const bool tile_frags = dirty && frag.key != nullptr;
    // tile-relative bounds of the requested range of starts
    const uint32_t q_lo = first_start > tile0 ? (uint32_t)((first_start - tile0 < (uint64_t)kPackTile)
                                                               ? first_start - tile0 : kPackTile) : 0u;
    const uint32_t q_hi = end_start > tile0 ? (uint32_t)((end_start - tile0 < (uint64_t)kPackTile)
                                                             ? end_start - tile0 : kPackTile) : 0u;
    if (!dirty && q_lo == 0 && q_hi == (uint32_t)kPackTile && valid_len <= 32 && key_len <= valid_len) {
        // ---- fast tile: every position starts a pure window inside one record -------------------------------
        // (no '$' and no other symbol in the tile or its halo, so nothing to validate, classify or list).
        // A thread cuts its 16 consecutive windows out of three code words in registers ...
        const uint32_t c0 = s_codes[t], c1 = s_codes[t + 1], c2 = s_codes[t + 2];
        const uint64_t hi64 = ((uint64_t)c0 << 32) | c1;
        const uint32_t sh = 64u - 2u * key_len;
#pragma unroll
        for (int j = 0; j < kPackPerThread; ++j) {
            const uint64_t x = (j == 0) ? hi64 : ((hi64 << (2 * j)) | ((uint64_t)c2 >> (32 - 2 * j)));
            const uint64_t value = x >> sh;
            const uint64_t key = class_bit ? ((value << 1) | 1ull) : value;
            s_keys[j * kPackKeyRow + t] = key;
#pragma unroll
            for (int p = 0; p < kPackHistPasses; ++p) {
                if (p < (HP >= 0 ? HP : hist_passes)) {
                    const int lo = hist_begin_bit + 8 * p;
                    const int bits = (hist_end_bit - lo < 8) ? hist_end_bit - lo : 8;
                    atomicAdd(&s_hist[p][(uint32_t)(key >> lo) & ((1u << bits) - 1u)], 1u);
                }
            }
        }
        __syncthreads();
        // ... and the tile leaves in position order: consecutive threads store consecutive pairs
        const uint64_t pos0 = tile0 - (uint64_t)valid_len * s_seg0 - out_base;  // (modular arithmetic)
#pragma unroll
        for (int j = 0; j < kPackPerThread; ++j) {
            const uint32_t q = t + j * kPackThreads;
            keys_out[pos0 + q] = s_keys[(q & 15u) * kPackKeyRow + (q >> 4)];
            idx_out[pos0 + q] = (IdxT)(tile0 + q);
        }
        __syncthreads();  // the staged streams and keys are rewritten by the next tile
        continue;
    }
    if (tile_frags) {
        // "next byte differs" bits: a window equals the window one position to its left iff its key_len + 1
        // bytes from that position on are all the same symbol
        for (uint32_t c = t; c < kPackChunks; c += kPackThreads) {
            const uint4 q = *reinterpret_cast<const uint4 *>(s_bytes + 16 * c);
            const uint32_t nxt = (c + 1 < kPackChunks) ? (uint32_t)s_bytes[16 * c + 16] : 0x100u;  // unknown: differs
            const uint32_t w[5] = {q.x, q.y, q.z, q.w, nxt};
            uint32_t ne = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t a = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
                const uint32_t b = (i == 15) ? nxt : ((w[(i + 1) >> 2] >> (8 * ((i + 1) & 3))) & 0xFFu);
                ne |= (a != b ? 1u : 0u) << i;
            }
            s_ne[c] = (uint16_t)ne;
        }
    }

    // exclusive prefix of '$' counts per 32-position word (one warp, kPackChunks/2 words)
    const uint32_t *amb32 = reinterpret_cast<const uint32_t *>(s_amb);
    const uint32_t *sep32 = reinterpret_cast<const uint32_t *>(s_sep);
    // a window of valid_len <= 32 symbols may start at position b iff no '$' (or array end, staged as '$')
    // lies in [b, b + valid_len): smear the '$' bits to the right by valid_len - 1 with doubling shifts
    const bool fast_valid = valid_len <= 32;
    if (fast_valid && t >= 32 && t < 32 + kPackTile / 32) {
        const uint32_t w = t - 32;
        uint64_t y = ((uint64_t)sep32[w + 1] << 32) | sep32[w];
        for (uint32_t done = 1; done < valid_len;) {
            const uint32_t d = (done < valid_len - done) ? done : valid_len - done;
            y |= y >> d;
            done += d;
        }
        s_valid[w] = ~(uint32_t)y;
    }
    if (t < 32) {
        constexpr int n_words = kPackChunks / 2;          // 129
        constexpr int per_lane = (n_words + 31) / 32;     // 5
        uint32_t local[per_lane];
        uint32_t sum = 0;
#pragma unroll
        for (int i = 0; i < per_lane; ++i) {
            int wi = t * per_lane + i;
            local[i] = sum;
            sum += (wi < n_words) ? __popc(sep32[wi]) : 0;
        }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (t >= (uint32_t)o) inc += v;
        }
        uint32_t excl = inc - sum;
#pragma unroll
        for (int i = 0; i < per_lane; ++i) {
            int wi = t * per_lane + i;
            if (wi < n_words) s_sep_pre[wi] = excl + local[i];
        }
    }
    __syncthreads();

    // ---- cut windows ---------------------------------------------------------------------------
    const uint32_t seg0 = s_seg0;
    const uint64_t key_mask = (key_len >= 32) ? 0xFFFFFFFFull : ((1ull << key_len) - 1ull);
    // the output slot of tile position 0
    const uint64_t tile_pos0 = tile0 - (uint64_t)valid_len * seg0 - out_base;  // (modular arithmetic)
#pragma unroll 4
    for (int j = 0; j < kPackPerThread; ++j) {
        const uint32_t q = t + j * kPackThreads;
        if (q < q_lo || q >= q_hi) continue;
        const uint64_t i = tile0 + q;
        const uint32_t mw = q >> 5, mo = q & 31u;
        const uint32_t seg_rel = s_sep_pre[mw] + __popc(sep32[mw] & ((1u << mo) - 1u));
        if (fast_valid) {
            if (!((s_valid[mw] >> mo) & 1u)) continue;
        } else {
            const uint32_t seg = seg0 + seg_rel;
            if (seg >= n_seg) continue;
            const uint64_t seg_end = (seg + 1 < n_seg) ? seg_starts[seg + 1] - 1 : sba_len;
            if (i + valid_len > seg_end) continue;  // also rejects '$' positions themselves
        }

        const uint32_t cw = q >> 4, co = 2u * (q & 15u);
        const uint32_t c0 = s_codes[cw], c1 = s_codes[cw + 1], c2 = s_codes[cw + 2];
        const uint32_t hi = __funnelshift_l(c1, c0, co);
        const uint32_t lo = __funnelshift_l(c2, c1, co);
        uint64_t value = (((uint64_t)hi << 32) | lo) >> (64 - 2 * key_len);

        const uint64_t m64 = (((uint64_t)amb32[mw + 1] << 32) | amb32[mw]) >> mo;
        const uint32_t m = (uint32_t)(m64 & key_mask);
        const bool pure = (m == 0);
        if (!pure) {
            ++n_amb;
            const uint32_t j0 = __ffs(m) - 1;                  // first non-ACGT symbol
            const uint32_t below = acgt_below(s_bytes[q + j0]);
            const uint32_t rem = 2u * (key_len - j0);          // bits from symbol j0 to the end
            const uint64_t prefix = (rem >= 64) ? 0ull : (value >> rem);
            value = ((rem >= 64) ? 0ull : (prefix << rem)) + ((uint64_t)below << (rem - 2));
        }
        const uint64_t key = class_bit ? ((value << 1) | (pure ? 1ull : 0ull)) : value;
        const uint64_t pos = tile_pos0 + q - (uint64_t)valid_len * seg_rel;
        keys_out[pos] = key;
        idx_out[pos] = (IdxT)i;
        if (!pure && tile_frags) {
            // head of a block of identical windows?  (the window at q - 1 must be one this launch emits too)
            const uint32_t *ne32 = reinterpret_cast<const uint32_t *>(s_ne);
            bool cont = false;
            if (q > q_lo) {
                const uint32_t u = q - 1;
                const uint64_t ne64 = (((uint64_t)ne32[(u >> 5) + 1] << 32) | ne32[u >> 5]) >> (u & 31u);
                cont = (ne64 & key_mask) == 0;
            }
            if (!cont) {
                // the block runs until the first differing byte pair at p*: windows q .. p* - key_len + 1
                uint32_t word = q >> 5;
                uint32_t bits = ne32[word] & ~((1u << (q & 31u)) - 1u);
                while (bits == 0) bits = ne32[++word];   // (the pad words are all ones)
                const int p_star = (int)(word * 32u + (uint32_t)__ffs((int)bits) - 1u);
                int cnt = p_star - (int)q - (int)key_len + 2;
                if (cnt < 1) cnt = 1;
                if (cnt > (int)(q_hi - q)) cnt = (int)(q_hi - q);
                if (valid_len > key_len) {  // identical key prefixes, but the record may end first
                    const uint32_t seg = seg0 + seg_rel;
                    const uint64_t seg_end = (seg + 1 < n_seg) ? seg_starts[seg + 1] - 1 : sba_len;
                    const uint64_t left = seg_end - valid_len - i + 1;   // valid starts from i on
                    if ((uint64_t)cnt > left) cnt = (int)left;
                }
                uint64_t w0 = 0, w1 = 0;
                for (uint32_t j = 0; j < key_len; ++j) {
                    const uint64_t code = rank4(s_bytes[q + j]);
                    if (code == 0) break;   // '$' / end of array: the k-mer ends here (kmers.py:360-378)
                    if (j < 16) w0 |= code << (4u * (15u - j));
                    else w1 |= code << (4u * (31u - j));
                }
                const unsigned long long slot = atomicAdd(frag.counter, 1ull);
                if (slot < frag.capacity) {
                    frag.key[slot] = key;
                    frag.w0[slot] = w0;
                    frag.w1[slot] = w1;
                    frag.start[slot] = i;
                    frag.count[slot] = (uint32_t)cnt;
                }
            }
        }
        // HP >= 0: the number of counted digit positions is a compile-time constant (no branches)
#pragma unroll
        for (int p = 0; p < kPackHistPasses; ++p) {
            if (p < (HP >= 0 ? HP : hist_passes)) {
                const int lo = hist_begin_bit + 8 * p;
                const int bits = (hist_end_bit - lo < 8) ? hist_end_bit - lo : 8;
                atomicAdd(&s_hist[p][(uint32_t)(key >> lo) & ((1u << bits) - 1u)], 1u);
            }
        }
    }
    __syncthreads();  // the staged streams are rewritten by the next tile
    }
    if (g_hist) {
        __syncthreads();
        for (int i = t; i < hist_passes * 256; i += kPackThreads) {
            const uint32_t c = (&s_hist[0][0])[i];
            if (c) atomicAdd(&g_hist[i], (unsigned long long)c);
        }
    }
    if (n_amb_out) {
        n_amb = warp_sum(n_amb);
        if (lane_id() == 0 && n_amb) atomicAdd(n_amb_out, (unsigned long long)n_amb);
    }
}

// Both 4-bit rank words (16 symbols per word, most significant first) of a window of max_len <= 32 symbols
// in one read of its bytes; symbols at or after a '$' / the end of the array are 0, so a shorter k-mer sorts
// first (kmers.py:360-378).  Pure windows (class bit set in class_keys) get zeros: their radix key already
// says everything.
template <typename IdxT>
__global__ void __launch_bounds__(256)
pack4_words_kernel(const uint8_t *__restrict__ sba, uint64_t sba_len, const IdxT *__restrict__ idx, uint64_t n,
                   uint32_t max_len, const uint64_t *__restrict__ class_keys, uint64_t *__restrict__ w0_out,
                   uint64_t *__restrict__ w1_out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        uint64_t w0 = 0, w1 = 0;
        if (!(class_keys && (class_keys[r] & 1ull))) {
            const uint64_t s = (uint64_t)idx[r];
            for (uint32_t j = 0; j < max_len; ++j) {
                const uint64_t p = s + j;
                const uint64_t code = (p < sba_len) ? rank4(sba[p]) : 0u;
                if (code == 0) break;  // '$' / end of array: the k-mer ends here and sorts first
                if (j < 16) w0 |= code << (4u * (15u - j));
                else w1 |= code << (4u * (31u - j));
            }
        }
        w0_out[r] = w0;
        if (w1_out) w1_out[r] = w1;
    }
}

int pack4_words_device(const uint8_t *d_sba, uint64_t sba_len, const void *d_idx, int idx_bytes, uint64_t n,
                       uint32_t max_len, const uint64_t *d_class_keys, uint64_t *d_w0, uint64_t *d_w1,
                       cudaStream_t st)
{
    if (n == 0) return GK_OK;
    if (max_len > 32) {
        set_error("pack4_words: at most 32 symbols");
        return GK_ERR_ARG;
    }
    uint64_t blocks = (n + 255) / 256;
    uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (idx_bytes == 4)
        pack4_words_kernel<uint32_t><<<(unsigned)blocks, 256, 0, st>>>(
            d_sba, sba_len, (const uint32_t *)d_idx, n, max_len, d_class_keys, d_w0, d_w1);
    else
        pack4_words_kernel<uint64_t><<<(unsigned)blocks, 256, 0, st>>>(
            d_sba, sba_len, (const uint64_t *)d_idx, n, max_len, d_class_keys, d_w0, d_w1);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

// ---- 4-bit rank stream of the whole byte array ------------------------------------------------------------
// word j holds the rank4 codes of symbols 16j .. 16j+15, symbol 16j in the top nibble.  With it, the two
// refinement words of a window are three aligned 64-bit loads and two funnel shifts instead of a byte loop.
__global__ void __launch_bounds__(256)
rank4_stream_kernel(const uint8_t *__restrict__ sba, uint64_t sba_len, uint64_t n_words,
                    uint64_t *__restrict__ stream)
{
    const bool aligned = (reinterpret_cast<uintptr_t>(sba) & 15u) == 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_words; j += stride) {
        const uint64_t g = 16 * j;
        uint64_t w = 0;
        if (aligned && g + 16 <= sba_len) {
            const uint4 q = *reinterpret_cast<const uint4 *>(sba + g);
            const uint32_t v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 16; ++i)
                w |= (uint64_t)rank4((v[i >> 2] >> (8 * (i & 3))) & 0xFFu) << (4 * (15 - i));
        } else {
            for (int i = 0; i < 16; ++i)
                if (g + i < sba_len) w |= (uint64_t)rank4(sba[g + i]) << (4 * (15 - i));
        }
        stream[j] = w;
    }
}

// Both refinement words of windows of max_len <= 32 symbols that lie inside one record (no '$').
template <typename IdxT>
__global__ void __launch_bounds__(256)
pack4_words_stream_kernel(const uint64_t *__restrict__ stream, const IdxT *__restrict__ idx, uint64_t n,
                          uint32_t max_len, const uint64_t *__restrict__ class_keys,
                          uint64_t *__restrict__ w0_out, uint64_t *__restrict__ w1_out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t n0 = max_len < 16 ? max_len : 16, n1 = max_len > 16 ? max_len - 16 : 0;
    const uint64_t m0 = n0 == 16 ? ~0ull : ~(~0ull >> (4 * n0));
    const uint64_t m1 = n1 == 0 ? 0ull : (n1 == 16 ? ~0ull : ~(~0ull >> (4 * n1)));
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        uint64_t w0 = 0, w1 = 0;
        if (!(class_keys && (class_keys[r] & 1ull))) {
            const uint64_t s = (uint64_t)idx[r];
            const uint64_t a = s >> 4;
            const uint32_t off = 4u * (uint32_t)(s & 15u);
            const uint64_t x0 = stream[a], x1 = stream[a + 1], x2 = stream[a + 2];
            w0 = (off ? (x0 << off) | (x1 >> (64 - off)) : x0) & m0;
            w1 = (off ? (x1 << off) | (x2 >> (64 - off)) : x1) & m1;
        }
        w0_out[r] = w0;
        if (w1_out) w1_out[r] = w1;
    }
}

int rank4_stream_device(const uint8_t *d_sba, uint64_t sba_len, uint64_t *d_stream, cudaStream_t st)
{
    const uint64_t n_words = sba_len / 16 + 3;  // two words of zero padding behind the last symbol
    uint64_t blocks = (n_words + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    rank4_stream_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_sba, sba_len, n_words, d_stream);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int pack4_words_stream_device(const uint64_t *d_stream, const void *d_idx, int idx_bytes, uint64_t n,
                              uint32_t max_len, const uint64_t *d_class_keys, uint64_t *d_w0, uint64_t *d_w1,
                              cudaStream_t st)
{
    if (n == 0) return GK_OK;
    if (max_len > 32) {
        set_error("pack4_words: at most 32 symbols");
        return GK_ERR_ARG;
    }
    uint64_t blocks = (n + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (idx_bytes == 4)
        pack4_words_stream_kernel<uint32_t><<<(unsigned)blocks, 256, 0, st>>>(
            d_stream, (const uint32_t *)d_idx, n, max_len, d_class_keys, d_w0, d_w1);
    else
        pack4_words_stream_kernel<uint64_t><<<(unsigned)blocks, 256, 0, st>>>(
            d_stream, (const uint64_t *)d_idx, n, max_len, d_class_keys, d_w0, d_w1);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int pack_hist_passes_max() { return kPackHistPasses; }

int pack_keys_device(const uint8_t *d_sba, uint64_t sba_len, const uint64_t *d_seg_starts,
                     uint32_t n_seg, uint32_t valid_len, uint32_t key_len, int class_bit,
                     uint64_t first_start, uint64_t end_start, uint64_t out_base,
                     uint64_t *d_keys_out, int idx_bytes, void *d_idx_out,
                     unsigned long long *d_n_amb, int hist_begin_bit, int hist_end_bit,
                     unsigned long long *d_hist, cudaStream_t st, const FragOut *frag)
{
    // valid_len < key_len is the variable-length mode: windows shorter than the key end at their record's
    // '$', which the key treats like any other non-ACGT symbol (it sorts below A); needs the class bit
    if (key_len < 1 || key_len > 32 || valid_len < 1 || (valid_len < key_len && !class_bit) ||
        (class_bit && key_len > 31)) {
        set_error("pack_keys: key_len %u / valid_len %u / class_bit %d out of range", key_len,
                  valid_len, class_bit);
        return GK_ERR_ARG;
    }
    if (end_start > sba_len) end_start = sba_len;
    if (first_start >= end_start) return GK_OK;
    const uint64_t tile_first = first_start / kPackTile;
    const uint64_t tile_last = (end_start - 1) / kPackTile;
    const uint64_t n_tiles = tile_last - tile_first + 1;
    if (d_hist && (hist_end_bit - hist_begin_bit + 7) / 8 > kPackHistPasses) {
        set_error("pack_keys: at most %d digit positions can be counted", kPackHistPasses);
        return GK_ERR_ARG;
    }
    // persistent CTAs (4 fit an SM: 64 registers x 256 threads, 45 KB of shared memory each) so that the shared
    // histograms are flushed ~600 times, not per tile
    uint64_t grid = (uint64_t)sm_count() * 4;
    if (grid > n_tiles) grid = n_tiles;
    const int hp = d_hist ? (hist_end_bit - hist_begin_bit + 7) / 8 : 0;
    const FragOut fr = frag ? *frag : FragOut();
    constexpr int kDynSmem = kPackPerThread * kPackKeyRow * 8;
    const char *tma = getenv("GK_PACK_TMA");
    const int use_bulk = (tma && tma[0] == '0') ? 0 : 1;
#define GK_PACK_LAUNCH(IDX, HPV)                                                                       \
    do {                                                                                               \
        GK_CUDA(cudaFuncSetAttribute(pack_keys_kernel<IDX, HPV>,                                       \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));          \
        pack_keys_kernel<IDX, HPV><<<(unsigned)grid, kPackThreads, kDynSmem, st>>>(                    \
            d_sba, sba_len, d_seg_starts, n_seg, valid_len, key_len, class_bit, first_start,           \
            end_start, out_base, d_keys_out, (IDX *)d_idx_out, d_n_amb, n_tiles, hist_begin_bit,       \
            hist_end_bit, d_hist, fr, use_bulk);                                                       \
    } while (0)
    if (idx_bytes == 4) {
        if (hp == 0) GK_PACK_LAUNCH(uint32_t, 0);
        else if (hp == 4) GK_PACK_LAUNCH(uint32_t, 4);
        else if (hp == 5) GK_PACK_LAUNCH(uint32_t, 5);
        else GK_PACK_LAUNCH(uint32_t, -1);
    } else {
        if (hp == 0) GK_PACK_LAUNCH(uint64_t, 0);
        else if (hp == 4) GK_PACK_LAUNCH(uint64_t, 4);
        else if (hp == 5) GK_PACK_LAUNCH(uint64_t, 5);
        else GK_PACK_LAUNCH(uint64_t, -1);
    }
#undef GK_PACK_LAUNCH
    GK_LAUNCH_CHECK();
    return GK_OK;
}

// ---- keys for an arbitrary list of starts -------------------------------------------------------------------
// Kmers.kmer_sba_start_indices may be assigned by the caller and sort() orders whatever the array holds
// (kmers.py:1648).  Same key definition as pack_keys_kernel, one thread per start, bytes read in place; a
// window that meets its record's '$' (or the end of the array) inside the key is a non-ACGT window whose
// first non-ACGT symbol sorts below 'A'.
template <typename IdxT>
__global__ void __launch_bounds__(256)
pack_keys_list_kernel(const uint8_t *__restrict__ sba, uint64_t sba_len, const IdxT *__restrict__ idx, uint64_t n,
                      uint32_t key_len, int class_bit, uint64_t *__restrict__ keys_out,
                      unsigned long long *__restrict__ n_amb_out)
{
    uint32_t n_amb = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        const uint64_t s = (uint64_t)idx[r];
        uint64_t value = 0;
        bool pure = true;
        for (uint32_t j = 0; j < key_len; ++j) {
            const uint32_t b = (s + j < sba_len) ? sba[s + j] : kSep;
            if (is_acgt(b)) {
                value = (value << 2) | code2(b);
            } else {
                const uint32_t rem = 2u * (key_len - j);
                value = ((rem >= 64) ? 0ull : (value << rem)) + ((uint64_t)acgt_below(b) << (rem - 2));
                pure = false;
                break;
            }
        }
        if (!pure) ++n_amb;
        keys_out[r] = class_bit ? ((value << 1) | (pure ? 1ull : 0ull)) : value;
    }
    if (n_amb_out) {
        n_amb = warp_sum(n_amb);
        if (lane_id() == 0 && n_amb) atomicAdd(n_amb_out, (unsigned long long)n_amb);
    }
}

template <typename IdxT>
__global__ void __launch_bounds__(256)
widen_indices_kernel(const IdxT *__restrict__ idx, uint64_t n, uint64_t *__restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) out[r] = (uint64_t)idx[r];
}

// report[0] += starts that do not begin a k-mer of min_len symbols inside one record,
// report[1] += positions r > 0 with idx[r] <= idx[r - 1] (the list is not strictly ascending)
template <typename IdxT>
__global__ void __launch_bounds__(256)
check_starts_kernel(const IdxT *__restrict__ idx, uint64_t n, const uint64_t *__restrict__ seg_starts,
                    uint32_t n_seg, uint64_t sba_len, uint32_t min_len, unsigned long long *__restrict__ report)
{
    uint32_t bad = 0, unordered = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        const uint64_t s = (uint64_t)idx[r];
        bool ok = s < sba_len;
        if (ok) {
            const uint32_t seg = upper_seg(seg_starts, n_seg, s);
            const uint64_t seg_end = (seg + 1 < n_seg) ? seg_starts[seg + 1] - 1 : sba_len;  // position of '$'
            ok = s + min_len <= seg_end;
        }
        if (!ok) ++bad;
        if (r > 0 && (uint64_t)idx[r - 1] >= s) ++unordered;
    }
    bad = warp_sum(bad);
    unordered = warp_sum(unordered);
    if (lane_id() == 0) {
        if (bad) atomicAdd(&report[0], (unsigned long long)bad);
        if (unordered) atomicAdd(&report[1], (unsigned long long)unordered);
    }
}

// Keys of evenly spaced windows of a slice (the multi-GPU driver picks its splitters from them): sample j is
// window number base + j * n_slice / n_samples of the index; its start comes from the same slot rule as
// init_indices_kernel, its key from the same definition as pack_keys_kernel.
__global__ void __launch_bounds__(256)
sample_keys_kernel(const uint8_t *__restrict__ sba, uint64_t sba_len, const uint64_t *__restrict__ seg_starts,
                   uint32_t n_seg, uint32_t k, uint32_t key_len, int class_bit, uint64_t base, uint64_t n_slice,
                   uint32_t n_samples, uint64_t *__restrict__ keys_out)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_samples) return;
    const uint64_t p = base + (uint64_t)(((unsigned __int128)j * n_slice) / n_samples);
    uint32_t lo = 0, hi = n_seg;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (p < seg_starts[mid] - (uint64_t)mid * k) hi = mid; else lo = mid + 1;
    }
    const uint64_t s = p + (uint64_t)(lo - 1) * k;
    uint64_t value = 0;
    bool pure = true;
    for (uint32_t i = 0; i < key_len; ++i) {
        const uint32_t b = (s + i < sba_len) ? sba[s + i] : kSep;
        if (is_acgt(b)) {
            value = (value << 2) | code2(b);
        } else {
            const uint32_t rem = 2u * (key_len - i);
            value = ((rem >= 64) ? 0ull : (value << rem)) + ((uint64_t)acgt_below(b) << (rem - 2));
            pure = false;
            break;
        }
    }
    keys_out[j] = class_bit ? ((value << 1) | (pure ? 1ull : 0ull)) : value;
}

static unsigned list_grid(uint64_t n)
{
    uint64_t blocks = (n + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

int pack_keys_list_device(const uint8_t *d_sba, uint64_t sba_len, const void *d_idx, int idx_bytes, uint64_t n,
                          uint32_t key_len, int class_bit, uint64_t *d_keys_out, unsigned long long *d_n_amb,
                          cudaStream_t st)
{
    if (n == 0) return GK_OK;
    if (key_len < 1 || key_len > 32 || (class_bit && key_len > 31)) {
        set_error("pack_keys_list: key_len %u / class_bit %d out of range", key_len, class_bit);
        return GK_ERR_ARG;
    }
    if (idx_bytes == 4)
        pack_keys_list_kernel<uint32_t><<<list_grid(n), 256, 0, st>>>(d_sba, sba_len, (const uint32_t *)d_idx, n,
                                                                      key_len, class_bit, d_keys_out, d_n_amb);
    else
        pack_keys_list_kernel<uint64_t><<<list_grid(n), 256, 0, st>>>(d_sba, sba_len, (const uint64_t *)d_idx, n,
                                                                      key_len, class_bit, d_keys_out, d_n_amb);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int widen_indices_device(const void *d_idx, int idx_bytes, uint64_t n, uint64_t *d_out, cudaStream_t st)
{
    if (n == 0) return GK_OK;
    if (idx_bytes == 4)
        widen_indices_kernel<uint32_t><<<list_grid(n), 256, 0, st>>>((const uint32_t *)d_idx, n, d_out);
    else
        widen_indices_kernel<uint64_t><<<list_grid(n), 256, 0, st>>>((const uint64_t *)d_idx, n, d_out);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

int check_starts_device(const void *d_idx, int idx_bytes, uint64_t n, const uint64_t *d_seg_starts, uint32_t n_seg,
                        uint64_t sba_len, uint32_t min_len, unsigned long long *d_report, cudaStream_t st)
{
    if (n == 0) return GK_OK;
    if (idx_bytes == 4)
        check_starts_kernel<uint32_t><<<list_grid(n), 256, 0, st>>>((const uint32_t *)d_idx, n, d_seg_starts, n_seg,
                                                                    sba_len, min_len, d_report);
    else
        check_starts_kernel<uint64_t><<<list_grid(n), 256, 0, st>>>((const uint64_t *)d_idx, n, d_seg_starts, n_seg,
                                                                    sba_len, min_len, d_report);
    GK_LAUNCH_CHECK();
    return GK_OK;
}

// number of valid_len-windows that start before `start` (host; mirrors the kernel's slot rule)
uint64_t windows_before(const uint64_t *h_seg_starts, uint32_t n_seg, uint64_t sba_len,
                        uint32_t valid_len, uint64_t start)
{
    uint64_t n = 0;
    for (uint32_t s = 0; s < n_seg; ++s) {
        uint64_t a = h_seg_starts[s];
        uint64_t e_excl = (s + 1 < n_seg) ? h_seg_starts[s + 1] - 1 : sba_len;
        if (start <= a) break;
        uint64_t last_excl = e_excl - valid_len + 1;  // one past the last valid start
        uint64_t upto = start < last_excl ? start : last_excl;
        if (upto > a) n += upto - a;
    }
    return n;
}

}  // namespace gk

using namespace gk;

extern "C" int gk_pack_keys(const uint8_t *d_sba, uint64_t sba_len, const uint64_t *h_seg_starts,
                            uint32_t n_seg, uint32_t valid_len, uint32_t key_len, int class_bit,
                            uint64_t first_start, uint64_t end_start, uint64_t *d_keys_out,
                            int idx_bytes, void *d_idx_out, uint64_t out_capacity,
                            uint64_t *h_n_out, uint64_t *h_n_ambiguous, void *stream)
{
    if (!d_sba || !h_seg_starts || !d_keys_out || !d_idx_out || n_seg == 0 ||
        (idx_bytes != 4 && idx_bytes != 8)) {
        set_error("gk_pack_keys: bad argument");
        return GK_ERR_ARG;
    }
    cudaStream_t st = as_stream(stream);
    if (end_start > sba_len) end_start = sba_len;
    const uint64_t base = windows_before(h_seg_starts, n_seg, sba_len, valid_len, first_start);
    const uint64_t upto = windows_before(h_seg_starts, n_seg, sba_len, valid_len, end_start);
    const uint64_t n = upto - base;
    if (n > out_capacity) {
        set_error("gk_pack_keys: %llu windows do not fit the output capacity %llu",
                  (unsigned long long)n, (unsigned long long)out_capacity);
        return GK_ERR_ARG;
    }
    DeviceBuffer segs, amb;
    GK_TRY(segs.alloc((size_t)n_seg * 8, st));
    GK_CUDA(cudaMemcpyAsync(segs.ptr, h_seg_starts, (size_t)n_seg * 8, cudaMemcpyHostToDevice, st));
    GK_TRY(amb.alloc(8, st));
    GK_CUDA(cudaMemsetAsync(amb.ptr, 0, 8, st));
    GK_TRY(pack_keys_device(d_sba, sba_len, segs.as<uint64_t>(), n_seg, valid_len, key_len,
                            class_bit, first_start, end_start, base, d_keys_out, idx_bytes,
                            d_idx_out, amb.as<unsigned long long>(), 0, 0, nullptr, st, nullptr));
    uint64_t n_amb = 0;
    GK_CUDA(cudaMemcpyAsync(&n_amb, amb.ptr, 8, cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    if (h_n_out) *h_n_out = n;
    if (h_n_ambiguous) *h_n_ambiguous = n_amb;
    return GK_OK;
}

/* Multi-GPU producer (no synchronise): pack the windows whose start lies in [first_start, end_start) and list
 * the ambiguous-window fragments of that slice.  d_frag: frag_capacity * 36 bytes laid out as
 * key[cap] w0[cap] w1[cap] start[cap] (u64) then count[cap] (u32); d_counters: 4 x u64, zeroed here:
 * [0] ambiguous windows, [2] fragments found (may exceed the capacity: the list is then incomplete). */
extern "C" int gk_pack_slice(const uint8_t *d_sba, uint64_t sba_len, const uint64_t *h_seg_starts, uint32_t n_seg,
                             uint32_t kmer_len, int class_bit, uint64_t first_start, uint64_t end_start,
                             uint64_t *d_keys_out, int idx_bytes, void *d_idx_out, uint64_t out_capacity,
                             uint64_t *h_n_out, void *d_frag, uint64_t frag_capacity, uint64_t *d_counters,
                             void *stream)
{
    if (!d_sba || !h_seg_starts || !d_keys_out || !d_idx_out || !d_counters || n_seg == 0 ||
        (idx_bytes != 4 && idx_bytes != 8)) {
        set_error("gk_pack_slice: bad argument");
        return GK_ERR_ARG;
    }
    cudaStream_t st = as_stream(stream);
    if (end_start > sba_len) end_start = sba_len;
    const uint64_t base = windows_before(h_seg_starts, n_seg, sba_len, kmer_len, first_start);
    const uint64_t upto = windows_before(h_seg_starts, n_seg, sba_len, kmer_len, end_start);
    const uint64_t n = upto - base;
    if (n > out_capacity) {
        set_error("gk_pack_slice: %llu windows do not fit the output capacity %llu", (unsigned long long)n,
                  (unsigned long long)out_capacity);
        return GK_ERR_ARG;
    }
    DeviceBuffer segs;
    GK_TRY(segs.alloc((size_t)n_seg * 8, st));
    GK_CUDA(cudaMemcpyAsync(segs.ptr, h_seg_starts, (size_t)n_seg * 8, cudaMemcpyHostToDevice, st));
    GK_CUDA(cudaMemsetAsync(d_counters, 0, 32, st));
    FragOut frag;
    if (d_frag && frag_capacity && class_bit) {
        uint64_t *b = reinterpret_cast<uint64_t *>(d_frag);
        frag.key = b; frag.w0 = b + frag_capacity; frag.w1 = b + 2 * frag_capacity; frag.start = b + 3 * frag_capacity;
        frag.count = reinterpret_cast<uint32_t *>(b + 4 * frag_capacity);
        frag.counter = reinterpret_cast<unsigned long long *>(d_counters) + 2;
        frag.capacity = frag_capacity;
    }
    // k-mers longer than one key word: the key covers the first 31 symbols, the windows are still kmer_len long
    const uint32_t key_len = (class_bit && kmer_len > 31) ? 31u : kmer_len;
    GK_TRY(pack_keys_device(d_sba, sba_len, segs.as<uint64_t>(), n_seg, kmer_len, key_len, class_bit, first_start,
                            end_start, base, d_keys_out, idx_bytes, d_idx_out,
                            reinterpret_cast<unsigned long long *>(d_counters), 0, 0, nullptr, st,
                            frag.key ? &frag : nullptr));
    if (h_n_out) *h_n_out = n;
    return GK_OK;   // (the pageable segment table was staged by the runtime before cudaMemcpyAsync returned)
}

/* Keys of n_samples evenly spaced windows of the slice [first_start, end_start) (fewer when the slice holds
 * fewer windows; *h_n_out tells).  No synchronise. */
extern "C" int gk_sample_keys(const uint8_t *d_sba, uint64_t sba_len, const uint64_t *h_seg_starts, uint32_t n_seg,
                              uint32_t kmer_len, int class_bit, uint64_t first_start, uint64_t end_start,
                              uint32_t n_samples, uint64_t *d_keys_out, uint32_t *h_n_out, void *stream)
{
    if (!d_sba || !h_seg_starts || !d_keys_out || n_seg == 0 || kmer_len < 1 || (!class_bit && kmer_len > 32)) {
        set_error("gk_sample_keys: bad argument");
        return GK_ERR_ARG;
    }
    const uint32_t key_len = (class_bit && kmer_len > 31) ? 31u : kmer_len;   // as gk_pack_slice
    cudaStream_t st = as_stream(stream);
    if (end_start > sba_len) end_start = sba_len;
    const uint64_t base = windows_before(h_seg_starts, n_seg, sba_len, kmer_len, first_start);
    const uint64_t upto = windows_before(h_seg_starts, n_seg, sba_len, kmer_len, end_start);
    const uint64_t n_slice = upto - base;
    if ((uint64_t)n_samples > n_slice) n_samples = (uint32_t)n_slice;
    if (h_n_out) *h_n_out = n_samples;
    if (n_samples == 0) return GK_OK;
    DeviceBuffer segs;
    GK_TRY(segs.alloc((size_t)n_seg * 8, st));
    GK_CUDA(cudaMemcpyAsync(segs.ptr, h_seg_starts, (size_t)n_seg * 8, cudaMemcpyHostToDevice, st));
    sample_keys_kernel<<<(n_samples + 255) / 256, 256, 0, st>>>(d_sba, sba_len, segs.as<uint64_t>(), n_seg, kmer_len,
                                                                key_len, class_bit, base, n_slice, n_samples,
                                                                d_keys_out);
    GK_LAUNCH_CHECK();
    return GK_OK;
}
