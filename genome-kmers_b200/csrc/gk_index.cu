// gk_index.cu -- the device-resident k-mer index: the state behind the reference's `Kmers`
// object (kmers.py:651-760) and the two seams the reference hands to numba:
//   gk_index_sort          <- Kmers.sort()                    kmers.py:1624-1652
//   gk_index_group_counts  <- get_kmer_group_size_hist()       kmers.py:454-520 (:1072, :1166)
// Host-side orchestration only; the kernels live in gk_pack.cu / gk_sort.cu / gk_group.cu.
#include <stdlib.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "gk_common.cuh"

namespace gk {

void trace_report(const char *what);
void trace_point(const char *name);
void set_deferred_sort_error_word(int *d_err);
// kernels / device drivers from the other translation units
int kmer_count_host(const uint64_t *, uint32_t, uint64_t, uint32_t, uint64_t *);
int init_indices_device(const uint64_t *, uint32_t, uint32_t, uint64_t, int, void *, cudaStream_t);
int pack_keys_device(const uint8_t *, uint64_t, const uint64_t *, uint32_t, uint32_t, uint32_t, int,
                     uint64_t, uint64_t, uint64_t, uint64_t *, int, void *, unsigned long long *, int, int,
                     unsigned long long *, cudaStream_t, const FragOut *);
int pack_hist_passes_max();
int pack_keys_list_device(const uint8_t *, uint64_t, const void *, int, uint64_t, uint32_t, int, uint64_t *,
                          unsigned long long *, cudaStream_t);
int widen_indices_device(const void *, int, uint64_t, uint64_t *, cudaStream_t);
int check_starts_device(const void *, int, uint64_t, const uint64_t *, uint32_t, uint64_t, uint32_t,
                        unsigned long long *, cudaStream_t);
int scan_alphabet_async(const uint8_t *, uint64_t, unsigned long long *, cudaStream_t);
int verify_order_device(const uint8_t *, uint64_t, const void *, int, uint64_t, uint32_t, const uint64_t *, uint32_t,
                        uint32_t, const uint8_t *, uint64_t *, uint32_t *, cudaStream_t);
struct FragSorted;
int frag_sort_device(const FragOut &, uint64_t, uint32_t, int, int, FragSorted &, cudaStream_t);
int frag_expand_device(FragSorted &, const uint64_t *, uint64_t, int, void *, uint8_t *, const unsigned int *,
                       const unsigned long long *, uint64_t, int *, cudaStream_t);
int frag_filter_device(const void *, const unsigned long long *, uint32_t, uint64_t, uint64_t, uint64_t,
                       const FragOut &, cudaStream_t);
int frag_sort_copy_device(const FragOut &, uint64_t, uint32_t, int, const FragOut &, cudaStream_t);
int frag_merge_device(const void *, const unsigned long long *, uint32_t, uint64_t, uint64_t, uint64_t, uint64_t,
                      unsigned long long *, FragSorted &, FragOut *, cudaStream_t);
int frag_placeholders_device(const FragOut &, unsigned long long *, uint64_t, uint64_t, uint64_t *, void *, int, int *,
                             cudaStream_t);
int subset_rep_flags_device(const uint64_t *, const uint64_t *, const uint64_t *, uint64_t, uint8_t *,
                            cudaStream_t);
int gather2_u64_device(const uint64_t *, const void *, const void *, uint64_t, int, uint64_t *, cudaStream_t);
int iota_device(void *, uint64_t, int, cudaStream_t);
int block_offsets_device(const void *, uint64_t, uint64_t, const void *, int, unsigned long long *, cudaStream_t);
int subset_expand_device(const void *, const void *, const uint64_t *, const uint64_t *, const uint64_t *,
                         const void *, const void *, const unsigned long long *, uint64_t, uint64_t, int, int,
                         void *, uint8_t *, cudaStream_t);
int radix_sort_pairs_device(uint64_t *, uint64_t *, void *, void *, int, uint64_t, int, int, int *,
                            cudaStream_t, SortTiming *, const unsigned long long *d_pre_hist = nullptr,
                            void *d_vals_final = nullptr);
int key_flags_device(const uint64_t *, uint64_t, int, uint8_t *, cudaStream_t);
int tie_fix_flags_device(uint64_t *, void *, int, uint64_t, int, int, uint8_t *, unsigned int *,
                         unsigned long long *, cudaStream_t, unsigned long long *d_descent_list = nullptr);
int range_key_flags_device(const uint64_t *, const unsigned long long *, uint32_t, uint64_t, int, uint8_t *,
                           cudaStream_t);
int sort_segments_device(uint64_t *, uint64_t *, void *, void *, int, const unsigned long long *,
                         const unsigned long long *, uint32_t, int, cudaStream_t);
int repair_buckets_on_device(uint64_t *, uint64_t *, void *, void *, int, uint64_t, int, int, uint8_t *,
                             const unsigned int *, const unsigned long long *, int *, unsigned long long *,
                             cudaStream_t);
int repair_big_ambiguous_buckets_on_device(uint64_t *, void *, int, int, uint8_t *, unsigned long long *, int *,
                                           cudaStream_t);
int select_pairs_count(const uint8_t *, uint64_t, uint8_t, DeviceBuffer &, uint64_t *, cudaStream_t);
int select_pairs_write(const uint8_t *, uint64_t, uint8_t, const DeviceBuffer &, int, void *, const uint64_t *,
                       uint64_t *, const void *, void *, cudaStream_t);
int pack4_words_device(const uint8_t *, uint64_t, const void *, int, uint64_t, uint32_t, const uint64_t *,
                       uint64_t *, uint64_t *, cudaStream_t);
int rank4_stream_device(const uint8_t *, uint64_t, uint64_t *, cudaStream_t);
int pack4_words_stream_device(const uint64_t *, const void *, int, uint64_t, uint32_t, const uint64_t *, uint64_t *,
                              uint64_t *, cudaStream_t);
int sba_flags_device(const uint8_t *, uint64_t, const void *, int, uint64_t, uint32_t, const void *,
                     uint8_t, uint8_t *, cudaStream_t);
int group_hist_device(const void *, int, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, int64_t *,
                      int64_t *, int64_t *, cudaStream_t);
uint64_t last_hist_top_bin();
const std::vector<unsigned long long> &last_hist_pairs();
void set_last_hist_single(uint64_t, uint64_t);
void set_last_hist_pairs(const std::vector<unsigned long long> &, uint64_t);
int flag_group_spectrum_device(const uint8_t *, uint64_t, unsigned long long *, uint64_t, cudaStream_t);
int spectrum_small_bins();
int flag_group_hist_device(const uint8_t *, uint64_t, uint8_t, uint64_t, uint64_t, uint64_t, int64_t *,
                           int64_t *, int64_t *, cudaStream_t);
int filter_flags_device(const uint8_t *, uint64_t, const void *, int, uint64_t, const gk_filter &,
                        uint8_t *, cudaStream_t);
int select_flagged(const uint8_t *, uint64_t, uint8_t, int, void *, const void *, void *, const void *,
                   void *, uint64_t *, cudaStream_t);
int select_flagged_with_bytes(const uint8_t *, uint64_t, uint8_t, int, const void *, void *, const uint8_t *,
                              uint8_t *, uint8_t, uint64_t *, cudaStream_t);
int head_positions_device(const uint8_t *, const uint32_t *, uint64_t, uint32_t *, uint32_t *, cudaStream_t);
int valid_flags_device(const uint32_t *, uint64_t, const uint64_t *, uint32_t, uint64_t, uint32_t, uint8_t *,
                       cudaStream_t);
int gid_flags_device(const uint32_t *, uint64_t, uint8_t *, cudaStream_t);
int key2_scatter_device(const uint64_t *, const uint32_t *, const uint32_t *, uint64_t, uint32_t *, uint8_t *,
                        cudaStream_t);
int pair_keys_var_device(const uint32_t *, const uint32_t *, uint64_t, const uint32_t *, uint32_t, const uint64_t *,
                         uint32_t, uint64_t, uint64_t *, cudaStream_t);
int pair_keys_words_device(const void *, int, const uint32_t *, uint64_t, const uint8_t *, uint64_t, uint64_t, uint32_t,
                           uint64_t *, cudaStream_t);
int key2_scatter_any_device(const uint64_t *, const void *, const uint32_t *, uint64_t, int, void *, uint8_t *,
                            cudaStream_t);
int scatter_sorted_subset_device(const void *, const uint64_t *, const void *, uint64_t, int, int, uint64_t *, void *,
                                 uint8_t *, cudaStream_t);
int subset_rank_update_device(const uint32_t *, const uint32_t *, const uint32_t *, uint64_t, uint32_t *, uint32_t *,
                              cudaStream_t);
int gather_u32_device(const uint32_t *, const uint32_t *, uint64_t, uint32_t *, cudaStream_t);
int multi_flags_device(const uint8_t *, uint64_t, uint8_t *, cudaStream_t);
int clear_finished_multi_device(const uint32_t *, uint64_t, const uint64_t *, uint32_t, uint64_t, uint64_t, uint8_t *,
                                cudaStream_t);


// index-lifetime device allocation; stream-ordered like the scratch buffers so that creating and
// destroying an index per query does not pay a device-wide synchronising cudaFree
struct Owned {
    void *ptr = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;
    ~Owned() { reset(); }
    void reset()
    {
        if (ptr) pool_free(ptr, bytes, stream);
        ptr = nullptr;
        bytes = 0;
    }
    int alloc(size_t n, cudaStream_t st = nullptr)
    {
        reset();
        stream = st;
        if (n == 0) return GK_OK;
        const double t0 = trace_now_ms();
        GK_TRY(pool_alloc(&ptr, n, st));
        trace_alloc(trace_now_ms() - t0, n);
        bytes = n;
        return GK_OK;
    }
    void swap(Owned &o)
    {
        void *p = ptr; ptr = o.ptr; o.ptr = p;
        size_t b = bytes; bytes = o.bytes; o.bytes = b;
        cudaStream_t t = stream; stream = o.stream; o.stream = t;
    }
};

struct EventTimer {
    cudaEvent_t ev[16];
    int n = 0;
    cudaStream_t st;
    explicit EventTimer(cudaStream_t s) : st(s) {}
    ~EventTimer() { for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]); }
    int mark()
    {
        if (n >= 16) return n - 1;
        if (cudaEventCreate(&ev[n]) != cudaSuccess) return -1;
        cudaEventRecord(ev[n], st);
        return n++;
    }
    float ms(int a, int b)
    {
        float t = 0.f;
        if (a < 0 || b < 0) return 0.f;
        cudaEventElapsedTime(&t, ev[a], ev[b]);
        return t;
    }
};

}  // namespace gk

using namespace gk;

struct gk_index {
    const uint8_t *d_sba = nullptr;
    uint64_t sba_len = 0;
    std::vector<uint64_t> h_segs;
    Owned d_segs;                       // device copy of h_segs, uploaded by the first call that has a stream
    cudaStream_t segs_stream = nullptr;
    cudaEvent_t segs_ev = nullptr;
    ~gk_index() { if (segs_ev) cudaEventDestroy(segs_ev); }
    uint32_t min_len = 1, max_len = 0;  // max_len 0 == None
    int idx_bytes = 4;
    uint64_t n = 0;
    Owned d_idx;                        // start indices (init order until sorted)
    bool idx_ready = false;
    bool user_idx = false;              // d_idx was assigned by the caller (gk_index_set_indices)
    bool sorted = false;
    Owned d_flags;                      // head/amb flags of the sorted order for flags_kmer_len
    bool flags_valid = false;
    bool flags_mark_amb = false;        // kFlagAmb bits are meaningful (single-level sort only)
    uint32_t flags_kmer_len = 0;
    bool alphabet_known = false;
    uint64_t n_bad = 0, n_sep = 0, n_amb_letters = 0;
    // group-size spectrum of the sorted order for flags_kmer_len: (size, number of groups), ascending size
    std::vector<std::pair<uint64_t, uint64_t>> spectrum;
    bool spectrum_valid = false;
};

// The segment table on the device.  gk_index_create has no stream, and a synchronous copy there would wait for
// whatever the caller has queued (the both-strand layout) before the sort could be enqueued; the first call that
// brings a stream uploads it there (a few hundred bytes, staged at once), later calls on another stream wait for
// that upload's event.
static int ensure_segs(gk_index *ix, cudaStream_t st)
{
    if (!ix->d_segs.ptr) {
        GK_TRY(ix->d_segs.alloc(ix->h_segs.size() * 8, st));
        GK_CUDA(cudaMemcpyAsync(ix->d_segs.ptr, ix->h_segs.data(), ix->h_segs.size() * 8, cudaMemcpyHostToDevice, st));
        GK_CUDA(cudaEventCreateWithFlags(&ix->segs_ev, cudaEventDisableTiming));
        GK_CUDA(cudaEventRecord(ix->segs_ev, st));
        ix->segs_stream = st;
    } else if (st != ix->segs_stream && ix->segs_ev) {
        GK_CUDA(cudaStreamWaitEvent(st, ix->segs_ev, 0));
    }
    return GK_OK;
}

// Group-size spectrum behind a sort: kernels and the device-to-host copy are enqueued (no synchronise); finish()
// turns the copy into ix->spectrum after the caller's own synchronise.
constexpr uint64_t kSpectrumBig = 4096;   // large groups listed with their exact size; more: no spectrum
struct SpectrumPending {
    DeviceBuffer dev;
    std::vector<unsigned long long> host;
    bool armed = false;
    int start(const uint8_t *d_flags, uint64_t n, cudaStream_t st)
    {
        const size_t words = (size_t)spectrum_small_bins() + 1 + kSpectrumBig;
        GK_TRY(dev.alloc(words * 8, st));
        GK_TRY(flag_group_spectrum_device(d_flags, n, dev.as<unsigned long long>(), kSpectrumBig, st));
        host.assign(words, 0);
        GK_CUDA(cudaMemcpyAsync(host.data(), dev.ptr, words * 8, cudaMemcpyDeviceToHost, st));
        armed = true;
        return GK_OK;
    }
    void finish(gk_index *ix)
    {
        ix->spectrum.clear();
        ix->spectrum_valid = false;
        if (!armed) return;
        const size_t small = (size_t)spectrum_small_bins();
        const uint64_t n_big = host[small];
        if (n_big > kSpectrumBig) return;   // too many large groups to list: queries take the device path
        for (size_t i = 1; i < small; ++i)
            if (host[i]) ix->spectrum.emplace_back((uint64_t)i, (uint64_t)host[i]);
        std::vector<uint64_t> big(host.begin() + small + 1, host.begin() + small + 1 + n_big);
        std::sort(big.begin(), big.end());
        for (size_t i = 0; i < big.size();) {
            size_t j = i;
            while (j < big.size() && big[j] == big[i]) ++j;
            ix->spectrum.emplace_back(big[i], (uint64_t)(j - i));
            i = j;
        }
        ix->spectrum_valid = true;
    }
};

static int ensure_alphabet(gk_index *ix, cudaStream_t st)
{
    if (ix->alphabet_known) return GK_OK;
    uint64_t c[3];
    GK_TRY(gk_sba_scan_alphabet(ix->d_sba, ix->sba_len, c, st));
    ix->n_bad = c[0];
    ix->n_sep = c[1];
    ix->n_amb_letters = c[2];
    ix->alphabet_known = true;
    return GK_OK;
}

static int ensure_indices(gk_index *ix, cudaStream_t st)
{
    GK_TRY(ensure_segs(ix, st));
    if (ix->idx_ready) return GK_OK;
    GK_TRY(ix->d_idx.alloc((size_t)ix->n * ix->idx_bytes, st));
    GK_TRY(init_indices_device((const uint64_t *)ix->d_segs.ptr, (uint32_t)ix->h_segs.size(),
                               ix->min_len, ix->n, ix->idx_bytes, ix->d_idx.ptr, st));
    ix->idx_ready = true;
    return GK_OK;
}

// Stage timing marks of one sort; turned into milliseconds after the final synchronise.
struct StageMarks {
    int pack0 = -1, pack1 = -1, fix0 = -1, fix1 = -1;
    SortTiming main_sort;
    uint64_t n_amb = 0;
    uint64_t n_frag = 0;
    uint32_t refine_flags = 0;  // 1 fragment path, 2 element-wise repair ran, 4 descent, fragment error bits << 8
    int key_bits = 0;
    int levels = 1;
};

// A second stream for work that does not depend on the main sort (reading back the pack kernel's counters,
// sorting the fragment list): one per host thread and device, created on first use.
static int side_stream(cudaStream_t *out)
{
    static thread_local cudaStream_t streams[64] = {nullptr};
    int dev = 0;
    GK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return GK_ERR_ARG; }
    if (!streams[dev]) {
        // highest priority: its few small kernels (fragment sort) must slip in between the CTAs of the main
        // stream's radix passes instead of queueing behind whole kernels
        int least = 0, greatest = 0;
        GK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        const char *e = getenv("GK_SIDE_PRIORITY");
        GK_CUDA(cudaStreamCreateWithPriority(&streams[dev], cudaStreamNonBlocking,
                                             (e && e[0] == '0') ? least : greatest));
    }
    *out = streams[dev];
    return GK_OK;
}

struct ScopedEvent {
    cudaEvent_t ev = nullptr;
    ~ScopedEvent() { if (ev) cudaEventDestroy(ev); }
    int create() { GK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); return GK_OK; }
};

static bool spectrum_enabled()
{
    const char *e = getenv("GK_SPECTRUM");
    return !(e && e[0] == '0');
}
static bool fragments_enabled()
{
    const char *e = getenv("GK_FRAGMENTS");
    return !(e && e[0] == '0');
}
constexpr uint64_t kFragCapacity = 1ull << 21;
constexpr uint64_t kFragReserve = 1ull << 17;   // fragment-sort scratch reserved on the main stream up front  // fragments listed by the pack kernel; beyond: element-wise path

// First key bit the LSD passes of the main sort cover.  0 = plain LSD over the whole key.  Otherwise only
// the top 8*ceil((log2(n)+4)/8) bits are sorted -- 16 times more prefix buckets than k-mers, so about one
// k-mer in twenty shares its bucket -- and tie_fix_flags orders each bucket by the remaining low bits while
// it computes the head flags.  GK_SORT_HYBRID=0 switches this off; GK_SORT_PREFIX_BITS=b
// forces a prefix width (tests use it to drive tiny inputs through the tie-repair and long-run paths).
static int prefix_begin_bit(uint64_t n, int key_bits)
{
    const char *off = getenv("GK_SORT_HYBRID");
    if (off && off[0] == '0') return 0;
    const int total_passes = (key_bits + 7) / 8;
    int prefix_passes;
    const char *force = getenv("GK_SORT_PREFIX_BITS");
    if (force && *force) {
        prefix_passes = (atoi(force) + 7) / 8;
        if (prefix_passes < 1) prefix_passes = 1;
    } else {
        if (n < (1ull << 16)) return 0;
        int need = 4;
        while ((1ull << (need - 4)) < n && need < 68) ++need;  // 2^need >= 16 n
        prefix_passes = (need + 7) / 8;
    }
    if (prefix_passes >= total_passes) return 0;
    return key_bits - 8 * prefix_passes;
}

// After the main sort and its flags pass, two kinds of slots may still hold the wrong element:
//   * ambiguous windows: in the right slots as a set, ordered by `value` only -> order them by their
//     terminator-aware 4-bit rank words (16 symbols per word);
//   * members of long prefix runs (kFlagLong), when the flags pass saw one of them out of order.
// Both sets are mostly huge blocks of identical elements (N runs, exact repeats), so they are sorted in
// run-length compressed form (gk_refine.cu): cost follows the number of blocks, not of elements.
// d_idx holds the sorted starts and is updated in place, and so are the flags of the repaired slots.
static int refine_subset(gk_index *ix, const uint64_t *keys_sorted, void *d_idx, uint8_t *d_flags, uint64_t n,
                         int class_bit, uint32_t key_len, int key_bits, uint64_t n_amb, bool descent,
                         bool terminated, cudaStream_t st)
{
    const int ib = ix->idx_bytes;
    const uint8_t mask = (uint8_t)((n_amb ? kFlagAmb : 0) | (descent ? kFlagLong : 0));
    if (!mask || n == 0) return GK_OK;
    uint64_t m = 0;
    DeviceBuffer sel_temp;
    GK_TRY(select_pairs_count(d_flags, n, mask, sel_temp, &m, st));
    if (m < n_amb) {
        set_error("ambiguous window count mismatch: packed %llu, selected %llu", (unsigned long long)n_amb,
                  (unsigned long long)m);
        return GK_ERR_INTERNAL;
    }
    if (m == 0) return GK_OK;
    DeviceBuffer pos, key, idx, w0, w1, rh, rep_start;
    GK_TRY(pos.alloc((size_t)m * ib, st));
    GK_TRY(key.alloc((size_t)m * 8, st));
    GK_TRY(idx.alloc((size_t)m * ib, st));
    GK_TRY(select_pairs_write(d_flags, n, mask, sel_temp, ib, pos.ptr, keys_sorted, key.as<uint64_t>(), d_idx,
                              idx.ptr, st));
    const int words = n_amb ? ((int)key_len + 15) / 16 : 0;
    const uint64_t *w0p = nullptr, *w1p = nullptr;
    if (words >= 1) {
        GK_TRY(w0.alloc((size_t)m * 8, st));
        if (words >= 2) GK_TRY(w1.alloc((size_t)m * 8, st));
        if (m >= ix->sba_len / 64 && !terminated) {
            // many windows: 4-bit rank stream of the whole byte array (half a byte per position), then
            // two funnel shifts per window
            DeviceBuffer stream;
            GK_TRY(stream.alloc((size_t)(ix->sba_len / 16 + 3) * 8, st));
            GK_TRY(rank4_stream_device(ix->d_sba, ix->sba_len, stream.as<uint64_t>(), st));
            GK_TRY(pack4_words_stream_device(stream.as<uint64_t>(), idx.ptr, ib, m, key_len, key.as<uint64_t>(),
                                             w0.as<uint64_t>(), words >= 2 ? w1.as<uint64_t>() : nullptr, st));
        } else {
            // few windows, or windows that may end at a '$' (variable-length mode): read their bytes
            // directly; this kernel stops at the terminator (kmers.py:360-378)
            GK_TRY(pack4_words_device(ix->d_sba, ix->sba_len, idx.ptr, ib, m, key_len, key.as<uint64_t>(),
                                      w0.as<uint64_t>(), words >= 2 ? w1.as<uint64_t>() : nullptr, st));
        }
        w0p = w0.as<uint64_t>();
        if (words >= 2) w1p = w1.as<uint64_t>();
    }
    GK_TRY(rh.alloc((size_t)((m + 15) & ~15ull), st));
    GK_TRY(subset_rep_flags_device(key.as<uint64_t>(), w0p, w1p, m, rh.as<uint8_t>(), st));
    GK_TRY(rep_start.alloc((size_t)m * ib, st));
    uint64_t R = 0;
    GK_TRY(select_flagged(rh.as<uint8_t>(), m, kFlagHead, ib, rep_start.ptr, nullptr, nullptr, nullptr, nullptr,
                          &R, st));

    // stable LSD over the words of the block representatives, least significant word first
    DeviceBuffer perm_a, perm_b, rk_a, rk_b, out_off;
    GK_TRY(perm_a.alloc((size_t)R * ib, st));
    GK_TRY(perm_b.alloc((size_t)R * ib, st));
    GK_TRY(rk_a.alloc((size_t)(R + 1) * 8, st));
    GK_TRY(rk_b.alloc((size_t)(R + 1) * 8, st));
    GK_TRY(iota_device(perm_a.ptr, R, ib, st));
    void *cur = perm_a.ptr, *alt = perm_b.ptr;
    auto sort_word = [&](const uint64_t *word, int begin_bit, int end_bit) -> int {
        GK_TRY(gather2_u64_device(word, rep_start.ptr, cur, R, ib, rk_a.as<uint64_t>(), st));
        int alt_has = 0;
        GK_TRY(radix_sort_pairs_device(rk_a.as<uint64_t>(), rk_b.as<uint64_t>(), cur, alt, ib, R, begin_bit,
                                       end_bit, &alt_has, st, nullptr));
        if (alt_has) { void *t = cur; cur = alt; alt = t; }
        return GK_OK;
    };
    if (R > 1) {
        if (words >= 2) GK_TRY(sort_word(w1p, 64 - 4 * ((int)key_len - 16), 64));
        if (words >= 1) GK_TRY(sort_word(w0p, 64 - 4 * ((int)key_len < 16 ? (int)key_len : 16), 64));
        GK_TRY(sort_word(key.as<uint64_t>(), 0, key_bits));
    }
    GK_TRY(out_off.alloc((size_t)R * 8, st));
    GK_TRY(block_offsets_device(rep_start.ptr, R, m, cur, ib, out_off.as<unsigned long long>(), st));
    GK_TRY(subset_expand_device(pos.ptr, idx.ptr, key.as<uint64_t>(), w0p, w1p, rep_start.ptr, cur,
                                out_off.as<unsigned long long>(), R, m, class_bit, ib, d_idx, d_flags, st));
    return GK_OK;  // the scratch above is released in stream order (cudaFreeAsync): no synchronise needed
}

// Too many out-of-order long prefix runs for the bucket-wise repair (a repeat-rich genome: millions of k-mers
// that agree on their first 16 symbols): every member of every long run is taken out, the set is sorted by the
// full key (the members come out in prefix order, so a stable sort puts every run's members back into the
// run's own slots) and written back with fresh flags -- keys too, so that the fragments can be expanded
// afterwards.  Cost: one ordered select, 8 radix passes over the set, one scatter.  No synchronise.
static int repair_long_runs(uint64_t *keys_sorted, void *d_idx, uint8_t *d_flags, uint64_t n, int ib, int class_bit,
                            int key_bits, uint64_t *m_out, cudaStream_t st)
{
    uint64_t m = 0;
    DeviceBuffer sel_temp;
    GK_TRY(select_pairs_count(d_flags, n, kFlagLong, sel_temp, &m, st));
    if (m_out) *m_out = m;
    if (m == 0) return GK_OK;
    DeviceBuffer pos, key, key_alt, idx, idx_alt;
    GK_TRY(pos.alloc((size_t)m * ib, st));
    GK_TRY(key.alloc((size_t)m * 8, st));
    GK_TRY(key_alt.alloc((size_t)m * 8, st));
    GK_TRY(idx.alloc((size_t)m * ib, st));
    GK_TRY(idx_alt.alloc((size_t)m * ib, st));
    GK_TRY(select_pairs_write(d_flags, n, kFlagLong, sel_temp, ib, pos.ptr, keys_sorted, key.as<uint64_t>(), d_idx,
                              idx.ptr, st));
    int in_alt = 0;
    GK_TRY(radix_sort_pairs_device(key.as<uint64_t>(), key_alt.as<uint64_t>(), idx.ptr, idx_alt.ptr, ib, m, 0, key_bits,
                                   &in_alt, st, nullptr));
    GK_TRY(scatter_sorted_subset_device(pos.ptr, in_alt ? key_alt.as<uint64_t>() : key.as<uint64_t>(),
                                        in_alt ? idx_alt.ptr : idx.ptr, m, class_bit, ib, keys_sorted, d_idx, d_flags,
                                        st));
    return GK_OK;
}

// Out-of-order prefix buckets that were too long for the device-side repair (one CTA per bucket): sort each
// by its low key bits with the radix passes, in place, and recompute its flags.  h_ranges: n_ranges (lo, hi)
// slot ranges as listed by repair_buckets_kernel.  No synchronise.
static int repair_big_buckets(uint64_t *keys, uint64_t *keys_tmp, void *idx, void *idx_tmp, int ib, int lo_bits,
                              int class_bit, uint8_t *d_flags, const unsigned long long *d_ranges,
                              const unsigned long long *h_ranges, uint32_t n_ranges, cudaStream_t st)
{
    uint64_t longest = 0;
    for (uint32_t i = 0; i < n_ranges; ++i) {
        const uint64_t len = h_ranges[2 * i + 1] - h_ranges[2 * i];
        longest = len > longest ? len : longest;
    }
    GK_TRY(sort_segments_device(keys, keys_tmp, idx, idx_tmp, ib, d_ranges, h_ranges, n_ranges, lo_bits, st));
    GK_TRY(range_key_flags_device(keys, d_ranges, n_ranges, longest, class_bit, d_flags, st));
    return GK_OK;
}

// device counters of one sort: [0] ambiguous windows, [1] low half: descents, high half: repair status (0 = the
// order is final), [2] fragments, [3] fragment error bits (int), [4] number of big out-of-order buckets, then
// their kBigBucketCap (lo, hi) ranges, [kCounterWords ...] the first kDescentCap descent positions
constexpr int kCounterWords = 5 + 2 * kBigBucketCap;
constexpr size_t kCounterBytes = (size_t)(kCounterWords + kDescentCap) * 8;

// (key, start) pairs ready to be sorted, and what is known about them.
struct PackedPairs {
    uint64_t *keys_a = nullptr, *keys_b = nullptr;   // ping-pong key buffers, keys in keys_a
    void *idx_a = nullptr, *idx_b = nullptr;         // ping-pong start buffers, starts in idx_a
    void *idx_final = nullptr;     // optional third buffer that must receive the sorted starts
    uint64_t n = 0;
    uint32_t key_len = 0;
    int key_bits = 0, class_bit = 0;
    bool terminated = false;       // windows may end at a '$' inside the key (variable-length mode)
    const unsigned long long *d_pre_hist = nullptr;   // digit histograms counted by the producer, or null
    FragOut frag;                  // fragment list of the ambiguous windows (counter = d_counters + 2)
    bool want_frag = false;
    FragSorted *fs = nullptr;
    bool fs_ready = false;         // *fs already holds the sorted fragments (merged from sorted lists): only F is open
    unsigned long long *d_counters = nullptr;   // kCounterBytes, zeroed before the producer ran
    bool producer_counted_amb = true;   // d_counters[0] is final when e_producer fires (else the flags pass counts)
    cudaEvent_t e_producer = nullptr;   // recorded on the stream behind the producer of keys and fragments
    int start_bits = 32;
    unsigned long long *d_alpha = nullptr, *h_alpha = nullptr;   // alphabet counters still in flight (optional)
    cudaEvent_t e_alpha = nullptr;      // ... on another stream: recorded behind the scan
    SpectrumPending *spectrum = nullptr;   // optional: group-size spectrum of the final order, same synchronise
    const int *d_extra_err = nullptr;   // optional device int that must be zero at the end (partition look-back)
    // results
    uint64_t *keys_sorted = nullptr;
    void *idx_sorted = nullptr;
};

// Sort packed pairs, repair what the prefix sort leaves open, expand the fragments, leave starts and flags in
// their final order.  Synchronises once, at the end.
//
// Stream plan (nothing on the main stream ever waits for the host):
//   main   producer (pack + fragment list) | E1 | histogram scan, radix passes, tie repair + flags,
//          bucket-wise repair | wait E2 | fragment expand
//   side   wait E1 | counters -> host | fragment sort | E2
// The host blocks on the side stream only: that copy completes when the producer does, milliseconds before the
// main stream runs dry, and tells how many fragments there are to sort.
static int sort_packed_pairs(gk_index *ix, PackedPairs &pp, uint8_t *d_flags, StageMarks &marks, EventTimer &tm,
                             cudaStream_t st)
{
    const int ib = ix->idx_bytes;
    const uint64_t n = pp.n;
    unsigned long long *d_counters = pp.d_counters;
    unsigned int *d_descent = reinterpret_cast<unsigned int *>(d_counters + 1);
    int *d_status = reinterpret_cast<int *>(d_descent + 1);
    int *d_frag_err = reinterpret_cast<int *>(d_counters + 3);
    const int begin_bit = prefix_begin_bit(n, pp.key_bits);
    marks.key_bits = pp.key_bits;

    int in_alt = 0;
    bool amb_counted = false;
    trace_point("pairs_ready");
    // (inlined sort_pairs_and_flag: the third start buffer needs the passes' own bookkeeping)
    GK_CUDA(cudaMemsetAsync(d_descent, 0, 4, st));
    GK_TRY(radix_sort_pairs_device(pp.keys_a, pp.keys_b, pp.idx_a, pp.idx_b, ib, n, begin_bit, pp.key_bits, &in_alt, st,
                                   &marks.main_sort, pp.d_pre_hist, pp.idx_final));
    trace_point("passes_enqueued");
    uint64_t *keys_sorted = in_alt ? pp.keys_b : pp.keys_a;
    uint64_t *keys_other = in_alt ? pp.keys_a : pp.keys_b;
    void *idx_sorted = pp.idx_final ? pp.idx_final : (in_alt ? pp.idx_b : pp.idx_a);
    // scratch of the same size for the in-place repairs: a start buffer that does not hold the result
    void *idx_other = (idx_sorted == pp.idx_b) ? pp.idx_a : pp.idx_b;
    pp.keys_sorted = keys_sorted;
    pp.idx_sorted = idx_sorted;
    unsigned long long *d_n_amb = pp.producer_counted_amb ? nullptr : d_counters;
    if (begin_bit == 0) {
        GK_TRY(key_flags_device(keys_sorted, n, pp.class_bit, d_flags, st));
    } else {
        GK_TRY(tie_fix_flags_device(keys_sorted, idx_sorted, ib, n, begin_bit, pp.class_bit, d_flags, d_descent,
                                    d_n_amb, st, d_counters + kCounterWords));
        amb_counted = d_n_amb != nullptr;
    }
    marks.fix0 = tm.mark();
    // long prefix runs that came out of order are re-sorted bucket by bucket, driven from the device list
    if (begin_bit > 0)
        GK_TRY(repair_buckets_on_device(keys_sorted, keys_other, idx_sorted, idx_other, ib, n, begin_bit,
                                        pp.class_bit, d_flags, d_descent, d_counters + kCounterWords, d_status,
                                        d_counters + 4, st));

    // ---- side stream: the producer's counters, then the fragment sort -----------------------------------------
    unsigned long long h_counters[kCounterWords] = {0};
    uint64_t n_amb = 0, n_frag = 0;
    bool use_frag = false;
    ScopedEvent e_frag;
    if (pp.want_frag) {
        cudaStream_t side = nullptr;
        GK_TRY(side_stream(&side));
        GK_TRY(e_frag.create());
        trace_point("repair_enqueued");
        GK_CUDA(cudaStreamWaitEvent(side, pp.e_producer, 0));
        GK_CUDA(cudaMemcpyAsync(h_counters, d_counters, 32, cudaMemcpyDeviceToHost, side));
        GK_CUDA(cudaStreamSynchronize(side));
        trace_point("side_sync");
        n_amb = h_counters[0];
        n_frag = h_counters[2];
        use_frag = n_amb > 0 && n_frag > 0 && n_frag <= pp.frag.capacity;
        if (use_frag) {
            // big out-of-order buckets made of one ambiguous key and a few strangers (the all-N bucket): only the
            // strangers move, the fragments rewrite the rest
            if (begin_bit > 0)
                GK_TRY(repair_big_ambiguous_buckets_on_device(keys_sorted, idx_sorted, ib, pp.class_bit, d_flags,
                                                              d_counters + 4, d_status, st));
            if (pp.fs_ready) {
                pp.fs->F = n_frag;
            } else {
                GK_TRY(frag_sort_device(pp.frag, n_frag, pp.key_len, pp.key_bits, pp.start_bits, *pp.fs, side));
                GK_CUDA(cudaEventRecord(e_frag.ev, side));
                GK_CUDA(cudaStreamWaitEvent(st, e_frag.ev, 0));
            }
            GK_TRY(frag_expand_device(*pp.fs, keys_sorted, n, ib, idx_sorted, d_flags,
                                      reinterpret_cast<const unsigned int *>(d_status), d_counters, n_amb, d_frag_err,
                                      st));
            pp.fs->rebind(st);
        }
    }
    marks.fix1 = tm.mark();
    if (pp.spectrum) GK_TRY(pp.spectrum->start(d_flags, n, st));
    trace_point("all_enqueued");

    // ---- the one synchronise: descents, repair status, fragment check, alphabet counters -----------------------
    int h_extra_err = 0;
    GK_CUDA(cudaMemcpyAsync(h_counters, d_counters, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
    if (pp.d_alpha) {
        if (pp.e_alpha) GK_CUDA(cudaStreamWaitEvent(st, pp.e_alpha, 0));
        GK_CUDA(cudaMemcpyAsync(pp.h_alpha, pp.d_alpha, 24, cudaMemcpyDeviceToHost, st));
    }
    if (pp.d_extra_err) GK_CUDA(cudaMemcpyAsync(&h_extra_err, pp.d_extra_err, 4, cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    trace_point("final_sync");
    if (h_extra_err) {
        set_error("partition: decoupled look-back timed out");
        return GK_ERR_INTERNAL;
    }
    if (!pp.want_frag) {
        n_amb = h_counters[0];
        if (pp.class_bit && n && !pp.producer_counted_amb && !amb_counted)   // plain LSD: count the marked slots
            GK_TRY(select_flagged(d_flags, n, kFlagAmb, ib, nullptr, nullptr, nullptr, nullptr, nullptr, &n_amb, st));
    }
    marks.n_amb = n_amb;
    marks.n_frag = n_frag;
    const uint32_t n_descent = (uint32_t)(h_counters[1] & 0xffffffffull);
    const int status = (int)(h_counters[1] >> 32);
    int frag_err = (int)(h_counters[3] & 0xffffffffull);
    marks.refine_flags = (use_frag ? 1u : 0u) | (n_descent ? 4u : 0u) | (n_descent && !status ? 16u : 0u) |
                         ((uint32_t)frag_err << 8);
    if (getenv("GK_TRACE"))
        fprintf(stderr, "[gk trace] sort: n=%llu ambiguous=%llu fragments=%llu use_frag=%d descents=%u status=%d "
                        "frag_err=%d\n", (unsigned long long)n, (unsigned long long)n_amb, (unsigned long long)n_frag,
                (int)use_frag, n_descent, status, frag_err);
    if (getenv("GK_TRACE") && status)
        for (unsigned i = 0; i < (unsigned)h_counters[4] && i < (unsigned)kBigBucketCap; ++i)
            fprintf(stderr, "[gk trace]   big bucket %u: [%llu, %llu)\n", i, h_counters[5 + 2 * i], h_counters[6 + 2 * i]);
    bool elementwise_long = false;
    if (status) {
        // the device-side repair left work behind: a bucket too long for one CTA (sorted from here with the
        // radix passes), or more descents than the list holds (element-wise repair of every long run)
        const int f0 = tm.mark();
        if (!(status & 2)) {
            const uint32_t n_big = (uint32_t)h_counters[4];
            GK_TRY(repair_big_buckets(keys_sorted, keys_other, idx_sorted, idx_other, ib, begin_bit, pp.class_bit,
                                      d_flags, d_counters + 5, h_counters + 5, n_big, st));
            marks.refine_flags |= 32u;
            if (use_frag && !frag_err) {   // the fragment kernels did nothing while the keys were out of order
                GK_CUDA(cudaMemsetAsync(d_status, 0, 4, st));
                GK_TRY(frag_expand_device(*pp.fs, keys_sorted, n, ib, idx_sorted, d_flags,
                                          reinterpret_cast<const unsigned int *>(d_status), d_counters, n_amb,
                                          d_frag_err, st));
            }
            GK_CUDA(cudaMemcpyAsync(h_counters, d_counters, 32, cudaMemcpyDeviceToHost, st));
            GK_CUDA(cudaStreamSynchronize(st));
            frag_err = (int)(h_counters[3] & 0xffffffffull);
        } else if (use_frag && !frag_err) {
            // more out-of-order long runs than the device list holds: re-sort all long runs by their keys, then
            // let the fragments write the ambiguous slots (they did nothing while the keys were out of order)
            uint64_t m_long = 0;
            GK_TRY(repair_long_runs(keys_sorted, idx_sorted, d_flags, n, ib, pp.class_bit, pp.key_bits, &m_long, st));
            marks.refine_flags |= 64u;
            GK_CUDA(cudaMemsetAsync(d_status, 0, 4, st));
            GK_TRY(frag_expand_device(*pp.fs, keys_sorted, n, ib, idx_sorted, d_flags,
                                      reinterpret_cast<const unsigned int *>(d_status), d_counters, n_amb, d_frag_err,
                                      st));
            GK_CUDA(cudaMemcpyAsync(h_counters, d_counters, 32, cudaMemcpyDeviceToHost, st));
            GK_CUDA(cudaStreamSynchronize(st));
            frag_err = (int)(h_counters[3] & 0xffffffffull);
        } else {
            elementwise_long = true;
        }
        marks.fix0 = f0;
        marks.fix1 = tm.mark();
    }
    if (elementwise_long || (n_amb > 0 && !(use_frag && !frag_err))) {
        // element-wise repair (gk_refine.cu): the ambiguous windows by their 4-bit rank words when there is no
        // fragment list for them, and the members of every long prefix run when too many are out of order
        const int f0 = tm.mark();
        GK_TRY(refine_subset(ix, keys_sorted, idx_sorted, d_flags, n, pp.class_bit, pp.key_len, pp.key_bits, n_amb,
                             elementwise_long, pp.terminated, st));
        GK_CUDA(cudaStreamSynchronize(st));
        if (!status) marks.fix0 = f0;
        marks.fix1 = tm.mark();
        marks.refine_flags |= 2u;
    }
    if (pp.spectrum && (status || (marks.refine_flags & 2u))) {   // the order changed after the first spectrum
        GK_TRY(pp.spectrum->start(d_flags, n, st));
        GK_CUDA(cudaStreamSynchronize(st));
    }
    return GK_OK;
}

// Level 1: every window of valid_len symbols, ordered by its first key_len <= 32 symbols.
// Leaves the sorted starts in out_idx and the head flags (for key_len) in out_flags; synchronises once, at
// the end.  d_alpha (optional): the three alphabet counters of a scan that is still in flight on `st`; they
// come back with the other counters (h_alpha).  d_list != nullptr: sort these starts (ascending, n of them)
// instead of every window of the byte array.
static int sort_level1(gk_index *ix, uint32_t valid_len, uint32_t key_len, int class_bit, uint64_t n,
                       Owned &out_idx, Owned &out_flags, StageMarks &marks, EventTimer &tm, cudaStream_t st,
                       unsigned long long *d_alpha, unsigned long long *h_alpha, const void *d_list = nullptr,
                       SpectrumPending *spectrum = nullptr, cudaEvent_t e_alpha = nullptr)
{
    const int ib = ix->idx_bytes;
    PackedPairs pp;
    pp.spectrum = spectrum;
    pp.e_alpha = e_alpha;
    pp.n = n;
    pp.key_len = key_len;
    pp.class_bit = class_bit;
    // value <= 4^key_len needs 2*key_len+1 bits when ambiguous windows exist, plus the class bit
    pp.key_bits = 2 * (int)key_len + (class_bit ? 2 : 0);
    pp.terminated = valid_len < key_len;
    pp.d_alpha = d_alpha;
    pp.h_alpha = h_alpha;

    DeviceBuffer keys_a, keys_b, counters, pre_hist, frag_mem;
    Owned idx_b;
    FragSorted fs;
    GK_TRY(keys_a.alloc((size_t)n * 8, st));
    GK_TRY(keys_b.alloc((size_t)n * 8, st));
    GK_TRY(idx_b.alloc((size_t)n * ib, st));
    GK_TRY(out_idx.alloc((size_t)n * ib, st));
    GK_TRY(counters.alloc(kCounterBytes, st));
    GK_CUDA(cudaMemsetAsync(counters.ptr, 0, (size_t)kCounterWords * 8, st));
    pp.d_counters = counters.as<unsigned long long>();
    pp.keys_a = keys_a.as<uint64_t>();
    pp.keys_b = keys_b.as<uint64_t>();
    pp.idx_a = out_idx.ptr;
    pp.idx_b = idx_b.ptr;
    pp.fs = &fs;
    // the pack kernel counts the digits of the radix passes while it writes the keys
    const int begin_bit = prefix_begin_bit(n, pp.key_bits);
    GK_TRY(pre_hist.alloc(8 * 256 * sizeof(unsigned long long), st));
    GK_CUDA(cudaMemsetAsync(pre_hist.ptr, 0, pre_hist.bytes, st));
    // fragment list of the ambiguous windows (not for an arbitrary list of starts: no neighbours there)
    pp.want_frag = class_bit && !d_list && fragments_enabled();
    if (pp.want_frag) {
        const uint64_t cap = n < kFragCapacity ? n : kFragCapacity;
        GK_TRY(frag_mem.alloc((size_t)cap * 36, st));
        uint64_t *base = frag_mem.as<uint64_t>();
        pp.frag.key = base; pp.frag.w0 = base + cap; pp.frag.w1 = base + 2 * cap; pp.frag.start = base + 3 * cap;
        pp.frag.count = reinterpret_cast<uint32_t *>(base + 4 * cap);
        pp.frag.counter = pp.d_counters + 2;
        pp.frag.capacity = cap;
        GK_TRY(fs.reserve(cap < kFragReserve ? cap : kFragReserve, st));
    }
    pp.start_bits = 1;
    while (pp.start_bits < 64 && (ix->sba_len >> pp.start_bits)) ++pp.start_bits;

    marks.pack0 = tm.mark();
    if (d_list) {
        GK_CUDA(cudaMemcpyAsync(out_idx.ptr, d_list, (size_t)n * ib, cudaMemcpyDeviceToDevice, st));
        GK_TRY(pack_keys_list_device(ix->d_sba, ix->sba_len, d_list, ib, n, key_len, class_bit, pp.keys_a,
                                     pp.d_counters, st));
    } else {
        // (more digit positions than the pack kernel has counters for -- plain LSD over a long key on a small
        // input: the sort counts its digits itself)
        const char *nohist = getenv("GK_PACK_NOHIST");   // (tuning: what do the pack kernel's histograms cost?)
        const bool count_digits = (pp.key_bits - begin_bit + 7) / 8 <= pack_hist_passes_max() &&
                                  !(nohist && nohist[0] == '1');
        GK_TRY(pack_keys_device(ix->d_sba, ix->sba_len, (const uint64_t *)ix->d_segs.ptr,
                                (uint32_t)ix->h_segs.size(), valid_len, key_len, class_bit, 0, ix->sba_len, 0,
                                pp.keys_a, ib, out_idx.ptr, pp.d_counters, begin_bit, pp.key_bits,
                                count_digits ? pre_hist.as<unsigned long long>() : nullptr, st,
                                pp.want_frag ? &pp.frag : nullptr));
        if (count_digits) pp.d_pre_hist = pre_hist.as<unsigned long long>();
    }
    marks.pack1 = tm.mark();
    ScopedEvent e_pack;
    GK_TRY(e_pack.create());
    GK_CUDA(cudaEventRecord(e_pack.ev, st));
    pp.e_producer = e_pack.ev;
    GK_TRY(out_flags.alloc((size_t)((n + 15) & ~15ull), st));
    GK_TRY(sort_packed_pairs(ix, pp, (uint8_t *)out_flags.ptr, marks, tm, st));
    if (pp.idx_sorted == idx_b.ptr) out_idx.swap(idx_b);
    return GK_OK;
}

// Prefix doubling (gk_refine.cu) for k-mers longer than one key word and for the variable-length mode.
// In: every start of every record, sorted by its first h0 symbols with a window that reaches its record's
// '$' first sorting first; flags = groups of equal h0-prefixes.  A round h -> h2 <= 2h orders the members
// of every group with more than one element by the pair (group, rank_h[start + h2 - h]); an empty second
// half (the window ended) sorts first.  Only the O(n) set-up touches every k-mer: rank_h of every start
// (= sorted position of the first member of its group) and the list of members of multi-element groups.
// The rounds then work on that shrinking list only and scatter their results into the global order.
static int doubling_rounds(gk_index *ix, Owned &cur_idx, Owned &cur_flags, uint64_t n_cur, uint64_t h0,
                           uint64_t target, int *levels, cudaStream_t st)
{
    if (h0 >= target || n_cur < 2) return GK_OK;
    uint32_t *d_idx = (uint32_t *)cur_idx.ptr;
    uint8_t *d_flags = (uint8_t *)cur_flags.ptr;
    // members of groups with more than one element, straight from the head flags: when there are none
    // (every k-mer already differs within the first h0 symbols) no rank is ever looked up
    DeviceBuffer rank, gid, mflags;
    GK_TRY(mflags.alloc((size_t)((n_cur + 15) & ~15ull), st));
    GK_TRY(multi_flags_device(d_flags, n_cur, mflags.as<uint8_t>(), st));
    GK_TRY(clear_finished_multi_device(d_idx, n_cur, (const uint64_t *)ix->d_segs.ptr, (uint32_t)ix->h_segs.size(),
                                       ix->sba_len, h0, mflags.as<uint8_t>(), st));
    uint64_t m = 0;
    GK_TRY(select_flagged(mflags.as<uint8_t>(), n_cur, kFlagMulti, 4, nullptr, nullptr, nullptr, nullptr, nullptr,
                          &m, st));
    if (m == 0) return GK_OK;
    GK_TRY(rank.alloc((size_t)ix->sba_len * 4, st));
    GK_TRY(gid.alloc((size_t)n_cur * 4, st));
    GK_TRY(head_positions_device(d_flags, d_idx, n_cur, gid.as<uint32_t>(), rank.as<uint32_t>(), st));
    DeviceBuffer slots, sub_idx, sub_gid;
    if (m) {
        GK_TRY(slots.alloc((size_t)m * 4, st));
        GK_TRY(sub_idx.alloc((size_t)m * 4, st));
        GK_TRY(sub_gid.alloc((size_t)m * 4, st));
        GK_TRY(select_flagged(mflags.as<uint8_t>(), n_cur, kFlagMulti, 4, slots.ptr, d_idx, sub_idx.ptr, gid.ptr,
                              sub_gid.ptr, nullptr, st));
    }
    gid.release();
    mflags.release();
    uint64_t h = h0;
    while (h < target && m > 0) {
        const uint64_t h2 = (2 * h < target) ? 2 * h : target;
        const uint32_t delta = (uint32_t)(h2 - h);
        DeviceBuffer keys, keys_alt, idx_alt, hf, gsub, gslot, mf;
        GK_TRY(keys.alloc((size_t)m * 8, st));
        GK_TRY(keys_alt.alloc((size_t)m * 8, st));
        GK_TRY(idx_alt.alloc((size_t)m * 4, st));
        GK_TRY(pair_keys_var_device(sub_idx.as<uint32_t>(), sub_gid.as<uint32_t>(), m, rank.as<uint32_t>(), delta,
                                    (const uint64_t *)ix->d_segs.ptr, (uint32_t)ix->h_segs.size(), ix->sba_len,
                                    keys.as<uint64_t>(), st));
        int in_alt = 0;
        GK_TRY(radix_sort_pairs_device(keys.as<uint64_t>(), keys_alt.as<uint64_t>(), sub_idx.ptr, idx_alt.ptr, 4, m,
                                       0, 64, &in_alt, st, nullptr));
        const uint64_t *K = in_alt ? keys_alt.as<uint64_t>() : keys.as<uint64_t>();
        const uint32_t *I = in_alt ? idx_alt.as<uint32_t>() : sub_idx.as<uint32_t>();
        GK_TRY(key2_scatter_device(K, I, slots.as<uint32_t>(), m, d_idx, d_flags, st));
        // the groups inside the list after this round, the new ranks, the members that are still tied
        GK_TRY(hf.alloc((size_t)((m + 15) & ~15ull), st));
        GK_TRY(key_flags_device(K, m, 0, hf.as<uint8_t>(), st));
        GK_TRY(gsub.alloc((size_t)m * 4, st));
        GK_TRY(head_positions_device(hf.as<uint8_t>(), I, m, gsub.as<uint32_t>(), nullptr, st));
        GK_TRY(gslot.alloc((size_t)m * 4, st));
        GK_TRY(subset_rank_update_device(slots.as<uint32_t>(), gsub.as<uint32_t>(), I, m, gslot.as<uint32_t>(),
                                         rank.as<uint32_t>(), st));
        GK_TRY(mf.alloc((size_t)((m + 15) & ~15ull), st));
        GK_TRY(gid_flags_device(gsub.as<uint32_t>(), m, mf.as<uint8_t>(), st));
        GK_TRY(clear_finished_multi_device(I, m, (const uint64_t *)ix->d_segs.ptr, (uint32_t)ix->h_segs.size(),
                                           ix->sba_len, h2, mf.as<uint8_t>(), st));
        uint64_t m2 = 0;
        GK_TRY(select_flagged(mf.as<uint8_t>(), m, kFlagMulti, 4, nullptr, nullptr, nullptr, nullptr, nullptr, &m2,
                              st));
        DeviceBuffer pos, new_slots, new_idx, new_gid;
        if (m2) {
            GK_TRY(pos.alloc((size_t)m2 * 4, st));
            GK_TRY(new_slots.alloc((size_t)m2 * 4, st));
            GK_TRY(new_idx.alloc((size_t)m2 * 4, st));
            GK_TRY(new_gid.alloc((size_t)m2 * 4, st));
            GK_TRY(select_flagged(mf.as<uint8_t>(), m, kFlagMulti, 4, pos.ptr, I, new_idx.ptr, gslot.ptr, new_gid.ptr,
                                  nullptr, st));
            GK_TRY(gather_u32_device(slots.as<uint32_t>(), pos.as<uint32_t>(), m2, new_slots.as<uint32_t>(), st));
        }
        GK_CUDA(cudaStreamSynchronize(st));  // the round's scratch is released in stream order after this
        slots.release(); sub_idx.release(); sub_gid.release();
        if (m2) {
            // adopt the new lists (move the pointers: DeviceBuffer has no move assignment)
            slots.ptr = new_slots.ptr; slots.bytes = new_slots.bytes; slots.stream = st; new_slots.ptr = nullptr;
            sub_idx.ptr = new_idx.ptr; sub_idx.bytes = new_idx.bytes; sub_idx.stream = st; new_idx.ptr = nullptr;
            sub_gid.ptr = new_gid.ptr; sub_gid.bytes = new_gid.bytes; sub_gid.stream = st; new_gid.ptr = nullptr;
        }
        m = m2;
        h = h2;
        if (levels) ++*levels;
    }
    return GK_OK;
}

// Word rounds: fixed-length k-mers longer than one key word where the ranks of other starts are not at hand -- a
// multi-GPU shard (the index holds one key range) -- or do not fit 32 bits (a byte array of 2^32 positions or
// more).  In: the windows sorted by their first h0 symbols, flags = groups of equal h0-prefixes.  The members of
// groups that are still tied are ordered by (group, 4-bit ranks of the next 8 symbols), read from the bytes, 8
// symbols per round, on the shrinking list of tied members.  Start indices of either width; positions inside the
// index are 32-bit (n < 2^32).
static int word_rounds(gk_index *ix, Owned &cur_idx, Owned &cur_flags, uint64_t n, uint64_t h0, uint64_t target,
                       int *levels, cudaStream_t st)
{
    if (h0 >= target || n < 2) return GK_OK;
    if (n >= (1ull << 32)) {
        set_error("k-mers longer than one key word: %llu windows in one index (limit 2^32 - 1)", (unsigned long long)n);
        return GK_ERR_UNSUPPORTED;
    }
    const int ib = ix->idx_bytes;
    void *d_idx = cur_idx.ptr;
    uint8_t *d_flags = (uint8_t *)cur_flags.ptr;
    DeviceBuffer gid, mflags;
    GK_TRY(mflags.alloc((size_t)((n + 15) & ~15ull), st));
    GK_TRY(multi_flags_device(d_flags, n, mflags.as<uint8_t>(), st));
    uint64_t m = 0;
    GK_TRY(select_flagged(mflags.as<uint8_t>(), n, kFlagMulti, 4, nullptr, nullptr, nullptr, nullptr, nullptr, &m, st));
    if (m == 0) return GK_OK;
    GK_TRY(gid.alloc((size_t)n * 4, st));
    GK_TRY(head_positions_device(d_flags, nullptr, n, gid.as<uint32_t>(), nullptr, st));
    DeviceBuffer slots, sub_idx, sub_gid;
    GK_TRY(slots.alloc((size_t)m * 4, st));
    GK_TRY(sub_idx.alloc((size_t)m * ib, st));
    GK_TRY(sub_gid.alloc((size_t)m * 4, st));
    GK_TRY(select_flagged(mflags.as<uint8_t>(), n, kFlagMulti, 4, slots.ptr, gid.ptr, sub_gid.ptr, nullptr, nullptr,
                          nullptr, st));
    GK_TRY(select_flagged(mflags.as<uint8_t>(), n, kFlagMulti, ib, nullptr, d_idx, sub_idx.ptr, nullptr, nullptr,
                          nullptr, st));
    gid.release();
    mflags.release();
    uint64_t h = h0;
    while (h < target && m > 0) {
        const uint64_t h2 = (h + 8 < target) ? h + 8 : target;
        DeviceBuffer keys, keys_alt, idx_alt, hf, gsub, gslot, mf;
        GK_TRY(keys.alloc((size_t)m * 8, st));
        GK_TRY(keys_alt.alloc((size_t)m * 8, st));
        GK_TRY(idx_alt.alloc((size_t)m * ib, st));
        GK_TRY(pair_keys_words_device(sub_idx.ptr, ib, sub_gid.as<uint32_t>(), m, ix->d_sba, ix->sba_len, h,
                                      (uint32_t)(h2 - h), keys.as<uint64_t>(), st));
        int in_alt = 0;
        GK_TRY(radix_sort_pairs_device(keys.as<uint64_t>(), keys_alt.as<uint64_t>(), sub_idx.ptr, idx_alt.ptr, ib, m, 0,
                                       64, &in_alt, st, nullptr));
        const uint64_t *K = in_alt ? keys_alt.as<uint64_t>() : keys.as<uint64_t>();
        const void *I = in_alt ? idx_alt.ptr : sub_idx.ptr;
        GK_TRY(key2_scatter_any_device(K, I, slots.as<uint32_t>(), m, ib, d_idx, d_flags, st));
        // the groups inside the list after this round and the members that are still tied
        GK_TRY(hf.alloc((size_t)((m + 15) & ~15ull), st));
        GK_TRY(key_flags_device(K, m, 0, hf.as<uint8_t>(), st));
        GK_TRY(gsub.alloc((size_t)m * 4, st));
        GK_TRY(head_positions_device(hf.as<uint8_t>(), nullptr, m, gsub.as<uint32_t>(), nullptr, st));
        GK_TRY(gslot.alloc((size_t)m * 4, st));
        GK_TRY(subset_rank_update_device(slots.as<uint32_t>(), gsub.as<uint32_t>(), nullptr, m, gslot.as<uint32_t>(),
                                         nullptr, st));
        GK_TRY(mf.alloc((size_t)((m + 15) & ~15ull), st));
        GK_TRY(gid_flags_device(gsub.as<uint32_t>(), m, mf.as<uint8_t>(), st));
        uint64_t m2 = 0;
        GK_TRY(select_flagged(mf.as<uint8_t>(), m, kFlagMulti, 4, nullptr, nullptr, nullptr, nullptr, nullptr, &m2, st));
        DeviceBuffer pos, new_slots, new_idx, new_gid;
        if (m2 && h2 < target) {
            GK_TRY(pos.alloc((size_t)m2 * 4, st));
            GK_TRY(new_slots.alloc((size_t)m2 * 4, st));
            GK_TRY(new_idx.alloc((size_t)m2 * ib, st));
            GK_TRY(new_gid.alloc((size_t)m2 * 4, st));
            GK_TRY(select_flagged(mf.as<uint8_t>(), m, kFlagMulti, 4, pos.ptr, gslot.ptr, new_gid.ptr, nullptr, nullptr,
                                  nullptr, st));
            GK_TRY(select_flagged(mf.as<uint8_t>(), m, kFlagMulti, ib, nullptr, I, new_idx.ptr, nullptr, nullptr,
                                  nullptr, st));
            GK_TRY(gather_u32_device(slots.as<uint32_t>(), pos.as<uint32_t>(), m2, new_slots.as<uint32_t>(), st));
        }
        GK_CUDA(cudaStreamSynchronize(st));  // the round's scratch is released in stream order after this
        slots.release(); sub_idx.release(); sub_gid.release();
        if (m2 && h2 < target) {
            slots.ptr = new_slots.ptr; slots.bytes = new_slots.bytes; slots.stream = st; new_slots.ptr = nullptr;
            sub_idx.ptr = new_idx.ptr; sub_idx.bytes = new_idx.bytes; sub_idx.stream = st; new_idx.ptr = nullptr;
            sub_gid.ptr = new_gid.ptr; sub_gid.bytes = new_gid.bytes; sub_gid.stream = st; new_gid.ptr = nullptr;
        }
        m = m2;
        h = h2;
        if (levels) ++*levels;
    }
    return GK_OK;
}

// Drop the windows shorter than min_len from a sorted order (variable-length mode sorts every start of
// every record so that all ranks exist).  Equal windows have equal lengths, so groups go or stay whole.
static int drop_short_windows(gk_index *ix, uint32_t min_len, Owned &cur_idx, Owned &cur_flags, uint64_t &n_cur,
                              cudaStream_t st)
{
    DeviceBuffer vflags;
    GK_TRY(vflags.alloc((size_t)((n_cur + 15) & ~15ull), st));
    GK_TRY(valid_flags_device((const uint32_t *)cur_idx.ptr, n_cur, (const uint64_t *)ix->d_segs.ptr,
                              (uint32_t)ix->h_segs.size(), ix->sba_len, min_len, vflags.as<uint8_t>(), st));
    Owned new_idx, new_flags;
    GK_TRY(new_idx.alloc((size_t)n_cur * 4, st));
    GK_TRY(new_flags.alloc((size_t)((n_cur + 15) & ~15ull), st));
    uint64_t n_new = 0;
    // groups go or stay whole, so the head flag of a kept window is still right: starts and head flags move
    // together in one ordered select (tiles that keep every window are copied straight through)
    GK_TRY(select_flagged_with_bytes(vflags.as<uint8_t>(), n_cur, kFlagPass, 4, cur_idx.ptr, new_idx.ptr,
                                     (const uint8_t *)cur_flags.ptr, (uint8_t *)new_flags.ptr, kFlagHead, &n_new,
                                     st));
    cur_idx.swap(new_idx);
    cur_flags.swap(new_flags);
    n_cur = n_new;
    return GK_OK;
}

extern "C" {

int gk_index_create(const uint8_t *d_sba, uint64_t sba_len, const uint64_t *h_seg_starts,
                    uint32_t n_seg, uint32_t min_kmer_len, uint32_t max_kmer_len, gk_index **out)
{
    if (!out || !d_sba || !h_seg_starts || n_seg == 0 || sba_len == 0) {
        set_error("gk_index_create: empty sequence collection or null pointer");
        return GK_ERR_ARG;
    }
    if (min_kmer_len < 1) {
        set_error("min_kmer_len (%u) must be greater than zero", min_kmer_len);
        return GK_ERR_ARG;
    }
    if (max_kmer_len != 0 && max_kmer_len < min_kmer_len) {
        set_error("max_kmer_len (%u) is less than min_kmer_len (%u)", max_kmer_len, min_kmer_len);
        return GK_ERR_ARG;
    }
    uint64_t n = 0;
    GK_TRY(kmer_count_host(h_seg_starts, n_seg, sba_len, min_kmer_len, &n));
    gk_index *ix = new gk_index();
    ix->d_sba = d_sba;
    ix->sba_len = sba_len;
    ix->h_segs.assign(h_seg_starts, h_seg_starts + n_seg);
    ix->min_len = min_kmer_len;
    ix->max_len = max_kmer_len;
    ix->n = n;
    // the reference refuses more than 2^32-1 k-mers (kmers.py:805-808); here the index widens
    ix->idx_bytes = (sba_len > 0xFFFFFFFFull) ? 8 : 4;
    {   // GK_FORCE_IDX64=1 (tests): exercise the 64-bit start-index path on small inputs
        const char *f = getenv("GK_FORCE_IDX64");
        if (f && *f && *f != '0') ix->idx_bytes = 8;
    }
    *out = ix;   // (the segment table goes to the device with the first call that brings a stream: ensure_segs)
    return GK_OK;
}

void gk_index_destroy(gk_index *ix) { delete ix; }

uint64_t gk_index_size(const gk_index *ix) { return ix ? ix->n : 0; }
int gk_index_idx_bytes(const gk_index *ix) { return ix ? ix->idx_bytes : 0; }
int gk_index_is_sorted(const gk_index *ix) { return ix && ix->sorted; }

int gk_index_set_indices(gk_index *ix, const void *h_idx, uint64_t n, int idx_bytes, int sorted,
                         void *stream)
{
    if (!ix || (!h_idx && n) || idx_bytes != ix->idx_bytes) {
        set_error("gk_index_set_indices: bad argument (index uses %d-byte starts)", ix ? ix->idx_bytes : 0);
        return GK_ERR_ARG;
    }
    cudaStream_t st = as_stream(stream);
    GK_TRY(ix->d_idx.alloc((size_t)n * idx_bytes, st));
    if (n) GK_CUDA(cudaMemcpyAsync(ix->d_idx.ptr, h_idx, (size_t)n * idx_bytes, cudaMemcpyHostToDevice, st));
    GK_CUDA(cudaStreamSynchronize(st));
    ix->n = n;
    ix->idx_ready = true;
    ix->user_idx = true;
    ix->sorted = sorted != 0;
    ix->flags_valid = false;
    ix->flags_mark_amb = false;
    ix->spectrum_valid = false;
    return GK_OK;
}

// Start indices assigned by the caller (gk_index_set_indices): order them by value, check that every one
// starts a k-mer of at least min_kmer_len symbols (the reference's sort raises through its validation
// otherwise, kmers.py:1716-1727), and tell whether they are exactly the full set of the index.
static int prepare_user_indices(gk_index *ix, Owned &list, bool *is_full_set, cudaStream_t st)
{
    const int ib = ix->idx_bytes;
    const uint64_t n = ix->n;
    *is_full_set = false;
    uint64_t n_full = 0;
    GK_TRY(kmer_count_host(ix->h_segs.data(), (uint32_t)ix->h_segs.size(), ix->sba_len, ix->min_len, &n_full));
    DeviceBuffer keys_a, keys_b, vals_b, report;
    const size_t pad = (size_t)((n + 1) & ~1ull);
    GK_TRY(keys_a.alloc(pad * 8, st));
    GK_TRY(keys_b.alloc(pad * 8, st));
    GK_TRY(vals_b.alloc((size_t)n * ib, st));
    GK_TRY(list.alloc((size_t)n * ib, st));
    GK_CUDA(cudaMemcpyAsync(list.ptr, ix->d_idx.ptr, (size_t)n * ib, cudaMemcpyDeviceToDevice, st));
    GK_TRY(widen_indices_device(list.ptr, ib, n, keys_a.as<uint64_t>(), st));
    int bits = 1;
    while (bits < 64 && (ix->sba_len >> bits)) ++bits;
    int in_alt = 0;
    GK_TRY(radix_sort_pairs_device(keys_a.as<uint64_t>(), keys_b.as<uint64_t>(), list.ptr, vals_b.ptr, ib, n, 0,
                                   ib == 4 ? (bits < 32 ? bits : 32) : 64, &in_alt, st, nullptr));
    if (in_alt) GK_CUDA(cudaMemcpyAsync(list.ptr, vals_b.ptr, (size_t)n * ib, cudaMemcpyDeviceToDevice, st));
    GK_TRY(report.alloc(16, st));
    GK_CUDA(cudaMemsetAsync(report.ptr, 0, 16, st));
    GK_TRY(check_starts_device(list.ptr, ib, n, (const uint64_t *)ix->d_segs.ptr, (uint32_t)ix->h_segs.size(),
                               ix->sba_len, ix->min_len, report.as<unsigned long long>(), st));
    unsigned long long h_report[2] = {0, 0};  // [0] starts that begin no k-mer of min_len symbols, [1] not ascending
    GK_CUDA(cudaMemcpyAsync(h_report, report.ptr, 16, cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    if (h_report[0]) {
        set_error("kmers compared were less than min_kmer_len (%u).  Was kmer_sba_start_indices "
                  "initialized correctly?", ix->min_len);
        return GK_ERR_INVALID_KMERS;
    }
    *is_full_set = (n == n_full && h_report[1] == 0);
    return GK_OK;
}

int gk_index_sort(gk_index *ix, gk_sort_stats *stats_out, void *stream)
{
    if (!ix) return GK_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    GK_TRY(ensure_segs(ix, st));
    gk_sort_stats stats;
    memset(&stats, 0, sizeof(stats));
    const uint64_t launches0 = gk_launch_count(0);
    trace_point("begin");
    EventTimer tm(st);
    StageMarks marks;
    const int t0 = tm.mark();
    // one device error word for every sort of this call, checked once at the end
    DeviceBuffer sort_err;
    GK_TRY(sort_err.alloc(4, st));
    GK_CUDA(cudaMemsetAsync(sort_err.ptr, 0, 4, st));
    struct ErrWordGuard {
        explicit ErrWordGuard(int *p) { set_deferred_sort_error_word(p); }
        ~ErrWordGuard() { set_deferred_sort_error_word(nullptr); }
    } err_guard(sort_err.as<int>());

    const bool fixed = ix->max_len != 0 && ix->max_len == ix->min_len;
    const uint32_t k = ix->min_len;
    const uint32_t max_len = ix->max_len;  // 0 = None
    // Modes that fit one key word always carry the class bit, so nothing has to wait for the alphabet scan:
    // it runs at the head of the stream and its counters come back with the sort's own.
    const bool one_word = (fixed && k <= 31) || (!fixed && max_len != 0 && max_len <= 31);
    DeviceBuffer alpha;
    unsigned long long h_alpha[3] = {0, 0, 0};
    const bool alpha_async = one_word && !ix->alphabet_known;
    ScopedEvent e_alpha0, e_alpha1;
    if (alpha_async) {
        // on the second stream, beside the pack kernel and the radix passes
        cudaStream_t side = nullptr;
        GK_TRY(side_stream(&side));
        GK_TRY(alpha.alloc(24, st));
        GK_TRY(e_alpha0.create());
        GK_TRY(e_alpha1.create());
        GK_CUDA(cudaEventRecord(e_alpha0.ev, st));
        GK_CUDA(cudaStreamWaitEvent(side, e_alpha0.ev, 0));
        GK_TRY(scan_alphabet_async(ix->d_sba, ix->sba_len, alpha.as<unsigned long long>(), side));
        GK_CUDA(cudaEventRecord(e_alpha1.ev, side));
    } else {
        GK_TRY(ensure_alphabet(ix, st));
    }
    auto check_separators = [&]() -> int {
        if (ix->n_sep != ix->h_segs.size() - 1) {
            // a '$' inside a record: the reference's sort raises through its validation
            set_error("kmers compared were less than min_kmer_len (%u).  Was kmer_sba_start_indices "
                      "initialized correctly?", ix->min_len);
            return GK_ERR_INVALID_KMERS;
        }
        return GK_OK;
    };
    if (!alpha_async) GK_TRY(check_separators());

    // start indices assigned by the caller: the reference sorts whatever the array holds (kmers.py:1648)
    Owned user_list;
    bool subset = false;
    if (ix->user_idx && ix->n > 0) {
        bool full = false;
        GK_TRY(prepare_user_indices(ix, user_list, &full, st));
        subset = !full;
    }
    uint64_t n_sort = ix->n;
    if (ix->user_idx && !subset)   // the full set (or an empty array): sort every window of the byte array
        GK_TRY(kmer_count_host(ix->h_segs.data(), (uint32_t)ix->h_segs.size(), ix->sba_len, ix->min_len, &n_sort));
    if (ix->user_idx && ix->n == 0) n_sort = 0;
    stats.n_windows = n_sort;

    // the new order is built in local buffers and adopted only when everything has succeeded
    Owned new_idx, new_flags;
    bool new_mark_amb = false;
    int t_ref0 = -1, t_ref1 = -1;
    // the group-size spectrum of the new order rides along with the sort's own synchronise
    SpectrumPending spectrum;
    SpectrumPending *spec = spectrum_enabled() ? &spectrum : nullptr;
    if (n_sort == 0) {
        // nothing to sort
    } else if (one_word) {
        const uint32_t key_len = fixed ? k : max_len;
        new_mark_amb = fixed;
        GK_TRY(sort_level1(ix, k, key_len, 1, n_sort, new_idx, new_flags, marks, tm, st,
                           alpha_async ? alpha.as<unsigned long long>() : nullptr, h_alpha,
                           subset ? user_list.ptr : nullptr, spec, alpha_async ? e_alpha1.ev : nullptr));
    } else if (fixed && k == 32 && ix->n_amb_letters == 0 && ix->n_bad == 0) {
        new_mark_amb = true;
        GK_TRY(sort_level1(ix, k, k, 0, n_sort, new_idx, new_flags, marks, tm, st, nullptr, nullptr,
                           subset ? user_list.ptr : nullptr, spec));
    } else {
        // Longer than one key word -- fixed k > 32, max_kmer_len > 31, or None (suffix order inside each
        // record).  Sort EVERY start of every record by its first 31 symbols, terminator-aware, so that
        // every position has a rank; double the compared length until the target is covered or nothing is
        // tied any more; finally drop the windows shorter than min_kmer_len.
        if (subset) {
            set_error("sorting an assigned subset of start indices is available for k-mers of one key word "
                      "(max_kmer_len <= 31) only");
            return GK_ERR_UNSUPPORTED;
        }
        if (ix->idx_bytes != 4 && !fixed) {
            set_error("the variable-length mode on a byte array of 2^32 or more positions is not available");
            return GK_ERR_UNSUPPORTED;
        }
        if (ix->idx_bytes != 4) {
            // fixed k > one key word with 64-bit start indices: the rank table of the doubling is 32-bit, so the
            // windows are sorted by their first 31 symbols and the tied ones by the following symbols, read from
            // the bytes (word_rounds)
            GK_TRY(sort_level1(ix, k, 31, 1, n_sort, new_idx, new_flags, marks, tm, st, nullptr, nullptr));
            t_ref0 = tm.mark();
            GK_TRY(word_rounds(ix, new_idx, new_flags, n_sort, 31, k, &marks.levels, st));
            t_ref1 = tm.mark();
        } else {
        uint64_t longest = 0;
        for (size_t r = 0; r < ix->h_segs.size(); ++r) {
            const uint64_t e = (r + 1 < ix->h_segs.size()) ? ix->h_segs[r + 1] - 1 : ix->sba_len;
            if (e - ix->h_segs[r] > longest) longest = e - ix->h_segs[r];
        }
        const uint64_t target = (max_len != 0 && max_len < longest) ? max_len : longest;
        uint64_t n_cur = 0;
        GK_TRY(kmer_count_host(ix->h_segs.data(), (uint32_t)ix->h_segs.size(), ix->sba_len, 1, &n_cur));
        GK_TRY(sort_level1(ix, 1, 31, 1, n_cur, new_idx, new_flags, marks, tm, st, nullptr, nullptr));
        t_ref0 = tm.mark();
        GK_TRY(doubling_rounds(ix, new_idx, new_flags, n_cur, 31, target, &marks.levels, st));
        if (ix->min_len > 1) GK_TRY(drop_short_windows(ix, ix->min_len, new_idx, new_flags, n_cur, st));
        t_ref1 = tm.mark();
        if (n_cur != n_sort) {
            set_error("doubling kept %llu windows, expected %llu", (unsigned long long)n_cur,
                      (unsigned long long)n_sort);
            return GK_ERR_INTERNAL;
        }
        }
    }
    const int t1 = tm.mark();
    if (n_sort > 0 && spec && !spectrum.armed)   // (the multi-level path: its flags are final only now)
        GK_TRY(spectrum.start((const uint8_t *)new_flags.ptr, n_sort, st));
    int h_sort_err = 0;
    GK_CUDA(cudaMemcpyAsync(&h_sort_err, sort_err.ptr, 4, cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    marks.main_sort.resolve();
    if (h_sort_err) {
        set_error("radix sort: decoupled look-back timed out");
        return GK_ERR_INTERNAL;
    }
    if (alpha_async) {
        ix->n_bad = h_alpha[0];
        ix->n_sep = h_alpha[1];
        ix->n_amb_letters = h_alpha[2];
        ix->alphabet_known = true;
        GK_TRY(check_separators());   // (the index keeps its previous order)
    }
    // adopt the result
    if (n_sort > 0) {
        ix->d_idx.swap(new_idx);
        ix->d_flags.swap(new_flags);
    }
    ix->n = n_sort;
    spectrum.finish(ix);
    ix->user_idx = subset;
    ix->idx_ready = true;
    ix->flags_mark_amb = new_mark_amb;
    ix->flags_valid = n_sort > 0;
    ix->flags_kmer_len = fixed ? k : ix->max_len;  // (None -> 0: such queries take the comparator path)
    ix->sorted = true;
    stats.pack_ms = tm.ms(marks.pack0, marks.pack1);
    stats.hist_ms = marks.main_sort.hist_ms;
    stats.sort_ms = marks.main_sort.passes_ms;
    stats.sort_passes = marks.main_sort.passes;
    stats.fixup_ms = tm.ms(marks.fix0, marks.fix1) + tm.ms(t_ref0, t_ref1);
    stats.key_bits = marks.key_bits;
    stats.levels = marks.levels;
    stats.n_ambiguous = marks.n_amb;
    stats.n_fragments = marks.n_frag;
    stats.refine_flags = marks.refine_flags;
    stats.total_ms = tm.ms(t0, t1);
    stats.gpu_launches = (int32_t)(gk_launch_count(0) - launches0);
    if (stats_out) *stats_out = stats;
    trace_point("end");
    trace_report("gk_index_sort");
    return GK_OK;
}

// Multi-GPU shard: adopt the (key, start) pairs this rank received from the exchange (keys made by gk_pack_keys /
// gk_pack_slice with key_len = valid_len = min_kmer_len, possibly relative to the start of this rank's key
// range), sort them and build the same state gk_index_sort leaves behind, for this rank's key range only.
// With a gathered fragment list the ambiguous windows of the range did not travel as pairs: they are generated
// here from the fragments (n_ambiguous of them, appended behind the n_pure received pairs) and ordered by the
// fragment path.  The pair buffers are scratch; the sorted starts land in the index's own buffer.
int gk_index_sort_shard(gk_index *ix, uint64_t *d_keys, uint64_t *d_keys_alt, void *d_idx, void *d_idx_alt,
                        uint64_t n_pure, uint64_t n_ambiguous, int class_bit, int key_bits,
                        const void *d_frag_gathered, const uint64_t *d_frag_counts, uint32_t n_sources,
                        uint64_t frag_capacity, int frag_presorted, uint64_t key_lo, uint64_t key_hi, const int *d_err,
                        gk_sort_stats *stats_out, void *stream)
{
    const uint64_t n_local = n_pure + n_ambiguous;
    if (!ix || (n_local && (!d_keys || !d_keys_alt || !d_idx || !d_idx_alt))) {
        set_error("gk_index_sort_shard: null buffer");
        return GK_ERR_ARG;
    }
    const uint32_t k = ix->min_len;
    if (ix->max_len != ix->min_len) {
        set_error("gk_index_sort_shard: fixed-length k-mers only");
        return GK_ERR_UNSUPPORTED;
    }
    // k-mers longer than one key word: the pairs carry the first 31 symbols (gk_pack_slice), the rest is compared
    // from the bytes in word rounds after the sort
    const bool long_k = k > 31 && !(k == 32 && !class_bit);
    if (long_k && !class_bit) {
        set_error("gk_index_sort_shard: k-mers longer than one key word need class-bit keys");
        return GK_ERR_UNSUPPORTED;
    }
    const uint32_t key_len = long_k ? 31u : k;
    const int full_bits = 2 * (int)key_len + (class_bit ? 2 : 0);
    if (key_bits <= 0 || key_bits > full_bits) key_bits = full_bits;
    const bool with_frag = d_frag_gathered != nullptr && class_bit && n_sources > 0 && frag_capacity > 0;
    if (!with_frag && n_ambiguous) {
        set_error("gk_index_sort_shard: ambiguous windows to generate but no fragment list");
        return GK_ERR_ARG;
    }
    cudaStream_t st = as_stream(stream);
    GK_TRY(ensure_segs(ix, st));
    gk_sort_stats stats;
    memset(&stats, 0, sizeof(stats));
    const uint64_t launches0 = gk_launch_count(0);
    EventTimer tm(st);
    StageMarks marks;
    const int t0 = tm.mark();
    const int ib = ix->idx_bytes;
    DeviceBuffer sort_err;
    GK_TRY(sort_err.alloc(4, st));
    GK_CUDA(cudaMemsetAsync(sort_err.ptr, 0, 4, st));
    struct ErrWordGuard {
        explicit ErrWordGuard(int *p) { set_deferred_sort_error_word(p); }
        ~ErrWordGuard() { set_deferred_sort_error_word(nullptr); }
    } err_guard(sort_err.as<int>());

    Owned new_idx, new_flags;
    DeviceBuffer counters, frag_mem, frag_off;
    FragSorted fs;
    PackedPairs pp;
    SpectrumPending spectrum;
    if (spectrum_enabled() && !long_k) pp.spectrum = &spectrum;
    if (n_local) {
        GK_TRY(new_idx.alloc((size_t)n_local * ib, st));
        GK_TRY(new_flags.alloc((size_t)((n_local + 15) & ~15ull), st));
        GK_TRY(counters.alloc(kCounterBytes, st));
        GK_CUDA(cudaMemsetAsync(counters.ptr, 0, (size_t)kCounterWords * 8, st));
        pp.n = n_local;
        pp.key_len = key_len;
        pp.key_bits = key_bits;
        pp.class_bit = class_bit;
        pp.keys_a = d_keys; pp.keys_b = d_keys_alt;
        pp.idx_a = d_idx; pp.idx_b = d_idx_alt;
        pp.idx_final = new_idx.ptr;
        pp.d_counters = counters.as<unsigned long long>();
        pp.d_extra_err = d_err;
        pp.fs = &fs;
        pp.start_bits = 1;
        while (pp.start_bits < 64 && (ix->sba_len >> pp.start_bits)) ++pp.start_bits;
        ScopedEvent e_ready;
        GK_TRY(e_ready.create());
        marks.pack0 = tm.mark();
        if (with_frag) {
            // the fragments of this key range, keys relative like the received pairs; then their windows as
            // placeholder pairs behind the received ones
            uint64_t cap = (uint64_t)n_sources * frag_capacity;
            if (cap > kFragCapacity) cap = kFragCapacity;
            GK_TRY(frag_off.alloc((size_t)(cap + 1) * 8, st));
            pp.want_frag = fragments_enabled() || n_ambiguous > 0;
            if (frag_presorted) {
                // every rank sorted its own list (gk_frag_sort_local): merge the lists instead of sorting their union
                GK_TRY(frag_merge_device(d_frag_gathered, reinterpret_cast<const unsigned long long *>(d_frag_counts),
                                         n_sources, frag_capacity, key_lo, key_hi, cap, pp.d_counters + 2, fs, &pp.frag,
                                         st));
                pp.fs_ready = true;
            } else {
                GK_TRY(frag_mem.alloc((size_t)cap * 36, st));
                uint64_t *base = frag_mem.as<uint64_t>();
                pp.frag.key = base; pp.frag.w0 = base + cap; pp.frag.w1 = base + 2 * cap;
                pp.frag.start = base + 3 * cap;
                pp.frag.count = reinterpret_cast<uint32_t *>(base + 4 * cap);
                pp.frag.counter = pp.d_counters + 2;
                pp.frag.capacity = cap;
                GK_TRY(fs.reserve(cap < kFragReserve ? cap : kFragReserve, st));
                GK_TRY(frag_filter_device(d_frag_gathered, reinterpret_cast<const unsigned long long *>(d_frag_counts),
                                          n_sources, frag_capacity, key_lo, key_hi, pp.frag, st));
            }
            GK_TRY(frag_placeholders_device(pp.frag, frag_off.as<unsigned long long>(), n_pure, n_ambiguous, d_keys,
                                            d_idx, ib, reinterpret_cast<int *>(pp.d_counters + 3), st));
            // d_counters[0] = ambiguous windows of the shard: known to the caller, the expand step checks it
            GK_CUDA(cudaMemcpyAsync(pp.d_counters, &n_ambiguous, 8, cudaMemcpyHostToDevice, st));
            pp.producer_counted_amb = true;
        } else {
            pp.producer_counted_amb = false;   // ambiguous windows are among the pairs: the flags pass counts them
        }
        marks.pack1 = tm.mark();
        GK_CUDA(cudaEventRecord(e_ready.ev, st));
        pp.e_producer = e_ready.ev;
        GK_TRY(sort_packed_pairs(ix, pp, (uint8_t *)new_flags.ptr, marks, tm, st));
    }
    int levels = 1;
    int t_ref0 = -1, t_ref1 = -1;
    if (long_k && n_local) {
        t_ref0 = tm.mark();
        GK_TRY(word_rounds(ix, new_idx, new_flags, n_local, key_len, k, &levels, st));
        t_ref1 = tm.mark();
        if (spectrum_enabled()) GK_TRY(spectrum.start((const uint8_t *)new_flags.ptr, n_local, st));
    }
    const int t1 = tm.mark();
    int h_sort_err = 0;
    GK_CUDA(cudaMemcpyAsync(&h_sort_err, sort_err.ptr, 4, cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    marks.main_sort.resolve();
    if (h_sort_err) {
        set_error("radix sort: decoupled look-back timed out");
        return GK_ERR_INTERNAL;
    }
    if (marks.refine_flags & (8u << 8)) {
        set_error("gk_index_sort_shard: the fragments of this key range do not add up to %llu ambiguous windows",
                  (unsigned long long)n_ambiguous);
        return GK_ERR_INTERNAL;
    }
    ix->d_idx.swap(new_idx);
    ix->d_flags.swap(new_flags);
    ix->n = n_local;
    spectrum.finish(ix);
    ix->user_idx = false;
    ix->idx_ready = true;
    ix->flags_mark_amb = true;
    ix->flags_valid = n_local > 0;
    ix->flags_kmer_len = k;
    ix->sorted = true;
    stats.pack_ms = tm.ms(marks.pack0, marks.pack1);
    stats.hist_ms = marks.main_sort.hist_ms;
    stats.sort_ms = marks.main_sort.passes_ms;
    stats.sort_passes = marks.main_sort.passes;
    stats.fixup_ms = tm.ms(marks.fix0, marks.fix1) + tm.ms(t_ref0, t_ref1);
    stats.key_bits = key_bits;
    stats.levels = levels;
    stats.n_windows = n_local;
    stats.n_ambiguous = marks.n_amb;
    stats.n_fragments = marks.n_frag;
    stats.refine_flags = marks.refine_flags;
    stats.total_ms = tm.ms(t0, t1);
    stats.gpu_launches = (int32_t)(gk_launch_count(0) - launches0);
    if (stats_out) *stats_out = stats;
    return GK_OK;
}

int gk_index_sort_pairs(gk_index *ix, uint64_t *d_keys, uint64_t *d_keys_alt, void *d_idx, void *d_idx_alt,
                        uint64_t n_local, int class_bit, gk_sort_stats *stats_out, void *stream)
{
    return gk_index_sort_shard(ix, d_keys, d_keys_alt, d_idx, d_idx_alt, n_local, 0, class_bit, 0, nullptr, nullptr, 0,
                               0, 0, 0, 0, nullptr, stats_out, stream);
}

/* Sort a fragment list (gk_pack_slice layout, n_frag <= frag_capacity entries used) by window, ties by start,
 * into d_frag_sorted (same layout and capacity).  No synchronise. */
int gk_frag_sort_local(const void *d_frag, uint64_t frag_capacity, uint64_t n_frag, uint32_t kmer_len,
                       uint64_t sba_len, void *d_frag_sorted, void *stream)
{
    if (!d_frag || !d_frag_sorted || frag_capacity == 0 || n_frag > frag_capacity || kmer_len < 1) {
        set_error("gk_frag_sort_local: bad argument");
        return GK_ERR_ARG;
    }
    auto view = [&](const void *p) {
        FragOut f;
        uint64_t *b = reinterpret_cast<uint64_t *>(const_cast<void *>(p));
        f.key = b; f.w0 = b + frag_capacity; f.w1 = b + 2 * frag_capacity; f.start = b + 3 * frag_capacity;
        f.count = reinterpret_cast<uint32_t *>(b + 4 * frag_capacity);
        f.capacity = frag_capacity;
        return f;
    };
    int start_bits = 1;
    while (start_bits < 64 && (sba_len >> start_bits)) ++start_bits;
    const uint32_t key_len = kmer_len > 31 ? 31u : kmer_len;   // as gk_pack_slice
    return frag_sort_copy_device(view(d_frag), n_frag, key_len, start_bits, view(d_frag_sorted), as_stream(stream));
}

int gk_index_device_indices(gk_index *ix, const void **d_idx_out, void *stream)
{
    if (!ix || !d_idx_out) return GK_ERR_ARG;
    GK_TRY(ensure_indices(ix, as_stream(stream)));
    *d_idx_out = ix->d_idx.ptr;
    return GK_OK;
}

int gk_index_copy_indices(gk_index *ix, void *h_dst, void *stream)
{
    if (!ix || (!h_dst && ix->n)) return GK_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    GK_TRY(ensure_indices(ix, st));
    if (ix->n)
        GK_CUDA(cudaMemcpyAsync(h_dst, ix->d_idx.ptr, (size_t)ix->n * ix->idx_bytes,
                                cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    return GK_OK;
}

// head flags of the current order for kmer_len (cached when they come from the sort)
static int flags_for(gk_index *ix, uint32_t kmer_len, const uint8_t **d_flags, DeviceBuffer &scratch,
                     cudaStream_t st)
{
    if (ix->flags_valid && ix->flags_kmer_len == kmer_len && kmer_len != 0) {
        *d_flags = (const uint8_t *)ix->d_flags.ptr;
        return GK_OK;
    }
    GK_TRY(scratch.alloc((size_t)((ix->n + 15) & ~15ull), st));
    GK_TRY(sba_flags_device(ix->d_sba, ix->sba_len, ix->d_idx.ptr, ix->idx_bytes, ix->n, kmer_len,
                            nullptr, 0, scratch.as<uint8_t>(), st));
    *d_flags = scratch.as<uint8_t>();
    return GK_OK;
}

// shared body: h_hist_out must already be zero; only bins [0, last_hist_top_bin()] are written
static int index_group_counts(gk_index *ix, uint32_t kmer_len, const gk_filter *filter, uint64_t min_group,
                              uint64_t max_group, uint64_t max_bin, int64_t *h_hist_out, int64_t *h_total_out,
                              void *stream)
{
    if (!ix) return GK_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    GK_TRY(ensure_segs(ix, st));
    gk_filter keep_all = {GK_FILTER_KEEP_ALL, 0, 0, 0};
    const gk_filter f = filter ? *filter : keep_all;
    if (h_total_out) *h_total_out = 0;
    set_last_hist_single(0, 0);
    if (ix->n == 0) return GK_OK;
    GK_TRY(ensure_indices(ix, st));
    const int ib = ix->idx_bytes;
    const uint64_t n = ix->n;

    // fast path: sorted, same kmer_len as the sort, filter uniform over a group of equal k-mers
    const bool cached = ix->sorted && ix->flags_valid && ix->flags_kmer_len == kmer_len && kmer_len != 0;
    if (cached && f.id == GK_FILTER_KEEP_ALL && ix->spectrum_valid) {
        // the sort left the whole group-size spectrum on the host: any limits and any bin clamp from there
        if (min_group < 1) { set_error("min_group_size (%llu) must be >= 1", (unsigned long long)min_group); return GK_ERR_ARG; }
        if (max_group != 0 && max_group < min_group) { set_error("max_group_size must be >= min_group_size"); return GK_ERR_ARG; }
        if (max_bin < 1) { set_error("max_counts_bin (%llu) must be >= 1", (unsigned long long)max_bin); return GK_ERR_ARG; }
        std::vector<unsigned long long> pairs;
        uint64_t total = 0, top = 0;
        for (const auto &sc : ix->spectrum) {
            const uint64_t size = sc.first, count = sc.second;
            if (size < min_group || (max_group != 0 && size > max_group)) continue;
            const uint64_t bin = size < max_bin ? size : max_bin;
            total += size * count;
            top = bin > top ? bin : top;
            if (!pairs.empty() && pairs[pairs.size() - 2] == bin) pairs.back() += count;
            else { pairs.push_back(bin); pairs.push_back(count); }
            if (h_hist_out) h_hist_out[bin] += (int64_t)count;
        }
        if (h_total_out) *h_total_out = (int64_t)total;
        set_last_hist_pairs(pairs, top);
        return GK_OK;
    }
    const bool no_amb_same_len =
        f.id == GK_FILTER_NO_AMBIGUOUS && (uint64_t)f.p0 == kmer_len && ix->flags_mark_amb;
    if (cached && (f.id == GK_FILTER_KEEP_ALL || no_amb_same_len))
        return flag_group_hist_device((const uint8_t *)ix->d_flags.ptr, n, no_amb_same_len ? kFlagAmb : 0,
                                      min_group, max_group, max_bin, h_hist_out, h_total_out, nullptr, st);

    // general path: the reference's walk, data-parallel.  Drop k-mers that fail the filter
    // (order preserved), then compare neighbours with the '$'-terminated comparator.
    DeviceBuffer pass_flags, kept;
    const void *d_list = ix->d_idx.ptr;
    uint64_t m = n;
    if (f.id != GK_FILTER_KEEP_ALL) {
        GK_TRY(pass_flags.alloc((size_t)((n + 15) & ~15ull), st));
        GK_TRY(filter_flags_device(ix->d_sba, ix->sba_len, ix->d_idx.ptr, ib, n, f,
                                   pass_flags.as<uint8_t>(), st));
        GK_TRY(kept.alloc((size_t)n * ib, st));
        GK_TRY(select_flagged(pass_flags.as<uint8_t>(), n, kFlagPass, ib, nullptr, ix->d_idx.ptr, kept.ptr,
                              nullptr, nullptr, &m, st));
        d_list = kept.ptr;
    }
    if (m == 0) return GK_OK;
    if (!ix->sorted) {
        // get_kmer_count on unsorted data: every passing k-mer is its own group (kmers.py:1061-1064)
        if (min_group > 1) return GK_OK;
        if (h_hist_out) h_hist_out[1 < max_bin ? 1 : max_bin] = (int64_t)m;
        if (h_total_out) *h_total_out = (int64_t)m;
        set_last_hist_single(1 < max_bin ? 1 : max_bin, m);
        return GK_OK;
    }
    DeviceBuffer flags;
    GK_TRY(flags.alloc((size_t)((m + 15) & ~15ull), st));
    GK_TRY(sba_flags_device(ix->d_sba, ix->sba_len, d_list, ib, m, kmer_len, nullptr, 0,
                            flags.as<uint8_t>(), st));
    return flag_group_hist_device(flags.as<uint8_t>(), m, 0, min_group, max_group, max_bin, h_hist_out,
                                  h_total_out, nullptr, st);
}

int gk_index_group_counts(gk_index *ix, uint32_t kmer_len, const gk_filter *filter, uint64_t min_group,
                          uint64_t max_group, uint64_t max_bin, int64_t *h_hist_out, int64_t *h_total_out,
                          void *stream)
{
    if (h_hist_out) memset(h_hist_out, 0, (size_t)(max_bin + 1) * 8);
    return index_group_counts(ix, kmer_len, filter, min_group, max_group, max_bin, h_hist_out, h_total_out,
                              stream);
}

int gk_index_group_counts_zeroed(gk_index *ix, uint32_t kmer_len, const gk_filter *filter, uint64_t min_group,
                                 uint64_t max_group, uint64_t max_bin, int64_t *h_hist_zeroed,
                                 int64_t *h_total_out, uint64_t *h_top_bin_out, void *stream)
{
    if (h_top_bin_out) *h_top_bin_out = 0;
    const int rc = index_group_counts(ix, kmer_len, filter, min_group, max_group, max_bin, h_hist_zeroed,
                                      h_total_out, stream);
    if (rc == GK_OK && h_top_bin_out) *h_top_bin_out = last_hist_top_bin();
    return rc;
}

int gk_index_group_counts_sparse(gk_index *ix, uint32_t kmer_len, const gk_filter *filter, uint64_t min_group,
                                 uint64_t max_group, uint64_t max_bin, uint64_t *h_bins_out,
                                 int64_t *h_counts_out, uint64_t capacity, uint64_t *h_n_pairs_out,
                                 int64_t *h_total_out, void *stream)
{
    if (!h_n_pairs_out) return GK_ERR_ARG;
    *h_n_pairs_out = 0;
    GK_TRY(index_group_counts(ix, kmer_len, filter, min_group, max_group, max_bin, nullptr, h_total_out, stream));
    std::vector<std::pair<unsigned long long, unsigned long long>> pairs;
    const std::vector<unsigned long long> &flat = last_hist_pairs();
    for (size_t i = 0; i + 1 < flat.size(); i += 2) pairs.emplace_back(flat[i], flat[i + 1]);
    std::sort(pairs.begin(), pairs.end());
    *h_n_pairs_out = pairs.size();
    if (pairs.size() > capacity || (pairs.size() && (!h_bins_out || !h_counts_out))) {
        set_error("gk_index_group_counts_sparse: %llu occupied bins do not fit the capacity %llu",
                  (unsigned long long)pairs.size(), (unsigned long long)capacity);
        return GK_ERR_ARG;
    }
    for (size_t i = 0; i < pairs.size(); ++i) {
        h_bins_out[i] = pairs[i].first;
        h_counts_out[i] = (int64_t)pairs[i].second;
    }
    return GK_OK;
}

int gk_index_groups(gk_index *ix, uint32_t kmer_len, uint64_t *h_n_groups, uint64_t *h_offsets_out,
                    uint64_t *h_sizes_out, void *stream)
{
    if (!ix || !h_n_groups) return GK_ERR_ARG;
    if (!ix->sorted) {
        set_error("The kmers must be sorted when calling gk_index_groups");
        return GK_ERR_STATE;
    }
    cudaStream_t st = as_stream(stream);
    GK_TRY(ensure_segs(ix, st));
    *h_n_groups = 0;
    if (ix->n == 0) return GK_OK;
    const uint64_t n = ix->n;
    const uint8_t *d_flags = nullptr;
    DeviceBuffer scratch, offsets;
    GK_TRY(flags_for(ix, kmer_len, &d_flags, scratch, st));
    GK_TRY(offsets.alloc((size_t)n * 8, st));
    uint64_t n_groups = 0;
    GK_TRY(select_flagged(d_flags, n, kFlagHead, 8, offsets.ptr, nullptr, nullptr, nullptr, nullptr,
                          &n_groups, st));
    *h_n_groups = n_groups;
    if (h_offsets_out && n_groups) {
        GK_CUDA(cudaMemcpyAsync(h_offsets_out, offsets.ptr, (size_t)n_groups * 8, cudaMemcpyDeviceToHost, st));
        GK_CUDA(cudaStreamSynchronize(st));
        if (h_sizes_out)
            for (uint64_t g = 0; g < n_groups; ++g)
                h_sizes_out[g] = ((g + 1 < n_groups) ? h_offsets_out[g + 1] : n) - h_offsets_out[g];
    }
    return GK_OK;
}

// Group table of the k-mers that pass `filter`, in the current order (kmers.py:523-648: a k-mer that fails
// is skipped and the next passing one is compared with the previous PASSING one).  Two-call protocol: with
// NULL outputs only the counts are returned.  h_kept_pos_out[j] = position in the (sorted) index of the
// j-th passing k-mer; offsets index that list.  On an unsorted index every passing k-mer is its own group.
int gk_index_groups_filtered(gk_index *ix, uint32_t kmer_len, const gk_filter *filter, uint64_t *h_n_kept,
                             uint64_t *h_n_groups, uint64_t *h_kept_pos_out, uint64_t *h_offsets_out,
                             uint64_t *h_sizes_out, void *stream)
{
    if (!ix || !filter || !h_n_kept || !h_n_groups) return GK_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    GK_TRY(ensure_segs(ix, st));
    *h_n_kept = *h_n_groups = 0;
    if (ix->n == 0) return GK_OK;
    GK_TRY(ensure_indices(ix, st));
    const int ib = ix->idx_bytes;
    const uint64_t n = ix->n;
    DeviceBuffer pass_flags, kept, kept_pos, flags, offsets;
    GK_TRY(pass_flags.alloc((size_t)((n + 15) & ~15ull), st));
    GK_TRY(filter_flags_device(ix->d_sba, ix->sba_len, ix->d_idx.ptr, ib, n, *filter, pass_flags.as<uint8_t>(), st));
    GK_TRY(kept.alloc((size_t)n * ib, st));
    GK_TRY(kept_pos.alloc((size_t)n * ib, st));
    uint64_t m = 0;
    GK_TRY(select_flagged(pass_flags.as<uint8_t>(), n, kFlagPass, ib, kept_pos.ptr, ix->d_idx.ptr, kept.ptr, nullptr,
                          nullptr, &m, st));
    *h_n_kept = m;
    if (m == 0) return GK_OK;
    uint64_t n_groups = m;
    if (ix->sorted) {
        GK_TRY(flags.alloc((size_t)((m + 15) & ~15ull), st));
        GK_TRY(sba_flags_device(ix->d_sba, ix->sba_len, kept.ptr, ib, m, kmer_len, nullptr, 0, flags.as<uint8_t>(),
                                st));
        GK_TRY(offsets.alloc((size_t)m * 8, st));
        GK_TRY(select_flagged(flags.as<uint8_t>(), m, kFlagHead, 8, offsets.ptr, nullptr, nullptr, nullptr, nullptr,
                              &n_groups, st));
    }
    *h_n_groups = n_groups;
    if (h_kept_pos_out) {
        std::vector<unsigned char> tmp((size_t)m * ib);
        GK_CUDA(cudaMemcpyAsync(tmp.data(), kept_pos.ptr, (size_t)m * ib, cudaMemcpyDeviceToHost, st));
        GK_CUDA(cudaStreamSynchronize(st));
        for (uint64_t j = 0; j < m; ++j)
            h_kept_pos_out[j] = ib == 4 ? (uint64_t)reinterpret_cast<const uint32_t *>(tmp.data())[j]
                                        : reinterpret_cast<const uint64_t *>(tmp.data())[j];
    }
    if (h_offsets_out) {
        if (ix->sorted) {
            GK_CUDA(cudaMemcpyAsync(h_offsets_out, offsets.ptr, (size_t)n_groups * 8, cudaMemcpyDeviceToHost, st));
            GK_CUDA(cudaStreamSynchronize(st));
        } else {
            for (uint64_t g = 0; g < n_groups; ++g) h_offsets_out[g] = g;
        }
        if (h_sizes_out)
            for (uint64_t g = 0; g < n_groups; ++g)
                h_sizes_out[g] = ((g + 1 < n_groups) ? h_offsets_out[g + 1] : m) - h_offsets_out[g];
    }
    return GK_OK;
}

int gk_index_verify(gk_index *ix, uint32_t kmer_len, uint64_t *h_report8, uint32_t *d_seen_bitmap, void *stream)
{
    if (!ix || !h_report8) return GK_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    GK_TRY(ensure_segs(ix, st));
    GK_TRY(ensure_indices(ix, st));
    const bool cached = ix->sorted && ix->flags_valid && ix->flags_kmer_len == kmer_len && kmer_len != 0;
    return verify_order_device(ix->d_sba, ix->sba_len, ix->d_idx.ptr, ix->idx_bytes, ix->n, kmer_len,
                               (const uint64_t *)ix->d_segs.ptr, (uint32_t)ix->h_segs.size(), ix->min_len,
                               cached ? (const uint8_t *)ix->d_flags.ptr : nullptr, h_report8, d_seen_bitmap, st);
}

int gk_sort_count_host(const uint8_t *h_sba, uint64_t sba_len, const uint64_t *h_seg_starts,
                       uint32_t n_seg, uint32_t kmer_len, int strands, int idx_bytes, void *h_idx_out,
                       uint64_t max_bin, int64_t *h_hist_out, int64_t *h_total_out,
                       uint64_t *h_n_kmers_out, gk_sort_stats *stats_out)
{
    if (!h_sba || !h_seg_starts || n_seg == 0 || sba_len == 0 || (strands != 0 && strands != 2)) {
        set_error("gk_sort_count_host: bad argument");
        return GK_ERR_ARG;
    }
    cudaStream_t st = nullptr;
    const uint64_t dev_len = strands == 2 ? 2 * sba_len + 1 : sba_len;
    Owned d_fwd, d_both;
    GK_TRY(d_fwd.alloc((size_t)sba_len));
    GK_CUDA(cudaMemcpyAsync(d_fwd.ptr, h_sba, (size_t)sba_len, cudaMemcpyHostToDevice, st));
    std::vector<uint64_t> segs(h_seg_starts, h_seg_starts + n_seg);
    const uint8_t *d_sba = (const uint8_t *)d_fwd.ptr;
    if (strands == 2) {
        GK_TRY(d_both.alloc((size_t)dev_len));
        GK_TRY(gk_sba_both_strands((const uint8_t *)d_fwd.ptr, sba_len, (uint8_t *)d_both.ptr, st));
        d_sba = (const uint8_t *)d_both.ptr;
        // mirrored segment table of the reverse strand (sequence_collection.py:905-928)
        for (uint32_t s = 0; s < n_seg; ++s) {
            const uint32_t src = n_seg - 1 - s;
            const uint64_t end = (src + 1 < n_seg) ? h_seg_starts[src + 1] - 2 : sba_len - 1;
            segs.push_back(sba_len + 1 + (sba_len - 1 - end));
        }
    }
    gk_index *ix = nullptr;
    GK_TRY(gk_index_create(d_sba, dev_len, segs.data(), (uint32_t)segs.size(), kmer_len, kmer_len, &ix));
    int rc = GK_OK;
    if (idx_bytes != 0 && idx_bytes != ix->idx_bytes && h_idx_out) {
        set_error("gk_sort_count_host: this input needs %d-byte start indices", ix->idx_bytes);
        rc = GK_ERR_ARG;
    }
    if (rc == GK_OK) rc = gk_index_sort(ix, stats_out, st);
    if (rc == GK_OK && (h_hist_out || h_total_out))
        rc = gk_index_group_counts(ix, kmer_len, nullptr, 1, 0, max_bin, h_hist_out, h_total_out, st);
    if (rc == GK_OK && h_idx_out) rc = gk_index_copy_indices(ix, h_idx_out, st);
    if (rc == GK_OK && h_n_kmers_out) *h_n_kmers_out = ix->n;
    gk_index_destroy(ix);
    return rc;
}

}  // extern "C"
