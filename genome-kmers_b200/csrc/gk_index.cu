// gk_index.cu -- the device-resident k-mer index: the state behind the reference's `Kmers`
// object (kmers.py:651-760) and the two seams the reference hands to numba:
//   gk_index_sort          <- Kmers.sort()                    kmers.py:1624-1652
//   gk_index_group_counts  <- get_kmer_group_size_hist()       kmers.py:454-520 (:1072, :1166)
// Host-side orchestration only; the kernels live in gk_pack.cu / gk_sort.cu / gk_group.cu.
#include <vector>

#include "gk_common.cuh"

namespace gk {

// kernels / device drivers from the other translation units
int kmer_count_host(const uint64_t *, uint32_t, uint64_t, uint32_t, uint64_t *);
int init_indices_device(const uint64_t *, uint32_t, uint32_t, uint64_t, int, void *, cudaStream_t);
int pack_keys_device(const uint8_t *, uint64_t, const uint64_t *, uint32_t, uint32_t, uint32_t, int,
                     uint64_t, uint64_t, uint64_t, uint64_t *, int, void *, unsigned long long *,
                     cudaStream_t);
int pack4_gather_device(const uint8_t *, uint64_t, const void *, int, uint64_t, uint32_t, uint32_t,
                        uint64_t *, cudaStream_t);
int radix_sort_pairs_device(uint64_t *, uint64_t *, void *, void *, int, uint64_t, int, int, int *,
                            cudaStream_t, SortTiming *);
int key_flags_device(const uint64_t *, uint64_t, int, uint8_t *, cudaStream_t);
int sba_flags_device(const uint8_t *, uint64_t, const void *, int, uint64_t, uint32_t, const void *,
                     uint8_t, uint8_t *, cudaStream_t);
int scatter_device(const void *, const void *, uint64_t, int, void *, cudaStream_t);
int group_hist_device(const void *, int, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, int64_t *,
                      int64_t *, int64_t *, cudaStream_t);
int group_hist_masked_device(const void *, int, uint64_t, uint64_t, const uint8_t *, uint8_t, uint64_t,
                             uint64_t, uint64_t, int64_t *, int64_t *, cudaStream_t);
int filter_flags_device(const uint8_t *, uint64_t, const void *, int, uint64_t, const gk_filter &,
                        uint8_t *, cudaStream_t);
template <typename PosT, typename PayT>
int select_flagged_device(const uint8_t *, uint64_t, uint8_t, PosT *, const PayT *, PayT *,
                          uint64_t *, cudaStream_t);

constexpr uint8_t kFlagHead = 1;
constexpr uint8_t kFlagAmb = 2;
constexpr uint8_t kFlagPass = 4;

// index-lifetime device allocation; stream-ordered like the scratch buffers so that creating and
// destroying an index per query does not pay a device-wide synchronising cudaFree
struct Owned {
    void *ptr = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;
    ~Owned() { reset(); }
    void reset()
    {
        if (ptr) cudaFreeAsync(ptr, stream);
        ptr = nullptr;
        bytes = 0;
    }
    int alloc(size_t n, cudaStream_t st = nullptr)
    {
        reset();
        stream = st;
        if (n == 0) return GK_OK;
        GK_TRY(ensure_pool_configured());
        GK_CUDA(cudaMallocAsync(&ptr, n, st));
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
    cudaEvent_t ev[8];
    int n = 0;
    cudaStream_t st;
    explicit EventTimer(cudaStream_t s) : st(s) {}
    ~EventTimer() { for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]); }
    int mark()
    {
        if (n >= 8) return n - 1;
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
    Owned d_segs;
    uint32_t min_len = 1, max_len = 0;  // max_len 0 == None
    int idx_bytes = 4;
    uint64_t n = 0;
    Owned d_idx;                        // start indices (init order until sorted)
    bool idx_ready = false;
    bool sorted = false;
    Owned d_flags;                      // head/amb flags of the sorted order for flags_kmer_len
    bool flags_valid = false;
    uint32_t flags_kmer_len = 0;
    bool alphabet_known = false;
    uint64_t n_bad = 0, n_sep = 0, n_amb_letters = 0;
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
    if (ix->idx_ready) return GK_OK;
    GK_TRY(ix->d_idx.alloc((size_t)ix->n * ix->idx_bytes, st));
    GK_TRY(init_indices_device((const uint64_t *)ix->d_segs.ptr, (uint32_t)ix->h_segs.size(),
                               ix->min_len, ix->n, ix->idx_bytes, ix->d_idx.ptr, st));
    ix->idx_ready = true;
    return GK_OK;
}

// Single-word sort: key over key_len <= 32 symbols, all windows of valid_len.
static int sort_single_level(gk_index *ix, uint32_t key_len, bool has_amb, gk_sort_stats *stats,
                             EventTimer &tm, cudaStream_t st)
{
    const uint64_t n = ix->n;
    const int ib = ix->idx_bytes;
    const int class_bit = has_amb ? 1 : 0;
    const int key_bits = 2 * (int)key_len + (class_bit ? 2 : 0);

    DeviceBuffer keys_a, keys_b, n_amb_dev;
    Owned idx_b;
    GK_TRY(keys_a.alloc((size_t)n * 8, st));
    GK_TRY(keys_b.alloc((size_t)n * 8, st));
    GK_TRY(idx_b.alloc((size_t)n * ib, st));
    GK_TRY(ix->d_idx.alloc((size_t)n * ib, st));
    GK_TRY(n_amb_dev.alloc(8, st));
    GK_CUDA(cudaMemsetAsync(n_amb_dev.ptr, 0, 8, st));

    const int e0 = tm.mark();
    GK_TRY(pack_keys_device(ix->d_sba, ix->sba_len, (const uint64_t *)ix->d_segs.ptr,
                            (uint32_t)ix->h_segs.size(), ix->min_len, key_len, class_bit, 0,
                            ix->sba_len, 0, keys_a.as<uint64_t>(), ib, ix->d_idx.ptr,
                            n_amb_dev.as<unsigned long long>(), st));
    const int e1 = tm.mark();
    int in_alt = 0;
    SortTiming timing;
    GK_TRY(radix_sort_pairs_device(keys_a.as<uint64_t>(), keys_b.as<uint64_t>(), ix->d_idx.ptr,
                                   idx_b.ptr, ib, n, 0, key_bits, &in_alt, st, &timing));
    const int e2 = tm.mark();
    uint64_t *keys_sorted = in_alt ? keys_b.as<uint64_t>() : keys_a.as<uint64_t>();
    if (in_alt) ix->d_idx.swap(idx_b);
    idx_b.reset();

    uint64_t n_amb = 0;
    GK_CUDA(cudaMemcpyAsync(&n_amb, n_amb_dev.ptr, 8, cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));

    // head flags of the sorted order; ambiguous slots are refined below
    GK_TRY(ix->d_flags.alloc((size_t)((n + 15) & ~15ull), st));
    GK_TRY(key_flags_device(keys_sorted, n, class_bit, (uint8_t *)ix->d_flags.ptr, st));

    if (n_amb > 0) {
        // ambiguous windows sit in the right slots as a set; order them among themselves by
        // their full 4-bit keys (16 symbols per word, least significant word first, stable)
        DeviceBuffer slot_pos, amb_a, amb_b, akeys_a, akeys_b;
        GK_TRY(slot_pos.alloc((size_t)n_amb * ib, st));
        GK_TRY(amb_a.alloc((size_t)n_amb * ib, st));
        GK_TRY(amb_b.alloc((size_t)n_amb * ib, st));
        GK_TRY(akeys_a.alloc((size_t)n_amb * 8, st));
        GK_TRY(akeys_b.alloc((size_t)n_amb * 8, st));
        uint64_t found = 0;
        if (ib == 4)
            GK_TRY((select_flagged_device<uint32_t, uint32_t>(
                (const uint8_t *)ix->d_flags.ptr, n, kFlagAmb, slot_pos.as<uint32_t>(),
                (const uint32_t *)ix->d_idx.ptr, amb_a.as<uint32_t>(), &found, st)));
        else
            GK_TRY((select_flagged_device<uint64_t, uint64_t>(
                (const uint8_t *)ix->d_flags.ptr, n, kFlagAmb, slot_pos.as<uint64_t>(),
                (const uint64_t *)ix->d_idx.ptr, amb_a.as<uint64_t>(), &found, st)));
        if (found != n_amb) {
            set_error("ambiguous window count mismatch: packed %llu, selected %llu",
                      (unsigned long long)n_amb, (unsigned long long)found);
            return GK_ERR_INTERNAL;
        }
        void *cur = amb_a.ptr, *alt = amb_b.ptr;
        const int words = ((int)key_len + 15) / 16;
        for (int w = words - 1; w >= 0; --w) {
            const int syms = ((int)key_len - 16 * w < 16) ? (int)key_len - 16 * w : 16;
            GK_TRY(pack4_gather_device(ix->d_sba, ix->sba_len, cur, ib, n_amb, (uint32_t)w, key_len,
                                       akeys_a.as<uint64_t>(), st));
            int alt_has = 0;
            GK_TRY(radix_sort_pairs_device(akeys_a.as<uint64_t>(), akeys_b.as<uint64_t>(), cur, alt,
                                           ib, n_amb, 64 - 4 * syms, 64, &alt_has, st, nullptr));
            if (alt_has) { void *t = cur; cur = alt; alt = t; }
        }
        GK_TRY(scatter_device(cur, slot_pos.ptr, n_amb, ib, ix->d_idx.ptr, st));
        GK_TRY(sba_flags_device(ix->d_sba, ix->sba_len, cur, ib, n_amb, key_len, slot_pos.ptr,
                                kFlagAmb, (uint8_t *)ix->d_flags.ptr, st));
    }
    const int e3 = tm.mark();
    if (stats) {
        stats->pack_ms = tm.ms(e0, e1);
        stats->hist_ms = timing.hist_ms;
        stats->sort_ms = timing.passes_ms;
        stats->fixup_ms = tm.ms(e2, e3);
        stats->sort_passes = timing.passes;
        stats->key_bits = key_bits;
        stats->levels = 1;
        stats->n_ambiguous = n_amb;
    }
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
    int rc = ix->d_segs.alloc((size_t)n_seg * 8);
    if (rc == GK_OK) {
        cudaError_t e = cudaMemcpy(ix->d_segs.ptr, h_seg_starts, (size_t)n_seg * 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            set_error("segment table upload failed: %s", cudaGetErrorString(e));
            rc = GK_ERR_CUDA;
        }
    }
    if (rc != GK_OK) {
        delete ix;
        return rc;
    }
    *out = ix;
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
    ix->sorted = sorted != 0;
    ix->flags_valid = false;
    return GK_OK;
}

int gk_index_sort(gk_index *ix, gk_sort_stats *stats_out, void *stream)
{
    if (!ix) return GK_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    gk_sort_stats stats;
    memset(&stats, 0, sizeof(stats));
    const uint64_t launches0 = gk_launch_count(0);
    EventTimer tm(st);
    const int t0 = tm.mark();

    GK_TRY(ensure_alphabet(ix, st));
    if (ix->n_sep != ix->h_segs.size() - 1) {
        // a '$' inside a record: the reference's sort raises through its validation
        set_error("kmers compared were less than min_kmer_len (%u).  Was kmer_sba_start_indices "
                  "initialized correctly?", ix->min_len);
        return GK_ERR_INVALID_KMERS;
    }
    const bool fixed = ix->max_len != 0 && ix->max_len == ix->min_len;
    if (!fixed) {
        set_error("sort() with min_kmer_len != max_kmer_len (variable-length / suffix order) is not "
                  "available on the GPU path yet");
        return GK_ERR_UNSUPPORTED;
    }
    const bool has_amb = ix->n_amb_letters > 0 || ix->n_bad > 0;
    const uint32_t k = ix->min_len;
    stats.n_windows = ix->n;
    if (ix->n == 0) {
        ix->sorted = true;
    } else if (k <= 31 || (k == 32 && !has_amb)) {
        GK_TRY(sort_single_level(ix, k, has_amb, &stats, tm, st));
        ix->idx_ready = true;
        ix->flags_valid = true;
        ix->flags_kmer_len = k;
        ix->sorted = true;
    } else {
        set_error("kmer_len %u needs multi-word keys (prefix-doubling), not available yet", k);
        return GK_ERR_UNSUPPORTED;
    }
    const int t1 = tm.mark();
    GK_CUDA(cudaStreamSynchronize(st));
    stats.total_ms = tm.ms(t0, t1);
    stats.gpu_launches = (int32_t)(gk_launch_count(0) - launches0);
    if (stats_out) *stats_out = stats;
    return GK_OK;
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

int gk_index_group_counts(gk_index *ix, uint32_t kmer_len, const gk_filter *filter, uint64_t min_group,
                          uint64_t max_group, uint64_t max_bin, int64_t *h_hist_out,
                          int64_t *h_total_out, void *stream)
{
    if (!ix) return GK_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    gk_filter keep_all = {GK_FILTER_KEEP_ALL, 0, 0, 0};
    const gk_filter f = filter ? *filter : keep_all;
    if (h_hist_out) memset(h_hist_out, 0, (size_t)(max_bin + 1) * 8);
    if (h_total_out) *h_total_out = 0;
    if (ix->n == 0) return GK_OK;
    GK_TRY(ensure_indices(ix, st));
    const int ib = ix->idx_bytes;
    const uint64_t n = ix->n;

    // fast path: sorted, same kmer_len as the sort, filter uniform over a group of equal k-mers
    const bool cached = ix->sorted && ix->flags_valid && ix->flags_kmer_len == kmer_len && kmer_len != 0;
    const bool no_amb_same_len = f.id == GK_FILTER_NO_AMBIGUOUS && (uint64_t)f.p0 == kmer_len;
    if (cached && (f.id == GK_FILTER_KEEP_ALL || no_amb_same_len)) {
        DeviceBuffer offsets;
        GK_TRY(offsets.alloc((size_t)n * ib, st));
        uint64_t n_groups = 0;
        if (ib == 4)
            GK_TRY((select_flagged_device<uint32_t, uint32_t>((const uint8_t *)ix->d_flags.ptr, n, kFlagHead,
                                                              offsets.as<uint32_t>(), nullptr, nullptr,
                                                              &n_groups, st)));
        else
            GK_TRY((select_flagged_device<uint64_t, uint64_t>((const uint8_t *)ix->d_flags.ptr, n, kFlagHead,
                                                              offsets.as<uint64_t>(), nullptr, nullptr,
                                                              &n_groups, st)));
        return group_hist_masked_device(offsets.ptr, ib, n_groups, n, (const uint8_t *)ix->d_flags.ptr,
                                        no_amb_same_len ? kFlagAmb : 0, min_group, max_group, max_bin,
                                        h_hist_out, h_total_out, st);
    }

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
        if (ib == 4)
            GK_TRY((select_flagged_device<uint32_t, uint32_t>(pass_flags.as<uint8_t>(), n, kFlagPass, nullptr,
                                                              (const uint32_t *)ix->d_idx.ptr,
                                                              kept.as<uint32_t>(), &m, st)));
        else
            GK_TRY((select_flagged_device<uint64_t, uint64_t>(pass_flags.as<uint8_t>(), n, kFlagPass, nullptr,
                                                              (const uint64_t *)ix->d_idx.ptr,
                                                              kept.as<uint64_t>(), &m, st)));
        d_list = kept.ptr;
    }
    if (m == 0) return GK_OK;
    if (!ix->sorted) {
        // get_kmer_count on unsorted data: every passing k-mer is its own group (kmers.py:1061-1064)
        if (min_group > 1) return GK_OK;
        if (h_hist_out) h_hist_out[1 < max_bin ? 1 : max_bin] = (int64_t)m;
        if (h_total_out) *h_total_out = (int64_t)m;
        return GK_OK;
    }
    DeviceBuffer flags, offsets;
    GK_TRY(flags.alloc((size_t)((m + 15) & ~15ull), st));
    GK_TRY(sba_flags_device(ix->d_sba, ix->sba_len, d_list, ib, m, kmer_len, nullptr, 0,
                            flags.as<uint8_t>(), st));
    GK_TRY(offsets.alloc((size_t)m * ib, st));
    uint64_t n_groups = 0;
    if (ib == 4)
        GK_TRY((select_flagged_device<uint32_t, uint32_t>(flags.as<uint8_t>(), m, kFlagHead,
                                                          offsets.as<uint32_t>(), nullptr, nullptr,
                                                          &n_groups, st)));
    else
        GK_TRY((select_flagged_device<uint64_t, uint64_t>(flags.as<uint8_t>(), m, kFlagHead,
                                                          offsets.as<uint64_t>(), nullptr, nullptr,
                                                          &n_groups, st)));
    return group_hist_device(offsets.ptr, ib, n_groups, m, min_group, max_group, max_bin, h_hist_out,
                             h_total_out, nullptr, st);
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
    *h_n_groups = 0;
    if (ix->n == 0) return GK_OK;
    const uint64_t n = ix->n;
    const uint8_t *d_flags = nullptr;
    DeviceBuffer scratch, offsets;
    GK_TRY(flags_for(ix, kmer_len, &d_flags, scratch, st));
    GK_TRY(offsets.alloc((size_t)n * 8, st));
    uint64_t n_groups = 0;
    GK_TRY((select_flagged_device<uint64_t, uint64_t>(d_flags, n, kFlagHead, offsets.as<uint64_t>(), nullptr,
                                                      nullptr, &n_groups, st)));
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
