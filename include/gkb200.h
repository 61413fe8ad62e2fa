/*
 * gkb200.h -- C ABI of the B200-native genome-kmers hot path (libgkb200.so).
 *
 * The reference (mrperkett/genome-kmers) is pure Python + numba and has no FFI; the seams this
 * library replaces are the two places where the reference hands work to numba-compiled code
 * (file:line relative to /root/reference/src/genome_kmers):
 *
 *   seam 1  kmers.py:1648   jit_sort_func(self.kmer_sba_start_indices)       -> gk_index_sort
 *   seam 2  kmers.py:1072, :1166  get_kmer_group_size_hist(sba, ..., max_counts_bin)
 *                                                                            -> gk_index_group_counts
 * plus the data producers either side of them:
 *   sequence_collection.py:42-73   reverse_complement_sba                   -> gk_sba_revcomp
 *   sequence_collection.py:693-697 alphabet check                           -> gk_sba_scan_alphabet
 *   kmers.py:789-835               _initialize_single_pass                  -> gk_index_create /
 *                                                                              gk_kmer_init_indices
 *
 * Conventions: every function returns a gk_status (0 = ok); gk_last_error() gives the text of
 * the last failure on the calling thread.  Pointers named d_* are device pointers on the current
 * CUDA device, h_* are host pointers.  `stream` is a cudaStream_t passed as void* (NULL = the
 * legacy default stream).  No torch types appear anywhere.  Index arrays hold sequence byte array
 * start positions as uint32 (idx_bytes == 4, the reference's dtype, kmers.py:811) or uint64
 * (idx_bytes == 8, the extension needed once a byte array exceeds 2^32 - 1 positions).
 *
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * GK_ERR_CUDA.
 */
#ifndef GKB200_H
#define GKB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GKB200_VERSION 100 /* 0.1.0 */

typedef enum {
    GK_OK = 0,
    GK_ERR_CUDA = 1,          /* a CUDA runtime call failed (see gk_last_error) */
    GK_ERR_ARG = 2,           /* invalid argument */
    GK_ERR_UNSUPPORTED = 3,   /* valid in the reference API but not implemented on this path */
    GK_ERR_INTERNAL = 4,      /* device-side consistency check failed (e.g. look-back timeout) */
    GK_ERR_INVALID_KMERS = 5, /* k-mers shorter than min_kmer_len (kmers.py:1724-1727) */
    GK_ERR_STATE = 6          /* call order violated (e.g. group counts on an unsorted index) */
} gk_status;

/* k-mer filters with device implementations (kmers.py:14-259).  p0..p2 as documented. */
typedef enum {
    GK_FILTER_KEEP_ALL = 0,     /* kmers.py:14-16 */
    GK_FILTER_NO_AMBIGUOUS = 1, /* kmers.py:195-229   p0 = kmer_len */
    GK_FILTER_MIN_LENGTH = 2,   /* kmers.py:19-34     p0 = min_kmer_len */
    GK_FILTER_HOMOPOLYMER = 3,  /* kmers.py:37-100    p0 = max_homopolymer_size, p1 = kmer_len */
    GK_FILTER_GC_COUNT = 4,     /* kmers.py:103-192   p0 = min_gc_count, p1 = max_gc_count, p2 = kmer_len */
    GK_FILTER_NGG_PAM = 5       /* kmers.py:232-259 */
} gk_filter_id;

typedef struct {
    int32_t id; /* gk_filter_id */
    int64_t p0, p1, p2;
} gk_filter;

/* Per-stage device times of the last gk_index_sort, in milliseconds (CUDA events). */
typedef struct {
    float pack_ms;        /* key-pack kernel(s) */
    float hist_ms;        /* up-front digit histogram + scan */
    float sort_ms;        /* all onesweep passes of the main sort */
    float fixup_ms;       /* ambiguous-window refinement + long-k refinement rounds */
    float total_ms;       /* whole sort, first launch to last */
    int32_t sort_passes;  /* onesweep passes launched for the main sort */
    int32_t key_bits;     /* radix key width of the main sort */
    int32_t levels;       /* 1 + number of prefix-doubling refinement rounds */
    int32_t gpu_launches; /* kernels launched by the call */
    uint64_t n_windows;   /* k-mers sorted */
    uint64_t n_ambiguous; /* windows holding a non-ACGT symbol within the key */
    uint64_t n_fragments; /* blocks of identical ambiguous windows listed by the pack kernel (0: not used) */
    uint64_t refine_flags; /* 1 fragment path used, 2 element-wise repair ran, 4 a long prefix run was out of
                              order, 16 repaired bucket-wise on the device, 32 big buckets re-sorted from the host,
                              64 all long runs re-sorted by key (too many were out of order for the bucket list),
                              fragment consistency error bits << 8 */
} gk_sort_stats;

typedef struct gk_index gk_index;

/* ---- library ---------------------------------------------------------------------------- */
int gk_version(void);
const char *gk_status_string(int status);
const char *gk_last_error(void);
/* sm_count / cc / memory of the current device; GK_ERR_CUDA when there is none. */
int gk_device_info(int *sm_count, int *cc_major, int *cc_minor, uint64_t *total_mem_bytes);
/* number of kernels launched by this library on the calling thread since the last reset */
uint64_t gk_launch_count(int reset);

/* ---- sequence byte array helpers ------------------------------ */
/* counts[0] = bytes outside the IUPAC+'$' alphabet (sequence_collection.py:441-459),
 * counts[1] = '$' bytes, counts[2] = ambiguous IUPAC letters (allowed, not A/C/G/T/'$'). */
int gk_sba_scan_alphabet(const uint8_t *d_sba, uint64_t len, uint64_t *h_counts3, void *stream);
/* the same counters left in device memory (zeroed by the call), no synchronise */
int gk_sba_scan_alphabet_async(const uint8_t *d_sba, uint64_t len, uint64_t *d_counts3, void *stream);
/* out[len-1-i] = complement(in[i]) (sequence_collection.py:42-73, :402-433); in != out. */
int gk_sba_revcomp(const uint8_t *d_in, uint64_t len, uint8_t *d_out, void *stream);
/* out = in || '$' || revcomp(in), 2*len+1 bytes: this library's definition of both strands. */
int gk_sba_both_strands(const uint8_t *d_in, uint64_t len, uint8_t *d_out, void *stream);
/* Bytes [begin, end) of that layout only: a GPU of a multi-GPU sort first builds the slice it packs. */
int gk_sba_both_strands_range(const uint8_t *d_in, uint64_t len, uint8_t *d_out, uint64_t begin, uint64_t end,
                              void *stream);

/* ---- k-mer start indices (row A3, kmers.py:789-835) --------------------------------------- */
/* Number of windows of kmer_len bases that fit inside a record. */
int gk_kmer_count(const uint64_t *h_seg_starts, uint32_t n_seg, uint64_t sba_len,
                  uint32_t kmer_len, uint64_t *n_out);
int gk_kmer_init_indices(const uint64_t *h_seg_starts, uint32_t n_seg, uint64_t sba_len,
                         uint32_t kmer_len, int idx_bytes, void *d_idx_out, void *stream);

/* ---- building blocks (also used by the multi-GPU driver and the unit tests) ---------------- */
/* Key-pack (north_star subsystem 1).  For every window of `valid_len` bases inside a record, in
 * ascending start order, emit a 64-bit radix key over its first key_len <= 32 symbols and its
 * start index.  Pure A/C/G/T windows get the 2-bit code (A<C<G<T); a window whose first
 * key_len symbols hold another symbol gets the count of pure windows that sort below it, so a
 * single integer sort orders both classes (DESIGN.md "two-class keys").  With class_bit = 1
 * the key is (value << 1) | is_pure; it is required whenever ambiguous symbols may occur and
 * needs key_len <= 31.  Only windows whose start lies in [first_start, end_start) are emitted
 * (a GPU's slice of the byte array); *h_n_out receives their number, *h_n_ambiguous the number
 * of non-pure ones. */
int gk_pack_keys(const uint8_t *d_sba, uint64_t sba_len, const uint64_t *h_seg_starts,
                 uint32_t n_seg, uint32_t valid_len, uint32_t key_len, int class_bit,
                 uint64_t first_start, uint64_t end_start, uint64_t *d_keys_out, int idx_bytes,
                 void *d_idx_out, uint64_t out_capacity, uint64_t *h_n_out,
                 uint64_t *h_n_ambiguous, void *stream);
/* Stable LSD onesweep radix sort of (key, value) pairs on key bits [begin_bit, end_bit)
 * (north_star subsystem 2).  Buffers ping-pong; *result_in_alt tells where the result is. */
int gk_radix_sort_pairs(uint64_t *d_keys, uint64_t *d_keys_alt, void *d_vals, void *d_vals_alt,
                        int val_bytes, uint64_t n, int begin_bit, int end_bit,
                        int *result_in_alt, void *stream);
/* The same sort for 32-bit keys on key bits [begin_bit, end_bit) of 32: with 32-bit values a pair is 8 bytes,
 * so a pass moves 16 bytes per pair instead of 24 (the planned split-key path sorts
 * (top key word, start index) pairs).  Same contract as gk_radix_sort_pairs. */
int gk_radix_sort_pairs32(uint32_t *d_keys, uint32_t *d_keys_alt, void *d_vals, void *d_vals_alt,
                          int val_bytes, uint64_t n, int begin_bit, int end_bit,
                          int *result_in_alt, void *stream);
/* Multi-GPU partition step: stable split of the pairs by destination rank = number of splitters
 * (sorted, n_parts-1 of them, device memory) that are <= key.  One onesweep pass whose "digit" is
 * the destination.  Output in the *_out buffers; h_counts_out[d] = pairs for destination d. */
int gk_partition_pairs(uint64_t *d_keys, uint64_t *d_keys_out, void *d_vals, void *d_vals_out,
                       int val_bytes, uint64_t n, const uint64_t *d_splitters, uint32_t n_parts,
                       uint64_t *h_counts_out, void *stream);
/* Fused partition + exchange (multi-GPU): the same stable partition pass, but every pair
 * is written straight into its destination rank's receive buffer over NVLink peer memory, so the
 * exchange overlaps the partition tile by tile and no separate all-to-all of the pairs is needed.
 *   gk_partition_count       destination counts only (the ranks exchange them to place the segments)
 *   gk_partition_pairs_peer  h_dst_keys / h_dst_vals: n_parts <= 16 device pointers, local or opened with
 *                            gk_peer_open; h_dst_offsets[d]: first element of this rank's segment in
 *                            destination d.  Consumers on other ranks must be ordered behind this call
 *                            by the caller (a stream-ordered collective). */
int gk_partition_count(const uint64_t *d_keys, uint64_t n, const uint64_t *d_splitters, uint32_t n_parts,
                       uint64_t *h_counts_out, void *stream);
/* The same counts left on the DEVICE (no synchronise): d_counts_out[d] = pure pairs for destination d,
 * d_counts_out[n_parts + d] = pairs of class 0 (ambiguous windows) when class_bit is set. */
int gk_partition_count_split(const uint64_t *d_keys, uint64_t n, const uint64_t *d_splitters, uint32_t n_parts,
                             int class_bit, uint64_t *d_counts_out, void *stream);
/* h_key_base (optional, n_parts entries): subtracted from every key sent to destination d, so that a rank
 * sorts keys relative to the start of its own key range.  skip_ambiguous: pairs of class 0 are dropped (their
 * windows travel as run-length fragments, gk_pack_slice).  d_err (optional device int): receives the look-back
 * guard's verdict and the call does not synchronise. */
int gk_partition_pairs_peer(const uint64_t *d_keys, const void *d_vals, int val_bytes, uint64_t n,
                            const uint64_t *d_splitters, uint32_t n_parts, uint64_t *const *h_dst_keys,
                            void *const *h_dst_vals, const uint64_t *h_dst_offsets, const uint64_t *h_key_base,
                            int skip_ambiguous, int *d_err, void *stream);
/* Peer-visible device memory for those buffers: cudaMalloc + CUDA IPC handles (64 bytes), one process
 * per GPU on one box. */
int gk_peer_alloc(uint64_t bytes, void **d_ptr_out);
int gk_peer_free(void *d_ptr);
int gk_peer_export(void *d_ptr, uint8_t *handle64_out);
int gk_peer_open(const uint8_t *handle64, void **d_ptr_out);
int gk_peer_close(void *d_ptr);
/* Run-length pass over sorted keys (north_star subsystem 3): group offsets (positions where a
 * new key starts; d_offsets_out has room for n entries), number of groups. */
int gk_rle_keys(const uint64_t *d_keys_sorted, uint64_t n, uint64_t *d_offsets_out,
                uint64_t *h_n_groups, void *stream);
/* counts_by_group_size[min(size, max_bin)] += 1 and total += size over groups with
 * min_group <= size <= max_group (0 = no maximum) -- kmers.py:514-518, :612-614. */
int gk_group_size_hist(const uint64_t *d_offsets, uint64_t n_groups, uint64_t n,
                       uint64_t min_group, uint64_t max_group, uint64_t max_bin,
                       int64_t *h_hist_out, int64_t *h_total_out, void *stream);

/* ---- the index object: drop-in for the state behind reference `Kmers` --------------------- */
/* Borrow d_sba (caller keeps it alive and unchanged), copy the segment table.  max_kmer_len 0
 * means None.  Mirrors Kmers.__init__ validation that depends on the data (kmers.py:743-748,
 * :805-808 is lifted: more than 2^32-1 k-mers switch the index to uint64). */
int gk_index_create(const uint8_t *d_sba, uint64_t sba_len, const uint64_t *h_seg_starts,
                    uint32_t n_seg, uint32_t min_kmer_len, uint32_t max_kmer_len,
                    gk_index **out);
void gk_index_destroy(gk_index *ix);
uint64_t gk_index_size(const gk_index *ix);   /* number of k-mers, kmers.py:863-864 */
int gk_index_idx_bytes(const gk_index *ix);   /* 4 or 8 */
int gk_index_is_sorted(const gk_index *ix);
/* Replace the start indices (e.g. Kmers.load); `sorted` states what the caller knows. */
int gk_index_set_indices(gk_index *ix, const void *h_idx, uint64_t n, int idx_bytes, int sorted,
                         void *stream);
/* seam 1: sort the start indices lexicographically by k-mer, ties by ascending start. */
int gk_index_sort(gk_index *ix, gk_sort_stats *stats_out, void *stream);
/* Multi-GPU shard of seam 1: adopt the (key, start) pairs this rank received from the exchange
 * (made by gk_pack_keys / gk_pack_slice with key_len = valid_len = min_kmer_len and the same class_bit), sort
 * them and leave the index describing this rank's key range only (gk_index_size() == n_local).  The
 * pair buffers are scratch. */
int gk_index_sort_pairs(gk_index *ix, uint64_t *d_keys, uint64_t *d_keys_alt, void *d_idx,
                        void *d_idx_alt, uint64_t n_local, int class_bit, gk_sort_stats *stats_out,
                        void *stream);
/* k-mers longer than one key word (min_kmer_len > 31, class-bit keys, starts of either width): the pairs carry the first 31
 * symbols; after the sort the members of groups that are still tied are ordered by the remaining symbols, read
 * from the bytes eight at a time (the prefix doubling of gk_index_sort needs the ranks of other starts, which
 * live on other GPUs).
 * The same with everything the sharded driver knows: key_bits = width of the received keys when they are
 * relative to the start of this rank's key range (0: full width); d_frag_gathered: the all-gathered fragment
 * lists of all n_sources ranks (gk_pack_slice layout, frag_capacity entries each, d_frag_counts[s] used), from
 * which the n_ambiguous windows of the key range [key_lo, key_hi) (key_hi 0: unbounded) are generated behind
 * the n_pure received pairs -- the four buffers hold n_pure + n_ambiguous pairs; frag_presorted: every list was
 * sorted by gk_frag_sort_local; d_err: the device word the
 * partition step wrote its look-back verdict to (optional). */
int gk_index_sort_shard(gk_index *ix, uint64_t *d_keys, uint64_t *d_keys_alt, void *d_idx, void *d_idx_alt,
                        uint64_t n_pure, uint64_t n_ambiguous, int class_bit, int key_bits,
                        const void *d_frag_gathered, const uint64_t *d_frag_counts, uint32_t n_sources,
                        uint64_t frag_capacity, int frag_presorted, uint64_t key_lo, uint64_t key_hi,
                        const int *d_err, gk_sort_stats *stats_out, void *stream);
/* Sort one rank's fragment list (gk_pack_slice layout, n_frag entries used) by window and start into
 * d_frag_sorted (same layout and capacity); no synchronise.  With frag_presorted = 1 gk_index_sort_shard merges
 * the gathered lists (one binary search per fragment and list) instead of sorting their union. */
int gk_frag_sort_local(const void *d_frag, uint64_t frag_capacity, uint64_t n_frag, uint32_t kmer_len,
                       uint64_t sba_len, void *d_frag_sorted, void *stream);
/* Multi-GPU producers (no synchronise).  gk_pack_slice: pack the windows (kmer_len symbols; the key covers the
 * first min(kmer_len, 31) of them when class_bit is set) whose start lies in
 * [first_start, end_start) and list the ambiguous-window fragments of that slice; d_frag: frag_capacity * 36
 * bytes laid out key[cap] w0[cap] w1[cap] start[cap] (u64) count[cap] (u32); d_counters: 4 x u64, zeroed by the
 * call: [0] ambiguous windows, [2] fragments found (above the capacity the list is incomplete).
 * gk_sample_keys: keys of n_samples evenly spaced windows of the slice, for the splitter choice. */
int gk_pack_slice(const uint8_t *d_sba, uint64_t sba_len, const uint64_t *h_seg_starts, uint32_t n_seg,
                  uint32_t kmer_len, int class_bit, uint64_t first_start, uint64_t end_start,
                  uint64_t *d_keys_out, int idx_bytes, void *d_idx_out, uint64_t out_capacity,
                  uint64_t *h_n_out, void *d_frag, uint64_t frag_capacity, uint64_t *d_counters, void *stream);
int gk_sample_keys(const uint8_t *d_sba, uint64_t sba_len, const uint64_t *h_seg_starts, uint32_t n_seg,
                   uint32_t kmer_len, int class_bit, uint64_t first_start, uint64_t end_start,
                   uint32_t n_samples, uint64_t *d_keys_out, uint32_t *h_n_out, void *stream);
/* Device pointer to the current (init or sorted) start indices; owned by the index. */
int gk_index_device_indices(gk_index *ix, const void **d_idx_out, void *stream);
/* Copy the start indices to a host buffer of gk_index_size() * idx_bytes bytes. */
int gk_index_copy_indices(gk_index *ix, void *h_dst, void *stream);
/* seam 2: group-size histogram and total (kmers.py:454-520).  kmer_len 0 means None.
 * sorted semantics follow get_kmer_count: on an unsorted index every passing k-mer is its own
 * group (kmers.py:1061-1064). h_hist_out may be NULL (get_kmer_count). */
int gk_index_group_counts(gk_index *ix, uint32_t kmer_len, const gk_filter *filter,
                          uint64_t min_group, uint64_t max_group, uint64_t max_bin,
                          int64_t *h_hist_out, int64_t *h_total_out, void *stream);
/* Same query for a caller whose table is ALREADY all zero (fresh calloc / numpy.zeros pages): only bins
 * [0, *h_top_bin_out] are written, so the reference's default 8 MB table (max_counts_bin = 1000000,
 * kmers.py:1091) is never touched beyond the occupied bins. */
int gk_index_group_counts_zeroed(gk_index *ix, uint32_t kmer_len, const gk_filter *filter,
                                 uint64_t min_group, uint64_t max_group, uint64_t max_bin,
                                 int64_t *h_hist_zeroed, int64_t *h_total_out, uint64_t *h_top_bin_out,
                                 void *stream);
/* Same query, sparse: the occupied bins as (bin, count) pairs in ascending bin order.  This is what
 * the multi-GPU driver exchanges between ranks.  GK_ERR_ARG (with *h_n_pairs_out set) when the
 * pairs do not fit `capacity`. */
int gk_index_group_counts_sparse(gk_index *ix, uint32_t kmer_len, const gk_filter *filter,
                                 uint64_t min_group, uint64_t max_group, uint64_t max_bin,
                                 uint64_t *h_bins_out, int64_t *h_counts_out, uint64_t capacity,
                                 uint64_t *h_n_pairs_out, int64_t *h_total_out, void *stream);
/* Group table of the sorted index for kmer_len: number of groups, and (optionally, host
 * buffers of n_groups entries obtained by a first call with NULLs) offsets into the sorted
 * order and sizes.  This is the unique-k-mer set: one entry per distinct k-mer. */
int gk_index_groups(gk_index *ix, uint32_t kmer_len, uint64_t *h_n_groups,
                    uint64_t *h_offsets_out, uint64_t *h_sizes_out, void *stream);

/* Group table of the k-mers that pass `filter` (kmers.py:586-601: failing k-mers are skipped, a passing
 * k-mer is compared with the previous PASSING one).  Two-call protocol as above.  h_kept_pos_out[j] =
 * position in the index of the j-th passing k-mer; offsets and sizes refer to that list.  On an unsorted
 * index every passing k-mer is its own group (kmers.py:1061-1064).  Behind Kmers.get_kmers(filter). */
int gk_index_groups_filtered(gk_index *ix, uint32_t kmer_len, const gk_filter *filter, uint64_t *h_n_kept,
                             uint64_t *h_n_groups, uint64_t *h_kept_pos_out, uint64_t *h_offsets_out,
                             uint64_t *h_sizes_out, void *stream);

/* Self-check of the current order against the sequence bytes, independent of how it was produced: the
 * reference's '$'-terminated comparator (kmers.py:306-397) on every pair of neighbours, ties in ascending
 * start order (kmers.py:1710-1711), every start a valid, distinct k-mer start (kmers.py:814-826).
 * h_report8: [0] k-mers checked, [1] neighbours out of order, [2] ties not in ascending start order,
 * [3] invalid starts, [4] duplicate starts, [5] groups of equal k-mers counted from the bytes,
 * [6] cached head flags that disagree with the bytes, [7] 1 when cached flags were compared.
 * A correct sort of the init set has [1..4] == 0, [6] == 0 and [0] == gk_kmer_count(). kmer_len 0 = None.
 * d_seen_bitmap (optional): a zeroed device bitmap of sba_len / 32 + 1 words that receives one bit per start
 * seen; the shards of a multi-GPU index sum their bitmaps (disjoint bitmaps add without carries) and count
 * the bits with gk_popcount_words to prove that together they hold every start exactly once. */
int gk_index_verify(gk_index *ix, uint32_t kmer_len, uint64_t *h_report8, uint32_t *d_seen_bitmap, void *stream);
int gk_popcount_words(const uint32_t *d_words, uint64_t n_words, uint64_t *h_count_out, void *stream);

/* ---- one-shot host entry point (host buffers in, host buffers out) ------------------------ */
/* sba (forward strand, records joined by '$') -> sorted start indices + histogram.  strands:
 * 0 forward, 2 both (index space of forward || '$' || revcomp).  h_idx_out may be NULL.
 * This is what a non-Python host would bind; bench.py's e2e leg times the Python equivalent. */
int gk_sort_count_host(const uint8_t *h_sba, uint64_t sba_len, const uint64_t *h_seg_starts,
                       uint32_t n_seg, uint32_t kmer_len, int strands, int idx_bytes,
                       void *h_idx_out, uint64_t max_bin, int64_t *h_hist_out,
                       int64_t *h_total_out, uint64_t *h_n_kmers_out, gk_sort_stats *stats_out);

#ifdef __cplusplus
}
#endif
#endif /* GKB200_H */
