/*
 * gk_oracle.c -- CPU restatement of the genome-kmers hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA path in genome-kmers_b200/csrc.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call it.
 * The product (genome_kmers.Kmers) never does: it fails loudly when the CUDA library is
 * missing.
 *
 * Parity pin: every function here is checked against golden vectors produced by importing
 * the real reference (/root/reference, numba 0.65) in tests/golden/make_golden.py, and against
 * the reference's own known-answer tests (tests/test_oracle_golden.py).
 *
 * Each function cites the reference lines (relative to /root/reference/src/genome_kmers) it
 * restates.  The algorithm is the reference's: a byte-wise, '$'-terminated lexicographic
 * comparator driving a median-of-3 quicksort with an insertion-sort cutoff, followed by a
 * linear group walk.  The only addition is the optional index tie-break, which is the
 * reference's own break_ties=True comparator (kmers.py:1710-1711) and yields the canonical
 * order that parity is stated in (the reference's default tie order is an accident of
 * numba's quicksort, SURVEY.md "Read this first" #2).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GKO_SEP 36 /* ord('$'), sequence_collection.py:689-691 */

/* ------------------------------------------------------------------------------------------
 * A1. sequence byte array: records joined by '$' (sequence_collection.py:663-699) and the
 * segment start table (sequence_collection.py:702-726).  Input is the concatenation of all
 * record bytes plus the record lengths.  Returns the sba length, or -1 when a record is
 * empty (sequence_collection.py:654-658) or a byte is outside the IUPAC+'$' alphabet
 * (sequence_collection.py:441-459, :693-697).
 * ---------------------------------------------------------------------------------------- */
static int gko_allowed(uint8_t b)
{
    switch (b) {
    case 'A': case 'C': case 'G': case 'T': case 'R': case 'Y': case 'S': case 'W':
    case 'K': case 'M': case 'B': case 'D': case 'H': case 'V': case 'N': case '$':
        return 1;
    default:
        return 0;
    }
}

int64_t gko_build_sba(const uint8_t *bases, const int64_t *rec_len, int64_t n_rec,
                      uint8_t *sba_out, uint32_t *seg_starts_out)
{
    int64_t w = 0, r = 0, src = 0;
    for (r = 0; r < n_rec; ++r) {
        if (rec_len[r] <= 0)
            return -1;
        seg_starts_out[r] = (uint32_t)w;
        for (int64_t i = 0; i < rec_len[r]; ++i) {
            uint8_t b = bases[src++];
            if (!gko_allowed(b))
                return -1;
            sba_out[w++] = b;
        }
        if (r != n_rec - 1)
            sba_out[w++] = GKO_SEP;
    }
    return w;
}

/* ------------------------------------------------------------------------------------------
 * A2. reverse complement of an sba (sequence_collection.py:42-73) with the IUPAC complement
 * table (sequence_collection.py:402-433), and the mirrored segment starts
 * (sequence_collection.py:905-928): the new start of a segment is the mirrored old end, and
 * segments appear in reversed record order.
 * ---------------------------------------------------------------------------------------- */
static uint8_t gko_complement(uint8_t b)
{
    switch (b) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'R': return 'Y'; case 'Y': return 'R'; case 'S': return 'S'; case 'W': return 'W';
    case 'K': return 'M'; case 'M': return 'K'; case 'B': return 'V'; case 'D': return 'H';
    case 'H': return 'D'; case 'V': return 'B'; case 'N': return 'N'; case '$': return '$';
    default: return 0; /* the reference's table is zero for bytes it does not know */
    }
}

void gko_revcomp(const uint8_t *sba, int64_t len, uint8_t *out)
{
    for (int64_t i = 0; i < len; ++i)
        out[len - 1 - i] = gko_complement(sba[i]);
}

void gko_revcomp_seg_starts(const uint32_t *starts, int64_t n_rec, int64_t sba_len,
                            uint32_t *out)
{
    for (int64_t s = 0; s < n_rec; ++s) {
        int64_t end = (s == n_rec - 1) ? sba_len - 1 : (int64_t)starts[s + 1] - 2;
        out[n_rec - 1 - s] = (uint32_t)(sba_len - 1 - end);
    }
}

/* forward || '$' || revcomp: the definition of source_strand="both" used by this repo
 * (SURVEY.md section 8c): running the reference's forward path over a collection made of the
 * forward records followed by the reverse-complemented records in reversed order gives
 * exactly this byte array. */
void gko_both_strands(const uint8_t *sba, int64_t len, uint8_t *out /* 2*len+1 */)
{
    memcpy(out, sba, (size_t)len);
    out[len] = GKO_SEP;
    gko_revcomp(sba, len, out + len + 1);
}

/* ------------------------------------------------------------------------------------------
 * A3. k-mer start index initialisation (kmers.py:789-835, count :837-861): every start whose
 * min_kmer_len bases lie inside one record, ascending.  Returns the count; writes when
 * out != NULL.  Starts are written as uint64 so the >2^32 extension can be checked too.
 * ---------------------------------------------------------------------------------------- */
int64_t gko_init_indices(const uint64_t *seg_starts, int64_t n_rec, int64_t sba_len,
                         int64_t min_kmer_len, uint64_t *out)
{
    int64_t n = 0;
    for (int64_t s = 0; s < n_rec; ++s) {
        int64_t a = (int64_t)seg_starts[s];
        int64_t e = (s == n_rec - 1) ? sba_len - 1 : (int64_t)seg_starts[s + 1] - 2;
        int64_t cnt = (e - a + 1) - min_kmer_len + 1;
        for (int64_t i = 0; i < cnt; ++i) {
            if (out)
                out[n] = (uint64_t)(a + i);
            ++n;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * A4. the comparator (kmers.py:306-397).  max_len <= 0 means None (compare to the end of
 * the record).  A '$' or the end of the array terminates a k-mer and the shorter one sorts
 * first (:360-378).  Bytes compare by raw value (:381-388).  Returns -1/0/+1 and the last
 * k-mer offset compared through *last (the reference raises when nothing could be compared;
 * here *last = -1).
 * ---------------------------------------------------------------------------------------- */
static inline int gko_compare_last(const uint8_t *sba, int64_t len, int64_t a, int64_t b,
                                   int64_t max_len, int64_t *last)
{
    int64_t j = 0;
    for (;;) {
        int64_t ia = a + j, ib = b + j;
        int a_out = (ia >= len) || sba[ia] == GKO_SEP;
        int b_out = (ib >= len) || sba[ib] == GKO_SEP;
        if (a_out || b_out) {
            *last = j - 1;
            if (a_out && !b_out) return -1;
            if (b_out && !a_out) return 1;
            return 0;
        }
        if (sba[ia] < sba[ib]) { *last = j; return -1; }
        if (sba[ia] > sba[ib]) { *last = j; return 1; }
        if (max_len > 0 && j == max_len - 1) { *last = j; return 0; }
        ++j;
    }
}

int gko_compare(const uint8_t *sba, int64_t len, int64_t a, int64_t b, int64_t max_len)
{
    int64_t last;
    return gko_compare_last(sba, len, a, b, max_len, &last);
}

/* kmers.py:262-282 */
static inline int gko_has_required_len(const uint8_t *sba, int64_t len, int64_t start,
                                       int64_t need)
{
    for (int64_t i = start; i < start + need; ++i)
        if (i >= len || sba[i] == GKO_SEP)
            return 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * A5. the sort (kmers.py:1624-1652 driving numba/misc/quicksort.py:165-197).  is_less_than
 * follows kmers.py:1690-1729 including the min_kmer_len validation (:1716-1727), which sets
 * ctx->invalid instead of raising.  break_ties selects the canonical (k-mer, start) order.
 *
 * numba 0.65 (pinned ^0.59.1 in the reference's pyproject.toml:21) is not vendored in
 * /root/reference; its published algorithm is restated here: explicit-stack quicksort,
 * median-of-three pivot, Hoare-style partition, insertion sort for partitions < 16 elements,
 * smaller partition processed first.  With break_ties the comparator is a strict total order,
 * so the result does not depend on these details.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t *sba;
    int64_t len;
    int64_t min_len;
    int64_t max_len; /* <=0: None */
    int break_ties;
    int validate;
    volatile int invalid;
} gko_sort_ctx;

static inline int gko_lt(gko_sort_ctx *c, uint64_t a, uint64_t b)
{
    int64_t last;
    int cmp = gko_compare_last(c->sba, c->len, (int64_t)a, (int64_t)b, c->max_len, &last);
    int lt = cmp < 0 ? 1 : (cmp > 0 ? 0 : (c->break_ties ? a < b : 0));
    if (c->validate) {
        int64_t need = c->min_len - (last + 1);
        if (!gko_has_required_len(c->sba, c->len, (int64_t)a + last + 1, need) ||
            !gko_has_required_len(c->sba, c->len, (int64_t)b + last + 1, need))
            c->invalid = 1;
    }
    return lt;
}

#define GKO_SMALL 15

static void gko_insertion(gko_sort_ctx *c, uint64_t *v, int64_t lo, int64_t hi)
{
    for (int64_t i = lo + 1; i <= hi; ++i) {
        uint64_t x = v[i];
        int64_t j = i;
        while (j > lo && gko_lt(c, x, v[j - 1])) {
            v[j] = v[j - 1];
            --j;
        }
        v[j] = x;
    }
}

static int64_t gko_partition(gko_sort_ctx *c, uint64_t *v, int64_t lo, int64_t hi)
{
    int64_t mid = lo + ((hi - lo) >> 1);
    uint64_t t;
    /* median of three into v[mid] */
    if (gko_lt(c, v[mid], v[lo])) { t = v[lo]; v[lo] = v[mid]; v[mid] = t; }
    if (gko_lt(c, v[hi], v[mid])) {
        t = v[hi]; v[hi] = v[mid]; v[mid] = t;
        if (gko_lt(c, v[mid], v[lo])) { t = v[lo]; v[lo] = v[mid]; v[mid] = t; }
    }
    uint64_t pivot = v[mid];
    t = v[mid]; v[mid] = v[hi]; v[hi] = t; /* park the pivot at the end */
    int64_t i = lo, j = hi - 1;
    for (;;) {
        while (i < hi && gko_lt(c, v[i], pivot)) ++i;
        while (j >= lo && gko_lt(c, pivot, v[j])) --j;
        if (i >= j) break;
        t = v[i]; v[i] = v[j]; v[j] = t;
        ++i; --j;
    }
    t = v[i]; v[i] = v[hi]; v[hi] = t;
    return i;
}

static void gko_quicksort_serial(gko_sort_ctx *c, uint64_t *v, int64_t lo, int64_t hi)
{
    /* explicit stack, smaller side first: depth <= 64 */
    int64_t stack_lo[128], stack_hi[128];
    int sp = 0;
    stack_lo[sp] = lo; stack_hi[sp] = hi; ++sp;
    while (sp > 0) {
        --sp;
        lo = stack_lo[sp]; hi = stack_hi[sp];
        while (hi - lo >= GKO_SMALL) {
            int64_t p = gko_partition(c, v, lo, hi);
            if (hi - p > p - lo) {
                stack_lo[sp] = p + 1; stack_hi[sp] = hi; ++sp;
                hi = p - 1;
            } else {
                stack_lo[sp] = lo; stack_hi[sp] = p - 1; ++sp;
                lo = p + 1;
            }
        }
        gko_insertion(c, v, lo, hi);
    }
}

#ifdef _OPENMP
static void gko_quicksort_tasks(gko_sort_ctx *c, uint64_t *v, int64_t lo, int64_t hi,
                                int64_t grain)
{
    while (hi - lo >= grain) {
        int64_t p = gko_partition(c, v, lo, hi);
#pragma omp task default(none) firstprivate(c, v, lo, p, grain)
        gko_quicksort_tasks(c, v, lo, p - 1, grain);
        lo = p + 1;
    }
    if (hi > lo)
        gko_quicksort_serial(c, v, lo, hi);
}
#endif

/* Sort n start indices in place.  threads <= 1: the reference's single-threaded algorithm.
 * threads > 1: the same partitioning run as OpenMP tasks (the "all host threads" arm of
 * bench.py --impl reference).  Returns 0, or 1 when the validation of kmers.py:1716-1727
 * would have raised. */
int gko_sort(const uint8_t *sba, int64_t len, uint64_t *idx, int64_t n, int64_t min_len,
             int64_t max_len, int break_ties, int validate, int threads)
{
    gko_sort_ctx c = {sba, len, min_len, max_len, break_ties, validate, 0};
    if (n < 2)
        return 0;
#ifdef _OPENMP
    if (threads > 1) {
        int64_t grain = n / (64 * (int64_t)threads);
        if (grain < 4096) grain = 4096;
#pragma omp parallel num_threads(threads)
#pragma omp single nowait
        gko_quicksort_tasks(&c, idx, 0, n - 1, grain);
        return c.invalid;
    }
#endif
    (void)threads;
    gko_quicksort_serial(&c, idx, 0, n - 1);
    return c.invalid;
}

/* ------------------------------------------------------------------------------------------
 * k-mer filters (kmers.py:14-259).  id: 0 keep_all (:14-16), 1 no_ambiguous_bases(p0=k)
 * (:195-229), 2 min_length(p0=len) (:19-34), 3 homopolymer(p0=max, p1=k) (:37-100),
 * 4 gc_count(p0=min_count, p1=max_count, p2=k) (:103-192, counts precomputed by the caller
 * with the reference's ceil/floor at :143-144), 5 crispr_ngg_pam (:232-259).
 * Returns 1 pass, 0 fail, -1 where the reference would raise.
 * ---------------------------------------------------------------------------------------- */
static int gko_filter(const uint8_t *sba, int64_t len, int64_t s, int id, int64_t p0,
                      int64_t p1, int64_t p2)
{
    switch (id) {
    case 0:
        return 1;
    case 1:
        if (s + p0 > len) return -1;
        for (int64_t i = 0; i < p0; ++i) {
            uint8_t b = sba[s + i];
            if (b == GKO_SEP) return -1;
            if (b != 'A' && b != 'T' && b != 'G' && b != 'C') return 0;
        }
        return 1;
    case 2:
        return gko_has_required_len(sba, len, s, p0);
    case 3: {
        if (s + p1 - 1 >= len) return -1;
        if (p1 < p0) return 1;
        int64_t run = 1;
        for (int64_t i = 1; i < p1; ++i) {
            if (sba[s + i] == GKO_SEP) return -1;
            if (sba[s + i] == sba[s + i - 1]) {
                if (++run > p0) return 0;
            } else {
                run = 1;
            }
        }
        return 1;
    }
    case 4: {
        if (p1 < p0) return 0;
        int64_t gc = 0;
        for (int64_t i = 0; i < p2; ++i) {
            uint8_t b = sba[s + i];
            if (b == GKO_SEP) return -1;
            if (b == 'G' || b == 'C')
                if (++gc > p1) return 0;
        }
        return (p0 <= gc && gc <= p1) ? 1 : 0;
    }
    case 5:
        if (s + 23 > len) return -1;
        return (sba[s + 21] == 'G' && sba[s + 22] == 'G') ? 1 : 0;
    default:
        return -1;
    }
}

/* ------------------------------------------------------------------------------------------
 * A6 + A7. the group walk (kmers.py:523-648) feeding the group-size histogram
 * (kmers.py:454-520): a k-mer joins the current group iff it compares equal (kmer_len bytes,
 * '$'-terminated) to the previous k-mer that passed the filter (:597-601); a finished group is
 * counted iff min_group <= size <= max_group (:612-614, :633-635); hist[min(size, max_bin)]++
 * and total += size (:516-518).  sorted == 0 restates get_kmer_count on unsorted data, where
 * every passing k-mer is its own group (:1061-1064).  max_group <= 0 means None.
 *
 * Optional outputs (NULL to skip) describe the groups that were counted, in order:
 * group_first[g] = position in idx of the group's first passing k-mer, group_size[g].
 * Returns the number of counted groups, or -1 on a filter error.
 * ---------------------------------------------------------------------------------------- */
int64_t gko_group_hist(const uint8_t *sba, int64_t len, const uint64_t *idx, int64_t n,
                       int64_t kmer_len, int sorted, int filter_id, int64_t p0, int64_t p1,
                       int64_t p2, int64_t min_group, int64_t max_group, int64_t max_bin,
                       int64_t *hist /* max_bin+1, zeroed here */, int64_t *total_out,
                       int64_t *group_first, int64_t *group_size)
{
    int64_t total = 0, n_groups = 0, size = 0, first = -1, prev = -1;
    if (hist)
        memset(hist, 0, (size_t)(max_bin + 1) * sizeof(int64_t));
    for (int64_t p = 0; p <= n; ++p) {
        int same = 0, flush = 0;
        if (p < n) {
            int64_t s = (int64_t)idx[p];
            int pass = gko_filter(sba, len, s, filter_id, p0, p1, p2);
            if (pass < 0)
                return -1;
            if (!pass)
                continue;
            if (prev < 0)
                same = 1;
            else
                same = sorted ? (gko_compare(sba, len, prev, s, kmer_len) == 0) : 0;
            prev = s;
            if (same) {
                if (size == 0)
                    first = p;
                ++size;
                continue;
            }
            flush = 1;
        } else {
            flush = 1;
        }
        if (flush) {
            if (size >= min_group && (max_group <= 0 || size <= max_group)) {
                if (size > 0) {
                    total += size;
                    if (hist)
                        hist[size < max_bin ? size : max_bin] += 1;
                    if (group_first) group_first[n_groups] = first;
                    if (group_size) group_size[n_groups] = size;
                    ++n_groups;
                }
            }
            if (p < n) {
                size = 1;
                first = p;
            }
        }
    }
    if (total_out)
        *total_out = total;
    return n_groups;
}

int gko_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
