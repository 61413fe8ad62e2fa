"""
Known answers of the reference's own test-suite for the hot path, on its fixtures (`seq_list_2`:
chr1 ATCGAATTAG, chr2 GGATCTTGCATT, chr3 GTGATTGACCCCT), run against this repo's drop-in through the GPU:
  * the sorted 3-mer list and its groups        reference tests/test_kmers.py:984-1038 (docs/overview.rst:46-74)
  * 3-mer counts 16 / 10 / 3 / 13 / 26           :1855-1887
  * 1-mer total 35 with A=8 T=12 G=8 C=7         :1889-1916
  * 5-mers: 23, all unique                       :1918-1944
  * group generator sweeps over min/max group size and yield_first_n, against Python grouping   :1466-1568
  * the module-level seams called the way the reference's tests call them   :454-648 (kmers.py)
The expectations are either the reference's literal numbers or a plain Python enumeration of the fixture,
as in the reference's tests.
"""
import itertools

import numpy as np
import pytest

from genome_kmers import kmers as gk
from genome_kmers.kmers import Kmers
from genome_kmers.sequence_collection import SequenceCollection

pytestmark = pytest.mark.gpu

SEQ_LIST_2 = [("chr1", "ATCGAATTAG"), ("chr2", "GGATCTTGCATT"), ("chr3", "GTGATTGACCCCT")]
SORTED_3MERS = ["AAT", "ACC", "ATC", "ATC", "ATT", "ATT", "ATT", "CAT", "CCC", "CCC", "CCT", "CGA", "CTT", "GAA",
                "GAC", "GAT", "GAT", "GCA", "GGA", "GTG", "TAG", "TCG", "TCT", "TGA", "TGA", "TGC", "TTA", "TTG",
                "TTG"]
KMER_NUMS_BY_GROUP = [[0], [1], [2, 3], [4, 5, 6], [7], [8, 9], [10], [11], [12], [13], [14], [15, 16], [17], [18],
                      [19], [20], [21], [22], [23, 24], [25], [26], [27, 28]]


def _sorted_kmers(k):
    sc = SequenceCollection(sequence_list=SEQ_LIST_2, strands_to_load="forward")
    km = Kmers(sc, min_kmer_len=k, max_kmer_len=k, source_strand="forward")
    km.sort()
    return sc, km


def _python_groups(k):
    """Plain Python: every k-mer of the fixture, sorted, grouped (what the reference's tests enumerate)."""
    kmers = sorted(seq[i:i + k] for _, seq in SEQ_LIST_2 for i in range(len(seq) - k + 1))
    return [list(g) for _, g in itertools.groupby(kmers)]


def test_sorted_3mers_and_groups():
    sc, km = _sorted_kmers(3)
    assert [km.get_kmer_str(i, 3) for i in range(len(km))] == SORTED_3MERS
    groups = {}
    for kmer_num, n_yield, size in km.get_kmers(3):
        groups.setdefault(SORTED_3MERS[kmer_num], []).append(kmer_num)
        assert size == n_yield
    assert sorted(groups.values()) == KMER_NUMS_BY_GROUP
    offsets, sizes = km.get_kmer_groups(3)
    assert [int(s) for s in sizes] == [len(g) for g in KMER_NUMS_BY_GROUP]
    assert [int(o) for o in offsets] == [g[0] for g in KMER_NUMS_BY_GROUP]


def test_3mer_counts_with_group_size_limits():
    _, km = _sorted_kmers(3)
    answers = {(1, 1): 16, (2, 2): 10, (3, 3): 3, (2, 3): 13, (1, 2): 26, (1, None): 29}
    for (lo, hi), want in answers.items():
        assert km.get_kmer_count(kmer_len=3, min_group_size=lo, max_group_size=hi) == want


def test_1mer_and_5mer_group_counts():
    _, km = _sorted_kmers(1)
    hist, total = km.get_kmer_group_counts(kmer_len=1, max_counts_bin=15)
    assert total == 35
    assert hist.tolist() == [0, 0, 0, 0, 0, 0, 0, 1, 2, 0, 0, 0, 1, 0, 0, 0]     # C=7, A=8, G=8, T=12
    _, km5 = _sorted_kmers(5)
    hist5, total5 = km5.get_kmer_group_counts(kmer_len=5, max_counts_bin=4)
    assert total5 == 23 and hist5.tolist() == [0, 23, 0, 0, 0]


@pytest.mark.parametrize("max_counts_bin", [1, 2, 3, 4, 10])
@pytest.mark.parametrize("k", [1, 2, 3, 4, 8])
def test_histogram_against_python_grouping(k, max_counts_bin):
    _, km = _sorted_kmers(k)
    sizes = [len(g) for g in _python_groups(k)]
    want = np.zeros(max_counts_bin + 1, dtype=np.int64)
    for s in sizes:
        want[min(s, max_counts_bin)] += 1
    hist, total = km.get_kmer_group_counts(kmer_len=k, max_counts_bin=max_counts_bin)
    assert total == sum(sizes) and np.array_equal(hist, want)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 8])
def test_group_generator_sweeps(k):
    _, km = _sorted_kmers(k)
    groups = _python_groups(k)
    for min_group, max_group, first_n in itertools.product((1, 2, 3, 4), (1, 2, 3, 4, 7, None),
                                                           (1, 2, 3, 4, 7, None)):
        if max_group is not None and max_group < min_group:
            continue
        want, pos = [], 0
        for g in groups:
            size = len(g)
            if size >= min_group and (max_group is None or size <= max_group):
                n_yield = size if first_n is None else min(size, first_n)
                want += [(pos + i, n_yield, size) for i in range(n_yield)]
            pos += size
        got = list(km.get_kmers(k, min_group_size=min_group, max_group_size=max_group, yield_first_n=first_n))
        assert got == want, (k, min_group, max_group, first_n)


def test_module_level_seams_the_way_the_reference_calls_them():
    """kmers.py:1072, :1166 (histogram) and :959-990 (generator): arrays in, not a Kmers object."""
    sc, km = _sorted_kmers(3)
    sba, idx = sc.forward_sba, km.kmer_sba_start_indices
    cmp3 = gk.get_compare_sba_kmers_func(3)
    hist, total = gk.get_kmer_group_size_hist(sba, "forward", 3, idx, cmp3, gk.kmer_filter_keep_all, 1, None, 10)
    assert total == 29 and hist.tolist() == [0, 16, 5, 1, 0, 0, 0, 0, 0, 0, 0]
    hist2, total2 = gk.get_kmer_group_size_hist(sba, "forward", 3, idx, cmp3, gk.kmer_filter_keep_all, 2, 3, 2)
    assert total2 == 13 and hist2.tolist() == [0, 0, 6]
    # unsorted semantics: every k-mer is its own group (kmers.py:1061-1064)
    init = Kmers(sc, 3, 3).kmer_sba_start_indices
    hist3, total3 = gk.get_kmer_group_size_hist(sba, "forward", 3, init, gk.compare_sba_kmers_always_less_than,
                                                gk.kmer_filter_keep_all, 1, None, 4)
    assert total3 == 29 and hist3.tolist() == [0, 29, 0, 0, 0]
    minimal = list(gk.kmer_info_by_group_generator(sba, "forward", 3, idx, cmp3, gk.kmer_filter_keep_all,
                                                   gk.get_kmer_info_minimal, 2, 3, 2))
    assert minimal == [(2, 2, 2), (3, 2, 2), (4, 2, 3), (5, 2, 3), (8, 2, 2), (9, 2, 2), (15, 2, 2), (16, 2, 2),
                       (23, 2, 2), (24, 2, 2), (27, 2, 2), (28, 2, 2)]
    sizes = list(gk.kmer_info_by_group_generator(sba, "forward", 3, idx, cmp3, gk.kmer_filter_keep_all,
                                                 gk.get_kmer_info_group_size_only, 1, None, 1))
    assert sizes == [len(g) for g in KMER_NUMS_BY_GROUP]
    full = km.generate_get_kmer_info_func(one_based_seq_index=True)
    rows = list(gk.kmer_info_by_group_generator(sba, "forward", 3, idx, cmp3, gk.kmer_filter_keep_all, full, 3, 3))
    assert [(r[1], r[2], r[3], r[4], r[5], r[6]) for r in rows] == [
        ("+", "chr1", 6, 3, 3, 3), ("+", "chr2", 10, 3, 3, 3), ("+", "chr3", 4, 3, 3, 3)]          # ATT x 3
    with pytest.raises(ValueError, match="max_counts_bin"):
        gk.get_kmer_group_size_hist(sba, "forward", 3, idx, cmp3, gk.kmer_filter_keep_all, 1, None, 0)
    with pytest.raises(ValueError, match="yield_first_n"):
        list(gk.kmer_info_by_group_generator(sba, "forward", 3, idx, cmp3, gk.kmer_filter_keep_all,
                                             gk.get_kmer_info_minimal, 1, None, 0))


def test_is_less_than_is_the_order_sort_produces():
    """kmers.py:1654-1731 with break_ties=True, all pairs of the fixture, against the GPU order."""
    sc, km = _sorted_kmers(3)
    lt = km.get_is_less_than_func(validate_kmers=True, break_ties=True)
    order = [int(v) for v in km.kmer_sba_start_indices]
    for a, b in zip(order[:-1], order[1:]):
        assert lt(a, b) and not lt(b, a)
    lt_plain = km.get_is_less_than_func(validate_kmers=True, break_ties=False)
    assert not lt_plain(order[2], order[3]) and not lt_plain(order[3], order[2])       # ATC == ATC
    with pytest.raises(AssertionError, match="less than min_kmer_len"):
        lt(9, 0)                                                                        # 'G$': too short
