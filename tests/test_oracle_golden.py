"""
Pins the CPU oracle (oracle/gk_oracle.c and oracle/oracle_np.py) to the real reference: every
golden case in tests/golden was produced by running mrperkett/genome-kmers itself
(tests/golden/make_golden.py).  CPU only.
"""
import numpy as np
import pytest

import oracle
from oracle import oracle_np
from conftest import dense_hist, golden_case, golden_case_names

ALL = golden_case_names()
COMP = str.maketrans("ACGTRYSWKMBDHVN", "TGCAYRSWMKVHDBN")


def _filter(spec):
    if spec is None:
        return (oracle.FILTER_KEEP_ALL, 0, 0, 0)
    assert spec[0] == "no_ambiguous"
    return (oracle.FILTER_NO_AMBIGUOUS, spec[1], 0, 0)


def _build(case):
    sba, starts = oracle.build_sba([s for _, s in case["seq_list"]])
    starts = starts.astype(np.uint64)
    if case["strands"] == "both":
        sba, starts = oracle.both_strands(sba, starts)
    return sba, starts


@pytest.mark.parametrize("name", ALL)
def test_sba_layout_matches_reference(name):
    """A1/A2: byte array and segment starts (sequence_collection.py:663-726, :42-73, :905-928)."""
    case = golden_case(name)
    sba, starts = _build(case)
    assert np.array_equal(sba, case["sba"])
    assert np.array_equal(starts, case["seg_starts"].astype(np.uint64))


@pytest.mark.parametrize("name", ALL)
def test_init_indices_match_reference(name):
    """A3: kmers.py:789-835."""
    case = golden_case(name)
    got = oracle.init_indices(case["seg_starts"], len(case["sba"]), case["min_len"])
    assert np.array_equal(got, case["init"].astype(np.uint64))


@pytest.mark.parametrize("name", ALL)
def test_c_sort_matches_reference(name):
    """A4+A5: comparator + quicksort, canonical tie order (kmers.py:306-397, :1624-1731)."""
    case = golden_case(name)
    got = oracle.sort_indices(case["sba"], case["init"], case["min_len"], case["max_len"])
    assert np.array_equal(got, case["sorted"].astype(np.uint64))
    if len(case["init"]) > 3000:
        par = oracle.sort_indices(case["sba"], case["init"], case["min_len"], case["max_len"],
                                  threads=4)
        assert np.array_equal(par, got)


@pytest.mark.parametrize("name", ALL)
def test_numpy_sort_matches_reference(name):
    case = golden_case(name)
    got = oracle_np.sort_indices(case["sba"], case["init"], case["seg_starts"], case["max_len"])
    assert np.array_equal(got, case["sorted"].astype(np.uint64))


@pytest.mark.parametrize("name", ALL)
def test_group_counts_match_reference(name):
    """A6+A7: group walk + histogram incl. filters and group-size limits (kmers.py:454-648)."""
    case = golden_case(name)
    for qu, ans in zip(case["queries"], case["answers"]):
        hist, total = oracle.group_hist(
            case["sba"], case["sorted"], qu["kmer_len"], True, _filter(qu["filter"]),
            qu["min_group"], qu["max_group"], qu["max_bin"])
        assert total == ans["total"], qu
        assert np.array_equal(hist, dense_hist(ans, qu["max_bin"])), qu
        if qu["filter"] is None:
            sizes = oracle_np.group_sizes(case["sba"], case["sorted"], case["seg_starts"],
                                          qu["kmer_len"])
            h2, t2 = oracle_np.group_hist(sizes, qu["min_group"], qu["max_group"], qu["max_bin"])
            assert t2 == ans["total"] and np.array_equal(h2, hist), qu


def test_unsorted_count_is_number_of_kmers():
    """get_kmer_count on unsorted data: every k-mer is its own group (kmers.py:1061-1064)."""
    case = golden_case("sl2_k3")
    hist, total = oracle.group_hist(case["sba"], case["init"], 3, sorted_=False, max_bin=4)
    assert total == len(case["init"]) and hist[1] == len(case["init"])


def test_reference_known_answers():
    """Known answers the reference's own tests hold (tests/test_kmers.py:1889-1944)."""
    case = golden_case("sl2_k1")
    hist, total = oracle.group_hist(case["sba"], case["sorted"], 1, max_bin=20)
    assert total == 35
    assert sorted(np.repeat(np.arange(21), hist).tolist()) == [7, 8, 8, 12]  # C, A, G, T
    case = golden_case("sl2_k5")
    hist, total = oracle.group_hist(case["sba"], case["sorted"], 5, max_bin=5)
    assert total == 23 and hist[1] == 23


def test_reference_golden_revcomp():
    """tests/test_sequence_collection.py:40-44."""
    sba, starts = oracle.build_sba(["ATCGAATTAG", "GGATCTTGCATT", "GTGATTGACCCCT"])
    assert bytes(sba) == b"ATCGAATTAG$GGATCTTGCATT$GTGATTGACCCCT"
    assert starts.tolist() == [0, 11, 24]
    assert bytes(oracle.revcomp(sba)) == b"AGGGGTCAATCAC$AATGCAAGATCC$CTAATTCGAT"
    assert oracle.revcomp_seg_starts(starts, len(sba)).tolist() == [0, 14, 27]


def test_sorted_3mers_match_reference_docs():
    """Sorted 3-mer list of seq_list_2 (tests/test_kmers.py:984-1014, docs/overview.rst:46-74)."""
    case = golden_case("sl2_k3")
    sba = case["sba"]
    kmers = [bytes(sba[i:i + 3]).decode() for i in case["sorted"]]
    assert kmers == sorted(kmers)
    assert kmers[:5] == ["AAT", "ACC", "AGG", "ATC", "ATC"] or kmers[0] == "AAT"
    assert len(kmers) == 29


# ---------------------------------------------------------------------------------------------------
# the oracle's group walk and all six filters against the tuples the real reference's get_kmers yielded
# (tests/golden/golden_get_kmers.json, tests/golden/make_golden_get_kmers.py)
# ---------------------------------------------------------------------------------------------------
def _oracle_filter(spec):
    if spec is None:
        return (oracle.FILTER_KEEP_ALL, 0, 0, 0)
    kind, args = spec[0], spec[1:]
    if kind == "no_ambiguous":
        return (oracle.FILTER_NO_AMBIGUOUS, args[0], 0, 0)
    if kind == "min_length":
        return (oracle.FILTER_MIN_LENGTH, args[0], 0, 0)
    if kind == "homopolymer":
        return (oracle.FILTER_HOMOPOLYMER, args[0], args[1], 0)
    if kind == "gc":  # the reference converts the fractions to counts once (kmers.py:143-144)
        lo, hi, k = args
        return (oracle.FILTER_GC_COUNT, int(np.ceil(k * lo)), int(np.floor(k * hi)), k)
    if kind == "ngg_pam":
        return (oracle.FILTER_NGG_PAM, 0, 0, 0)
    raise ValueError(spec)


from conftest import get_kmers_entries, get_kmers_entry_id  # noqa: E402


@pytest.mark.parametrize("entry", get_kmers_entries(), ids=get_kmers_entry_id)
def test_oracle_group_walk_matches_reference_get_kmers(entry):
    """A6 + N1: groups (first member, size) of kmer_info_by_group_generator (kmers.py:523-648) with every
    built-in filter (kmers.py:14-259), group limits and the unsorted mode."""
    case, qu = golden_case(entry["case"]), entry["query"]
    sba, _ = _build(case)
    idx = case["sorted"] if qu["sorted"] else case["init"]
    _, total, first, size = oracle.group_hist(
        sba, idx, qu["kmer_len"], sorted_=qu["sorted"], filt=_oracle_filter(qu["filter"]),
        min_group=qu["min_group"], max_group=qu["max_group"], max_bin=8, want_groups=True)
    # the reference's tuples end with (group_size_yielded, group_size_total); a group's first tuple carries
    # the kmer_num of its first passing member
    want_first, want_size, left = [], [], 0
    for t in entry["tuples"]:
        if left == 0:
            want_first.append(t[0])
            want_size.append(t[-1])
            left = t[-2]
        left -= 1
    assert first.tolist() == want_first
    assert size.tolist() == want_size
    assert total == sum(want_size)


# ---------------------------------------------------------------------------------------------------
# seeded random cross-check of the two independent oracles (comparator quicksort in C against the 4-bit rank
# encoding + lexsort in NumPy) on inputs the golden cases do not enumerate: many short records, every IUPAC
# letter, fixed and variable window lengths, both strands
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(24))
def test_c_oracle_and_numpy_oracle_agree_on_random_collections(seed):
    rng = np.random.default_rng(1000 + seed)
    alphabet = np.frombuffer(b"ACGT" * 6 + b"RYSWKMBDHVN", dtype=np.uint8)     # mostly A/C/G/T
    n_rec = int(rng.integers(1, 9))
    lengths = rng.integers(1, 120, n_rec) if seed % 3 else rng.integers(40, 400, n_rec)
    records = []
    for n in lengths:
        seq = alphabet[rng.integers(0, len(alphabet), int(n))].copy()
        if n > 30 and seed % 2:
            seq[5:5 + int(n) // 3] = ord("N")                                  # an N run
        if n > 20 and seed % 4 == 0:
            seq[:] = seq[int(rng.integers(0, n))]                              # a homopolymer record
        records.append(seq)
    sba, starts = oracle.build_sba(records)
    starts = starts.astype(np.uint64)
    if seed % 5 == 0:
        sba, starts = oracle.both_strands(sba, starts)
    shortest = int(min(len(r) for r in records))
    min_len = int(rng.integers(1, shortest + 1))
    max_len = [min_len, None, min_len + int(rng.integers(0, 40))][seed % 3]
    init = oracle.init_indices(starts, len(sba), min_len)
    got_c = oracle.sort_indices(sba, init, min_len, max_len, break_ties=True, threads=1 + seed % 3)
    got_np = oracle_np.sort_indices(sba, init, starts, max_len)
    assert np.array_equal(got_c, got_np)
    for kmer_len in {min_len, max_len, max(1, min_len // 2)}:
        hist_c, total_c, first, size = oracle.group_hist(sba, got_c, kmer_len, max_bin=12, want_groups=True)
        sizes_np = oracle_np.group_sizes(sba, got_c, starts, kmer_len)
        hist_np, total_np = oracle_np.group_hist(sizes_np, max_bin=12)
        assert size.tolist() == sizes_np.tolist()
        assert total_c == total_np == len(init) and np.array_equal(hist_c, hist_np)
