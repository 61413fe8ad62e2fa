"""
GPU parity tests proper: the drop-in API (SequenceCollection + Kmers, which call libgkb200 through
the C ABI) against (1) the golden vectors produced by the real reference and (2) the CPU oracle on
larger seeded inputs, (3) size-independent properties at BASELINE.json's full sizes.
Bit-exact: start-index order with ties ascending (the reference's break_ties=True order).
"""
import os

import numpy as np
import pytest

import oracle
from conftest import (dense_hist, expected_get_kmers_tuples, filter_from_spec, get_kmers_entries,
                      get_kmers_entry_id, golden_case, golden_case_names, is_fixed_k)
from genome_kmers.kmers import (Kmers, gen_kmer_gc_content_filter_func, gen_kmer_homopolymer_filter_func,
                                gen_no_ambiguous_bases_filter, kmer_filter_keep_all)
from genome_kmers.sequence_collection import SequenceCollection
import gpu_utils as gu

pytestmark = pytest.mark.gpu

FIXED = golden_case_names(is_fixed_k)
VARIABLE = golden_case_names(lambda c: not is_fixed_k(c))


def _filter(spec):
    return kmer_filter_keep_all if spec is None else gen_no_ambiguous_bases_filter(spec[1])


def _kmers_for(case):
    sc = SequenceCollection(sequence_list=[tuple(r) for r in case["seq_list"]],
                            strands_to_load=case["strands"])
    return sc, Kmers(sc, min_kmer_len=case["min_len"], max_kmer_len=case["max_len"],
                     source_strand=case["strands"])


@pytest.mark.parametrize("name", FIXED)
def test_golden_init_sort_and_counts(name):
    case = golden_case(name)
    sc, km = _kmers_for(case)
    assert len(km) == case["n_kmers"]
    assert np.array_equal(km.kmer_sba_start_indices, case["init"])          # A3
    assert km.kmer_sba_start_indices.dtype == (np.uint64 if os.environ.get("GK_FORCE_IDX64") else np.uint32)
    try:
        km.sort()                                                             # A4 + A5
    except NotImplementedError as exc:
        pytest.skip(f"not on the GPU path yet: {exc}")
    assert km._is_sorted
    got = km.kmer_sba_start_indices
    bad = np.flatnonzero(got != case["sorted"])
    assert len(bad) == 0, f"{name}: first mismatches at {bad[:8]}: got {got[bad[:8]]} want {case['sorted'][bad[:8]]}"
    for qu, ans in zip(case["queries"], case["answers"]):                    # A6 + A7
        hist, total = km.get_kmer_group_counts(
            qu["kmer_len"], kmer_filter_func=_filter(qu["filter"]), min_group_size=qu["min_group"],
            max_group_size=qu["max_group"], max_counts_bin=qu["max_bin"])
        assert total == ans["total"], (name, qu)
        assert np.array_equal(hist, dense_hist(ans, qu["max_bin"])), (name, qu)
        assert km.get_kmer_count(qu["kmer_len"], _filter(qu["filter"]), qu["min_group"],
                                 qu["max_group"]) == ans["total"]


HYBRID_CASES = [n for n in FIXED if n.split("_k")[0] in
                ("rand5k", "iupac4k", "lowcomplex", "edgeT", "repeat", "repeatN", "iupac4k_both", "rand120k",
                 "rand120kN_both", "lowcomplex_both", "repeat_both")]


@pytest.mark.parametrize("prefix_bits", [8, 16, 24])
@pytest.mark.parametrize("name", HYBRID_CASES)
def test_golden_with_forced_prefix_sort(name, prefix_bits, monkeypatch):
    """The hybrid sort (LSD passes over the top bits + in-place tie repair + long-run fallback) must give
    the same order as plain LSD whatever the prefix width; tiny prefixes force the long-run path."""
    monkeypatch.setenv("GK_SORT_PREFIX_BITS", str(prefix_bits))
    case = golden_case(name)
    sc, km = _kmers_for(case)
    km.sort()
    got = km.kmer_sba_start_indices
    bad = np.flatnonzero(got != case["sorted"])
    assert len(bad) == 0, f"{name}: first mismatches at {bad[:8]}"
    for qu, ans in zip(case["queries"], case["answers"]):
        hist, total = km.get_kmer_group_counts(
            qu["kmer_len"], kmer_filter_func=_filter(qu["filter"]), min_group_size=qu["min_group"],
            max_group_size=qu["max_group"], max_counts_bin=qu["max_bin"])
        assert total == ans["total"], (name, qu)
        assert np.array_equal(hist, dense_hist(ans, qu["max_bin"])), (name, qu)


@pytest.mark.parametrize("mode", ["0", "auto", "16", "32"])
def test_hybrid_and_plain_sort_agree_on_skewed_input(mode, monkeypatch):
    """Low-complexity + repeats + N runs: long prefix runs, some out of order, some uniform."""
    if mode == "0":
        monkeypatch.setenv("GK_SORT_HYBRID", "0")
    elif mode != "auto":
        monkeypatch.setenv("GK_SORT_PREFIX_BITS", mode)
    rng = np.random.default_rng(7)
    unit = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 700)]
    parts = []
    for i in range(60):                       # 60 diverged copies of one 700-mer + poly-A + N runs
        copy = unit.copy()
        copy[rng.integers(0, 700, 3)] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 3)]
        parts += [copy, np.full(int(rng.integers(5, 80)), ord("A"), dtype=np.uint8)]
        if i % 7 == 0:
            parts.append(np.full(int(rng.integers(40, 400)), ord("N"), dtype=np.uint8))
    parts.append(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 120_000)])
    seq = np.concatenate(parts)
    _oracle_compare([("chr0", seq[:len(seq) // 2]), ("chr1", seq[len(seq) // 2:])], 31, "both")


@pytest.mark.parametrize("name", ["rand5k_k21", "iupac4k_both_k31", "lowcomplex_k31", "rand120kN_both_k31", "sl2_k3"])
def test_golden_with_64bit_start_indices(name, monkeypatch):
    """Byte arrays of 2^32 or more positions switch the start indices to uint64 (the reference refuses,
    kmers.py:805-808); GK_FORCE_IDX64 drives small inputs through that path."""
    monkeypatch.setenv("GK_FORCE_IDX64", "1")
    case = golden_case(name)
    sc, km = _kmers_for(case)
    km.sort()
    got = km.kmer_sba_start_indices
    assert got.dtype == np.uint64
    assert np.array_equal(got, case["sorted"].astype(np.uint64))
    for qu, ans in zip(case["queries"], case["answers"]):
        hist, total = km.get_kmer_group_counts(
            qu["kmer_len"], kmer_filter_func=_filter(qu["filter"]), min_group_size=qu["min_group"],
            max_group_size=qu["max_group"], max_counts_bin=qu["max_bin"])
        assert total == ans["total"] and np.array_equal(hist, dense_hist(ans, qu["max_bin"])), (name, qu)


def test_seeded_random_with_64bit_start_indices(monkeypatch):
    monkeypatch.setenv("GK_FORCE_IDX64", "1")
    rng = np.random.default_rng(64)
    recs = gu.random_genome(rng, 700_000, 4, n_runs=5, run_lo=50, run_hi=3000, n_scatter=20)
    _oracle_compare(recs, 31, "both")


@pytest.mark.parametrize("k", [33, 40, 64, 100])
def test_long_kmers_with_64bit_start_indices(k, monkeypatch):
    """k-mers longer than one key word on a byte array of 2^32 or more positions: the rank table of the prefix
    doubling is 32-bit, so the tied members are ordered by the symbols read from the bytes (word rounds)."""
    monkeypatch.setenv("GK_FORCE_IDX64", "1")
    rng = np.random.default_rng(640 + k)
    recs = gu.random_genome(rng, 300_000, 3, n_runs=4, run_lo=50, run_hi=3000, n_scatter=15)
    for _ in range(80):   # copies: k-mers that agree on the first 31 symbols and differ later, or never
        seq = recs[int(rng.integers(0, len(recs)))][1]
        ln = int(rng.integers(33, 500))
        src, dst = (int(v) for v in rng.integers(0, len(seq) - ln, 2))
        seq[dst:dst + ln] = seq[src:src + ln].copy()
    _oracle_compare(recs, k, "both")


@pytest.mark.parametrize("name", VARIABLE)
def test_golden_variable_length_modes(name):
    """sort() with min_kmer_len != max_kmer_len."""
    case = golden_case(name)
    sc, km = _kmers_for(case)
    assert np.array_equal(km.kmer_sba_start_indices, case["init"])
    try:
        km.sort()
    except NotImplementedError:
        pytest.skip("variable-length / suffix order is not on the GPU path yet")
    assert np.array_equal(km.kmer_sba_start_indices, case["sorted"])
    for qu, ans in zip(case["queries"], case["answers"]):
        hist, total = km.get_kmer_group_counts(qu["kmer_len"], max_counts_bin=qu["max_bin"],
                                               min_group_size=qu["min_group"],
                                               max_group_size=qu["max_group"])
        assert total == ans["total"] and np.array_equal(hist, dense_hist(ans, qu["max_bin"]))


@pytest.mark.parametrize("name", ["sl2_k3", "rand5k_k21", "iupac4k_both_k31", "rand120kN_both_k31"])
def test_one_shot_host_entry_point(name):
    """gk_sort_count_host: host buffers in, host buffers out -- what a non-Python host binds."""
    import ctypes

    from genome_kmers import _native

    case = golden_case(name)
    lib = _native.lib()
    sc = SequenceCollection(sequence_list=[tuple(r) for r in case["seq_list"]], strands_to_load="forward")
    sba = np.ascontiguousarray(sc.forward_sba)
    starts = np.ascontiguousarray(sc._forward_sba_seg_starts, dtype=np.uint64)
    k = case["min_len"]
    n = case["n_kmers"]
    idx = np.zeros(n, dtype=np.uint32)
    max_bin = 64
    hist = np.zeros(max_bin + 1, dtype=np.int64)
    total, n_out = ctypes.c_int64(0), ctypes.c_uint64(0)
    stats = _native.GkSortStats()
    _native.check(lib.gk_sort_count_host(
        _native.host_ptr(sba), len(sba), _native.host_ptr(starts), len(starts), k,
        2 if case["strands"] == "both" else 0, 4, _native.host_ptr(idx), max_bin, _native.host_ptr(hist),
        ctypes.byref(total), ctypes.byref(n_out), ctypes.byref(stats)))
    assert n_out.value == n and total.value == n
    assert np.array_equal(idx, case["sorted"])
    qu = next((q, a) for q, a in zip(case["queries"], case["answers"])
              if q["kmer_len"] == k and q["filter"] is None and q["min_group"] == 1 and q["max_group"] is None)
    want = dense_hist(qu[1], qu[0]["max_bin"])
    m = min(len(want), len(hist)) - 1
    assert np.array_equal(hist[:m], want[:m])


def test_counts_with_other_kmer_len_and_unsorted():
    case = golden_case("rand5k_k21_count11")
    sc, km = _kmers_for(case)
    assert km.get_kmer_count(21) == case["n_kmers"]                          # unsorted: kmers.py:1061-1064
    with pytest.raises(AssertionError):
        km.get_kmer_group_counts(21)
    with pytest.raises(ValueError):
        km.get_kmer_count(21, min_group_size=2)
    km.sort()
    for qu, ans in zip(case["queries"], case["answers"]):
        hist, total = km.get_kmer_group_counts(qu["kmer_len"], min_group_size=qu["min_group"],
                                               max_group_size=qu["max_group"], max_counts_bin=qu["max_bin"])
        assert total == ans["total"] and np.array_equal(hist, dense_hist(ans, qu["max_bin"]))


def test_group_table_is_the_unique_kmer_set():
    case = golden_case("sl2_k3")
    sc, km = _kmers_for(case)
    km.sort()
    offsets, sizes = km.get_kmer_groups(3)
    idx = km.kmer_sba_start_indices
    kmers = [sc.forward_sba[i:i + 3].tobytes().decode() for i in idx]
    uniq = [kmers[o] for o in offsets]
    assert uniq == sorted(set(kmers))
    assert [kmers.count(u) for u in uniq] == sizes.tolist()
    minimal = list(km.get_kmers(3, min_group_size=2, yield_first_n=1))
    assert [kmers[num] for num, _, _ in minimal] == [u for u, s in zip(uniq, sizes) if s >= 2]


@pytest.mark.parametrize("name,make_filter", [
    ("iupac4k_k21", lambda k: gen_no_ambiguous_bases_filter(k)),
    ("rand5k_k21", lambda k: gen_kmer_gc_content_filter_func(0.4, 0.6, k)),
    ("lowcomplex_k8", lambda k: gen_kmer_homopolymer_filter_func(3, k)),
    ("rand5k_k3", lambda k: gen_kmer_gc_content_filter_func(0.3, 0.7, k)),
])
def test_get_kmers_with_device_filters(name, make_filter):
    """get_kmers(filter): a failing k-mer is skipped and the next passing one is compared with the previous
    PASSING one (kmers.py:586-601); checked against a host walk of the sorted order."""
    case = golden_case(name)
    sc, km = _kmers_for(case)
    k = case["min_len"]
    filt = make_filter(k)
    sba = sc.forward_sba
    # unsorted: every passing k-mer is its own group
    want = [(num, 1, 1) for num, s in enumerate(km.kmer_sba_start_indices) if filt(sba, "forward", int(s))]
    assert list(km.get_kmers(k, kmer_filter_func=filt)) == want
    km.sort()
    groups, prev = [], None
    for num, s in enumerate(km.kmer_sba_start_indices):
        if not filt(sba, "forward", int(s)):
            continue
        kmer = sba[int(s):int(s) + k].tobytes()
        if prev is not None and kmer == prev:
            groups[-1].append(num)
        else:
            groups.append([num])
            prev = kmer
    assert len(groups) > 0
    for min_g, max_g, first_n in [(1, None, None), (2, None, 1), (1, 3, 2), (1, 1, None)]:
        want = [(num, min(len(g), first_n or len(g)), len(g)) for g in groups
                if len(g) >= min_g and (max_g is None or len(g) <= max_g) for num in g[:first_n or len(g)]]
        got = list(km.get_kmers(k, kmer_filter_func=filt, min_group_size=min_g, max_group_size=max_g,
                                yield_first_n=first_n))
        assert got == want, (min_g, max_g, first_n)
    # and the counts agree with the same walk
    assert km.get_kmer_count(k, filt) == sum(len(g) for g in groups)


@pytest.mark.parametrize("entry", get_kmers_entries(), ids=get_kmers_entry_id)
def test_golden_get_kmers(entry):
    """Row N2: Kmers.get_kmers against tuples the real reference yielded (tests/golden/make_golden_get_kmers.py)
    -- minimum and full info, 0/1-based, filters, group limits, yield_first_n, sorted and unsorted.
    source_strand='both': the reference ran its forward path over forward + '<name>_rc' records; a k-mer of a
    reverse-complemented record is ('-', name, forward sequence index) here, the reference's own convention for
    its reverse-complement strand (sequence_collection.py:101-153)."""
    case, qu = golden_case(entry["case"]), entry["query"]
    sc, km = _kmers_for(case)
    if qu["sorted"]:
        km.sort()
        assert np.array_equal(km.kmer_sba_start_indices, case["sorted"])
    got = list(km.get_kmers(qu["kmer_len"], one_based_seq_index=qu["one_based"],
                            kmer_filter_func=filter_from_spec(qu["filter"]), kmer_info_to_yield=qu["info"],
                            min_group_size=qu["min_group"], max_group_size=qu["max_group"],
                            yield_first_n=qu["first_n"]))
    assert got == expected_get_kmers_tuples(entry, case)


def _oracle_compare(records, k, strands, threads=8, filt_k=None):
    sc = SequenceCollection.from_arrays(records, strands_to_load=strands)
    km = Kmers(sc, k, k, source_strand=strands)
    km.sort()
    got = km.kmer_sba_start_indices.astype(np.uint64)
    sba, starts = sc.forward_sba, sc._forward_sba_seg_starts.astype(np.uint64)
    if strands == "both":
        sba, starts = oracle.both_strands(sba, starts)
    init = oracle.init_indices(starts, len(sba), k)
    want = oracle.sort_indices(sba, init, k, k, threads=threads)
    bad = np.flatnonzero(got != want)
    assert len(bad) == 0, f"first mismatches at {bad[:8]}: got {got[bad[:8]]} want {want[bad[:8]]}"
    hist, total = km.get_kmer_group_counts(k, max_counts_bin=1000)
    o_hist, o_total = oracle.group_hist(sba, want, k, max_bin=1000)
    assert total == o_total and np.array_equal(hist, o_hist)
    fk = filt_k or k
    hist, total = km.get_kmer_group_counts(fk, gen_no_ambiguous_bases_filter(fk), max_counts_bin=1000)
    o_hist, o_total = oracle.group_hist(sba, want, fk, filt=(oracle.FILTER_NO_AMBIGUOUS, fk, 0, 0), max_bin=1000)
    assert total == o_total and np.array_equal(hist, o_hist)
    return km, sba, want


@pytest.mark.parametrize("k,strands,n_bases,n_rec,runs", [
    (21, "forward", 400_000, 1, 0),
    (31, "both", 300_000, 5, 6),
    (12, "both", 500_000, 3, 4),      # 4^12 < windows: heavy duplication
    (31, "forward", 1_000_000, 10, 20),
    (8, "forward", 200_000, 2, 3),
    (32, "forward", 250_000, 3, 0),
])
def test_seeded_random_vs_oracle(k, strands, n_bases, n_rec, runs):
    rng = np.random.default_rng(42 + k + n_bases)
    recs = gu.random_genome(rng, n_bases, n_rec, n_runs=runs, run_lo=50, run_hi=3000,
                            n_scatter=30 if runs else 0)
    _oracle_compare(recs, k, strands)


@pytest.mark.parametrize("min_len,max_len,strands,n_bases,n_rec,runs", [
    (5, None, "forward", 60_000, 4, 3),       # suffix order inside each record
    (1, None, "forward", 30_000, 7, 0),
    (3, 20, "forward", 200_000, 5, 4),        # one key word, windows shorter than the key near record ends
    (1, 31, "both", 150_000, 3, 2),
    (10, 40, "forward", 200_000, 5, 4),       # two doubling rounds, short windows dropped at the end
    (20, 100, "both", 120_000, 6, 3),
])
def test_variable_length_modes_vs_oracle(min_len, max_len, strands, n_bases, n_rec, runs):
    """sort() with min_kmer_len < max_kmer_len / max_kmer_len None (kmers.py:360-378), SURVEY.md 8f N5."""
    rng = np.random.default_rng(7 * min_len + (max_len or 0) + n_bases)
    recs = gu.random_genome(rng, n_bases, n_rec, n_runs=runs, run_lo=20, run_hi=400, n_scatter=10 if runs else 0)
    # plant repeats so that many windows stay tied for hundreds of symbols
    recs = [(nm, seq.copy()) for nm, seq in recs]
    unit = recs[0][1][100:700].copy()
    for nm, seq in recs[1:]:
        seq[50:650] = unit
        seq[-300:] = unit[:300]               # identical record tails: suffixes equal up to the terminator
    sc = SequenceCollection.from_arrays(recs, strands_to_load=strands)
    km = Kmers(sc, min_len, max_len, source_strand=strands)
    km.sort()
    got = km.kmer_sba_start_indices.astype(np.uint64)
    sba, starts = sc.forward_sba, sc._forward_sba_seg_starts.astype(np.uint64)
    if strands == "both":
        sba, starts = oracle.both_strands(sba, starts)
    want = oracle.sort_indices(sba, oracle.init_indices(starts, len(sba), min_len), min_len, max_len, threads=8)
    bad = np.flatnonzero(got != want)
    assert len(bad) == 0, f"first mismatches at {bad[:8]}: got {got[bad[:8]]} want {want[bad[:8]]}"
    for kmer_len in (max_len, min_len):
        hist, total = km.get_kmer_group_counts(kmer_len, max_counts_bin=500)
        o_hist, o_total = oracle.group_hist(sba, want, kmer_len or 0, max_bin=500)
        assert total == o_total and np.array_equal(hist, o_hist), kmer_len


def _rec(text):
    return np.frombuffer(text.encode(), dtype=np.uint8).copy()


@pytest.mark.parametrize("label,records,k,strands", [
    ("one k-mer", [("a", _rec("ACGTACGTAC"))], 10, "forward"),
    ("record of exactly k next to a long one", [("a", _rec("ACGTTGCA")), ("b", _rec("ACGTTGCAACGTTGCATT"))], 8, "both"),
    ("poly-A: one pure giant group", [("a", np.full(150_000, ord("A"), dtype=np.uint8))], 31, "forward"),
    ("only N: one ambiguous giant group", [("a", np.full(90_000, ord("N"), dtype=np.uint8))], 21, "both"),
    ("two-letter low complexity", [("a", np.frombuffer(b"AT" * 60_000, dtype=np.uint8).copy()),
                                   ("b", np.frombuffer(b"ATT" * 30_000, dtype=np.uint8).copy())], 31, "both"),
    ("k=32 without ambiguous symbols (no class bit)", [("a", np.frombuffer(b"ACGT", dtype=np.uint8)[
        np.random.default_rng(5).integers(0, 4, 120_000)].copy())], 32, "both"),
    ("k=1", [("a", _rec("ACGTNNRYACGT" * 2000)), ("b", _rec("TTTTGGGG" * 500))], 1, "both"),
    ("every IUPAC letter", [("a", np.frombuffer(b"ACGTRYSWKMBDHVN", dtype=np.uint8)[
        np.random.default_rng(6).integers(0, 15, 80_000)].copy())], 12, "both"),
])
def test_edge_inputs_vs_oracle(label, records, k, strands):
    _oracle_compare(records, k, strands)


def test_small_max_counts_bin_and_group_limits():
    """counts_by_group_size[min(size, max_counts_bin)] with a giant group and tiny tables (kmers.py:514-518)."""
    seq = np.concatenate([np.full(5000, ord("N"), dtype=np.uint8),
                          np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(9).integers(0, 4, 20_000)],
                          np.full(300, ord("A"), dtype=np.uint8)])
    sc = SequenceCollection.from_arrays([("a", seq)], strands_to_load="forward")
    km = Kmers(sc, 8, 8)
    km.sort()
    sba, starts = sc.forward_sba, sc._forward_sba_seg_starts.astype(np.uint64)
    want = oracle.sort_indices(sba, oracle.init_indices(starts, len(sba), 8), 8, 8)
    assert np.array_equal(km.kmer_sba_start_indices.astype(np.uint64), want)
    for max_bin in (1, 2, 7, 5000):
        for min_g, max_g in ((1, None), (2, None), (1, 3), (290, 5000), (3, 3)):
            hist, total = km.get_kmer_group_counts(8, min_group_size=min_g, max_group_size=max_g,
                                                   max_counts_bin=max_bin)
            o_hist, o_total = oracle.group_hist(sba, want, 8, min_group=min_g, max_group=max_g, max_bin=max_bin)
            assert total == o_total and np.array_equal(hist, o_hist), (max_bin, min_g, max_g)
            assert km.get_kmer_count(8, min_group_size=min_g, max_group_size=max_g) == o_total


def test_config1_shape_vs_oracle():
    """BASELINE.json configs[0]: 4.6 Mbp, one record, forward, k=21 (the reference's CPU case)."""
    rng = np.random.default_rng(42)
    recs = gu.random_genome(rng, 4_600_000, 1)
    km, sba, want = _oracle_compare(recs, 21, "forward")
    assert len(km) == 4_599_980


def _check_sorted_properties(km, sc, k, strands):
    """Size-independent checks, all on the device with torch as the checker."""
    torch = gu.torch_mod()
    idx = km.device_start_indices().to(torch.int64)
    n = idx.numel()
    # (1) a permutation of the init set
    fresh = Kmers(sc, k, k, source_strand=strands)
    init = fresh.device_start_indices().to(torch.int64)
    assert torch.equal(torch.sort(idx).values, init)
    # (2) adjacent windows non-decreasing by raw byte order, ties ascending in start
    d_sba = km._d_sba
    chunk = 1 << 22
    carry = None
    n_groups = 0
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk + 1)
        starts = idx[lo:hi]
        win = d_sba[(starts[:, None] + torch.arange(k, device="cuda")[None, :])]
        a, b = win[:-1].to(torch.int16), win[1:].to(torch.int16)
        diff = (a != b)
        first = torch.where(diff.any(dim=1), diff.to(torch.int8).argmax(dim=1), torch.full((len(a),), k - 1, device="cuda"))
        av = a.gather(1, first[:, None]).squeeze(1)
        bv = b.gather(1, first[:, None]).squeeze(1)
        assert bool((av <= bv).all()), "k-mers are not in non-decreasing order"
        tie = ~diff.any(dim=1)
        assert bool((starts[:-1][tie] < starts[1:][tie]).all()), "ties are not in ascending start order"
        n_groups += int((~tie).sum())
    return n_groups + 1


def test_config2_full_size_properties():
    """BASELINE.json configs[1]: 100 Mbp, 10 records, N runs, both strands, k=31 -- the bench workload.
    The oracle cannot reach this size; check permutation, order, tie order and histogram identities."""
    n_bases = int(os.environ.get("GK_TEST_C2_BASES", 100_000_000))
    rng = np.random.default_rng(42)
    recs = gu.random_genome(rng, n_bases, 10, n_runs=200, run_lo=1000, run_hi=100_000)
    sc = SequenceCollection.from_arrays(recs, strands_to_load="both")
    km = Kmers(sc, 31, 31, source_strand="both")
    km.sort()
    n = len(km)
    assert n == 2 * (n_bases - 10 * 30)
    n_groups = _check_sorted_properties(km, sc, 31, "both")
    hist, total = km.get_kmer_group_counts(31)
    assert total == n
    assert int(hist.sum()) == n_groups
    sizes = np.flatnonzero(hist)
    assert int((hist[sizes] * np.minimum(sizes, 1000000)).sum()) <= n
    offsets, group_sizes = km.get_kmer_groups(31)
    assert len(offsets) == n_groups and int(group_sizes.sum()) == n
    # the fast (key-flag) and general (byte comparator) grouping paths agree
    h2, t2 = km.get_kmer_group_counts(30)   # forces the comparator path with another length
    assert t2 == n and int(h2.sum()) <= n_groups
    pure_hist, pure_total = km.get_kmer_group_counts(31, gen_no_ambiguous_bases_filter(31))
    assert pure_total == n - km.last_sort_stats["n_ambiguous"]


# ---- full-size property tests of the other BASELINE.json configs (no CPU oracle reaches these sizes: the order
# ---- is checked on the device against the bytes, gk_index_verify) ------------------------------------------------
def _device_genome(n_bases, n_records, seed, n_runs=0, plant=None):
    """Records generated on the device (torch is the generator here, not the thing under test)."""
    torch = gu.torch_mod()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
    bases = lut[torch.randint(0, 4, (n_bases,), generator=gen, device="cuda")]
    rng = np.random.default_rng(seed)
    for _ in range(n_runs):
        ln = int(np.exp(rng.uniform(np.log(1e3), np.log(1e5))))
        st = int(rng.integers(0, n_bases - ln))
        bases[st:st + ln] = ord("N")
    if plant is not None:
        plant(bases, rng)
    host = bases.cpu().numpy()
    avg = n_bases // n_records
    bounds = [i * avg for i in range(n_records)] + [n_bases]
    return [(f"chr{i}", host[bounds[i]:bounds[i + 1]]) for i in range(n_records)]


def _plant_diverged_repeats(bases, rng):
    """Copies of 3 kb units whose first differences lie 35..90 symbols apart: k-mers of 64 and 100 symbols tie
    on their first 31 symbols and need the prefix-doubling rounds to be separated."""
    torch = gu.torch_mod()
    n = bases.numel()
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    for _ in range(40):
        unit = lut[rng.integers(0, 4, 3000)]
        for _ in range(int(rng.integers(5, 40))):
            copy = unit.copy()
            pos = np.cumsum(rng.integers(35, 90, 60))
            pos = pos[pos < 3000]
            copy[pos] = lut[rng.integers(0, 4, len(pos))]
            st = int(rng.integers(0, n - 3000))
            bases[st:st + 3000] = torch.from_numpy(copy).to("cuda")


@pytest.mark.parametrize("k", [64, 100])
def test_config4_long_kmers_full_size_properties(k):
    """BASELINE.json configs[3]: 100 Mbp, both strands, k = 64 / 100, with planted diverged repeats so that the
    doubling rounds have ties to break."""
    n_bases = int(os.environ.get("GK_TEST_C4_BASES", 100_000_000))
    recs = _device_genome(n_bases, 10, 4, plant=_plant_diverged_repeats)
    sc = SequenceCollection.from_arrays(recs, strands_to_load="both")
    km = Kmers(sc, k, k, source_strand="both")
    km.sort()
    assert len(km) == 2 * (n_bases - 10 * (k - 1))
    assert km.last_sort_stats["levels"] >= 2, "the doubling rounds did not run"
    rep = km.verify_order(k)
    assert rep["ok"] and rep["kmers"] == len(km), rep
    hist, total = km.get_kmer_group_counts(k)
    assert total == len(km) and int(hist.sum()) == rep["groups"]
    assert int(hist[2:].sum()) > 100, "the planted repeats should give groups of equal k-mers"


@pytest.mark.parametrize("k", [15, 31, 63])
def test_config5_one_gbp_full_size_properties(k):
    """BASELINE.json configs[4]: 1 Gbp, both strands, k from the sweep (15: more windows than possible k-mers,
    the grouping stress case; 63: two key words)."""
    n_bases = int(os.environ.get("GK_TEST_C5_BASES", 1_000_000_000))
    recs = _device_genome(n_bases, 10, 5)
    sc = SequenceCollection.from_arrays(recs, strands_to_load="both")
    km = Kmers(sc, k, k, source_strand="both")
    km.sort()
    n = len(km)
    assert n == 2 * (n_bases - 10 * (k - 1))
    rep = km.verify_order(k)
    assert rep["ok"] and rep["kmers"] == n, rep
    hist, total = km.get_kmer_group_counts(k)
    assert total == n and int(hist.sum()) == rep["groups"]
    if k == 15:
        assert rep["groups"] < 4 ** 15 and int(hist[1]) < n // 2      # heavy duplication


def test_repeat_rich_workload_full_size_properties():
    """100 Mbp with ~40 % of the bases in diverged repeat families and low-complexity tracts (bench.py
    --workload repeats), both strands, k = 31: long prefix buckets everywhere."""
    import bench

    n_bases = int(os.environ.get("GK_TEST_REPEAT_BASES", 100_000_000))
    sba, starts, names = bench.make_repeat_genome(n_bases, 10, 20, 42)
    sc = SequenceCollection.from_sba(sba, starts.astype(np.uint32), names, strands_to_load="both", validate=False)
    km = Kmers(sc, 31, 31, source_strand="both")
    km.sort()
    rep = km.verify_order(31)
    assert rep["ok"] and rep["kmers"] == len(km), (rep, km.last_sort_stats)
    hist, total = km.get_kmer_group_counts(31)
    assert total == len(km) and int(hist.sum()) == rep["groups"]
    assert int(hist[2:].sum()) > 100_000
