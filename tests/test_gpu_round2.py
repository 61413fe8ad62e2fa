"""
GPU tests of the round-2 paths: ambiguous-window fragments (pack-time run-length blocks) against the
element-wise refinement, the single-CTA small sort, sorting an assigned subset of start indices
(reference kmers.py:1648), the device-side order check (gk_index_verify), reverse complement on unaligned
buffers.  Bit-exact against the CPU oracle / the golden vectors of the real reference.
"""
import ctypes

import numpy as np
import pytest

import oracle
from conftest import dense_hist, golden_case, golden_case_names, is_fixed_k
from genome_kmers import _native
from genome_kmers.kmers import Kmers, gen_no_ambiguous_bases_filter, kmer_filter_keep_all
from genome_kmers.sequence_collection import SequenceCollection
import gpu_utils as gu

pytestmark = pytest.mark.gpu

AMBIGUOUS_GOLDEN = [n for n in golden_case_names(is_fixed_k)
                    if any(t in n for t in ("iupac", "repeatN", "rand120kN", "amb", "N_both", "every"))]


def _oracle_sorted(recs, k, strands, min_len=None, max_len=None):
    sc = SequenceCollection.from_arrays(recs, strands_to_load=strands)
    fwd, starts = sc.forward_sba, sc._forward_sba_seg_starts.astype(np.uint64)
    if strands == "both":
        sba, seg = oracle.both_strands(fwd, starts)
    else:
        sba, seg = fwd, starts
    lo, hi = (k, k) if min_len is None else (min_len, max_len)
    init = oracle.init_indices(seg, len(sba), lo)
    want = oracle.sort_indices(sba, init, lo, hi, threads=min(8, oracle.max_threads()))
    return sc, sba, seg, want


@pytest.mark.parametrize("frag", ["0", "1"])
@pytest.mark.parametrize("name", AMBIGUOUS_GOLDEN)
def test_golden_ambiguous_cases_with_and_without_fragments(name, frag, monkeypatch):
    monkeypatch.setenv("GK_FRAGMENTS", frag)
    case = golden_case(name)
    sc = SequenceCollection(sequence_list=[tuple(r) for r in case["seq_list"]], strands_to_load=case["strands"])
    km = Kmers(sc, case["min_len"], case["max_len"], source_strand=case["strands"])
    km.sort()
    got = km.kmer_sba_start_indices
    bad = np.flatnonzero(got != case["sorted"])
    assert len(bad) == 0, f"{name}: first mismatches at {bad[:8]}"
    if frag == "1" and km.last_sort_stats["n_ambiguous"]:
        assert km.last_sort_stats["n_fragments"] > 0
    for qu, ans in zip(case["queries"], case["answers"]):
        flt = kmer_filter_keep_all if qu["filter"] is None else gen_no_ambiguous_bases_filter(qu["filter"][1])
        hist, total = km.get_kmer_group_counts(qu["kmer_len"], flt, qu["min_group"], qu["max_group"], qu["max_bin"])
        assert total == ans["total"] and np.array_equal(hist, dense_hist(ans, qu["max_bin"])), (name, qu)
    assert km.verify_order(case["max_len"])["ok"]


@pytest.mark.parametrize("k,strands", [(31, "both"), (12, "forward"), (5, "both"), (17, "both")])
def test_fragments_on_long_runs_of_every_kind(k, strands):
    """N runs longer than several pack tiles, runs of other IUPAC letters, runs that touch record ends, runs
    shorter than k, and scattered single letters: the fragment path against the oracle."""
    rng = np.random.default_rng(100 + k)
    recs = gu.random_genome(rng, 260_000, 3, n_runs=0)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for name, seq in recs:
        n = len(seq)
        seq[0:9000] = ord("N")                       # run at the start of a record, > 2 pack tiles
        seq[n - 5000:n] = ord("N")                   # run at the end of a record
        seq[20_000:20_000 + k - 1] = ord("N")        # shorter than k: no all-N window
        seq[30_000:30_000 + k] = ord("N")            # exactly one all-N window
        seq[40_000:40_000 + k + 1] = ord("R")        # two identical windows of another letter
        seq[50_000:50_700] = ord("Y")
        seq[60_000:60_300] = ord("N")
        seq[60_300:60_600] = ord("R")                # two runs back to back
        seq[rng.integers(0, n, 30)] = ord("W")
        seq[70_000:70_000 + 3 * k] = acgt[0]         # a pure homopolymer: never a fragment
    sc, sba, seg, want = _oracle_sorted(recs, k, strands)
    km = Kmers(sc, k, k, source_strand=strands)
    km.sort()
    got = km.kmer_sba_start_indices.astype(np.uint64)
    bad = np.flatnonzero(got != want)
    assert len(bad) == 0, f"first mismatches at {bad[:8]}: got {got[bad[:8]]} want {want[bad[:8]]}"
    st = km.last_sort_stats
    assert st["n_fragments"] > 0 and st["n_fragments"] < st["n_ambiguous"]
    assert st["refine_flags"] & 1 and not st["refine_flags"] & 2, st   # fragment path, no element-wise repair
    hist, total = km.get_kmer_group_counts(k, max_counts_bin=50)
    o_hist, o_total = oracle.group_hist(sba, want, k, max_bin=50)
    assert total == o_total and np.array_equal(hist, o_hist)
    rep = km.verify_order(k)
    assert rep["ok"] and rep["groups"] == int(hist.sum()) and rep["flags_compared"] == 1


@pytest.mark.parametrize("n", [2, 31, 33, 1000, 1025, 40_000, 65_536])
@pytest.mark.parametrize("val_dtype", [np.uint32, np.uint64])
def test_small_sort_matches_stable_numpy(n, val_dtype, monkeypatch):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << 63, n, dtype=np.uint64)
    keys[rng.integers(0, n, n // 3)] &= np.uint64(0xFF00FF)          # many ties and constant digits
    vals = np.arange(n).astype(val_dtype)
    for begin, end in ((0, 64), (8, 40), (3, 21)):
        order = np.argsort((keys >> np.uint64(begin)) & np.uint64((1 << (end - begin)) - 1), kind="stable")
        for small in ("1", "0"):
            monkeypatch.setenv("GK_SMALL_SORT", small)
            k_out, v_out = gu.radix_sort_pairs(keys, vals, begin, end)
            assert np.array_equal(v_out, vals[order]), (n, begin, end, small)
            assert np.array_equal(k_out, keys[order])


def test_sort_of_an_assigned_subset_of_start_indices():
    """reference kmers.py:1648: sort() orders whatever kmer_sba_start_indices holds."""
    rng = np.random.default_rng(5)
    recs = gu.random_genome(rng, 50_000, 3, n_runs=4, run_lo=40, run_hi=400, n_scatter=10)
    k = 21
    sc, sba, seg, want_all = _oracle_sorted(recs, k, "forward")
    init = oracle.init_indices(seg, len(sba), k)
    subset = rng.permutation(init)[:7000].astype(np.uint32)          # arbitrary order, no duplicates
    km = Kmers(sc, k, k)
    km.kmer_sba_start_indices = subset
    km.sort()
    want = oracle.sort_indices(sba, np.sort(subset).astype(np.uint64), k, k, threads=1)
    got = km.kmer_sba_start_indices
    assert len(got) == len(subset) and np.array_equal(got.astype(np.uint64), want)
    hist, total = km.get_kmer_group_counts(k, max_counts_bin=20)
    o_hist, o_total = oracle.group_hist(sba, want, k, max_bin=20)
    assert total == o_total == len(subset) and np.array_equal(hist, o_hist)
    # sorting again is idempotent; assigning the whole set (shuffled) gives the full order
    km.sort()
    assert np.array_equal(km.kmer_sba_start_indices.astype(np.uint64), want)
    km.kmer_sba_start_indices = rng.permutation(init).astype(np.uint32)
    km.sort()
    assert np.array_equal(km.kmer_sba_start_indices.astype(np.uint64), want_all)
    # a start that begins no k-mer of min_kmer_len symbols: the reference's validation raises
    bad = subset.copy()
    bad[3] = np.uint32(int(seg[1]) - 2)                               # last base of record 0
    km.kmer_sba_start_indices = bad
    with pytest.raises(Exception, match="less than min_kmer_len"):
        km.sort()


def test_verify_order_reports_what_is_wrong():
    rng = np.random.default_rng(11)
    recs = gu.random_genome(rng, 30_000, 2, n_runs=2, run_lo=50, run_hi=200)
    k = 15
    sc = SequenceCollection.from_arrays(recs)
    km = Kmers(sc, k, k)
    assert not km.verify_order(k)["ok"]                                # init order is not sorted
    km.sort()
    rep = km.verify_order(k)
    assert rep["ok"] and rep["kmers"] == len(km) and rep["flags_compared"] == 1
    good = km.kmer_sba_start_indices.copy()
    swapped = good.copy()
    swapped[[100, 20_000]] = swapped[[20_000, 100]]
    km2 = Kmers(sc, k, k)
    km2.kmer_sba_start_indices = swapped
    km2._is_sorted = True
    rep = km2.verify_order(k)
    assert not rep["ok"] and rep["out_of_order"] >= 1
    dup = good.copy()
    dup[5] = dup[6]
    km2.kmer_sba_start_indices = dup
    rep = km2.verify_order(k)
    assert rep["duplicate_starts"] == 1


@pytest.mark.parametrize("n,shift", [(100003, 1), (4099, 7), (65, 3), (1 << 20, 5)])
def test_revcomp_on_unaligned_input(n, shift):
    torch = gu.torch_mod()
    rng = np.random.default_rng(n)
    sba = np.frombuffer(b"ACGTRYSWKMBDHVN$", dtype=np.uint8)[rng.integers(0, 16, n + shift)].copy()
    d_in = gu.dev(sba)[shift:]
    d_out = torch.zeros(n + 11, dtype=torch.uint8, device="cuda")
    _native.check(_native.lib().gk_sba_revcomp(d_in.data_ptr(), n, d_out[11:].data_ptr(), gu.stream()))
    assert np.array_equal(d_out[11:].cpu().numpy(), oracle.revcomp(sba[shift:]))


def test_pure_kmer_inside_a_run_of_equal_ambiguous_keys_is_repaired_bucket_wise():
    """Every window 'A' + 30 x 'N' has the radix key of the pure 31-mer 'AT' + 29 x 'A' minus the class bit,
    so they share a prefix bucket; a pure k-mer that starts before them leaves that long bucket out of order.
    Only that bucket is re-sorted (refine_flags bit 16), and the fragments fill the ambiguous slots."""
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seq = acgt[rng.integers(0, 4, 150_000)].copy()
    seq[1000:1031] = np.frombuffer(b"AT" + b"A" * 29, dtype=np.uint8)
    for i in range(12):                                   # twelve 'A' + N-run edges: a long bucket of equal keys
        p = 20_000 + 9_000 * i
        seq[p] = ord("A")
        seq[p + 1:p + 200] = ord("N")
    sc, sba, seg, want = _oracle_sorted([("chr0", seq)], 31, "forward")
    km = Kmers(sc, 31, 31)
    km.sort()
    got = km.kmer_sba_start_indices.astype(np.uint64)
    assert np.array_equal(got, want)
    st = km.last_sort_stats
    assert st["refine_flags"] & 4 and st["refine_flags"] & 16 and st["refine_flags"] & 1, st
    assert not st["refine_flags"] & 2, st
    hist, total = km.get_kmer_group_counts(31, max_counts_bin=300)
    o_hist, o_total = oracle.group_hist(sba, want, 31, max_bin=300)
    assert total == o_total and np.array_equal(hist, o_hist)
    assert km.verify_order(31)["ok"]


def test_repeat_rich_input_is_repaired_without_the_elementwise_path():
    """Diverged copies of one unit: many prefix buckets longer than eight, some out of order."""
    rng = np.random.default_rng(78)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    unit = acgt[rng.integers(0, 4, 900)]
    parts = []
    for _ in range(150):
        c = unit.copy()
        c[rng.integers(0, 900, 4)] = acgt[rng.integers(0, 4, 4)]
        parts += [c, acgt[rng.integers(0, 4, int(rng.integers(10, 200)))]]
    seq = np.concatenate(parts)
    sc, sba, seg, want = _oracle_sorted([("chr0", seq[:70_000]), ("chr1", seq[70_000:])], 31, "both")
    km = Kmers(sc, 31, 31, source_strand="both")
    km.sort()
    assert np.array_equal(km.kmer_sba_start_indices.astype(np.uint64), want)
    assert km.verify_order(31)["ok"]
    hist, total = km.get_kmer_group_counts(31, max_counts_bin=400)
    o_hist, o_total = oracle.group_hist(sba, want, 31, max_bin=400)
    assert total == o_total and np.array_equal(hist, o_hist)


def test_many_out_of_order_long_runs_are_resorted_by_key_and_fragments_still_apply():
    """Hundreds of diverged copies of a long unit: far more out-of-order members of long prefix runs than the
    device-side bucket list holds.  All long runs are then re-sorted by key (refine_flags bit 64) and the
    ambiguous windows still come from the fragments -- no element-wise repair (bit 2)."""
    rng = np.random.default_rng(79)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    unit = acgt[rng.integers(0, 4, 2500)]
    parts = []
    for c_i in range(300):
        c = unit.copy()
        c[rng.integers(0, len(unit), 60)] = acgt[rng.integers(0, 4, 60)]
        parts += [c, acgt[rng.integers(0, 4, int(rng.integers(10, 200)))]]
        if c_i % 40 == 0:
            parts.append(np.full(int(rng.integers(50, 3000)), ord("N"), dtype=np.uint8))
    seq = np.concatenate(parts)
    seq[rng.integers(0, len(seq), 12)] = ord("R")
    cut = len(seq) // 2
    sc, sba, seg, want = _oracle_sorted([("chr0", seq[:cut]), ("chr1", seq[cut:])], 31, "both")
    km = Kmers(sc, 31, 31, source_strand="both")
    km.sort()
    st = km.last_sort_stats
    assert np.array_equal(km.kmer_sba_start_indices.astype(np.uint64), want)
    assert st["refine_flags"] & 64 and st["refine_flags"] & 1 and not st["refine_flags"] & 2, st
    assert km.verify_order(31)["ok"]
    hist, total = km.get_kmer_group_counts(31, max_counts_bin=700)
    o_hist, o_total = oracle.group_hist(sba, want, 31, max_bin=700)
    assert total == o_total and np.array_equal(hist, o_hist)


@pytest.mark.parametrize("strands", ["forward", "both"])
def test_stranger_in_the_all_n_bucket_is_moved_without_sorting_the_bucket(strands):
    """A bucket of more than 65536 equal ambiguous keys (one long N run) with pure k-mers that share its 32-bit
    prefix ('T' + A's: the all-N key is the count of pure k-mers below 'N...', i.e. 'TAAA...A'), some starting
    before the run and some after it: the device-side repair moves only the strangers (no host round trip:
    refine_flags bit 32 stays clear) and the fragments rewrite the ambiguous slots."""
    rng = np.random.default_rng(79)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seq = acgt[rng.integers(0, 4, 400_000)].copy()
    stranger = np.frombuffer(b"T" + b"A" * 17 + b"CCGTA" + b"ACGTACGTA", dtype=np.uint8)   # 32 symbols
    seq[5_000:5_032] = stranger
    seq[100_000:180_000] = ord("N")
    seq[300_000:300_032] = stranger
    seq[350_000:350_031] = np.frombuffer(b"T" + b"A" * 30, dtype=np.uint8)
    sc, sba, seg, want = _oracle_sorted([("chr0", seq)], 31, strands)
    km = Kmers(sc, 31, 31, source_strand=strands)
    km.sort()
    got = km.kmer_sba_start_indices.astype(np.uint64)
    bad = np.flatnonzero(got != want)
    assert len(bad) == 0, f"first mismatches at {bad[:8]}: got {got[bad[:8]]} want {want[bad[:8]]}"
    st = km.last_sort_stats
    assert st["refine_flags"] & 4 and st["refine_flags"] & 16 and st["refine_flags"] & 1, st
    assert not st["refine_flags"] & (2 | 32), st
    hist, total = km.get_kmer_group_counts(31, max_counts_bin=100)
    o_hist, o_total = oracle.group_hist(sba, want, 31, max_bin=100)
    assert total == o_total and np.array_equal(hist, o_hist)
    assert km.verify_order(31)["ok"]


@pytest.mark.parametrize("spectrum", ["0", "1"])
@pytest.mark.parametrize("name", ["sl2_k3", "rand5k_k21_count11", "lowcomplex_k8", "rand120kN_both_k31", "iupac4k_k15"])
def test_group_counts_from_the_cached_spectrum_and_from_the_device_agree(name, spectrum, monkeypatch):
    """sort() leaves the group-size spectrum on the host (GK_SPECTRUM=1, default) and later queries are answered
    from it; GK_SPECTRUM=0 keeps every query on the device histogram kernels.  Same answers as the reference."""
    monkeypatch.setenv("GK_SPECTRUM", spectrum)
    case = golden_case(name)
    sc = SequenceCollection(sequence_list=[tuple(r) for r in case["seq_list"]], strands_to_load=case["strands"])
    km = Kmers(sc, case["min_len"], case["max_len"], source_strand=case["strands"])
    km.sort()
    for qu, ans in zip(case["queries"], case["answers"]):
        flt = kmer_filter_keep_all if qu["filter"] is None else gen_no_ambiguous_bases_filter(qu["filter"][1])
        hist, total = km.get_kmer_group_counts(qu["kmer_len"], flt, qu["min_group"], qu["max_group"], qu["max_bin"])
        assert total == ans["total"] and np.array_equal(hist, dense_hist(ans, qu["max_bin"])), (name, qu)
        assert km.get_kmer_count(qu["kmer_len"], flt, qu["min_group"], qu["max_group"]) == ans["total"]
    for max_bin in (1, 2, 5, 1000):
        h, t = km.get_kmer_group_counts(case["max_len"], max_counts_bin=max_bin)
        assert t == len(km) and int((h * np.arange(max_bin + 1)).sum()) <= len(km)


@pytest.mark.parametrize("n", [5, 8192, 8195, 40_001])
@pytest.mark.parametrize("skip", [0, 1])
def test_partition_to_destination_buffers_with_key_base_and_dropped_ambiguous_pairs(n, skip):
    """gk_partition_pairs_peer into local buffers: stable per destination, keys minus the destination's base,
    class-0 pairs dropped when asked, and not one slot written beyond each destination's count -- also when the
    last (partial) tile holds dropped pairs (its padding must not leak into a destination)."""
    torch = gu.torch_mod()
    lib = _native.lib()
    rng = np.random.default_rng(n + skip)
    keys = rng.integers(1 << 40, 1 << 62, n, dtype=np.uint64) | np.uint64(1)
    amb = rng.random(n) < 0.2
    amb[-3:] = True                                   # ambiguous pairs at the very end of the last tile
    keys[amb] &= ~np.uint64(1)
    vals = np.arange(n, dtype=np.uint32)
    splitters = np.sort(rng.integers(1 << 40, 1 << 62, 3, dtype=np.uint64)) & ~np.uint64(1)
    base = np.concatenate([np.zeros(1, dtype=np.uint64), splitters])
    dest = np.searchsorted(splitters, keys, side="right")
    keep = ~amb if skip else np.ones(n, dtype=bool)
    counts = np.bincount(dest[keep], minlength=4)
    SENT = np.uint64(0xDEADBEEFDEADBEEF)
    d_keys, d_vals, d_split = gu.dev(keys), gu.dev(vals), gu.dev(splitters)
    outs_k = [torch.full((int(c) + 8,), int(SENT.view(np.int64)), dtype=torch.int64, device="cuda") for c in counts]
    outs_v = [torch.full((int(c) + 8,), -1, dtype=torch.int32, device="cuda") for c in counts]
    kp = np.array([t.data_ptr() for t in outs_k], dtype=np.uint64)
    vp = np.array([t.data_ptr() for t in outs_v], dtype=np.uint64)
    off = np.zeros(4, dtype=np.uint64)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    _native.check(lib.gk_partition_pairs_peer(d_keys.data_ptr(), d_vals.data_ptr(), 4, n, d_split.data_ptr(), 4,
                                              _native.host_ptr(kp), _native.host_ptr(vp), _native.host_ptr(off),
                                              _native.host_ptr(base), skip, err.data_ptr(), gu.stream()))
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    for d in range(4):
        sel = keep & (dest == d)
        got_k = gu.host(outs_k[d], np.uint64)
        got_v = gu.host(outs_v[d], np.uint32)
        c = int(counts[d])
        assert np.array_equal(got_k[:c], keys[sel] - base[d]) and np.array_equal(got_v[:c], vals[sel]), (n, skip, d)
        assert bool((got_k[c:] == SENT).all()) and bool((got_v[c:] == np.uint32(0xFFFFFFFF)).all()), \
            f"destination {d}: slots beyond its {c} pairs were written"
    # the split counts agree with what the partition does
    d_counts = torch.zeros(8, dtype=torch.int64, device="cuda")
    _native.check(lib.gk_partition_count_split(d_keys.data_ptr(), n, d_split.data_ptr(), 4, 1, d_counts.data_ptr(),
                                               gu.stream()))
    got_counts = d_counts.cpu().numpy()
    assert np.array_equal(got_counts[:4], np.bincount(dest[~amb], minlength=4))
    assert np.array_equal(got_counts[4:], np.bincount(dest[amb], minlength=4))


@pytest.mark.parametrize("k,strands,wide", [(33, "forward", False), (40, "both", False), (64, "both", False),
                                            (100, "forward", False), (40, "both", True)])
def test_shard_sort_of_kmers_longer_than_one_key_word(k, strands, wide, tmp_path, monkeypatch):
    """gk_index_sort_shard with k > 31: the pairs carry the first 31 symbols, the rest is compared from the
    bytes in word rounds (the doubling of the single-GPU path needs ranks that live on other GPUs).  One rank,
    so the whole sharded path -- slice pack, staging partition, fragments, shard sort -- runs in this process."""
    import torch.distributed as dist

    from genome_kmers.distributed import ShardedKmers

    if wide:
        monkeypatch.setenv("GK_FORCE_IDX64", "1")     # 64-bit start indices on a small input
    rng = np.random.default_rng(k)
    recs = gu.random_genome(rng, 300_000, 3, n_runs=5, run_lo=40, run_hi=3000, n_scatter=15)
    for _ in range(80):   # copies: k-mers that agree on the first 31 symbols and differ later, or never
        seq = recs[int(rng.integers(0, len(recs)))][1]
        ln = int(rng.integers(33, 500))
        src, dst = (int(v) for v in rng.integers(0, len(seq) - ln, 2))
        seq[dst:dst + ln] = seq[src:src + ln].copy()
    sba = np.concatenate([np.concatenate([seq, np.array([36], dtype=np.uint8)]) for _, seq in recs])[:-1]
    starts = np.cumsum([0] + [len(seq) + 1 for _, seq in recs[:-1]]).astype(np.uint64)
    full, full_starts = oracle.both_strands(sba, starts) if strands == "both" else (sba, starts)
    want = oracle.sort_indices(full, oracle.init_indices(full_starts, len(full), k), k, k,
                               threads=min(8, oracle.max_threads()))
    hist_want, total_want = oracle.group_hist(full, want, k, max_bin=1000)
    store = dist.FileStore(str(tmp_path / "store"), 1)
    dist.init_process_group("gloo", store=store, rank=0, world_size=1)
    try:
        sk = ShardedKmers(sba, starts, k, strands)
        sk.sort()
        hist, total = sk.get_kmer_group_counts(k, max_counts_bin=1000)
        got = sk.local_start_indices()
        ver = sk.verify(hist, len(want))
        levels = sk.stats["levels"]
        sk.close()
    finally:
        dist.destroy_process_group()
    assert np.array_equal(got.astype(np.uint64), want)
    assert total == total_want and np.array_equal(hist, hist_want)
    assert all(ver["checks"].values()), ver
    assert levels > 1, "no word round ran: the input has no k-mers tied on 31 symbols?"


@pytest.mark.parametrize("n", [1, 37, 4099, 300_001])
def test_both_strand_layout_built_in_ranges_equals_the_whole(n):
    """gk_sba_both_strands_range: any partition of [0, 2n + 1) into ranges gives the layout of
    gk_sba_both_strands (the multi-GPU path builds a rank's own slice first, the rest on another stream)."""
    torch = gu.torch_mod()
    lib = _native.lib()
    rng = np.random.default_rng(n)
    sba = np.frombuffer(b"ACGTRYSWKMBDHVN$", dtype=np.uint8)[rng.integers(0, 16, n)].copy()
    d_in = gu.dev(sba)
    whole = torch.zeros(2 * n + 1, dtype=torch.uint8, device="cuda")
    _native.check(lib.gk_sba_both_strands(d_in.data_ptr(), n, whole.data_ptr(), gu.stream()))
    total = 2 * n + 1
    for trial in range(6):
        cuts = sorted({0, total, *(int(c) for c in rng.integers(0, total + 1, 4)), n, min(total, n + 1)})
        out = torch.full((total,), 255, dtype=torch.uint8, device="cuda")
        for b, e in zip(cuts[:-1], cuts[1:]):
            _native.check(lib.gk_sba_both_strands_range(d_in.data_ptr(), n, out.data_ptr(), b, e, gu.stream()))
        assert torch.equal(out, whole), f"cuts {cuts}"
    # a range writes nothing outside itself
    out = torch.full((total,), 255, dtype=torch.uint8, device="cuda")
    b, e = total // 3, (2 * total) // 3 + 1
    _native.check(lib.gk_sba_both_strands_range(d_in.data_ptr(), n, out.data_ptr(), b, e, gu.stream()))
    got = out.cpu().numpy()
    assert np.array_equal(got[b:e], whole.cpu().numpy()[b:e])
    assert (got[:b] == 255).all() and (got[e:] == 255).all()


@pytest.mark.parametrize("k", [12, 31, 64])
def test_pack_slice_reads_nothing_beyond_two_tiles_around_its_slice(k):
    """The multi-GPU path builds bytes [first - 8192, end + 8192 + k) of the both-strand layout before it packs
    the slice [first, end) and the rest later: whatever the other bytes hold then ('$' is the worst garbage: it
    shifts the record count of a tile) must not change the pairs or the fragments of the slice."""
    torch = gu.torch_mod()
    lib = _native.lib()
    rng = np.random.default_rng(k)
    recs = gu.random_genome(rng, 400_000, 5, n_runs=6, run_lo=40, run_hi=6000, n_scatter=20)
    sba = np.concatenate([np.concatenate([seq, np.array([36], dtype=np.uint8)]) for _, seq in recs])[:-1]
    starts = np.cumsum([0] + [len(seq) + 1 for _, seq in recs[:-1]]).astype(np.uint64)
    n = len(sba)
    cap_frag = 1 << 14

    def pack(d_bytes, first, end):
        keys = torch.zeros(end - first, dtype=torch.int64, device="cuda")
        idx = torch.zeros(end - first, dtype=torch.int32, device="cuda")
        frag = torch.zeros(cap_frag * 36, dtype=torch.uint8, device="cuda")
        counters = torch.zeros(4, dtype=torch.int64, device="cuda")
        n_out = ctypes.c_uint64(0)
        _native.check(lib.gk_pack_slice(d_bytes.data_ptr(), n, _native.host_ptr(starts), len(starts), k, 1, first, end,
                                        keys.data_ptr(), 4, idx.data_ptr(), end - first, ctypes.byref(n_out),
                                        frag.data_ptr(), cap_frag, counters.data_ptr(), gu.stream()))
        torch.cuda.synchronize()
        m = n_out.value
        c = counters.cpu().numpy()
        f = frag.cpu().numpy()
        nf = int(c[2])
        words = f[:cap_frag * 32].view(np.uint64).reshape(4, cap_frag)[:, :nf]
        cnt = f[cap_frag * 32:].view(np.uint32)[:nf]
        order = np.argsort(words[3], kind="stable")          # fragments are listed in any order: by start
        return keys[:m].cpu().numpy(), idx[:m].cpu().numpy(), int(c[0]), words[:, order], cnt[order]

    full = gu.dev(sba)
    for first, end in ((0, 100_000), (100_001, 233_333), (n - 90_000, n), (4096 * 30, 4096 * 40)):
        poisoned = np.full(n, 36, dtype=np.uint8)
        lo, hi = max(0, first - 8192), min(n, end + 8192 + k)
        poisoned[lo:hi] = sba[lo:hi]
        want = pack(full, first, end)
        got = pack(gu.dev(poisoned), first, end)
        for a, b in zip(got, want):
            assert np.array_equal(a, b), (first, end)
        assert len(want[0]) > 0 and want[2] > 0


@pytest.mark.parametrize("presort", ["1", "0"])
def test_sharded_path_with_merged_and_with_resorted_fragment_lists(presort, tmp_path, monkeypatch):
    """GK_FRAG_PRESORT=1 (default): every rank sorts its own fragment list and the owner of a key range merges the
    lists; 0: the owner sorts the gathered lists itself.  Same order either way (one rank here; two ranks on one
    GPU in test_gpu_multi.py)."""
    import torch.distributed as dist

    from genome_kmers.distributed import ShardedKmers

    monkeypatch.setenv("GK_FRAG_PRESORT", presort)
    rng = np.random.default_rng(31)
    recs = gu.random_genome(rng, 400_000, 4, n_runs=12, run_lo=31, run_hi=9000, n_scatter=40)
    sba = np.concatenate([np.concatenate([seq, np.array([36], dtype=np.uint8)]) for _, seq in recs])[:-1]
    starts = np.cumsum([0] + [len(seq) + 1 for _, seq in recs[:-1]]).astype(np.uint64)
    full, full_starts = oracle.both_strands(sba, starts)
    want = oracle.sort_indices(full, oracle.init_indices(full_starts, len(full), 31), 31, 31,
                               threads=min(8, oracle.max_threads()))
    store = dist.FileStore(str(tmp_path / "store"), 1)
    dist.init_process_group("gloo", store=store, rank=0, world_size=1)
    try:
        sk = ShardedKmers(sba, starts, 31, "both")
        sk.sort()
        got = sk.local_start_indices()
        stats = dict(sk.stats)
        hist, total = sk.get_kmer_group_counts(31)
        ver = sk.verify(hist, len(want))
        sk.close()
    finally:
        dist.destroy_process_group()
    assert np.array_equal(got.astype(np.uint64), want)
    assert stats["n_fragments"] > 0 and stats["refine_flags"] & 1 and not stats["refine_flags"] & 2, stats
    assert all(ver["checks"].values()), ver
