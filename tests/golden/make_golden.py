#!/usr/bin/env python
"""
Generate tests/golden/golden.npz by running the REAL reference (mrperkett/genome-kmers,
/root/reference, numba) in the build container.  The reference cannot travel to the GPU box, so
the vectors are committed; this script is the record of how they were made.

    python tests/golden/make_golden.py            # ~5 min on 6 processes (numba recompiles per call)

Every case records
  * the input records and the Kmers parameters,
  * init:     Kmers.kmer_sba_start_indices right after construction (kmers.py:789-835),
  * sorted:   the start indices after the reference's own quicksort driven by its own
              comparator with break_ties=True (kmers.py:1654-1731) -- the canonical order,
  * default_sort_same_groups: that the reference's default Kmers.sort() output (unstable tie
              order) equals `sorted` after sorting the indices inside every group of equal k-mers,
  * counts:   get_kmer_group_counts / get_kmer_count answers for a list of queries.

source_strand="both" is not implemented by the reference (kmers.py:689-696); those cases run the
reference's forward path over forward records + reverse-complemented records in reversed order,
whose sba is forward || '$' || revcomp.
"""
import json
import os
import sys
import types

import numpy as np

sys.modules.setdefault("h5py", types.ModuleType("h5py"))  # absent in the image; only save/load use it
sys.path.insert(0, "/root/reference/src")

import numba as nb  # noqa: E402
from numba.misc import quicksort  # noqa: E402

from genome_kmers import kmers as ref_kmers  # noqa: E402
from genome_kmers.kmers import Kmers  # noqa: E402
from genome_kmers.sequence_collection import SequenceCollection  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
IUPAC = "RYSWKMBDHVN"
COMP = str.maketrans("ACGTRYSWKMBDHVN", "TGCAYRSWMKVHDBN")


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(np.array(list(alphabet))[rng.integers(0, len(alphabet), n)])


def with_iupac(rng, seq, n_runs, run_lo, run_hi, n_scatter):
    s = list(seq)
    for _ in range(n_runs):
        ln = int(rng.integers(run_lo, run_hi + 1))
        st = int(rng.integers(0, max(1, len(s) - ln)))
        s[st:st + ln] = "N" * min(ln, len(s) - st)
    for _ in range(n_scatter):
        s[int(rng.integers(0, len(s)))] = IUPAC[int(rng.integers(0, len(IUPAC)))]
    return "".join(s)


def repeat_rich(rng, unit_len, n_units, mut_rate):
    unit = rand_seq(rng, unit_len)
    out = []
    for _ in range(n_units):
        u = list(unit)
        for i in range(len(u)):
            if rng.random() < mut_rate:
                u[i] = "ACGT"[int(rng.integers(0, 4))]
        out.append("".join(u))
    return "".join(out)


def both_strand_list(seq_list):
    rc = [(name + "_rc", seq.translate(COMP)[::-1]) for name, seq in reversed(seq_list)]
    return list(seq_list) + rc


def build_cases():
    rng = np.random.default_rng(20261018)
    sl1 = [("chr1", "ATCGAATTAG")]
    sl2 = [("chr1", "ATCGAATTAG"), ("chr2", "GGATCTTGCATT"), ("chr3", "GTGATTGACCCCT")]
    amb = [("chr1", "ATCGAATTAGNNNRYACGTTGCAWSACGT"), ("chr2", "NNNNNNNNACGTNACGTRACGNNNNNN")]
    cases = []

    def add(name, seq_list, k, strands="forward", max_len="same", queries=None):
        cases.append(dict(name=name, seq_list=seq_list, min_len=k,
                          max_len=(k if max_len == "same" else max_len),
                          strands=strands, queries=queries or []))

    def q(kmer_len, filt=None, min_group=1, max_group=None, max_bin=1000000):
        return dict(kmer_len=kmer_len, filter=filt, min_group=min_group, max_group=max_group,
                    max_bin=max_bin)

    for k in range(1, 10):
        add(f"sl1_k{k}", sl1, k, queries=[q(k), q(k, max_bin=2), q(k, min_group=2)])
    for k in range(1, 11):
        add(f"sl2_k{k}", sl2, k,
            queries=[q(k), q(k, max_bin=3), q(k, min_group=2, max_group=3), q(k, max_group=1)])
    add("sl2_both_k3", sl2, 3, strands="both", queries=[q(3), q(3, min_group=2)])
    add("sl2_both_k5", sl2, 5, strands="both", queries=[q(5)])
    for k in (1, 2, 3, 4, 7, 8):
        add(f"amb_k{k}", amb, k, queries=[q(k), q(k, filt=["no_ambiguous", k]), q(k, min_group=2)])
    add("amb_both_k4", amb, 4, strands="both", queries=[q(4), q(4, filt=["no_ambiguous", 4])])

    r3 = [(f"chr{i}", rand_seq(rng, n)) for i, n in enumerate((1700, 1650, 1650))]
    for k in (3, 11, 16, 21, 31, 32):
        add(f"rand5k_k{k}", r3, k, queries=[q(k), q(k, max_bin=4), q(k, min_group=2, max_group=10)])
    add("rand5k_k21_count11", r3, 21, queries=[q(11), q(5, min_group=3), q(21)])

    i3 = [(f"c{i}", with_iupac(rng, rand_seq(rng, n), 3, 20, 120, 25))
          for i, n in enumerate((1500, 1200, 1300))]
    for k in (4, 15, 16, 17, 21, 31, 32):
        add(f"iupac4k_k{k}", i3, k,
            queries=[q(k), q(k, filt=["no_ambiguous", k]), q(k, min_group=2, max_bin=50)])
    for k in (21, 31):
        add(f"iupac4k_both_k{k}", i3, k, strands="both",
            queries=[q(k), q(k, filt=["no_ambiguous", k])])

    # edge: windows whose first ambiguous symbol follows an all-T prefix (upper key bound overflow)
    edge = [("e1", "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTVACGTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTW"),
            ("e2", "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTYAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAB"),
            ("e3", "AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAATTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT")]
    for k in (2, 8, 31, 32):
        add(f"edgeT_k{k}", edge, k, queries=[q(k), q(k, filt=["no_ambiguous", k])])

    low = [("p1", "A" * 700 + "AC" * 300 + rand_seq(rng, 400) + "T" * 300),
           ("p2", "ACG" * 250 + "A" * 500 + rand_seq(rng, 300, "AC"))]
    for k in (2, 8, 31):
        add(f"lowcomplex_k{k}", low, k,
            queries=[q(k), q(k, max_bin=5), q(k, min_group=5, max_group=200, max_bin=100)])
    add("lowcomplex_both_k31", low, 31, strands="both", queries=[q(31)])

    rep = [("r1", repeat_rich(rng, 300, 12, 0.004)), ("r2", repeat_rich(rng, 450, 6, 0.002))]
    repn = [(n, with_iupac(rng, s, 2, 10, 80, 10)) for n, s in rep]
    for k in (33, 40, 62, 64, 100, 150):
        add(f"repeat_k{k}", rep, k, queries=[q(k), q(k, min_group=2)])
    for k in (33, 64, 100):
        add(f"repeatN_k{k}", repn, k, queries=[q(k), q(k, filt=["no_ambiguous", k])])
    add("repeat_both_k64", rep, 64, strands="both", queries=[q(64)])
    add("repeatN_both_k40", repn, 40, strands="both", queries=[q(40)])

    # variable-length / suffix modes of sort()
    for mn, mx in ((1, None), (2, 5), (3, None), (1, 4)):
        add(f"sl2_var_{mn}_{mx}", sl2, mn, max_len=mx, queries=[q(mx), q(2)])
    add("amb_var_1_None", amb, 1, max_len=None, queries=[q(None)])
    add("rand5k_var_5_None", r3, 5, max_len=None, queries=[q(None), q(12)])
    add("repeat_var_20_None", rep, 20, max_len=None, queries=[q(None), q(50)])
    add("iupac4k_var_8_40", i3, 8, max_len=40, queries=[q(40), q(8)])

    big = [(f"chr{i}", rand_seq(rng, 24000)) for i in range(5)]
    add("rand120k_k21", big, 21, queries=[q(21), q(21, max_bin=2)])
    bign = [(n, with_iupac(rng, s, 4, 100, 1500, 30)) for n, s in big]
    add("rand120kN_both_k31", bign, 31, strands="both",
        queries=[q(31), q(31, filt=["no_ambiguous", 31])])
    return cases


def make_filter(spec):
    if spec is None:
        return ref_kmers.kmer_filter_keep_all
    if spec[0] == "no_ambiguous":
        return ref_kmers.gen_no_ambiguous_bases_filter(spec[1])
    raise ValueError(spec)


def canonical_sort(km: Kmers) -> np.ndarray:
    lt = km.get_is_less_than_func(validate_kmers=True, break_ties=True)
    qs = quicksort.make_jit_quicksort(lt=lt, is_argsort=False)
    fn = nb.njit(qs.run_quicksort)
    arr = km.kmer_sba_start_indices.copy()
    fn(arr)
    return arr


def run_case(case):
    seq_list = both_strand_list(case["seq_list"]) if case["strands"] == "both" else case["seq_list"]
    sc = SequenceCollection(sequence_list=seq_list, strands_to_load="forward")
    km = Kmers(sc, min_kmer_len=case["min_len"], max_kmer_len=case["max_len"])
    init = km.kmer_sba_start_indices.copy()
    canon = canonical_sort(km)
    km.sort()
    default = km.kmer_sba_start_indices.copy()

    # canonicalise the default (unstable) order group by group and compare
    cmp_fn = ref_kmers.compare_sba_kmers_lexicographically
    sba = sc.forward_sba
    fixed = default.copy()
    lo = 0
    for p in range(1, len(default) + 1):
        if p == len(default) or cmp_fn(sba, sba, int(default[p - 1]), int(default[p]),
                                       case["max_len"])[0] != 0:
            fixed[lo:p] = np.sort(default[lo:p])
            lo = p
    same = bool(np.array_equal(fixed, canon))
    assert same, case["name"]

    answers = []
    km.kmer_sba_start_indices = canon.copy()
    for qu in case["queries"]:
        hist, total = km.get_kmer_group_counts(
            qu["kmer_len"], kmer_filter_func=make_filter(qu["filter"]),
            min_group_size=qu["min_group"], max_group_size=qu["max_group"],
            max_counts_bin=qu["max_bin"])
        if qu is case["queries"][0]:  # get_kmer_count is the same walk; check it once per case
            cnt = km.get_kmer_count(
                qu["kmer_len"], kmer_filter_func=make_filter(qu["filter"]),
                min_group_size=qu["min_group"], max_group_size=qu["max_group"])
            assert cnt == total
        nz = np.flatnonzero(hist)
        answers.append(dict(total=int(total), hist_bins=nz.tolist(), hist_counts=hist[nz].tolist()))
    return init, canon, same, answers, np.asarray(sc.forward_sba), np.asarray(sc._forward_sba_seg_starts)


def main():
    cases = build_cases()
    arrays, meta = {}, []
    import multiprocessing as mp
    with mp.get_context("fork").Pool(int(os.environ.get("GOLDEN_PROCS", "6"))) as pool:
        results = pool.map(run_case, cases, chunksize=1)
    for i, (case, res) in enumerate(zip(cases, results)):
        init, canon, same, answers, sba, starts = res
        arrays[f"{case['name']}__init"] = init
        arrays[f"{case['name']}__sorted"] = canon
        arrays[f"{case['name']}__sba"] = sba
        arrays[f"{case['name']}__seg_starts"] = starts
        m = dict(case)
        m["answers"] = answers
        m["default_sort_same_groups"] = same
        m["n_kmers"] = int(len(init))
        meta.append(m)
        print(f"[{i + 1}/{len(cases)}] {case['name']}: n={len(init)}", flush=True)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(dict(reference="mrperkett/genome-kmers v1.0.1", numba=nb.__version__,
                       numpy=np.__version__, cases=meta), f)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
