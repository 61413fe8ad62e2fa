#!/usr/bin/env python
"""
Generate tests/golden/golden_get_kmers.json: outputs of the REAL reference's Kmers.get_kmers generator
(kmers.py:869-992, tuples built by kmers.py:400-425 and :1180-1264) for a list of queries on cases of
golden.npz.  Run in the build container (the reference cannot travel to the GPU box):

    python tests/golden/make_golden_get_kmers.py        # ~3 min (numba recompiles per query)

The k-mer order is the canonical one stored in golden.npz (`sorted`), assigned to the reference object
before the queries, so kmer_num means the same thing on both sides.  Unsorted queries use `init`.

source_strand="both" is not implemented by the reference; as in make_golden.py those cases run the
reference's forward path over forward records + reverse-complemented records (named <name>_rc) in reversed
order.  The tuples are stored exactly as the reference yields them ('+', '<name>_rc', index inside the
reverse-complemented record); tests/test_gpu_parity.py maps them to this repo's ('-', '<name>', forward
sequence index) convention, which is the reference's own convention for its reverse-complement strand
(sequence_collection.py:101-153).
"""
import json
import os
import sys
import types

import numpy as np

sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, "/root/reference/src")

from genome_kmers import kmers as ref_kmers  # noqa: E402
from genome_kmers.kmers import Kmers  # noqa: E402
from genome_kmers.sequence_collection import SequenceCollection  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import both_strand_list  # noqa: E402


def make_filter(spec):
    if spec is None:
        return ref_kmers.kmer_filter_keep_all
    kind = spec[0]
    if kind == "no_ambiguous":
        return ref_kmers.gen_no_ambiguous_bases_filter(spec[1])
    if kind == "gc":
        return ref_kmers.gen_kmer_gc_content_filter_func(spec[1], spec[2], spec[3])
    if kind == "homopolymer":
        return ref_kmers.gen_kmer_homopolymer_filter_func(spec[1], spec[2])
    if kind == "min_length":
        return ref_kmers.gen_kmer_length_filter_func(spec[1])
    if kind == "ngg_pam":
        return ref_kmers.crispr_ngg_pam_filter
    raise ValueError(spec)


def q(kmer_len, info="minimum", one_based=False, filt=None, min_group=1, max_group=None, first_n=None,
      is_sorted=True):
    return dict(kmer_len=kmer_len, info=info, one_based=one_based, filter=filt, min_group=min_group,
                max_group=max_group, first_n=first_n, sorted=is_sorted)


# case name in golden.json -> queries
QUERIES = {
    "sl2_k3": [q(3), q(3, "full"), q(3, "full", one_based=True), q(3, min_group=2, max_group=3, first_n=1),
               q(3, "full", min_group=2), q(3, "full", is_sorted=False), q(3, is_sorted=False),
               q(2, "full", first_n=2), q(3, "full", filt=["gc", 0.3, 0.7, 3])],
    "sl2_var_1_None": [q(None, "full"), q(None), q(4, "full", filt=["min_length", 4]),
                       q(None, "full", one_based=True, first_n=1)],
    "sl2_both_k3": [q(3, "full"), q(3, "full", one_based=True, min_group=2), q(3, first_n=1)],
    "amb_k4": [q(4, "full", filt=["no_ambiguous", 4]), q(4, "full", min_group=2),
               q(4, filt=["homopolymer", 2, 4])],
    "amb_both_k4": [q(4, "full", filt=["no_ambiguous", 4], first_n=1)],
    "rand5k_k21": [q(5, "full", first_n=1, min_group=8), q(21, filt=["gc", 0.4, 0.6, 21], max_group=1),
                   q(6, "full", one_based=True, min_group=3, max_group=4)],
    "rand5k_k31": [q(23, "full", filt=["ngg_pam"], is_sorted=False)],
    "lowcomplex_k8": [q(8, "full", filt=["homopolymer", 3, 8], min_group=2, first_n=2),
                      q(8, min_group=5, max_group=200)],
    "iupac4k_both_k21": [q(6, "full", filt=["no_ambiguous", 6], min_group=4, first_n=2),
                         q(21, "full", one_based=True, first_n=1, max_group=1, filt=["gc", 0.45, 0.55, 21])],
}


def main():
    meta = json.load(open(os.path.join(HERE, "golden.json")))
    arrays = np.load(os.path.join(HERE, "golden.npz"))
    cases = {c["name"]: c for c in meta["cases"]}
    out = []
    for name, queries in QUERIES.items():
        case = cases[name]
        seq_list = [tuple(r) for r in case["seq_list"]]
        if case["strands"] == "both":
            seq_list = both_strand_list(seq_list)
        sc = SequenceCollection(sequence_list=seq_list, strands_to_load="forward")
        for qu in queries:
            km = Kmers(sc, min_kmer_len=case["min_len"], max_kmer_len=case["max_len"])
            assert np.array_equal(km.kmer_sba_start_indices, arrays[f"{name}__init"])
            if qu["sorted"]:
                km.sort()
                km.kmer_sba_start_indices = arrays[f"{name}__sorted"].copy()
            tuples = list(km.get_kmers(
                qu["kmer_len"], one_based_seq_index=qu["one_based"], kmer_filter_func=make_filter(qu["filter"]),
                kmer_info_to_yield=qu["info"], min_group_size=qu["min_group"], max_group_size=qu["max_group"],
                yield_first_n=qu["first_n"]))
            tuples = [[v if isinstance(v, str) else int(v) for v in t] for t in tuples]
            out.append(dict(case=name, query=qu, tuples=tuples))
            print(f"{name} {qu}: {len(tuples)} tuples", flush=True)
    with open(os.path.join(HERE, "golden_get_kmers.json"), "w") as f:
        json.dump(dict(reference="mrperkett/genome-kmers v1.0.1", entries=out), f)
    print("wrote", len(out), "entries")


if __name__ == "__main__":
    main()
