import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genome-kmers_b200"))

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _load_golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        meta = json.load(f)
    arrays = np.load(os.path.join(GOLDEN_DIR, "golden.npz"))
    return meta, arrays


_GOLDEN = None


def golden():
    global _GOLDEN
    if _GOLDEN is None:
        _GOLDEN = _load_golden()
    return _GOLDEN


def golden_case_names(pred=None):
    meta, _ = golden()
    return [c["name"] for c in meta["cases"] if pred is None or pred(c)]


def golden_case(name):
    meta, arrays = golden()
    case = next(c for c in meta["cases"] if c["name"] == name)
    out = dict(case)
    for key in ("init", "sorted", "sba", "seg_starts"):
        out[key] = arrays[f"{name}__{key}"]
    return out


def is_fixed_k(case):
    return case["max_len"] is not None and case["max_len"] == case["min_len"]


def dense_hist(answer, max_bin):
    hist = np.zeros(max_bin + 1, dtype=np.int64)
    hist[np.asarray(answer["hist_bins"], dtype=np.int64)] = answer["hist_counts"]
    return hist


def get_kmers_entries():
    """Outputs of the real reference's Kmers.get_kmers (tests/golden/make_golden_get_kmers.py)."""
    with open(os.path.join(GOLDEN_DIR, "golden_get_kmers.json")) as f:
        return json.load(f)["entries"]


def get_kmers_entry_id(entry):
    return "%s-%s" % (entry["case"], "-".join(str(v) for v in entry["query"].values()))


def filter_from_spec(spec):
    from genome_kmers import kmers as gk

    if spec is None:
        return gk.kmer_filter_keep_all
    kind, args = spec[0], spec[1:]
    return {"no_ambiguous": gk.gen_no_ambiguous_bases_filter, "gc": gk.gen_kmer_gc_content_filter_func,
            "homopolymer": gk.gen_kmer_homopolymer_filter_func, "min_length": gk.gen_kmer_length_filter_func,
            "ngg_pam": lambda: gk.crispr_ngg_pam_filter}[kind](*args)


def expected_get_kmers_tuples(entry, case):
    """The reference's tuples in this repo's convention: a k-mer of a '<name>_rc' record (the reference ran
    its forward path over forward + reverse-complemented records) is ('-', name, forward sequence index),
    the reference's own convention for its reverse-complement strand (sequence_collection.py:101-153)."""
    qu = entry["query"]
    rec_len = {name: len(seq) for name, seq in case["seq_list"]}
    ob = 1 if qu["one_based"] else 0
    want = []
    for t in entry["tuples"]:
        if qu["info"] == "full" and t[2].endswith("_rc"):
            name = t[2][:-3]
            t = [t[0], "-", name, rec_len[name] - 1 - (t[3] - ob) + ob] + t[4:]
        want.append(tuple(t))
    return want


@pytest.fixture(scope="session")
def golden_meta():
    return golden()[0]
