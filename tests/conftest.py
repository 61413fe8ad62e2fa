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


@pytest.fixture(scope="session")
def golden_meta():
    return golden()[0]
