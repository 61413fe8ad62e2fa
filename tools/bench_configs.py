#!/usr/bin/env python
"""
Side measurements for the other BASELINE.json configs (not the bench line): sort + count through the
public API on one GPU, device-resident timing (CUDA events around sort() + get_kmer_group_counts()).

    python tools/bench_configs.py [c1] [c4] [c4k64] [c5] [c5k63] [c5big]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genome-kmers_b200"))

import torch  # noqa: E402

from bench import make_genome  # noqa: E402
from genome_kmers.kmers import Kmers  # noqa: E402
from genome_kmers.sequence_collection import SequenceCollection  # noqa: E402


def run(name, n_bases, n_rec, runs, strands, k, reps=3):
    sba, starts, names = make_genome(n_bases, n_rec, runs, 42)
    sc = SequenceCollection.from_sba(sba, starts.astype(np.uint32), names, strands_to_load=strands, validate=False)
    best, stats, n, distinct = None, None, 0, 0
    for _ in range(reps):
        km = Kmers(sc, k, k, source_strand=strands)
        km._ensure_device()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        km.sort()
        t1 = time.perf_counter()
        hist, total = km.get_kmer_group_counts(k)
        t2 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        wall = (round(1e3 * (t1 - t0), 2), round(1e3 * (t2 - t1), 2))
        n, distinct = total, int(hist.sum())
        if best is None or ms < best:
            best, stats = ms, dict(km.last_sort_stats, wall_sort_count_ms=wall)
        del km
    print(json.dumps({"config": name, "bases": n_bases, "records": n_rec, "strands": strands, "k": k, "kmers": n,
                      "distinct": distinct, "ms": round(best, 3), "Gkmer_per_s": round(n / best / 1e6, 2),
                      "sort_passes": stats["sort_passes"], "key_bits": stats["key_bits"], "levels": stats["levels"],
                      "launches": stats["gpu_launches"],
                      "stage_ms": {k_: round(stats[k_], 2) for k_ in ("pack_ms", "sort_ms", "fixup_ms", "total_ms")},
                      "wall_sort_count_ms": stats["wall_sort_count_ms"]}), flush=True)


def main():
    which = sys.argv[1:] or ["c1", "c4", "c5"]
    if "c1" in which:
        run("C1 4.6 Mbp forward k=21", 4_600_000, 1, 0, "forward", 21)
    if "c4" in which:
        for k in (64, 100):
            run(f"C4 100 Mbp both k={k}", 100_000_000, 10, 0, "both", k)
    if "c4k64" in which:            # one configuration, two repetitions: launch lists under ncu
        run("C4 100 Mbp both k=64", 100_000_000, 10, 0, "both", 64, reps=2)
    if "c5" in which:
        for k in (15, 21, 27, 31, 32, 33, 47, 63):
            run(f"k-sweep 100 Mbp both k={k}", 100_000_000, 10, 0, "both", k)
    if "c5k63" in which:
        run("C5 1 Gbp both k=63", 1_000_000_000, 10, 0, "both", 63, reps=2)
    if "c5big" in which:
        for k in (15, 31, 63):
            run(f"C5 1 Gbp both k={k}", 1_000_000_000, 10, 0, "both", k, reps=2)


if __name__ == "__main__":
    main()
