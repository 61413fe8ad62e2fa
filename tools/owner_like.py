#!/usr/bin/env python
"""Single-GPU stand-in for the rank that owns the all-N group of an 8-GPU sort (tuning aid): a 100 Mbp
genome of which `frac` is one kind of N run, both strands, k = 31 -- stage times and repair flags of the sort.
    python tools/owner_like.py [frac]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genome-kmers_b200")]


def main():
    import torch

    import bench
    from genome_kmers.kmers import Kmers
    from genome_kmers.sequence_collection import SequenceCollection

    frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.35
    sba, starts, names = bench.make_genome(100_000_000, 10, 0, 1)
    rng = np.random.default_rng(2)
    bounds = list(starts.astype(np.int64)) + [len(sba) + 1]
    for r in range(10):                     # 20 runs per record covering `frac` of it
        ln = int((bounds[r + 1] - 1 - bounds[r]) * frac / 20)
        for s in rng.integers(bounds[r], bounds[r + 1] - 1 - ln, 20):
            sba[int(s):int(s) + ln] = ord("N")
    recs = [(names[r], sba[bounds[r]:bounds[r + 1] - 1]) for r in range(10)]
    sc = SequenceCollection.from_arrays(recs, strands_to_load="both")
    km = Kmers(sc, 31, 31, source_strand="both")
    for it in range(4):
        km._is_sorted = False
        torch.cuda.synchronize()
        km.sort()
        st = km.last_sort_stats
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}, flush=True)
    rep = km.verify_order(31)
    print("verify", rep, flush=True)
    del km
    # the same input through the sharded path with one rank: slice pack, staging partition, placeholders from
    # the fragments, shard sort -- what the owner of the all-N group runs
    import tempfile

    import torch.distributed as dist

    from genome_kmers.distributed import ShardedKmers

    store = dist.FileStore(os.path.join(tempfile.mkdtemp(), "store"), 1)
    dist.init_process_group("gloo", store=store, rank=0, world_size=1)
    sk = ShardedKmers(sc.forward_sba, sc._forward_sba_seg_starts.astype(np.uint64), 31, "both")
    for it in range(4):
        torch.cuda.synchronize()
        sk.sort()
        torch.cuda.synchronize()
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in sk.stats.items()},
              {k: round(v, 3) for k, v in sk.phase_ms().items()}, flush=True)
    hist, total = sk.get_kmer_group_counts(31)
    print("sharded verify", sk.verify(hist, total)["checks"], flush=True)
    sk.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
