#!/usr/bin/env python
"""Debug helper: sharded sort of a bench genome under torchrun, per-rank verify reports, first violation.
    torchrun --nproc-per-node 2 tools/shard_verify.py [--same-gpu] [--bases N]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genome-kmers_b200")]


def main():
    import torch
    import torch.distributed as dist

    import bench
    from genome_kmers import distributed as gkd

    rank = int(os.environ.get("RANK", "0"))
    same = "--same-gpu" in sys.argv
    bases = int(sys.argv[sys.argv.index("--bases") + 1]) if "--bases" in sys.argv else 20_000_000
    if same:
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    world = dist.get_world_size()
    k = 31
    sba, starts, _ = bench.make_genome(bases * world, 10 * world, 20, 42)
    sk = gkd.ShardedKmers(sba, starts, k, "both")
    sk.sort()
    hist, total = sk.get_kmer_group_counts(k)
    rep, first, last, bits = sk.engine.shard_verify(sk, k)
    print(f"rank {rank}: report {rep} stats {sk.stats}", flush=True)
    if rep[1]:
        idx = sk.local_start_indices().astype(np.int64)
        both = sk.d_sba.cpu().numpy()
        ar = np.arange(k)
        chunk = 1 << 21
        for lo in range(0, len(idx) - 1, chunk):
            hi = min(len(idx), lo + chunk + 1)
            win = both[idx[lo:hi, None] + ar[None, :]].astype(np.int16)
            d = win[:-1] != win[1:]
            first_d = np.where(d.any(axis=1), d.argmax(axis=1), k - 1)
            av = np.take_along_axis(win[:-1], first_d[:, None], 1)[:, 0]
            bv = np.take_along_axis(win[1:], first_d[:, None], 1)[:, 0]
            bad = np.flatnonzero(av > bv)
            for j in bad[:5]:
                r = lo + j + 1
                print(f"rank {rank}: violation at slot {r} of {len(idx)}: "
                      f"{both[idx[r-1]:idx[r-1]+k].tobytes()} > {both[idx[r]:idx[r]+k].tobytes()} starts {idx[r-1]} {idx[r]}"
                      f" next {both[idx[r+1]:idx[r+1]+k].tobytes() if r + 1 < len(idx) else None}", flush=True)
    dist.barrier()
    gkd.PeerExchange.close_all()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
