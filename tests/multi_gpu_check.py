#!/usr/bin/env python
"""
Multi-GPU parity check (run under torchrun, one rank per GPU; not collected by pytest directly --
tests/test_gpu_multi.py launches it when the box has at least two GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu_check.py

Every rank builds the same seeded genome, the ranks sort + count it as ONE key-range sharded index
(both exchange modes), and rank 0 compares the concatenated shards and the histogram with the CPU oracle.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "genome-kmers_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import gpu_utils as gu
    import oracle
    from genome_kmers import distributed as gkd

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    same_gpu = "--same-gpu" in sys.argv
    if same_gpu:
        # every rank on cuda:0 (the driver's single-GPU box): gloo carries the small collectives (NCCL refuses two
        # ranks on one device), CUDA IPC carries the pairs exactly as it does between GPUs
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world = dist.get_world_size()
    failures = []
    cases = [(31, 600_000, 6, 8, False), (21, 400_000, 3, 0, False), (12, 300_000, 2, 3, False),
             (31, 500_000, 5, 6, True),                      # 64-bit start indices (GK_FORCE_IDX64)
             (40, 400_000, 4, 6, False), (64, 300_000, 3, 4, False), (33, 200_000, 2, 0, False),   # k > one key word
             (47, 250_000, 3, 3, True)]                      # ... with 64-bit start indices
    for k, n_bases, n_rec, runs, wide in cases:
        os.environ["GK_FORCE_IDX64"] = "1" if wide else "0"
        rng = np.random.default_rng(1000 + k)
        recs = gu.random_genome(rng, n_bases, n_rec, n_runs=runs, run_lo=50, run_hi=5000,
                                n_scatter=20 if runs else 0)
        if k > 31:
            # copies of 35..400 symbols: k-mers that agree on the first 31 symbols (one key word) and differ
            # later, or not at all
            for _ in range(60):
                seq = recs[int(rng.integers(0, len(recs)))][1]
                ln = int(rng.integers(35, 400))
                src, dst = (int(v) for v in rng.integers(0, len(seq) - ln, 2))
                seq[dst:dst + ln] = seq[src:src + ln].copy()
        sba = np.concatenate([np.concatenate([seq, np.array([36], dtype=np.uint8)]) for _, seq in recs])[:-1]
        starts = np.cumsum([0] + [len(seq) + 1 for _, seq in recs[:-1]]).astype(np.uint64)
        both, both_starts = oracle.both_strands(sba, starts)
        want = hist_want = total_want = None
        if rank == 0:
            want = oracle.sort_indices(both, oracle.init_indices(both_starts, len(both), k), k, k,
                                       threads=min(8, oracle.max_threads()))
            hist_want, total_want = oracle.group_hist(both, want, k, max_bin=1000)
        for mode in ("1", "0"):
            os.environ["GK_PEER_EXCHANGE"] = mode
            sk = gkd.ShardedKmers(sba, starts, k, "both")
            sk.sort()
            hist, total = sk.get_kmer_group_counts(k, max_counts_bin=1000)
            got = sk.gather_start_indices(0)
            used = sk.exchange_mode
            ver = sk.verify(hist, total)
            sk.close()
            if rank == 0:
                ok = (len(got) == len(want) and np.array_equal(got.astype(np.uint64), want)
                      and total == total_want and np.array_equal(hist, hist_want)
                      and all(ver["checks"].values()))
                if not all(ver["checks"].values()):
                    print("verify failed:", [name for name, v in ver["checks"].items() if not v], flush=True)
                print(f"k={k} world={world} exchange={used} idx64={wide}: {'ok' if ok else 'MISMATCH'} "
                      f"({len(got)} k-mers, {int(hist.sum())} distinct)", flush=True)
                if not ok:
                    failures.append((k, used))
                if mode == "1" and used != "peer":
                    failures.append((k, "peer exchange was not used"))
    # the drop-in class itself with num_gpus: same answers as the oracle on every rank
    from genome_kmers.kmers import Kmers, gen_no_ambiguous_bases_filter
    from genome_kmers.sequence_collection import SequenceCollection

    os.environ["GK_FORCE_IDX64"] = "0"
    os.environ["GK_PEER_EXCHANGE"] = "1"
    rng = np.random.default_rng(5)
    recs = gu.random_genome(rng, 300_000, 3, n_runs=4, run_lo=50, run_hi=3000, n_scatter=10)
    sc = SequenceCollection.from_arrays(recs, strands_to_load="both")
    km = Kmers(sc, 21, 21, source_strand="both", num_gpus=world)
    km.sort()
    hist, total = km.get_kmer_group_counts(21, max_counts_bin=100)
    p_hist, p_total = km.get_kmer_group_counts(21, gen_no_ambiguous_bases_filter(21), max_counts_bin=100)
    got = km.kmer_sba_start_indices
    both, both_starts = oracle.both_strands(sc.forward_sba, sc._forward_sba_seg_starts.astype(np.uint64))
    want = oracle.sort_indices(both, oracle.init_indices(both_starts, len(both), 21), 21, 21,
                               threads=min(8, oracle.max_threads()))
    o_hist, o_total = oracle.group_hist(both, want, 21, max_bin=100)
    q_hist, q_total = oracle.group_hist(both, want, 21, filt=(oracle.FILTER_NO_AMBIGUOUS, 21, 0, 0), max_bin=100)
    ok = (np.array_equal(got.astype(np.uint64), want) and total == o_total and np.array_equal(hist, o_hist)
          and p_total == q_total and np.array_equal(p_hist, q_hist) and len(km.local_start_indices()) <= len(want))
    print(f"rank {rank}: Kmers(num_gpus={world}): {'ok' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        print(f"rank {rank}: order {np.array_equal(got.astype(np.uint64), want)} ({len(got)} / {len(want)}), counts "
              f"{total == o_total} {np.array_equal(hist, o_hist)}, pure counts {p_total == q_total} ({p_total} / {q_total}) "
              f"{np.array_equal(p_hist, q_hist)}", flush=True)
        if rank == 0:
            bad = np.flatnonzero(got.astype(np.uint64) != want)
            print(f"rank 0: {len(bad)} mismatching slots, first {bad[:6]}, last {bad[-3:]}", flush=True)
            for b in bad[:6]:
                print(f"   slot {b}: got {got[b]} {both[int(got[b]):int(got[b]) + 21].tobytes()} want {want[b]} "
                      f"{both[int(want[b]):int(want[b]) + 21].tobytes()}", flush=True)
            print(f"rank 0: stats {km.last_sort_stats}", flush=True)
            x = int(want[bad[0]])
            where = np.flatnonzero(got.astype(np.uint64) == x)
            vals, cnts = np.unique(got, return_counts=True)
            print(f"rank 0: start {x} sits at slots {where.tolist()} of got; duplicated starts {vals[cnts > 1][:5].tolist()}; "
                  f"last slots got {got[-3:].tolist()} want {want[-3:].tolist()}; splitters {km._sk.splitters.tolist()}",
                  flush=True)
    print(f"rank {rank}: shard stats refine_flags {km.last_sort_stats.get('refine_flags')} n_shard "
          f"{km.last_sort_stats.get('n_shard')} amb {km.last_sort_stats.get('n_ambiguous')} frag "
          f"{km.last_sort_stats.get('n_fragments')} key_bits {km.last_sort_stats.get('key_bits')}", flush=True)
    if not ok:
        failures.append(("Kmers num_gpus", rank))
    # a larger case than the oracle can answer: checked on the devices against the bytes (ShardedKmers.verify)
    os.environ["GK_FORCE_IDX64"] = "0"
    os.environ["GK_PEER_EXCHANGE"] = "1"
    sys.path.insert(0, ROOT)
    import bench

    big, big_starts, _ = bench.make_genome(12_000_000 * world, 4 * world, 6, 42)
    sk = gkd.ShardedKmers(big, big_starts, 31, "both")
    sk.sort()
    hist, total = sk.get_kmer_group_counts(31)
    ver = sk.verify(hist, 2 * (12_000_000 * world - 4 * world * 30))
    sk.close()
    if rank == 0:
        bad = [name for name, v in ver["checks"].items() if not v]
        print(f"verify at {total} k-mers over {world} ranks: {'ok' if not bad else bad} {ver['report']}", flush=True)
        if bad:
            failures.append(("verify", bad))
    flag = torch.tensor([len(failures)])
    if not same_gpu:
        flag = flag.cuda()
    dist.all_reduce(flag)
    gkd.PeerExchange.close_all()
    dist.barrier()
    dist.destroy_process_group()
    if int(flag.item()):
        print("FAILED", failures, flush=True)
        sys.exit(1)
    if rank == 0:
        print("multi-GPU parity ok", flush=True)


if __name__ == "__main__":
    main()
