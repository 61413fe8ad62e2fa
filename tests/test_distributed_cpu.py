"""
World-size-2 (and 3) gloo tests of the multi-GPU orchestration on CPU: slicing, splitter
choice, count + pair exchange, shard concatenation, histogram all-reduce.  The compute steps are served by
tests/numpy_engine.py (a test double); the GPU kernels behind the same calls are covered by the gpu tests.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, strands, k, out_dir, kind="mixed"):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "genome-kmers_b200"), HERE]
    import torch.distributed as dist

    import oracle
    from genome_kmers.distributed import ShardedKmers
    from numpy_engine import NumpyEngine

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        recs = []
        for n in (700, 900, 650):
            seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
            seq[100:140] = ord("N")
            seq[rng.integers(0, n, 3)] = ord("R")
            recs.append(seq)
        recs.append(np.frombuffer(b"ACAC" * 60, dtype=np.uint8))     # a low-complexity record: big groups
        if kind == "one_group":      # every k-mer is the same: one rank owns everything, the others get nothing
            recs = [np.full(400, ord("A"), dtype=np.uint8), np.full(300, ord("A"), dtype=np.uint8)]
        elif kind == "tiny":         # fewer k-mers than ranks can share evenly
            recs = [np.frombuffer(b"ACGTTGCA"[:k + 1], dtype=np.uint8)]
        sba, starts = oracle.build_sba(recs)
        sk = ShardedKmers(sba, starts, k, strands, engine=NumpyEngine())
        sk.sort()
        hist, total = sk.get_kmer_group_counts(k, max_counts_bin=50)
        everything = sk.gather_start_indices(0)
        shard = sk.local_start_indices()
        np.save(os.path.join(out_dir, f"shard{rank}.npy"), shard)
        if rank == 0:
            full_sba, full_starts = (oracle.both_strands(sba, starts.astype(np.uint64)) if strands == "both"
                                     else (sba, starts.astype(np.uint64)))
            want = oracle.sort_indices(full_sba, oracle.init_indices(full_starts, len(full_sba), k), k, k)
            o_hist, o_total = oracle.group_hist(full_sba, want, k, max_bin=50)
            assert np.array_equal(everything.astype(np.uint64), want), "concatenated shards != global order"
            assert total == o_total and np.array_equal(hist, o_hist)
            np.save(os.path.join(out_dir, "ok.npy"), np.array([len(want)]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,strands,k", [(2, "forward", 11), (2, "both", 21), (3, "both", 5),
                                             (2, "both", 40)])      # 40: longer than one key word
def test_sharded_sort_count_matches_oracle(tmp_path, world, strands, k):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, strands, k, str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(tmp_path / "ok.npy")
    sizes = [len(np.load(tmp_path / f"shard{r}.npy")) for r in range(world)]
    assert sum(sizes) == int(np.load(tmp_path / "ok.npy")[0])
    assert min(sizes) > 0, f"a rank received nothing: {sizes}"


@pytest.mark.parametrize("world,strands,k,kind", [(2, "forward", 7, "one_group"), (3, "both", 7, "one_group"),
                                                  (2, "forward", 5, "tiny")])
def test_sharded_sort_count_degenerate_inputs(tmp_path, world, strands, k, kind):
    """A rank that receives nothing (one giant group, or almost no k-mers) must still take part in every
    collective and report an empty shard."""
    port = _free_port()
    mp.spawn(_worker, args=(world, port, strands, k, str(tmp_path), kind), nprocs=world, join=True)
    assert os.path.exists(tmp_path / "ok.npy")
    sizes = [len(np.load(tmp_path / f"shard{r}.npy")) for r in range(world)]
    assert sum(sizes) == int(np.load(tmp_path / "ok.npy")[0])


def test_splitters_and_slices():
    sys.path[:0] = [os.path.join(ROOT, "genome-kmers_b200")]
    from genome_kmers.distributed import choose_splitters, slice_bounds

    samples = np.arange(100, dtype=np.uint64) * np.uint64(7)
    sp = choose_splitters(samples, 4)
    assert sp.tolist() == [175, 350, 525]
    assert choose_splitters(samples, 1).size == 0
    bounds = [slice_bounds(1001, 4, r) for r in range(4)]
    assert bounds[0][0] == 0 and bounds[-1][1] == 1001
    assert all(a[1] == b[0] for a, b in zip(bounds[:-1], bounds[1:]))


def test_cost_balanced_splitters_around_a_giant_group():
    """One N run = millions of ambiguous windows with ONE key (class bit 0).  A key splitter cannot cut that
    group; the block-aware, bottleneck-optimal cuts give its owner fewer other keys instead (with any weight
    for an ambiguous window: 1 now that they sort as fragments, 2.5 when they were refined one by one)."""
    sys.path[:0] = [os.path.join(ROOT, "genome-kmers_b200")]
    from genome_kmers.distributed import choose_splitters

    rng = np.random.default_rng(3)
    m, parts = 8 * 4096, 8
    pure = (np.sort(rng.integers(0, 1 << 62, m, dtype=np.uint64)) << np.uint64(1)) | np.uint64(1)
    giant = np.full(int(0.044 * m), np.uint64(3) << np.uint64(61), dtype=np.uint64)   # 4.4 % of the samples
    samples = np.sort(np.concatenate([pure, giant]))
    for cost in (1.0, 2.5):
        sp = choose_splitters(samples, parts, class_bit=1, ambiguous_cost=cost)
        assert len(sp) == parts - 1 and bool((sp[1:] >= sp[:-1]).all())
        dest = np.searchsorted(sp, samples, side="right")
        weight = np.where((samples & np.uint64(1)) == 0, cost, 1.0)
        share = np.array([weight[dest == r].sum() for r in range(parts)]) / weight.sum() * parts
        assert share.max() < 1.08, share                       # even quantiles by count give the owner 1.3+
        assert len(set(dest[samples == giant[0]].tolist())) == 1    # the group stays on one rank
    # without heavy keys no part is more than 3 % above an even share
    sp0 = choose_splitters(pure, parts, class_bit=1)
    d0 = np.searchsorted(sp0, pure, side="right")
    assert np.bincount(d0, minlength=parts).max() < 1.03 * m / parts
    # degenerate inputs
    assert choose_splitters(np.full(16, 5, dtype=np.uint64), 4, class_bit=1).size == 3
    assert choose_splitters(np.zeros(0, dtype=np.uint64), 4, class_bit=1).size == 0


def test_key_ranges_are_exact_in_uint64():
    """Splitters are 64-bit keys: the bounds must not pass through float64 (53 bits), or a key equal to its
    rank's first key would wrap around when the first key is subtracted."""
    sys.path[:0] = [os.path.join(ROOT, "genome-kmers_b200")]
    from genome_kmers.distributed import key_ranges

    sp = np.array([(1 << 61) + 2, (1 << 63) + 1022, (1 << 64) - 4], dtype=np.uint64)
    lo, hi, bits = key_ranges(sp, 64)
    assert lo.dtype == np.uint64 and hi.dtype == np.uint64
    assert [int(v) for v in lo] == [0, (1 << 61) + 2, (1 << 63) + 1022, (1 << 64) - 4]
    assert [int(v) for v in hi] == [(1 << 61) + 2, (1 << 63) + 1022, (1 << 64) - 4, 0]
    assert bits == [62, 63, 63, 2]
    lo1, hi1, bits1 = key_ranges(np.zeros(0, dtype=np.uint64), 44)
    assert [int(v) for v in lo1] == [0] and bits1 == [44]
