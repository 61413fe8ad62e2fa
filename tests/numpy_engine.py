"""NumPy stand-in for genome_kmers.distributed.NativeEngine -- a TEST DOUBLE for the CPU (gloo) tests of
the multi-GPU orchestration.  It lives under tests/ on purpose: the product has no CPU path.  It has no
peer memory and no fragment lists, so the driver takes the staging + all-to-all exchange and sends the
ambiguous windows as pairs; keys are made relative to the destination's key range like the CUDA kernel does."""
import ctypes

import numpy as np
import torch

import oracle
from genome_kmers.distributed import PackedSlice
from gpu_utils import expected_key


class NumpyEngine:
    supports_fragments = False

    def to_device(self, host_u8):
        return torch.from_numpy(np.ascontiguousarray(host_u8))

    def both_strands(self, d_fwd):
        sba = d_fwd.numpy()
        out, _ = oracle.both_strands(sba, np.array([0], dtype=np.uint64))
        return torch.from_numpy(out)

    def alphabet_async(self, d_sba):
        sba = d_sba.numpy()
        acgt = np.isin(sba, np.frombuffer(b"ACGT", dtype=np.uint8))
        sep = sba == 36
        allowed = np.isin(sba, np.frombuffer(b"ACGTRYSWKMBDHVN$", dtype=np.uint8))
        return torch.tensor([int((~allowed).sum()), int(sep.sum()), int((allowed & ~acgt & ~sep).sum())],
                            dtype=torch.int64)

    def _slice_starts(self, d_sba, seg_starts, k, first, end):
        starts = oracle.init_indices(seg_starts, len(d_sba), k)
        return starts[(starts >= first) & (starts < end)]

    def pack_slice(self, d_sba, seg_starts, k, class_bit, first, end, idx_bytes):
        raw = d_sba.numpy().tobytes()
        starts = self._slice_starts(d_sba, seg_starts, k, first, end)
        kl = min(k, 31) if class_bit else k   # k > 31: the key covers the first 31 symbols
        keys = np.array([expected_key(raw[int(i):int(i) + kl], class_bit) for i in starts], dtype=np.uint64)
        idx = starts.astype(np.uint32 if idx_bytes == 4 else np.uint64)
        n_amb = int((keys & np.uint64(1) == 0).sum()) if class_bit else 0
        return PackedSlice(torch.from_numpy(keys.view(np.int64).copy()),
                           torch.from_numpy(idx.view(np.int32 if idx_bytes == 4 else np.int64).copy()), len(keys),
                           None, torch.tensor([n_amb, 0, 0, 0], dtype=torch.int64))

    def sample_keys_host(self, d_sba, seg_starts, k, class_bit, first, end, n_samples):
        raw = d_sba.numpy().tobytes()
        starts = self._slice_starts(d_sba, seg_starts, k, first, end)
        n = min(n_samples, len(starts))
        pick = [starts[(j * len(starts)) // n] for j in range(n)]
        kl = min(k, 31) if class_bit else k
        return np.array([expected_key(raw[int(i):int(i) + kl], class_bit) for i in pick], dtype=np.uint64)

    def splitters_to_device(self, splitters_host):
        return torch.from_numpy(np.ascontiguousarray(splitters_host).view(np.int64).copy())

    @staticmethod
    def _dest(keys_u64, splitters):
        sp = splitters.numpy().view(np.uint64) if splitters is not None else np.zeros(0, np.uint64)
        return np.searchsorted(sp, keys_u64, side="right")

    def partition_counts_host(self, pk, splitters, n_parts, class_bit, extra=()):
        keys = pk.keys.numpy().view(np.uint64)
        dest = self._dest(keys, splitters)
        amb = (keys & np.uint64(1)) == 0 if class_bit else np.zeros(len(keys), dtype=bool)
        pure = np.bincount(dest[~amb], minlength=n_parts)
        ambc = np.bincount(dest[amb], minlength=n_parts)
        parts = [pure, ambc, [int(pk.counters[2])], [int(pk.counters[0])]] + [np.asarray(e) for e in extra]
        return np.concatenate(parts).astype(np.int64)

    def partition_to_staging(self, pk, splitters, n_parts, send_counts, key_base, skip_amb):
        assert not skip_amb
        keys = pk.keys.numpy().view(np.uint64)
        dest = self._dest(keys, splitters)
        order = np.argsort(dest, kind="stable")
        assert np.array_equal(np.bincount(dest, minlength=n_parts), np.asarray(send_counts))
        rel = keys - np.asarray(key_base, dtype=np.uint64)[dest]
        return (torch.from_numpy(rel[order].view(np.int64).copy()), pk.idx[torch.from_numpy(order)])

    def recv_buffers(self, ref_idx, n):
        return torch.empty(max(1, n), dtype=torch.int64), torch.empty(max(1, n), dtype=ref_idx.dtype)

    @staticmethod
    def tensor_ptr(t):
        return t          # the double passes tensors where the CUDA engine passes device pointers

    def shard_sort(self, d_sba, seg_starts, k, keys, idx, idx_bytes, n_pure, n_amb, class_bit, key_bits, frag_all,
                   frag_counts, key_lo, key_hi):
        assert frag_all is None and n_amb == 0
        sba = d_sba.numpy()
        rel = keys.numpy().view(np.uint64)[:n_pure]
        assert n_pure == 0 or int(rel.max()).bit_length() <= key_bits, "keys are not relative to the rank's range"
        starts = idx.numpy().view(np.uint32 if idx_bytes == 4 else np.uint64)[:n_pure].astype(np.uint64)
        # received pairs arrive grouped by source rank with ascending starts inside each group; a stable
        # sort by k-mer then leaves ties in ascending start order, like the GPU path
        srt = oracle.sort_indices(sba, starts, k, k)
        return {"handle": None, "n": len(srt), "sorted": srt, "sba": sba, "stats": {"sort_passes": 0, "sort_ms": 0.0},
                "idx_bytes": idx_bytes}

    def shard_counts(self, shard, k, filt, min_group, max_group, max_bin):
        return oracle.group_hist(shard["sba"], shard["sorted"], k, min_group=min_group, max_group=max_group,
                                 max_bin=max_bin)

    def shard_indices_host(self, shard):
        return shard["sorted"].astype(np.uint32 if shard["idx_bytes"] == 4 else np.uint64)

    def shard_free(self, shard):
        pass

    def from_host_i64(self, arr):
        return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64).copy())
