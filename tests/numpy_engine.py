"""NumPy stand-in for genome_kmers.distributed.NativeEngine -- a TEST DOUBLE for the CPU (gloo) tests of
the multi-GPU orchestration.  It lives under tests/ on purpose: the product has no CPU path."""
import numpy as np
import torch

import oracle
from gpu_utils import expected_key


class NumpyEngine:
    def to_device(self, host_u8):
        return torch.from_numpy(np.ascontiguousarray(host_u8))

    def both_strands(self, d_fwd):
        sba = d_fwd.numpy()
        out, _ = oracle.both_strands(sba, np.array([0], dtype=np.uint64))
        return torch.from_numpy(out)

    def alphabet(self, d_sba):
        sba = d_sba.numpy()
        acgt = np.isin(sba, np.frombuffer(b"ACGT", dtype=np.uint8))
        sep = sba == 36
        allowed = np.isin(sba, np.frombuffer(b"ACGTRYSWKMBDHVN$", dtype=np.uint8))
        return np.array([(~allowed).sum(), sep.sum(), (allowed & ~acgt & ~sep).sum()], dtype=np.uint64)

    def pack_slice(self, d_sba, seg_starts, k, class_bit, first, end, idx_bytes):
        sba = d_sba.numpy()
        raw = sba.tobytes()
        starts = oracle.init_indices(seg_starts, len(sba), k)
        starts = starts[(starts >= first) & (starts < end)]
        keys = np.array([expected_key(raw[int(i):int(i) + k], class_bit) for i in starts], dtype=np.uint64)
        idx = starts.astype(np.uint32 if idx_bytes == 4 else np.uint64)
        return (torch.from_numpy(keys.view(np.int64).copy()),
                torch.from_numpy(idx.view(np.int32 if idx_bytes == 4 else np.int64).copy()))

    def sort_keys(self, keys):
        return torch.from_numpy(np.sort(keys.numpy().view(np.uint64)).view(np.int64).copy())

    def partition(self, keys, idx, splitters, n_parts):
        k = keys.numpy().view(np.uint64)
        sp = splitters.numpy().view(np.uint64) if splitters is not None else np.zeros(0, np.uint64)
        dest = np.searchsorted(sp, k, side="right")
        order = np.argsort(dest, kind="stable")
        counts = np.bincount(dest, minlength=n_parts).astype(np.int64)
        return keys[torch.from_numpy(order)], idx[torch.from_numpy(order)], counts

    def empty_like_n(self, ref, n):
        return torch.empty(n, dtype=ref.dtype)

    def from_host_i64(self, arr):
        return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64).copy())

    def shard_index(self, d_sba, seg_starts, k, keys, idx, class_bit):
        sba = d_sba.numpy()
        starts = idx.numpy().view(np.uint32 if idx.element_size() == 4 else np.uint64).astype(np.uint64)
        # received pairs arrive grouped by source rank with ascending starts inside each group; a stable
        # sort by k-mer then leaves ties in ascending start order, like the GPU path
        srt = oracle.sort_indices(sba, starts, k, k)
        return {"handle": None, "n": len(srt), "sorted": srt, "sba": sba, "stats": {"sort_passes": 0, "sort_ms": 0.0},
                "idx_bytes": idx.element_size()}

    def shard_counts(self, shard, k, filt, min_group, max_group, max_bin):
        return oracle.group_hist(shard["sba"], shard["sorted"], k, min_group=min_group, max_group=max_group,
                                 max_bin=max_bin)

    def shard_indices_host(self, shard):
        return shard["sorted"].astype(np.uint32 if shard["idx_bytes"] == 4 else np.uint64)

    def shard_free(self, shard):
        pass
