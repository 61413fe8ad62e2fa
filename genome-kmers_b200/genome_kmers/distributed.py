"""
Multi-GPU driver for the hot path: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch) for the plumbing, libgkb200 kernels for every compute step.

The reference has no parallel code at all; this is new design.  The k-mer windows of the indexed byte
array are range-partitioned by key so that equal k-mers always meet on one GPU:

  1. every rank holds the whole byte array (a few GB at most) and packs the windows of ITS slice of
     start positions into (key, start) pairs                                   gk_pack_keys
  2. evenly spaced key samples are all-gathered; every rank sorts them and picks the same G-1 splitters
  3. one stable partition pass groups the pairs by destination rank            gk_partition_pairs
  4. pair counts are exchanged (tiny all-to-all), then the pairs themselves: ONE variable-size
     all-to-all for the keys and one for the starts                            all_to_all_single
  5. every rank sorts its key range, refines ambiguous windows, flags groups   gk_index_sort_pairs
  6. histograms are summed with an all-reduce; the global sorted order is the concatenation of the
     ranks' shards in rank order.
Splitters are keys (not (key, start) pairs), so a group of equal k-mers never straddles two ranks and
no boundary fix-up is needed; the price is that one giant group cannot be split (documented skew).
Ties stay in ascending start order: source ranks hold ascending slices, the partition and the sort are
stable, and all_to_all_single concatenates by source rank.

The compute steps go through an `engine` object.  The default engine calls the CUDA library and needs
a GPU; tests on CPU (gloo, world_size 2) inject a NumPy engine to exercise the orchestration only.
"""
import ctypes
import os
from typing import Optional

import numpy as np

from genome_kmers import _native

SAMPLES_PER_RANK = 2048


def _torch():
    import torch

    return torch


class NativeEngine:
    """Compute steps on the current CUDA device through the C ABI."""

    def __init__(self):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("genome_kmers.distributed needs a CUDA device per rank (no CPU fallback)")
        self.torch = torch
        self.lib = _native.lib()
        self.device = torch.device("cuda", torch.cuda.current_device())

    def stream(self):
        return int(self.torch.cuda.current_stream().cuda_stream)

    def mark(self):
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(self.torch.cuda.current_stream())
        return ev

    def to_device(self, host_u8: np.ndarray):
        return self.torch.from_numpy(np.ascontiguousarray(host_u8)).to(self.device, non_blocking=True)

    def both_strands(self, d_fwd):
        out = self.torch.empty(2 * d_fwd.numel() + 1, dtype=self.torch.uint8, device=self.device)
        _native.check(self.lib.gk_sba_both_strands(d_fwd.data_ptr(), d_fwd.numel(), out.data_ptr(), self.stream()))
        return out

    def alphabet(self, d_sba):
        counts = np.zeros(3, dtype=np.uint64)
        _native.check(self.lib.gk_sba_scan_alphabet(d_sba.data_ptr(), d_sba.numel(), _native.host_ptr(counts),
                                                    self.stream()))
        return counts

    def pack_slice(self, d_sba, seg_starts, k, class_bit, first, end, idx_bytes):
        torch = self.torch
        cap = max(1, end - first)
        keys = torch.empty(cap, dtype=torch.int64, device=self.device)
        idx = torch.empty(cap, dtype=torch.int32 if idx_bytes == 4 else torch.int64, device=self.device)
        n_out, n_amb = ctypes.c_uint64(0), ctypes.c_uint64(0)
        _native.check(self.lib.gk_pack_keys(d_sba.data_ptr(), d_sba.numel(), _native.host_ptr(seg_starts),
                                            len(seg_starts), k, k, class_bit, first, end, keys.data_ptr(),
                                            idx_bytes, idx.data_ptr(), cap, ctypes.byref(n_out),
                                            ctypes.byref(n_amb), self.stream()))
        return keys[:n_out.value], idx[:n_out.value]

    def sort_keys(self, keys):
        """Ascending (unsigned) order of a small key tensor."""
        torch = self.torch
        n = keys.numel()
        if n < 2:
            return keys
        a = torch.empty(n + (n & 1), dtype=torch.int64, device=self.device)
        a[:n] = keys
        b = torch.empty_like(a)
        v0 = torch.zeros(n, dtype=torch.int32, device=self.device)
        v1 = torch.empty_like(v0)
        in_alt = ctypes.c_int(0)
        _native.check(self.lib.gk_radix_sort_pairs(a.data_ptr(), b.data_ptr(), v0.data_ptr(), v1.data_ptr(), 4, n,
                                                   0, 64, ctypes.byref(in_alt), self.stream()))
        return (b if in_alt.value else a)[:n]

    def partition(self, keys, idx, splitters, n_parts):
        torch = self.torch
        n = keys.numel()
        k_out, i_out = torch.empty_like(keys), torch.empty_like(idx)
        counts = np.zeros(n_parts, dtype=np.uint64)
        sp = splitters.data_ptr() if splitters is not None and splitters.numel() else None
        _native.check(self.lib.gk_partition_pairs(keys.data_ptr(), k_out.data_ptr(), idx.data_ptr(),
                                                  i_out.data_ptr(), idx.element_size(), n, sp, n_parts,
                                                  _native.host_ptr(counts), self.stream()))
        return k_out, i_out, counts.astype(np.int64)

    def empty_like_n(self, ref, n):
        return self.torch.empty(n, dtype=ref.dtype, device=self.device)

    # ---- fused partition + exchange over peer memory ------------------------------------------------
    def peer_exchange(self, dist, group, capacity, idx_bytes):
        return PeerExchange.get(self, dist, group, capacity, idx_bytes)

    def partition_count(self, keys, splitters, n_parts):
        counts = np.zeros(n_parts, dtype=np.uint64)
        sp = splitters.data_ptr() if splitters is not None and splitters.numel() else None
        _native.check(self.lib.gk_partition_count(keys.data_ptr(), keys.numel(), sp, n_parts,
                                                  _native.host_ptr(counts), self.stream()))
        return counts.astype(np.int64)

    def partition_peer(self, keys, idx, splitters, n_parts, px, offsets):
        sp = splitters.data_ptr() if splitters is not None and splitters.numel() else None
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        _native.check(self.lib.gk_partition_pairs_peer(
            keys.data_ptr(), idx.data_ptr(), idx.element_size(), keys.numel(), sp, n_parts,
            _native.host_ptr(px.key_ptrs), _native.host_ptr(px.idx_ptrs), _native.host_ptr(off), self.stream()))

    def shard_index_ptr(self, d_sba, seg_starts, k, keys_ptr, idx_ptr, n, idx_bytes, class_bit):
        """Sort pairs that already sit in library-owned buffers (the peer receive buffers)."""
        torch = self.torch
        handle = ctypes.c_void_p()
        _native.check(self.lib.gk_index_create(d_sba.data_ptr(), d_sba.numel(), _native.host_ptr(seg_starts),
                                               len(seg_starts), k, k, ctypes.byref(handle)))
        stats = _native.GkSortStats()
        k_alt = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        i_alt = torch.empty(max(n, 1), dtype=torch.int32 if idx_bytes == 4 else torch.int64, device=self.device)
        try:
            _native.check(self.lib.gk_index_sort_pairs(handle, keys_ptr, k_alt.data_ptr(), idx_ptr,
                                                       i_alt.data_ptr(), n, class_bit, ctypes.byref(stats),
                                                       self.stream()))
        except Exception:
            self.lib.gk_index_destroy(handle)
            raise
        return {"handle": handle, "n": n, "stats": stats.as_dict(), "idx_bytes": idx_bytes}

    def shard_index(self, d_sba, seg_starts, k, keys, idx, class_bit):
        """Sort the received pairs and return an opaque shard handle."""
        torch = self.torch
        handle = ctypes.c_void_p()
        _native.check(self.lib.gk_index_create(d_sba.data_ptr(), d_sba.numel(), _native.host_ptr(seg_starts),
                                               len(seg_starts), k, k, ctypes.byref(handle)))
        stats = _native.GkSortStats()
        n = keys.numel()
        k_alt, i_alt = torch.empty_like(keys), torch.empty_like(idx)
        try:
            _native.check(self.lib.gk_index_sort_pairs(handle, keys.data_ptr(), k_alt.data_ptr(), idx.data_ptr(),
                                                       i_alt.data_ptr(), n, class_bit, ctypes.byref(stats),
                                                       self.stream()))
        except Exception:
            self.lib.gk_index_destroy(handle)
            raise
        return {"handle": handle, "n": n, "stats": stats.as_dict(), "idx_bytes": idx.element_size()}

    def shard_counts(self, shard, k, filt, min_group, max_group, max_bin):
        hist = np.zeros(max_bin + 1, dtype=np.int64)
        total, top = ctypes.c_int64(0), ctypes.c_uint64(0)
        flt = filt if filt is not None else _native.GkFilter(0, 0, 0, 0)
        _native.check(self.lib.gk_index_group_counts_zeroed(shard["handle"], k, ctypes.byref(flt), min_group,
                                                            max_group or 0, max_bin, _native.host_ptr(hist),
                                                            ctypes.byref(total), ctypes.byref(top), self.stream()))
        return hist, int(total.value)

    def shard_counts_sparse(self, shard, k, filt, min_group, max_group, max_bin):
        """(bins uint64, counts int64, total): the occupied histogram bins only."""
        flt = filt if filt is not None else _native.GkFilter(0, 0, 0, 0)
        cap = 1 << 12
        while True:
            bins = np.empty(cap, dtype=np.uint64)
            counts = np.empty(cap, dtype=np.int64)
            n_pairs, total = ctypes.c_uint64(0), ctypes.c_int64(0)
            rc = self.lib.gk_index_group_counts_sparse(
                shard["handle"], k, ctypes.byref(flt), min_group, max_group or 0, max_bin, _native.host_ptr(bins),
                _native.host_ptr(counts), cap, ctypes.byref(n_pairs), ctypes.byref(total), self.stream())
            if rc == _native.GK_ERR_ARG and n_pairs.value > cap:
                cap = int(n_pairs.value)
                continue
            _native.check(rc)
            return bins[:n_pairs.value], counts[:n_pairs.value], int(total.value)

    def shard_indices_host(self, shard):
        torch = self.torch
        wide = shard["idx_bytes"] == 8
        # pinned staging (torch caches pinned blocks) so that the D2H copy runs at PCIe speed
        buf = torch.empty(shard["n"], dtype=torch.int64 if wide else torch.int32, pin_memory=True)
        if shard["n"]:
            _native.check(self.lib.gk_index_copy_indices(shard["handle"], buf.data_ptr(), self.stream()))
        return buf.numpy().view(np.uint64 if wide else np.uint32)

    def shard_free(self, shard):
        if shard and shard.get("handle") is not None:
            self.lib.gk_index_destroy(shard["handle"])
            shard["handle"] = None

    def as_dist_tensor(self, t):
        return t

    def from_host_i64(self, arr):
        return self.torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64)).to(self.device)


class PeerExchange:
    """Receive buffers of every rank, mapped into every rank (CUDA IPC over NVLink peer memory).

    The fused partition kernel (gk_partition_pairs_peer) writes each (key, start) pair directly into the
    buffer of the rank that owns its key range.  Buffers are allocated once per (group, capacity) and
    reused by later sorts; `capacity` counts pairs."""

    _cache = {}
    MAX_BYTES = 64 << 30   # per rank; beyond this the NCCL all-to-all path is used

    def __init__(self, engine, dist, group, capacity: int, idx_bytes: int):
        torch, lib = engine.torch, engine.lib
        self.lib, self.capacity, self.idx_bytes = lib, int(capacity), idx_bytes
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.my_keys, self.my_idx = ctypes.c_void_p(), ctypes.c_void_p()
        _native.check(lib.gk_peer_alloc(self.capacity * 8, ctypes.byref(self.my_keys)))
        _native.check(lib.gk_peer_alloc(self.capacity * idx_bytes, ctypes.byref(self.my_idx)))
        handles = np.zeros(128, dtype=np.uint8)
        _native.check(lib.gk_peer_export(self.my_keys, _native.host_ptr(handles)))
        _native.check(lib.gk_peer_export(self.my_idx, _native.host_ptr(handles[64:])))
        mine = torch.from_numpy(handles).to(engine.device)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=group)
        self.key_ptrs = np.zeros(self.world, dtype=np.uint64)
        self.idx_ptrs = np.zeros(self.world, dtype=np.uint64)
        self._opened = []
        for r, g in enumerate(gathered):
            if r == self.rank:
                self.key_ptrs[r], self.idx_ptrs[r] = self.my_keys.value, self.my_idx.value
                continue
            h = np.ascontiguousarray(g.cpu().numpy())
            pk, pi = ctypes.c_void_p(), ctypes.c_void_p()
            _native.check(lib.gk_peer_open(_native.host_ptr(h), ctypes.byref(pk)))
            _native.check(lib.gk_peer_open(_native.host_ptr(h[64:]), ctypes.byref(pi)))
            self._opened += [pk, pi]
            self.key_ptrs[r], self.idx_ptrs[r] = pk.value, pi.value

    @classmethod
    def get(cls, engine, dist, group, capacity: int, idx_bytes: int):
        key = (id(group), dist.get_world_size(group), idx_bytes)
        px = cls._cache.get(key)
        if px is None or px.capacity < capacity:   # every rank computes the same capacity: collective-safe
            if px is not None:
                px.close()
            px = cls(engine, dist, group, capacity, idx_bytes)
            cls._cache[key] = px
        return px

    def close(self):
        for p in self._opened:
            self.lib.gk_peer_close(p)
        self._opened = []
        if self.my_keys:
            self.lib.gk_peer_free(self.my_keys)
            self.lib.gk_peer_free(self.my_idx)
            self.my_keys = self.my_idx = None

    @classmethod
    def close_all(cls):
        for px in cls._cache.values():
            px.close()
        cls._cache.clear()


AMBIGUOUS_COST = 2.5   # local-sort cost of an ambiguous window relative to a pure one (refinement), measured


def choose_splitters(sorted_samples: np.ndarray, n_parts: int, class_bit: int = 0) -> np.ndarray:
    """n_parts-1 splitters at the even quantiles of the pooled, sorted samples (uint64).

    With class_bit the quantiles are taken over COST, not count: a key with class bit 0 is an ambiguous
    window, which also goes through the refinement.  One N run makes millions of them with ONE key, which no
    key splitter can cut, so the rank that gets that group is given correspondingly fewer other keys."""
    m = len(sorted_samples)
    if n_parts <= 1 or m == 0:
        return np.zeros(0, dtype=np.uint64)
    if class_bit:
        # blocks of equal keys (a part can only begin where a key begins) and their costs
        weight = np.where((sorted_samples & np.uint64(1)) == 0, AMBIGUOUS_COST, 1.0)
        first = np.flatnonzero(np.concatenate([[True], sorted_samples[1:] != sorted_samples[:-1]]))
        run = np.concatenate([[0.0], np.cumsum(weight)])      # run[i] = cost of samples [0, i)
        cum = np.concatenate([run[first], run[-1:]])          # cum[b] = cost of blocks [0, b)
        cost = np.diff(cum)
        n_blocks = len(cost)
        find = cum.searchsorted

        def pack(limit):
            """Greedy: fill every part up to `limit`; returns the first block of parts 1, 2, ..."""
            cuts, b = [], 0
            while b < n_blocks and len(cuts) < n_parts:
                e = int(find(cum[b] + limit, "right")) - 1    # blocks [b, e) fit
                if e <= b:
                    e = b + 1
                if e >= n_blocks:
                    return cuts, True
                cuts.append(e)
                b = e
            return cuts, False

        lo, hi = max(float(cost.max()), cum[-1] / n_parts), float(cum[-1])
        if float(cost.max()) * 50 < cum[-1] / n_parts:        # no heavy block: even shares, nothing to search
            lo = hi = 1.02 * cum[-1] / n_parts
        for _ in range(14 if hi > lo else 0):                 # smallest bottleneck that needs <= n_parts parts
            mid = 0.5 * (lo + hi)
            cuts, ok = pack(mid)
            if ok and len(cuts) <= n_parts - 1:
                hi = mid
            else:
                lo = mid
        cuts, _ = pack(hi)
        # a block heavier than an even share sets the bottleneck and leaves parts unused: split the heaviest
        # parts that still hold more than one block, nearest to their middle, until every rank has work
        bounds = [0] + cuts + [len(cost)]
        while len(bounds) - 1 < n_parts:
            spans = [(cum[bounds[i + 1]] - cum[bounds[i]], i) for i in range(len(bounds) - 1)
                     if bounds[i + 1] - bounds[i] > 1]
            if not spans:
                break
            _, i = max(spans)
            a, b = bounds[i], bounds[i + 1]
            half = int(np.searchsorted(cum, 0.5 * (cum[a] + cum[b]), side="left"))
            bounds.insert(i + 1, min(max(half, a + 1), b - 1))
        cuts = (bounds[1:-1] + [len(cost) - 1] * n_parts)[:n_parts - 1]
        pos = first[np.asarray(cuts, dtype=np.int64)]
    else:
        pos = (np.arange(1, n_parts, dtype=np.int64) * m) // n_parts
    return np.ascontiguousarray(sorted_samples[pos], dtype=np.uint64)


def slice_bounds(total_len: int, world: int, rank: int):
    """Contiguous slice of start positions owned by `rank` (windows may read k-1 bytes past it)."""
    return (total_len * rank) // world, (total_len * (rank + 1)) // world


class ShardedKmers:
    """Sort + count the k-mers of one byte array across the ranks of a process group.

    Every rank passes the same forward byte array (uint8, records joined by '$'), segment starts, k and
    strands ("forward" or "both").  After sort(), each rank owns the k-mers of one key range:
    local_start_indices() is that shard; concatenating the shards in rank order gives the global order.
    """

    def __init__(self, forward_sba, seg_starts, kmer_len: int, strands: str = "forward", group=None,
                 engine=None):
        import torch.distributed as dist

        if strands not in ("forward", "both"):
            raise ValueError(f"strands ({strands}) must be 'forward' or 'both'")
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.engine = engine or NativeEngine()
        self.k = int(kmer_len)
        self.strands = strands
        eng = self.engine
        d_fwd = forward_sba if not isinstance(forward_sba, np.ndarray) else eng.to_device(forward_sba)
        starts = np.ascontiguousarray(seg_starts, dtype=np.uint64)
        n = int(d_fwd.numel()) if hasattr(d_fwd, "numel") else len(d_fwd)
        if strands == "both":
            ends = np.concatenate([starts[1:].astype(np.int64) - 2, [n - 1]])
            rc = (n - 1 - ends[::-1]).astype(np.uint64) + np.uint64(n + 1)
            self.seg_starts = np.ascontiguousarray(np.concatenate([starts, rc]), dtype=np.uint64)
            self.d_sba = eng.both_strands(d_fwd)
            self.total_len = 2 * n + 1
        else:
            self.seg_starts = starts
            self.d_sba = d_fwd
            self.total_len = n
        self.idx_bytes = 8 if self.total_len > 0xFFFFFFFF else 4
        if os.environ.get("GK_FORCE_IDX64", "0") not in ("", "0"):   # tests: 64-bit starts on small inputs
            self.idx_bytes = 8
        self.shard = None
        self.stats = {}
        self._marks = []
        self._is_sorted = False

    # -------------------------------------------------------------------------------------------
    def _mark(self, name):
        """Phase boundary: a CUDA event on the current stream (no synchronise); see phase_ms()."""
        mk = getattr(self.engine, "mark", None)
        if mk is not None:
            self._marks.append((name, mk()))

    def phase_ms(self):
        """Device time between consecutive phase marks of the last sort() (call after a synchronise)."""
        out = {}
        for (_, a), (name, b) in zip(self._marks[:-1], self._marks[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    def sort(self):
        eng, dist = self.engine, self.dist
        k, world, rank = self.k, self.world, self.rank
        self._marks = []
        self._mark("start")
        if k > 31:
            raise NotImplementedError("the multi-GPU path handles single-word k-mers (k <= 31)")
        first, end = slice_bounds(self.total_len, world, rank)
        # (16-byte aligned cut points, so that the scan keeps its 128-bit loads)
        a16 = (first // 16) * 16 if rank > 0 else 0
        b16 = (end // 16) * 16 if rank < world - 1 else self.total_len
        counts = eng.alphabet(self.d_sba[a16:b16] if world > 1 else self.d_sba)
        if world > 1:   # every rank scans its own slice of the byte array; the three counters are summed
            tot = eng.from_host_i64(counts.astype(np.int64))
            dist.all_reduce(tot, group=self.group)
            counts = self._to_host_i64(tot).astype(np.uint64)
        n_sep_expected = len(self.seg_starts) - 1
        if int(counts[1]) != n_sep_expected:
            raise AssertionError("kmers compared were less than min_kmer_len: '$' inside a record")
        class_bit = 1 if (counts[2] > 0 or counts[0] > 0) else 0
        self._mark("alphabet")
        keys, idx = eng.pack_slice(self.d_sba, self.seg_starts, k, class_bit, first, end, self.idx_bytes)
        n_local_in = int(keys.numel())
        self._mark("pack")

        # ---- splitters from evenly spaced samples --------------------------------------------------
        if world > 1:
            step = max(1, n_local_in // SAMPLES_PER_RANK)
            sample = keys[::step][:SAMPLES_PER_RANK]
            padded = eng.empty_like_n(keys, SAMPLES_PER_RANK + 1)
            padded[0] = int(sample.numel())
            padded[1:1 + sample.numel()] = sample
            if sample.numel() < SAMPLES_PER_RANK:
                padded[1 + sample.numel():] = 0
            gathered = [eng.empty_like_n(keys, SAMPLES_PER_RANK + 1) for _ in range(world)]
            dist.all_gather(gathered, padded, group=self.group)
            # one device-to-host copy of all samples; 32 k keys are sorted faster on the host than through
            # eight tiny radix passes and their synchronisations
            table = self._to_host_u64(self._cat(gathered)).reshape(world, SAMPLES_PER_RANK + 1)
            pooled = np.sort(np.concatenate([row[1:1 + int(row[0])] for row in table]))
            splitters_host = choose_splitters(pooled, world, class_bit)
            splitters = eng.from_host_i64(splitters_host.view(np.int64))
        else:
            splitters_host, splitters = np.zeros(0, dtype=np.uint64), None
        self.splitters = splitters_host
        self._mark("splitters")

        # ---- partition by destination, exchange -----------------------------------------------------
        use_peer = (world > 1 and hasattr(eng, "peer_exchange")
                    and os.environ.get("GK_PEER_EXCHANGE", "1") != "0")
        recv_ptrs = None
        if use_peer:
            # fused: counts first (placement), then ONE kernel partitions and writes every pair straight
            # into its destination rank's receive buffer over NVLink peer memory
            send_counts = eng.partition_count(keys, splitters, world)
            mine = eng.from_host_i64(send_counts)
            rows = [eng.empty_like_n(mine, world) for _ in range(world)]
            dist.all_gather(rows, mine, group=self.group)
            matrix = np.stack([self._to_host_i64(r) for r in rows])          # [source, destination]
            recv_total = matrix.sum(axis=0)
            # every rank sees the same matrix, so every rank computes the same capacity: the largest
            # receive count plus 10 % head-room (the buffers are cached and only ever grow)
            capacity = int(1.1 * int(recv_total.max())) + (1 << 20)
            self._mark("partition")
            if capacity * (8 + self.idx_bytes) <= PeerExchange.MAX_BYTES:
                px = eng.peer_exchange(dist, self.group, capacity, self.idx_bytes)
                offsets = matrix[:rank, :].sum(axis=0)
                eng.partition_peer(keys, idx, splitters, world, px, offsets)
                token = eng.from_host_i64(np.zeros(1, dtype=np.int64))
                dist.all_reduce(token, group=self.group)     # every rank's writes have landed behind this
                recv_ptrs = (px.my_keys, px.my_idx, int(recv_total[rank]))
                self.exchange_bytes_sent = int((send_counts.sum() - send_counts[rank]) * (8 + self.idx_bytes))
                del keys, idx
                self._mark("exchange")
            else:
                use_peer = False                             # skew beyond the buffers: NCCL path below
        if not use_peer:
            keys_p, idx_p, send_counts = eng.partition(keys, idx, splitters, world)
            del keys, idx
            self._mark("partition")
            if world > 1:
                send_t = eng.from_host_i64(send_counts)
                recv_t = eng.empty_like_n(send_t, world)
                dist.all_to_all_single(recv_t, send_t, group=self.group)
                recv_counts = self._to_host_i64(recv_t)
                n_recv = int(recv_counts.sum())
                keys_r = eng.empty_like_n(keys_p, n_recv)
                idx_r = eng.empty_like_n(idx_p, n_recv)
                in_splits, out_splits = [int(c) for c in send_counts], [int(c) for c in recv_counts]
                dist.all_to_all_single(keys_r, keys_p, output_split_sizes=out_splits, input_split_sizes=in_splits,
                                       group=self.group)
                dist.all_to_all_single(idx_r, idx_p, output_split_sizes=out_splits, input_split_sizes=in_splits,
                                       group=self.group)
                self.exchange_bytes_sent = int((send_counts.sum() - send_counts[rank])
                                               * (8 + idx_p.element_size()))
            else:
                keys_r, idx_r = keys_p, idx_p
                self.exchange_bytes_sent = 0
            del keys_p, idx_p
            self._mark("exchange")

        # ---- local sort + refinement + flags ------------------------------------------------------------
        if self.shard is not None:
            eng.shard_free(self.shard)
        if recv_ptrs is not None:
            self.shard = eng.shard_index_ptr(self.d_sba, self.seg_starts, k, recv_ptrs[0], recv_ptrs[1],
                                             recv_ptrs[2], self.idx_bytes, class_bit)
        else:
            self.shard = eng.shard_index(self.d_sba, self.seg_starts, k, keys_r, idx_r, class_bit)
        self.exchange_mode = "peer" if recv_ptrs is not None else "nccl"
        self.stats = dict(self.shard["stats"])
        self.stats.update(n_packed=n_local_in, n_shard=int(self.shard["n"]), class_bit=class_bit)
        self._mark("local_sort")
        self._is_sorted = True

    # -------------------------------------------------------------------------------------------
    def get_kmer_group_counts(self, kmer_len: Optional[int] = None, filt=None, min_group_size: int = 1,
                              max_group_size: Optional[int] = None, max_counts_bin: int = 1000000):
        """Global (counts_by_group_size, total): every rank returns the same answer."""
        if not self._is_sorted:
            raise AssertionError("The kmers must be sorted when calling get_kmer_group_counts")
        kmer_len = self.k if kmer_len is None else kmer_len
        if kmer_len != self.k:
            raise NotImplementedError("the multi-GPU path counts groups for the sort length only")
        sparse = getattr(self.engine, "shard_counts_sparse", None)
        if sparse is not None:
            bins, counts, total = sparse(self.shard, kmer_len, filt, min_group_size, max_group_size,
                                         max_counts_bin)
        else:
            dense, total = self.engine.shard_counts(self.shard, kmer_len, filt, min_group_size, max_group_size,
                                                    max_counts_bin)
            bins = np.flatnonzero(dense).astype(np.uint64)
            counts = dense[bins.astype(np.int64)]
        hist = np.zeros(max_counts_bin + 1, dtype=np.int64)
        if self.world == 1:
            hist[bins.astype(np.int64)] = counts
            return hist, total
        # the occupied bins of every rank are all-gathered as (bin, count) pairs: a few hundred bytes,
        # instead of all-reducing the 8 MB table the reference's default max_counts_bin implies
        head = self.engine.from_host_i64(np.array([len(bins), total], dtype=np.int64))
        heads = [self.engine.empty_like_n(head, 2) for _ in range(self.world)]
        self.dist.all_gather(heads, head, group=self.group)
        heads = np.stack([self._to_host_i64(h) for h in heads])
        width = int(heads[:, 0].max())
        total = int(heads[:, 1].sum())
        if width:
            mine = np.zeros(2 * width, dtype=np.int64)
            mine[:len(bins)] = bins.astype(np.int64)
            mine[width:width + len(bins)] = counts
            mine_t = self.engine.from_host_i64(mine)
            parts = [self.engine.empty_like_n(mine_t, 2 * width) for _ in range(self.world)]
            self.dist.all_gather(parts, mine_t, group=self.group)
            for r, part in enumerate(parts):
                arr, m = self._to_host_i64(part), int(heads[r, 0])
                np.add.at(hist, arr[:m], arr[width:width + m])
        return hist, total

    def local_start_indices(self) -> np.ndarray:
        """This rank's shard of the globally sorted start indices (host array)."""
        return self.engine.shard_indices_host(self.shard)

    def gather_start_indices(self, dst: int = 0):
        """The whole sorted index on rank `dst` (None elsewhere)."""
        local = self.local_start_indices()
        if self.world == 1:
            return local
        gathered = [None] * self.world if self.rank == dst else None
        self.dist.gather_object(local, gathered, dst=dst, group=self.group)
        return np.concatenate(gathered) if self.rank == dst else None

    def close(self):
        if self.shard is not None:
            self.engine.shard_free(self.shard)
            self.shard = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- small helpers that work for CUDA tensors and for the NumPy test engine's CPU tensors ---------
    def _cat(self, parts):
        return _torch().cat(list(parts)) if parts else parts

    @staticmethod
    def _to_host_i64(t) -> np.ndarray:
        return t.detach().cpu().numpy().astype(np.int64)

    @staticmethod
    def _to_host_u64(t) -> np.ndarray:
        return t.detach().cpu().numpy().view(np.uint64)


# ---------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): weak scaling, 100 Mbp of genome per GPU, one index over all of it
# ---------------------------------------------------------------------------------------------------
def bench_main(args, rank, world, make_genome, workload_config, ClockSampler, metric, unit, k, n_records,
               runs_per_record, max_bin, measured_hbm_peak):
    import json
    import os
    import time

    import torch
    import torch.distributed as dist

    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = NativeEngine()
    n_bases = args.bases

    # every rank generates its own 100 Mbp (10 records) and the ranks all-gather the forward byte array;
    # the trailing '$' of a rank's chunk separates it from the next rank's first record
    chunk_len = n_bases + n_records
    pinned = torch.empty(chunk_len, dtype=torch.uint8).pin_memory()
    host = pinned.numpy()
    sba, starts, _ = make_genome(n_bases, n_records, runs_per_record, 42 + rank, out=host[:chunk_len - 1])
    host[chunk_len - 1] = ord("$")
    all_starts = np.concatenate([starts + np.uint64(r * chunk_len) for r in range(world)])
    total_fwd = world * chunk_len - 1
    n_total = 2 * (n_bases * world - n_records * world * (k - 1))

    def load_inputs():
        d_chunk = pinned.to("cuda", non_blocking=True)
        d_all = torch.empty(world * chunk_len, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(d_all, d_chunk)
        return d_all[:total_fwd]

    d_fwd = load_inputs()
    hist = None

    def step(d_forward):
        m0 = eng.mark()
        sk = ShardedKmers(d_forward, all_starts, k, "both", engine=eng)
        m1 = eng.mark()
        sk.sort()
        m2 = eng.mark()
        h, total = sk.get_kmer_group_counts(k, max_counts_bin=max_bin)
        assert total == n_total, (total, n_total)
        stats, sent = dict(sk.stats), sk.exchange_bytes_sent
        stats["_mode"] = sk.exchange_mode
        sk.close()
        m3 = eng.mark()
        stats["_sk_marks"] = [("begin", m0), ("both_strands", m1)] + sk._marks[1:] + [("count_allreduce", m3)]
        del m2
        return h, stats, sent

    for _ in range(args.warmup):
        step(d_fwd)
    launches0 = _native.launch_count(reset=True)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    stream = torch.cuda.current_stream()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    per_step = []
    for _ in range(args.steps):
        hist, stats, sent = step(d_fwd)
        if rank == 0:
            clocks.sample()
        per_step.append((stats, sent))
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = _native.launch_count()
    ms_per_step = float(ms.item()) / args.steps
    clock_info = clocks.stop() if rank == 0 else None

    # e2e: host chunk -> H2D -> all-gather -> sort/count -> shard of sorted starts back on the host
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        shard_bytes = 0

        def e2e_step():
            d_in = load_inputs()
            sk = ShardedKmers(d_in, all_starts, k, "both", engine=eng)
            sk.sort()
            h, total = sk.get_kmer_group_counts(k, max_counts_bin=max_bin)
            local = sk.local_start_indices()     # D2H of this rank's shard of the sorted starts (pinned)
            sk.close()
            assert total == n_total
            return int(local.nbytes)

        for _ in range(2):                       # warm the pinned-buffer cache and the NCCL channels
            e2e_step()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            shard_bytes = e2e_step()
        torch.cuda.synchronize()
        dist.barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": n_total / float(dt.item()) / 1e9, "unit": unit, "ms_per_step": 1e3 * float(dt.item()),
               "steps": e2e_steps, "h2d_bytes_per_step": int(chunk_len) * world,
               "d2h_bytes_per_step": shard_bytes * world,
               "api": "ShardedKmers(...).sort(); get_kmer_group_counts(); local_start_indices() on every rank"}

    eng_idx_bytes = 8 if (2 * (world * chunk_len - 1) + 1 > 0xFFFFFFFF
                          or os.environ.get("GK_FORCE_IDX64", "0") not in ("", "0")) else 4
    last = per_step[-1][0]
    mine = [last.get("total_ms", 0.0), last.get("fixup_ms", 0.0), float(last.get("n_shard", 0)),
            float(last.get("n_ambiguous", 0))]
    per_rank = [None] * world
    dist.all_gather_object(per_rank, mine)
    sent_all = torch.tensor([float(np.mean([s for _, s in per_step]))], dtype=torch.float64, device="cuda")
    dist.all_reduce(sent_all, op=dist.ReduceOp.SUM)
    phase_ms = {}
    exchange_mode = per_step[-1][0].pop("_mode", "nccl")
    for stats, _ in per_step:
        stats.pop("_mode", None)
        marks = stats.pop("_sk_marks", [])
        for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
            phase_ms[name] = phase_ms.get(name, 0.0) + a.elapsed_time(b) / len(per_step)
    if rank == 0:
        passes = per_step[-1][0]["sort_passes"]
        pass_ms = float(np.mean([s["sort_ms"] for s, _ in per_step])) / max(passes, 1)
        n_shard = per_step[-1][0]["n_shard"]
        peak, peak_src = measured_hbm_peak()
        pair_bytes = 8 + eng_idx_bytes
        achieved = 2 * pair_bytes * n_shard / (pass_ms * 1e-3) / 1e9
        line = {
            "metric": metric, "value": n_total / (ms_per_step * 1e-3) / 1e9, "unit": unit, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": dict(workload_config(world), **({} if n_bases == 100_000_000 else {
                "workload": f"NOT the bench workload: {n_bases} bp per GPU x {world} GPUs = "
                            f"{n_bases * world / 1e9:.2f} Gbp, {n_records} records per GPU, N runs, both strands, "
                            f"k={k} (BASELINE.json configs[2] size when bases x GPUs = 3.1e9)",
                "bases_per_gpu": n_bases, "kmers_total": int(n_total)})),
            "e2e": e2e, "gpu_launches": int(launches - 0),
            "roofline": {"bound": "hbm", "kernel": "gk::onesweep_kernel on rank 0's key range",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "avg_launch_ms": pass_ms,
                         "launches_per_step": passes, "pairs_on_rank0": int(n_shard),
                         "pair_bytes": pair_bytes},
            "exchange": {"bytes_over_nvlink_per_step": float(sent_all.item()), "mode": exchange_mode,
                         # rank 0's exchange phase: fused partition + peer writes + the ordering all-reduce
                         "exchange_ms_rank0": round(float(phase_ms.get("exchange", 0.0)), 3),
                         "nvlink_gbs_per_gpu_outbound": (
                             round(float(sent_all.item()) / world / (phase_ms["exchange"] * 1e-3) / 1e9, 1)
                             if phase_ms.get("exchange", 0.0) > 0 else None),
                         "nvlink_peak_gbs_per_direction": 900.0,
                         "note": "(G-1)/G of all (u64 key, u32 start) pairs cross NVLink once: written by the "
                                 "partition kernel into peer memory (mode peer) or one NCCL all-to-all (mode nccl)"},
            "cpu_baseline": None, "clocks": clock_info,
            "phase_ms_rank0": {k_: round(v, 3) for k_, v in phase_ms.items()},
            "per_rank_local_sort": {"total_ms": [round(r[0], 3) for r in per_rank],
                                    "refine_ms": [round(r[1], 3) for r in per_rank],
                                    "pairs": [int(r[2]) for r in per_rank],
                                    "ambiguous": [int(r[3]) for r in per_rank]},
            "local_sort_stats_rank0": {k_: v for k_, v in per_step[-1][0].items()},
            "result": {"kmers": int(n_total), "distinct_kmers": int(hist.sum())},
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    PeerExchange.close_all()
    dist.destroy_process_group()
