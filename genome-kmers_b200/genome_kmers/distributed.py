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
HIST_PAIRS_PER_MESSAGE = 512   # occupied histogram bins per rank in the first all-gather round


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

    def shard_verify(self, sk, k):
        """(report8, first k-mer bytes, last k-mer bytes, bits set in the summed start bitmaps)."""
        torch, lib = self.torch, self.lib
        shard = sk.shard
        words = sk.total_len // 32 + 1
        seen = torch.zeros(words, dtype=torch.int32, device=self.device)
        report = np.zeros(8, dtype=np.uint64)
        _native.check(lib.gk_index_verify(shard["handle"], k, _native.host_ptr(report), seen.data_ptr(),
                                          self.stream()))
        if sk.world > 1:
            sk.dist.all_reduce(seen, group=sk.group)     # disjoint bitmaps add without carries
        bits = ctypes.c_uint64(0)
        _native.check(lib.gk_popcount_words(seen.data_ptr(), words, ctypes.byref(bits), self.stream()))
        first = last = bytes(k)
        if shard["n"]:
            ptr = ctypes.c_void_p()
            _native.check(lib.gk_index_device_indices(shard["handle"], ctypes.byref(ptr), self.stream()))
            from genome_kmers.kmers import _tensor_from_ptr

            idx = _tensor_from_ptr(torch, ptr.value, shard["n"], shard["idx_bytes"])
            ends = idx[[0, shard["n"] - 1]].to(torch.int64).cpu().numpy()
            first = sk.d_sba[int(ends[0]):int(ends[0]) + k].cpu().numpy().tobytes()
            last = sk.d_sba[int(ends[1]):int(ends[1]) + k].cpu().numpy().tobytes()
        return [int(v) for v in report], first, last, int(bits.value)

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


def backend_is_nccl(dist, group) -> bool:
    try:
        return str(dist.get_backend(group)).lower() == "nccl"
    except Exception:
        return False


def gather_small(dist, group, engine, arr: np.ndarray) -> np.ndarray:
    """All-gather of a small host array (same shape and dtype on every rank) -> [world, ...] on the host.
    NCCL moves it through device tensors; any other backend (gloo in the CPU tests and in the
    two-ranks-on-one-GPU test) through CPU tensors."""
    torch = _torch()
    arr = np.ascontiguousarray(arr)
    world = dist.get_world_size(group)
    flat = torch.from_numpy(arr.reshape(-1).view(np.uint8).copy())
    if backend_is_nccl(dist, group):
        flat = flat.to(engine.device)
    out = torch.empty(world * flat.numel(), dtype=torch.uint8, device=flat.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    host = out.cpu().numpy().view(arr.dtype)
    return host.reshape((world,) + arr.shape)


class PeerExchange:
    """Receive buffers of every rank, mapped into every rank (CUDA IPC over NVLink peer memory).

    The fused partition kernel (gk_partition_pairs_peer) writes each (key, start) pair directly into the
    buffer of the rank that owns its key range.  Buffers are allocated once per (group, capacity) and
    reused by later sorts; `capacity` counts pairs.  The mapping is only attempted when every rank sits on
    one host and the group has at most 16 ranks (the kernel's destination table); get() returns None when
    any rank fails to map a peer's buffers, and every rank then takes the NCCL all-to-all path."""

    _cache = {}
    MAX_BYTES = 64 << 30   # per rank; beyond this the NCCL all-to-all path is used
    MAX_RANKS = 16         # kMaxPeers in csrc/gk_sort.cu

    def __init__(self, engine, dist, group, capacity: int, idx_bytes: int):
        lib = engine.lib
        self.lib, self.capacity, self.idx_bytes = lib, int(capacity), idx_bytes
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.my_keys, self.my_idx = ctypes.c_void_p(), ctypes.c_void_p()
        self._opened = []
        self.ok = False
        ok = 1
        handles = np.zeros(128, dtype=np.uint8)
        try:
            _native.check(lib.gk_peer_alloc(self.capacity * 8, ctypes.byref(self.my_keys)))
            _native.check(lib.gk_peer_alloc(self.capacity * idx_bytes, ctypes.byref(self.my_idx)))
            _native.check(lib.gk_peer_export(self.my_keys, _native.host_ptr(handles)))
            _native.check(lib.gk_peer_export(self.my_idx, _native.host_ptr(handles[64:])))
        except Exception:
            ok = 0
        gathered = gather_small(dist, group, engine, handles)
        self.key_ptrs = np.zeros(self.world, dtype=np.uint64)
        self.idx_ptrs = np.zeros(self.world, dtype=np.uint64)
        for r in range(self.world):
            if not ok:
                break
            if r == self.rank:
                self.key_ptrs[r], self.idx_ptrs[r] = self.my_keys.value, self.my_idx.value
                continue
            h = np.ascontiguousarray(gathered[r])
            pk, pi = ctypes.c_void_p(), ctypes.c_void_p()
            try:
                _native.check(lib.gk_peer_open(_native.host_ptr(h), ctypes.byref(pk)))
                self._opened.append(pk)
                _native.check(lib.gk_peer_open(_native.host_ptr(h[64:]), ctypes.byref(pi)))
                self._opened.append(pi)
            except Exception:
                ok = 0
                break
            self.key_ptrs[r], self.idx_ptrs[r] = pk.value, pi.value
        # every rank must be able to reach every other one, or nobody uses the mapping
        self.ok = bool(gather_small(dist, group, engine, np.array([ok], dtype=np.int64)).min())

    @staticmethod
    def available(engine, dist, group) -> bool:
        """One host, at most MAX_RANKS ranks (agreed by all ranks)."""
        import socket
        import zlib

        world = dist.get_world_size(group)
        if world > PeerExchange.MAX_RANKS:
            return False
        host = np.array([zlib.crc32(socket.gethostname().encode())], dtype=np.int64)
        hosts = gather_small(dist, group, engine, host)
        return bool((hosts == hosts[0]).all())

    @classmethod
    def get(cls, engine, dist, group, capacity: int, idx_bytes: int):
        key = (id(group), dist.get_world_size(group), idx_bytes)
        if key in cls._cache and cls._cache[key] is None:
            return None                                # the mapping failed before: stay on the NCCL path
        px = cls._cache.get(key)
        if px is None or px.capacity < capacity:   # every rank computes the same capacity: collective-safe
            if px is not None:
                px.close()
            elif not cls.available(engine, dist, group):
                cls._cache[key] = None
                return None
            px = cls(engine, dist, group, capacity, idx_bytes)
            if not px.ok:
                px.close()
                px = None
            cls._cache[key] = px
        return px

    def close(self):
        """Collective: unmap every peer's buffers, wait until every rank has done so, then free the local ones
        (freeing exported memory that another process still has open is undefined in CUDA IPC)."""
        _torch().cuda.synchronize()
        for p in self._opened:
            self.lib.gk_peer_close(p)
        self._opened = []
        self.dist.barrier(group=self.group)
        if self.my_keys:
            self.lib.gk_peer_free(self.my_keys)
            self.my_keys = None
        if self.my_idx:
            self.lib.gk_peer_free(self.my_idx)
            self.my_idx = None

    @classmethod
    def close_all(cls):
        for px in cls._cache.values():
            if px is not None:
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

    # ---- small collectives through the host (see gather_small) -------------------------------------------
    def _gather(self, arr) -> np.ndarray:
        return gather_small(self.dist, self.group, self.engine, np.asarray(arr))

    def _order_after_peer_writes(self):
        """Every rank's peer writes have landed behind this point of the current stream."""
        if backend_is_nccl(self.dist, self.group):
            token = self.engine.from_host_i64(np.zeros(1, dtype=np.int64))
            self.dist.all_reduce(token, group=self.group)      # stream-ordered: no host synchronise
        else:
            _torch().cuda.synchronize()
            self.dist.barrier(group=self.group)

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
            counts = self._gather(counts.astype(np.int64)).sum(axis=0).astype(np.uint64)
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
            sample = self._to_host_u64(keys[::step][:SAMPLES_PER_RANK])
            padded = np.zeros(SAMPLES_PER_RANK + 1, dtype=np.uint64)
            padded[0] = len(sample)
            padded[1:1 + len(sample)] = sample
            # one small all-gather; 32 k keys are sorted faster on the host than through eight tiny radix
            # passes and their synchronisations
            table = self._gather(padded)
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
            matrix = self._gather(send_counts.astype(np.int64))             # [source, destination]
            recv_total = matrix.sum(axis=0)
            # every rank sees the same matrix, so every rank computes the same capacity: the largest
            # receive count plus 10 % head-room (the buffers are cached and only ever grow)
            capacity = int(1.1 * int(recv_total.max())) + (1 << 20)
            self._mark("partition")
            px = None
            if capacity * (8 + self.idx_bytes) <= PeerExchange.MAX_BYTES:
                px = eng.peer_exchange(dist, self.group, capacity, self.idx_bytes)
            if px is not None:
                offsets = matrix[:rank, :].sum(axis=0)
                eng.partition_peer(keys, idx, splitters, world, px, offsets)
                self._order_after_peer_writes()
                recv_ptrs = (px.my_keys, px.my_idx, int(recv_total[rank]))
                self.exchange_bytes_sent = int((send_counts.sum() - send_counts[rank]) * (8 + self.idx_bytes))
                del keys, idx
                self._mark("exchange")
            else:
                use_peer = False         # skew beyond the buffers, several hosts, or no peer access: NCCL path
        if not use_peer:
            keys_p, idx_p, send_counts = eng.partition(keys, idx, splitters, world)
            del keys, idx
            self._mark("partition")
            if world > 1:
                recv_counts = self._gather(send_counts.astype(np.int64))[:, rank]
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
        # the occupied bins of every rank are all-gathered as (bin, count) pairs in ONE fixed-size message
        # (a few KB), instead of all-reducing the 8 MB table the reference's default max_counts_bin implies;
        # a rank with more occupied bins than the message holds triggers a second, wider round
        width = HIST_PAIRS_PER_MESSAGE
        while True:
            mine = np.zeros(2 + 2 * width, dtype=np.int64)
            mine[0], mine[1] = len(bins), total
            m = min(len(bins), width)
            mine[2:2 + m] = bins[:m].astype(np.int64)
            mine[2 + width:2 + width + m] = counts[:m]
            table = self._gather(mine)
            widest = int(table[:, 0].max())
            if widest <= width:
                break
            width = widest
        for row in table:
            m = int(row[0])
            np.add.at(hist, row[2:2 + m], row[2 + width:2 + width + m])
        return hist, int(table[:, 1].sum())

    def verify(self, hist=None, n_total: Optional[int] = None) -> dict:
        """Check the sharded order against the sequence bytes (collective; gk_index_verify per shard):
        every shard sorted with ties in ascending start order and valid, distinct starts; shard boundaries in
        strictly increasing k-mer order; all shards together hold every window start exactly once (their start
        bitmaps add up without a carry to n_total bits); the global histogram counts the groups the bytes show.
        Returns {"checked": True, "checks": {name: bool}, "report": ...}; every rank gets the same answer."""
        eng, k = self.engine, self.k
        rep, first_kmer, last_kmer, bits_total = eng.shard_verify(self, k)
        reports = self._gather(np.array([rep[i] for i in range(8)], dtype=np.int64))
        edges = self._gather(np.frombuffer(first_kmer + last_kmer, dtype=np.uint8))   # [world, 2k]
        n_shards = reports[:, 0]
        total = int(n_shards.sum())
        boundaries_ok = True
        prev_last = None
        for r in range(self.world):
            if n_shards[r] == 0:
                continue
            f, l = edges[r, :k].tobytes(), edges[r, k:].tobytes()
            if prev_last is not None and not prev_last < f:
                boundaries_ok = False                 # a group would straddle two ranks, or order is broken
            prev_last = l
        checks = {
            "every shard: neighbours in non-decreasing order under the reference's byte comparator":
                int(reports[:, 1].sum()) == 0,
            "every shard: equal k-mers in ascending start order (break_ties=True order)": int(reports[:, 2].sum()) == 0,
            "every shard: starts valid and distinct": int(reports[:, 3].sum()) == 0 and int(reports[:, 4].sum()) == 0,
            "every shard: head flags of the sort agree with the bytes":
                int(reports[:, 6].sum()) == 0 and bool((reports[n_shards > 0, 7] == 1).all()),
            "shard boundaries: last k-mer of a rank < first k-mer of the next (no group straddles ranks)":
                boundaries_ok,
            "all shards together hold every window start exactly once (summed start bitmaps)":
                bits_total == total,
        }
        if n_total is not None:
            checks["number of k-mers over all shards"] = total == int(n_total)
        if hist is not None:
            checks["histogram: number of groups equals the groups counted from the bytes"] = \
                int(np.asarray(hist).sum()) == int(reports[n_shards > 0, 5].sum())
        return {"checked": True, "checks": checks,
                "report": {"kmers_per_rank": [int(v) for v in n_shards], "groups": int(reports[n_shards > 0, 5].sum()),
                           "start_bits_set": int(bits_total)}}

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
