"""
Multi-GPU driver for the hot path: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch) for the plumbing, libgkb200 kernels for every compute step.

The reference has no parallel code at all; this is new design.  The k-mer windows of the indexed byte
array are range-partitioned by key so that equal k-mers always meet on one GPU:

  1. every rank holds the whole byte array (a few GB at most) and packs the windows of ITS slice of start
     positions into (key, start) pairs; the ambiguous windows of the slice (N runs, IUPAC letters) are also
     listed as run-length fragments                                                gk_pack_slice
  2. meanwhile (second stream) keys of evenly spaced windows of the slice are sampled, all-gathered, and
     every rank picks the same G-1 splitters                                       gk_sample_keys
  3. destination counts (pure / ambiguous pairs per rank) are exchanged in ONE small all-gather together
     with the fragment counts and the alphabet counters                            gk_partition_count_split
  4. ONE kernel partitions the pure pairs and writes each straight into the receive buffer of the rank that
     owns its key range, over NVLink peer memory, with the key made relative to the start of that range;
     ambiguous pairs are not sent at all: the fragment lists are all-gathered (a few MB) while that kernel
     runs, and each rank regenerates the ambiguous windows of its own key range    gk_partition_pairs_peer
  5. every rank sorts its key range, repairs prefix buckets, expands the fragments gk_index_sort_shard
  6. the occupied histogram bins are all-gathered; the global sorted order is the concatenation of the
     ranks' shards in rank order.
Splitters are keys (not (key, start) pairs), so a group of equal k-mers never straddles two ranks and
no boundary fix-up is needed; the price is that one giant group cannot be split (documented skew).
Ties stay in ascending start order: source ranks hold ascending slices, the partition and the sort are
stable, and segments are laid out by source rank.
Without peer access (several hosts, more than 16 ranks) step 4 partitions into a local staging buffer and
the pairs travel through one NCCL all-to-all per array.

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
FRAG_SHARE = 1 << 16           # fragments per rank that the fixed-size fragment all-gather carries
FRAG_BYTES = 36                # key, w0, w1, start (u64) + count (u32)
# weight of an ambiguous window against a pure one when the splitters balance the ranks: its pair does not cross
# NVLink, but the owner generates it, carries it through the passes and rewrites its slot (measured at 8 GPUs:
# the owner of the 70 M-window all-N group took 0.7 ms longer than the others at weight 1)
AMBIGUOUS_COST = float(os.environ.get("GK_AMBIGUOUS_COST", "1.3"))


def _torch():
    import torch

    return torch


class PackedSlice:
    """What pack_slice leaves on the device: the (key, start) pairs of a rank's slice, its fragment list and
    the counters [ambiguous windows, -, fragments, -]."""

    def __init__(self, keys, idx, n, frag=None, counters=None):
        self.keys, self.idx, self.n, self.frag, self.counters = keys, idx, int(n), frag, counters


class NativeEngine:
    """Compute steps on the current CUDA device through the C ABI."""

    supports_fragments = True

    def __init__(self):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("genome_kmers.distributed needs a CUDA device per rank (no CPU fallback)")
        self.torch = torch
        self.lib = _native.lib()
        self.device = torch.device("cuda", torch.cuda.current_device())
        prio = 0 if os.environ.get("GK_SIDE_PRIORITY", "1") == "0" else -1
        self.side = torch.cuda.Stream(device=self.device, priority=prio)   # ahead of the bulk kernels
        self.err_word = torch.zeros(1, dtype=torch.int32, device=self.device)

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

    def both_strands_sliced(self, d_fwd, begin, end):
        """The same layout, bytes [begin, end) on the caller's stream now.  Returns (tensor, finish): finish()
        enqueues the rest on a third stream behind whatever the caller's stream holds at that moment (the pack
        kernel: the rest then fills the gaps of the host-driven phases) and returns the event to wait for."""
        torch = self.torch
        n = d_fwd.numel()
        out = torch.empty(2 * n + 1, dtype=torch.uint8, device=self.device)
        _native.check(self.lib.gk_sba_both_strands_range(d_fwd.data_ptr(), n, out.data_ptr(), begin, end,
                                                         self.stream()))
        if getattr(self, "bg", None) is None:
            self.bg = torch.cuda.Stream(device=self.device)

        def finish():
            self.bg.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.bg):
                for b, e in ((0, begin), (end, 2 * n + 1)):
                    if e > b:
                        _native.check(self.lib.gk_sba_both_strands_range(d_fwd.data_ptr(), n, out.data_ptr(), b, e,
                                                                         self.stream()))
                ev = torch.cuda.Event()
                ev.record(self.bg)
            d_fwd.record_stream(self.bg)
            out.record_stream(self.bg)
            return ev

        return out, finish

    def alphabet_async(self, d_sba):
        """Device tensor int64[3] (bad bytes, '$', ambiguous letters); no synchronise."""
        counts = self.torch.empty(3, dtype=self.torch.int64, device=self.device)
        _native.check(self.lib.gk_sba_scan_alphabet_async(d_sba.data_ptr(), d_sba.numel(), counts.data_ptr(),
                                                          self.stream()))
        return counts

    def pack_slice(self, d_sba, seg_starts, k, class_bit, first, end, idx_bytes):
        torch = self.torch
        cap = max(1, end - first)
        keys = torch.empty(cap, dtype=torch.int64, device=self.device)
        idx = torch.empty(cap, dtype=torch.int32 if idx_bytes == 4 else torch.int64, device=self.device)
        frag = torch.empty(FRAG_SHARE * FRAG_BYTES, dtype=torch.uint8, device=self.device)
        counters = torch.empty(4, dtype=torch.int64, device=self.device)
        n_out = ctypes.c_uint64(0)
        _native.check(self.lib.gk_pack_slice(d_sba.data_ptr(), d_sba.numel(), _native.host_ptr(seg_starts),
                                             len(seg_starts), k, class_bit, first, end, keys.data_ptr(), idx_bytes,
                                             idx.data_ptr(), cap, ctypes.byref(n_out), frag.data_ptr(), FRAG_SHARE,
                                             counters.data_ptr(), self.stream()))
        return PackedSlice(keys[:n_out.value], idx[:n_out.value], n_out.value, frag, counters)

    def sample_keys_host(self, d_sba, seg_starts, k, class_bit, first, end, n_samples, after=None):
        """Keys of evenly spaced windows of the slice, on the host.  Runs on the engine's second stream so that
        the caller's stream (busy with the pack kernel) is not waited for."""
        torch = self.torch
        out = torch.empty(max(1, n_samples), dtype=torch.int64, device=self.device)
        n_out = ctypes.c_uint32(0)
        if after is not None:
            self.side.wait_event(after)      # d_sba was produced on the caller's stream, before this event
        else:
            self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            _native.check(self.lib.gk_sample_keys(d_sba.data_ptr(), d_sba.numel(), _native.host_ptr(seg_starts),
                                                  len(seg_starts), k, class_bit, first, end, n_samples,
                                                  out.data_ptr(), ctypes.byref(n_out), self.stream()))
            host = out[:n_out.value].cpu()
        return host.numpy().view(np.uint64)

    def side_context(self):
        """Work issued inside runs on the second stream (small collectives that must not queue behind the bulk
        kernel the caller's stream is busy with)."""
        return self.torch.cuda.stream(self.side)

    def splitters_to_device(self, splitters_host):
        return self.torch.from_numpy(np.ascontiguousarray(splitters_host).view(np.int64).copy()).to(self.device)

    def partition_counts_host(self, pk, splitters, n_parts, class_bit, extra=()):
        """[pure pairs per destination, ambiguous pairs per destination, fragments, ambiguous windows, *extra]
        as int64 on the host: one device-to-host copy behind the pack kernel."""
        torch = self.torch
        counts = torch.empty(2 * n_parts, dtype=torch.int64, device=self.device)
        sp = splitters.data_ptr() if splitters is not None and splitters.numel() else None
        _native.check(self.lib.gk_partition_count_split(pk.keys.data_ptr(), pk.n, sp, n_parts, class_bit,
                                                        counts.data_ptr(), self.stream()))
        parts = [counts, pk.counters[2:3], pk.counters[0:1]] + [e for e in extra]
        return torch.cat(parts).cpu().numpy()

    def partition_counts_gathered(self, pk, splitters, n_parts, class_bit, dist, group, extra=()):
        """The same header from every rank, [world, ...] on the host: under NCCL the header is gathered on the
        device and crosses to the host once (no round trip in between)."""
        torch = self.torch
        counts = torch.empty(2 * n_parts, dtype=torch.int64, device=self.device)
        sp = splitters.data_ptr() if splitters is not None and splitters.numel() else None
        _native.check(self.lib.gk_partition_count_split(pk.keys.data_ptr(), pk.n, sp, n_parts, class_bit,
                                                        counts.data_ptr(), self.stream()))
        mine = torch.cat([counts, pk.counters[2:3], pk.counters[0:1]] + [e for e in extra])
        world = dist.get_world_size(group)
        out = torch.empty(world * mine.numel(), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(out, mine, group=group)
        return out.cpu().numpy().reshape(world, mine.numel())

    def peer_exchange(self, dist, group, capacity, idx_bytes):
        return PeerExchange.get(self, dist, group, capacity, idx_bytes)

    def _partition(self, pk, splitters, n_parts, key_ptrs, idx_ptrs, offsets, key_base, skip_amb):
        sp = splitters.data_ptr() if splitters is not None and splitters.numel() else None
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        base = np.ascontiguousarray(key_base, dtype=np.uint64)
        _native.check(self.lib.gk_partition_pairs_peer(
            pk.keys.data_ptr(), pk.idx.data_ptr(), pk.idx.element_size(), pk.n, sp, n_parts,
            _native.host_ptr(key_ptrs), _native.host_ptr(idx_ptrs), _native.host_ptr(off), _native.host_ptr(base),
            int(bool(skip_amb)), self.err_word.data_ptr(), self.stream()))

    def partition_to_peers(self, pk, splitters, n_parts, px, offsets, key_base, skip_amb):
        """One kernel: every pair straight into its destination rank's receive buffer (peer memory)."""
        self.err_word.zero_()
        self._partition(pk, splitters, n_parts, px.key_ptrs, px.idx_ptrs, offsets, key_base, skip_amb)

    def partition_to_staging(self, pk, splitters, n_parts, send_counts, key_base, skip_amb):
        """The same kernel into a local staging buffer laid out by destination (for the all-to-all path)."""
        torch = self.torch
        total = int(np.sum(send_counts))
        keys = torch.empty(max(1, total), dtype=torch.int64, device=self.device)
        idx = torch.empty(max(1, total), dtype=pk.idx.dtype, device=self.device)
        offsets = np.concatenate([[0], np.cumsum(send_counts)[:-1]]).astype(np.uint64)
        key_ptrs = np.full(n_parts, keys.data_ptr(), dtype=np.uint64)
        idx_ptrs = np.full(n_parts, idx.data_ptr(), dtype=np.uint64)
        self.err_word.zero_()
        self._partition(pk, splitters, n_parts, key_ptrs, idx_ptrs, offsets, key_base, skip_amb)
        return keys[:total], idx[:total]

    def gather_fragments(self, pk, dist, group, n_frag, k, sba_len):
        """All ranks' fragment lists, each sorted by its own rank (gk_frag_sort_local), back to back on the device;
        issued on the second stream so that sort and gather run beside the partition kernel.  The owner of a key
        range merges the sorted lists.  Returns (tensor, event to wait for)."""
        torch = self.torch
        world = dist.get_world_size(group)
        main = torch.cuda.current_stream()
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            self.frag_presorted = os.environ.get("GK_FRAG_PRESORT", "1") != "0"   # (0: A/B switch, tests)
            if self.frag_presorted:
                mine = torch.empty_like(pk.frag)
                _native.check(self.lib.gk_frag_sort_local(pk.frag.data_ptr(), FRAG_SHARE, int(n_frag), k, sba_len,
                                                          mine.data_ptr(), self.stream()))
            else:
                mine = pk.frag
            pk.frag.record_stream(self.side)
            if world == 1:
                out = mine
            elif backend_is_nccl(dist, group):
                out = torch.empty(world * mine.numel(), dtype=torch.uint8, device=self.device)
                dist.all_gather_into_tensor(out, mine, group=group)
            else:   # gloo cannot gather CUDA tensors: through the host (tests with two ranks on one GPU)
                mine_host = mine.cpu()
                host = torch.empty(world * mine_host.numel(), dtype=torch.uint8)
                dist.all_gather_into_tensor(host, mine_host, group=group)
                out = host.to(self.device)
            out.record_stream(main)
            ev = torch.cuda.Event()
            ev.record(self.side)
        return out, ev

    def recv_buffers(self, ref_idx, n):
        torch = self.torch
        return (torch.empty(max(1, n), dtype=torch.int64, device=self.device),
                torch.empty(max(1, n), dtype=ref_idx.dtype, device=self.device))

    def shard_sort(self, d_sba, seg_starts, k, keys_ptr, idx_ptr, idx_bytes, n_pure, n_amb, class_bit, key_bits,
                   frag_all, frag_counts, key_lo, key_hi, keep_alive=()):
        """Sort the received pairs (+ the ambiguous windows regenerated from the fragments) into a shard."""
        torch = self.torch
        handle = ctypes.c_void_p()
        _native.check(self.lib.gk_index_create(d_sba.data_ptr(), d_sba.numel(), _native.host_ptr(seg_starts),
                                               len(seg_starts), k, k, ctypes.byref(handle)))
        stats = _native.GkSortStats()
        n = n_pure + n_amb
        k_alt = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        i_alt = torch.empty(max(n, 1), dtype=torch.int32 if idx_bytes == 4 else torch.int64, device=self.device)
        counts_dev = None
        if frag_all is not None:
            counts_dev = torch.from_numpy(np.ascontiguousarray(frag_counts, dtype=np.int64)).to(self.device)
        try:
            _native.check(self.lib.gk_index_sort_shard(
                handle, keys_ptr, k_alt.data_ptr(), idx_ptr, i_alt.data_ptr(), n_pure, n_amb, class_bit, key_bits,
                frag_all.data_ptr() if frag_all is not None else None,
                counts_dev.data_ptr() if counts_dev is not None else None,
                len(frag_counts) if frag_all is not None else 0, FRAG_SHARE,
                int(getattr(self, "frag_presorted", False)), int(key_lo), int(key_hi),
                self.err_word.data_ptr(), ctypes.byref(stats), self.stream()))
        except Exception:
            self.lib.gk_index_destroy(handle)
            raise
        return {"handle": handle, "n": n, "stats": stats.as_dict(), "idx_bytes": idx_bytes}

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

    def from_host_i64(self, arr):
        return self.torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64)).to(self.device)

    @staticmethod
    def tensor_ptr(t):
        return ctypes.c_void_p(t.data_ptr())

    def wait_event(self, ev):
        self.torch.cuda.current_stream().wait_event(ev)


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



def choose_splitters(sorted_samples: np.ndarray, n_parts: int, class_bit: int = 0,
                     ambiguous_cost: float = 1.0) -> np.ndarray:
    """n_parts-1 splitters at the even quantiles of the pooled, sorted samples (uint64).

    With class_bit the parts are built from BLOCKS of equal keys: one N run makes millions of windows with ONE
    key, which no key splitter can cut, so the rank that gets that group is given correspondingly fewer other
    keys (bottleneck-optimal cuts).  ambiguous_cost weighs a key with class bit 0 against a pure one; it is 1
    since the ambiguous windows travel and sort as run-length fragments (it was 2.5 when each of them went
    through the element-wise refinement)."""
    m = len(sorted_samples)
    if n_parts <= 1 or m == 0:
        return np.zeros(0, dtype=np.uint64)
    if class_bit:
        # blocks of equal keys (a part can only begin where a key begins) and their costs
        weight = np.where((sorted_samples & np.uint64(1)) == 0, float(ambiguous_cost), 1.0)
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


def key_ranges(splitters: np.ndarray, full_bits: int):
    """(first key, end key (0: no upper bound), key width in bits) of every rank's key range, exact in uint64.
    Rank d owns the keys in [splitters[d-1], splitters[d]); it receives them minus its first key, so its local
    sort only covers the bits of its own range."""
    sp = np.ascontiguousarray(splitters, dtype=np.uint64)
    zero = np.zeros(1, dtype=np.uint64)      # (a Python int would promote the concatenation to float64)
    key_lo = np.concatenate([zero, sp])
    key_hi = np.concatenate([sp, zero])
    bits = []
    for d in range(len(sp) + 1):
        hi = int(key_hi[d]) if d < len(sp) else (1 << full_bits)
        bits.append(max(1, (hi - int(key_lo[d]) - 1).bit_length()))
    return key_lo, key_hi, bits


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
            self.total_len = 2 * n + 1
            self._layout_ev = self._layout_finish = None
            sliced = getattr(eng, "both_strands_sliced", None)
            if sliced is not None and self.world > 1 and os.environ.get("GK_LAZY_LAYOUT", "1") != "0":
                # this rank's slice of the start positions first -- with two pack tiles either side: the pack kernel
                # reads whole 4096-byte tiles (aligned down) and counts the '$' of a tile from its first byte; the
                # rest of the layout is built behind the pack kernel and is waited for before the local sort
                first, end = slice_bounds(self.total_len, self.world, self.rank)
                self.d_sba, self._layout_finish = sliced(d_fwd, max(0, first - 8192),
                                                         min(self.total_len, end + 8192 + self.k))
            else:
                self.d_sba = eng.both_strands(d_fwd)
        else:
            self.seg_starts = starts
            self.d_sba = d_fwd
            self.total_len = n
            self._layout_ev = self._layout_finish = None
        self.idx_bytes = 8 if self.total_len > 0xFFFFFFFF else 4
        if os.environ.get("GK_FORCE_IDX64", "0") not in ("", "0"):   # tests: 64-bit starts on small inputs
            self.idx_bytes = 8
        self.shard = None
        self.stats = {}
        self._marks = []
        self._is_sorted = False
        self._keep = None

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
        # every key carries the class bit, so nothing waits for the alphabet scan: its counters travel with the
        # destination counts.  k > 31: the key covers the first 31 symbols (a key range still holds whole groups);
        # the local sort compares the remaining symbols from the bytes (gk_index_sort_shard, word rounds)
        class_bit = 1
        full_bits = 2 * min(k, 31) + 2
        first, end = slice_bounds(self.total_len, world, rank)
        # (16-byte aligned cut points, so that the scan keeps its 128-bit loads)
        a16 = (first // 16) * 16 if rank > 0 else 0
        b16 = (end // 16) * 16 if rank < world - 1 else self.total_len
        alpha = eng.alphabet_async(self.d_sba[a16:b16] if world > 1 else self.d_sba)
        bytes_ready = eng.mark() if world > 1 and hasattr(eng, "mark") else None   # (before the pack kernel)
        pk = eng.pack_slice(self.d_sba, self.seg_starts, k, class_bit, first, end, self.idx_bytes)
        self._mark("pack")
        if self._layout_finish is not None:   # the rest of the both-strand layout: behind the pack kernel
            self._layout_ev = self._layout_finish()
            self._layout_finish = None

        # ---- splitters from evenly spaced samples (second stream: runs beside the pack kernel) --------------
        if world > 1:
            sample = eng.sample_keys_host(self.d_sba, self.seg_starts, k, class_bit, first, end, SAMPLES_PER_RANK,
                                          **({"after": bytes_ready} if bytes_ready is not None else {}))
            padded = np.zeros(SAMPLES_PER_RANK + 1, dtype=np.uint64)
            padded[0] = len(sample)
            padded[1:1 + len(sample)] = sample
            side = getattr(eng, "side_context", None)
            if side is not None:      # beside the pack kernel, not behind it
                with side():
                    table = self._gather(padded)
            else:
                table = self._gather(padded)
            pooled = np.sort(np.concatenate([row[1:1 + int(row[0])] for row in table]))
            # even splitters: a key range then starts at an even key, so subtracting it keeps the class bit
            splitters_host = choose_splitters(pooled, world, class_bit, AMBIGUOUS_COST) & ~np.uint64(1)
            splitters = eng.splitters_to_device(splitters_host)
        else:
            splitters_host, splitters = np.zeros(0, dtype=np.uint64), None
        self.splitters = splitters_host
        self._mark("splitters")

        # ---- ONE small all-gather: destination counts, fragment count, ambiguous windows, alphabet counters -----
        if world > 1 and hasattr(eng, "partition_counts_gathered") and backend_is_nccl(dist, self.group):
            table = eng.partition_counts_gathered(pk, splitters, world, class_bit, dist, self.group, extra=(alpha,))
        else:
            header = eng.partition_counts_host(pk, splitters, world, class_bit, extra=(alpha,))
            table = self._gather(header.astype(np.int64))
        alpha_all = table[:, 2 * world + 2:].sum(axis=0)
        if int(alpha_all[1]) != len(self.seg_starts) - 1:
            raise AssertionError("kmers compared were less than min_kmer_len: '$' inside a record")
        n_frag_src = table[:, 2 * world]
        frag_ok = (eng.supports_fragments and os.environ.get("GK_FRAGMENTS", "1") != "0"
                   and int(n_frag_src.max()) <= FRAG_SHARE)
        pure, amb = table[:, :world].copy(), table[:, world:2 * world].copy()     # [source, destination]
        if not frag_ok:             # the ambiguous windows travel as pairs like everything else
            pure += amb
            amb[:] = 0
        recv_pure, recv_amb = pure.sum(axis=0), amb.sum(axis=0)
        n_pure, n_amb = int(recv_pure[rank]), int(recv_amb[rank])
        key_lo, key_hi, bits_of = key_ranges(splitters_host, full_bits)
        key_bits = bits_of[rank] if world > 1 else full_bits
        # every rank sees the same table, so every rank computes the same capacity: the largest shard plus
        # 10 % head-room (the peer buffers are cached and only ever grow)
        capacity = int(1.1 * int((recv_pure + recv_amb).max())) + (1 << 20)
        self._mark("counts")

        # ---- partition by destination + exchange -------------------------------------------------------------
        px = None
        if (world > 1 and hasattr(eng, "peer_exchange") and os.environ.get("GK_PEER_EXCHANGE", "1") != "0"
                and capacity * (8 + self.idx_bytes) <= PeerExchange.MAX_BYTES):
            px = eng.peer_exchange(dist, self.group, capacity, self.idx_bytes)
        frag_all = frag_ev = None
        if px is not None:
            # fused: ONE kernel partitions and writes every pair straight into its destination rank's receive
            # buffer over NVLink peer memory; the fragment lists are gathered beside it
            eng.partition_to_peers(pk, splitters, world, px, pure[:rank, :].sum(axis=0), key_lo, frag_ok)
            self._mark("partition_kernel")
            if frag_ok:
                frag_all, frag_ev = eng.gather_fragments(pk, dist, self.group, n_frag_src[rank], k, self.total_len)
            self._order_after_peer_writes()
            keys_ptr, idx_ptr = px.my_keys, px.my_idx
            self._keep = None
            self.exchange_mode = "peer"
        else:
            # several hosts, no peer access, or a single rank: the same kernel into a local staging buffer, then
            # one all-to-all per array
            keys_s, idx_s = eng.partition_to_staging(pk, splitters, world, pure[rank], key_lo, frag_ok)
            if frag_ok:
                frag_all, frag_ev = eng.gather_fragments(pk, dist, self.group, n_frag_src[rank], k, self.total_len)
            keys_r, idx_r = eng.recv_buffers(idx_s, n_pure + n_amb)
            if world > 1:
                in_splits, out_splits = [int(c) for c in pure[rank]], [int(c) for c in pure[:, rank]]
                dist.all_to_all_single(keys_r[:n_pure], keys_s, output_split_sizes=out_splits,
                                       input_split_sizes=in_splits, group=self.group)
                dist.all_to_all_single(idx_r[:n_pure], idx_s, output_split_sizes=out_splits,
                                       input_split_sizes=in_splits, group=self.group)
            else:
                keys_r[:n_pure] = keys_s
                idx_r[:n_pure] = idx_s
            keys_ptr, idx_ptr = eng.tensor_ptr(keys_r), eng.tensor_ptr(idx_r)
            self._keep = (keys_r, idx_r)
            self.exchange_mode = "nccl"
        self.exchange_bytes_sent = int((pure[rank].sum() - pure[rank, rank]) * (8 + self.idx_bytes))
        del pk.keys, pk.idx
        self._mark("exchange")

        # ---- local sort: radix passes over the rank's own key range, bucket repair, fragments, flags ---------
        if self.shard is not None:
            eng.shard_free(self.shard)
        if frag_ev is not None:
            eng.wait_event(frag_ev)
        if self._layout_ev is not None:     # the rest of the both-strand layout (second stream)
            eng.wait_event(self._layout_ev)
            self._layout_ev = None
        self.shard = eng.shard_sort(self.d_sba, self.seg_starts, k, keys_ptr, idx_ptr, self.idx_bytes, n_pure, n_amb,
                                    class_bit, key_bits, frag_all, n_frag_src if frag_all is not None else None,
                                    int(key_lo[rank]), int(key_hi[rank]))
        self._keep = None
        self.stats = dict(self.shard["stats"])
        self.stats.update(n_packed=pk.n, n_shard=int(self.shard["n"]), class_bit=class_bit, key_bits_shard=key_bits,
                          fragments=bool(frag_ok))
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
        hist = _native.zeros_int64(max_counts_bin + 1)
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
