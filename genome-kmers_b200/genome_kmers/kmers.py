"""
Kmers: drop-in mirror of the reference class (/root/reference/src/genome_kmers/kmers.py:651-1737)
whose hot path -- sort() and the group counting behind get_kmer_count / get_kmer_group_counts --
runs as hand-written sm_100a kernels in libgkb200.so (include/gkb200.h).  This module is host
logic only: argument validation with the reference's messages, lazy host<->device movement, and
thin calls through ctypes.  There is no CPU implementation of the hot path here.

Differences from the reference, all additive (defaults behave the same):
  * source_strand "reverse_complement" and "both" are implemented (the reference raises
    NotImplementedError, kmers.py:693-696).  "both" indexes the byte array
    forward || '$' || revcomp: start i < L is a forward k-mer, i > L is revcomp_sba[i - L - 1].
  * more than 2^32 - 1 k-mers switch the start indices to uint64 instead of raising
    (kmers.py:805-808).
  * ties (equal k-mers) are always ordered by ascending start index -- the reference's
    break_ties=True order (kmers.py:1710-1711); its default order is unspecified.
  * filters are `KmerFilter` objects (callable like the reference's numba closures, and carrying
    the parameters the device predicate needs).  Arbitrary Python callables cannot run on the GPU.
"""
import shelve
from pathlib import Path
from typing import Callable, Generator, Optional, Union

import ctypes
import numpy as np

from genome_kmers import _native
from genome_kmers.sequence_collection import SequenceCollection

_SEP = 36


# ---------------------------------------------------------------------------------------------
# filters (kmers.py:14-259)
# ---------------------------------------------------------------------------------------------
class KmerFilter:
    """A k-mer predicate with a device implementation.  Calling it evaluates one k-mer on the host
    with the reference's semantics: filter(sba, sba_strand, kmer_sba_start_idx) -> bool."""

    def __init__(self, filter_id: int, p0: int = 0, p1: int = 0, p2: int = 0, name: str = ""):
        self.filter_id, self.p0, self.p1, self.p2 = int(filter_id), int(p0), int(p1), int(p2)
        self.name = name or f"filter{filter_id}"

    def native(self) -> _native.GkFilter:
        return _native.GkFilter(self.filter_id, self.p0, self.p1, self.p2)

    def __repr__(self):
        return f"KmerFilter({self.name}, {self.p0}, {self.p1}, {self.p2})"

    def __call__(self, sba: np.ndarray, sba_strand: str, kmer_sba_start_idx: int) -> bool:
        s, n = int(kmer_sba_start_idx), len(sba)
        fid = self.filter_id
        if fid == _native.FILTER_KEEP_ALL:
            return True
        if fid == _native.FILTER_NO_AMBIGUOUS:
            k = self.p0
            if s + k > n:
                raise ValueError(f"kmer_len ({k}) is invalid. It extends beyond len(sba)")
            for b in sba[s:s + k]:
                if b == _SEP:
                    raise ValueError(f"end of segment was reached. kmer_len ({k}) invalid.")
                if b not in (65, 84, 71, 67):
                    return False
            return True
        if fid == _native.FILTER_MIN_LENGTH:
            return kmer_has_required_len(sba, s, self.p0)
        if fid == _native.FILTER_HOMOPOLYMER:
            max_run, k = self.p0, self.p1
            msg = f"The kmer_len ({k}) requested is too large for kmer_sba_start_idx ({s})"
            if s + k - 1 >= n:
                raise ValueError(msg)
            if k < max_run:
                return True
            run = 1
            for i in range(s + 1, s + k):
                if sba[i] == _SEP:
                    raise ValueError(msg)
                run = run + 1 if sba[i] == sba[i - 1] else 1
                if run > max_run:
                    return False
            return True
        if fid == _native.FILTER_GC_COUNT:
            lo, hi, k = self.p0, self.p1, self.p2
            if hi < lo:
                return False
            gc = 0
            for b in sba[s:s + k]:
                if b == _SEP:
                    raise ValueError(
                        f"The kmer_len ({k}) requested is too larger for kmer_sba_start_idx ({s})")
                if b in (71, 67):
                    gc += 1
                    if gc > hi:
                        return False
            return lo <= gc <= hi
        if fid == _native.FILTER_NGG_PAM:
            if s + 23 > n:
                raise ValueError("The guide defined at this start index extends beyond the sba")
            return bool(sba[s + 21] == 71 and sba[s + 22] == 71)
        raise ValueError(f"unknown filter id {fid}")


kmer_filter_keep_all = KmerFilter(_native.FILTER_KEEP_ALL, name="keep_all")
crispr_ngg_pam_filter = KmerFilter(_native.FILTER_NGG_PAM, name="crispr_ngg_pam")


def gen_kmer_length_filter_func(min_kmer_len: int) -> KmerFilter:
    return KmerFilter(_native.FILTER_MIN_LENGTH, min_kmer_len, name="min_length")


def gen_kmer_homopolymer_filter_func(max_homopolymer_size: int, kmer_len: int) -> KmerFilter:
    if max_homopolymer_size < 1:
        raise ValueError(f"max_homopolymer_size ({max_homopolymer_size}) must be >= 1")
    if kmer_len < 1:
        raise ValueError(f"kmer_len ({kmer_len}) must be >= 1")
    return KmerFilter(_native.FILTER_HOMOPOLYMER, max_homopolymer_size, kmer_len, name="homopolymer")


def gen_kmer_gc_content_filter_func(min_allowed_gc_frac: float, max_allowed_gc_frac: float,
                                    kmer_len: int) -> KmerFilter:
    if min_allowed_gc_frac > max_allowed_gc_frac:
        raise ValueError(
            f"min_allowed_gc_frac ({min_allowed_gc_frac}) must be <= max_allowed_gc_frac ({max_allowed_gc_frac})")
    if min_allowed_gc_frac < 0.0 or min_allowed_gc_frac > 1.0:
        raise ValueError(f"min_allowed_gc_frac ({min_allowed_gc_frac}) must be in the range [0.0, 1.0]")
    if max_allowed_gc_frac < 0.0 or max_allowed_gc_frac > 1.0:
        raise ValueError(f"max_allowed_gc_frac ({max_allowed_gc_frac}) must be in the range [0.0, 1.0]")
    # the reference converts the fractions to counts once (kmers.py:143-144)
    lo = int(np.ceil(kmer_len * min_allowed_gc_frac))
    hi = int(np.floor(kmer_len * max_allowed_gc_frac))
    return KmerFilter(_native.FILTER_GC_COUNT, lo, hi, kmer_len, name="gc_content")


def gen_no_ambiguous_bases_filter(kmer_len: int) -> KmerFilter:
    return KmerFilter(_native.FILTER_NO_AMBIGUOUS, kmer_len, name="no_ambiguous_bases")


# ---------------------------------------------------------------------------------------------
# scalar helpers kept for API parity (kmers.py:262-397); host-side, one pair at a time
# ---------------------------------------------------------------------------------------------
def kmer_has_required_len(sba: np.ndarray, sba_start_idx: int, min_kmer_len: int) -> bool:
    for idx in range(sba_start_idx, sba_start_idx + min_kmer_len):
        if idx >= len(sba) or sba[idx] == _SEP:
            return False
    return True


def compare_sba_kmers_lexicographically(sba_a, sba_b, kmer_sba_start_idx_a: int,
                                        kmer_sba_start_idx_b: int,
                                        max_kmer_len: Union[int, None] = None):
    """(comparison, last_kmer_index_compared) with the reference's '$' rules (kmers.py:306-397)."""
    j = 0
    while True:
        ia, ib = kmer_sba_start_idx_a + j, kmer_sba_start_idx_b + j
        a_out = ia >= len(sba_a) or sba_a[ia] == _SEP
        b_out = ib >= len(sba_b) or sba_b[ib] == _SEP
        if a_out or b_out:
            if j - 1 < 0:
                raise AssertionError("There were no valid kmer bases to compare")
            return (-1 if a_out and not b_out else 1 if b_out and not a_out else 0), j - 1
        if sba_a[ia] != sba_b[ib]:
            return (-1 if sba_a[ia] < sba_b[ib] else 1), j
        if max_kmer_len is not None and j == max_kmer_len - 1:
            return 0, j
        j += 1


def get_compare_sba_kmers_func(kmer_len):
    def compare_sba_kmers_func(sba_a, sba_b, kmer_sba_start_idx_a, kmer_sba_start_idx_b):
        return compare_sba_kmers_lexicographically(
            sba_a, sba_b, kmer_sba_start_idx_a, kmer_sba_start_idx_b, max_kmer_len=kmer_len)

    compare_sba_kmers_func.kmer_len = kmer_len      # the device group walk reads the length from here
    return compare_sba_kmers_func


def compare_sba_kmers_always_less_than(sba_a, sba_b, kmer_sba_start_idx_a: int, kmer_sba_start_idx_b: int,
                                       max_kmer_len: Union[int, None] = None):
    """Every k-mer is its own group: what the reference passes for an unsorted index (kmers.py:295-303)."""
    return -1, 0


def get_kmer_info_minimal(kmer_num: int, kmer_sba_start_indices, sba, kmer_len: Union[int, None],
                          group_size_yielded: int, group_size_total: int):
    """(kmer_num, group_size_yielded, group_size_total) -- kmers.py:400-425."""
    return kmer_num, group_size_yielded, group_size_total


def get_kmer_info_group_size_only(kmer_num: int, kmer_sba_start_indices, sba, kmer_len: Union[int, None],
                                  group_size_yielded: int, group_size_total: int):
    """group_size_total only -- kmers.py:428-451."""
    return group_size_total


class _ArrayIndex:
    """A native index over caller-supplied arrays (sequence byte array + start indices): what the module-level
    seams of the reference take instead of a Kmers object (kmers.py:454-648).  Records are found from the '$'
    separators.  The arrays are uploaded once; the group walk and the histogram run on the GPU."""

    def __init__(self, sba: np.ndarray, kmer_start_indices: np.ndarray, is_sorted: bool):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("genome_kmers needs a CUDA device: the k-mer hot path has no CPU fallback")
        self.lib = _native.lib()
        sba = np.ascontiguousarray(sba, dtype=np.uint8)
        self.d_sba = torch.from_numpy(sba).to("cuda")
        seps = np.flatnonzero(sba == _SEP)
        starts = np.ascontiguousarray(np.concatenate([[0], seps + 1]), dtype=np.uint64)
        self.handle = ctypes.c_void_p()
        _native.check(self.lib.gk_index_create(self.d_sba.data_ptr(), len(sba), _native.host_ptr(starts), len(starts),
                                               1, 0, ctypes.byref(self.handle)))
        want = np.uint32 if self.lib.gk_index_idx_bytes(self.handle) == 4 else np.uint64
        idx = np.ascontiguousarray(kmer_start_indices, dtype=want)
        _native.check(self.lib.gk_index_set_indices(self.handle, _native.host_ptr(idx), len(idx), idx.itemsize,
                                                    int(is_sorted), self.stream()))

    @staticmethod
    def stream() -> int:
        return int(_torch().cuda.current_stream().cuda_stream)

    def close(self):
        if self.handle is not None:
            self.lib.gk_index_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _comparison_mode(kmer_comparison_func, kmer_len):
    """(is_sorted, kmer_len) for the device group walk from the reference's comparison-function argument."""
    if kmer_comparison_func is compare_sba_kmers_always_less_than:
        return False, kmer_len
    if hasattr(kmer_comparison_func, "kmer_len"):
        return True, kmer_comparison_func.kmer_len
    raise NotImplementedError(
        "kmer_comparison_func must be compare_sba_kmers_always_less_than or come from "
        "get_compare_sba_kmers_func(kmer_len): arbitrary Python callables cannot run on the GPU")


def _check_group_limits(min_group_size, max_group_size, yield_first_n=None):
    if min_group_size < 1:
        raise ValueError(f"min_group_size ({min_group_size}) must be >= 1")
    if max_group_size is not None and max_group_size < min_group_size:
        raise ValueError(
            f"if max_group_size ({max_group_size}) is specified, it must be >= min_group_size ({min_group_size})")
    if yield_first_n is not None and yield_first_n < 1:
        raise ValueError(f"if yield_first_n ({yield_first_n}) is specified, it must be > 0")


def get_kmer_group_size_hist(sba, sba_strand: str, kmer_len: Union[int, None], kmer_start_indices,
                             kmer_comparison_func: Callable, kmer_filter_func: Callable, min_group_size: int = 1,
                             max_group_size: Union[int, None] = None, max_counts_bin: int = 1000000):
    """(counts_by_group_size, total_kmer_count) for caller-supplied arrays -- the reference's internal seam 2
    (kmers.py:454-520), computed by gk_index_group_counts on the GPU."""
    if max_counts_bin <= 0:
        raise ValueError(f"max_counts_bin ({max_counts_bin}) must be >= 1")
    _check_group_limits(min_group_size, max_group_size)
    is_sorted, cmp_len = _comparison_mode(kmer_comparison_func, kmer_len)
    flt = Kmers._native_filter(kmer_filter_func)
    hist = np.zeros(max_counts_bin + 1, dtype=np.int64)
    if len(kmer_start_indices) == 0:
        return hist, 0
    ai = _ArrayIndex(sba, kmer_start_indices, is_sorted)
    try:
        total, top = ctypes.c_int64(0), ctypes.c_uint64(0)
        _native.check(ai.lib.gk_index_group_counts_zeroed(
            ai.handle, cmp_len or 0, ctypes.byref(flt), min_group_size, max_group_size or 0, max_counts_bin,
            _native.host_ptr(hist), ctypes.byref(total), ctypes.byref(top), ai.stream()))
    finally:
        ai.close()
    return hist, int(total.value)


def kmer_info_by_group_generator(sba, sba_strand: str, kmer_len: Union[int, None], kmer_start_indices,
                                 kmer_comparison_func: Callable, kmer_filter_func: Callable,
                                 kmer_info_func: Callable, min_group_size: int = 1,
                                 max_group_size: Union[int, None] = None, yield_first_n: Union[int, None] = None):
    """The reference's group walk over caller-supplied arrays (kmers.py:523-648): k-mers that fail the filter are
    skipped, a passing k-mer is compared with the previous PASSING one, and the first yield_first_n members of
    every group within the size limits are passed to kmer_info_func.  The filter and the grouping run on the GPU
    (gk_index_groups_filtered); only the calls of kmer_info_func are a host loop."""
    _check_group_limits(min_group_size, max_group_size, yield_first_n)
    is_sorted, cmp_len = _comparison_mode(kmer_comparison_func, kmer_len)
    flt = Kmers._native_filter(kmer_filter_func)
    if len(kmer_start_indices) == 0:
        return
    ai = _ArrayIndex(sba, kmer_start_indices, is_sorted)
    try:
        n_kept, n_groups = ctypes.c_uint64(0), ctypes.c_uint64(0)
        _native.check(ai.lib.gk_index_groups_filtered(ai.handle, cmp_len or 0, ctypes.byref(flt), ctypes.byref(n_kept),
                                                      ctypes.byref(n_groups), None, None, None, ai.stream()))
        kept_pos = np.zeros(n_kept.value, dtype=np.uint64)
        offsets = np.zeros(n_groups.value, dtype=np.uint64)
        sizes = np.zeros(n_groups.value, dtype=np.uint64)
        if n_kept.value:
            _native.check(ai.lib.gk_index_groups_filtered(
                ai.handle, cmp_len or 0, ctypes.byref(flt), ctypes.byref(n_kept), ctypes.byref(n_groups),
                _native.host_ptr(kept_pos), _native.host_ptr(offsets), _native.host_ptr(sizes), ai.stream()))
    finally:
        ai.close()
    for off, size in zip(offsets.tolist(), sizes.tolist()):
        if size < min_group_size or (max_group_size is not None and size > max_group_size):
            continue
        n_yield = size if yield_first_n is None else min(size, yield_first_n)
        for member in range(off, off + n_yield):
            yield kmer_info_func(int(kept_pos[member]), kmer_start_indices, sba, kmer_len, n_yield, size)


def _torch():
    import torch  # deferred: importing torch costs seconds and is only needed for device work

    return torch


def _unsorted_group_msg(name: str, value) -> str:
    return ("Returning group parameters is not supported when kmers has not been"
            f" sorted. {name} ({value}) cannot be specified. Did you"
            " mean to run sort() before getting kmers?")


class Kmers:
    """Memory-efficient k-mer calculations on a genome, with the hot path on a B200."""

    def __init__(self, seq_coll: Union[SequenceCollection, None] = None, min_kmer_len: int = 1,
                 max_kmer_len: Union[int, None] = None, source_strand: str = "forward",
                 track_strands_separately: bool = False, method: str = "single_pass",
                 device: Optional[Union[int, str]] = None, num_gpus: Optional[int] = None) -> None:
        if track_strands_separately:
            raise NotImplementedError(
                f"This function has not been implemented for track_strands_separately = '{track_strands_separately}'")
        if source_strand not in ("forward", "reverse_complement", "both"):
            raise ValueError(f"source_strand ({source_strand}) not recognized")
        if min_kmer_len < 1:
            raise ValueError(f"min_kmer_len ({min_kmer_len}) must be greater than zero")
        if max_kmer_len is not None:
            if max_kmer_len < 1:
                raise ValueError(f"max_kmer_len ({max_kmer_len}) must be greater than zero")
            if max_kmer_len < min_kmer_len:
                raise ValueError(f"max_kmer_len ({max_kmer_len}) is less than min_kmer_len ({min_kmer_len})")

        self.min_kmer_len = min_kmer_len
        self.max_kmer_len = max_kmer_len
        self.kmer_source_strand = source_strand
        self.track_strands_separately = track_strands_separately
        self._is_initialized = False
        self._is_set = False
        self._is_sorted = False
        self.last_sort_stats = None

        self._device = device
        # num_gpus > 1 (or GK_NUM_GPUS): sort() and the group counts run key-range sharded over the ranks of the
        # default torch.distributed process group, one rank per GPU (genome_kmers.distributed.ShardedKmers);
        # every rank constructs the same Kmers and makes the same calls.  Default: this process's GPU only.
        self._num_gpus = num_gpus
        self._sk = None            # ShardedKmers once sort() took the multi-GPU path
        self._ix = None            # gk_index handle
        self._d_sba = None         # torch uint8 tensor the index borrows
        self._host_idx = None      # cached host copy of the start indices
        self._host_idx_dirty = False  # host copy was assigned by the user and not yet uploaded
        self._n_kmers = None

        if seq_coll is None:
            return

        # record lengths first, strand agreement last: the reference's order of checks (kmers.py:730-754)
        loaded = seq_coll.strands_loaded()
        lengths = [end - start + 1
                   for _, start, end in seq_coll.iter_records("forward" if loaded == "both" else None)]
        if len(lengths) == 0:
            raise ValueError("sequence_collection is empty")
        if min_kmer_len > min(lengths):
            raise ValueError(
                f"min_kmer_len ({min_kmer_len}) must be <= the shortest sequence length ({min(lengths)})")
        if loaded != source_strand:
            raise ValueError(
                f"source_strand ({source_strand}) does not match sequence_collection loaded strand ({loaded})")

        self.seq_coll = seq_coll
        self._initialize(method=method)

    # ------------------------------------------------------------------ initialisation
    def _initialize(self, kmer_filters=[], method: str = "single_pass"):
        if kmer_filters != []:
            raise NotImplementedError("kmer_filters have not been implemented")
        if method == "double_pass":
            raise NotImplementedError(f"method '{method}' has not been implemented")
        if method != "single_pass":
            raise ValueError(f"method '{method}' not recognized")
        self._n_kmers = self._get_unfiltered_kmer_count()
        self._is_initialized = True

    def _strand_layout(self):
        """(host byte array or None, uint64 segment starts, total length) of the indexed array."""
        sc = self.seq_coll
        if self.kmer_source_strand == "forward":
            return sc.forward_sba, sc._forward_sba_seg_starts.astype(np.uint64), len(sc.forward_sba)
        if self.kmer_source_strand == "reverse_complement":
            return sc.revcomp_sba, sc._revcomp_sba_seg_starts.astype(np.uint64), len(sc.revcomp_sba)
        n = len(sc.forward_sba)
        starts = np.concatenate([sc._forward_sba_seg_starts.astype(np.uint64),
                                 sc._revcomp_sba_seg_starts.astype(np.uint64) + np.uint64(n + 1)])
        return None, starts, 2 * n + 1

    def _get_unfiltered_kmer_count(self) -> int:
        """Sum over records of (len - min_kmer_len + 1) (kmers.py:837-861)."""
        _, starts, total = self._strand_layout()
        if len(starts) == 0:
            raise ValueError("SequenceCollection does not have any records")
        ends_excl = np.concatenate([starts[1:].astype(np.int64) - 1, [total]])
        return int((ends_excl - starts.astype(np.int64) - self.min_kmer_len + 1).sum())

    # ------------------------------------------------------------------ device plumbing
    def _stream(self) -> int:
        return int(_torch().cuda.current_stream().cuda_stream)

    def _ensure_device(self) -> None:
        """Upload the byte array (once) and create the native index.  Fails without a GPU."""
        if self._ix is not None:
            # every later call allocates and launches on the current device: it must be the one that holds
            # the byte array and the index
            current = _torch().cuda.current_device()
            if current != self._device_index:
                raise RuntimeError(f"this Kmers object lives on cuda:{self._device_index} but the current CUDA device "
                                   f"is cuda:{current}; call torch.cuda.set_device({self._device_index}) first")
            return
        torch = _torch()
        lib = _native.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("genome_kmers needs a CUDA device: the k-mer hot path has no CPU fallback")
        if self._device is not None:
            torch.cuda.set_device(self._device)
        self._device_index = torch.cuda.current_device()
        host_sba, starts, total = self._strand_layout()
        stream = self._stream()
        if host_sba is not None:
            self._d_sba = torch.from_numpy(np.ascontiguousarray(host_sba)).to("cuda", non_blocking=True)
        else:
            fwd = torch.from_numpy(np.ascontiguousarray(self.seq_coll.forward_sba)).to("cuda", non_blocking=True)
            self._d_sba = torch.empty(total, dtype=torch.uint8, device="cuda")
            _native.check(lib.gk_sba_both_strands(fwd.data_ptr(), fwd.numel(), self._d_sba.data_ptr(), stream))
            torch.cuda.current_stream().synchronize()
            del fwd
        starts = np.ascontiguousarray(starts, dtype=np.uint64)
        handle = ctypes.c_void_p()
        _native.check(lib.gk_index_create(
            self._d_sba.data_ptr(), total, _native.host_ptr(starts), len(starts), self.min_kmer_len,
            self.max_kmer_len or 0, ctypes.byref(handle)))
        self._ix = handle
        self._n_kmers = int(lib.gk_index_size(self._ix))

    def _push_host_indices(self) -> None:
        if not self._host_idx_dirty:
            return
        self._ensure_device()
        lib = _native.lib()
        want = np.uint32 if lib.gk_index_idx_bytes(self._ix) == 4 else np.uint64
        arr = np.ascontiguousarray(self._host_idx, dtype=want)
        _native.check(lib.gk_index_set_indices(self._ix, _native.host_ptr(arr), len(arr), arr.itemsize,
                                               int(self._is_sorted), self._stream()))
        self._n_kmers = len(arr)
        self._host_idx_dirty = False

    def __del__(self):
        try:
            if self._ix is not None:
                _native.lib().gk_index_destroy(self._ix)
                self._ix = None
        except Exception:
            pass

    # the reference exposes the index array as a plain attribute (tests/test_kmers.py:146)
    @property
    def kmer_sba_start_indices(self):
        if not self._is_initialized and self._host_idx is None:
            return None
        if self._host_idx is None and self._sk is not None:
            # multi-GPU path: every rank receives the whole sorted array (the shards in rank order)
            import torch.distributed as dist

            parts = [None] * self._sk.world
            dist.all_gather_object(parts, self._sk.local_start_indices())
            self._host_idx = np.concatenate(parts)
        if self._host_idx is None:
            self._ensure_device()
            lib = _native.lib()
            torch = _torch()
            wide = lib.gk_index_idx_bytes(self._ix) == 8
            # pinned staging (torch caches pinned blocks) so the D2H copy runs at PCIe speed
            buf = torch.empty(self._n_kmers, dtype=torch.int64 if wide else torch.int32, pin_memory=True)
            _native.check(lib.gk_index_copy_indices(self._ix, buf.data_ptr(), self._stream()))
            self._host_idx = buf.numpy().view(np.uint64 if wide else np.uint32)
        return self._host_idx

    @kmer_sba_start_indices.setter
    def kmer_sba_start_indices(self, value):
        self._host_idx = None if value is None else np.asarray(value)
        self._host_idx_dirty = value is not None
        if value is not None:
            self._n_kmers = len(self._host_idx)

    def device_start_indices(self):
        """The (sorted) start indices as a torch tensor on the device: a copy, so it stays valid when a later
        sort() or load() replaces the index's own buffer."""
        self._ensure_device()
        self._push_host_indices()
        torch, lib = _torch(), _native.lib()
        ptr = ctypes.c_void_p()
        _native.check(lib.gk_index_device_indices(self._ix, ctypes.byref(ptr), self._stream()))
        nbytes = lib.gk_index_idx_bytes(self._ix)
        return _tensor_from_ptr(torch, ptr.value, self._n_kmers, nbytes).clone()

    def __len__(self):
        if self._host_idx is not None:
            return len(self._host_idx)
        return self._n_kmers

    def __getitem__(self):
        pass

    # ------------------------------------------------------------------ the hot path
    def _sharded_world(self) -> int:
        """Number of ranks when this object is to take the multi-GPU path, else 0."""
        import os

        n = self._num_gpus if self._num_gpus is not None else int(os.environ.get("GK_NUM_GPUS", "0") or 0)
        if n <= 1:
            return 0
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("num_gpus > 1 needs an initialised torch.distributed process group with one rank "
                               "per GPU (launch with torchrun)")
        if dist.get_world_size() != n:
            raise ValueError(f"num_gpus ({n}) does not match the process group's world size ({dist.get_world_size()})")
        return n

    def _sort_sharded(self):
        from genome_kmers.distributed import ShardedKmers

        if self.max_kmer_len != self.min_kmer_len:
            raise NotImplementedError("the multi-GPU path sorts fixed-length k-mers (min_kmer_len == max_kmer_len)")
        if self._host_idx_dirty:
            raise NotImplementedError("assigned start indices cannot be sorted on the multi-GPU path")
        sc = self.seq_coll
        if self.kmer_source_strand == "reverse_complement":
            sba, starts, strands = sc.revcomp_sba, sc._revcomp_sba_seg_starts, "forward"
        else:
            sba, starts = sc.forward_sba, sc._forward_sba_seg_starts
            strands = "both" if self.kmer_source_strand == "both" else "forward"
        if self._sk is not None:
            self._sk.close()
        self._sk = ShardedKmers(np.ascontiguousarray(sba), starts.astype(np.uint64), self.min_kmer_len, strands)
        self._sk.sort()
        self.last_sort_stats = dict(self._sk.stats)
        self._host_idx = None
        self._is_sorted = True

    def local_start_indices(self) -> np.ndarray:
        """Multi-GPU path: this rank's shard of the sorted start indices (the shards of ranks 0, 1, ... in that
        order are the global order).  Single GPU: the whole array."""
        return self._sk.local_start_indices() if self._sk is not None else self.kmer_sba_start_indices

    def sort(self):
        """Sort the start indices in place by k-mer (kmers.py:1624-1652).  Runs on the GPU."""
        if self._sharded_world():
            return self._sort_sharded()
        self._ensure_device()
        self._push_host_indices()
        stats = _native.GkSortStats()
        _native.check(_native.lib().gk_index_sort(self._ix, ctypes.byref(stats), self._stream()))
        self.last_sort_stats = stats.as_dict()
        self._host_idx = None
        self._is_sorted = True

    def _check_group_args(self, kmer_len, min_group_size, max_group_size, yield_first_n="unset"):
        if kmer_len is not None and kmer_len < 1:
            raise ValueError(f"kmer_len ({kmer_len}) must be > 0")
        if not self._is_sorted:
            if min_group_size != 1:
                raise ValueError(_unsorted_group_msg("min_group_size", min_group_size))
            if max_group_size is not None:
                raise ValueError(_unsorted_group_msg("max_group_size", max_group_size))
            if yield_first_n != "unset" and yield_first_n is not None:
                raise ValueError(_unsorted_group_msg("yield_first_n", yield_first_n))

    @staticmethod
    def _native_filter(kmer_filter_func) -> _native.GkFilter:
        if isinstance(kmer_filter_func, KmerFilter):
            return kmer_filter_func.native()
        raise NotImplementedError(
            "kmer_filter_func must be one of this module's KmerFilter objects (kmer_filter_keep_all, "
            "gen_no_ambiguous_bases_filter(...), ...): arbitrary Python callables cannot run on the GPU")

    def _group_counts(self, kmer_len, kmer_filter_func, min_group_size, max_group_size, max_counts_bin,
                      want_hist):
        if max_counts_bin <= 0:
            raise ValueError(f"max_counts_bin ({max_counts_bin}) must be >= 1")
        if min_group_size < 1:
            raise ValueError(f"min_group_size ({min_group_size}) must be >= 1")
        if max_group_size is not None and max_group_size < min_group_size:
            raise ValueError(
                f"if max_group_size ({max_group_size}) is specified, it must be >= min_group_size ({min_group_size})")
        flt = self._native_filter(kmer_filter_func)
        if self._sk is not None and not self._host_idx_dirty:
            if kmer_len != self.min_kmer_len:
                raise NotImplementedError("the multi-GPU path counts groups for the sort length only")
            hist, total = self._sk.get_kmer_group_counts(kmer_len, flt, min_group_size, max_group_size,
                                                         max_counts_bin)
            return (hist if want_hist else None), total
        self._ensure_device()
        self._push_host_indices()
        # np.zeros hands out untouched zero pages: the library writes the occupied bins only
        hist = _native.zeros_int64(max_counts_bin + 1) if want_hist else None
        total, top = ctypes.c_int64(0), ctypes.c_uint64(0)
        _native.check(_native.lib().gk_index_group_counts_zeroed(
            self._ix, kmer_len or 0, ctypes.byref(flt), min_group_size, max_group_size or 0,
            max_counts_bin, None if hist is None else _native.host_ptr(hist), ctypes.byref(total),
            ctypes.byref(top), self._stream()))
        return hist, int(total.value)

    def get_kmer_count(self, kmer_len: Union[int, None], kmer_filter_func: Callable = kmer_filter_keep_all,
                       min_group_size: int = 1, max_group_size: Union[int, None] = None) -> int:
        """Total number of k-mers in groups that pass the filter and size limits (kmers.py:994-1083)."""
        self._check_group_args(kmer_len, min_group_size, max_group_size)
        _, total = self._group_counts(kmer_len, kmer_filter_func, min_group_size, max_group_size,
                                      1000000, want_hist=False)
        return total

    def get_kmer_group_counts(self, kmer_len: Union[int, None],
                              kmer_filter_func: Callable = kmer_filter_keep_all, min_group_size: int = 1,
                              max_group_size: Union[int, None] = None,
                              max_counts_bin: int = 1000000):
        """(counts_by_group_size int64[max_counts_bin+1], total_kmer_count) (kmers.py:1085-1178)."""
        self._check_group_args(kmer_len, min_group_size, max_group_size)
        if not self._is_sorted:
            raise AssertionError("The kmers must be sorted when calling get_kmer_group_counts")
        return self._group_counts(kmer_len, kmer_filter_func, min_group_size, max_group_size,
                                  max_counts_bin, want_hist=True)

    def get_kmer_groups(self, kmer_len: Union[int, None]):
        """(offsets, sizes) uint64 arrays: one entry per distinct k-mer of the sorted index, i.e.
        the unique-k-mer table (positions into kmer_sba_start_indices and multiplicities)."""
        if not self._is_sorted:
            raise AssertionError("The kmers must be sorted when calling get_kmer_groups")
        self._ensure_device()
        self._push_host_indices()
        lib = _native.lib()
        n_groups = ctypes.c_uint64(0)
        _native.check(lib.gk_index_groups(self._ix, kmer_len or 0, ctypes.byref(n_groups), None, None,
                                          self._stream()))
        offsets = np.zeros(n_groups.value, dtype=np.uint64)
        sizes = np.zeros(n_groups.value, dtype=np.uint64)
        if n_groups.value:
            _native.check(lib.gk_index_groups(self._ix, kmer_len or 0, ctypes.byref(n_groups),
                                              _native.host_ptr(offsets), _native.host_ptr(sizes),
                                              self._stream()))
        return offsets, sizes

    def verify_order(self, kmer_len: Union[int, None] = None) -> dict:
        """Self-check of the current order on the device, straight from the sequence bytes (gk_index_verify):
        neighbours compared with the reference's '$'-terminated comparator, ties in ascending start order,
        every start a valid and distinct k-mer start.  `ok` is True for a correctly sorted index."""
        self._ensure_device()
        self._push_host_indices()
        if kmer_len is None:
            kmer_len = self.max_kmer_len
        report = np.zeros(8, dtype=np.uint64)
        _native.check(_native.lib().gk_index_verify(self._ix, kmer_len or 0, _native.host_ptr(report), None,
                                                    self._stream()))
        names = ("kmers", "out_of_order", "tie_order", "invalid_starts", "duplicate_starts", "groups",
                 "flag_mismatches", "flags_compared")
        out = {k: int(v) for k, v in zip(names, report)}
        out["ok"] = not any(out[k] for k in ("out_of_order", "tie_order", "invalid_starts", "duplicate_starts",
                                             "flag_mismatches"))
        return out

    # ------------------------------------------------------------------ host-side accessors
    def get_kmers(self, kmer_len: Union[int, None], one_based_seq_index: bool = False,
                  kmer_filter_func: Callable = kmer_filter_keep_all, kmer_info_to_yield: str = "minimum",
                  min_group_size: int = 1, max_group_size: Union[int, None] = None,
                  yield_first_n: Union[int, None] = None) -> Generator[tuple, None, None]:
        """Generator over k-mers group by group (kmers.py:869-992).  The grouping is computed on the
        GPU (group table); only the tuple construction is a host loop."""
        self._check_group_args(kmer_len, min_group_size, max_group_size, yield_first_n)
        if kmer_info_to_yield not in ("minimum", "full"):
            raise ValueError(f"kmer_info_to_yield ({kmer_info_to_yield}) not recognized")
        if min_group_size < 1:
            raise ValueError(f"min_group_size ({min_group_size}) must be >= 1")
        if max_group_size is not None and max_group_size < min_group_size:
            raise ValueError(
                f"if max_group_size ({max_group_size}) is specified, it must be >= min_group_size ({min_group_size})")
        if yield_first_n is not None and yield_first_n < 1:
            raise ValueError(f"if yield_first_n ({yield_first_n}) is specified, it must be > 0")
        flt = kmer_filter_func
        if not isinstance(flt, KmerFilter):
            self._native_filter(flt)
        idx = self.kmer_sba_start_indices
        kept_pos = None                      # positions (kmer_num) of the k-mers that pass the filter
        if flt.filter_id != _native.FILTER_KEEP_ALL:
            kept_pos, offsets, sizes = self._filtered_groups(kmer_len, flt)
        elif self._is_sorted:
            offsets, sizes = self.get_kmer_groups(kmer_len)
        else:
            offsets = np.arange(len(idx), dtype=np.uint64)
            sizes = np.ones(len(idx), dtype=np.uint64)
        locate = None
        if kmer_info_to_yield == "full":
            locate = self._locate_all(idx, one_based_seq_index)
        for off, size in zip(offsets.tolist(), sizes.tolist()):
            if size < min_group_size or (max_group_size is not None and size > max_group_size):
                continue
            n_yield = size if yield_first_n is None else min(size, yield_first_n)
            for member in range(off, off + n_yield):
                kmer_num = member if kept_pos is None else int(kept_pos[member])
                if locate is None:
                    yield kmer_num, n_yield, size
                else:
                    strand, chrom, seq_idx, seg_end = locate(kmer_num)
                    if kmer_len is None:
                        this_len = seg_end - int(idx[kmer_num]) + 1
                    else:
                        this_len = kmer_len
                        if int(idx[kmer_num]) + kmer_len - 1 > seg_end:
                            raise ValueError(
                                f"kmer_len ({kmer_len}) for kmer_num ({kmer_num}) extends beyond the end of the segment")
                    yield kmer_num, strand, chrom, seq_idx, this_len, n_yield, size

    def _filtered_groups(self, kmer_len, flt: KmerFilter):
        """(kept_pos, offsets, sizes): group table over the k-mers that pass the filter (device-side)."""
        self._ensure_device()
        self._push_host_indices()
        lib = _native.lib()
        nat = flt.native()
        n_kept, n_groups = ctypes.c_uint64(0), ctypes.c_uint64(0)
        _native.check(lib.gk_index_groups_filtered(self._ix, kmer_len or 0, ctypes.byref(nat), ctypes.byref(n_kept),
                                                   ctypes.byref(n_groups), None, None, None, self._stream()))
        kept_pos = np.zeros(n_kept.value, dtype=np.uint64)
        offsets = np.zeros(n_groups.value, dtype=np.uint64)
        sizes = np.zeros(n_groups.value, dtype=np.uint64)
        if n_kept.value:
            _native.check(lib.gk_index_groups_filtered(
                self._ix, kmer_len or 0, ctypes.byref(nat), ctypes.byref(n_kept), ctypes.byref(n_groups),
                _native.host_ptr(kept_pos), _native.host_ptr(offsets), _native.host_ptr(sizes), self._stream()))
        return kept_pos, offsets, sizes

    def _locate_all(self, idx, one_based):
        """kmer_num -> (strand symbol, record name, forward sequence index, segment end)."""
        sc = self.seq_coll
        strand = self.kmer_source_strand
        if strand == "both":
            return self._locate_all_both(idx, one_based)
        seg, seq_idx = sc.locate_sba_indices(idx, strand, one_based)
        sba, starts, names = sc._strand_arrays(strand)
        ends = np.concatenate([starts[1:].astype(np.int64) - 2, [len(sba) - 1]])
        symbol = "+" if strand == "forward" else "-"

        def locate(kmer_num):
            s = int(seg[kmer_num])
            return symbol, names[s], int(seq_idx[kmer_num]), int(ends[s])

        return locate

    def _locate_all_both(self, idx, one_based):
        """source_strand='both': start i < L is forward_sba[i] ('+'), i > L is revcomp_sba[i - L - 1] ('-'),
        each located exactly as the reference locates an index of that strand
        (sequence_collection.py:930-978).  The segment end is in the coordinates of the indexed array."""
        sc = self.seq_coll
        n_fwd = len(sc.forward_sba)
        idx = np.asarray(idx, dtype=np.int64)
        is_rc = idx > n_fwd
        seg = np.empty(len(idx), dtype=np.int64)
        seq_idx = np.empty(len(idx), dtype=np.int64)
        seg[~is_rc], seq_idx[~is_rc] = sc.locate_sba_indices(idx[~is_rc], "forward", one_based)
        seg[is_rc], seq_idx[is_rc] = sc.locate_sba_indices(idx[is_rc] - (n_fwd + 1), "reverse_complement",
                                                         one_based)
        ends, names = {}, {}
        for rc, strand in ((False, "forward"), (True, "reverse_complement")):
            sba, starts, names[rc] = sc._strand_arrays(strand)
            ends[rc] = (np.concatenate([starts[1:].astype(np.int64) - 2, [len(sba) - 1]])
                        + (n_fwd + 1 if rc else 0))

        def locate(kmer_num):
            rc = bool(is_rc[kmer_num])
            s = int(seg[kmer_num])
            return "-" if rc else "+", names[rc][s], int(seq_idx[kmer_num]), int(ends[rc][s])

        return locate

    def generate_get_kmer_info_func(self, one_based_seq_index: bool) -> Callable:
        """get_kmer_info(kmer_num, kmer_sba_start_indices, sba, kmer_len, group_size_yielded, group_size_total) ->
        (kmer_num, strand, chrom, seq_start_idx, kmer_len, yielded, total) -- kmers.py:1180-1264."""
        get_record_info_from_sba_index = self.seq_coll.generate_get_record_info_from_sba_index_func(
            one_based_seq_index)

        def get_kmer_info(kmer_num, kmer_sba_start_indices, sba, kmer_len, group_size_yielded, group_size_total):
            if kmer_num < 0:
                raise ValueError(f"kmer_num ({kmer_num}) cannot be less than zero")
            if kmer_num >= len(kmer_sba_start_indices):
                raise ValueError(
                    f"kmer_num ({kmer_num}) is out of bounds (num kmers = {len(kmer_sba_start_indices)})")
            sba_idx = int(kmer_sba_start_indices[kmer_num])
            _, _, seg_end, seq_strand, seq_chrom, seq_start_idx = get_record_info_from_sba_index(sba_idx)
            if kmer_len is None:
                kmer_len = seg_end - sba_idx + 1
            elif sba_idx + kmer_len - 1 > seg_end:
                raise ValueError(
                    f"kmer_len ({kmer_len}) for kmer_num ({kmer_num}) extends beyond the end of the segment")
            return kmer_num, seq_strand, seq_chrom, seq_start_idx, kmer_len, group_size_yielded, group_size_total

        return get_kmer_info

    def get_is_less_than_func(self, validate_kmers: bool = True, break_ties: bool = False) -> Callable:
        """is_less_than(start_a, start_b) with the reference's semantics (kmers.py:1654-1731), one pair at a time
        on the host: the scalar definition of the order that sort() produces on the GPU (with break_ties=True;
        sort() does not call it)."""
        if self.kmer_source_strand != "forward" or self.seq_coll.strands_loaded() != "forward":
            raise NotImplementedError(
                f"both kmer_source_strand ({self.kmer_source_strand}) and "
                "sequence_collection.strands_loaded() must be 'forward'")
        sba = self.seq_coll.forward_sba
        min_kmer_len, max_kmer_len = self.min_kmer_len, self.max_kmer_len

        def is_less_than(kmer_sba_start_idx_a: int, kmer_sba_start_idx_b: int) -> bool:
            comparison, last = compare_sba_kmers_lexicographically(
                sba, sba, kmer_sba_start_idx_a, kmer_sba_start_idx_b, max_kmer_len=max_kmer_len)
            if comparison < 0:
                a_lt_b = True
            elif comparison > 0:
                a_lt_b = False
            else:
                a_lt_b = bool(break_ties and kmer_sba_start_idx_a < kmer_sba_start_idx_b)
            if validate_kmers:
                rest = min_kmer_len - (last + 1)
                if not (kmer_has_required_len(sba, kmer_sba_start_idx_a + last + 1, rest)
                        and kmer_has_required_len(sba, kmer_sba_start_idx_b + last + 1, rest)):
                    raise AssertionError(
                        f"kmers compared were less than min_kmer_len ({min_kmer_len}).  Was "
                        "kmer_sba_start_indices initialized correctly?")
            return a_lt_b

        return is_less_than

    def _indexed_bytes(self):
        host_sba, _, _ = self._strand_layout()
        if host_sba is None:
            sc = self.seq_coll
            host_sba = np.concatenate([sc.forward_sba, np.array([_SEP], dtype=np.uint8), sc.revcomp_sba])
        return host_sba

    def get_kmer_str_no_checks(self, kmer_num: int, kmer_strand: str, kmer_len: int) -> str:
        if kmer_strand == "-":
            raise NotImplementedError("Only implemented for kmer_strand='+'")
        if kmer_strand != "+":
            raise ValueError(f"kmer_strand ({kmer_strand}) not recognized")
        start = int(self.kmer_sba_start_indices[kmer_num])
        return self.seq_coll.forward_sba[start:start + kmer_len].tobytes().decode("utf-8")

    def get_kmer_str(self, kmer_num: int, kmer_len: Union[int, None] = None) -> str:
        """The kmer_num'th k-mer of the current order as a string (kmers.py:1561-1622)."""
        if kmer_num < 0:
            raise ValueError(f"kmer_num ({kmer_num}) cannot be less than zero")
        if kmer_num >= len(self):
            raise ValueError(f"kmer_num ({kmer_num}) is out of bounds (num kmers = {len(self)})")
        if kmer_len is not None and kmer_len < self.min_kmer_len:
            raise ValueError(f"kmer_len ({kmer_len}) is less than min_kmer_len ({self.min_kmer_len})")
        if self.max_kmer_len is not None and kmer_len is not None and kmer_len > self.max_kmer_len:
            raise ValueError(f"kmer_len ({kmer_len}) is greater than max_kmer_len ({self.max_kmer_len})")
        sba = self._indexed_bytes()
        _, starts, total = self._strand_layout()
        start = int(self.kmer_sba_start_indices[kmer_num])
        seg = int(np.searchsorted(starts, start, side="right")) - 1
        seg_end = total - 1 if seg == len(starts) - 1 else int(starts[seg + 1]) - 2
        if kmer_len is None:
            longest = seg_end - start + 1
            kmer_len = longest if self.max_kmer_len is None else min(self.max_kmer_len, longest)
        if start + kmer_len - 1 > seg_end:
            raise ValueError(
                f"kmer_len ({kmer_len}) for kmer_num ({kmer_num}) extends beyond the end of the segment")
        return sba[start:start + kmer_len].tobytes().decode("utf-8")

    # ------------------------------------------------------------------ equality / persistence
    def __ne__(self, other):
        return not self.__eq__(other)

    def __eq__(self, other):
        for attr in ("min_kmer_len", "max_kmer_len", "kmer_source_strand", "track_strands_separately",
                     "_is_initialized", "_is_set", "_is_sorted"):
            if getattr(self, attr) != getattr(other, attr):
                return False
        a, b = self.kmer_sba_start_indices, other.kmer_sba_start_indices
        if (a is None) != (b is None) or (a is not None and not np.array_equal(a, b)):
            return False
        return self.seq_coll == other.seq_coll

    _PERSISTED = ("min_kmer_len", "max_kmer_len", "kmer_source_strand", "track_strands_separately",
                  "_is_initialized", "_is_set", "_is_sorted")

    def save(self, save_file_path: Path, include_sequence_collection: bool = False, format: str = "hdf5",
             mode: str = "w") -> None:
        """Same on-disk layout as the reference (kmers.py:1400-1499)."""
        if format == "hdf5":
            from genome_kmers.sequence_collection import _require_h5py

            h5py = _require_h5py()
            with h5py.File(save_file_path, mode) as file:
                grp = file.create_group("kmers")
                for attr in self._PERSISTED:
                    value = getattr(self, attr)
                    grp[attr] = 0 if value is None else value
                idx = self.kmer_sba_start_indices
                grp["kmer_sba_start_indices"] = np.array([], dtype=np.uint32) if idx is None else idx
            if include_sequence_collection:
                self.seq_coll.save(save_file_path, mode="a", format="hdf5")
        elif format == "shelve":
            with shelve.open(str(save_file_path)) as db:
                for attr in self._PERSISTED:
                    db[attr] = getattr(self, attr)
                db["kmer_sba_start_indices"] = self.kmer_sba_start_indices
            if include_sequence_collection:
                self.seq_coll.save(save_file_path, format="shelve")
        else:
            raise ValueError(f"format ({format}) not recognized")

    def load(self, load_file_path: Path, seq_coll: Union[SequenceCollection, None] = None,
             format: str = "hdf5") -> None:
        if format == "hdf5":
            from genome_kmers.sequence_collection import _require_h5py

            h5py = _require_h5py()
            with h5py.File(load_file_path, "r") as file:
                grp = file["kmers"]
                self.min_kmer_len = int(grp["min_kmer_len"][()])
                max_len = int(grp["max_kmer_len"][()])
                self.max_kmer_len = None if max_len == 0 else max_len
                self.kmer_source_strand = grp["kmer_source_strand"][()].decode("utf-8")
                self.track_strands_separately = bool(grp["track_strands_separately"][()])
                self._is_initialized = bool(grp["_is_initialized"][()])
                self._is_set = bool(grp["_is_set"][()])
                self._is_sorted = bool(grp["_is_sorted"][()])
                idx = grp["kmer_sba_start_indices"][:]
                idx = None if idx.shape == (0,) else idx
        elif format == "shelve":
            with shelve.open(str(load_file_path)) as db:
                for attr in self._PERSISTED:
                    setattr(self, attr, db[attr])
                idx = db["kmer_sba_start_indices"]
        else:
            raise ValueError(f"format ({format}) not recognized")
        if seq_coll is not None:
            self.seq_coll = seq_coll
        else:
            self.seq_coll = SequenceCollection()
            self.seq_coll.load(load_file_path, format=format)
        if self._ix is not None:
            _native.lib().gk_index_destroy(self._ix)
            self._ix = None
        self.kmer_sba_start_indices = idx

    def to_csv(self, kmer_len, output_file_path, fields=["kmer"]):
        pass


def _tensor_from_ptr(torch, ptr: int, n: int, itemsize: int):
    """Zero-copy torch view of library-owned device memory (valid while the index lives)."""
    dtype = torch.int32 if itemsize == 4 else torch.int64

    class _Holder:
        __cuda_array_interface__ = {
            "shape": (n,), "typestr": "<i4" if itemsize == 4 else "<i8", "data": (ptr, False),
            "version": 3,
        }

    return torch.as_tensor(_Holder(), device="cuda", dtype=dtype)
