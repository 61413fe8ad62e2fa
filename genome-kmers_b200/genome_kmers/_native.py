"""
ctypes binding of libgkb200.so (include/gkb200.h).  There is NO fallback: if the CUDA library
is missing or a call fails, an exception is raised.
"""
import ctypes
import os

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_PKG_DIR), "lib", "libgkb200.so")

GK_OK = 0
GK_ERR_CUDA = 1
GK_ERR_ARG = 2
GK_ERR_UNSUPPORTED = 3
GK_ERR_INTERNAL = 4
GK_ERR_INVALID_KMERS = 5
GK_ERR_STATE = 6

FILTER_KEEP_ALL = 0
FILTER_NO_AMBIGUOUS = 1
FILTER_MIN_LENGTH = 2
FILTER_HOMOPOLYMER = 3
FILTER_GC_COUNT = 4
FILTER_NGG_PAM = 5


class GkFilter(ctypes.Structure):
    _fields_ = [("id", ctypes.c_int32), ("p0", ctypes.c_int64), ("p1", ctypes.c_int64),
                ("p2", ctypes.c_int64)]


class GkSortStats(ctypes.Structure):
    _fields_ = [
        ("pack_ms", ctypes.c_float), ("hist_ms", ctypes.c_float), ("sort_ms", ctypes.c_float),
        ("fixup_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
        ("sort_passes", ctypes.c_int32), ("key_bits", ctypes.c_int32), ("levels", ctypes.c_int32),
        ("gpu_launches", ctypes.c_int32), ("n_windows", ctypes.c_uint64),
        ("n_ambiguous", ctypes.c_uint64),
        ("n_fragments", ctypes.c_uint64),
        ("refine_flags", ctypes.c_uint64),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class NativeError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


_lib = None

# name -> (restype, argtypes); every symbol include/gkb200.h declares
_vp, _u64, _u32, _int = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
_p = ctypes.POINTER
SIGNATURES = {
    "gk_version": (_int, []),
    "gk_status_string": (ctypes.c_char_p, [_int]),
    "gk_last_error": (ctypes.c_char_p, []),
    "gk_device_info": (_int, [_p(_int), _p(_int), _p(_int), _p(_u64)]),
    "gk_launch_count": (_u64, [_int]),
    "gk_sba_scan_alphabet": (_int, [_vp, _u64, _vp, _vp]),
    "gk_sba_scan_alphabet_async": (_int, [_vp, _u64, _vp, _vp]),
    "gk_sba_revcomp": (_int, [_vp, _u64, _vp, _vp]),
    "gk_sba_both_strands": (_int, [_vp, _u64, _vp, _vp]),
    "gk_sba_both_strands_range": (_int, [_vp, _u64, _vp, _u64, _u64, _vp]),
    "gk_kmer_count": (_int, [_vp, _u32, _u64, _u32, _p(_u64)]),
    "gk_kmer_init_indices": (_int, [_vp, _u32, _u64, _u32, _int, _vp, _vp]),
    "gk_pack_keys": (_int, [_vp, _u64, _vp, _u32, _u32, _u32, _int, _u64, _u64, _vp, _int, _vp,
                            _u64, _p(_u64), _p(_u64), _vp]),
    "gk_radix_sort_pairs": (_int, [_vp, _vp, _vp, _vp, _int, _u64, _int, _int, _p(_int), _vp]),
    "gk_radix_sort_pairs32": (_int, [_vp, _vp, _vp, _vp, _int, _u64, _int, _int, _p(_int), _vp]),
    "gk_partition_pairs": (_int, [_vp, _vp, _vp, _vp, _int, _u64, _vp, _u32, _vp, _vp]),
    "gk_partition_count": (_int, [_vp, _u64, _vp, _u32, _vp, _vp]),
    "gk_partition_pairs_peer": (_int, [_vp, _vp, _int, _u64, _vp, _u32, _vp, _vp, _vp, _vp, _int, _vp, _vp]),
    "gk_partition_count_split": (_int, [_vp, _u64, _vp, _u32, _int, _vp, _vp]),
    "gk_peer_alloc": (_int, [_u64, _p(_vp)]),
    "gk_peer_free": (_int, [_vp]),
    "gk_peer_export": (_int, [_vp, _vp]),
    "gk_peer_open": (_int, [_vp, _p(_vp)]),
    "gk_peer_close": (_int, [_vp]),
    "gk_rle_keys": (_int, [_vp, _u64, _vp, _p(_u64), _vp]),
    "gk_group_size_hist": (_int, [_vp, _u64, _u64, _u64, _u64, _u64, _vp, _p(ctypes.c_int64), _vp]),
    "gk_index_create": (_int, [_vp, _u64, _vp, _u32, _u32, _u32, _p(_vp)]),
    "gk_index_destroy": (None, [_vp]),
    "gk_index_size": (_u64, [_vp]),
    "gk_index_idx_bytes": (_int, [_vp]),
    "gk_index_is_sorted": (_int, [_vp]),
    "gk_index_set_indices": (_int, [_vp, _vp, _u64, _int, _int, _vp]),
    "gk_index_sort": (_int, [_vp, _p(GkSortStats), _vp]),
    "gk_index_sort_pairs": (_int, [_vp, _vp, _vp, _vp, _vp, _u64, _int, _p(GkSortStats), _vp]),
    "gk_index_sort_shard": (_int, [_vp, _vp, _vp, _vp, _vp, _u64, _u64, _int, _int, _vp, _vp, _u32, _u64, _int, _u64,
                                   _u64, _vp, _p(GkSortStats), _vp]),
    "gk_frag_sort_local": (_int, [_vp, _u64, _u64, _u32, _u64, _vp, _vp]),
    "gk_pack_slice": (_int, [_vp, _u64, _vp, _u32, _u32, _int, _u64, _u64, _vp, _int, _vp, _u64, _p(_u64), _vp, _u64,
                             _vp, _vp]),
    "gk_sample_keys": (_int, [_vp, _u64, _vp, _u32, _u32, _int, _u64, _u64, _u32, _vp, _p(_u32), _vp]),
    "gk_index_device_indices": (_int, [_vp, _p(_vp), _vp]),
    "gk_index_copy_indices": (_int, [_vp, _vp, _vp]),
    "gk_index_group_counts": (_int, [_vp, _u32, _p(GkFilter), _u64, _u64, _u64, _vp,
                                     _p(ctypes.c_int64), _vp]),
    "gk_index_group_counts_zeroed": (_int, [_vp, _u32, _p(GkFilter), _u64, _u64, _u64, _vp,
                                            _p(ctypes.c_int64), _p(_u64), _vp]),
    "gk_index_group_counts_sparse": (_int, [_vp, _u32, _p(GkFilter), _u64, _u64, _u64, _vp, _vp, _u64,
                                            _p(_u64), _p(ctypes.c_int64), _vp]),
    "gk_index_groups": (_int, [_vp, _u32, _p(_u64), _vp, _vp, _vp]),
    "gk_index_groups_filtered": (_int, [_vp, _u32, _p(GkFilter), _p(_u64), _p(_u64), _vp, _vp, _vp, _vp]),
    "gk_index_verify": (_int, [_vp, _u32, _vp, _vp, _vp]),
    "gk_popcount_words": (_int, [_vp, _u64, _p(_u64), _vp]),
    "gk_sort_count_host": (_int, [_vp, _u64, _vp, _u32, _u32, _int, _int, _vp, _u64, _vp,
                                  _p(ctypes.c_int64), _p(_u64), _p(GkSortStats)]),
}


def lib():
    """Load libgkb200.so (once).  Raises if it has not been built -- no CPU fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C genome-kmers_b200/csrc`.  genome_kmers has no CPU fallback."
            )
        L = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def last_error() -> str:
    return lib().gk_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    """Map a gk_status to the exception type the reference raises for the same condition."""
    if status == GK_OK:
        return
    msg = last_error() or lib().gk_status_string(status).decode()
    if status == GK_ERR_ARG:
        raise ValueError(msg)
    if status == GK_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if status in (GK_ERR_INVALID_KMERS, GK_ERR_STATE):
        raise AssertionError(msg)
    raise NativeError(status, f"libgkb200: {msg}")


def host_ptr(arr: np.ndarray) -> int:
    assert arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data


def launch_count(reset: bool = False) -> int:
    return int(lib().gk_launch_count(1 if reset else 0))


def zeros_int64(n: int):
    """np.zeros(n, int64) whose pages are untouched zero pages even after many calls: glibc serves repeated
    multi-megabyte calloc requests from its heap (and then clears them by hand, 0.3 ms for the reference's
    default 8 MB histogram); an anonymous mapping is lazily zero every time."""
    import mmap

    import numpy as np

    nbytes = int(n) * 8
    if nbytes < (1 << 20):
        return np.zeros(n, dtype=np.int64)
    return np.frombuffer(mmap.mmap(-1, nbytes), dtype=np.int64)
