"""Helpers for the `-m gpu` tests: thin wrappers that call libgkb200 primitives through the C ABI
with torch tensors as device memory, plus NumPy expectations."""
import ctypes

import numpy as np

from genome_kmers import _native


def torch_mod():
    import torch

    return torch


def dev(arr: np.ndarray):
    torch = torch_mod()
    t = torch.from_numpy(np.ascontiguousarray(arr).view(_signed(arr.dtype)))
    return t.to("cuda")


def _signed(dtype):
    return {np.dtype(np.uint64): np.int64, np.dtype(np.uint32): np.int32,
            np.dtype(np.uint8): np.uint8}.get(np.dtype(dtype), dtype)


def host(t, dtype):
    return t.cpu().numpy().view(dtype)


def stream():
    return int(torch_mod().cuda.current_stream().cuda_stream)


def segs_ptr(seg_starts):
    arr = np.ascontiguousarray(seg_starts, dtype=np.uint64)
    return arr, _native.host_ptr(arr)


def radix_sort_pairs(keys: np.ndarray, vals: np.ndarray, begin_bit: int, end_bit: int):
    torch = torch_mod()
    lib = _native.lib()
    k0, v0 = dev(keys.astype(np.uint64)), dev(vals)
    k1, v1 = torch.empty_like(k0), torch.empty_like(v0)
    in_alt = ctypes.c_int(0)
    _native.check(lib.gk_radix_sort_pairs(k0.data_ptr(), k1.data_ptr(), v0.data_ptr(), v1.data_ptr(),
                                          vals.dtype.itemsize, len(keys), begin_bit, end_bit,
                                          ctypes.byref(in_alt), stream()))
    torch.cuda.synchronize()
    kk, vv = (k1, v1) if in_alt.value else (k0, v0)
    return host(kk, np.uint64), host(vv, vals.dtype)


def radix_sort_pairs32(keys: np.ndarray, vals: np.ndarray, begin_bit: int, end_bit: int):
    torch = torch_mod()
    lib = _native.lib()
    k0, v0 = dev(keys.astype(np.uint32)), dev(vals)
    k1, v1 = torch.empty_like(k0), torch.empty_like(v0)
    in_alt = ctypes.c_int(0)
    _native.check(lib.gk_radix_sort_pairs32(k0.data_ptr(), k1.data_ptr(), v0.data_ptr(), v1.data_ptr(),
                                            vals.dtype.itemsize, len(keys), begin_bit, end_bit,
                                            ctypes.byref(in_alt), stream()))
    torch.cuda.synchronize()
    kk, vv = (k1, v1) if in_alt.value else (k0, v0)
    return host(kk, np.uint32), host(vv, vals.dtype)


def pack_keys(sba: np.ndarray, seg_starts, valid_len, key_len, class_bit, idx_dtype=np.uint32,
              first=0, end=None):
    torch = torch_mod()
    lib = _native.lib()
    d_sba = dev(sba)
    segs, segs_p = segs_ptr(seg_starts)
    end = len(sba) if end is None else end
    n_cap = len(sba)
    keys = torch.zeros(n_cap, dtype=torch.int64, device="cuda")
    idx = torch.zeros(n_cap, dtype=torch.int32 if idx_dtype == np.uint32 else torch.int64, device="cuda")
    n_out, n_amb = ctypes.c_uint64(0), ctypes.c_uint64(0)
    _native.check(lib.gk_pack_keys(d_sba.data_ptr(), len(sba), segs_p, len(segs), valid_len, key_len,
                                   class_bit, first, end, keys.data_ptr(), np.dtype(idx_dtype).itemsize,
                                   idx.data_ptr(), n_cap, ctypes.byref(n_out), ctypes.byref(n_amb),
                                   stream()))
    torch.cuda.synchronize()
    n = n_out.value
    return host(keys, np.uint64)[:n], host(idx, idx_dtype)[:n], n_amb.value


_ACGT_CODE = {65: 0, 67: 1, 71: 2, 84: 3}


def expected_key(window: bytes, class_bit: int) -> int:
    """Python restatement of the key definition in csrc/gk_pack.cu (slow; small inputs only)."""
    k = len(window)
    value, pure = 0, True
    for j, b in enumerate(window):
        if b in _ACGT_CODE:
            value = (value << 2) | _ACGT_CODE[b]
        else:
            below = sum(1 for c in (65, 67, 71, 84) if c < b)
            value = (value << (2 * (k - j))) + (below << (2 * (k - j - 1)))
            pure = False
            break
    return ((value << 1) | int(pure)) if class_bit else value


def expected_pack(sba: np.ndarray, seg_starts, valid_len, key_len, class_bit):
    sba_b = sba.tobytes()
    starts = list(np.asarray(seg_starts, dtype=np.int64))
    keys, idx = [], []
    for s, a in enumerate(starts):
        e_excl = starts[s + 1] - 1 if s + 1 < len(starts) else len(sba)
        for i in range(a, e_excl - valid_len + 1):
            keys.append(expected_key(sba_b[i:i + key_len], class_bit))
            idx.append(i)
    return np.array(keys, dtype=np.uint64), np.array(idx, dtype=np.uint64)


def random_genome(rng, n_bases, n_records, n_runs=0, run_lo=10, run_hi=100, n_scatter=0):
    """uint8 records of random ACGT with optional N runs / scattered IUPAC letters."""
    # equal-length records, remainder in the last one (reference generator, profiling.py:27-53)
    avg = n_bases // n_records
    bounds = [i * avg for i in range(n_records)] + [n_bases]
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n_bases)].copy()
    for _ in range(n_runs):
        ln = int(rng.integers(run_lo, run_hi + 1))
        st = int(rng.integers(0, max(1, n_bases - ln)))
        bases[st:st + ln] = ord("N")
    if n_scatter:
        letters = np.frombuffer(b"RYSWKMBDHVN", dtype=np.uint8)
        pos = rng.integers(0, n_bases, n_scatter)
        bases[pos] = letters[rng.integers(0, len(letters), n_scatter)]
    return [(f"chr{i}", bases[bounds[i]:bounds[i + 1]]) for i in range(n_records)]
