"""
GPU unit tests of the libgkb200 building blocks, called through the C ABI and compared with NumPy /
the CPU oracle.  Bit-exact: everything here is integer / byte work.
"""
import ctypes

import numpy as np
import pytest

import oracle
from genome_kmers import _native
import gpu_utils as gu

pytestmark = pytest.mark.gpu


def test_library_reports_blackwell():
    lib = _native.lib()
    sm, major, minor, mem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_uint64()
    _native.check(lib.gk_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor),
                                     ctypes.byref(mem)))
    assert major.value == 10, f"libgkb200 is built for sm_100a, found cc {major.value}.{minor.value}"
    assert sm.value > 0


@pytest.mark.parametrize("n", [1, 15, 16, 17, 1000, 4096, 100003])
def test_revcomp_and_both_strands(n):
    torch = gu.torch_mod()
    rng = np.random.default_rng(n)
    sba = np.frombuffer(b"ACGTRYSWKMBDHVN$", dtype=np.uint8)[rng.integers(0, 16, n)].copy()
    lib = _native.lib()
    d_in = gu.dev(sba)
    d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
    _native.check(lib.gk_sba_revcomp(d_in.data_ptr(), n, d_out.data_ptr(), gu.stream()))
    assert np.array_equal(d_out.cpu().numpy(), oracle.revcomp(sba))
    d_both = torch.zeros(2 * n + 1, dtype=torch.uint8, device="cuda")
    _native.check(lib.gk_sba_both_strands(d_in.data_ptr(), n, d_both.data_ptr(), gu.stream()))
    expect, _ = oracle.both_strands(sba, np.array([0], dtype=np.uint32))
    assert np.array_equal(d_both.cpu().numpy(), expect)


def test_scan_alphabet_counts():
    rng = np.random.default_rng(7)
    sba = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 70001)].copy()
    sba[[5, 1000, 69999]] = ord("$")
    sba[[6, 7, 4000]] = ord("N")
    sba[[9]] = ord("R")
    sba[[10, 20000]] = ord("x")
    counts = np.zeros(3, dtype=np.uint64)
    d = gu.dev(sba)
    _native.check(_native.lib().gk_sba_scan_alphabet(d.data_ptr(), len(sba), _native.host_ptr(counts),
                                                     gu.stream()))
    assert counts.tolist() == [2, 3, 4]


@pytest.mark.parametrize("k", [1, 3, 21, 31])
def test_init_indices(k):
    torch = gu.torch_mod()
    rng = np.random.default_rng(k)
    recs = gu.random_genome(rng, 50000, 7)
    sba, starts = oracle.build_sba([r for _, r in recs])
    expect = oracle.init_indices(starts, len(sba), k)
    segs, segs_p = gu.segs_ptr(starts)
    out = torch.zeros(len(expect), dtype=torch.int32, device="cuda")
    _native.check(_native.lib().gk_kmer_init_indices(segs_p, len(segs), len(sba), k, 4, out.data_ptr(),
                                                     gu.stream()))
    assert np.array_equal(gu.host(out, np.uint32).astype(np.uint64), expect)


@pytest.mark.parametrize("k,class_bit,amb", [(1, 0, False), (5, 1, True), (16, 1, True), (21, 0, False),
                                             (31, 1, True), (32, 0, False), (31, 0, False)])
def test_pack_keys_matches_definition(k, class_bit, amb):
    rng = np.random.default_rng(100 + k)
    recs = gu.random_genome(rng, 9000, 3, n_runs=4 if amb else 0, run_lo=5, run_hi=60,
                            n_scatter=40 if amb else 0)
    sba, starts = oracle.build_sba([r for _, r in recs])
    keys, idx, n_amb = gu.pack_keys(sba, starts, k, k, class_bit)
    exp_keys, exp_idx = gu.expected_pack(sba, starts, k, k, class_bit)
    assert np.array_equal(idx.astype(np.uint64), exp_idx)
    assert np.array_equal(idx.astype(np.uint64), oracle.init_indices(starts, len(sba), k))
    bad = np.flatnonzero(keys != exp_keys)
    assert len(bad) == 0, f"first mismatch at window {bad[:5]}: got {keys[bad[:5]]}, want {exp_keys[bad[:5]]}"
    if class_bit:
        assert n_amb == int((exp_keys & np.uint64(1) == 0).sum())


def test_pack_keys_slices_concatenate():
    """A GPU's slice [first, end) of the byte array packs exactly its share, in order."""
    rng = np.random.default_rng(5)
    recs = gu.random_genome(rng, 30000, 4, n_runs=3, n_scatter=10)
    sba, starts = oracle.build_sba([r for _, r in recs])
    full_k, full_i, _ = gu.pack_keys(sba, starts, 21, 21, 1)
    cuts = [0, 4096, 9999, 17000, len(sba)]
    parts = [gu.pack_keys(sba, starts, 21, 21, 1, first=a, end=b) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), full_k)
    assert np.array_equal(np.concatenate([p[1] for p in parts]), full_i)


@pytest.mark.parametrize("n", [1, 2, 33, 4095, 4096, 4097, 50000, 1 << 20, 3_000_001])
@pytest.mark.parametrize("bits", [(0, 64), (0, 43), (3, 29), (56, 64), (0, 8), (0, 5)])
def test_radix_sort_pairs_is_a_stable_sort(n, bits):
    rng = np.random.default_rng(n + bits[0] * 131 + bits[1])
    keys = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    if n > 1000:
        keys[rng.integers(0, n, n // 3)] = keys[0]  # long ties exercise stability
    vals = np.arange(n, dtype=np.uint32)
    got_k, got_v = gu.radix_sort_pairs(keys, vals, bits[0], bits[1])
    width = bits[1] - bits[0]
    mask = np.uint64((1 << width) - 1) if width < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    field = (keys >> np.uint64(bits[0])) & mask
    order = np.argsort(field, kind="stable")
    assert np.array_equal(got_v, vals[order])
    assert np.array_equal(got_k, keys[order])


def test_radix_sort_pairs_u64_values_and_skewed_digits():
    rng = np.random.default_rng(11)
    n = 700_001
    keys = (rng.integers(0, 4, n, dtype=np.uint64) << np.uint64(40)) | rng.integers(0, 3, n, dtype=np.uint64)
    vals = rng.integers(0, 1 << 62, n, dtype=np.uint64)
    got_k, got_v = gu.radix_sort_pairs(keys, vals, 0, 48)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(got_k, keys[order]) and np.array_equal(got_v, vals[order])


@pytest.mark.parametrize("cfg", ["0", "1", "2", "3"])
@pytest.mark.parametrize("n,bits,val_dtype", [
    (1, (0, 32), np.uint32), (2, (0, 32), np.uint32), (8191, (0, 32), np.uint32), (8193, (0, 32), np.uint32),
    (12289, (5, 21), np.uint32), (500_003, (0, 32), np.uint32), (500_003, (24, 32), np.uint32),
    (300_001, (0, 32), np.uint64), (2_000_003, (0, 16), np.uint32),
])
def test_radix_sort_pairs32_is_a_stable_sort(n, bits, val_dtype, cfg, monkeypatch):
    """The 32-bit-key variant of the onesweep sort, every tile shape (GK_SORT32_CFG is read per call)."""
    monkeypatch.setenv("GK_SORT32_CFG", cfg)
    rng = np.random.default_rng(n + bits[0] * 131 + bits[1])
    keys = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    if n > 1000:
        keys[rng.integers(0, n, n // 3)] = keys[0]  # long ties exercise stability
    vals = (np.arange(n, dtype=np.uint64) * np.uint64(3 if val_dtype == np.uint64 else 1)).astype(val_dtype)
    got_k, got_v = gu.radix_sort_pairs32(keys, vals, bits[0], bits[1])
    field = (keys >> np.uint32(bits[0])) & np.uint32((1 << (bits[1] - bits[0])) - 1)
    order = np.argsort(field, kind="stable")
    assert np.array_equal(got_v, vals[order])
    assert np.array_equal(got_k, keys[order])


@pytest.mark.parametrize("n", [1, 7, 8, 9, 4096, 100_000, 1_234_567])
def test_rle_and_group_hist(n):
    torch = gu.torch_mod()
    rng = np.random.default_rng(n)
    keys = np.sort(rng.integers(0, max(2, n // 3), n, dtype=np.uint64))
    if n > 5000:
        keys[100:3000] = keys[100]  # one big group (> the shared-memory histogram range)
        keys = np.sort(keys)
    d_keys = gu.dev(keys)
    d_off = torch.zeros(n, dtype=torch.int64, device="cuda")
    n_groups = ctypes.c_uint64(0)
    lib = _native.lib()
    _native.check(lib.gk_rle_keys(d_keys.data_ptr(), n, d_off.data_ptr(), ctypes.byref(n_groups), gu.stream()))
    uniq, first, counts = np.unique(keys, return_index=True, return_counts=True)
    assert n_groups.value == len(uniq)
    assert np.array_equal(gu.host(d_off, np.uint64)[:len(uniq)], first.astype(np.uint64))
    for min_g, max_g, max_bin in [(1, 0, 1000000), (2, 0, 10), (1, 3, 2), (2, 5000, 2500)]:
        hist = np.zeros(max_bin + 1, dtype=np.int64)
        total = ctypes.c_int64(0)
        _native.check(lib.gk_group_size_hist(d_off.data_ptr(), n_groups.value, n, min_g, max_g, max_bin,
                                             _native.host_ptr(hist), ctypes.byref(total), gu.stream()))
        keep = counts[(counts >= min_g) & ((max_g == 0) | (counts <= max_g))]
        exp = np.bincount(np.minimum(keep, max_bin), minlength=max_bin + 1)
        assert total.value == int(keep.sum())
        assert np.array_equal(hist, exp)
