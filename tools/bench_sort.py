#!/usr/bin/env python
"""Micro-benchmark of gk_radix_sort_pairs (tuning aid, not the bench): one process per tile config.
    python tools/bench_sort.py [n] [configs...]"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "genome-kmers_b200"))


def child(n):
    import torch
    from genome_kmers import _native

    if os.environ.get("GK_BENCH_LIB"):   # a second build of the library (an experiment compiled with -D...)
        _native.LIB_PATH = os.environ["GK_BENCH_LIB"]
    lib = _native.lib()
    g = torch.Generator(device="cuda").manual_seed(1)
    keys = torch.randint(-(1 << 62), 1 << 62, (n,), dtype=torch.int64, device="cuda", generator=g)
    vals = torch.arange(n, dtype=torch.int32, device="cuda")
    k0, v0 = keys.clone(), vals.clone()
    k1, v1 = torch.empty_like(keys), torch.empty_like(vals)
    sp = int(torch.cuda.current_stream().cuda_stream)
    in_alt = ctypes.c_int(0)
    times = []
    for it in range(6):
        k0.copy_(keys); v0.copy_(vals)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _native.check(lib.gk_radix_sort_pairs(k0.data_ptr(), k1.data_ptr(), v0.data_ptr(), v1.data_ptr(), 4, n,
                                              0, 64, ctypes.byref(in_alt), sp))
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    out = k1 if in_alt.value else k0
    u = out ^ (-(1 << 63))  # unsigned order == signed order after flipping the sign bit
    ok = bool((u[1:] >= u[:-1]).all())
    best = min(times[2:])
    print(f"cfg={os.environ.get('GK_SORT_CFG', 'default')} n={n} sort_ms={best:.3f} per_pass_ms={best / 8:.3f} "
          f"GBps_per_pass={24 * n / (best / 8) / 1e6:.0f} (incl. histogram) sorted={ok}", flush=True)


def variants(n, env_name, values):
    """One process, every value of an environment switch the library reads per call (GK_SORT_CFG, or a
    switch added for an experiment): a stability and value check against torch's stable sort on small
    inputs, then the timing at n.  Back-to-back A/B/A/B in one process removes the box-to-box noise."""
    import torch
    from genome_kmers import _native

    lib = _native.lib()
    sp = int(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(1)

    def run(keys, vals, bits):
        k0, v0 = keys.clone(), vals.clone()
        k1, v1 = torch.empty_like(keys), torch.empty_like(vals)
        in_alt = ctypes.c_int(0)
        vb = 4 if vals.dtype == torch.int32 else 8
        _native.check(lib.gk_radix_sort_pairs(k0.data_ptr(), k1.data_ptr(), v0.data_ptr(), v1.data_ptr(), vb,
                                              len(keys), 0, bits, ctypes.byref(in_alt), sp))
        torch.cuda.synchronize()
        return (k1, v1) if in_alt.value else (k0, v0)

    for value in values:
        os.environ[env_name] = value
        ok = True
        for m, bits, vdt in ((1, 64, torch.int32), (5000, 64, torch.int32), (8192 * 3, 16, torch.int32),
                             (10_000_019, 64, torch.int32), (10_000_019, 24, torch.int32),
                             (3_000_001, 64, torch.int64), (40_000_003, 8, torch.int32)):
            hi = 1 << (bits - 1) if bits < 64 else 1 << 62
            keys = torch.randint(0, hi, (m,), dtype=torch.int64, device="cuda", generator=g)
            if bits == 24:
                keys = keys & 0xFF00FF          # few distinct digits: long runs inside a tile
            vals = torch.arange(m, dtype=vdt, device="cuda")
            gk, gv = run(keys, vals, bits)
            ek, perm = torch.sort(keys, stable=True)
            good = bool(torch.equal(gk, ek)) and bool(torch.equal(gv.long(), perm))
            if not good:
                print(f"{env_name}={value} MISMATCH at m={m} bits={bits} vals={vdt}", flush=True)
            ok = ok and good
        keys = torch.randint(-(1 << 62), 1 << 62, (n,), dtype=torch.int64, device="cuda", generator=g)
        vals = torch.arange(n, dtype=torch.int32, device="cuda")
        k0, v0 = keys.clone(), vals.clone()
        k1, v1 = torch.empty_like(keys), torch.empty_like(vals)
        in_alt = ctypes.c_int(0)
        times = []
        for it in range(6):
            k0.copy_(keys); v0.copy_(vals)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _native.check(lib.gk_radix_sort_pairs(k0.data_ptr(), k1.data_ptr(), v0.data_ptr(), v1.data_ptr(), 4, n,
                                                  0, 64, ctypes.byref(in_alt), sp))
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        best = min(times[2:])
        print(f"{env_name}={value} n={n} sort_ms={best:.3f} per_pass_ms={best / 8:.3f} "
              f"GBps_per_pass={24 * n / (best / 8) / 1e6:.0f} (incl. histogram) parity={ok}", flush=True)
        del keys, vals, k0, v0, k1, v1


def keys32(n, cfgs):
    """gk_radix_sort_pairs32 (u32 key, u32 value: 8-byte pairs), 4 passes over 32 bits, every tile shape."""
    import torch
    from genome_kmers import _native

    lib = _native.lib()
    sp = int(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(1)
    keys = torch.randint(-(1 << 31), 1 << 31, (n,), dtype=torch.int32, device="cuda", generator=g)
    vals = torch.arange(n, dtype=torch.int32, device="cuda")
    k0, v0 = keys.clone(), vals.clone()
    k1, v1 = torch.empty_like(keys), torch.empty_like(vals)
    in_alt = ctypes.c_int(0)
    for cfg in cfgs:
        os.environ["GK_SORT32_CFG"] = cfg
        times = []
        for it in range(6):
            k0.copy_(keys); v0.copy_(vals)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _native.check(lib.gk_radix_sort_pairs32(k0.data_ptr(), k1.data_ptr(), v0.data_ptr(), v1.data_ptr(), 4, n,
                                                    0, 32, ctypes.byref(in_alt), sp))
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        out = k1 if in_alt.value else k0
        u = out.long() & 0xFFFFFFFF
        ok = bool((u[1:] >= u[:-1]).all())
        best = min(times[2:])
        print(f"keys32 cfg={cfg} n={n} sort_ms={best:.3f} per_pass_ms={best / 4:.3f} "
              f"GBps_per_pass={16 * n / (best / 4) / 1e6:.0f} (incl. histogram) sorted={ok}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--keys32":
        # python tools/bench_sort.py --keys32 200000000 0 1 2 3
        keys32(int(sys.argv[2]), sys.argv[3:] or ["0", "1", "2", "3"])
    elif len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    elif len(sys.argv) > 1 and sys.argv[1] == "--env":
        # python tools/bench_sort.py --env GK_SORT_CFG 200000000 7 6 7 6
        variants(int(sys.argv[3]), sys.argv[2], sys.argv[4:])
    else:
        n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
        cfgs = sys.argv[2:] or ["0", "1", "2", "3", "4", "5"]
        for c in cfgs:
            env = dict(os.environ, GK_SORT_CFG=c)
            subprocess.run([sys.executable, __file__, "--child", str(n)], env=env)
