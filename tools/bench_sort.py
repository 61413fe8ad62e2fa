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


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    else:
        n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
        cfgs = sys.argv[2:] or ["0", "1", "2", "3", "4", "5"]
        for c in cfgs:
            env = dict(os.environ, GK_SORT_CFG=c)
            subprocess.run([sys.executable, __file__, "--child", str(n)], env=env)
