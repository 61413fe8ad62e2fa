#!/usr/bin/env python
"""Where a sort pass spends its time (tuning aid): runs gk_radix_sort_pairs from the library built with
`make -C genome-kmers_b200/csrc probe` (-DGK_PROBE) and prints, per tile, the cycles thread 0 spends in each
phase of onesweep_kernel and how far bin 0's look-back walks.
    python tools/probe_onesweep.py [n]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "genome-kmers_b200"))


def main():
    import numpy as np
    import torch
    from genome_kmers import _native

    _native.LIB_PATH = os.path.join(ROOT, "genome-kmers_b200", "lib", "libgkb200_probe.so")
    lib = _native.lib()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 199_999_400
    g = torch.Generator(device="cuda").manual_seed(1)
    keys = torch.randint(-(1 << 62), 1 << 62, (n,), dtype=torch.int64, device="cuda", generator=g)
    vals = torch.arange(n, dtype=torch.int32, device="cuda")
    k0, v0 = keys.clone(), vals.clone()
    k1, v1 = torch.empty_like(keys), torch.empty_like(vals)
    sp = int(torch.cuda.current_stream().cuda_stream)
    in_alt = ctypes.c_int(0)
    probe = np.zeros(16, dtype=np.uint64)
    lib.gk_probe_read.argtypes = [ctypes.c_void_p]
    for it in range(4):
        k0.copy_(keys); v0.copy_(vals)
        torch.cuda.synchronize()
        lib.gk_probe_read(probe.ctypes.data)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _native.check(lib.gk_radix_sort_pairs(k0.data_ptr(), k1.data_ptr(), v0.data_ptr(), v1.data_ptr(), 4, n,
                                              32, 64, ctypes.byref(in_alt), sp))
        e1.record()
        torch.cuda.synchronize()
        lib.gk_probe_read(probe.ctypes.data)
        tiles = max(int(probe[10]), 1)
        names = ["A load+count", "B publish+scan", "C rank+scatter", "D look-back (bin 0)", "D wait for other bins",
                 "E stream out"]
        cyc = [int(probe[i]) / tiles for i in range(6)]
        print(f"run {it}: 4 passes {e0.elapsed_time(e1):.3f} ms, tiles {tiles}, cycles per tile: "
              + ", ".join(f"{nm} {c:.0f}" for nm, c in zip(names, cyc))
              + f" | sum {sum(cyc):.0f}; look-back walk avg {int(probe[8]) / tiles:.1f} max {int(probe[11])} "
              f"status words, unpublished polls avg {int(probe[9]) / tiles:.2f}", flush=True)


if __name__ == "__main__":
    main()
