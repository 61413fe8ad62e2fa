#!/usr/bin/env python
"""Per-kernel SASS summary of libgkb200.so (cuobjdump -sass): instruction count, registers are in -Xptxas -v;
here the mnemonics that say how a kernel moves data -- bulk copies through the copy engine (UBLKCP = cp.async.bulk,
SYNCS = mbarrier), vector width of global loads/stores, shared-memory atomics, warp votes.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "genome-kmers_b200", "lib", "libgkb200.so")
WATCH = ["UBLKCP", "SYNCS", "UTMALDG", "UTMASTG", "LDG.E.128", "LDG.E.64", "LDG.E", "STG.E.128", "STG.E.64", "STG.E",
         "LDGSTS", "LDS", "STS", "ATOMS", "RED", "ATOMG", "VOTE", "SHFL", "MATCH", "BAR", "PRMT", "SHF", "POPC"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("gk::", "")
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            kernels[name]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[name][w] += 1
                    break
    arch = re.search(r"arch = (sm_\w+)", out)
    print(f"libgkb200.so: {len(kernels)} kernels, {arch.group(1) if arch else '?'} cubins (cuobjdump -sass)")
    print("kernels that use the copy engine (UBLKCP = cp.async.bulk, SYNCS = mbarrier):",
          sorted({k.split('<')[0] for k, c in kernels.items() if c['UBLKCP']}) or "none")
    print()
    seen = set()
    for k, c in kernels.items():
        base = k.split("<")[0]
        if base in seen:
            continue            # one instantiation per kernel template
        seen.add(base)
        parts = [f"{w}={c[w]}" for w in WATCH if c[w]]
        print(f"{k[:100]}\n    {c['_total']} instructions: " + " ".join(parts))


if __name__ == "__main__":
    main()
