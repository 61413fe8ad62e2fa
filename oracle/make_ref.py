#!/usr/bin/env python
"""
Recipe for oracle/_ref: the UNMODIFIED reference package, vendored from where it lies
(/root/reference/src/genome_kmers, pure Python + numba) so that it can travel to the GPU box.

    python oracle/make_ref.py          # run by __graft_entry__.build() where /root/reference exists

oracle/_ref/ is git-ignored (the reference's sources never enter this repository's history) but not
gpurun-ignored.  bench.py --impl reference and bench.py's cpu_baseline leg import it when numba is
importable (kind "reference") and fall back to the C port oracle/gk_oracle.c otherwise (kind "port").
Nothing in the product imports it.
"""
import os
import shutil
import sys

SRC = "/root/reference/src/genome_kmers"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "genome_kmers")


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"make_ref: {SRC} does not exist here; keeping whatever oracle/_ref already holds")
        return 0
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__"))
    with open(os.path.join(HERE, "_ref", "README"), "w") as f:
        f.write("Verbatim copy of /root/reference/src/genome_kmers (mrperkett/genome-kmers 1.0.1), made by "
                "oracle/make_ref.py.\nNot part of this repository's history; used only as the CPU baseline.\n")
    print(f"make_ref: copied {SRC} -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
