#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (tcgen05 MMA, TMEM loads / stores, TMA
loads / stores / reduce-adds) in the built libvla_b200.so.  python scripts/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vla_adapter_b200", "lib", "libvla_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2",
             "FFMA2", "LDL", "STL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, kernel = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kernel = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kernel = re.sub(r"\(anonymous namespace\)::", "", kernel).split("(")[0].replace("void vla::", "")
            counts[kernel] = collections.Counter()
            continue
        if kernel is None:
            continue
        for mn in MNEMONICS:
            if re.search(r"\b" + re.escape(mn) + r"\b", line):
                counts[kernel][mn] += 1
    print(f"SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)")
    print("UTCHMMA = tcgen05.mma (bf16), LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG / UTMAREDG = TMA load /")
    print("store / reduce-add, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = mma.sync, LDL / STL = local-memory spills\n")
    print(f"{'kernel':58s} " + " ".join(f"{m:>8s}" for m in MNEMONICS))
    tot = collections.Counter()
    for k, c in counts.items():
        if not any(c.values()):
            continue
        print(f"{k[:58]:58s} " + " ".join(f"{c[m]:8d}" for m in MNEMONICS))
        tot.update(c)
    print(f"{'total':58s} " + " ".join(f"{tot[m]:8d}" for m in MNEMONICS))
    out = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
    libs = sorted({l.split()[0] for l in out.splitlines() if l.strip()})
    print("\nldd: " + ", ".join(libs))
    print("(no cuBLAS / cuDNN / NCCL / torch: every kernel on the path is in this library)")


if __name__ == "__main__":
    sys.exit(main())
