#!/usr/bin/env python
"""Opcode histogram per kernel of libgdtb.so (cuobjdump -sass, no GPU needed): which kernels carry TMA bulk copies
(UBLKCP), mbarrier waits (SYNCS), FP64 FMAs (DFMA), FP64 tensor-core MMAs (DMMA), atomics (ATOM / RED), shuffles.

  python tools/sass_histogram.py [dune-gdt_b200/lib/libgdtb.so] > profiles/rNN_sass_histogram.txt
"""
import collections
import re
import subprocess
import sys

OPS = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "DMMA", "DFMA", "DMUL", "DADD", "MUFU", "ATOM", "ATOMG", "RED", "SHFL",
       "LDG", "STG", "LDS", "STS", "LDC", "BAR", "MEMBAR", "NANOSLEEP"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "dune-gdt_b200/lib/libgdtb.so"
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = collections.Counter()
            kernels[m.group(1)] = cur
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    print(f"# {lib}: SASS opcode counts per kernel (static instruction counts; cuobjdump -sass, sm_100a)")
    print("# " + " ".join(f"{o:>7s}" for o in ["total"] + OPS) + "  kernel")
    for (name, c), dn in zip(kernels.items(), demangled):
        total.update(c)
        dn = re.sub(r"(gdtb::)?\(anonymous namespace\)::", "", dn)
        dn = re.sub(r"^void ", "", dn)
        dn = re.sub(r"\((gdtb::|long|double|int|unsigned|float|void|char|bool).*", "", dn)
        print("  " + " ".join(f"{c[o]:7d}" for o in ["_total"] + OPS) + "  " + dn[:110])
    print("# whole library")
    print("  " + " ".join(f"{total[o]:7d}" for o in ["_total"] + OPS))


if __name__ == "__main__":
    main()
