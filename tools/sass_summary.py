#!/usr/bin/env python
"""SASS evidence for the hot kernels: per kernel the resource usage (registers, shared memory, spills) and an opcode
histogram of the sm_100a code in libfdc_b200.so (cuobjdump; runs without a GPU).

    python tools/sass_summary.py [kernel-name-substring ...] > profiles/r1_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gr-fdc_b200", "lib", "libfdc_b200.so")
DEFAULT = ["k_fwd_cols<256, 256, 16, true, 16>", "k_fwd_rows<256, 256, 16, true, 16>", "k_extract32<512, 8>", "k_fwd_small32<8192, 1>",
           "k_extract8<128, 16, true>", "k_extract<256, 16, true>", "k_jobs<512", "k_edges", "k_group_power", "k_band_power"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    wanted = sys.argv[1:] or DEFAULT
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\n\s*(REG:\d+[^\n]*)", res):
        usage[m.group(1)] = m.group(2).strip()
    dm = demangle(list(usage))
    pick = [k for k in usage if any(w in dm[k] for w in wanted)]
    print("# %s -- sm_100a SASS, %d kernels in the library, %d shown" % (os.path.relpath(LIB, ROOT), len(usage), len(pick)))
    for k in sorted(pick, key=lambda k: dm[k]):
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", k, LIB], capture_output=True, text=True).stdout
        ops = collections.Counter()
        for line in sass.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m:
                ops[m.group(1)] += 1
        total = sum(ops.values())
        fam = collections.Counter()
        for o, c in ops.items():
            fam[o.split(".")[0]] += c
        print("\n## %s\n   %s\n   %d instructions" % (dm[k].replace("fdc::", "")[:150], usage[k], total))
        print("   by family: " + ", ".join("%s %d" % (o, c) for o, c in fam.most_common(14)))
        mem = [(o, c) for o, c in ops.most_common() if o.split(".")[0] in ("LDG", "STG", "LDS", "STS", "LDC", "LDL", "STL", "UBLKCP", "SYNCS", "BAR", "ACQBULK", "UTMALDG")]
        print("   memory / sync ops: " + ", ".join("%s %d" % (o, c) for o, c in mem))


if __name__ == "__main__":
    main()
