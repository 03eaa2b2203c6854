#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that identify Blackwell-native code paths in libgsage_sm100.so
(B200_PROFILING.md "What proves a Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
UTMALDG = TMA tensor loads, SYNCS = mbarrier ops, LDG.E.128 / RED.E.ADD.F32x4-style = 128-bit loads and vector
reductions, .SYS-scoped loads/stores = peer (NVLink) traffic.  Runs without a GPU (cuobjdump on the in-tree .so).

    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "graphsage-simple_b200", "graphsage", "lib", "libgsage_sm100.so")
PATTERNS = [("tcgen05.mma", r"\bUTC\w*MMA\b"), ("tcgen05.st (A operand -> TMEM)", r"\bSTTM\b"), ("tcgen05.ld (epilogue)", r"\bLDTM\b"),
            ("tcgen05.commit", r"\bUTCBAR\b"), ("TMA tensor load", r"\bUTMALDG\b"), ("mbarrier", r"\bSYNCS\b"),
            ("elect.sync", r"\bELECT\b"), ("cp.async (LDGSTS: rows gathered into the stage ring)", r"\bLDGSTS\b"), ("128-bit global load", r"\bLDG\.E\.(?:\w+\.)*128\b"),
            ("128-bit shared load", r"\bLDS\.128\b"), ("vector reduction red.v4.f32", r"\bRED\.E\.ADD\.F32x4\b|\bREDG\.E\.ADD\.F32x4\b|\bRED\.E\.ADD\.\w*F32\w*\.V?4|RED\S*128"),
            ("system-scope ld/st (.SYS: peer memory, or volatile ticket reads)", r"\b(?:LDG|STG|LD|ST)\.E\.\S*SYS\b"), ("legacy HMMA (must be 0)", r"\bHMMA\b")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    print("# SASS evidence for %s (sm_100a)" % os.path.relpath(LIB, ROOT))
    print("# kernel: count of each mnemonic class (only non-zero classes shown)\n")
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        dem = re.sub(r"\(anonymous namespace\)::", "", dem)
        dem = dem.split("(")[0][:70]
        counts = collections.OrderedDict()
        for label, pat in PATTERNS:
            n = len(re.findall(pat, f))
            if n:
                counts[label] = n
        line = ", ".join("%s x%d" % kv for kv in counts.items()) or "-"
        print("%-72s %s" % (dem, line))


if __name__ == "__main__":
    sys.exit(main())
