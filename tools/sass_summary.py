#!/usr/bin/env python3
"""SASS mnemonic counts per kernel of the shipped library (profiles/r2_sass_summary.txt):
    python tools/sass_summary.py > profiles/r2_sass_summary.txt
Proves which kernels are TMA / mbarrier / cp.async / packed-float32 code (B200_PROFILING.md lists the mnemonics)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "wicca_b200" / "lib" / "libwicca_b200.so"
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "UTMACMDFLUSH", "SYNCS", "LDGSTS", "FFMA2", "FADD2", "PRMT", "LDG", "STG", "LDS", "STS",
         "ATOMS", "SHFL", "BAR", "REDUX"]
out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
kern, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("wicca::", "")
        kern = re.sub(r"^void ", "", re.sub(r"\(.*", "", kern))
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[kern][w] += 1
print(f"# cuobjdump -sass {LIB.name} (sm_100a): instructions per kernel and the mnemonics that matter")
print(f"# {'kernel':58s} {'instr':>6s}  " + " ".join(f"{w:>7s}" for w in WATCH))
for k, c in counts.items():
    print(f"{k[:60]:60s} {total[k]:6d}  " + " ".join(f"{c.get(w, 0):7d}" for w in WATCH))
