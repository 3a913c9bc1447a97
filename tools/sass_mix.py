#!/usr/bin/env python3
"""SASS instruction mix of the kernels of libhc_b200.so (cuobjdump -sass): static counts of the mnemonics that the
design notes refer to, per kernel.  No GPU needed.
    python tools/sass_mix.py > profiles/r02_sass_mix.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "huffman-codec_b200", "libhc_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ("CREDUX", "VOTE", "SHFL", "IDP", "PRMT", "POPC", "FLO", "SHF", "LOP3", "LDS", "STS", "ATOMS", "LDG", "STG", "BAR", "BRA", "BSSY", "CCTL", "LDL", "STL")
kern, mix, total = None, {}, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.split("(")[0].strip()
        kern = name.replace("hcd::", "")
        mix[kern], total[kern] = collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        total[kern] += 1
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                mix[kern][w] += 1
                break
print("static SASS instruction counts per kernel (sm_100a), %s" % os.path.basename(lib))
print("%-28s %6s  %s" % ("kernel", "total", "  ".join("%5s" % w for w in WATCH)))
for k in sorted(mix, key=lambda k: -total[k]):
    print("%-28s %6d  %s" % (k[:28], total[k], "  ".join("%5d" % mix[k][w] for w in WATCH)))
