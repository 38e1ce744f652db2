#!/usr/bin/env python
"""SASS opcode histogram per kernel of libshopformer_b200.so (cuobjdump -sass): the mnemonics that prove what each kernel
runs on -- UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk = TMA 1-D), UTMALDG (tensor-map TMA),
SYNCS (mbarrier), HFMA2 / FFMA / FFMA2, F2FP (packed conversions), LDS / STS / LDG / STG.
    python profiles/sass_histogram.py > profiles/r2_sass_histogram.md"""
import collections, re, subprocess, sys
from pathlib import Path
so = Path(__file__).resolve().parent.parent / "computer-vision-shoplifting-detection_b200" / "shopformer_b200" / "libshopformer_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"^void ", "", kern)
        kern = re.sub(r"\((?!anonymous).*", "", kern).replace("sf::", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0]] += 1
keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "HFMA2", "FFMA", "FFMA2", "FADD2", "F2FP", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "MATCH", "DFMA"]
print("# SASS opcode histogram per kernel (round 2 build)\n")
print("`python profiles/sass_histogram.py` over `cuobjdump -sass libshopformer_b200.so` (static instruction counts).\n")
print("| kernel | total | " + " | ".join(keys) + " |")
print("|---|---|" + "---|" * len(keys))
for k, h in hist.items():
    if sum(h.values()) < 50:
        continue
    print(f"| `{k[:70]}` | {sum(h.values())} | " + " | ".join(str(h.get(x, 0)) for x in keys) + " |")
