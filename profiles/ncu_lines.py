#!/usr/bin/env python
"""Top CUDA source lines of an .ncu-rep by stall samples / instructions: python profiles/ncu_lines.py <rep> [n]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur, agg = None, []
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] == "Line No" or r[2] != "-":
        continue
    try:
        agg.append((cur, int(r[0]), r[1].strip()[:100], int(r[6]), int(r[7])))
    except ValueError:
        pass
ts, ti = sum(a[3] for a in agg) or 1, sum(a[4] for a in agg) or 1
print(f"total samples {ts}  warp-instructions {ti}")
for a in sorted(agg, key=lambda a: -a[3])[:n]:
    print(f"{a[0][:20]:20s} {a[1]:4d}  smp {100*a[3]/ts:5.1f}%  ins {100*a[4]/ti:5.1f}%  {a[2]}")
