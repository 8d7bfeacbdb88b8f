#!/usr/bin/env python
"""Execution-weighted SASS opcode histogram of an .ncu-rep (source page), optionally restricted to a file:line range.
usage: python tools/sass_hist.py rep.ncu-rep [file:lo-hi]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else None
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, h, line_ok, hist, tot = None, None, True, collections.Counter(), 0
if flt:
    ff, lh = flt.split(":"); lo, hi = map(int, lh.split("-"))
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 2 and r[0] == "Line No":
        h = r
    elif len(r) > 2 and h:
        if r[0] != "":
            line_ok = (not flt) or (cur == ff and lo <= int(r[0]) <= hi)
        elif line_ok and r[3] not in ("", "..."):
            try:
                n = int(r[h.index("Instructions Executed")])
            except ValueError:
                continue
            op = r[3].split()[0]
            if op.startswith("@"):
                op = r[3].split()[1]
            op = op.split(".")[0] + ("." + r[3].split()[0 if not r[3].split()[0].startswith("@") else 1].split(".")[1] if op.split(".")[0] in ("F2F", "MUFU", "LDS", "STS", "LDG", "STG") and "." in r[3].split()[0 if not r[3].split()[0].startswith("@") else 1] else "")
            hist[op] += n; tot += n
print("total", tot)
for op, n in hist.most_common(40):
    print(f"{op:14s} {n:12d} {100*n/tot:5.1f}%")
