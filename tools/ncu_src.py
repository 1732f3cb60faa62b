#!/usr/bin/env python3
"""Instructions executed per SOURCE LINE of one kernel, in file order, with the source text the report embeds
(needs -lineinfo and ncu --import-source on).   usage: ncu_src.py report.ncu-rep kernel-substring [file-substring] [min-instr]"""
import collections
import csv
import io
import subprocess
import sys

rep, kn = sys.argv[1], sys.argv[2]
fsel = sys.argv[3] if len(sys.argv) > 3 else "pipeline"
minn = float(sys.argv[4]) if len(sys.argv) > 4 else 2e5
txt = {}
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
blocks, cur, fpath = [], None, ""
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "File Path":
        fpath = row[1].split("/")[-1]
        continue
    if row[0] == "Function Name":
        cur = {"fn": row[1], "file": fpath, "hdr": None, "rows": []}
        blocks.append(cur)
        continue
    if cur is None:
        continue
    if cur["hdr"] is None:
        cur["hdr"] = row
    else:
        cur["rows"].append(row)
agg, smp = collections.Counter(), collections.Counter()
for b in blocks:
    if kn not in b["fn"]:
        continue
    h = b["hdr"]
    iI, iS = h.index("Instructions Executed"), h.index("# Samples")
    for r in b["rows"]:
        try:
            ln, n, s = int(r[0]), int(r[iI] or 0), int(r[iS] or 0)
        except ValueError:
            continue
        agg[(b["file"], ln)] += n
        txt[(b["file"], ln)] = r[1]
        smp[(b["file"], ln)] += s
tot = sum(agg.values())
print(f"{kn}: {tot / 1e6:.1f} M warp instructions (all captured launches)")
byfile = collections.Counter()
for (f, ln), n in agg.items():
    byfile[f] += n
print({f: round(n / 1e6, 1) for f, n in byfile.most_common()})
for (f, ln) in sorted(agg):
    if fsel in f and agg[(f, ln)] >= minn:
        print(f"{f[:18]:18s}{ln:5d} {agg[(f, ln)] / 1e6:7.2f}M {100 * agg[(f, ln)] / tot:5.1f}% s={smp[(f, ln)]:5d} | {txt.get((f, ln), '?').strip()[:105]}")
