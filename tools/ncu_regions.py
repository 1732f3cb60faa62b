#!/usr/bin/env python3
"""Aggregate an .ncu-rep source page by line ranges: usage ncu_regions.py rep kernel 'file:a-b:name' ..."""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    f, rng, name = a.split(":")
    lo, hi = rng.split("-")
    regions.append((f, int(lo), int(hi), name))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fpath, hdr, infn = "", None, False
agg, samp, other = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "File Path":
        fpath = row[1].split("/")[-1]
        continue
    if row[0] == "Function Name":
        infn = kern in row[1]
        hdr = None
        continue
    if hdr is None:
        hdr = row
        ci = {n: i for i, n in enumerate(hdr)}
        continue
    if not infn or row[0] == "":
        continue
    try:
        ln = int(row[0]); inst = int(float(row[ci["Instructions Executed"]])); sm = int(float(row[ci["# Samples"]]))
    except (ValueError, KeyError):
        continue
    tot += inst
    for f, a, b, name in regions:
        if f in fpath and a <= ln <= b:
            agg[name] += inst; samp[name] += sm
            break
    else:
        other[(fpath, ln)] += inst
print("total warp-instr", tot)
for k, v in agg.most_common():
    print(f"  {k:28s} {v/1e6:8.1f}M ({100*v/max(tot,1):4.1f}%)  samples {samp[k]}")
print("  other", round(sum(other.values()) / 1e6, 1), "M", other.most_common(6))
