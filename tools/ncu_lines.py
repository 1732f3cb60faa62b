#!/usr/bin/env python3
"""Per-source-line hot spots from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep [kernel-substring] [top N]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
# the output is a sequence of blocks: File Path / Function Name / header / rows
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
        continue
    cur["rows"].append(row)
seen = set()
merged = {}
for b in blocks:
    if want not in b["fn"]:
        continue
    m = merged.setdefault(b["fn"], {"fn": b["fn"], "hdr": b["hdr"], "rows": []})
    for r in b["rows"]:
        if r and r[0] != "":
            m["rows"].append([b["file"] + ":" + r[0]] + r[1:])
for b in merged.values():
    h = b["hdr"]
    # first "Source" column = CUDA source text (when present), second = SASS
    ci = {n: i for i, n in enumerate(h)}
    i_line = ci.get("Line No")
    src_cols = [i for i, n in enumerate(h) if n == "Source"]
    i_inst = ci["Instructions Executed"]
    i_samp = ci["# Samples"]
    i_thr = ci["Thread Instructions Executed"]
    per = defaultdict(lambda: [0, 0, 0, ""])
    tot_i = tot_s = 0
    for r in b["rows"]:
        try:
            ln = r[i_line]
            inst = int(float(r[i_inst] or 0)); samp = int(float(r[i_samp] or 0)); thr = int(float(r[i_thr] or 0))
        except (ValueError, IndexError):
            continue
        key = ln
        per[key][0] += inst; per[key][1] += samp; per[key][2] += thr
        tot_i += inst; tot_s += samp
    print(f"=== {b['fn'][:70]}  warp-instr {tot_i:,}  samples {tot_s:,}")
    for ln, (inst, samp, thr, _) in sorted(per.items(), key=lambda kv: -kv[1][int(__import__('os').environ.get('SORTCOL','1'))])[:top]:
        print(f"   line {ln:>28s}  instr {inst:>12,} ({100.0*inst/max(tot_i,1):5.1f}%)  samples {samp:>8,} ({100.0*samp/max(tot_s,1):5.1f}%)  thr/instr {thr/max(inst,1):5.1f}")
