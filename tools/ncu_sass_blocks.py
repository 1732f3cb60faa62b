#!/usr/bin/env python3
"""Hot SASS blocks of one kernel in an .ncu-rep: consecutive instructions with the same
execution count, with their opcode mix and the source lines they come from.
usage: ncu_sass_blocks.py report.ncu-rep kernel-regex [min share %]"""
import csv, io, subprocess, sys
from collections import Counter
rep, kern = sys.argv[1], sys.argv[2]
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
data = []
ninst = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        ninst += 1
        if ninst > 1:
            break
        continue
    if r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr and len(r) > 5:
        try:
            n = int(r[hdr["Instructions Executed"]])
        except ValueError:
            continue
        data.append((r[hdr["Address"]], r[hdr["Source"]].strip(), n, int(r[hdr["# Samples"]] or 0)))
tot = sum(d[2] for d in data)
print("total warp-instr", tot, "static", len(data))
seg, cur = [], None
for a, s, n, sm in data:
    op = s.split()[1] if s.startswith("@") else s.split()[0]
    if cur and abs(n - cur["n"]) <= max(2, 0.02 * cur["n"]):
        cur["k"] += 1; cur["tot"] += n; cur["samp"] += sm; cur["ops"].append(op)
    else:
        cur = {"a": a, "n": n, "k": 1, "tot": n, "samp": sm, "ops": [op]}
        seg.append(cur)
for s in seg:
    if s["tot"] >= tot * minshare / 100:
        c = Counter(o.split(".")[0] for o in s["ops"]).most_common(7)
        print(f"{s['a'][-5:]} n={s['n']:>8} len={s['k']:>4} tot={s['tot']/1e6:6.2f}M ({100*s['tot']/tot:4.1f}%) samp={s['samp']:>5} {c}")
