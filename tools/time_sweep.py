#!/usr/bin/env python3
"""cfg2 sweep through the one-pass sweep classify (mc33cu_classify_sweep + extract_set) against
eight separate extractions; events around `reps` sweeps.  usage: time_sweep.py [reps]"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from mc33_c_library_b200 import _cabi as cabi
from mc33_c_library_b200.device import Extractor

n = 512
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
grid = bench.gyroid_device(n, 0, n, n, dev)
isos = bench.ISOS
ex = Extractor(cabi.make_desc(cabi.F32, n - 1, n - 1, n - 1), 0)
ex.bind(grid)
s = torch.cuda.Stream()
res = {}
with torch.cuda.stream(s):
    ex.use_stream(s)
    ks = [ex.count(i) for i in isos]
    buf = ex.alloc(max(int(c.nV) for c in ks) + 16, max(int(c.nT) for c in ks) + 16)

    def single():
        for i in isos:
            ex.extract_async(i, buf)

    def sweep():
        ex.classify_sweep(isos)
        for j in range(len(isos)):
            ex.extract_set_async(j, buf)

    def classify_only():
        ex.classify_sweep(isos)

    for name, fn in (("single", single), ("sweep", sweep), ("classify_sweep_only", classify_only)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            fn()
        e1.record(s)
        torch.cuda.synchronize()
        ex.sync()
        res[name + "_ms_per_sweep"] = round(e0.elapsed_time(e1) / reps, 4)
    # counts of the last set must equal the single path's
    ex.classify_sweep(isos)
    ok = True
    for j in range(len(isos)):
        ex.extract_set_async(j, buf)
        k = ex.sync()
        ok = ok and (int(k.nV), int(k.nT)) == (int(ks[j].nV), int(ks[j].nT))
res["counts_match"] = ok
res["ms_per_iso_single"] = round(res["single_ms_per_sweep"] / len(isos), 4)
res["ms_per_iso_sweep"] = round(res["sweep_ms_per_sweep"] / len(isos), 4)
print(json.dumps(res))
