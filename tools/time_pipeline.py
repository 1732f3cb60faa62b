#!/usr/bin/env python3
"""Device-pipeline time of the cfg2 sweep (512^3 gyroid, 8 isovalues), grid resident:
ms per isosurface (events around `reps` sweeps) and the per-kernel split of one sweep.
Environment overrides (MC33_B200_*) are picked up by the library at context creation, so
this is the harness for A/B runs.  usage: time_pipeline.py [n_side] [reps] [tag]"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

import bench
from mc33_c_library_b200 import _cabi as cabi
from mc33_c_library_b200.device import Extractor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
tag = sys.argv[3] if len(sys.argv) > 3 else ""
dev = torch.device("cuda", 0)
grid = bench.gyroid_device(n, 0, n, n, dev)
isos = [float(v) for v in os.environ["MC33_TP_ISOS"].split(",")] if os.environ.get("MC33_TP_ISOS") else bench.ISOS
ex = Extractor(cabi.make_desc(cabi.F32, n - 1, n - 1, n - 1), 0)
ex.bind(grid)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ex.use_stream(s)
    ks = [ex.count(i) for i in isos]
    buf = ex.alloc(max(int(c.nV) for c in ks) + 16, max(int(c.nT) for c in ks) + 16)
    for _ in range(2):
        for i in isos:
            ex.extract_async(i, buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps):
        for i in isos:
            ex.extract_async(i, buf)
    e1.record(s)
    torch.cuda.synchronize()
    ex.sync()
    ms = e0.elapsed_time(e1) / (reps * len(isos))
    ex.timing(True)
    kt = [0.0] * 5
    per_iso = []
    for i in isos:
        ex.extract_async(i, buf)
        torch.cuda.synchronize()
        one = ex.kernel_times()
        per_iso.append([round(x, 4) for x in one])
        kt = [a + b / len(isos) for a, b in zip(kt, one)]
    ex.timing(False)
env = {k: v for k, v in os.environ.items() if k.startswith("MC33_B200_")}
print(json.dumps({"tag": tag, "env": env, "n": n, "ms_per_iso": round(ms, 4),
                  "serial_kernel_ms": dict(zip(["classify", "count", "rowscan", "emit_cells", "emit_vertices"], [round(x, 4) for x in kt])),
                  "per_iso_kernel_ms": per_iso, "nV": [int(c.nV) for c in ks], "nT": [int(c.nT) for c in ks]}))
