#!/usr/bin/env python3
"""Device-pipeline time of one bench workload on one GPU, with the per-kernel split (no CPU legs, no e2e):
the quick A/B companion of bench.py.   usage: time_workload.py cfgN [steps]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import torch

import bench
import workloads

name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
W = workloads.make(name)
torch.cuda.set_device(0)
rig = bench.Rig(W, 0, 1, 0)
ms, launches, _ = rig.timed(steps, 3)
kt, single = rig.kernel_times()
per_iso = []
rig.ex.timing(True)
for iso in rig.isos:
    rig.ex.extract_async(iso, rig.buf)
    torch.cuda.synchronize()
    per_iso.append([round(float(v), 4) for v in rig.ex.kernel_times()])
rig.ex.timing(False)
print(json.dumps({"workload": name, "ms_per_step": round(ms, 4), "ms_per_iso": round(ms / len(rig.isos), 4), "launches": launches,
                  "kernel_ms": {k: round(float(v), 4) for k, v in zip(bench.KNAMES, kt)},
                  "per_iso_kernel_ms": per_iso if len(per_iso) <= 2 else None, "nV": [int(k.nV) for k in rig.cnt][:2]}))
rig.close()
