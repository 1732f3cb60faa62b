#!/usr/bin/env python3
"""Per-kernel device times of ONE z-slab of a multi-GPU workload on one GPU (no NCCL): what rank `rank` of `world`
would run.  usage: time_slab.py [workload=cfg4] [world=2] [rank=0] [reps=2]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import torch

import bench
import workloads

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
W = workloads.make(name)
torch.cuda.set_device(0)
rig = bench.Rig(W, rank, world, 0)
kt, _ = rig.kernel_times(reps)
print(json.dumps({"workload": name, "world": world, "rank": rank, "slab": [rig.sl.z_lo, rig.sl.z_hi],
                  "kernel_ms": {k: round(float(v), 4) for k, v in zip(bench.KNAMES, kt)},
                  "nV": int(rig.cnt[0].nV), "nT": int(rig.cnt[0].nT)}))
rig.close()
