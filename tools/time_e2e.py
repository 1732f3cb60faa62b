#!/usr/bin/env python3
"""End-to-end time of one bench workload through the drop-in C API (host buffers; bench.py's `e2e` leg alone), for
A/B runs of the upload / download schedule.   usage: [MC33_B200_CHUNKS=k ...] time_e2e.py cfgN [gpus] [steps]"""
import json
import os
import sys
from pathlib import Path
from types import SimpleNamespace

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import torch

import bench
import workloads

name = sys.argv[1]
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
W = workloads.make(name)
torch.cuda.set_device(0)
r = bench.e2e_dropin(W, SimpleNamespace(steps=steps), gpus, 0)
r["env"] = {k: v for k, v in os.environ.items() if k.startswith("MC33_B200_")}
print(json.dumps(r))
