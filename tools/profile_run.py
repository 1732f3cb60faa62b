#!/usr/bin/env python3
"""Short single-GPU run for ncu: the cfg2 grid (512^3 gyroid), a few isosurfaces
through the device pipeline.  usage: profile_run.py [n_side] [n_iso] [kind]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch

import bench
from mc33_c_library_b200 import _cabi as cabi
from mc33_c_library_b200.device import Extractor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n_iso = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kind = sys.argv[3] if len(sys.argv) > 3 else "gyroid"
dev = torch.device("cuda", 0)
if kind == "gyroid":
    grid = bench.gyroid_device(n, 0, n, n, dev)
    isos = [0.0, -0.9, 0.6, -1.2][:n_iso]
else:
    from support import noise_grid
    grid = torch.from_numpy(noise_grid(n, "f32")).to(dev)
    isos = [0.0, 0.1][:n_iso]
ex = Extractor(cabi.make_desc(cabi.F32, n - 1, n - 1, n - 1), 0)
ex.bind(grid)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ex.use_stream(s)
    k = [ex.count(i) for i in isos]
    buf = ex.alloc(max(int(c.nV) for c in k) + 16, max(int(c.nT) for c in k) + 16)
    for i in isos:
        ex.extract_async(i, buf)
    torch.cuda.synchronize()
    c = ex.sync()
print("ok", int(c.nV), int(c.nT))
