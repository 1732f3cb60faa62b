#!/usr/bin/env python3
"""Short single-GPU run for ncu: the cfg2 grid (512^3 gyroid), a few isosurfaces
through the device pipeline.  usage: profile_run.py [n_side] [n_iso] [kind]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tools"))
import numpy as np
import torch

import workloads
from mc33_c_library_b200 import _cabi as cabi
from mc33_c_library_b200.device import Extractor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n_iso = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kind = sys.argv[3] if len(sys.argv) > 3 else "gyroid"
dev = torch.device("cuda", 0)
if kind == "gyroid":
    grid = workloads.make("cfg2", n=n).device_slab(0, n, dev)
    isos = [0.0, -0.9, 0.6, -1.2][:n_iso]
elif kind == "ct":
    # cfg3-like: u16 blobs + texture + noise 0..15, INTEGER isovalue (on-iso samples everywhere near the surface)
    g = torch.Generator(device=dev); g.manual_seed(5)
    ax = torch.linspace(-1, 1, n, device=dev)
    v = torch.full((n, n, n), 1000.0, device=dev)
    for c0, c1, c2, sg in ((0.1, -0.2, 0.05, 0.3), (-0.4, 0.3, -0.1, 0.25), (0.5, 0.5, 0.1, 0.2), (-0.3, -0.5, 0.6, 0.35)):
        v += 2500.0 / 3 * torch.exp(-((ax[None, None, :] - c0) ** 2 + (ax[None, :, None] - c1) ** 2 + (ax[:, None, None] - c2) ** 2) / (2 * sg * sg))
    v += 20.0 * torch.sin(37 * ax[None, None, :]) * torch.sin(29 * ax[None, :, None]) * torch.sin(31 * ax[:, None, None])
    v += torch.randint(0, 16, v.shape, device=dev, generator=g)
    grid = v.clamp(0, 65535).to(torch.int32).to(torch.uint16)
    del v
    isos = [1500.0, 1500.5][:n_iso]
else:
    from support import noise_grid
    grid = torch.from_numpy(noise_grid(n, "f32")).to(dev)
    isos = [0.0, 0.1][:n_iso]
ex = Extractor(cabi.make_desc(cabi.U16 if kind == "ct" else cabi.F32, n - 1, n - 1, n - 1), 0)
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
