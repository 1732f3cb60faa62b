#!/usr/bin/env python3
"""Short run for ncu: one mc33cu_classify_sweep of the cfg2 grid."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import torch
import workloads
from mc33_c_library_b200 import _cabi as cabi
from mc33_c_library_b200.device import Extractor
W = workloads.make("cfg2")
n = 512
grid = W.device_slab(0, n, torch.device("cuda", 0))
ex = Extractor(cabi.make_desc(cabi.F32, n - 1, n - 1, n - 1), 0)
ex.bind(grid)
for _ in range(2):
    ex.classify_sweep(list(W.isos))
torch.cuda.synchronize()
print("ok")
