#!/usr/bin/env python3
"""classify kernel alone, back to back (bandwidth check): classify_bw.py [n] [dtype]"""
import ctypes as C, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import bench
from mc33_c_library_b200 import _cabi as cabi
from mc33_c_library_b200.device import Extractor
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
grid = bench.gyroid_device(n, 0, n, n, dev)
ex = Extractor(cabi.make_desc(cabi.F32, n - 1, n - 1, n - 1), 0)
ex.bind(grid)
s = torch.cuda.Stream()
lib = cabi.load()
lib.mc33cu_debug_classify.argtypes = [C.c_void_p, C.c_double]
with torch.cuda.stream(s):
    ex.use_stream(s)
    for _ in range(3):
        lib.mc33cu_debug_classify(ex.h, 0.1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(20):
        lib.mc33cu_debug_classify(ex.h, 0.1)
    e1.record(s)
    torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1000 / 20
print(f"classify alone: {us:.1f} us per launch, {grid.numel()*4/us*1e-3:.0f} GB/s")
