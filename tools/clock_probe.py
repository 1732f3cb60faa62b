#!/usr/bin/env python3
"""Is the device pipeline clock-limited?  Times (a) the emit phase alone, repeated, (b) the
whole sweep, while sampling nvidia-smi clocks / power in a side thread."""
import json, subprocess, sys, threading, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from mc33_c_library_b200 import _cabi as cabi
from mc33_c_library_b200.device import Extractor

n = 512
dev = torch.device("cuda", 0)
grid = bench.gyroid_device(n, 0, n, n, dev)
ex = Extractor(cabi.make_desc(cabi.F32, n - 1, n - 1, n - 1), 0)
ex.bind(grid)
samples, stop = [], False
def sampler():
    while not stop:
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active",
                            "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        samples.append((time.time(), o))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ex.use_stream(s)
    k = ex.count(0.0)
    buf = ex.alloc(int(k.nV) + 16, int(k.nT) + 16)
    ex.extract_async(0.0, buf); torch.cuda.synchronize()
    th = threading.Thread(target=sampler); th.start()
    out = {}
    for name, reps, fn in (("emit_only_x200", 200, lambda: ex.emit(buf)), ("extract_x200", 200, lambda: ex.extract_async(0.0, buf)),
                           ("emit_only_x5", 5, lambda: ex.emit(buf))):
        torch.cuda.synchronize(); time.sleep(1.0)
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            fn()
        e1.record(s); torch.cuda.synchronize()
        out[name] = {"ms": e0.elapsed_time(e1) / reps, "t0": t0, "t1": time.time()}
    stop = True; th.join()
for name, d in out.items():
    d["smi"] = [o for t, o in samples if d["t0"] <= t <= d["t1"]][:6]
print(json.dumps(out, indent=1))
print("idle samples:", [o for t, o in samples][:2])
