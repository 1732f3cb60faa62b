#!/usr/bin/env python3
"""cfg2 sweep with TWO extractor contexts alternating isovalues on two streams (the sweep's
isosurfaces are independent): does classify of one overlap emit of the other?
usage: time_dual.py [reps]"""
import json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
sys.path.insert(0, str(ROOT / 'tools'))
import workloads
from mc33_c_library_b200 import _cabi as cabi
from mc33_c_library_b200.device import Extractor

n = 512
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
W = workloads.make('cfg2')
grid = W.device_slab(0, n, dev)
isos = list(W.isos)
nstreams = int(os.environ.get('DUAL_STREAMS', '2'))
exs = [Extractor(cabi.make_desc(cabi.F32, n - 1, n - 1, n - 1), 0) for _ in range(nstreams)]
streams = [torch.cuda.Stream() for _ in range(nstreams)]
bufs = []
for ex, s in zip(exs, streams):
    ex.bind(grid)
    with torch.cuda.stream(s):
        ex.use_stream(s)
        ks = [ex.count(i) for i in isos]
        bufs.append(ex.alloc(max(int(c.nV) for c in ks) + 16, max(int(c.nT) for c in ks) + 16))
torch.cuda.synchronize()

def sweep():
    for j, iso in enumerate(isos):
        exs[j % nstreams].extract_async(iso, bufs[j % nstreams])

for _ in range(2):
    sweep()
torch.cuda.synchronize()
main = torch.cuda.current_stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(main)
for s in streams:
    s.wait_event(e0)
for _ in range(reps):
    sweep()
evs = []
for s in streams:
    e = torch.cuda.Event(); e.record(s); evs.append(e)
for e in evs:
    main.wait_event(e)
e1.record(main)
torch.cuda.synchronize()
for ex in exs:
    ex.sync()
env = {k: v for k, v in os.environ.items() if k.startswith("MC33_B200_")}
print(json.dumps({"env": env, "streams": nstreams, "dual_ms_per_iso": round(e0.elapsed_time(e1) / (reps * len(isos)), 4)}))
