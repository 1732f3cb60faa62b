#!/usr/bin/env python3
"""Experiment: the isovalues of a sweep on several CUDA streams of ONE context (the sets of a sweep are independent
once the sweep classify has run; the context keeps one vertex-task buffer per stream).  Prints the time per
isosurface on 1 .. N streams and checks the meshes of the overlapped run against the single-stream run.

    usage: time_overlap.py [workload=cfg2] [max_streams=3] [steps=5]
    env:   MC33_B200_EMC_PER_SM / MC33_B200_EMV_PER_SM size the persistent grids (smaller grids leave room for
           the other stream's kernels on the same SM)
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import torch

import bench
import workloads

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
max_streams = int(sys.argv[2]) if len(sys.argv) > 2 else 3
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
W = workloads.make(name)
assert W.sweep, "a sweep workload"
torch.cuda.set_device(0)
rig = bench.Rig(W, 0, 1, 0)
ex, isos, main = rig.ex, rig.isos, rig.stream
n = len(isos)


def run(streams, bufs):
    ex.use_stream(main)
    ex.classify_sweep(isos)
    ev = main.record_event()
    for s in streams:
        s.wait_event(ev)
    for j in range(n):
        k = j % len(streams)
        ex.use_stream(streams[k])
        ex.extract_set_async(j, bufs[k])
    for s in streams:
        main.wait_event(s.record_event())
    ex.use_stream(main)


def timed(streams, bufs):
    for _ in range(3):
        run(streams, bufs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(steps):
        run(streams, bufs)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


# reference meshes of the last set of each residue class, from the single-stream schedule
out = {"workload": name, "ms_per_iso": {}}
out["ms_per_iso"]["1 (main stream)"] = round(timed([main], [rig.buf]) / n, 4)
ref = {}
for ns in range(2, max_streams + 1):
    streams = [torch.cuda.Stream(device=rig.dev) for _ in range(ns)]
    bufs = [ex.alloc(rig.capV, rig.capT) for _ in range(ns)]
    out["ms_per_iso"][str(ns)] = round(timed(streams, bufs) / n, 4)
    # the last set each stream handled is still in its buffer: the same set alone on the main stream must give the same mesh
    ok = True
    for k in range(ns):
        j = max(i for i in range(n) if i % ns == k)
        nV, nT = int(rig.cnt[j].nV), int(rig.cnt[j].nT)
        got = [bufs[k]["V"][:nV].clone(), bufs[k]["N"][:nV].clone(), bufs[k]["T"][:nT].clone()]
        ex.use_stream(main)
        ex.classify_sweep(isos)
        ex.extract_set_async(j, rig.buf)
        torch.cuda.synchronize()
        want = [rig.buf["V"][:nV], rig.buf["N"][:nV], rig.buf["T"][:nT]]
        ok = ok and all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(got, want))
    out.setdefault("meshes_equal", {})[str(ns)] = bool(ok)
    del streams, bufs
print(json.dumps(out))
rig.close()
