#!/usr/bin/env python3
"""Device-pipeline timings of the non-headline BASELINE configs (parity cases, not bench lines):
cfg1 201^3 cos-sum, cfg3 1024^3 u16 CT-like, cfg5 768^3 (or smaller) inclined white noise.
usage: time_configs.py [cfg1|cfg3|cfg5 ...]   -> one JSON line per config"""
import json, math, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
from mc33_c_library_b200 import _cabi as cabi
from mc33_c_library_b200.device import Extractor
from support import cfg1_grid, inclined_geom, make_desc

dev = torch.device("cuda", 0)
PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def run(name, ex, isos, sample_bytes, npts, real_bytes=4, reps=5):
    stream = torch.cuda.Stream()
    out = []
    with torch.cuda.stream(stream):
        ex.use_stream(stream)
        for iso in isos:
            k = ex.count(iso)
            nV, nT = int(k.nV), int(k.nT)
            buf = ex.alloc(nV + 16, nT + 16)
            for _ in range(2):
                ex.extract_async(iso, buf)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                ex.extract_async(iso, buf)
            e1.record(stream)
            torch.cuda.synchronize()
            ex.sync()
            ms = e0.elapsed_time(e1) / reps
            ex.timing(True); ex.extract_async(iso, buf); torch.cuda.synchronize(); kt = ex.kernel_times(); ex.timing(False)
            B = npts * sample_bytes + nV * (3 * real_bytes + 16) + nT * 12
            out.append(dict(iso=iso, nV=nV, nT=nT, ms=ms, gvox_s=npts / ms * 1e-6, mtri_s=nT / ms * 1e-3,
                            algorithmic_MB=B / 1e6, frac_of_hbm_peak=B / (ms * 1e-3) / 1e9 / PEAK,
                            kernel_ms=dict(zip(["classify", "count", "rowscan", "emit_cells", "emit_vertices"], [round(x, 4) for x in kt]))))
            del buf
    print(json.dumps({"config": name, "results": out}))


which = sys.argv[1:] or ["cfg1", "cfg3", "cfg5"]
if "cfg1" in which:
    F, geom = cfg1_grid()
    ex = Extractor(make_desc(F.shape, "f32", geom)); ex.bind(torch.from_numpy(F).to(dev))
    run("cfg1 201^3 f32 cos x+cos y+cos z (spnA)", ex, [0.0, 0.5], 4, F.size); ex.close()
if "cfg3" in which:
    n = 1024
    g = torch.Generator(device=dev); g.manual_seed(5)
    ax = torch.linspace(-1, 1, n, device=dev)
    vol = torch.empty((n, n, n), dtype=torch.uint16, device=dev)
    for z0 in range(0, n, 64):
        v = torch.full((64, n, n), 1000.0, device=dev)
        az = ax[z0:z0 + 64]
        for c0, c1, c2, sg in ((0.1, -0.2, 0.05, 0.3), (-0.4, 0.3, -0.1, 0.25), (0.5, 0.5, 0.1, 0.2), (-0.3, -0.5, 0.6, 0.35)):
            v += 2500.0 / 3 * torch.exp(-((ax[None, None, :] - c0) ** 2 + (ax[None, :, None] - c1) ** 2 + (az[:, None, None] - c2) ** 2) / (2 * sg * sg))
        v += 20.0 * torch.sin(37 * ax[None, None, :]) * torch.sin(29 * ax[None, :, None]) * torch.sin(31 * az[:, None, None])
        v += torch.randint(0, 16, v.shape, device=dev, generator=g)
        vol[z0:z0 + 64] = v.clamp(0, 65535).to(torch.int32).to(torch.uint16)
        del v
    ex = Extractor(cabi.make_desc(cabi.U16, n - 1, n - 1, n - 1)); ex.bind(vol)
    run("cfg3 1024^3 u16 CT-like", ex, [1500.0, 1500.5], 2, n ** 3); ex.close(); del vol
if "cfg5" in which:
    n = int([a for a in sys.argv if a.startswith("n=")][0][2:]) if any(a.startswith("n=") for a in sys.argv) else 768
    g = torch.Generator(device=dev); g.manual_seed(7)
    vol = torch.rand((n, n, n), device=dev, generator=g) * 2 - 1
    geom = inclined_geom()
    ex = Extractor(make_desc(vol.shape, "f32", geom)); ex.bind(vol)
    run(f"cfg5 {n}^3 f32 white noise, inclined grid (spnC)", ex, [0.0], 4, n ** 3, reps=2); ex.close()
