#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 Marching Cubes 33 path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg5]

One STEP = one iso sweep (8 isosurfaces) over the grid, which stays resident in HBM
(BASELINE.json configs[1]: 512^3 float gyroid, 8 isovalues).  Per rank the grid is a
512x512x512-cell... (513^3 would not be the named shape: the grid has 512^3 SAMPLES).
N > 1: weak scaling -- every rank owns a 512-slice z-slab (plus halo slices) of one
512 x 512 x (512*N) gyroid; the only exchange is the on-device all-gather of the
per-slab counts that turns local vertex ids into global ones.

Prints ONE JSON line (see the contract in the task description / DESIGN.md).
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ISOS = [-1.2, -0.9, -0.6, -0.3, 0.0, 0.3, 0.6, 0.9]
N_SIDE = 512


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------
# synthetic grids (same generator on host and device: products of per-axis
# sin/cos tables computed once in float64 on the host)
# --------------------------------------------------------------------------
def gyroid_tables(n_xy, z0, z1, nz_total, periods_xy=4.0):
    import numpy as np
    w = 2.0 * math.pi * periods_xy
    t = (np.arange(n_xy, dtype=np.float64) / (n_xy - 1) - 0.5) * w
    # along z the same spacing continues, so the N=1 grid is exactly the cfg2 grid
    tz = (np.arange(z0, z1, dtype=np.float64) / (n_xy - 1) - 0.5) * w
    return np.sin(t), np.cos(t), np.sin(tz), np.cos(tz)


def gyroid_host(n_xy, z0, z1, nz_total):
    import numpy as np
    s, c, sz, cz = gyroid_tables(n_xy, z0, z1, nz_total)
    return (s[None, None, :] * c[None, :, None] + s[None, :, None] * cz[:, None, None]
            + sz[:, None, None] * c[None, None, :]).astype(np.float32)


def gyroid_device(n_xy, z0, z1, nz_total, dev):
    import torch
    s, c, sz, cz = (torch.from_numpy(a).to(dev) for a in gyroid_tables(n_xy, z0, z1, nz_total))
    out = torch.empty((z1 - z0, n_xy, n_xy), dtype=torch.float32, device=dev)
    for k in range(0, z1 - z0, 64):     # chunked: keeps the float64 temporaries small
        e = min(k + 64, z1 - z0)
        out[k:e] = (s[None, None, :] * c[None, :, None] + s[None, :, None] * cz[k:e, None, None]
                    + sz[k:e, None, None] * c[None, None, :]).to(torch.float32)
    return out


# --------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # "under load": upper half of the samples (idle samples before/after the region drag the median down)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------
# reference arm: the unmodified reference (oracle/_ref, -Ofast build) on host cores
# --------------------------------------------------------------------------
def ref_binding():
    sys.path.insert(0, str(ROOT / "tests"))
    from support import MC33Lib, REF_DIR
    so = REF_DIR / "libMC33_ref_f32.so"
    if not so.exists():
        return None
    return MC33Lib(so, "f32")


def cpu_sweep(lib, grid, isos, threads):
    """calculate_isosurface for every iso; `threads` independent MC33 objects in
    parallel (the reference itself is single threaded; distinct MC33 are independent)."""
    import numpy as np
    results = [None] * len(isos)

    def work(i):
        G, keep = lib.make_grid(grid)
        M = lib.lib.create_MC33(G)
        S = lib.lib.calculate_isosurface(M, lib.real_c(isos[i]))
        results[i] = (int(S.contents.nV), int(S.contents.nT))
        lib.lib.free_surface_memory(S); lib.lib.free_MC33(M); lib.lib.free_memory_grd(G)
    t0 = time.perf_counter()
    if threads <= 1:
        for i in range(len(isos)):
            work(i)
    else:
        pending = list(range(len(isos)))
        lock = threading.Lock()

        def runner():
            while True:
                with lock:
                    if not pending:
                        return
                    i = pending.pop(0)
                work(i)
        ts = [threading.Thread(target=runner) for _ in range(threads)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    return time.perf_counter() - t0, results


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib = ref_binding()
    base = {"impl": "reference", "metric": "Gvoxels/s per isosurface (iso sweep)", "unit": "Gvoxels/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    if lib is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libMC33_ref_f32.so not built"}))
        return
    grid = gyroid_host(N_SIDE, 0, N_SIDE, N_SIDE)
    cores = max(1, min(len(ISOS), os.cpu_count() or 1))
    for _ in range(args.warmup):
        cpu_sweep(lib, grid, ISOS[:cores], cores)
    ts = []
    ntri = 0
    for _ in range(args.steps):
        t, res = cpu_sweep(lib, grid, ISOS, cores)
        ts.append(t)
        ntri = sum(r[1] for r in res)
    t = sum(ts) / len(ts)
    vox = len(ISOS) * N_SIDE ** 3
    val = vox / t * 1e-9
    base.update({"value": val, "ms_per_step": t * 1e3, "mtriangles_per_s": ntri / t * 1e-6,
                 "config": {"workload": "cfg2: 512^3 float gyroid (4 periods), iso sweep of 8 values, reference "
                            "calculate_isosurface (-Ofast) on host cores", "isovalues": ISOS},
                 "cpu_baseline": {"value": val, "unit": "Gvoxels/s", "cores": cores, "kind": "reference",
                                  "sample": f"full sweep, {len(ISOS)} isovalues, one MC33 per thread on {cores} threads"},
                 "e2e": {"value": val, "unit": "Gvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(base))


# --------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from mc33_c_library_b200 import _cabi as cabi, slabs
    from mc33_c_library_b200.device import Extractor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, peak_src = peaks()

    # ---- grid: this rank's z-slab of a 512 x 512 x (512*world) gyroid ------------
    NZ = N_SIDE * world                      # sample slices of the global grid
    nz = NZ - 1                              # cell layers
    parts = slabs.partition(nz, world)
    sl = parts[rank]
    d = cabi.make_desc(cabi.F32, N_SIDE - 1, N_SIDE - 1, nz, z_lo=sl.z_lo, z_hi=sl.z_hi, cell_z0=sl.cell_z0,
                       cell_z1=sl.cell_z1, is_last=sl.is_last)
    grid = gyroid_device(N_SIDE, sl.z_lo, sl.z_hi, NZ, dev)
    ex = Extractor(d, device=local)
    ex.bind(grid)
    # a non-default torch stream carries everything: the kernels (mc33cu_set_stream),
    # the NCCL all-gather and the timing events
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ex.use_stream(stream)

    # ---- size the outputs once (largest isosurface of the sweep) -------------------
    cnt = [ex.count(i) for i in ISOS]
    capV = max(int(k.nV) for k in cnt) + 1024
    capT = max(int(k.nT) for k in cnt) + 1024
    buf = ex.alloc(capV, capT)
    counts_dev = torch.zeros((len(ISOS), 4), dtype=torch.int32, device=dev)
    gathered = torch.zeros((world, len(ISOS), 4), dtype=torch.int32, device=dev)
    bases_dev = torch.zeros((len(ISOS), 2), dtype=torch.int32, device=dev)

    def sweep():
        # the sweep's isovalues share one pass over the samples (mc33cu_classify_sweep);
        # count / scan / emit then run per isovalue on its pre-classified bitmap set
        ex.classify_sweep(ISOS)
        if world == 1:
            for j in range(len(ISOS)):
                ex.extract_set_async(j, buf)
        else:
            # every set keeps its own count state: count them all, ONE all-gather of the sweep's
            # counts across the slabs, then emit them all
            for j in range(len(ISOS)):
                ex.count_set_async(j, counts_dev[j])
            dist.all_gather_into_tensor(gathered, counts_dev)
            for j in range(len(ISOS)):
                ex.slab_bases_strided(gathered[0, j], 4 * len(ISOS), rank, world, bases_dev[j])
                ex.emit_set(j, buf, dev_bases=bases_dev[j])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        sweep()
    barrier()
    ex.sync()

    # ---- timed region ----------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = ex.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        sweep()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ex.launches() - l0
    clocks = sampler.stop() if sampler else None
    ex.sync()
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps

    # ---- per-kernel times (separate pass, events between kernels) -------------------
    ex.timing(True)
    kt = np.zeros(5)
    reps = 3
    for _ in range(reps):
        for iso in ISOS:
            ex.extract_async(iso, buf)
            torch.cuda.synchronize()
            kt += np.array(ex.kernel_times())
    ex.timing(False)
    kt /= reps * len(ISOS)
    # the sweep classify on its own: its time is shared by the sweep's isovalues
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ex.classify_sweep(ISOS)
    c0.record(stream)
    for _ in range(reps):
        ex.classify_sweep(ISOS)
    c1.record(stream)
    torch.cuda.synchronize()
    classify_single_ms = float(kt[0])
    kt[0] = c0.elapsed_time(c1) / reps / len(ISOS)
    knames = ["classify", "count", "rowscan", "emit_cells", "emit_vertices"]
    dom = int(np.argmax(kt))

    # ---- algorithmic bytes: grid read once + mesh written once -------
    npts_rank = (sl.cell_z1 - sl.cell_z0 + (1 if sl.is_last else 0)) * N_SIDE * N_SIDE
    mesh_bytes = [int(k.nV) * 28 + int(k.nT) * 12 for k in cnt]
    B_iso = [npts_rank * 4 + m for m in mesh_bytes]
    B_step = sum(B_iso)
    nT_step = sum(int(k.nT) for k in cnt)
    tot = torch.tensor([npts_rank * len(ISOS), nT_step, B_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    vox_step, tri_step, bytes_step = (float(x) for x in tot.tolist())
    value = vox_step / (ms_step * 1e-3) * 1e-9
    pipeline_gbs = bytes_step / world / (ms_step * 1e-3) * 1e-9     # per GPU
    # dominant kernel: its own algorithmic bytes per launch
    nV_avg = sum(int(k.nV) for k in cnt) / len(cnt)
    nT_avg = sum(int(k.nT) for k in cnt) / len(cnt)
    nC_avg = sum(int(k.nCentre) for k in cnt) / len(cnt)
    grid_bytes_rank = grid.numel() * 4
    kbytes = {"classify": grid_bytes_rank / len(ISOS), "count": grid_bytes_rank / 32, "rowscan": 0, "unused": 0,
              "emit_vertices": (nV_avg - nC_avg) * 28, "emit_cells": nT_avg * 12 + nC_avg * 28}
    kb = kbytes[knames[dom]]
    achieved = kb / (kt[dom] * 1e-3) * 1e-9 if kt[dom] > 0 else 0.0
    per_kernel = {n: {"ms": float(t), "algorithmic_bytes": float(kbytes[n]),
                      "achieved_gbs": float(kbytes[n] / (t * 1e-3) * 1e-9) if t > 0 else 0.0} for n, t in zip(knames, kt)}
    traffic = ncu_traffic(knames[dom])

    out = None
    if rank == 0:
        out = {"metric": "Gvoxels/s per isosurface (iso sweep)", "value": value, "unit": "Gvoxels/s",
               "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": f"cfg2: 512^3 float gyroid (4 periods) per GPU, iso sweep of 8 values; "
                          f"global grid 512x512x{NZ} samples in {world} z-slab(s)",
                          "isovalues": ISOS, "l2": "inputs larger than L2 (537 MB grid per GPU vs 126 MB L2)",
                          "parallelism": f"zslab{world}", "step": "one 8-isovalue sweep, grid resident in HBM: the samples are read once "
                          "for the 8 isovalues (mc33cu_classify_sweep), then count / scan / emit per isovalue"},
               "mtriangles_per_s": tri_step / (ms_step * 1e-3) * 1e-6,
               "ms_per_isosurface": ms_step / len(ISOS),
               "pipeline": {"algorithmic_bytes_per_step_per_gpu": bytes_step / world, "achieved_gbs": pipeline_gbs,
                            "frac_of_hbm_peak": pipeline_gbs / hbm_peak, "peak_gbs": hbm_peak, "peak_source": peak_src,
                            "note": "SURVEY 8d bytes (grid read once PER ISOSURFACE + mesh written once); the sweep classify "
                                    "reads the grid once per SWEEP, counting it so gives frac_shared_read",
                            "frac_shared_read": (bytes_step / world - (len(ISOS) - 1) * npts_rank * 4) / (ms_step * 1e-3) * 1e-9 / hbm_peak},
               "kernel_ms": dict(zip(knames, [float(x) for x in kt])),
               "kernel_ms_note": "per isosurface; classify = the one-pass sweep classify (k_classify_sweep, all 8 isovalues) / 8; "
                                 f"a single-isovalue classify (k_classify_vec) takes {classify_single_ms:.4f} ms",
               "roofline": {"bound": "hbm", "kernel": knames[dom], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                            "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": kb, "per_kernel": per_kernel,
                            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel in "
                                              "profiles/r1_ncu_full_summary.csv (one ncu --set full capture, iso 0.0)"},
               "gpu_launches": int(launches), "clocks": clocks}
    # ---- e2e: the drop-in C API with host buffers (rank 0, one GPU) ----------------
    if rank == 0:
        out["e2e"] = e2e_dropin(args)
        out["cpu_baseline"] = cpu_baseline()
        # full-size parity property: vertex / triangle counts of all 8 isosurfaces against the reference's
        ref_counts = out["cpu_baseline"].pop("counts", None)
        if ref_counts is not None and world == 1:
            out["counts_match_reference"] = ref_counts == [[int(k.nV), int(k.nT)] for k in cnt]
        print(json.dumps(out))
    ex.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def e2e_dropin(args):
    """same sweep through include/marching_cubes_33.h with HOST memory: every
    calculate_isosurface uploads the grid (the reference reads it at call time) and
    returns malloc'ed host arrays."""
    sys.path.insert(0, str(ROOT / "tests"))
    from support import MC33Lib
    lib = MC33Lib(ROOT / "mc33_c_library_b200" / "lib" / "libMC33_b200_f32.so", "f32")
    grid = gyroid_host(N_SIDE, 0, N_SIDE, N_SIDE)
    G, keep = lib.make_grid(grid)
    M = lib.lib.create_MC33(G)
    assert M, "create_MC33 failed"

    def sweep():
        d2h = 0
        for iso in ISOS:
            S = lib.lib.calculate_isosurface(M, lib.real_c(iso))
            assert S
            d2h += int(S.contents.nV) * 28 + int(S.contents.nT) * 12
            lib.lib.free_surface_memory(S)
        return d2h
    sweep()
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        d2h = sweep()
    t = (time.perf_counter() - t0) / steps
    lib.lib.free_MC33(M); lib.lib.free_memory_grd(G)
    return {"value": len(ISOS) * N_SIDE ** 3 / t * 1e-9, "unit": "Gvoxels/s", "ms_per_step": t * 1e3,
            "h2d_bytes_per_step": len(ISOS) * grid.nbytes, "d2h_bytes_per_step": d2h,
            "api": "grid_from_data_pointer/create_MC33 once, then calculate_isosurface + free_surface_memory per isovalue"}


def cpu_baseline():
    lib = ref_binding()
    if lib is None:
        return {"value": None, "unit": "Gvoxels/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
    grid = gyroid_host(N_SIDE, 0, N_SIDE, N_SIDE)
    t, res = cpu_sweep(lib, grid, ISOS, 1)
    return {"value": len(ISOS) * N_SIDE ** 3 / t * 1e-9, "unit": "Gvoxels/s", "cores": 1, "kind": "reference",
            "sample": f"the full 8-isovalue sweep once on one core ({t:.1f} s), reference built -Ofast -funroll-loops",
            "mtriangles_per_s": sum(r[1] for r in res) / t * 1e-6, "counts": [list(r) for r in res]}


def ncu_traffic(kname):
    """DRAM bytes (read + write) of one launch of the kernel, from the committed ncu summary."""
    import csv
    p = ROOT / "profiles" / "r1_ncu_full_summary.csv"
    if not p.exists():
        return None
    def mb(v):
        x, u = v.split()[:2]
        return float(x) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    for r in csv.DictReader(open(p)):
        if ("k_" + kname) in r["kernel"]:
            try:
                return mb(r["dram__bytes_read.sum"]) + mb(r["dram__bytes_write.sum"])
            except (KeyError, ValueError):
                return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
