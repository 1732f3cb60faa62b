#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 Marching Cubes 33 path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cfg1|cfg2|cfg3|cfg4|cfg5]

Workloads are BASELINE.json's configs at their named sizes (tools/workloads.py).  Defaults:
  --gpus 1   cfg2: 512^3 float gyroid, one STEP = the 8-isovalue sweep, grid resident in HBM
  --gpus N>1 cfg4: 2048^3 float smooth noise cut into N z-slabs (one halo slice below, two above),
             STRONG scaling; one STEP = one isosurface.  The only exchange is the NCCL all-gather
             of the per-slab counts.  The round-1 weak-scaling curve (cfg2, one 512-slice slab per
             rank) is kept as the second key "weak_cfg2" of the same line.
One STEP of any other workload = one pass over its isovalues.

Every line carries: roofline (dominant kernel), cpu_baseline (the unmodified reference on one host
core, bounded sample), e2e (the drop-in C API with HOST buffers at N GPUs) and parity:
  N = 1  vertex / triangle counts against the reference's own size_of_isosurface on the same samples
         (whole grid when the CPU finishes it in seconds, else a z sub-volume extracted separately);
  N > 1  sum nV, sum nT and order-independent 64-bit digests of the canonical vertex / triangle
         records over the ranks against ONE single-context extraction of the same global grid.

Prints ONE JSON line (contract in the task description / DESIGN.md section 6).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

import workloads  # noqa: E402  (tools/workloads.py)

REF_DIR = ROOT / "oracle" / "_ref"
KNAMES = ["classify", "count", "rowscan", "emit_cells", "emit_vertices"]
NCU_SUMMARY = ROOT / "profiles" / "r2_ncu_full_summary.csv"
NCU_SUMMARY_OLD = ROOT / "profiles" / "r1_ncu_full_summary.csv"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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
# the reference on host cores (oracle/_ref: the unmodified library, -Ofast build)
# --------------------------------------------------------------------------
def ref_binding(variant):
    from mc33_c_library_b200.dropin import MC33Lib
    so = REF_DIR / f"libMC33_ref_{variant}.so"
    return MC33Lib(so, variant) if so.exists() else None


# z sub-range of the grid the CPU reference is timed / counted on (whole grid when None)
CPU_SAMPLE_SLICES = {"cfg1": None, "cfg2": None, "cfg3": 256, "cfg4": 48, "cfg5": 40}
# ... and the range full-grid counts are checked on at N = 1 (None: whole grid; the CPU finishes it in seconds)
PARITY_FULL = {"cfg1": True, "cfg2": True, "cfg3": True, "cfg4": False, "cfg5": False}


def sample_range(W):
    n = CPU_SAMPLE_SLICES[W.name]
    NZ = W.shape[0]
    if n is None or n >= NZ:
        return 0, NZ
    z0 = (NZ - n) // 2
    return z0, z0 + n


def cpu_run(lib, W, host, isos, threads, count_only=False):
    """reference calculate_isosurface (or size_of_isosurface) for every isovalue on `host` (z,y,x samples).
    The reference is single threaded; the one parallel use of it that still returns the SAME meshes is one
    independent MC33 object per isovalue (distinct MC33 on one grid are independent), so
    `threads` is capped at the number of isovalues.  (Cutting the grid into z-chunks would return pieces with
    duplicated seam vertices and separate index spaces: not the reference's output.)"""
    import ctypes as C
    geom = W.geometry()
    NZ = host.shape[0]
    nchunk = 1
    bounds = [round(i * (NZ - 1) / nchunk) for i in range(nchunk + 1)]
    tasks = [(i, bounds[c], bounds[c + 1] + 1) for i in range(len(isos)) for c in range(nchunk) if bounds[c + 1] > bounds[c]]
    res = [[0, 0] for _ in isos]
    lock = threading.Lock()

    def work(t):
        i, a, b = t
        G, keep = lib.make_grid(host[a:b], geom)
        M = lib.lib.create_MC33(G)
        if count_only:
            nV, nT = C.c_uint(0), C.c_uint(0)
            lib.lib.size_of_isosurface(M, lib.real_c(isos[i]), C.byref(nV), C.byref(nT))
            v, t_ = int(nV.value), int(nT.value)
        else:
            S = lib.lib.calculate_isosurface(M, lib.real_c(isos[i]))
            v, t_ = int(S.contents.nV), int(S.contents.nT)
            lib.lib.free_surface_memory(S)
        lib.lib.free_MC33(M); lib.lib.free_memory_grd(G)
        with lock:
            res[i][0] += v; res[i][1] += t_

    t0 = time.perf_counter()
    if threads <= 1:
        for t in tasks:
            work(t)
    else:
        pending = list(tasks)

        def runner():
            while True:
                with lock:
                    if not pending:
                        return
                    t = pending.pop(0)
                work(t)
        ts = [threading.Thread(target=runner) for _ in range(min(threads, len(tasks)))]
        [t.start() for t in ts]
        [t.join() for t in ts]
    return time.perf_counter() - t0, res, len(tasks), nchunk


def host_sample(W, dev):
    z0, z1 = sample_range(W)
    return (z0, z1), W.host_slab(z0, z1, dev)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = workloads.make(args.workload)
    lib = ref_binding(W.variant)
    if lib is None:
        print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref/libMC33_ref_{W.variant}.so not built"}))
        return
    import torch
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    (z0, z1), host = host_sample(W, dev)
    isos = list(W.isos)
    cores = max(1, min(len(isos), os.cpu_count() or 1))
    for _ in range(min(args.warmup, 1)):
        cpu_run(lib, W, host, isos, cores)
    ts, ntri, ntask = [], 0, 0
    for _ in range(args.steps):
        t, res, ntask, nchunk = cpu_run(lib, W, host, isos, cores)
        ts.append(t)
        ntri = sum(r[1] for r in res)
    t = sum(ts) / len(ts)
    vox = len(isos) * host.size
    val = vox / t * 1e-9
    used = min(cores, ntask)
    sample = (f"z slices [{z0},{z1}) of the grid ({host.shape[0]}x{host.shape[1]}x{host.shape[2]} samples), {len(isos)} isovalue(s), "
              f"one independent MC33 per isovalue on {used} host thread(s) (the reference is single threaded: this is all the "
              "parallelism that returns the same meshes)")
    print(json.dumps({"impl": "reference", "metric": "Gvoxels/s per isosurface", "value": val, "unit": "Gvoxels/s",
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
                      "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
                      "dtype": W.variant, "data": "synthetic", "mtriangles_per_s": ntri / t * 1e-6,
                      "config": {"workload": f"{W.name}: {W.describe}; reference calculate_isosurface (-Ofast) on host cores",
                                 "isovalues": isos, "sample": sample},
                      "cpu_baseline": {"value": val, "unit": "Gvoxels/s", "cores": used, "kind": "reference", "sample": sample},
                      "e2e": {"value": val, "unit": "Gvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
class Rig:
    """one rank: its z-slab of the workload's grid on its GPU, a context, output buffers"""

    def __init__(self, W, rank, world, local, keys=False):
        import torch
        from mc33_c_library_b200 import _cabi as cabi, slabs
        from mc33_c_library_b200.device import Extractor
        self.W, self.rank, self.world = W, rank, world
        self.dev = torch.device("cuda", local)
        NZ, NY, NX = W.shape
        self.sl = slabs.partition(NZ - 1, world)[rank]
        sl = self.sl
        geom = W.geometry()
        from mc33_c_library_b200.dropin import DTYPES
        code, _, real = DTYPES[W.variant]
        if geom is not None:
            store, O, D, ca, cb, A, Ai = geom.derived(real)
            d = cabi.make_desc(code, NX - 1, NY - 1, NZ - 1, store, O, D, ca, cb, A.flat, Ai.flat, geom.tsa, geom.normal_neg,
                               z_lo=sl.z_lo, z_hi=sl.z_hi, cell_z0=sl.cell_z0, cell_z1=sl.cell_z1, is_last=sl.is_last)
        else:
            d = cabi.make_desc(code, NX - 1, NY - 1, NZ - 1, z_lo=sl.z_lo, z_hi=sl.z_hi, cell_z0=sl.cell_z0, cell_z1=sl.cell_z1,
                               is_last=sl.is_last)
        self.grid = W.device_slab(sl.z_lo, sl.z_hi, self.dev)
        self.ex = Extractor(d, device=local)
        self.ex.bind(self.grid)
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.ex.use_stream(self.stream)
        self.isos = list(W.isos)
        self.cnt = [self.ex.count(i) for i in self.isos]
        self.capV = max(int(k.nV) for k in self.cnt) + 1024
        self.capT = max(int(k.nT) for k in self.cnt) + 1024
        self.buf = self.ex.alloc(self.capV, self.capT, keys=keys)
        n = len(self.isos)
        self.counts_dev = torch.zeros((n, 4), dtype=torch.int32, device=self.dev)
        self.gathered = torch.zeros((world, n, 4), dtype=torch.int32, device=self.dev)
        self.gath1 = [torch.zeros((world, 4), dtype=torch.int32, device=self.dev) for _ in range(n)]
        self.bases_dev = torch.zeros((n, 2), dtype=torch.int32, device=self.dev)
        self.npts_owned = (sl.cell_z1 - sl.cell_z0 + (1 if sl.is_last else 0)) * NY * NX
        # An iso sweep on one GPU: once the sweep classify has run its isovalues are independent (every set keeps its own
        # count state, the context one vertex-task buffer per stream), so they alternate between this rank's stream and a
        # second one, each with its own output buffers: the tail of one extraction's kernels overlaps the next one's
        # (tools/time_overlap.py: 0.360 -> 0.344 ms per isosurface at cfg2; more than two streams do not help).
        self.side, self.side_buf = None, None
        if world == 1 and W.sweep and n > 1 and int(os.environ.get("MC33_BENCH_STREAMS", "2")) > 1:
            self.side = torch.cuda.Stream(device=self.dev)
            self.side_buf = self.ex.alloc(self.capV, self.capT, keys=keys)
            self.ev_fork, self.ev_join = torch.cuda.Event(), torch.cuda.Event()

    def step(self):
        """one pass over the workload's isovalues"""
        import torch.distributed as dist
        ex, W, n = self.ex, self.W, len(self.isos)
        if W.sweep:
            ex.classify_sweep(self.isos)
        if self.world == 1 and self.side is not None:
            self.ev_fork.record(self.stream)
            self.side.wait_event(self.ev_fork)
            for j in range(n):
                if j & 1:
                    ex.use_stream(self.side)
                    ex.extract_set_async(j, self.side_buf)
                else:
                    ex.use_stream(self.stream)
                    ex.extract_set_async(j, self.buf)
            self.ev_join.record(self.side)
            self.stream.wait_event(self.ev_join)
            ex.use_stream(self.stream)
        elif self.world == 1:
            for j, iso in enumerate(self.isos):
                if W.sweep:
                    ex.extract_set_async(j, self.buf)
                else:
                    ex.extract_async(iso, self.buf)
        elif W.sweep:
            # every set keeps its own count state: count them all, ONE all-gather of the sweep's counts, emit them all
            for j in range(n):
                ex.count_set_async(j, self.counts_dev[j])
            dist.all_gather_into_tensor(self.gathered, self.counts_dev)
            for j in range(n):
                ex.slab_bases_strided(self.gathered[0, j], 4 * n, self.rank, self.world, self.bases_dev[j])
                ex.emit_set(j, self.buf, dev_bases=self.bases_dev[j])
        else:
            for j, iso in enumerate(self.isos):
                ex.count_async(iso, self.counts_dev[j])
                dist.all_gather_into_tensor(self.gath1[j], self.counts_dev[j])
                ex.slab_bases(self.gath1[j], self.rank, self.world, self.bases_dev[j])
                ex.emit(self.buf, dev_bases=self.bases_dev[j])

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, steps, warmup, sampler=None):
        import torch
        import torch.distributed as dist
        for _ in range(max(warmup, 3)):
            self.step()
        self.barrier()
        self.ex.sync()
        if sampler:
            sampler.start()
            time.sleep(0.3)
        l0 = self.ex.launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        for _ in range(steps):
            self.step()
        e1.record(self.stream)
        self.barrier()
        ms = e0.elapsed_time(e1)
        launches = self.ex.launches() - l0
        clocks = sampler.stop() if sampler else None
        self.ex.sync()
        tmax = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return float(tmax.item()) / steps, launches, clocks

    def kernel_times(self, reps=3):
        """per-kernel device times per isosurface (CUDA events between the kernels, separate pass)"""
        import numpy as np
        import torch
        ex = self.ex
        ex.timing(True)
        kt = np.zeros(5)
        for _ in range(reps):
            for iso in self.isos:
                ex.extract_async(iso, self.buf)
                torch.cuda.synchronize()
                kt += np.array(ex.kernel_times())
        ex.timing(False)
        kt /= reps * len(self.isos)
        single = float(kt[0])
        if self.W.sweep:
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ex.classify_sweep(self.isos)
            c0.record(self.stream)
            for _ in range(reps):
                ex.classify_sweep(self.isos)
            c1.record(self.stream)
            torch.cuda.synchronize()
            kt[0] = c0.elapsed_time(c1) / reps / len(self.isos)
        return kt, single

    def digests(self, j=0):
        """order-independent digests of this rank's part of isosurface j (buffers with keys), summed over the ranks"""
        import torch
        import torch.distributed as dist
        k = self.cnt[j]
        nV, nT, nS = int(k.nV), int(k.nT), int(k.nShared)
        b = self.buf
        vkey = b["vkey"][:nV]
        dv = workloads.vertex_digest(vkey, b["V"][:nV], b["N"][:nV])
        vb, vbn = (int(x) & 0xFFFFFFFF for x in self.bases_dev[j].tolist()) if self.world > 1 else (0, nV)
        # the vertices of the seam slice are numbered by the next slab: fetch their keys from it
        halo = int(k.nSharedHalo)
        halo_keys = torch.zeros(max(halo, 1), dtype=torch.int64, device=self.dev)
        if self.world > 1:
            # what this rank's predecessor needs = the keys of the first nSharedHalo(prev) vertices here
            need_prev = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
            mine = torch.tensor([halo], dtype=torch.int64, device=self.dev)
            dist.all_gather_into_tensor(need_prev, mine)
            ops = []
            if self.rank > 0 and int(need_prev[self.rank - 1]) > 0:
                ops.append(dist.P2POp(dist.isend, vkey[:int(need_prev[self.rank - 1])].contiguous(), self.rank - 1))
            if self.rank + 1 < self.world and halo > 0:
                ops.append(dist.P2POp(dist.irecv, halo_keys[:halo], self.rank + 1))
            if ops:
                for r in dist.batch_isend_irecv(ops):
                    r.wait()
            torch.cuda.synchronize()

        def key_of(ids):
            loc = ids - vb
            own = (loc >= 0) & (loc < nV)
            out = torch.where(own, vkey[loc.clamp(0, max(nV - 1, 0))], halo_keys[(ids - vbn).clamp(0, max(halo - 1, 0))])
            return out
        dt = workloads.triangle_digest(b["tcell"][:nT], b["T"][:nT], key_of)
        # sums over the ranks, kept as 4 x 16-bit limbs in int64 so that the all-reduce cannot overflow
        vals = [nV, nT, dv, dt]
        limbs = torch.tensor([(v >> (16 * i)) & 0xFFFF for v in vals for i in range(4)], dtype=torch.int64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(limbs)
        l = limbs.tolist()
        tot = [sum(l[4 * q + i] << (16 * i) for i in range(4)) for q in range(4)]
        return {"nV": tot[0], "nT": tot[1], "vertex_digest": f"{tot[2] & workloads.M64:016x}",
                "triangle_digest": f"{tot[3] & workloads.M64:016x}"}

    def close(self):
        self.ex.close()
        self.buf = None
        self.side_buf = None
        self.grid = None


def mesh_counts_equal(a, b):
    return [list(x) for x in a] == [list(x) for x in b]


def parity_single_gpu(W, rig, lib, sample, host, ref_counts):
    """N = 1: counts of every isosurface against the reference's on the same samples"""
    import torch
    from mc33_c_library_b200.device import Extractor
    from mc33_c_library_b200 import _cabi as cabi
    (z0, z1) = sample
    NZ = W.shape[0]
    ours_full = [[int(k.nV), int(k.nT)] for k in rig.cnt]
    out = {"kind": "vertex / triangle counts of every isovalue against the unmodified reference (oracle/_ref) on the same samples"}
    if lib is None:
        out.update({"match": None, "note": "oracle/_ref not built"})
        return out
    if (z0, z1) == (0, NZ):
        out.update({"scope": "whole grid", "ours": ours_full, "reference": ref_counts, "match": mesh_counts_equal(ours_full, ref_counts)})
        return out
    # a z sub-volume: extracted on its own by a second context on the same samples
    sub = rig.grid[z0:z1].contiguous()
    rig2 = Rig.__new__(Rig)
    geom = W.geometry()
    from mc33_c_library_b200.dropin import DTYPES
    code, _, real = DTYPES[W.variant]
    NY, NX = W.shape[1], W.shape[2]
    if geom is not None:
        store, O, D, ca, cb, A, Ai = geom.derived(real)
        d = cabi.make_desc(code, NX - 1, NY - 1, z1 - z0 - 1, store, O, D, ca, cb, A.flat, Ai.flat, geom.tsa, geom.normal_neg)
    else:
        d = cabi.make_desc(code, NX - 1, NY - 1, z1 - z0 - 1)
    ex = Extractor(d, device=rig.dev.index)
    ex.bind(sub)
    ours = [[int(k.nV), int(k.nT)] for k in (ex.count(i) for i in rig.isos)]
    ex.close()
    out.update({"scope": f"z sub-volume [{z0},{z1}) extracted separately (the CPU needs minutes for the whole grid)",
                "ours": ours, "reference": ref_counts, "match": mesh_counts_equal(ours, ref_counts), "whole_grid_counts": ours_full})
    return out


def parity_multi_gpu(W, rig):
    """N > 1: the sharded mesh of isosurface 0 against ONE single-context extraction of the same global grid"""
    import torch
    import torch.distributed as dist
    # the sharded mesh, with canonical keys
    keyed = rig.ex.alloc(rig.capV, rig.capT, keys=True)
    old = rig.buf
    rig.buf = keyed
    j = 0
    isos_keep, cnt_keep = rig.isos, rig.cnt
    rig.step()
    rig.barrier()
    rig.ex.sync()
    # (the last isovalue of the step is what the buffers hold)
    j = len(rig.isos) - 1
    sharded = rig.digests(j)
    rig.buf = old
    del keyed
    torch.cuda.empty_cache()
    single = None
    if rig.rank == 0:
        one = Rig(W, 0, 1, rig.dev.index, keys=True)
        one.isos = [isos_keep[j]]
        one.cnt = [one.ex.count(isos_keep[j])]
        one.ex.extract_async(isos_keep[j], one.buf)
        torch.cuda.synchronize()
        one.ex.sync()
        single = one.digests(0)
        one.close()
        del one
        torch.cuda.empty_cache()
        torch.cuda.set_stream(rig.stream)
    if rig.world > 1:
        dist.barrier()
    if rig.rank != 0:
        return None
    return {"kind": f"z-slab mesh over {rig.world} ranks against one single-context extraction of the same global grid (rank 0's GPU): "
                    "sum nV, sum nT, order-independent 64-bit digests of (canonical vertex key, position bits, normal bits) and of "
                    "(cell, three vertex keys)", "isovalue": isos_keep[j], "sharded": sharded, "single_context": single,
            "match": sharded == single}


def e2e_dropin(W, args, world, dev):
    """the same pass over the isovalues through include/marching_cubes_33.h with HOST memory: every
    calculate_isosurface brings the samples to the GPU(s) (the reference reads them at call time) and
    returns host arrays.  At N GPUs the drop-in cuts the grid into N z-slabs (MC33_B200_GPUS)."""
    from mc33_c_library_b200.dropin import dropin
    os.environ["MC33_B200_GPUS"] = str(world)
    lib = dropin(W.variant)
    NZ = W.shape[0]
    # bounded host footprint: the meshes of cfg4 / cfg5 are tens of GB; their e2e leg runs on a z range
    cap = {"cfg4": 128 * max(world, 1), "cfg5": 128}.get(W.name)
    z0, z1 = (0, NZ) if not cap or cap >= NZ else ((NZ - cap) // 2, (NZ - cap) // 2 + cap)
    host = W.host_slab(z0, z1, dev)
    G, keep = lib.make_grid(host, W.geometry())
    M = lib.lib.create_MC33(G)
    assert M, "create_MC33 failed"
    isos = list(W.isos)
    real_b = 8 if W.variant == "f64" else 4

    def one_pass():
        d2h = 0
        for iso in isos:
            S = lib.lib.calculate_isosurface(M, lib.real_c(iso))
            assert S
            d2h += int(S.contents.nV) * (3 * real_b + 16) + int(S.contents.nT) * 12
            lib.lib.free_surface_memory(S)
        return d2h
    # warm-up: the first passes page-lock the grid, grow the pooled result arrays and the per-slab device staging (the
    # drop-in sizes the result arrays of a call from the meshes it has produced before: two passes settle that)
    one_pass()
    one_pass()
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        d2h = one_pass()
    t = (time.perf_counter() - t0) / steps
    lib.lib.free_MC33(M); lib.lib.free_memory_grd(G)
    gpus_used, slabs_used = 1, 1
    try:
        gpus_used = int(lib.lib.mc33_dropin_gpus_last())
        slabs_used = int(lib.lib.mc33_dropin_slabs_last())
    except AttributeError:
        pass
    return {"value": len(isos) * host.size / t * 1e-9, "unit": "Gvoxels/s", "ms_per_step": t * 1e3,
            "h2d_bytes_per_step": len(isos) * host.nbytes, "d2h_bytes_per_step": d2h, "n_gpus": gpus_used, "z_chunks": slabs_used,
            "schedule": "the samples go up in z-chunks, one after the other per GPU; a chunk is counted, emitted and its part of the mesh "
                        "downloaded while the next chunks are still arriving" if slabs_used > gpus_used else "upload, extract, download",
            "sample": "whole grid" if (z0, z1) == (0, NZ) else f"z slices [{z0},{z1}) (bounded host memory: the whole mesh is tens of GB)",
            "api": "grid_from_data_pointer/create_MC33 once, then calculate_isosurface + free_surface_memory per isovalue"}


def ncu_traffic(kname):
    """DRAM bytes (read + write) of one launch of the kernel, from the committed ncu summary (cfg2, iso 0.0)."""
    import csv
    p = NCU_SUMMARY if NCU_SUMMARY.exists() else NCU_SUMMARY_OLD
    if not p.exists():
        return None, None

    def mb(v):
        x, u = v.split()[:2]
        return float(x) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    for r in csv.DictReader(open(p)):
        if ("k_" + kname) in r["kernel"]:
            try:
                return mb(r["dram__bytes_read.sum"]) + mb(r["dram__bytes_write.sum"]), p.name
            except (KeyError, ValueError):
                return None, None
    return None, None


def measure(W, args, rank, world, local, with_extras=True):
    """-> the JSON line's dict on rank 0 (None elsewhere)"""
    import numpy as np
    import torch
    import torch.distributed as dist
    hbm_peak, peak_src = peaks()
    rig = Rig(W, rank, world, local)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_step, launches, clocks = rig.timed(args.steps, args.warmup, sampler)
    kt, classify_single = rig.kernel_times()
    n_iso = len(rig.isos)
    real_b = 8 if W.variant == "f64" else 4
    vbytes = 3 * real_b + 16
    # algorithmic bytes: grid read once per isosurface + mesh written once
    mesh_bytes = [int(k.nV) * vbytes + int(k.nT) * 12 for k in rig.cnt]
    grid_b = rig.npts_owned * W.sample_bytes
    tot = torch.tensor([rig.npts_owned * n_iso, sum(int(k.nT) for k in rig.cnt), sum(mesh_bytes) + n_iso * grid_b,
                        sum(int(k.nV) for k in rig.cnt)], dtype=torch.float64, device=rig.dev)
    if world > 1:
        dist.all_reduce(tot)
    vox_step, tri_step, bytes_step, vert_step = (float(x) for x in tot.tolist())
    value = vox_step / (ms_step * 1e-3) * 1e-9
    agg_gbs = bytes_step / (ms_step * 1e-3) * 1e-9
    nV_avg = sum(int(k.nV) for k in rig.cnt) / n_iso
    nT_avg = sum(int(k.nT) for k in rig.cnt) / n_iso
    nC_avg = sum(int(k.nCentre) for k in rig.cnt) / n_iso
    slab_bytes = rig.grid.numel() * W.sample_bytes
    kbytes = {"classify": slab_bytes / (n_iso if W.sweep else 1), "count": rig.grid.numel() / 8, "rowscan": 0,
              "emit_vertices": (nV_avg - nC_avg) * vbytes, "emit_cells": nT_avg * 12 + nC_avg * vbytes}
    dom = int(np.argmax(kt))
    kb = kbytes[KNAMES[dom]]
    achieved = kb / (kt[dom] * 1e-3) * 1e-9 if kt[dom] > 0 else 0.0
    per_kernel = {n: {"ms": float(t), "algorithmic_bytes": float(kbytes[n]),
                      "achieved_gbs": float(kbytes[n] / (t * 1e-3) * 1e-9) if t > 0 else 0.0} for n, t in zip(KNAMES, kt)}
    traffic, traffic_file = ncu_traffic(KNAMES[dom]) if W.name == "cfg2" else (None, None)
    out = None
    if rank == 0:
        NZ, NY, NX = W.shape
        strong = world > 1 and not getattr(W, "weak", False)
        out = {"metric": "Gvoxels/s per isosurface", "value": value, "unit": "Gvoxels/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
               "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": W.variant, "data": "synthetic",
               "config": {"workload": f"{W.name}: {W.describe}; global grid {NX}x{NY}x{NZ} samples in {world} z-slab(s)",
                          "isovalues": rig.isos, "parallelism": f"zslab{world}",
                          "l2": f"inputs larger than L2 ({slab_bytes / 1e6:.0f} MB of samples per GPU vs 126 MB L2)"
                                if slab_bytes > 126e6 else "inputs fit L2; every step rewrites > L2 of bitmaps / prefixes / mesh in between",
                          "step": ("one 8-isovalue sweep, grid resident in HBM: the samples are read once for the 8 isovalues "
                                   "(mc33cu_classify_sweep), then count / scan / emit per isovalue"
                                   + (", the isovalues alternating between two CUDA streams of the context" if rig.side is not None else "")) if W.sweep else
                                  f"one pass over {n_iso} isovalue(s), grid resident in HBM: classify, count / scan, emit per isovalue"},
               "mtriangles_per_s": tri_step / (ms_step * 1e-3) * 1e-6,
               "ms_per_isosurface": ms_step / n_iso,
               "mesh": {"vertices_per_step": vert_step, "triangles_per_step": tri_step},
               "pipeline": {"algorithmic_bytes_per_step": bytes_step, "achieved_gbs_aggregate": agg_gbs,
                            "frac_of_hbm_peak": agg_gbs / (hbm_peak * world), "peak_gbs_per_gpu": hbm_peak, "peak_source": peak_src,
                            "note": "SURVEY 8d bytes: grid read once PER ISOSURFACE + mesh written once, over the aggregate peak of the N GPUs"},
               "kernel_ms": dict(zip(KNAMES, [float(x) for x in kt])),
               "kernel_ms_note": "rank 0, per isosurface" + ("; classify = the one-pass sweep classify / 8; a single-isovalue classify takes "
                                                               f"{classify_single:.4f} ms" if W.sweep else ""),
               "roofline": {"bound": "hbm", "kernel": KNAMES[dom], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                            "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": kb, "per_kernel": per_kernel,
                            "traffic_source": (f"dram__bytes_read.sum + dram__bytes_write.sum of this kernel in profiles/{traffic_file} "
                                               "(one ncu --set full capture, cfg2 iso 0.0)") if traffic else None},
               "gpu_launches": int(launches), "clocks": clocks}
        if W.sweep:
            # the sweep reads the grid once per SWEEP: the same fraction with the shared read counted once
            shared = bytes_step - (n_iso - 1) * grid_b * world
            out["pipeline"]["frac_shared_read"] = shared / (ms_step * 1e-3) * 1e-9 / (hbm_peak * world)
    if not with_extras:
        rig.close()
        return out
    # ---- parity -------------------------------------------------------------------------------
    if world > 1:
        par = parity_multi_gpu(W, rig)
        if rank == 0:
            out["parity"] = par
    elif rank == 0:
        lib = ref_binding(W.variant)
        sample, host = (0, W.shape[0]), None
        ref_counts, cpu = None, {"value": None, "unit": "Gvoxels/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
        if lib is not None:
            sample, host = host_sample(W, rig.dev)
            t, res, _, _ = cpu_run(lib, W, host, rig.isos, 1)
            cpu = {"value": n_iso * host.size / t * 1e-9, "unit": "Gvoxels/s", "cores": 1, "kind": "reference",
                   "sample": f"z slices [{sample[0]},{sample[1]}) of the grid, {n_iso} isovalue(s), calculate_isosurface once each on one core "
                             f"({t:.1f} s); reference built -Ofast -funroll-loops; CPU time is linear in voxels for fixed statistics",
                   "mtriangles_per_s": sum(r[1] for r in res) / t * 1e-6}
            ref_counts = res
            if PARITY_FULL[W.name] and sample != (0, W.shape[0]):
                # whole-grid counts from the reference's count-only twin (size_of_isosurface), all host threads
                sample = (0, W.shape[0])
                host = W.host_slab(0, W.shape[0], rig.dev)
                _, ref_counts, _, _ = cpu_run(lib, W, host, rig.isos, n_iso, count_only=True)     # one thread per isovalue, no z-chunks
        out["cpu_baseline"] = cpu
        out["parity"] = parity_single_gpu(W, rig, lib, sample, host, ref_counts)
        del host
    dev = rig.dev
    rig.close()
    del rig
    torch.cuda.empty_cache()
    # ---- e2e: the drop-in C API with host buffers (rank 0 drives all N GPUs) --------------------
    if rank == 0:
        out["e2e"] = e2e_dropin(W, args, world, dev)
    if world > 1:
        dist.barrier()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries ONE JSON line: whatever libraries print there while the job runs (NCCL announces its version on
    # rank 0's stdout) goes to stderr instead; the descriptor comes back for the line itself
    sys.stdout.flush()
    keep_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = workloads.make(args.workload)
    out = measure(W, args, rank, world, local)
    if world > 1 and args.workload == "cfg4" and not args.no_weak:
        # the round-1 curve: weak scaling of cfg2 (every rank one 512-slice slab of a 512 x 512 x 512N gyroid)
        W2 = workloads.make("cfg2", n=512, nz=512 * world)
        W2.weak = True
        w = measure(W2, args, rank, world, local, with_extras=False)
        if rank == 0:
            out["weak_cfg2"] = {k: w[k] for k in ("value", "unit", "ms_per_step", "ms_per_isosurface", "mtriangles_per_s", "gpu_launches")}
            out["weak_cfg2"]["scaling"] = "weak"
            out["weak_cfg2"]["workload"] = w["config"]["workload"]
    sys.stdout.flush()
    os.dup2(keep_stdout, 1)
    os.close(keep_stdout)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--no-weak", action="store_true", help="N > 1, cfg4: skip the secondary weak-scaling cfg2 measurement")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "cfg2" if args.gpus <= 1 else "cfg4"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
