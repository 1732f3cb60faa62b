#!/usr/bin/env python3
"""BASELINE config 4: n^3 float smooth-noise grid (default 2048), z-slab sharded over the
ranks of one box (halo slices + NCCL all-gather of the per-slab counts), device pipeline.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29512 tools/time_cfg4.py [n=2048] [iso=0.0] [reps=5]

The grid is a trilinear up-sampling (dyadic weights, exact in float) of a hashed coarse lattice
with 32-sample cells, so every rank generates exactly its own slices on its own GPU and a
single-GPU run of a smaller n gives the same function.  Besides the time it checks, on the
device, what must hold for any correct sharded mesh: every triangle index is below the global
vertex count, a slab's triangles only reference its own vertices or the first ones of the next
slab, and the seam accounting of the counts adds up.  One JSON line from rank 0."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist

from mc33_c_library_b200 import _cabi as cabi, slabs
from mc33_c_library_b200.device import Extractor

kv = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
n = int(kv.get("n", 2048))
iso = float(kv.get("iso", 0.0))
reps = int(kv.get("reps", 5))
CELL = 32

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def lattice(m):
    """(m+1)^3 values in [-1, 1) from an integer hash of the lattice coordinates (the same on every rank)"""
    i = torch.arange(m + 1, device=dev, dtype=torch.int64)
    h = (i[:, None, None] * 73856093) ^ (i[None, :, None] * 19349663) ^ (i[None, None, :] * 83492791)
    h = (h * 2654435761) & 0xFFFFFFFF
    h = ((h >> 13) ^ h) * 1274126177 & 0xFFFFFFFF
    return ((h >> 8).to(torch.float32) / float(1 << 23)) - 1.0          # [z][y][x]


def slab_grid(z0, z1):
    m = (n + CELL - 1) // CELL
    L = lattice(m)
    idx = torch.arange(n, device=dev)
    c0 = (idx // CELL).long()
    t = ((idx % CELL).to(torch.float32) / CELL)
    out = torch.empty((z1 - z0, n, n), dtype=torch.float32, device=dev)
    # interpolate along x and y once per pair of lattice slices, then along z per sample slice
    cache = {}

    def plane(k):            # lattice slice k up-sampled in y and x -> [n][n]
        if k not in cache:
            if len(cache) > 2:
                cache.pop(min(cache))
            P = L[k]
            px = P[:, c0] * (1 - t)[None, :] + P[:, c0 + 1] * t[None, :]          # [m+1][n]
            cache[k] = px[c0, :] * (1 - t)[:, None] + px[c0 + 1, :] * t[:, None]   # [n][n]
        return cache[k]

    for z in range(z0, z1):
        k, tz = z // CELL, (z % CELL) / CELL
        out[z - z0] = plane(k) * (1 - tz) + plane(k + 1) * tz
    return out


nz = n - 1
sl = slabs.partition(nz, world)[rank]
desc = cabi.make_desc(cabi.F32, n - 1, n - 1, nz, z_lo=sl.z_lo, z_hi=sl.z_hi, cell_z0=sl.cell_z0, cell_z1=sl.cell_z1,
                      is_last=sl.is_last)
grid = slab_grid(sl.z_lo, sl.z_hi)
ex = Extractor(desc, device=local)
ex.bind(grid)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ex.use_stream(stream)

k = ex.count(iso)
nV, nT = int(k.nV), int(k.nT)
buf = ex.alloc(nV + 1024, nT + 1024)
counts_dev = torch.zeros(4, dtype=torch.int32, device=dev)
gathered = torch.zeros((world, 4), dtype=torch.int32, device=dev)
bases_dev = torch.zeros(2, dtype=torch.int32, device=dev)


def extract():
    if world == 1:
        ex.extract_async(iso, buf)
    else:
        ex.count_async(iso, counts_dev)
        dist.all_gather_into_tensor(gathered, counts_dev)
        ex.slab_bases(gathered, rank, world, bases_dev)
        ex.emit(buf, dev_bases=bases_dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(2):
    extract()
barrier()
ex.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
barrier()
e0.record(stream)
for _ in range(reps):
    extract()
e1.record(stream)
barrier()
ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ex.sync()

# ---- consistency of the sharded mesh (device side) ------------------------------------
tot = torch.tensor([nV, nT], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(tot)
    vb, vbn = (int(x) for x in bases_dev.tolist())
else:
    vb, vbn = 0, nV
T = buf["T"][:nT].view(torch.int32).to(torch.int64) & 0xFFFFFFFF
tmin, tmax = (int(T.min()), int(T.max())) if nT else (0, 0)
halo = int(k.nSharedHalo)
ok_range = nT == 0 or (tmin >= vb and tmax < vbn + halo and tmax < int(tot[0]))
finite = bool(torch.isfinite(buf["V"][:nV]).all()) if nV else True
ex.timing(True)
extract(); torch.cuda.synchronize()
kt = ex.kernel_times() if world == 1 else None
ex.timing(False)
flags = torch.tensor([int(ok_range), int(finite)], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)

if rank == 0:
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    npts = float(n) ** 3
    B = npts * 4 + float(tot[0]) * 28 + float(tot[1]) * 12
    t = float(ms.item()) * 1e-3
    print(json.dumps({"config": f"cfg4: {n}^3 f32 smooth noise (hashed lattice, {CELL}-sample cells), iso {iso}, {world} z-slab(s)",
                      "n_gpus": world, "ms_per_isosurface": t * 1e3, "gvoxels_per_s": npts / t * 1e-9,
                      "mtriangles_per_s": float(tot[1]) / t * 1e-6, "nV": int(tot[0]), "nT": int(tot[1]),
                      "algorithmic_GB": B / 1e9, "aggregate_gbs": B / t * 1e-9, "frac_of_aggregate_hbm_peak": B / t * 1e-9 / (peak * world),
                      "checks": {"triangle_ids_in_range": bool(flags[0]), "positions_finite": bool(flags[1])},
                      "kernel_ms_rank0": kt}))
ex.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
