"""CPU checks of the boundary and of the multi-rank host logic:

* libmc33cu.so loads and exports every function include/mc33cu.h declares, the
  drop-in libraries export the marching_cubes_33.h API (no compute calls: no GPU here);
* the z-slab partition and index bases, including a world_size-2 run over
  torch.distributed (gloo) that mirrors what bench.py does over NCCL: per-rank
  counts -> all_gather -> bases -> per-rank emit -> the union equals the oracle mesh
  (the per-slab extraction itself is stepped on the CPU by tests/hostemu).
"""
import ctypes as C
import os
import re
import socket
from pathlib import Path

import numpy as np
import pytest

from mc33_c_library_b200 import _cabi, slabs

ROOT = Path(__file__).resolve().parent.parent
LIBDIR = ROOT / "mc33_c_library_b200" / "lib"


def test_cabi_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "mc33cu.h").read_text()
    declared = sorted(set(re.findall(r"\b(mc33cu_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 18
    lib = _cabi.load()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_cabi.EXPORTS) == declared


def test_cabi_argument_errors_without_gpu():
    lib = _cabi.load()
    h = C.c_void_p()
    assert lib.mc33cu_create(None, 0, C.byref(h)) == _cabi.ERR_ARG
    d = _cabi.make_desc(_cabi.F32, 0, 4, 4)
    assert lib.mc33cu_create(C.byref(d), 0, C.byref(h)) == _cabi.ERR_ARG       # empty grid
    d = _cabi.make_desc(_cabi.F32, 4, 4, 4, z_lo=2, z_hi=5, cell_z0=0, cell_z1=4)
    assert lib.mc33cu_create(C.byref(d), 0, C.byref(h)) == _cabi.ERR_ARG       # halo too small
    assert lib.mc33cu_count(None, 0.0, None) == _cabi.ERR_ARG
    assert b"null" in lib.mc33cu_last_error()
    # the sweep / slab entry points reject null contexts and arguments the same way
    isos = (C.c_double * 2)(0.0, 1.0)
    assert lib.mc33cu_classify_sweep(None, isos, 2) == _cabi.ERR_ARG
    assert lib.mc33cu_count_set_async(None, 0, None) == _cabi.ERR_ARG
    assert lib.mc33cu_extract_set_device(None, 0, None) == _cabi.ERR_ARG
    assert lib.mc33cu_emit_set_device(None, 0, None) == _cabi.ERR_ARG
    assert lib.mc33cu_slab_bases(None, None, 0, 1, None) == _cabi.ERR_ARG
    assert lib.mc33cu_slab_bases_strided(None, None, 4, 0, 1, None) == _cabi.ERR_ARG
    lib.mc33cu_destroy(None)


@pytest.mark.parametrize("variant", ["f32", "f64", "u8", "u16", "u32", "f32_ortho"])
def test_dropin_libraries_export_the_reference_api(variant):
    """the symbols SURVEY.md section 8b lists as must-export, per element-type variant"""
    lib = C.CDLL(str(LIBDIR / f"libMC33_b200_{variant}.so"), mode=os.RTLD_NOW | os.RTLD_LOCAL)
    for s in ["grid_from_data_pointer", "create_MC33", "calculate_isosurface", "size_of_isosurface", "free_surface_memory",
              "free_MC33", "free_memory_grd", "generate_grid_from_fn", "alloc_F", "adjustvectorlenght_s", "DefaultColorMC"]:
        assert hasattr(lib, s), s
    if variant != "f32_ortho":
        for s in ["mult_Abf", "_multA_bf", "_multTSA_bf"]:
            assert hasattr(lib, s), s
    assert C.c_int.in_dll(lib, "DefaultColorMC").value == -10724260          # 0xff5c5c5c (reference c:76-80)
    # host-only entry points behave like the reference without touching the GPU
    lib.grid_from_data_pointer.restype = C.c_void_p
    lib.grid_from_data_pointer.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_void_p]
    assert lib.grid_from_data_pointer(0, 3, 3, None) is None
    lib.free_surface_memory(None); lib.free_MC33(None); lib.free_memory_grd(None)


@pytest.mark.parametrize("nz,world", [(511, 1), (511, 2), (511, 8), (5, 8), (7, 3), (1023, 4)])
def test_partition_covers_every_layer_once(nz, world):
    parts = slabs.partition(nz, world)
    assert len(parts) == world
    live = [p for p in parts if p is not None]
    assert live[0].cell_z0 == 0 and live[-1].cell_z1 == nz and live[-1].is_last
    for a, b in zip(live, live[1:]):
        assert a.cell_z1 == b.cell_z0 and not a.is_last
    for p in live:
        assert p.z_lo <= max(p.cell_z0, 1) - 1 and p.z_hi >= min(p.cell_z1 + 2, nz + 1)
        assert 0 <= p.z_lo < p.z_hi <= nz + 1
    sizes = [p.cell_z1 - p.cell_z0 for p in live]
    assert max(sizes) - min(sizes) <= 1


def test_bases():
    assert slabs.bases([(10, 5), None, (7, 2), (0, 0)]) == [(0, 10), (10, 10), (10, 17), (17, 17)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from support import DTYPES, emu_count, emu_emit, make_desc, noise_grid
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    variant, iso = "u8", 2.0
    _, sdt, real = DTYPES[variant]
    a = noise_grid(0, variant, scale=4, shape=(19, 9, 40))          # every rank builds the same grid, keeps its slab
    sl = slabs.partition(a.shape[0] - 1, world)[rank]
    sub = np.ascontiguousarray(a[sl.z_lo:sl.z_hi])
    d = make_desc(a.shape, variant, None, sl)
    k = emu_count(sub, iso, d)
    mine = torch.tensor([int(k.nV), int(k.nT), int(k.nShared), int(k.nCentre)], dtype=torch.int64)
    gathered = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, mine)                                  # the path's only exchange
    vb, vbn = slabs.bases([(int(g[0]), int(g[1])) for g in gathered])[rank]
    m = emu_emit(sub, iso, d, k, real, vbase=vb, vbase_next=vbn)
    out = [None] * world if rank == 0 else None
    dist.gather_object(dict(V=m.V, N=m.N, T=m.T, vkey=m.vkey, tcell=m.tcell, nShared=int(k.nShared), nCentre=int(k.nCentre)),
                       out, dst=0)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_slab_pipeline_matches_oracle():
    import torch.multiprocessing as mp
    from support import Mesh, merge_slab_meshes, noise_grid, oracle_extract
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    parts = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    meshes = [Mesh(d["V"], d["N"], d["T"], vkey=d["vkey"], tcell=d["tcell"], nShared=d["nShared"], nCentre=d["nCentre"])
              for d in parts]
    m = merge_slab_meshes(meshes)
    whole = oracle_extract(noise_grid(0, "u8", scale=4, shape=(19, 9, 40)), 2.0, "u8")
    assert (m.nV, m.nT) == (whole.nV, whole.nT)
    assert np.array_equal(m.tcell, whole.tcell)
    om, ow = np.argsort(m.vkey, kind="stable"), np.argsort(whole.vkey, kind="stable")
    assert np.array_equal(m.vkey[om], whole.vkey[ow])
    inv = np.empty(m.nV, np.int64); inv[om] = ow
    assert np.array_equal(inv[m.T.astype(np.int64)], whole.T.astype(np.int64))
    assert np.array_equal(m.V[om], whole.V[ow])
