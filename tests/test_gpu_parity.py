"""GPU parity: the sm_100a kernels, called through the C-ABI (include/mc33cu.h) and
through the marching_cubes_33.h drop-in, against the oracle (oracle/) and the
compiled reference (oracle/_ref, when it travelled with the snapshot).

Bar: triangles bit exact (same triangles, same order, same winding, vertex ids
identical in canonical order), positions bit exact, normals within 1e-6."""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

from support import (DTYPES, Geometry, Mesh, MC33Lib, cfg1_grid, compare_exact, compare_to_reference, ct_grid,
                     gyroid_grid, have_ref, inclined_geom, make_desc, noise_grid, oracle_extract, ref_lib)

pytestmark = pytest.mark.gpu
KATS = json.loads((Path(__file__).parent / "golden" / "kats.json").read_text())
LIBDIR = Path(__file__).resolve().parent.parent / "mc33_c_library_b200" / "lib"


def gpu_extract(data, iso, variant="f32", geom=None, keys=True):
    import torch
    from mc33_c_library_b200.device import Extractor
    code, sdt, real = DTYPES[variant]
    data = np.ascontiguousarray(data, dtype=sdt)
    ex = Extractor(make_desc(data.shape, variant, geom))
    ex.upload(data)
    r = ex.extract(iso, keys=keys)
    ex.close()
    m = Mesh(r["V"], r["N"], r["T"], color=r["color"], vkey=r.get("vkey"), tcell=r.get("tcell"))
    m.counts = dict(nShared=int(r["counts"].nShared), nCentre=int(r["counts"].nCentre))
    return m


def _same(o, g):
    compare_exact(o, g, nrm_atol=1e-6)
    assert np.array_equal(o.vkey.astype(np.int64), g.vkey) and np.array_equal(o.tcell.astype(np.int64), g.tcell)
    assert (g.color == -10724260).all()


def test_extension_is_loaded():
    from mc33_c_library_b200 import _cabi
    lib = _cabi.load()
    assert lib.mc33cu_device_count() >= 1
    assert any("libmc33cu.so" in l for l in open("/proc/self/maps").read().splitlines())


@pytest.mark.parametrize("variant,iso,scale", [("f32", 0.0, 0), ("f32", 0.05, 0), ("f64", 0.0, 0), ("u8", 3.0, 6),
                                               ("u8", 2.5, 6), ("u16", 500.0, 1000), ("u16", 500.5, 1000),
                                               ("u16", 1.0, 3)])
def test_noise_vs_oracle(variant, iso, scale):
    a = noise_grid(64, variant, scale=scale)
    _same(oracle_extract(a, iso, variant), gpu_extract(a, iso, variant))


def test_u32_vs_oracle():
    a = noise_grid(0, "u16", scale=7, shape=(20, 30, 50)).astype(np.uint32)
    _same(oracle_extract(a, 3.0, "u32"), gpu_extract(a, 3.0, "u32"))


@pytest.mark.parametrize("shape", [(2, 2, 2), (3, 2, 33), (2, 5, 34), (9, 4, 65), (5, 5, 32), (4, 3, 97), (30, 42, 38),
                                   (3, 3, 300), (40, 3, 3), (3, 300, 3)])
def test_ragged_shapes(shape):
    a = noise_grid(0, "u8", scale=4, shape=shape)
    for iso in (1.0, 2.0, 1.5):
        _same(oracle_extract(a, iso, "u8"), gpu_extract(a, iso, "u8"))
    f = noise_grid(0, "f32", shape=shape)
    _same(oracle_extract(f, 0.0, "f32"), gpu_extract(f, 0.0, "f32"))


@pytest.mark.parametrize("geom", [Geometry(r0=(1, 2, 3), d=(.5, .5, .5)), Geometry(r0=(1, 2, 3), d=(.5, .25, 2)),
                                  inclined_geom(), inclined_geom(tsa=1, d=(.5, .25, 2), r0=(1, 2, 3)),
                                  Geometry(normal_neg=1)])
def test_store_variants(geom):
    for variant, iso in (("f32", 0.0), ("u8", 3.0), ("f64", 0.0)):
        a = noise_grid(0, variant, scale=6, shape=(24, 33, 47))
        _same(oracle_extract(a, iso, variant, geom), gpu_extract(a, iso, variant, geom))


def test_smooth_and_ct():
    g = gyroid_grid(96, periods=3)
    for iso in (-1.2, -0.3, 0.0, 0.9):
        _same(oracle_extract(g, iso), gpu_extract(g, iso))
    c = ct_grid(64)
    for iso in (1500.0, 1500.5):
        _same(oracle_extract(c, iso, "u16"), gpu_extract(c, iso, "u16"))


def test_sparse_on_iso_samples_mix_fast_and_generic_paths():
    """a smooth float grid with a few samples planted exactly on the isovalue: most row
    groups take the quad fast paths, the ones near a planted sample the generic walk"""
    g = gyroid_grid(80, periods=2).copy()
    rng = np.random.default_rng(11)
    iso = np.float32(0.25)
    idx = rng.integers(0, g.size, size=40)
    g.reshape(-1)[idx] = iso
    # also on the faces / corners of the grid
    g[0, 0, 0] = iso; g[-1, -1, -1] = iso; g[0, 40, 79] = iso; g[79, 0, 33] = iso
    _same(oracle_extract(g, float(iso)), gpu_extract(g, float(iso)))
    d = g.astype(np.float64)
    _same(oracle_extract(d, float(iso), "f64"), gpu_extract(d, float(iso), "f64"))


@pytest.mark.parametrize("variant,shape,iso,scale", [("f32", (5, 9, 128), 0.0, 0), ("f32", (4, 6, 512), 0.1, 0),
                                                      ("f64", (5, 7, 128), 0.0, 0), ("u16", (5, 6, 256), 500.5, 1000),
                                                      ("u16", (4, 5, 512), 3.0, 7), ("u8", (4, 5, 512), 2.5, 6),
                                                      ("u8", (3, 4, 1024), 3.0, 6)])
def test_vector_classify_shapes(variant, shape, iso, scale):
    """row lengths that take the 16-byte vector classify kernel (whole NW-word groups)"""
    a = noise_grid(0, variant, scale=scale, shape=shape)
    _same(oracle_extract(a, iso, variant), gpu_extract(a, iso, variant))


@pytest.mark.parametrize("variant,shape,iso,scale", [("f32", (3, 4, 4200), 0.0, 0), ("u8", (3, 3, 4500), 2.0, 5),
                                                      ("f32", (3, 3, 5000), 0.2, 0), ("f64", (2, 3, 4224), 0.0, 0)])
def test_long_rows(variant, shape, iso, scale):
    """rows of more than 128 words (a warp walks them in passes) and rows longer than a classify chunk"""
    a = noise_grid(0, variant, scale=scale, shape=shape)
    _same(oracle_extract(a, iso, variant), gpu_extract(a, iso, variant))


def test_cfg5_like_inclined_noise():
    """BASELINE config 5 in small: white noise (every ambiguous sub-case) on an inclined grid"""
    a = noise_grid(0, "f32", shape=(40, 37, 128))
    g = inclined_geom()
    _same(oracle_extract(a, 0.0, "f32", g), gpu_extract(a, 0.0, "f32", g))


@pytest.mark.parametrize("device_bases", [False, True])
@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("variant,iso,scale,shape", [("f32", 0.0, 0, (37, 20, 128)), ("u8", 2.0, 4, (23, 9, 40))])
def test_zslabs_on_gpu_match_single_extraction(world, variant, iso, scale, shape, device_bases):
    """BASELINE config 4's decomposition in small: every z-slab through its own context
    (halo slices, ids of the seam slice in the next slab's space), merged == oracle.
    device_bases: the multi-GPU flow of bench.py -- counts stay on the device
    (mc33cu_count_async), are 'all-gathered' (here: stacked), and mc33cu_slab_bases turns
    them into the global vertex bases without a host round trip."""
    import torch
    from mc33_c_library_b200 import slabs
    from mc33_c_library_b200.device import Extractor
    from support import merge_slab_meshes
    a = noise_grid(0, variant, scale=scale, shape=shape)
    whole = oracle_extract(a, iso, variant)
    parts = [s for s in slabs.partition(a.shape[0] - 1, world) if s is not None]
    exs, counts = [], []
    for s in parts:
        ex = Extractor(make_desc(a.shape, variant, None, s))
        ex.upload(np.ascontiguousarray(a[s.z_lo:s.z_hi]))
        exs.append(ex)
        counts.append(ex.count(iso))
    bases = slabs.bases([(int(k.nV), int(k.nT)) for k in counts])
    gathered = None
    if device_bases:
        dev = torch.device("cuda", 0)
        per = [torch.zeros(4, dtype=torch.int32, device=dev) for _ in exs]
        for ex, c4 in zip(exs, per):
            ex.count_async(iso, c4)
        torch.cuda.synchronize()
        gathered = torch.stack(per).contiguous()
    meshes = []
    for r, (ex, k, (vb, vbn)) in enumerate(zip(exs, counts, bases)):
        b = ex.alloc(int(k.nV), int(k.nT), keys=True)
        if device_bases:
            b2 = torch.zeros(2, dtype=torch.int32, device=gathered.device)
            ex.slab_bases(gathered, r, len(exs), b2)
            ex.emit(b, dev_bases=b2)
        else:
            ex.emit(b, vbase=vb, vbase_next=vbn)
        ex.sync()
        nV, nT = int(k.nV), int(k.nT)
        meshes.append(Mesh(b["V"][:nV].cpu().numpy(), b["N"][:nV].cpu().numpy(), b["T"][:nT].cpu().numpy().view(np.uint32),
                           vkey=b["vkey"][:nV].cpu().numpy().astype(np.uint64), tcell=b["tcell"][:nT].cpu().numpy().astype(np.uint64),
                           nShared=int(k.nShared), nCentre=int(k.nCentre)))
        ex.close()
    m = merge_slab_meshes(meshes)
    assert (m.nV, m.nT) == (whole.nV, whole.nT)
    assert np.array_equal(m.tcell, whole.tcell)
    om, ow = np.argsort(m.vkey, kind="stable"), np.argsort(whole.vkey, kind="stable")
    assert np.array_equal(m.vkey[om], whole.vkey[ow])
    inv = np.empty(m.nV, np.int64); inv[om] = ow
    assert np.array_equal(inv[m.T.astype(np.int64)], whole.T.astype(np.int64))
    assert np.array_equal(m.V[om], whole.V[ow])
    fa, fb = np.isfinite(m.N[om]), np.isfinite(whole.N[ow])        # zero gradients give NaN normals on both sides
    assert np.array_equal(fa, fb)
    assert np.abs(np.where(fa, m.N[om], 0) - np.where(fb, whole.N[ow], 0)).max() <= 1e-6


@pytest.mark.parametrize("variant,shape", [("u8", (20, 21, 128)),      # general classify kernel (every Z word rewritten)
                                           ("f32", (12, 13, 256)),     # vector kernel: Z units tracked by dirty bits
                                           ("u8", (6, 7, 512)), ("u16", (6, 9, 256))])
def test_repeated_extractions_reuse_state(variant, shape):
    """iso sweep on one context: bitmaps, on-iso hints, Z dirty bits and prefixes of an earlier isovalue
    must not leak (integer-valued samples: on-iso samples come and go with the isovalue)"""
    from mc33_c_library_b200.device import Extractor
    code, sdt, real = DTYPES[variant]
    a = noise_grid(0, "u8", scale=5, shape=shape).astype(sdt)
    ex = Extractor(make_desc(a.shape, variant))
    ex.upload(a)
    for iso in (2.0, 2.5, 1.0, 3.5, 2.0, 0.5, 3.0, 3.0, 1.5):
        r = ex.extract(iso, keys=True)
        g = Mesh(r["V"], r["N"], r["T"], color=r["color"], vkey=r["vkey"], tcell=r["tcell"])
        _same(oracle_extract(a, iso, variant), g)
    # the same context through the sweep sets and back
    isos = [3.0, 2.0, 2.5]
    for _ in range(2):
        ex.classify_sweep(isos)
        for j, iso in enumerate(isos):
            want = oracle_extract(a, iso, variant)
            b = ex.alloc(want.nV + 8, want.nT + 8, keys=True)
            ex.extract_set_async(j, b)
            k = ex.sync()
            assert (int(k.nV), int(k.nT)) == (want.nV, want.nT)
            assert np.array_equal(b["T"][:want.nT].cpu().numpy().view(np.uint32), want.T)
        isos = [1.0, 3.5, 0.5]
    r = ex.extract(2.0, keys=True)
    _same(oracle_extract(a, 2.0, variant), Mesh(r["V"], r["N"], r["T"], color=r["color"], vkey=r["vkey"], tcell=r["tcell"]))
    ex.close()


def test_plateaus_and_empty():
    rng = np.random.default_rng(3)
    a = rng.integers(0, 3, size=(20, 22, 67)).astype(np.uint8)
    for iso in (0.0, 1.0, 2.0):
        _same(oracle_extract(a, iso, "u8"), gpu_extract(a, iso, "u8"))
    z = np.zeros((5, 6, 7), np.float32)
    for iso in (1.0, -1.0, 0.0):
        g = gpu_extract(z, iso)
        assert g.nV == 0 and g.nT == 0


def test_kats_counts():
    """BASELINE.md K1..K6 known answers of the reference, counts through the GPU count pass"""
    from mc33_c_library_b200.device import Extractor

    def count(data, iso, variant, geom=None):
        ex = Extractor(make_desc(data.shape, variant, geom))
        ex.upload(np.ascontiguousarray(data))
        k = ex.count(iso)
        ex.close()
        return int(k.nV), int(k.nT), int(k.nCentre)
    F, geom = cfg1_grid()
    assert count(F, 0.0, "f32", geom)[:2] == (KATS["K1"]["nV"], KATS["K1"]["nT"])
    assert count(F, 0.5, "f32", geom)[:2] == (KATS["K2"]["nV"], KATS["K2"]["nT"])
    assert count(noise_grid(128, "f32"), 0.0, "f32")[:2] == (KATS["K3"]["nV"], KATS["K3"]["nT"])
    assert count(noise_grid(256, "f32"), 0.0, "f32")[:2] == (KATS["K4"]["nV"], KATS["K4"]["nT"])
    u = noise_grid(128, "u16", scale=1000)
    assert count(u, 500.0, "u16")[:2] == (KATS["K5a"]["nV"], KATS["K5a"]["nT"])
    assert count(u, 500.5, "u16")[:2] == (KATS["K5b"]["nV"], KATS["K5b"]["nT"])
    nV, nT, nC = count(noise_grid(128, "u8", scale=6), 3.0, "u8")
    assert (nV, nT, nC) == (KATS["K6"]["nV"], KATS["K6"]["nT"], KATS["K6"]["nCentre"])


def test_cfg1_full_mesh_vs_oracle():
    F, geom = cfg1_grid()
    _same(oracle_extract(F, 0.0, "f32", geom), gpu_extract(F, 0.0, "f32", geom))


def test_large_noise_vs_oracle():
    a = noise_grid(160, "f32")
    _same(oracle_extract(a, 0.0), gpu_extract(a, 0.0))


def test_full_size_properties():
    """512^3 gyroid (BASELINE config 2): size independent invariants + the
    reference's counts from BASELINE.md section 2"""
    import torch
    from mc33_c_library_b200.device import Extractor
    g = gyroid_grid(512, periods=4)
    ex = Extractor(make_desc(g.shape, "f32"))
    ex.upload(g)
    for iso, key in ((-1.2, "G512_-1.2"), (0.0, "G512_0.0")):
        r = ex.extract(iso, keys=True)
        k = r["counts"]
        nV, nT = int(k.nV), int(k.nT)
        assert (nV, nT) == (KATS[key]["nV"], KATS[key]["nT"])
        T = r["T"].astype(np.int64)
        assert T.max() == nV - 1 and T.min() == 0
        assert (T[:, 0] != T[:, 1]).all() and (T[:, 1] != T[:, 2]).all() and (T[:, 0] != T[:, 2]).all()
        assert np.bincount(T.reshape(-1), minlength=nV).min() >= 1          # every vertex used
        assert (np.diff(r["tcell"]) >= 0).all()                              # cell-major order
        assert (np.diff(r["vkey"][: int(k.nShared)] // 4 // g.shape[2]) >= 0).all()   # row-major order
        n = r["N"].astype(np.float64)
        assert np.abs(np.sqrt((n * n).sum(1)) - 1).max() < 1e-6
        # closed manifold away from the boundary: every edge is shared by two triangles
        e = np.sort(np.concatenate([T[:, [0, 1]], T[:, [1, 2]], T[:, [2, 0]]]), axis=1)
        _, cnt = np.unique(e[:, 0] * nV + e[:, 1], return_counts=True)
        assert cnt.max() == 2
    ex.close()


def _mesh_properties(r, nV, nT, nShared, NX, check_used=True):
    T = r["T"].astype(np.int64)
    assert T.shape == (nT, 3) and T.min() == 0 and T.max() == nV - 1
    assert (T[:, 0] != T[:, 1]).all() and (T[:, 1] != T[:, 2]).all() and (T[:, 0] != T[:, 2]).all()
    assert (np.diff(r["tcell"]) >= 0).all()                                  # sweep (cell-major) order
    assert (np.diff(r["vkey"][:nShared] // 4 // NX) >= 0).all()              # shared vertices in point-row order
    assert (np.diff(r["vkey"][nShared:]) > 0).all()                          # centre vertices in cell order
    if check_used:
        assert np.bincount(T.reshape(-1), minlength=nV).min() >= 1
    n = r["N"].astype(np.float64)
    ln = np.sqrt((n * n).sum(1))
    ok = np.isfinite(ln)
    assert np.abs(ln[ok] - 1).max() < 1e-5


def test_cfg3_like_large_u16_volume():
    """BASELINE config 3 at 1024 x 1024 x 160 uint16 (GRD_INTEGER, GRD_TYPE_SIZE 2): CT-like blobs + noise,
    integer isovalue (on-iso samples everywhere -> generic path) and half-integer (fast path);
    counts against the compiled reference's size_of_isosurface, size-independent mesh properties"""
    import torch
    from mc33_c_library_b200.device import Extractor
    NZ, NY, NX = 160, 1024, 1024
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(5)
    ax = torch.linspace(-1, 1, NX, device=dev)
    az = torch.linspace(-0.3, 0.3, NZ, device=dev)
    v = torch.full((NZ, NY, NX), 1000.0, device=dev)
    for c0, c1, c2, sg in ((0.1, -0.2, 0.05, 0.3), (-0.4, 0.3, -0.1, 0.25), (0.5, 0.5, 0.1, 0.2)):
        v += 2500.0 / 3 * torch.exp(-((ax[None, None, :] - c0) ** 2 + (ax[None, :, None] - c1) ** 2 + (az[:, None, None] - c2) ** 2) / (2 * sg * sg))
    v += torch.randint(0, 16, v.shape, device=dev, generator=g)
    vol = v.clamp(0, 65535).to(torch.int32).to(torch.uint16)
    del v
    host = vol.cpu().numpy()
    ex = Extractor(make_desc(host.shape, "u16"))
    ex.bind(vol)
    for iso in (1500.0, 1500.5):
        r = ex.extract(iso, keys=True)
        k = r["counts"]
        nV, nT = int(k.nV), int(k.nT)
        if have_ref("u16"):
            _, rV, rT = ref_lib("u16").size(host, iso)
            assert (nV, nT) == (rV, rT)
        assert nV > 100000
        _mesh_properties(r, nV, nT, int(k.nShared), NX, check_used=False)
    ex.close()


def test_cfg5_like_large_inclined_noise():
    """BASELINE config 5 at 320^3: float white noise on an inclined grid (dense: ~1.5 vertices and ~3.5
    triangles per voxel, half the cells through a face / interior test)"""
    import torch
    from mc33_c_library_b200.device import Extractor
    n = 320
    a = noise_grid(n, "f32")
    geom = inclined_geom()
    ex = Extractor(make_desc(a.shape, "f32", geom))
    ex.upload(a)
    r = ex.extract(0.0, keys=True)
    k = r["counts"]
    nV, nT = int(k.nV), int(k.nT)
    assert 1.4 < nV / n ** 3 < 1.7 and 3.2 < nT / n ** 3 < 3.7
    if have_ref("f32"):
        _, rV, rT = ref_lib("f32").size(a, 0.0, geom)
        assert (nV, nT) == (rV, rT)
    _mesh_properties(r, nV, nT, int(k.nShared), n)
    # a 40-slice sub-volume against the oracle, vertex by vertex
    sub = np.ascontiguousarray(a[:40])
    _same(oracle_extract(sub, 0.0, "f32", geom), gpu_extract(sub, 0.0, "f32", geom))
    ex.close()


# ---- the drop-in C API (include/marching_cubes_33.h) --------------------------
def dropin(variant):
    return MC33Lib(LIBDIR / f"libMC33_b200_{variant}.so", variant)


@pytest.mark.parametrize("variant,iso,scale", [("f32", 0.0, 0), ("f64", 0.0, 0), ("u8", 3.0, 6), ("u16", 500.0, 1000)])
def test_dropin_api_vs_oracle_and_reference(variant, iso, scale):
    a = noise_grid(48, variant, scale=scale)
    lib = dropin(variant)
    for geom in (None, Geometry(r0=(1, 2, 3), d=(.5, .25, 2)), inclined_geom()):
        mine = lib.extract(a, iso, geom)
        orc = oracle_extract(a, iso, variant, geom)
        compare_exact(orc, mine, nrm_atol=1e-6)
        assert (mine.color == -10724260).all() and mine.iso == pytest.approx(iso)
        if have_ref(variant):
            ref = ref_lib(variant).extract(a, iso, geom)
            compare_to_reference(ref, mine, exact_pos=geom is None or not geom.nonortho, pos_rtol=2e-6)
    sz, nV, nT = lib.size(a, iso)
    assert (nV, nT) == (orc.nV, orc.nT)
    assert sz == nV * (6 * (8 if variant == "f64" else 4) + 4) + nT * 12 + 64


def test_dropin_generate_grid_from_fn_cfg1():
    """BASELINE config 1 end to end through the C API: generate_grid_from_fn with a
    C callback (libm cos), create_MC33, calculate_isosurface -> K1"""
    lib = dropin("f32")
    libm = C.CDLL("libm.so.6")
    libm.cos.restype = C.c_double
    libm.cos.argtypes = [C.c_double]
    FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double, C.c_double)
    fn = FN(lambda x, y, z: libm.cos(x) + libm.cos(y) + libm.cos(z))
    L = lib.lib
    L.generate_grid_from_fn.argtypes = [C.c_double] * 9 + [FN]
    G = L.generate_grid_from_fn(-4, -4, -4, 4, 4, 4, .04, .04, .04, fn)
    assert G and tuple(G.contents.N) == (200, 200, 200) and G.contents.internal_data == 1
    M = L.create_MC33(G)
    S = L.calculate_isosurface(M, C.c_float(0.0))
    assert (S.contents.nV, S.contents.nT) == (KATS["K1"]["nV"], KATS["K1"]["nT"])
    L.free_surface_memory(S); L.free_MC33(M); L.free_memory_grd(G)


def test_dropin_empty_surface_and_null_safety():
    lib = dropin("f32")
    a = np.zeros((4, 5, 6), np.float32)
    m = lib.extract(a, 1.0)
    assert m.nV == 0 and m.nT == 0 and m.iso == 0.0   # zeroed struct (reference c:1880-1883)
    L = lib.lib
    L.free_surface_memory(None); L.free_MC33(None); L.free_memory_grd(None)
    assert not L.create_MC33(None)
    assert not L.grid_from_data_pointer(0, 4, 4, a.ctypes.data)


@pytest.mark.parametrize("variant,shape,scale,isos", [
    # float rows of whole 128-sample groups: the one-pass sweep kernel; integer-valued samples
    # and integer isovalues put on-iso samples in every set
    ("f32", (9, 12, 256), 6, [1.0, 2.0, 2.5, 3.0, 4.0, 0.5, 5.0, 3.5]),
    ("f32", (7, 10, 128), 0, [-0.5, -0.2, 0.0, 0.1, 0.3]),
    # other shapes / element types: one classify launch per set
    ("f32", (8, 9, 70), 0, [-0.3, 0.0, 0.2]),
    ("u8", (8, 9, 128), 6, [2.0, 2.5, 3.0]),
])
def test_classify_sweep_matches_single_extractions(variant, shape, scale, isos):
    """mc33cu_classify_sweep + mc33cu_extract_set_device == one mc33cu_extract_device per isovalue
    (== the oracle), sets taken in scrambled order, with a normal extraction in between"""
    import torch
    from mc33_c_library_b200.device import Extractor
    code, sdt, real = DTYPES[variant]
    if variant == "f32" and scale:       # integer-valued floats: many samples exactly on the integer isovalues
        a = noise_grid(0, "u8", scale=scale, shape=shape).astype(np.float32)
    else:
        a = noise_grid(0, variant, scale=scale, shape=shape) if scale else noise_grid(0, variant, shape=shape)
    ex = Extractor(make_desc(a.shape, variant, None))
    ex.upload(np.ascontiguousarray(a, dtype=sdt))
    ex.classify_sweep(isos)
    order = list(range(len(isos)))[::-1]
    for n_done, j in enumerate(order):
        want = oracle_extract(a, isos[j], variant)
        b = ex.alloc(want.nV + 8, want.nT + 8, keys=True)
        ex.extract_set_async(j, b)
        k = ex.sync()
        nV, nT = int(k.nV), int(k.nT)
        assert (nV, nT) == (want.nV, want.nT)
        g = Mesh(b["V"][:nV].cpu().numpy(), b["N"][:nV].cpu().numpy(), b["T"][:nT].cpu().numpy().view(np.uint32),
                 color=b["color"][:nV].cpu().numpy(), vkey=b["vkey"][:nV].cpu().numpy(), tcell=b["tcell"][:nT].cpu().numpy())
        _same(want, g)
    # the single-isovalue path still works on the same context afterwards, and the sets survive it
    r = ex.extract(isos[0], keys=True)
    want = oracle_extract(a, isos[0], variant)
    assert (int(r["counts"].nV), int(r["counts"].nT)) == (want.nV, want.nT)
    b = ex.alloc(want.nV + 8, want.nT + 8, keys=True)
    ex.extract_set_async(0, b)
    k = ex.sync()
    assert (int(k.nV), int(k.nT)) == (want.nV, want.nT)
    assert np.array_equal(b["T"][:want.nT].cpu().numpy().view(np.uint32), want.T)
    ex.close()


@pytest.mark.parametrize("nstreams", [2, 6])
def test_sweep_sets_on_several_streams(nstreams):
    """the sets of one sweep extracted side by side on several CUDA streams of ONE context (bench.py's cfg2
    schedule): the context keeps a vertex-task buffer per stream (four slots: with six streams slots change
    hands), mc33cu_sync waits for all of them and reports every set's counts; meshes == the oracle's"""
    import torch
    from mc33_c_library_b200.device import Extractor
    isos = [-0.6, -0.3, -0.1, 0.0, 0.1, 0.2, 0.4, 0.6]
    a = gyroid_grid(128, periods=3)
    ex = Extractor(make_desc(a.shape, "f32", None))
    ex.upload(a)
    wants = [oracle_extract(a, v, "f32") for v in isos]
    capV, capT = max(w.nV for w in wants) + 8, max(w.nT for w in wants) + 8
    main = torch.cuda.current_stream()
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    bufs = [ex.alloc(capV, capT, keys=True) for _ in isos]
    for rep in range(2):
        ex.use_stream(main)
        ex.classify_sweep(isos)
        ev = main.record_event()
        for j in range(len(isos)):
            s = streams[j % nstreams]
            s.wait_event(ev)
            ex.use_stream(s)
            ex.extract_set_async(j, bufs[j])
        ex.use_stream(main)
        ex.sync()                       # waits for the extractions on the other streams too
        for j, (want, b) in enumerate(zip(wants, bufs)):
            nV, nT = want.nV, want.nT
            g = Mesh(b["V"][:nV].cpu().numpy(), b["N"][:nV].cpu().numpy(), b["T"][:nT].cpu().numpy().view(np.uint32),
                     color=b["color"][:nV].cpu().numpy(), vkey=b["vkey"][:nV].cpu().numpy(), tcell=b["tcell"][:nT].cpu().numpy())
            _same(want, g)
    ex.close()


def test_sweep_sets_counted_first_then_emitted_across_slabs():
    """bench.py's multi-GPU sweep flow on one GPU: every slab classifies the sweep once and counts ALL its
    sets, the counts are 'all-gathered' as [slab][set][4], mc33cu_slab_bases_strided picks a set's column,
    and the sets are emitted afterwards from their own count state; merged meshes == oracle"""
    import torch
    from mc33_c_library_b200 import slabs
    from mc33_c_library_b200.device import Extractor
    from support import merge_slab_meshes
    variant, isos = "f32", [2.0, 3.0, 2.5, 1.0]
    a = noise_grid(0, "u8", scale=5, shape=(21, 10, 128)).astype(np.float32)
    parts = [s for s in slabs.partition(a.shape[0] - 1, 3) if s is not None]
    dev = torch.device("cuda", 0)
    exs = []
    for s in parts:
        ex = Extractor(make_desc(a.shape, variant, None, s))
        ex.upload(np.ascontiguousarray(a[s.z_lo:s.z_hi]))
        exs.append(ex)
    per = [torch.zeros((len(isos), 4), dtype=torch.int32, device=dev) for _ in exs]
    for ex, c in zip(exs, per):
        ex.classify_sweep(isos)
        for j in range(len(isos)):
            ex.count_set_async(j, c[j])
    torch.cuda.synchronize()
    gathered = torch.stack(per).contiguous()                  # [slab][set][4]
    for j in (2, 0, 3, 1):
        whole = oracle_extract(a, isos[j], variant)
        meshes = []
        for r, ex in enumerate(exs):
            nV, nT, nS, nC = (int(v) for v in gathered[r, j].tolist())
            b = ex.alloc(nV, nT, keys=True)
            b2 = torch.zeros(2, dtype=torch.int32, device=dev)
            ex.slab_bases_strided(gathered[0, j], 4 * len(isos), r, len(exs), b2)
            ex.emit_set(j, b, dev_bases=b2)
            ex.sync()
            meshes.append(Mesh(b["V"][:nV].cpu().numpy(), b["N"][:nV].cpu().numpy(), b["T"][:nT].cpu().numpy().view(np.uint32),
                               vkey=b["vkey"][:nV].cpu().numpy().astype(np.uint64), tcell=b["tcell"][:nT].cpu().numpy().astype(np.uint64),
                               nShared=nS, nCentre=nC))
        m = merge_slab_meshes(meshes)
        assert (m.nV, m.nT) == (whole.nV, whole.nT)
        assert np.array_equal(m.tcell, whole.tcell)
        om, ow = np.argsort(m.vkey, kind="stable"), np.argsort(whole.vkey, kind="stable")
        assert np.array_equal(m.vkey[om], whole.vkey[ow])
        inv = np.empty(m.nV, np.int64); inv[om] = ow
        assert np.array_equal(inv[m.T.astype(np.int64)], whole.T.astype(np.int64))
        assert np.array_equal(m.V[om], whole.V[ow])
    for ex in exs:
        ex.close()


def test_output_capacity_is_respected():
    """too small output buffers: mc33cu_sync reports MC33CU_ERR_CAPACITY and nothing is written past
    the stated capacities (vertices, vertex tasks, triangles, centre vertices)"""
    import torch
    from mc33_c_library_b200 import _cabi as cabi
    from mc33_c_library_b200.device import Extractor
    a = noise_grid(0, "f32", shape=(24, 25, 128))
    ex = Extractor(make_desc(a.shape, "f32"))
    ex.upload(a)
    k = ex.count(0.0)
    nV, nT = int(k.nV), int(k.nT)
    assert nV > 100 and nT > 100
    for capV, capT in ((nV // 2, nT), (nV, nT // 2), (nV // 3, nT // 3), (int(k.nShared) + 1, nT)):
        b = ex.alloc(nV, nT)
        for t in (b["V"], b["N"]):
            t.fill_(-77.0)
        b["color"].fill_(-77); b["T"].fill_(-77)
        b["capV"], b["capT"] = capV, capT
        ex.emit(b)
        with pytest.raises(cabi.Mc33CudaError) as e:
            ex.sync()
        assert e.value.code == cabi.ERR_CAPACITY
        assert bool((b["V"][capV:] == -77.0).all()) and bool((b["N"][capV:] == -77.0).all())
        assert bool((b["color"][capV:] == -77).all()) and bool((b["T"][capT:] == -77).all())
    b = ex.alloc(nV, nT)                       # exact capacities still work afterwards
    ex.emit(b)
    ex.sync()
    ex.close()


def test_sweep_api_state_errors():
    import ctypes as C
    import torch
    from mc33_c_library_b200 import _cabi as cabi
    from mc33_c_library_b200.device import Extractor
    a = noise_grid(0, "f32", shape=(6, 7, 128))
    ex = Extractor(make_desc(a.shape, "f32"))
    b = ex.alloc(16, 16)
    with pytest.raises(cabi.Mc33CudaError) as e:            # no grid bound yet
        ex.classify_sweep([0.0])
    assert e.value.code == cabi.ERR_STATE
    ex.upload(a)
    with pytest.raises(cabi.Mc33CudaError) as e:            # no pre-classified set yet
        ex.extract_set_async(0, b)
    assert e.value.code == cabi.ERR_STATE
    with pytest.raises(cabi.Mc33CudaError) as e:            # more than 8 isovalues
        ex.classify_sweep([0.1 * i for i in range(9)])
    assert e.value.code == cabi.ERR_ARG
    ex.classify_sweep([0.0, 0.2])
    with pytest.raises(cabi.Mc33CudaError) as e:            # set index beyond the sweep
        ex.extract_set_async(2, b)
    assert e.value.code == cabi.ERR_STATE
    c4 = torch.zeros(4, dtype=torch.int32, device="cuda")
    ex.count_set_async(1, c4)
    torch.cuda.synchronize()
    want = oracle_extract(a, 0.2, "f32")
    assert (int(c4[0]), int(c4[1])) == (want.nV, want.nT)
    ex.close()


def test_sync_reports_overflow_of_any_set_emitted_since_the_last_sync():
    """count every set, emit set 0 into too small buffers and set 1 into large enough ones, sync ONCE:
    the overflow of set 0 must not be hidden by the later emit (round-1 advisor finding)"""
    import torch
    from mc33_c_library_b200 import _cabi as cabi
    from mc33_c_library_b200.device import Extractor
    a = noise_grid(0, "f32", shape=(12, 13, 128))
    ex = Extractor(make_desc(a.shape, "f32"))
    ex.upload(a)
    isos = [0.0, 0.3]
    ex.classify_sweep(isos)
    c4 = torch.zeros((2, 4), dtype=torch.int32, device="cuda")
    for j in range(2):
        ex.count_set_async(j, c4[j])
    torch.cuda.synchronize()
    nV, nT = int(c4[:, 0].max()), int(c4[:, 1].max())
    small, big = ex.alloc(nV, nT), ex.alloc(nV, nT)
    small["capV"], small["capT"] = nV // 4, nT // 4
    ex.emit_set(0, small)
    ex.emit_set(1, big)
    with pytest.raises(cabi.Mc33CudaError) as e:
        ex.sync()
    assert e.value.code == cabi.ERR_CAPACITY
    # the same two emits with room for both: clean
    ex.emit_set(0, big); ex.sync()
    want = oracle_extract(a, isos[0], "f32")
    assert np.array_equal(big["T"][:want.nT].cpu().numpy().view(np.uint32), want.T)
    ex.emit_set(1, big); ex.sync()
    # a new sweep classify invalidates the sets' counts: emit without a new count is a state error, not a wrong mesh
    ex.classify_sweep([0.1, 0.2])
    with pytest.raises(cabi.Mc33CudaError) as e:
        ex.emit_set(0, big)
    assert e.value.code == cabi.ERR_STATE
    ex.count_set_async(0, c4[0])
    ex.emit_set(0, big)
    ex.sync()
    want = oracle_extract(a, 0.1, "f32")
    assert np.array_equal(big["T"][:want.nT].cpu().numpy().view(np.uint32), want.T)
    # counts fetched lazily after an *_async count (the host mirror is not stale)
    k = ex.sync()
    assert (int(k.nV), int(k.nT)) == (want.nV, want.nT)
    ex.close()


@pytest.mark.parametrize("variant,iso,scale,shape,nslab", [("f32", 0.0, 0, (41, 20, 128), 3), ("u8", 2.0, 4, (23, 9, 40), 2),
                                                           ("f32", 0.05, 0, (70, 33, 65), 8), ("u16", 500.0, 1000, (19, 17, 96), 4)])
def test_dropin_multi_slab_matches_single_slab_and_the_oracle(variant, iso, scale, shape, nslab, monkeypatch):
    """calculate_isosurface with the grid cut into z-slabs, one context each (MC33_B200_DEVICES names a device per
    slab; on a one-GPU box they all sit on device 0, on a multi-GPU box the same code runs one slab per GPU):
    the mesh equals the one-slab drop-in mesh and the oracle's after canonical ordering by key-free comparison:
    triangles in the same (sweep) order, vertex numbering by slab"""
    a = noise_grid(0, variant, scale=scale, shape=shape)
    want = oracle_extract(a, iso, variant)
    monkeypatch.setenv("MC33_B200_DEVICES", "0")
    one = dropin(variant).extract(a, iso)
    compare_exact(want, one, nrm_atol=1e-6)
    import torch
    ndev = max(1, torch.cuda.device_count())
    monkeypatch.setenv("MC33_B200_DEVICES", ",".join(str(i % ndev) for i in range(nslab)))      # one slab per GPU where there are several
    lib = dropin(variant)
    many = lib.extract(a, iso)
    assert lib.lib.mc33_dropin_slabs_last() == nslab
    assert lib.lib.mc33_dropin_gpus_last() == min(nslab, ndev)
    assert (many.nV, many.nT) == (one.nV, one.nT)
    # same triangles in the same order; the vertex numbering differs (per slab: shared vertices, then centres), so
    # corner k of triangle j induces the id map, which must be a bijection carrying positions / normals / colours over
    fwd = np.full(one.nV, -1, np.int64)
    fwd[one.T.reshape(-1).astype(np.int64)] = many.T.reshape(-1).astype(np.int64)
    used = fwd >= 0
    assert np.array_equal(fwd[one.T.reshape(-1).astype(np.int64)], many.T.reshape(-1).astype(np.int64))
    assert len(np.unique(fwd[used])) == used.sum()
    assert np.array_equal(one.V[used], many.V[fwd[used]])
    fa = np.isfinite(one.N[used])
    assert np.array_equal(np.where(fa, one.N[used], 0), np.where(fa, many.N[fwd[used]], 0))
    assert (many.color == -10724260).all()
    # vertices no triangle refers to (on-iso points whose triangles were all dropped) still exist on both sides
    assert sorted(map(tuple, np.round(one.V[~used].astype(np.float64), 5))) == \
        sorted(map(tuple, np.round(np.delete(many.V, fwd[used], axis=0).astype(np.float64), 5)))
    sz, nV, nT = lib.size(a, iso)
    assert (nV, nT) == (one.nV, one.nT)


def _same_mesh_up_to_slab_numbering(one, many):
    assert (many.nV, many.nT) == (one.nV, one.nT)
    if one.nT == 0:
        return
    fwd = np.full(one.nV, -1, np.int64)
    fwd[one.T.reshape(-1).astype(np.int64)] = many.T.reshape(-1).astype(np.int64)
    used = fwd >= 0
    assert np.array_equal(fwd[one.T.reshape(-1).astype(np.int64)], many.T.reshape(-1).astype(np.int64))
    assert len(np.unique(fwd[used])) == used.sum()
    assert np.array_equal(one.V[used], many.V[fwd[used]])
    fa = np.isfinite(one.N[used])
    assert np.array_equal(np.where(fa, one.N[used], 0), np.where(fa, many.N[fwd[used]], 0))
    assert (many.color == -10724260).all()


@pytest.mark.parametrize("spec,chunks", [("1.5", 4), ("0", 4), ("0.4", 5), ("1.5", 1), ("3", 7)])
def test_dropin_chunked_upload_and_result_arrays_sized_from_an_estimate(spec, chunks, monkeypatch):
    """MC33_B200_CHUNKS z-chunks per GPU: a chunk is emitted as soon as it has been counted, into result arrays sized
    before the counts are complete (from the slabs counted so far / the meshes this MC33 produced before).  Head room 0
    waits for all the counts (the round-1 flow), 0.4 makes every estimate too small (the arrays are replaced by exact
    ones and the part already downloaded is carried over); the sphere puts nothing into the first chunks, so the
    extrapolation starts late.  Several isovalues on ONE MC33, growing and shrinking meshes."""
    a = noise_grid(0, "f32", shape=(40, 24, 97))
    zz, yy, xx = np.meshgrid(*(np.arange(n, dtype=np.float32) for n in a.shape), indexing="ij")
    sphere = (np.sqrt((zz - 20) ** 2 + (yy - 12) ** 2 + (xx - 48) ** 2) - 9.0).astype(np.float32)
    isos = [0.3, -0.2, 0.0, 0.9, 0.0]
    monkeypatch.setenv("MC33_B200_DEVICES", "0")
    lib1 = dropin("f32")
    ones = lib1.extract_many(a, isos) + lib1.extract_many(sphere, [0.0, 5.0])
    for m, iso in zip(ones[:2], isos[:2]):
        compare_exact(oracle_extract(a, iso, "f32"), m, nrm_atol=1e-6)
    monkeypatch.delenv("MC33_B200_DEVICES")
    monkeypatch.setenv("MC33_B200_GPUS", "1")
    monkeypatch.setenv("MC33_B200_CHUNKS", str(chunks))
    monkeypatch.setenv("MC33_B200_SPECULATE", spec)
    lib = dropin("f32")
    manys = lib.extract_many(a, isos) + lib.extract_many(sphere, [0.0, 5.0])
    assert lib.lib.mc33_dropin_slabs_last() == chunks and lib.lib.mc33_dropin_gpus_last() == 1
    for one, many in zip(ones, manys):
        _same_mesh_up_to_slab_numbering(one, many)
        if spec == "0" or chunks == 1:
            assert (many.capv, many.capt) == (many.nV, max(many.nT, 1))
