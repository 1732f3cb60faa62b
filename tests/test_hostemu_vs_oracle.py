"""CPU check of the per-word device logic (mc33_core.cuh stepped on the host by
tests/hostemu, test infrastructure) against the oracle, including the z-slab
decomposition used for multi-GPU runs.  No GPU needed; the kernels proper are
covered by the -m gpu tests."""
import numpy as np
import pytest

from mc33_c_library_b200 import slabs
from support import (DTYPES, Geometry, compare_exact, emu_count, emu_emit, emu_extract, gyroid_grid,
                     inclined_geom, make_desc, merge_slab_meshes, noise_grid, oracle_extract)


def _same(o, e):
    compare_exact(o, e, nrm_atol=0)
    assert np.array_equal(o.vkey, e.vkey) and np.array_equal(o.tcell, e.tcell)


@pytest.mark.parametrize("variant,iso,scale", [("f32", 0.0, 0), ("f64", 0.0, 0), ("u8", 3.0, 6), ("u8", 2.5, 6),
                                               ("u16", 500.0, 1000), ("u16", 1.0, 3), ("u32", 2.0, 5)])
def test_noise(variant, iso, scale):
    if variant == "u32":
        a = noise_grid(0, "u16", scale=scale, shape=(14, 19, 70)).astype(np.uint32)
    else:
        a = noise_grid(0, variant, scale=scale, shape=(14, 19, 70))
    _same(oracle_extract(a, iso, variant), emu_extract(a, iso, variant))


@pytest.mark.parametrize("shape", [(2, 2, 2), (3, 2, 33), (2, 5, 34), (9, 4, 65), (5, 5, 32), (4, 3, 97)])
def test_word_boundaries(shape):
    a = noise_grid(0, "u8", scale=4, shape=shape)
    for iso in (1.0, 2.0, 1.5):
        _same(oracle_extract(a, iso, "u8"), emu_extract(a, iso, "u8"))
    f = noise_grid(0, "f32", shape=shape)
    _same(oracle_extract(f, 0.0, "f32"), emu_extract(f, 0.0, "f32"))


@pytest.mark.parametrize("geom", [Geometry(r0=(1, 2, 3), d=(.5, .5, .5)), Geometry(r0=(1, 2, 3), d=(.5, .25, 2)),
                                  inclined_geom(), inclined_geom(tsa=1, d=(.5, .25, 2), r0=(1, 2, 3)),
                                  Geometry(normal_neg=1)])
def test_geometry(geom):
    for variant, iso in (("f32", 0.0), ("u8", 3.0), ("f64", 0.0)):
        a = noise_grid(0, variant, scale=6, shape=(12, 17, 35))
        _same(oracle_extract(a, iso, variant, geom), emu_extract(a, iso, variant, geom))


def test_smooth():
    g = gyroid_grid(48, periods=2)
    for iso in (-0.9, 0.0, 0.3):
        _same(oracle_extract(g, iso), emu_extract(g, iso))


@pytest.mark.parametrize("world", [2, 3, 5, 8])
@pytest.mark.parametrize("variant,iso,scale", [("f32", 0.0, 0), ("u8", 2.0, 4)])
def test_slab_decomposition(world, variant, iso, scale):
    """every slab seam: per-slab extraction with halos + index bases == one slab"""
    _, sdt, real = DTYPES[variant]
    a = noise_grid(0, variant, scale=scale, shape=(23, 9, 40))
    whole = oracle_extract(a, iso, variant)
    parts = [s for s in slabs.partition(a.shape[0] - 1, world) if s is not None]
    descs, datas, counts = [], [], []
    for s in parts:
        sub = np.ascontiguousarray(a[s.z_lo:s.z_hi])
        d = make_desc(a.shape, variant, None, s)
        descs.append(d); datas.append(sub)
        counts.append(emu_count(sub, iso, d))
    b = slabs.bases([(int(k.nV), int(k.nT)) for k in counts])
    # the shared-vertex count a slab computes for its upper halo slice must be the
    # first-slice count of the next slab
    meshes = []
    for s, d, sub, k, (vb, vbn) in zip(parts, descs, datas, counts, b):
        meshes.append(emu_emit(sub, iso, d, k, real, vbase=vb, vbase_next=vbn))
    m = merge_slab_meshes(meshes)
    assert m.nV == whole.nV and m.nT == whole.nT
    assert np.array_equal(m.tcell, whole.tcell)
    # vertex order differs only by slab-major vs global order of centres; keys identify them
    order_m, order_w = np.argsort(m.vkey, kind="stable"), np.argsort(whole.vkey, kind="stable")
    assert np.array_equal(m.vkey[order_m], whole.vkey[order_w])
    inv = np.empty(m.nV, np.int64); inv[order_m] = order_w
    assert np.array_equal(inv[m.T.astype(np.int64)], whole.T.astype(np.int64))
    assert np.array_equal(m.V[order_m], whole.V[order_w])
    assert np.array_equal(m.N[order_m], whole.N[order_w], equal_nan=True)


def test_bit_parallel_triangle_count_matches_case_table():
    """count_simple_cells (k_count): for every one of the 256 indices, alone and packed 32
    per word in random order, complex cells are exactly those without a fixed
    triangulation and the simple ones sum to the table's triangle counts."""
    import ctypes as C
    from support import hostemu_lib
    emu = hostemu_lib()
    emu.mc33emu_count_simple.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32)]
    emu.mc33emu_count_simple.restype = C.c_uint32
    emu.mc33emu_simple256.restype = C.POINTER(C.c_uint16)
    simple = np.ctypeslib.as_array(emu.mc33emu_simple256(), shape=(256,)).copy()
    rng = np.random.default_rng(3)
    words = [np.array([i] * 32) for i in range(256)] + [rng.integers(0, 256, 32) for _ in range(400)]
    for idx in words:
        c = (C.c_uint32 * 8)()
        for k in range(8):                      # corner k <-> index bit 7-k
            c[k] = int(sum(((int(i) >> (7 - k)) & 1) << b for b, i in enumerate(idx)))
        act = int(sum(1 << b for b, i in enumerate(idx) if i not in (0, 255)))
        cx = C.c_uint32()
        nt = emu.mc33emu_count_simple(c, act, C.byref(cx))
        want_cx = int(sum(1 << b for b, i in enumerate(idx) if i not in (0, 255) and simple[i] == 0xFFFF))
        want_nt = int(sum(int(simple[i]) >> 12 for i in idx if i not in (0, 255) and simple[i] != 0xFFFF))
        assert cx.value == want_cx and nt == want_nt


@pytest.mark.parametrize("variant,shape,iso,scale", [("f32", (3, 4, 1100), 0.0, 0), ("u8", (3, 3, 1300), 2.0, 5), ("f32", (3, 3, 2500), 0.2, 0),
                                                      ("u8", (2, 3, 4500), 2.0, 5), ("f32", (4, 5, 129), 0.0, 0), ("u16", (5, 6, 257), 3.0, 7)])
def test_kernel_bodies_on_long_rows(variant, shape, iso, scale):
    """rows that the cell kernel walks in several x-segments (more than 32 words) and rows of more than 32 quads
    (the count kernel's passes with a carry), through the real kernel bodies under the fiber scheduler"""
    a = noise_grid(0, variant, scale=scale, shape=shape)
    _same(oracle_extract(a, iso, variant), emu_extract(a, iso, variant))


def test_kernel_bodies_on_ct_like_volume_with_on_iso_samples_everywhere():
    from support import ct_grid
    c = ct_grid(40)
    for iso in (1500.0, 1500.5, 1234.0):
        _same(oracle_extract(c, iso, "u16"), emu_extract(c, iso, "u16"))


def test_kernel_bodies_plateaus_and_planted_on_iso_samples():
    rng = np.random.default_rng(3)
    a = rng.integers(0, 3, size=(11, 13, 67)).astype(np.uint8)
    for iso in (0.0, 1.0, 2.0):
        _same(oracle_extract(a, iso, "u8"), emu_extract(a, iso, "u8"))
    g = gyroid_grid(40, periods=2).copy()
    iso = np.float32(0.25)
    g.reshape(-1)[rng.integers(0, g.size, size=30)] = iso
    g[0, 0, 0] = iso; g[-1, -1, -1] = iso; g[0, 20, 39] = iso; g[39, 0, 33] = iso
    _same(oracle_extract(g, float(iso)), emu_extract(g, float(iso)))


def test_kernel_bodies_lookback_over_many_blocks():
    """enough rows for hundreds of count blocks: the decoupled look-back walks several 128-block windows"""
    import os
    a = noise_grid(0, "u8", scale=4, shape=(210, 200, 9))
    want = oracle_extract(a, 2.0, "u8")
    _same(want, emu_extract(a, 2.0, "u8"))
    # ... and with the walk forced over aggregates all the way back to block 0 (on the device it normally ends
    # within the first window; the CPU emulation runs the blocks in order, so every prefix is there already)
    os.environ["MC33_EMU_LB_NOPREFIX"] = "1"
    try:
        _same(want, emu_extract(a, 2.0, "u8"))
    finally:
        del os.environ["MC33_EMU_LB_NOPREFIX"]


def test_kernel_bodies_task_free_vertex_kernel(monkeypatch):
    """the alternative vertex kernel body (MC33_B200_VTX=2): vertices straight from the bitmaps, no per-vertex task"""
    monkeypatch.setenv("MC33_EMU_VTX2", "1")
    a = noise_grid(0, "u8", scale=4, shape=(9, 11, 70))
    _same(oracle_extract(a, 2.0, "u8"), emu_extract(a, 2.0, "u8"))
    f = noise_grid(0, "f32", shape=(7, 9, 130))
    _same(oracle_extract(f, 0.0, "f32"), emu_extract(f, 0.0, "f32"))
    g = inclined_geom()
    _same(oracle_extract(f, 0.0, "f32", g), emu_extract(f, 0.0, "f32", g))


@pytest.mark.parametrize("variant,shape,scale", [("f32", (9, 11, 70), 0), ("f32", (6, 5, 129), 0), ("u8", (7, 9, 97), 9)])
def test_row_table_cell_arithmetic_matches_cell_fast(variant, shape, scale, monkeypatch):
    """the direct cell kernel is device-only, its per-cell arithmetic is not: make_rowt + cell_fast_rt (what
    k_emit_cells runs) against cell_fast on every grid point of every visited row, whole grid and z-slabs with halo
    slices (ids of the seam slice in the next slab's space); grids without on-iso samples, as on the kernel's fast path"""
    import ctypes as C
    from support import hostemu_lib
    emu = hostemu_lib()
    emu.mc33emu_rt_stats.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    monkeypatch.setenv("MC33_EMU_CHECK_RT", "1")
    iso = 0.05 if variant == "f32" else 4.5
    a = noise_grid(0, variant, scale=scale, shape=shape) if scale else noise_grid(0, variant, shape=shape)
    n, bad = C.c_uint64(), C.c_uint64()
    emu.mc33emu_rt_stats(C.byref(n), C.byref(bad))          # reset
    whole = emu_extract(a, iso, variant)
    emu.mc33emu_rt_stats(C.byref(n), C.byref(bad))
    assert n.value >= (shape[0] - 1) * shape[1] * shape[2] and bad.value == 0
    _same(oracle_extract(a, iso, variant), whole)
    parts = [s for s in slabs.partition(a.shape[0] - 1, 3) if s is not None]
    _, sdt, real = DTYPES[variant]
    descs = [make_desc(a.shape, variant, None, s) for s in parts]
    subs = [np.ascontiguousarray(a[s.z_lo:s.z_hi]) for s in parts]
    counts = [emu_count(sub, iso, d) for sub, d in zip(subs, descs)]
    b = slabs.bases([(int(k.nV), int(k.nT)) for k in counts])
    emu.mc33emu_rt_stats(C.byref(n), C.byref(bad))          # (the count-only runs do not know the bases: reset)
    for d, sub, k, (vb, vbn) in zip(descs, subs, counts, b):
        emu_emit(sub, iso, d, k, real, vbase=vb, vbase_next=vbn)
    emu.mc33emu_rt_stats(C.byref(n), C.byref(bad))
    assert n.value > 0 and bad.value == 0
