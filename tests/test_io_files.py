"""File functions of the marching_cubes_33.h API against the
compiled, unmodified reference (oracle/_ref): surfaces written by either library are
byte-identical, and grids read by either library from the same file agree sample for
sample.  Host-only: runs in the CPU suite (the reference surface comes from the
reference's own calculate_isosurface)."""
import ctypes as C
import os
import struct
from pathlib import Path

import numpy as np
import pytest

from support import DTYPES, MC33Lib, RefGRD, gyroid_grid, have_ref, noise_grid, ref_lib

ROOT = Path(__file__).resolve().parent.parent
LIBDIR = ROOT / "mc33_c_library_b200" / "lib"

pytestmark = pytest.mark.skipif(not have_ref("f32"), reason="oracle/_ref not built")


def ours(variant):
    return MC33Lib(LIBDIR / f"libMC33_b200_{variant}.so", variant)


def _io_protos(lib):
    L = lib.lib
    for n in ("write_bin_s", "write_txt_s", "write_obj_s"):
        getattr(L, n).argtypes = [C.POINTER(lib.Surface), C.c_char_p]
        getattr(L, n).restype = C.c_int
    L.write_ply_s.argtypes = [C.POINTER(lib.Surface), C.c_char_p, C.c_char_p, C.c_char_p]
    L.write_ply_s.restype = C.c_int
    L.read_bin_s.argtypes = [C.c_char_p]
    L.read_bin_s.restype = C.POINTER(lib.Surface)
    for n in ("read_grd", "read_grd_binary", "read_dat_file"):
        getattr(L, n).argtypes = [C.c_char_p]
        getattr(L, n).restype = C.POINTER(RefGRD)
    L.read_scanfiles.argtypes = [C.c_char_p, C.c_uint, C.c_int]
    L.read_scanfiles.restype = C.POINTER(RefGRD)
    L.read_raw_file.argtypes = [C.c_char_p, C.POINTER(C.c_uint), C.c_int, C.c_int]
    L.read_raw_file.restype = C.POINTER(RefGRD)
    L.adjustvectorlenght_s.argtypes = [C.POINTER(lib.Surface)]
    return L


def _ref_surface(variant, data, iso):
    """a real surface from the reference's calculate_isosurface (kept alive: returns (lib, S, cleanup))"""
    ref = ref_lib(variant)
    G, keep = ref.make_grid(data)
    M = ref.lib.create_MC33(G)
    S = ref.lib.calculate_isosurface(M, ref.real_c(iso))
    assert S and S.contents.nV > 0

    def done():
        ref.lib.free_surface_memory(S); ref.lib.free_MC33(M); ref.lib.free_memory_grd(G)
    return ref, S, done, keep


@pytest.mark.parametrize("variant", ["f32", "f64", "u16"])
def test_surface_files_are_byte_identical_to_the_reference(variant, tmp_path):
    code, sdt, real = DTYPES[variant]
    data = gyroid_grid(20, periods=1.5, dtype=np.float64)
    data = (data * 1000 + 3000).astype(sdt) if variant == "u16" else data.astype(sdt)
    iso = 3100.0 if variant == "u16" else 0.1
    ref, S, done, keep = _ref_surface(variant, data, iso)
    R, O = _io_protos(ref), _io_protos(ours(variant))
    # the surface struct and its malloc'ed arrays are layout-identical, so the SAME surface goes to both writers
    Sours = C.cast(S, C.POINTER(ours(variant).Surface))
    for name, args in (("write_bin_s", ()), ("write_txt_s", ()), ("write_obj_s", ()), ("write_ply_s", (b"me", b"thing")),
                       ("write_ply_s", (None, None))):
        a, b = tmp_path / f"ref_{name}", tmp_path / f"ours_{name}"
        assert getattr(R, name)(S, str(a).encode(), *args) == 0
        assert getattr(O, name)(C.cast(S, C.POINTER(getattr(O, name).argtypes[0]._type_)), str(b).encode(), *args) == 0
        assert a.read_bytes() == b.read_bytes(), name
    # read_bin_s: our reader on the reference's file and the other way round
    for reader, fname in ((O, "ref_write_bin_s"), (R, "ours_write_bin_s")):
        T = reader.read_bin_s(str(tmp_path / fname).encode())
        assert T
        s, t = S.contents, T.contents
        assert (t.nV, t.nT, t.iso) == (s.nV, s.nT, s.iso)
        for fld, n in (("T", s.nT * 12), ("V", s.nV * 3 * np.dtype(real).itemsize), ("N", s.nV * 12), ("color", s.nV * 4)):
            assert C.string_at(getattr(t, fld), n) == C.string_at(getattr(s, fld), n), fld
        reader.free_surface_memory(C.cast(T, reader.free_surface_memory.argtypes[0]))
    del Sours
    done()


def test_read_bin_s_converts_the_other_precision(tmp_path):
    data = gyroid_grid(16, periods=1.0)
    ref, S, done, keep = _ref_surface("f32", data, 0.0)
    R = _io_protos(ref)
    p = tmp_path / "f32.sup"
    assert R.write_bin_s(S, str(p).encode()) == 0
    nV, nT = S.contents.nV, S.contents.nT
    V32 = np.frombuffer(C.string_at(S.contents.V, nV * 12), np.float32).copy()
    done()
    for lib in (ours("f64"), ref_lib("f64")):
        L = _io_protos(lib)
        T = L.read_bin_s(str(p).encode())
        assert T and (T.contents.nV, T.contents.nT) == (nV, nT)
        V64 = np.frombuffer(C.string_at(T.contents.V, nV * 24), np.float64)
        assert np.array_equal(V64, V32.astype(np.float64))
        L.free_surface_memory(T)
    # bad magic / missing file
    (tmp_path / "junk").write_bytes(b"nope" + bytes(64))
    O = _io_protos(ours("f32"))
    assert not O.read_bin_s(str(tmp_path / "junk").encode()) and not O.read_bin_s(b"/nonexistent/file")
    assert O.write_bin_s(None, b"/tmp/x") == -1


def _grid_arrays(G, sdt):
    g = G.contents
    nx, ny, nz = g.N[0] + 1, g.N[1] + 1, g.N[2] + 1
    F = C.cast(g.F, C.POINTER(C.POINTER(C.c_void_p)))
    out = np.empty((nz, ny, nx), sdt)
    for k in range(nz):
        for j in range(ny):
            out[k, j] = np.frombuffer(C.string_at(F[k][j], nx * np.dtype(sdt).itemsize), sdt)
    meta = dict(N=tuple(g.N), r0=tuple(g.r0), d=tuple(g.d), L=tuple(g.L), nonortho=g.nonortho,
                A=np.array([[g._A[j][i] for i in range(3)] for j in range(3)]),
                Ai=np.array([[g.A_[j][i] for i in range(3)] for j in range(3)]))
    return out, meta


def _same_grid(a, b, sdt):
    Fa, ma = _grid_arrays(a, sdt)
    Fb, mb = _grid_arrays(b, sdt)
    assert ma["N"] == mb["N"] and ma["r0"] == mb["r0"] and ma["d"] == mb["d"] and ma["L"] == mb["L"]
    assert ma["nonortho"] == mb["nonortho"]
    assert np.allclose(ma["A"], mb["A"], rtol=0, atol=1e-15) and np.allclose(ma["Ai"], mb["Ai"], rtol=0, atol=1e-13)
    assert np.array_equal(Fa, Fb)
    return Fa, ma


@pytest.mark.parametrize("variant", ["f32", "u16", "f64", "u8"])
def test_raw_and_dat_readers_match_the_reference(variant, tmp_path):
    code, sdt, real = DTYPES[variant]
    R, O = _io_protos(ref_lib(variant)), _io_protos(ours(variant))
    rng = np.random.default_rng(3)
    shape = (5, 6, 7)                       # z, y, x
    N = (C.c_uint * 3)(7, 6, 5)
    cases = [("u8", np.uint8, 1, 0), ("u16le", "<u2", 2, 0), ("u16be", ">u2", -2, 0), ("u32le", "<u4", 4, 0), ("u32be", ">u4", -4, 0),
             ("f32le", "<f4", 4, 1), ("f64le", "<f8", 8, 1), ("f64be", ">f8", -8, 1), ("f32be", ">f4", -4, 1)]
    for name, dt, byte, isfloat in cases:
        hi = 200 if (variant == "u8" or name == "u8") else 60000
        vals = rng.integers(0, hi, size=shape)
        p = tmp_path / f"{name}.raw"
        np.asarray(vals, dtype=dt).tofile(p)
        a = R.read_raw_file(str(p).encode(), N, byte, isfloat)
        b = O.read_raw_file(str(p).encode(), N, byte, isfloat)
        assert a and b, name
        # a byte-swapped file of the float build's own element type: the reference has no conversion branch
        # (MC33_util_grd.c:467-493) and leaves its samples unset; ours is checked against the file alone
        ref_converts = not (variant in ("f32", "f64") and isfloat and byte == -np.dtype(sdt).itemsize)
        if ref_converts:
            Fa, _ = _same_grid(a, b, sdt)
        else:
            Fa, _ = _grid_arrays(b, sdt)
        assert np.array_equal(Fa, np.asarray(vals, dtype=dt).astype(sdt)), name
        assert b.contents.internal_data == 2           # one contiguous block
        R.free_memory_grd(a); O.free_memory_grd(b)
    assert not O.read_raw_file(str(p).encode(), N, 3, 0) and not O.read_raw_file(str(p).encode(), N, 2, 1)
    assert not O.read_raw_file(b"/nonexistent", N, 2, 0)
    # .dat: uint16 header + samples, first slice of the file on top
    vals = rng.integers(0, 200 if variant == "u8" else 4000, size=shape).astype("<u2")
    p = tmp_path / "v.dat"
    p.write_bytes(struct.pack("<3H", 7, 6, 5) + vals.tobytes())
    a, b = R.read_dat_file(str(p).encode()), O.read_dat_file(str(p).encode())
    Fa, _ = _same_grid(a, b, sdt)
    assert np.array_equal(Fa, vals[::-1].astype(sdt))
    R.free_memory_grd(a); O.free_memory_grd(b)


@pytest.mark.parametrize("variant", ["f32", "f64"])
def test_grd_text_and_binary_readers_match_the_reference(variant, tmp_path):
    code, sdt, real = DTYPES[variant]
    # (the strict build of the reference: its -Ofast build turns the header arithmetic d = L / N into
    # a reciprocal multiplication and lands one float ulp off)
    R, O = _io_protos(ref_lib(variant, strict=True)), _io_protos(ours(variant))
    rng = np.random.default_rng(5)
    for order, angles in ((1, (90, 90, 90)), (3, (90, 90, 90)), (1, (80.0, 75.0, 100.0))):
        nx, ny, nz = 4, 3, 2                 # intervals
        vals = rng.normal(size=(nz + 1, ny + 1, nx + 1)).round(4)
        lines = ["a test grid\n", "(1p,e12.5)\n", f"{8.0:8.4f} {6.0:8.4f} {3.0:8.4f} {angles[0]:8.4f} {angles[1]:8.4f} {angles[2]:8.4f}\n",
                 f"{nx:5d} {ny:5d} {nz:5d}\n", f"{order:5d} {0:5d} {nx:5d} {2:5d} {ny + 2:5d} {-1:5d} {nz - 1:5d}\n"]
        body = vals if order == 1 else vals.transpose(0, 2, 1)
        lines += [f"{v:12.5E}\n" for v in body.reshape(-1)]
        p = tmp_path / f"g{order}_{angles[0]}.grd"
        p.write_text("".join(lines))
        a, b = R.read_grd(str(p).encode()), O.read_grd(str(p).encode())
        assert a and b
        Fa, meta = _same_grid(a, b, sdt)
        assert np.allclose(Fa, vals, atol=1e-6) and meta["N"] == (nx, ny, nz)
        assert meta["nonortho"] == int(angles != (90, 90, 90))
        assert a.contents.periodic == b.contents.periodic
        R.free_memory_grd(a); O.free_memory_grd(b)
    # binary "_GRD" files, orthogonal and inclined
    for nono in (0, 1):
        vals = rng.normal(size=(3, 4, 5)).astype(sdt)
        title = b"binary grid"
        blob = struct.pack("<II", 0x4452475F, len(title)) + title + struct.pack("<3I", 4, 3, 2) + struct.pack("<3f", 4, 3, 2) + \
            struct.pack("<3d", 0.5, 0, 0) + struct.pack("<3d", 1, 1.5, 2) + struct.pack("<i", nono)
        if nono:
            A = np.array([[1, .3, .2], [0, .95, .1], [0, 0, .9]])
            blob += struct.pack("<3f", 80, 75, 100) + A.tobytes() + np.linalg.inv(A).tobytes()
        blob += vals.tobytes()
        p = tmp_path / f"b{nono}.bin"
        p.write_bytes(blob)
        a, b = R.read_grd_binary(str(p).encode()), O.read_grd_binary(str(p).encode())
        assert a and b
        Fa, meta = _same_grid(a, b, sdt)
        assert np.array_equal(Fa, vals) and meta["nonortho"] == nono
        R.free_memory_grd(a); O.free_memory_grd(b)
    (tmp_path / "bad.bin").write_bytes(b"XXXX" + bytes(100))
    assert not O.read_grd_binary(str(tmp_path / "bad.bin").encode())


def test_scanfiles_reader_matches_the_reference(tmp_path):
    R, O = _io_protos(ref_lib("u16")), _io_protos(ours("u16"))
    rng = np.random.default_rng(9)
    for nslices, order in ((5, 0), (4, 1)):
        res = 6
        d = tmp_path / f"scan{nslices}"
        d.mkdir()
        vals = rng.integers(0, 4000, size=(nslices, res, res))
        for k in range(nslices):
            np.asarray(vals[k], dtype=">u2" if order else "<u2").tofile(d / f"slice.{k + 3}")
        a = R.read_scanfiles(str(d / "slice.3").encode(), res, order)
        b = O.read_scanfiles(str(d / "slice.3").encode(), res, order)
        assert a and b
        Fa, meta = _same_grid(a, b, np.uint16)
        assert meta["N"] == (res - 1, res - 1, nslices - 1)
        assert np.array_equal(Fa[0], vals[-1]) and np.array_equal(Fa[-1], vals[0])
        R.free_memory_grd(a); O.free_memory_grd(b)
