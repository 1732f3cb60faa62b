"""Shared test support: ctypes bindings for the compiled reference (oracle/_ref),
the CPU restatement (oracle/_build) and mesh comparison helpers.

TEST INFRASTRUCTURE ONLY -- nothing in mc33_c_library_b200/ imports this.
Nothing here reads /root/reference at run time; the reference shared objects
are prebuilt by `make -C oracle ref` (or __graft_entry__.build()) and travel
with the snapshot.
"""
import ctypes as C
import math
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
REF_DIR = ORACLE_DIR / "_ref"
ORACLE_SO = ORACLE_DIR / "_build" / "libmc33_oracle.so"

import sys
sys.path.insert(0, str(ROOT))
from mc33_c_library_b200.dropin import DTYPES, SPN0, SPNA, SPNB, SPNC  # noqa: E402


def ensure_oracle_built():
    if not ORACLE_SO.exists():
        subprocess.check_call(["make", "-C", str(ORACLE_DIR), "oracle"])
    return ORACLE_SO


def have_ref(variant="f32", strict=False):
    return (REF_DIR / f"libMC33_ref_{variant}{'_strict' if strict else ''}.so").exists()


from mc33_c_library_b200.dropin import Geometry, MC33Lib, Mesh, RefGRD, _np_from, _surface_struct  # noqa: E402,F401


# --------------------------------------------------------------------------
# oracle (CPU restatement)
# --------------------------------------------------------------------------
class OGeom(C.Structure):
    _fields_ = [("store", C.c_int), ("normal_neg", C.c_int), ("tsa", C.c_int), ("pad_", C.c_int),
                ("O", C.c_double * 3), ("D", C.c_double * 3), ("ca", C.c_double), ("cb", C.c_double),
                ("A", C.c_double * 9), ("Ai", C.c_double * 9)]


class OMesh(C.Structure):
    _fields_ = [("nV", C.c_uint64), ("nT", C.c_uint64), ("nShared", C.c_uint64), ("nCentre", C.c_uint64),
                ("nPoint", C.c_uint64), ("nActive", C.c_uint64),
                ("V", C.c_void_p), ("N", C.c_void_p), ("T", C.c_void_p),
                ("vkey", C.c_void_p), ("tcell", C.c_void_p), ("tpat", C.c_void_p)]


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(str(ensure_oracle_built()))
        lib.mc33o_extract.restype = C.c_int
        lib.mc33o_extract.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double,
                                      C.POINTER(OGeom), C.c_int, C.POINTER(OMesh)]
        lib.mc33o_free.argtypes = [C.POINTER(OMesh)]
        lib.mc33o_cell_patterns.restype = C.c_int
        lib.mc33o_cell_patterns.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                            C.c_double, C.c_void_p]
        lib.mc33o_fill_xorshift.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_double]
        lib.mc33o_fill_xorshift.restype = None
        _oracle = lib
    return _oracle


def make_ogeom(geom, real):
    geom = geom or Geometry()
    store, O, D, ca, cb, A, Ai = geom.derived(real)
    g = OGeom()
    g.store, g.normal_neg, g.tsa = store, geom.normal_neg, geom.tsa
    for i in range(3):
        g.O[i], g.D[i] = O[i], D[i]
    g.ca, g.cb = ca, cb
    for i in range(9):
        g.A[i] = float(A.flat[i])
        g.Ai[i] = float(Ai.flat[i])
    return g


def oracle_extract(data, iso, variant="f32", geom=None, count_only=False):
    """data: (NZ,NY,NX) C-contiguous array of the variant's sample type."""
    code, sdt, real = DTYPES[variant]
    data = np.ascontiguousarray(data, dtype=sdt)
    NZ, NY, NX = data.shape
    lib = oracle_lib()
    g = make_ogeom(geom, real)
    m = OMesh()
    rc = lib.mc33o_extract(code, data.ctypes.data, NX - 1, NY - 1, NZ - 1, float(iso), C.byref(g),
                           int(count_only), C.byref(m))
    if rc != 0:
        raise RuntimeError("oracle failed")
    counts = dict(nShared=m.nShared, nCentre=m.nCentre, nPoint=m.nPoint, nActive=m.nActive)
    if count_only:
        out = Mesh(np.zeros((0, 3), real), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32), **counts)
        out.nV, out.nT = int(m.nV), int(m.nT)
        return out
    nV, nT = int(m.nV), int(m.nT)
    out = Mesh(_np_from(m.V, nV, real, (nV, 3)), _np_from(m.N, nV, np.float32, (nV, 3)),
               _np_from(m.T, nT, np.uint32, (nT, 3)),
               vkey=_np_from(m.vkey, nV, np.uint64, (nV,)), tcell=_np_from(m.tcell, nT, np.uint64, (nT,)),
               tpat=_np_from(m.tpat, nT, np.uint16, (nT,)), **counts)
    lib.mc33o_free(C.byref(m))
    return out


def oracle_cell_patterns(data, iso, variant="f32"):
    code, sdt, _ = DTYPES[variant]
    data = np.ascontiguousarray(data, dtype=sdt)
    NZ, NY, NX = data.shape
    pat = np.empty((NZ - 1, NY - 1, NX - 1), np.uint16)
    rc = oracle_lib().mc33o_cell_patterns(code, data.ctypes.data, NX - 1, NY - 1, NZ - 1, float(iso),
                                          pat.ctypes.data)
    assert rc == 0
    return pat


_refs = {}


def ref_lib(variant="f32", strict=False):
    key = (variant, strict)
    if key not in _refs:
        p = REF_DIR / f"libMC33_ref_{variant}{'_strict' if strict else ''}.so"
        _refs[key] = MC33Lib(p, variant)
    return _refs[key]


# --------------------------------------------------------------------------
# synthetic grids
# --------------------------------------------------------------------------
XS_SEED = 88172645463325252


def noise_grid(n, variant="f32", scale=1000.0, seed=XS_SEED, shape=None):
    """BASELINE.md K3..K6 generator (xorshift64, x fastest)."""
    code, sdt, _ = DTYPES[variant]
    shape = shape or (n, n, n)
    a = np.empty(shape, dtype=sdt)
    kind = {"f32": 0, "f64": 1, "u8": 2, "u16": 3}[variant]
    oracle_lib().mc33o_fill_xorshift(kind, a.ctypes.data, a.size, seed, float(scale))
    return a


def cfg1_grid():
    """generate_grid_from_fn(-4,-4,-4, 4,4,4, .04,.04,.04, cos x+cos y+cos z)
    restated (MC33_util_grd.c:630-686): N=(int)((xf-xi)/dx+.5), x += dx in double,
    libm cos, sample narrowed to float."""
    xi, xf, dx = -4.0, 4.0, 0.04
    n = int((xf - xi) / dx + 0.5)
    xs = []
    x = xi
    for _ in range(n + 1):
        xs.append(x)
        x += dx
    c = np.array([math.cos(v) for v in xs], dtype=np.float64)
    F = (c[None, None, :] + c[None, :, None]) + c[:, None, None]
    geom = Geometry(r0=(xi, xi, xi), d=(dx, dx, dx))
    return F.astype(np.float32), geom


def gyroid_grid(n, periods=4.0, dtype=np.float32):
    """cfg2 generator: w = 2*pi*periods, x = i/(n-1) - .5"""
    w = 2.0 * math.pi * periods
    t = (np.arange(n, dtype=np.float64) / (n - 1) - 0.5) * w
    s, c = np.sin(t), np.cos(t)
    X_s, X_c = s[None, None, :], c[None, None, :]
    Y_s, Y_c = s[None, :, None], c[None, :, None]
    Z_s, Z_c = s[:, None, None], c[:, None, None]
    return (X_s * Y_c + Y_s * Z_c + Z_s * X_c).astype(dtype)


def ct_grid(n, seed=7):
    """cfg3-like u16 volume: Gaussian blobs 1000..3500 + texture + 0..15 noise."""
    rng = np.random.default_rng(seed)
    ax = np.linspace(-1, 1, n)
    Z, Y, X = np.meshgrid(ax, ax, ax, indexing="ij")
    v = np.full((n, n, n), 1000.0)
    for _ in range(6):
        c = rng.uniform(-0.6, 0.6, 3)
        s = rng.uniform(0.15, 0.4)
        v += 2500.0 / 3 * np.exp(-((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2) / (2 * s * s))
    v += 20.0 * np.sin(37 * X) * np.sin(29 * Y) * np.sin(31 * Z)
    v += rng.integers(0, 16, size=v.shape)
    return np.clip(v, 0, 65535).astype(np.uint16)


INCLINED_A = np.array([[1.0, 0.3, 0.2], [0.0, 0.95, 0.1], [0.0, 0.0, 0.9]])


def inclined_geom(tsa=0, d=(1, 1, 1), r0=(0, 0, 0)):
    return Geometry(r0=r0, d=d, nonortho=1, A=INCLINED_A, Ai=np.linalg.inv(INCLINED_A), tsa=tsa)


# --------------------------------------------------------------------------
# comparisons
# --------------------------------------------------------------------------
def vertex_bijection(T_ref, T_new, nV, V_ref=None, V_new=None):
    """Triangles are listed in the same order by the reference and by the
    canonical order (cell-major, table order), so corner k of triangle j must
    denote the same vertex: that induces the map ref id -> new id.  Returns the
    map after asserting it is a well defined bijection.

    A vertex can legitimately be referenced by no triangle: an on-iso grid point
    whose triangles were all dropped as zero-area (marching_cubes_33.c:1235)
    still got its vertex stored.  Those are paired by position."""
    assert T_ref.shape == T_new.shape, (T_ref.shape, T_new.shape)
    a = T_ref.reshape(-1).astype(np.int64)
    b = T_new.reshape(-1).astype(np.int64)
    fwd = np.full(nV, -1, np.int64)
    fwd[a] = b
    assert np.array_equal(fwd[a], b), "one reference vertex maps to several new vertices"
    bwd = np.full(nV, -1, np.int64)
    bwd[b] = a
    assert np.array_equal(bwd[b], a), "several reference vertices map to one new vertex"
    ur, un = np.flatnonzero(fwd < 0), np.flatnonzero(bwd < 0)
    assert len(ur) == len(un), (len(ur), len(un))
    if len(ur):
        assert V_ref is not None, "unreferenced vertices and no positions to pair them"
        kr = np.lexsort(np.round(V_ref[ur].astype(np.float64), 4).T[::-1])
        kn = np.lexsort(np.round(V_new[un].astype(np.float64), 4).T[::-1])
        fwd[ur[kr]] = un[kn]
    return fwd


def unit(n):
    n = n.astype(np.float64)
    l = np.sqrt((n * n).sum(1, keepdims=True))
    l[l == 0] = 1
    return n / l


def compare_to_reference(ref, new, pos_rtol=1e-5, nrm_atol=1e-5, exact_pos=False):
    """Topology bit exact; positions within pos_rtol (relative to the coordinate
    magnitude, floor 1); normal directions within nrm_atol."""
    assert ref.nV == new.nV, (ref.nV, new.nV)
    assert ref.nT == new.nT, (ref.nT, new.nT)
    if ref.nT == 0:
        return None
    fwd = vertex_bijection(ref.T, new.T, ref.nV, ref.V, new.V)
    Vn = new.V[fwd]
    if exact_pos:
        assert np.array_equal(ref.V, Vn), "positions differ bitwise"
    err = np.abs(ref.V.astype(np.float64) - Vn.astype(np.float64))
    scale = np.maximum(1.0, np.abs(ref.V.astype(np.float64)))
    assert (err / scale).max() <= pos_rtol, (err / scale).max()
    # a zero gradient gives 0*inf = NaN normals in the reference (rsqrt(0)); the
    # same vertices must be NaN here, every other direction must agree
    bad_r = ~np.isfinite(ref.N).all(1)
    bad_n = ~np.isfinite(new.N[fwd]).all(1)
    assert np.array_equal(bad_r, bad_n), (bad_r.sum(), bad_n.sum())
    ok = ~bad_r
    if ok.any():
        nr, nn = unit(ref.N[ok]), unit(new.N[fwd][ok])
        nerr = np.abs(nr - nn).max()
        assert nerr <= nrm_atol, nerr
        lr = np.sqrt((ref.N[ok].astype(np.float64) ** 2).sum(1))
        ln = np.sqrt((new.N[fwd][ok].astype(np.float64) ** 2).sum(1))
        assert np.abs(lr - 1).max() <= 5e-4 and np.abs(ln - 1).max() <= 1e-6
    return fwd


def compare_exact(a, b, nrm_atol=2e-6):
    """Two meshes in this project's canonical order (oracle vs CUDA path)."""
    assert a.nV == b.nV and a.nT == b.nT, ((a.nV, a.nT), (b.nV, b.nT))
    assert np.array_equal(a.T, b.T), "triangle index arrays differ"
    assert np.array_equal(a.V, b.V), "positions differ bitwise"
    if a.nV:
        fa, fb = np.isfinite(a.N), np.isfinite(b.N)
        assert np.array_equal(fa, fb)
        d = np.abs(np.where(fa, a.N, 0).astype(np.float64) - np.where(fb, b.N, 0).astype(np.float64))
        assert d.max() <= nrm_atol, d.max()


# --------------------------------------------------------------------------
# C-ABI descriptions + the test-only host emulation of the device code
# --------------------------------------------------------------------------
HOSTEMU_SO = ROOT / "tests" / "hostemu" / "_build" / "libmc33_hostemu.so"
DEFAULT_COLOR = -10724260  # 0xff5c5c5c as int


def make_desc(shape, variant="f32", geom=None, slab=None):
    from mc33_c_library_b200 import _cabi as cabi
    code, _, real = DTYPES[variant]
    NZ, NY, NX = shape
    geom = geom or Geometry()
    store, O, D, ca, cb, A, Ai = geom.derived(real)
    kw = {}
    if slab is not None:
        kw = dict(z_lo=slab.z_lo, z_hi=slab.z_hi, cell_z0=slab.cell_z0, cell_z1=slab.cell_z1, is_last=slab.is_last)
    return cabi.make_desc(code, NX - 1, NY - 1, NZ - 1, store, O, D, ca, cb, A.flat, Ai.flat, geom.tsa,
                          geom.normal_neg, **kw)


_emu = None


def hostemu_lib():
    global _emu
    if _emu is None:
        from mc33_c_library_b200 import _cabi as cabi
        if not HOSTEMU_SO.exists():
            from mc33_c_library_b200 import build
            build.build_oracle()
        _emu = C.CDLL(str(HOSTEMU_SO))
        _emu.mc33emu_run.argtypes = [C.POINTER(cabi.Desc), C.c_void_p, C.c_double, C.POINTER(cabi.Out),
                                     C.POINTER(cabi.Counts)]
    return _emu


def emu_count(data_slab, iso, desc):
    from mc33_c_library_b200 import _cabi as cabi
    k = cabi.Counts()
    rc = hostemu_lib().mc33emu_run(C.byref(desc), data_slab.ctypes.data, float(iso), None, C.byref(k))
    assert rc == 0
    return k


def emu_emit(data_slab, iso, desc, k, real, vbase=0, vbase_next=0):
    from mc33_c_library_b200 import _cabi as cabi
    nV, nT = int(k.nV), int(k.nT)
    V = np.zeros((nV, 3), real); N = np.zeros((nV, 3), np.float32); col = np.zeros(nV, np.int32)
    T = np.zeros((nT, 3), np.uint32); vkey = np.zeros(nV, np.uint64); tcell = np.zeros(nT, np.uint64)
    o = cabi.Out()
    o.V, o.N, o.color, o.T = V.ctypes.data, N.ctypes.data, col.ctypes.data, T.ctypes.data
    o.vkey, o.tcell = vkey.ctypes.data, tcell.ctypes.data
    o.capV, o.capT, o.color_value = nV, nT, DEFAULT_COLOR
    o.vbase, o.vbase_next = vbase, vbase_next
    k2 = cabi.Counts()
    rc = hostemu_lib().mc33emu_run(C.byref(desc), data_slab.ctypes.data, float(iso), C.byref(o), C.byref(k2))
    assert rc == 0, rc
    return Mesh(V, N, T, color=col, vkey=vkey, tcell=tcell, nShared=int(k.nShared), nCentre=int(k.nCentre))


def emu_extract(data, iso, variant="f32", geom=None):
    code, sdt, real = DTYPES[variant]
    data = np.ascontiguousarray(data, dtype=sdt)
    d = make_desc(data.shape, variant, geom)
    k = emu_count(data, iso, d)
    return emu_emit(data, iso, d, k, real)


def merge_slab_meshes(meshes):
    """Concatenate per-slab meshes (already carrying global vertex ids) in rank
    order and bring them to the single-GPU canonical order: shared vertices of
    all slabs first, then all centre vertices."""
    nS = [m.counts["nShared"] for m in meshes]
    nC = [m.counts["nCentre"] for m in meshes]
    totS = sum(nS)
    remap_parts, off_s, off_c, vb = [], 0, totS, 0
    for m, s, c in zip(meshes, nS, nC):
        r = np.empty(s + c, np.int64)
        r[:s] = off_s + np.arange(s)
        r[s:] = off_c + np.arange(c)
        remap_parts.append(r)
        off_s += s
        off_c += c
    remap = np.concatenate(remap_parts) if remap_parts else np.zeros(0, np.int64)
    order = np.argsort(remap, kind="stable")
    V = np.concatenate([m.V for m in meshes])[order]
    N = np.concatenate([m.N for m in meshes])[order]
    vkey = np.concatenate([m.vkey for m in meshes])[order]
    T = remap[np.concatenate([m.T for m in meshes]).astype(np.int64)].astype(np.uint32)
    tcell = np.concatenate([m.tcell for m in meshes])
    return Mesh(V, N, T, vkey=vkey, tcell=tcell, nShared=totS, nCentre=sum(nC))
