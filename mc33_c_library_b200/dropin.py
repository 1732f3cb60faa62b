"""Host-side Python mirror of the reference's C interface (include/marching_cubes_33.h):
a ctypes binding of one build of the API -- _GRD / surface layouts, grid_from_data_pointer,
create_MC33, calculate_isosurface, size_of_isosurface, free_* (reference
include/marching_cubes_33.h:111-179, :189-329).  The same binding drives this project's
drop-in libraries (mc33_c_library_b200/lib/libMC33_b200_<variant>.so) and, in tests and
in bench.py's CPU arm, the compiled reference: they export identical symbols and layouts.

Plumbing only: no marching cubes arithmetic lives here.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_DIR = PKG_DIR / "lib"

# variant -> (C-ABI dtype code, sample dtype, MC33_real dtype)
DTYPES = {
    "f32": (0, np.float32, np.float32),
    "f64": (1, np.float64, np.float64),
    "u8": (2, np.uint8, np.float32),
    "u16": (3, np.uint16, np.float32),
    "u32": (4, np.uint32, np.float32),
}

SPN0, SPNA, SPNB, SPNC = 0, 1, 2, 3


# --------------------------------------------------------------------------
# geometry, exactly as create_MC33 derives it (marching_cubes_33.c:1762-1782)
# --------------------------------------------------------------------------
class Geometry:
    def __init__(self, r0=(0, 0, 0), d=(1, 1, 1), nonortho=0, A=None, Ai=None, tsa=0, normal_neg=0):
        self.r0 = [float(v) for v in r0]
        self.d = [float(v) for v in d]
        self.nonortho = int(nonortho)
        self.A = np.eye(3) if A is None else np.asarray(A, dtype=np.float64)
        self.Ai = np.eye(3) if Ai is None else np.asarray(Ai, dtype=np.float64)
        self.tsa = int(tsa)
        self.normal_neg = int(normal_neg)

    def derived(self, real):
        """-> (store, O, D, ca, cb, A', Ai') with O, D, ca, cb narrowed to `real`."""
        d, r0 = self.d, self.r0
        ca = cb = 1.0
        A = np.eye(3)
        Ai = np.eye(3)
        if self.nonortho:
            store = SPNC
            A = np.array([[self.A[j][i] * d[i] for i in range(3)] for j in range(3)])
            Ai = np.array([[self.Ai[j][i] / d[j] for i in range(3)] for j in range(3)])
        elif d[0] != d[1] or d[1] != d[2]:
            store = SPNB
            ca = float(real(d[2] / d[0]))
            cb = float(real(d[2] / d[1]))
        else:
            store = SPN0 if (d[0] == 1 and r0[0] == 0 and r0[1] == 0 and r0[2] == 0) else SPNA
        O = [float(real(v)) for v in r0]
        D = [float(real(v)) for v in d]
        return store, O, D, ca, cb, A, Ai


class Mesh:
    """Plain container: V (nV,3) real, N (nV,3) f32, T (nT,3) u32 (+ optional keys)."""

    def __init__(self, V, N, T, color=None, vkey=None, tcell=None, tpat=None, **counts):
        self.V, self.N, self.T, self.color = V, N, T, color
        self.vkey, self.tcell, self.tpat = vkey, tcell, tpat
        self.nV, self.nT = len(V), len(T)
        self.counts = counts


def _np_from(ptr, n, dtype, shape):
    if n == 0 or not ptr:
        return np.zeros(shape, dtype=dtype)
    buf = (C.c_char * (int(np.prod(shape)) * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()


# --------------------------------------------------------------------------
# the compiled, unmodified reference (oracle/_ref/libMC33_ref_<variant>.so)
# struct layouts: include/marching_cubes_33.h:111-152 (default, inclined-capable build)
# --------------------------------------------------------------------------
class RefGRD(C.Structure):
    _fields_ = [("F", C.c_void_p), ("N", C.c_uint * 3), ("r0", C.c_double * 3), ("d", C.c_double * 3),
                ("L", C.c_float * 3), ("Ang", C.c_float * 3), ("nonortho", C.c_int),
                ("_A", (C.c_double * 3) * 3), ("A_", (C.c_double * 3) * 3),
                ("periodic", C.c_int), ("internal_data", C.c_int), ("title", C.c_char * 160)]


def _surface_struct(real_ctype):
    class Surface(C.Structure):
        _fields_ = [("T", C.c_void_p), ("V", C.c_void_p), ("N", C.c_void_p), ("color", C.c_void_p),
                    ("nV", C.c_uint), ("nT", C.c_uint), ("capt", C.c_uint), ("capv", C.c_uint),
                    ("iso", real_ctype), ("user", C.c_longlong)]
    return Surface


class MC33Lib:
    """Binding of one build of the marching_cubes_33.h API -- used both for the
    reference (oracle/_ref) and for this project's drop-in library, which export
    identical symbols and struct layouts."""

    def __init__(self, path, variant):
        self.variant = variant
        self.code, self.sdt, self.real = DTYPES[variant]
        self.lib = C.CDLL(str(path), mode=os.RTLD_NOW | os.RTLD_LOCAL)
        self.real_c = C.c_double if variant == "f64" else C.c_float
        self.Surface = _surface_struct(self.real_c)
        L = self.lib
        L.grid_from_data_pointer.restype = C.POINTER(RefGRD)
        L.grid_from_data_pointer.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_void_p]
        L.create_MC33.restype = C.c_void_p
        L.create_MC33.argtypes = [C.POINTER(RefGRD)]
        L.calculate_isosurface.restype = C.POINTER(self.Surface)
        L.calculate_isosurface.argtypes = [C.c_void_p, self.real_c]
        L.size_of_isosurface.restype = C.c_ulonglong
        L.size_of_isosurface.argtypes = [C.c_void_p, self.real_c, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        L.free_surface_memory.argtypes = [C.POINTER(self.Surface)]
        L.free_MC33.argtypes = [C.c_void_p]
        L.free_memory_grd.argtypes = [C.POINTER(RefGRD)]
        L.generate_grid_from_fn.restype = C.POINTER(RefGRD)
        L.alloc_F.argtypes = [C.POINTER(RefGRD)]
        L.alloc_F.restype = C.c_int

    def make_grid(self, data, geom=None):
        data = np.ascontiguousarray(data, dtype=self.sdt)
        NZ, NY, NX = data.shape
        G = self.lib.grid_from_data_pointer(NX, NY, NZ, data.ctypes.data)
        assert G
        if geom is not None:
            g = G.contents
            for i in range(3):
                g.r0[i] = geom.r0[i]
                g.d[i] = geom.d[i]
            g.nonortho = geom.nonortho
            for j in range(3):
                for i in range(3):
                    g._A[j][i] = float(geom.A[j][i])
                    g.A_[j][i] = float(geom.Ai[j][i])
        return G, data

    def set_tsa(self, on):
        """Point mult_Abf at _multTSA_bf / _multA_bf (MC33_util_grd.c:86-114)."""
        fp = C.c_void_p.in_dll(self.lib, "mult_Abf")
        fn = self.lib._multTSA_bf if on else self.lib._multA_bf
        fp.value = C.cast(fn, C.c_void_p).value

    def extract(self, data, iso, geom=None, keep=False):
        G, data = self.make_grid(data, geom)
        if geom is not None:
            self.set_tsa(geom.tsa)
        M = self.lib.create_MC33(G)
        assert M
        S = self.lib.calculate_isosurface(M, self.real_c(iso))
        assert S, "calculate_isosurface returned NULL"
        s = S.contents
        nV, nT = int(s.nV), int(s.nT)
        mesh = Mesh(_np_from(s.V, nV, self.real, (nV, 3)), _np_from(s.N, nV, np.float32, (nV, 3)),
                    _np_from(s.T, nT, np.uint32, (nT, 3)), color=_np_from(s.color, nV, np.int32, (nV,)))
        mesh.iso = float(s.iso)
        self.lib.free_surface_memory(S)
        self.lib.free_MC33(M)
        self.lib.free_memory_grd(G)
        if geom is not None and geom.tsa:
            self.set_tsa(0)
        return mesh

    def extract_many(self, data, isos, geom=None):
        """several calculate_isosurface calls on ONE MC33 (an iso sweep through the reference API) -> list of Mesh"""
        G, data = self.make_grid(data, geom)
        M = self.lib.create_MC33(G)
        assert M
        out = []
        for iso in isos:
            S = self.lib.calculate_isosurface(M, self.real_c(iso))
            assert S, "calculate_isosurface returned NULL"
            s = S.contents
            nV, nT = int(s.nV), int(s.nT)
            assert int(s.capv) >= nV and int(s.capt) >= nT
            mesh = Mesh(_np_from(s.V, nV, self.real, (nV, 3)), _np_from(s.N, nV, np.float32, (nV, 3)),
                        _np_from(s.T, nT, np.uint32, (nT, 3)), color=_np_from(s.color, nV, np.int32, (nV,)))
            mesh.iso = float(s.iso)
            mesh.capv, mesh.capt = int(s.capv), int(s.capt)
            out.append(mesh)
            self.lib.free_surface_memory(S)
        self.lib.free_MC33(M)
        self.lib.free_memory_grd(G)
        return out

    def size(self, data, iso, geom=None):
        G, data = self.make_grid(data, geom)
        M = self.lib.create_MC33(G)
        nV, nT = C.c_uint(0), C.c_uint(0)
        sz = self.lib.size_of_isosurface(M, self.real_c(iso), C.byref(nV), C.byref(nT))
        self.lib.free_MC33(M)
        self.lib.free_memory_grd(G)
        return int(sz), int(nV.value), int(nT.value)




def dropin(variant="f32"):
    """this project's drop-in library for one element type (raises if it has not been built)"""
    p = LIB_DIR / f"libMC33_b200_{variant}.so"
    if not p.exists():
        raise ImportError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    return MC33Lib(p, variant)
