"""ctypes view of the C-ABI in include/mc33cu.h (libmc33cu.so).

Plumbing only: structures, prototypes, library loading.  The library is built
in-tree by build.py (nvcc, sm_100a); loading fails loudly if it is missing --
there is no CPU fallback.
"""
import ctypes as C
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libmc33cu.so"

F32, F64, U8, U16, U32 = range(5)
SPN0, SPNA, SPNB, SPNC = range(4)
OK, ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_CAPACITY, ERR_RANGE, ERR_STATE = 0, -1, -2, -3, -4, -5, -6


class Desc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("nx", C.c_uint32), ("ny", C.c_uint32), ("nz", C.c_uint32),
                ("z_lo", C.c_uint32), ("z_hi", C.c_uint32), ("cell_z0", C.c_uint32), ("cell_z1", C.c_uint32),
                ("is_last", C.c_int32), ("store", C.c_int32), ("normal_neg", C.c_int32), ("tsa", C.c_int32),
                ("O", C.c_double * 3), ("D", C.c_double * 3), ("ca", C.c_double), ("cb", C.c_double),
                ("A", C.c_double * 9), ("Ai", C.c_double * 9)]


class Counts(C.Structure):
    _fields_ = [("nV", C.c_uint64), ("nT", C.c_uint64), ("nShared", C.c_uint64), ("nCentre", C.c_uint64),
                ("nSharedHalo", C.c_uint64)]


class Out(C.Structure):
    _fields_ = [("V", C.c_void_p), ("N", C.c_void_p), ("color", C.c_void_p), ("T", C.c_void_p),
                ("vkey", C.c_void_p), ("tcell", C.c_void_p), ("capV", C.c_uint32), ("capT", C.c_uint32),
                ("vbase", C.c_uint32), ("vbase_next", C.c_uint32), ("color_value", C.c_int32),
                ("pad_", C.c_int32), ("dev_bases", C.c_void_p)]


EXPORTS = ["mc33cu_last_error", "mc33cu_device_count", "mc33cu_create", "mc33cu_destroy", "mc33cu_set_stream", "mc33cu_set_geometry",
           "mc33cu_grid_device", "mc33cu_grid_upload", "mc33cu_grid_upload_rows", "mc33cu_count", "mc33cu_count_async",
           "mc33cu_slab_bases", "mc33cu_slab_bases_strided", "mc33cu_emit_set_device", "mc33cu_emit_device", "mc33cu_extract_device", "mc33cu_classify_sweep", "mc33cu_count_set_async", "mc33cu_extract_set_device", "mc33cu_sync", "mc33cu_get_counts",
           "mc33cu_emit_host", "mc33cu_emit_host_async", "mc33cu_grid_upload_async", "mc33cu_grid_upload_rows_async", "mc33cu_grid_rows_block", "mc33cu_host_register", "mc33cu_host_unregister", "mc33cu_host_alloc", "mc33cu_host_free", "mc33cu_enable_timing", "mc33cu_kernel_times", "mc33cu_launch_count", "mc33cu_stream_wait"]

_lib = None


class Mc33CudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mc33cu error {code}: {msg}")
        self.code = code


def load():
    """Load libmc33cu.so (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    import os
    path = Path(os.environ.get("MC33_B200_LIB", LIB_PATH))    # A/B builds of the same C-ABI (tools/)
    if not path.exists():
        raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(str(path))
    vp, i32, u64 = C.c_void_p, C.c_int32, C.c_uint64
    lib.mc33cu_last_error.restype = C.c_char_p
    lib.mc33cu_device_count.restype = C.c_int
    lib.mc33cu_create.argtypes = [C.POINTER(Desc), C.c_int, C.POINTER(vp)]
    lib.mc33cu_destroy.argtypes = [vp]
    lib.mc33cu_destroy.restype = None
    lib.mc33cu_set_stream.argtypes = [vp, vp]
    lib.mc33cu_set_geometry.argtypes = [vp, C.POINTER(Desc)]
    lib.mc33cu_grid_device.argtypes = [vp, vp]
    lib.mc33cu_grid_upload.argtypes = [vp, vp]
    lib.mc33cu_grid_upload_rows.argtypes = [vp, vp]
    lib.mc33cu_count.argtypes = [vp, C.c_double, C.POINTER(Counts)]
    lib.mc33cu_count_async.argtypes = [vp, C.c_double, vp]
    lib.mc33cu_slab_bases.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    lib.mc33cu_slab_bases_strided.argtypes = [vp, vp, C.c_uint32, C.c_int, C.c_int, vp]
    lib.mc33cu_emit_set_device.argtypes = [vp, C.c_int, C.POINTER(Out)]
    lib.mc33cu_emit_device.argtypes = [vp, C.POINTER(Out)]
    lib.mc33cu_extract_device.argtypes = [vp, C.c_double, C.POINTER(Out)]
    lib.mc33cu_classify_sweep.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
    lib.mc33cu_count_set_async.argtypes = [vp, C.c_int, vp]
    lib.mc33cu_extract_set_device.argtypes = [vp, C.c_int, C.POINTER(Out)]
    lib.mc33cu_sync.argtypes = [vp]
    lib.mc33cu_get_counts.argtypes = [vp, C.POINTER(Counts)]
    lib.mc33cu_emit_host.argtypes = [vp, vp, vp, vp, vp, i32]
    lib.mc33cu_emit_host_async.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, i32]
    lib.mc33cu_grid_upload_async.argtypes = [vp, vp]
    lib.mc33cu_grid_upload_rows_async.argtypes = [vp, vp]
    lib.mc33cu_grid_rows_block.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.mc33cu_host_register.argtypes = [vp, C.c_size_t]
    lib.mc33cu_host_unregister.argtypes = [vp]
    lib.mc33cu_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    lib.mc33cu_host_free.argtypes = [vp]
    lib.mc33cu_enable_timing.argtypes = [vp, C.c_int]
    lib.mc33cu_kernel_times.argtypes = [vp, C.POINTER(C.c_float * 5)]
    lib.mc33cu_launch_count.argtypes = [vp]
    lib.mc33cu_launch_count.restype = u64
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise Mc33CudaError(rc, load().mc33cu_last_error().decode())


def make_desc(dtype, nx, ny, nz, store=SPN0, O=(0, 0, 0), D=(1, 1, 1), ca=1.0, cb=1.0, A=None, Ai=None,
              tsa=0, normal_neg=0, z_lo=0, z_hi=None, cell_z0=0, cell_z1=None, is_last=1):
    d = Desc()
    d.dtype, d.nx, d.ny, d.nz = dtype, nx, ny, nz
    d.z_lo, d.z_hi = z_lo, (nz + 1 if z_hi is None else z_hi)
    d.cell_z0, d.cell_z1 = cell_z0, (nz if cell_z1 is None else cell_z1)
    d.is_last = int(is_last)
    d.store, d.normal_neg, d.tsa = store, int(normal_neg), int(tsa)
    ident = (1, 0, 0, 0, 1, 0, 0, 0, 1)
    A = ident if A is None else [float(v) for v in A]
    Ai = ident if Ai is None else [float(v) for v in Ai]
    for i in range(3):
        d.O[i], d.D[i] = float(O[i]), float(D[i])
    d.ca, d.cb = float(ca), float(cb)
    for i in range(9):
        d.A[i], d.Ai[i] = A[i], Ai[i]
    return d
