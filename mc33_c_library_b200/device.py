"""Device-side driver used by bench.py and the GPU tests: torch owns the device
buffers and streams, the C-ABI (include/mc33cu.h) does all the work."""
import ctypes as C

import numpy as np
import torch

from . import _cabi as cabi

TORCH_SAMPLE = {cabi.F32: torch.float32, cabi.F64: torch.float64, cabi.U8: torch.uint8,
                cabi.U16: torch.uint16, cabi.U32: torch.uint32}
NP_SAMPLE = {cabi.F32: np.float32, cabi.F64: np.float64, cabi.U8: np.uint8, cabi.U16: np.uint16, cabi.U32: np.uint32}


class Extractor:
    """One mc33cu context = one grid (or z-slab of it) on one GPU."""

    def __init__(self, desc, device=0):
        self.lib = cabi.load()
        self.desc = desc
        self.device = int(device)
        self.real = torch.float64 if desc.dtype == cabi.F64 else torch.float32
        h = C.c_void_p()
        cabi.check(self.lib.mc33cu_create(C.byref(desc), self.device, C.byref(h)))
        self.h = h
        self._grid = None
        self._bufs = None

    def close(self):
        if self.h:
            self.lib.mc33cu_destroy(self.h)
            self.h = None

    __del__ = close

    # -- grid ---------------------------------------------------------------
    def bind(self, t):
        """t: contiguous CUDA tensor holding the slab's sample slices (borrowed)."""
        assert t.is_cuda and t.is_contiguous()
        self._grid = t
        cabi.check(self.lib.mc33cu_grid_device(self.h, C.c_void_p(t.data_ptr())))

    def upload(self, a):
        """a: contiguous host numpy array / pinned tensor (copied)."""
        ptr = a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr()
        cabi.check(self.lib.mc33cu_grid_upload(self.h, C.c_void_p(ptr)))

    def use_stream(self, stream):
        cabi.check(self.lib.mc33cu_set_stream(self.h, C.c_void_p(stream.cuda_stream if stream is not None else 0)))

    # -- extraction -----------------------------------------------------------
    def count(self, iso):
        k = cabi.Counts()
        cabi.check(self.lib.mc33cu_count(self.h, float(iso), C.byref(k)))
        return k

    def count_async(self, iso, dev_counts4):
        """classify + count + scan; {nV,nT,nShared,nCentre} land in the CUDA int32
        tensor dev_counts4 (no host synchronisation)."""
        cabi.check(self.lib.mc33cu_count_async(self.h, float(iso), C.c_void_p(dev_counts4.data_ptr())))

    def slab_bases(self, counts_all, rank, world, bases2):
        """{vbase, vbase_next} of this slab from the all-gathered counts (CUDA int32 tensors
        [world, 4] and [2]); one small kernel on the context stream, no host round trip."""
        cabi.check(self.lib.mc33cu_slab_bases(self.h, C.c_void_p(counts_all.data_ptr()), int(rank), int(world),
                                              C.c_void_p(bases2.data_ptr())))

    def alloc(self, capV, capT, keys=False):
        dev = torch.device("cuda", self.device)
        b = dict(V=torch.empty((max(capV, 1), 3), dtype=self.real, device=dev),
                 N=torch.empty((max(capV, 1), 3), dtype=torch.float32, device=dev),
                 color=torch.empty((max(capV, 1),), dtype=torch.int32, device=dev),
                 T=torch.empty((max(capT, 1), 3), dtype=torch.int32, device=dev), capV=capV, capT=capT)
        if keys:
            b["vkey"] = torch.empty((max(capV, 1),), dtype=torch.int64, device=dev)
            b["tcell"] = torch.empty((max(capT, 1),), dtype=torch.int64, device=dev)
        return b

    def _out(self, b, vbase=0, vbase_next=0, color=-10724260, dev_bases=None):
        o = cabi.Out()
        o.V, o.N, o.color, o.T = b["V"].data_ptr(), b["N"].data_ptr(), b["color"].data_ptr(), b["T"].data_ptr()
        o.vkey = b["vkey"].data_ptr() if "vkey" in b else None
        o.tcell = b["tcell"].data_ptr() if "tcell" in b else None
        o.capV, o.capT = b["capV"], b["capT"]
        o.vbase, o.vbase_next, o.color_value = vbase, vbase_next, color
        o.dev_bases = dev_bases.data_ptr() if dev_bases is not None else None
        return o

    def emit(self, b, **kw):
        o = self._out(b, **kw)
        cabi.check(self.lib.mc33cu_emit_device(self.h, C.byref(o)))

    def extract_async(self, iso, b, **kw):
        """classify + count + scan + emit with no host synchronisation."""
        o = self._out(b, **kw)
        cabi.check(self.lib.mc33cu_extract_device(self.h, float(iso), C.byref(o)))

    # -- iso sweep: classify once for up to 8 isovalues, then count / emit per set ----------
    def classify_sweep(self, isos):
        arr = (C.c_double * len(isos))(*[float(v) for v in isos])
        cabi.check(self.lib.mc33cu_classify_sweep(self.h, arr, len(isos)))

    def count_set_async(self, j, dev_counts4):
        cabi.check(self.lib.mc33cu_count_set_async(self.h, int(j), C.c_void_p(dev_counts4.data_ptr())))

    def extract_set_async(self, j, b, **kw):
        o = self._out(b, **kw)
        cabi.check(self.lib.mc33cu_extract_set_device(self.h, int(j), C.byref(o)))

    def emit_set(self, j, b, **kw):
        """emit pre-classified, already counted set j (all sets can be counted before any is emitted)"""
        o = self._out(b, **kw)
        cabi.check(self.lib.mc33cu_emit_set_device(self.h, int(j), C.byref(o)))

    def slab_bases_strided(self, counts_all, stride_words, rank, world, bases2):
        cabi.check(self.lib.mc33cu_slab_bases_strided(self.h, C.c_void_p(counts_all.data_ptr()), int(stride_words), int(rank),
                                                      int(world), C.c_void_p(bases2.data_ptr())))

    def sync(self):
        cabi.check(self.lib.mc33cu_sync(self.h))
        k = cabi.Counts()
        cabi.check(self.lib.mc33cu_get_counts(self.h, C.byref(k)))
        return k

    def extract(self, iso, keys=False):
        """count, size the buffers exactly, emit; returns host numpy arrays."""
        k = self.count(iso)
        nV, nT = int(k.nV), int(k.nT)
        b = self.alloc(nV, nT, keys)
        self.emit(b)
        self.sync()
        out = {n: b[n][: (nV if n != "T" and n != "tcell" else nT)].cpu().numpy() for n in b if n not in ("capV", "capT")}
        out["T"] = out["T"].view(np.uint32)
        out["counts"] = k
        return out

    def timing(self, on):
        cabi.check(self.lib.mc33cu_enable_timing(self.h, int(on)))

    def kernel_times(self):
        ms = (C.c_float * 5)()
        cabi.check(self.lib.mc33cu_kernel_times(self.h, C.byref(ms)))
        return list(ms)

    def launches(self):
        return int(self.lib.mc33cu_launch_count(self.h))
