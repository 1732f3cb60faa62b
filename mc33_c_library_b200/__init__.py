"""mc33_c_library_b200 -- B200-native Marching Cubes 33 (drop-in for MC33_c_library's
calculate_isosurface path).

The product is native code: mc33_c_library_b200/csrc (CUDA kernels for sm_100a +
the C-ABI of include/mc33cu.h + the plain-C marching_cubes_33.h API).  This
Python package is only the plumbing used by bench.py and the tests: ctypes
bindings, device buffers via torch, torch.distributed for the z-slab counts.
"""
from . import _cabi  # noqa: F401

__all__ = ["_cabi"]
