"""In-tree build of the native code: nvcc for sm_100a, gcc for the plain-C drop-in.

    python -m mc33_c_library_b200.build [--force]

Outputs (git-ignored, shipped to the GPU box by gpurun) under mc33_c_library_b200/lib/:
    libmc33cu.so                 CUDA kernels + C-ABI (include/mc33cu.h)
    libMC33_b200_<variant>.so    marching_cubes_33.h API, one per element type
"""
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "lib"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false",            # no FMA contraction: case selection must match the reference bit for bit
              "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
              "-Wno-deprecated-gpu-targets"]

VARIANTS = {
    "f32": [],
    "f64": ["-DGRD_TYPE_SIZE=8"],
    "u8": ["-DINTEGER_GRD", "-DGRD_INTEGER", "-DGRD_TYPE_SIZE=1"],
    "u16": ["-DINTEGER_GRD", "-DGRD_INTEGER", "-DGRD_TYPE_SIZE=2"],
    "u32": ["-DINTEGER_GRD", "-DGRD_INTEGER", "-DGRD_TYPE_SIZE=4"],
    "f32_ortho": ["-DGRD_ORTHOGONAL"],
}


def _newer(target, sources):
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def _run(cmd):
    print("+", " ".join(str(c) for c in cmd), flush=True)
    subprocess.check_call([str(c) for c in cmd])


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(p).exists():
        raise RuntimeError("nvcc not found")
    return p


def build_cuda(force=False, extra=()):
    LIB.mkdir(exist_ok=True)
    out = LIB / "libmc33cu.so"
    srcs = [CSRC / "mc33_kernels.cu", CSRC / "mc33_core.cuh", CSRC / "mc33_pipeline.cuh", CSRC / "mc33_simt.h", CSRC / "mc33_tables.h",
            ROOT / "include" / "mc33cu.h"]
    if force or not _newer(out, srcs):
        _run([nvcc_path(), *NVCC_FLAGS, *extra, CSRC / "mc33_kernels.cu", "-o", out])
    return out


def build_dropin(force=False):
    outs = []
    srcs = [CSRC / "mc33_api.c", CSRC / "mc33_io.c", CSRC / "mc33_internal.h", ROOT / "include" / "marching_cubes_33.h",
            ROOT / "include" / "mc33cu.h"]
    srcs = [s for s in srcs if s.exists()]
    csrcs = [s for s in srcs if s.suffix == ".c"]
    for name, defs in VARIANTS.items():
        out = LIB / f"libMC33_b200_{name}.so"
        if force or not _newer(out, srcs + [LIB / "libmc33cu.so"]):
            _run(["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra", "-ffp-contract=off", *defs, *csrcs,
                  "-o", out, f"-L{LIB}", "-lmc33cu", "-Wl,-rpath,$ORIGIN", "-lm"])
        outs.append(out)
    return outs


def build_oracle():
    """Test infrastructure: the CPU restatement, and the unmodified reference when
    its sources are present (only in the build container)."""
    _run(["make", "-s", "-C", ROOT / "oracle", "all"])
    emu = ROOT / "tests" / "hostemu"
    out = emu / "_build" / "libmc33_hostemu.so"
    srcs = [emu / "mc33_hostemu.cu", emu / "simt_emu.h", CSRC / "mc33_core.cuh", CSRC / "mc33_pipeline.cuh", CSRC / "mc33_simt.h",
            CSRC / "mc33_tables.h"]
    if not _newer(out, srcs):
        # plain g++ (host only): the kernel bodies of mc33_pipeline.cuh run as fibers under tests/hostemu/simt_emu.h
        out.parent.mkdir(exist_ok=True)
        _run(["g++", "-O1", "-g", "-std=c++17", "-x", "c++", "-fPIC", "-ffp-contract=off", "-shared", emu / "mc33_hostemu.cu", "-o", out])


def build_all(force=False):
    build_cuda(force)
    build_dropin(force)
    build_oracle()


if __name__ == "__main__":
    build_all("--force" in sys.argv)
