"""The drop-in boundary seen from C: tests/c/dropin_link.c is compiled against
include/marching_cubes_33.h with each variant's -D flags and linked with
-lMC33_b200_<variant>, the way a user of the reference links -lMC33
(reference README.md:84-88).  Compilation checks the struct layouts of SURVEY.md 8(a)
(_Static_assert); the run goes through the README flow and the file functions.

CPU suite: build + run; without a device create_MC33 returns NULL (NO_DEVICE) and the
host-only entry points are still exercised.  GPU suite: the counts the program prints
are compared with the reference's known answers / the oracle."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
LIBDIR = ROOT / "mc33_c_library_b200" / "lib"
SRC = ROOT / "tests" / "c" / "dropin_link.c"

VARIANTS = {
    "f32": [],
    "f64": ["-DGRD_TYPE_SIZE=8"],
    "u8": ["-DINTEGER_GRD", "-DGRD_INTEGER", "-DGRD_TYPE_SIZE=1"],
    "u16": ["-DINTEGER_GRD", "-DGRD_INTEGER", "-DGRD_TYPE_SIZE=2"],
    "u32": ["-DINTEGER_GRD", "-DGRD_INTEGER", "-DGRD_TYPE_SIZE=4"],
    "f32_ortho": ["-DGRD_ORTHOGONAL"],
}


def build_and_run(variant, tmp_path):
    exe = tmp_path / f"dropin_link_{variant}"
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", *VARIANTS[variant], str(SRC),
           "-o", str(exe), f"-L{LIBDIR}", f"-lMC33_b200_{variant}", f"-Wl,-rpath,{LIBDIR}", "-lm"]
    subprocess.check_call(cmd)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=600)
    return r


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_c_program_links_and_runs_host_side(variant, tmp_path):
    r = build_and_run(variant, tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "DONE fails=0" in r.stdout
    assert "RAW ok" in r.stdout and "DAT ok" in r.stdout
    assert "NO_DEVICE" in r.stdout or "SURFACE" in r.stdout


def _int_sphere(n, dtype):
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    r = np.sqrt((i - 11.5) ** 2 + (j - 11.5) ** 2 + (k - 11.5) ** 2)
    return np.clip(100.0 - 8.0 * r + 0.5, 0, None).astype(dtype)     # same expression as the C program (C casts truncate, as astype does)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", list(VARIANTS))
def test_c_program_readme_flow_on_gpu(variant, tmp_path):
    """K1 (150 096 vertices / 297 872 triangles for the README's cfg1 grid) through a
    compiled C caller; integer variants against the oracle on the program's own grid"""
    import json
    from support import oracle_extract
    r = build_and_run(variant, tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "NO_DEVICE" not in r.stdout and "DONE fails=0" in r.stdout and "BIN_ROUNDTRIP ok" in r.stdout and "EMPTY ok" in r.stdout
    m = re.search(r"SURFACE (\d+) (\d+) capv=(\d+) capt=(\d+)", r.stdout)
    nV, nT = int(m.group(1)), int(m.group(2))
    m2 = re.search(r"SURFACE2 (\d+) (\d+)", r.stdout)
    if variant in ("f32", "f64", "f32_ortho"):
        kats = json.loads((ROOT / "tests" / "golden" / "kats.json").read_text())
        if variant != "f64":      # the KATs were measured on the float build; the double build samples cos() in double
            assert (nV, nT) == (kats["K1"]["nV"], kats["K1"]["nT"])
            assert (int(m2.group(1)), int(m2.group(2))) == (kats["K2"]["nV"], kats["K2"]["nT"])
        else:
            assert abs(nV - kats["K1"]["nV"]) < 2000
    else:
        dt = {"u8": np.uint8, "u16": np.uint16, "u32": np.uint32}[variant]
        want = oracle_extract(_int_sphere(24, dt), 40.0, variant, count_only=True)
        assert (nV, nT) == (want.nV, want.nT)
    # the files the program wrote are well formed
    obj = (tmp_path / "s.obj").read_text().splitlines()
    assert sum(l.startswith("v ") for l in obj) == nV and sum(l.startswith("f ") for l in obj) == nT
    ply = (tmp_path / "s.ply").read_text().splitlines()
    assert ply[0] == "ply" and f"element vertex {nV}" in ply and f"element face {nT}" in ply[:16]
