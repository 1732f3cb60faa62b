"""Pins the CPU restatement (oracle/mc33_oracle.c) against the compiled,
unmodified reference (oracle/_ref) and against the known-answer counts of
BASELINE.md section 2 (K1..K7).  CPU only."""
import json
from pathlib import Path

import numpy as np
import pytest

from support import (Geometry, cfg1_grid, compare_to_reference, ct_grid, gyroid_grid, have_ref,
                     inclined_geom, noise_grid, oracle_extract, ref_lib)

KATS = json.loads((Path(__file__).parent / "golden" / "kats.json").read_text())
needs_ref = pytest.mark.skipif(not have_ref("f32"), reason="oracle/_ref not built")


def _check(data, iso, variant, geom=None, strict=False, **kw):
    ref = ref_lib(variant, strict).extract(data, iso, geom)
    orc = oracle_extract(data, iso, variant, geom)
    compare_to_reference(ref, orc, **kw)
    return ref, orc


def test_kat_counts_oracle_only():
    """K1..K6 counts from BASELINE.md, no reference binary needed."""
    F, geom = cfg1_grid()
    for iso, key in ((0.0, "K1"), (0.5, "K2")):
        m = oracle_extract(F, iso, "f32", geom, count_only=True)
        assert (m.nV, m.nT) == (KATS[key]["nV"], KATS[key]["nT"])
    a = noise_grid(128, "f32")
    m = oracle_extract(a, 0.0, "f32", count_only=True)
    assert (m.nV, m.nT) == (KATS["K3"]["nV"], KATS["K3"]["nT"])
    u = noise_grid(128, "u16", scale=1000)
    for iso, k in ((500.0, "K5a"), (500.5, "K5b")):
        m = oracle_extract(u, iso, "u16", count_only=True)
        assert (m.nV, m.nT) == (KATS[k]["nV"], KATS[k]["nT"])
    b = noise_grid(128, "u8", scale=6)
    m = oracle_extract(b, 3.0, "u8", count_only=True)
    assert (m.nV, m.nT) == (KATS["K6"]["nV"], KATS["K6"]["nT"])
    assert m.counts["nPoint"] == KATS["K6"]["nPoint"]
    assert m.counts["nCentre"] == KATS["K6"]["nCentre"]


@needs_ref
def test_cfg1_exact_positions():
    F, geom = cfg1_grid()
    ref, orc = _check(F, 0.0, "f32", geom, exact_pos=True)
    assert (ref.nV, ref.nT) == (KATS["K1"]["nV"], KATS["K1"]["nT"])
    assert (ref.color == np.int32(-10724260)).all()  # 0xff5c5c5c


@needs_ref
@pytest.mark.parametrize("variant,iso,scale", [("f32", 0.0, 0), ("f32", 0.05, 0), ("f64", 0.0, 0),
                                               ("u16", 500.0, 1000), ("u16", 500.5, 1000),
                                               ("u8", 3.0, 6), ("u8", 2.5, 6), ("u16", 1.0, 3)])
def test_noise_all_types(variant, iso, scale):
    a = noise_grid(64, variant, scale=scale)
    _check(a, iso, variant, exact_pos=True)


@needs_ref
def test_strict_and_fast_reference_agree_with_oracle():
    a = noise_grid(48, "f32")
    for strict in (False, True):
        _check(a, 0.0, "f32", strict=strict, exact_pos=True)


@needs_ref
@pytest.mark.parametrize("shape", [(30, 42, 38), (6, 8, 10), (2, 2, 2), (2, 9, 3), (17, 2, 5)])
def test_noncubic_and_tiny(shape):
    a = noise_grid(0, "f32", shape=shape)
    _check(a, 0.0, "f32", exact_pos=True)
    b = noise_grid(0, "u8", scale=5, shape=shape)
    _check(b, 2.0, "u8", exact_pos=True)


@needs_ref
@pytest.mark.parametrize("variant", ["f32", "f64", "u8"])
def test_store_variants(variant):
    scale = 6
    a = noise_grid(40, variant, scale=scale)
    iso = 3.0 if variant == "u8" else 0.0
    # spnA, spnB: r*D+O is evaluated identically -> bit exact positions
    _check(a, iso, variant, Geometry(r0=(1, 2, 3), d=(.5, .5, .5)), exact_pos=True)
    _check(a, iso, variant, Geometry(r0=(1, 2, 3), d=(.5, .25, 2)), exact_pos=True)
    # spnC: 3x3 double mat-vec then narrowed
    _check(a, iso, variant, inclined_geom(), pos_rtol=2e-6)
    _check(a, iso, variant, inclined_geom(tsa=1, d=(.5, .25, 2), r0=(1, 2, 3)), pos_rtol=2e-6)


@needs_ref
def test_smooth_fields():
    g = gyroid_grid(64, periods=2)
    for iso in (-1.2, -0.3, 0.0, 0.9):
        _check(g, iso, "f32", exact_pos=True)
    c = ct_grid(48)
    for iso in (1500.0, 1500.5):
        _check(c, iso, "u16", exact_pos=True)


@needs_ref
def test_empty_and_full():
    a = np.zeros((5, 6, 7), np.float32)
    for iso in (1.0, -1.0, 0.0):
        ref = ref_lib("f32").extract(a, iso)
        orc = oracle_extract(a, iso, "f32")
        assert ref.nV == orc.nV == 0 and ref.nT == orc.nT == 0
    # reference returns a zeroed struct for an empty surface (marching_cubes_33.c:1880-1883)
    assert ref_lib("f32").extract(a, 1.0).iso == 0.0


@needs_ref
def test_size_of_isosurface_agrees():
    a = noise_grid(40, "u8", scale=6)
    sz, nV, nT = ref_lib("u8").size(a, 3.0)
    m = oracle_extract(a, 3.0, "u8", count_only=True)
    assert (nV, nT) == (m.nV, m.nT)
    assert sz == nV * (6 * 4 + 4) + nT * 12 + 64


@needs_ref
def test_plateaus_on_iso():
    """large on-iso plateaus: many POINT vertices, dropped zero-area triangles"""
    rng = np.random.default_rng(3)
    a = rng.integers(0, 3, size=(20, 22, 24)).astype(np.uint8)
    for iso in (0.0, 1.0, 2.0):
        _check(a, iso, "u8", exact_pos=True)
    f = rng.integers(-1, 2, size=(18, 18, 18)).astype(np.float32)
    for iso in (0.0, 1.0, -1.0):
        _check(f, iso, "f32", exact_pos=True)
