"""The C-ABI shared library builds, loads and exports every symbol include/swm_orb.h declares; without a
GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "swm_orb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(swm_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(swm):
    lib = swm.load()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/swm_orb.h but not exported"
    bound = {s[0] for s in swm.SYMBOLS}
    assert set(declared) == bound, set(declared) ^ bound


def test_keypoint_layout_matches_cv_keypoint(swm):
    # cv::KeyPoint: Point2f pt; float size, angle, response; int octave, class_id -> 28 bytes
    assert swm.KP_DTYPE.itemsize == 28
    assert [swm.KP_DTYPE.fields[n][1] for n in ("x", "y", "size", "angle", "response", "octave", "class_id")] == \
        [0, 4, 8, 12, 16, 20, 24]


def test_version_string(swm):
    assert swm.load().swm_version().decode().startswith("swm_orb")


def test_invalid_config_rejected(swm):
    lib = swm.load()
    h = C.c_void_p()
    for cfg in (swm.OrbCfg(0, 1.2, 8, 20, 7, 1, 0), swm.OrbCfg(1000, 1.0, 8, 20, 7, 1, 0),
                swm.OrbCfg(1000, 1.2, 0, 20, 7, 1, 0), swm.OrbCfg(1000, 1.2, 8, 5, 7, 1, 0),
                swm.OrbCfg(1000, 1.2, 8, 20, 7, 0, 0)):
        assert lib.swm_orb_create(C.byref(cfg), 0, C.byref(h)) == -1
        assert not h.value


def test_no_cpu_fallback(swm):
    """On a box without a GPU the product must refuse to run rather than fall back to anything."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present: the loud-failure path is exercised on the CPU-only box")
    lib = swm.load()
    h = C.c_void_p()
    cfg = swm.OrbCfg(1000, 1.2, 8, 20, 7, 1, 0)
    assert lib.swm_orb_create(C.byref(cfg), 0, C.byref(h)) == -3  # SWM_E_NODEVICE
    assert b"no CPU fallback" in lib.swm_last_error(None)
    m = C.c_void_p()
    assert lib.swm_matcher_create(0, C.byref(m)) == -3
    a = np.zeros((4, 32), np.uint8)
    out = np.zeros((4, 4), np.uint16)
    assert lib.swm_hamming_matrix(swm.ptr(a), 4, swm.ptr(a), 4, swm.ptr(out), 0) == -3
    from swarmmap_b200.orb import ORBextractor
    with pytest.raises(swm.SwmError):
        ORBextractor(1000, 1.2, 8, 20, 7)


def test_product_never_imports_oracle():
    """Static check: nothing under swarmmap_b200/ references the oracle."""
    pkg = os.path.join(ROOT, "swarmmap_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".inc")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "orb_oracle" not in txt and "liborb_oracle" not in txt, f
