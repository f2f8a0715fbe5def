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
    # the later entry points: resident frames, vocabulary, distinctive descriptors, camera bounds
    f = C.c_void_p()
    assert lib.swm_frame_create(0, C.byref(f)) == -3 and not f.value
    from swarmmap_b200 import synth
    blob = np.frombuffer(synth.make_vocabulary(4, 2, seed=1), np.uint8)
    v = C.c_void_p()
    assert lib.swm_vocab_create(0, swm.ptr(blob), len(blob), C.byref(v)) == -3 and not v.value
    off = np.array([0, 4], np.int32)
    best = np.zeros(1, np.int32)
    assert lib.swm_distinctive_descriptors(swm.ptr(a), swm.ptr(off), 1, swm.ptr(best), None, 0) == -3
    cam = swm.Camera(458.654, 457.296, 367.215, 248.375, -0.28, 0.07, 0.0, 0.0, 0.0)
    b4 = np.zeros(4, np.float32)
    assert lib.swm_camera_bounds(0, C.byref(cam), 752, 480, swm.ptr(b4)) == -3
    from swarmmap_b200.bow import ORBVocabulary
    from swarmmap_b200.matcher import ResidentFrame
    with pytest.raises(swm.SwmError):
        ResidentFrame()
    with pytest.raises(swm.SwmError):
        ORBVocabulary(blob.tobytes())


def test_vocabulary_blob_validation(swm):
    """Malformed vocabulary files are rejected before any device work (host-side parser of the ORBvoc.bin layout)."""
    from swarmmap_b200 import synth
    lib = swm.load()
    v = C.c_void_p()
    good = bytearray(synth.make_vocabulary(4, 2, seed=1))
    short = np.frombuffer(bytes(good[:10]), np.uint8)
    assert lib.swm_vocab_create(0, swm.ptr(short), len(short), C.byref(v)) == -1
    bad_size = bytearray(good)
    bad_size[4:8] = np.array([40], np.uint32).tobytes()
    b = np.frombuffer(bytes(bad_size), np.uint8)
    assert lib.swm_vocab_create(0, swm.ptr(b), len(b), C.byref(v)) == -1
    assert b"record size" in lib.swm_vocab_last_error(None)
    bad_parent = bytearray(good)
    bad_parent[24:28] = np.array([10 ** 6], np.int32).tobytes()
    b = np.frombuffer(bytes(bad_parent), np.uint8)
    assert lib.swm_vocab_create(0, swm.ptr(b), len(b), C.byref(v)) == -1
    bad_scoring = bytearray(good)
    bad_scoring[16:20] = np.array([9], np.int32).tobytes()
    b = np.frombuffer(bytes(bad_scoring), np.uint8)
    assert lib.swm_vocab_create(0, swm.ptr(b), len(b), C.byref(v)) == -1
    assert not v.value


def test_product_never_imports_oracle():
    """Static check: nothing under swarmmap_b200/ references the oracle."""
    pkg = os.path.join(ROOT, "swarmmap_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".inc")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "orb_oracle" not in txt and "liborb_oracle" not in txt, f


def test_opencv_branch_compiles():
    """The SWM_HAVE_OPENCV branch of the drop-in extractor header (cv::InputArray / OutputArray, mvImagePyramid and
    mvImagePyramidBorder as cv::cuda::GpuMat headers over the device planes) compiles against headers that declare the
    OpenCV signatures it uses (the images have no OpenCV C++ headers; oracle/ref_shim stands in)."""
    import subprocess
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-DSWM_HAVE_OPENCV", "-I", os.path.join(ROOT, "oracle", "ref_shim"),
                        os.path.join(ROOT, "tests", "opencv_build_check.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def test_batch_and_sharded_entry_points_reject_bad_arguments(swm):
    """Argument validation of the round-2 entry points happens before any device work."""
    lib = swm.load()
    assert lib.swm_match_window_batch(None, None, 0) == -1
    assert lib.swm_match_init_batch(None, None, 0) == -1
    assert lib.swm_match_bow_batch(None, None, 0) == -1
    assert lib.swm_frames_from_extractor(None, 0, None, None, None, None) == -1
    assert lib.swm_db_query_sharded(None, None, 2, None, 0, 2, None, None, 50, None) == -1
    t = C.c_double(0)
    assert lib.swm_i8_peak(0, 7, 1000, C.byref(t)) == -1          # unknown mode
    assert lib.swm_i8_peak(0, 0, 1000, C.byref(t)) in (-3, 0)     # no device here / measured on a GPU box
    assert lib.swm_matcher_last_device_ms(None) == 0.0
