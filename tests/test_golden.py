"""Committed golden fixtures (tests/golden/, produced by make_golden.py from the oracle).
CPU: the oracle still reproduces them (guards the checker itself).  GPU: the CUDA path reproduces
them without the oracle in the loop."""
import hashlib
import os

import numpy as np
import pytest

from swarmmap_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _load(name):
    return np.load(os.path.join(GOLD, name))


@pytest.mark.parametrize("name", ["extract_euroc.npz", "extract_kitti.npz"])
def test_oracle_reproduces_extract_golden(oracle, name):
    g = _load(name)
    img = synth.make_frame(int(g["w"]), int(g["h"]), int(g["seed"]))
    assert sha(img) == str(g["frame_sha"]), "synthetic generator drifted"
    ex = oracle.Extractor(int(g["nfeatures"]), 1.2, 8, 20, 7)
    kps, desc = ex(img)
    assert kps.tobytes() == g["kps"].tobytes()
    np.testing.assert_array_equal(desc, g["desc"])
    for l in range(8):
        assert sha(ex.level(l, 0)) == str(g[f"plane_sha_{l}"])
        assert sha(ex.level(l, 1)) == str(g[f"blur_sha_{l}"])
        assert len(ex.level_fast(l)) == int(g[f"fast_count_{l}"])


def test_oracle_reproduces_match_golden(oracle):
    from swarmmap_b200.matcher import Frame
    g = _load("match_init.npz")
    seq = synth.make_sequence(4, 1241, 376, 20220405)
    assert sha(seq) == str(g["seq_sha"])
    ex = oracle.Extractor(4000, 1.2, 8, 20, 7)
    sf = oracle.scale_tables(1.2, 8)[0]
    fs = [Frame.from_keypoints(*ex(img), 1241, 376, sf) for img in seq]
    assert [f.N for f in fs] == g["n_kp"].tolist()
    prev = np.stack([fs[0].x, fs[0].y], 1).astype(np.float32)
    for k in (1, 2, 3):
        n, m12, prev = oracle.search_for_initialization(fs[0], fs[k], prev, 100, 0.9, True)
        assert n == int(g[f"n_{k}"])
        np.testing.assert_array_equal(m12, g[f"m12_{k}"])
        np.testing.assert_array_equal(prev, g[f"prev_{k}"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["extract_euroc.npz", "extract_kitti.npz"])
def test_gpu_reproduces_extract_golden(swm, name):
    from swarmmap_b200.orb import ORBextractor
    g = _load(name)
    img = synth.make_frame(int(g["w"]), int(g["h"]), int(g["seed"]))
    ex = ORBextractor(int(g["nfeatures"]), 1.2, 8, 20, 7)
    kps, desc = ex(img)
    gk = g["kps"]
    assert len(kps) == len(gk)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        np.testing.assert_array_equal(kps[f], gk[f])
    d = np.abs(kps["angle"] - gk["angle"])
    assert np.deg2rad(np.minimum(d, 360 - d).max()) <= 1e-4
    agree = 1.0 - np.unpackbits(desc ^ g["desc"]).sum() / float(desc.size * 8)
    assert agree >= 0.999
    for l in range(8):
        assert sha(ex.debug_plane(0, l, 0)) == str(g[f"plane_sha_{l}"])
        assert sha(ex.debug_plane(0, l, 1)) == str(g[f"blur_sha_{l}"])
        assert min(len(ex.debug_points(0, l, 0)), 10000) == int(g[f"fast_count_{l}"])


@pytest.mark.gpu
def test_gpu_extract_then_match_golden(swm):
    """extract (GPU) -> SearchForInitialization (GPU) end to end against the golden match indices.
    Descriptors may differ from the oracle's in <0.1 % of bits, so allow a handful of match flips."""
    from swarmmap_b200.matcher import Frame, ORBmatcher
    from swarmmap_b200.orb import ORBextractor
    g = _load("match_init.npz")
    seq = synth.make_sequence(4, 1241, 376, 20220405)
    ex = ORBextractor(4000, 1.2, 8, 20, 7)
    fs = [Frame.from_keypoints(*ex(img), 1241, 376) for img in seq]
    assert [f.N for f in fs] == g["n_kp"].tolist()
    m = ORBmatcher(0.9, True)
    prev = np.stack([fs[0].x, fs[0].y], 1).astype(np.float32).copy()
    for k in (1, 2, 3):
        n, m12 = m.SearchForInitialization(fs[0], fs[k], prev, 100)
        same = (m12 == g[f"m12_{k}"]).mean()
        assert same >= 0.995 and abs(n - int(g[f"n_{k}"])) <= 5, (k, n, int(g[f"n_{k}"]), same)
        prev = g[f"prev_{k}"].copy()
