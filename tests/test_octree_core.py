"""The product's block-parallel quadtree (swarmmap_b200/csrc/octree_core.cuh) compiled as a serial host
emulation and checked against the oracle's literal std::list restatement of DistributeOctTree."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness():
    so = os.path.join(ROOT, "tests", "libhost_harness.so")
    src = os.path.join(ROOT, "tests", "host_harness.cpp")
    deps = [src, os.path.join(ROOT, "swarmmap_b200", "csrc", "octree_core.cuh"),
            os.path.join(ROOT, "swarmmap_b200", "csrc", "swm_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-x", "c++", src, "-o", so], check=True)
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _run_both(oracle, harness, xs, ys, sc, W, H, N):
    n = len(xs)
    pts = np.zeros(n, oracle.FASTPT_DTYPE)
    pts["x"], pts["y"], pts["score"] = xs, ys, sc
    ref = oracle.octree(pts, 16, 16 + W, 16, 16 + H, N)
    packed = ((ys.astype(np.uint32) << 20) | (xs.astype(np.uint32) << 8) | sc.astype(np.uint32)).astype(np.uint32)
    out = np.zeros(n + 16, np.uint32)
    k = harness.hh_octree(_p(packed), n, W, H, N, _p(out), len(out))
    refp = (ref["y"].astype(np.uint32) << 20) | (ref["x"].astype(np.uint32) << 8) | ref["score"].astype(np.uint32)
    return refp, out[:k]


def test_quadtree_core_matches_oracle_random(oracle, harness):
    rng = np.random.default_rng(5)
    done = 0
    for trial in range(400):
        W, H = int(rng.integers(40, 1300)), int(rng.integers(40, 600))
        if not 1 <= round(W / H) <= 16:
            continue
        n = int(rng.integers(0, 3000)) if trial % 4 else int(rng.integers(0, 30))
        N = int(rng.integers(1, 1100)) if trial % 5 else int(rng.integers(0, 20))
        if trial % 4 == 2:  # clustered points
            cx, cy = rng.integers(3, W - 3, 5), rng.integers(3, H - 3, 5)
            xs = np.clip((cx[rng.integers(0, 5, n)] + rng.normal(0, 20, n)).astype(int), 3, W - 4)
            ys = np.clip((cy[rng.integers(0, 5, n)] + rng.normal(0, 20, n)).astype(int), 3, H - 4)
        else:
            xs, ys = rng.integers(3, W - 3, n), rng.integers(3, H - 3, n)
        key = np.unique(ys.astype(np.int64) * 4096 + xs)  # distinct pixels, raster order
        ys, xs = (key // 4096).astype(np.int16), (key % 4096).astype(np.int16)
        sc = rng.integers(7, 40 if trial % 4 == 3 else 256, len(key)).astype(np.int32)  # many response ties
        ref, got = _run_both(oracle, harness, xs, ys, sc, W, H, N)
        np.testing.assert_array_equal(got, ref, err_msg=f"trial {trial}: W={W} H={H} n={len(key)} N={N}")
        done += 1
    assert done > 300


def test_quadtree_core_on_real_candidates(oracle, harness):
    from swarmmap_b200 import synth
    for (w, h, nf, seed) in ((752, 480, 1000, 20220404), (1241, 376, 4000, 20220405)):
        ex = oracle.Extractor(nf)
        ex(synth.make_frame(w, h, seed))
        ws, hs = oracle.level_sizes(w, h)
        quotas = oracle.level_quotas(nf)
        for l in range(8):
            p = ex.level_fast(l)
            ref, got = _run_both(oracle, harness, p["x"], p["y"], p["score"], int(ws[l]) - 32, int(hs[l]) - 32,
                                 int(quotas[l]))
            np.testing.assert_array_equal(got, ref)
            sel = ex.level_selected(l)
            assert len(sel) == len(ref) and quotas[l] <= len(sel) <= quotas[l] + 3


def test_fast_score_core_matches_oracle(oracle, harness):
    from swarmmap_b200 import synth
    rng = np.random.default_rng(8)
    for img in (synth.make_frame(320, 240, 3), rng.integers(0, 256, (200, 300), dtype=np.uint8)):
        h, w = img.shape
        out = np.zeros_like(img)
        harness.hh_fast_score_map(_p(img), w, h, w, 7, _p(out))
        np.testing.assert_array_equal(out, oracle.fast_score_map(img, 7))
