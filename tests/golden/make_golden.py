#!/usr/bin/env python3
"""Generates the committed golden fixtures from the CPU oracle (the reference ships none).

    python tests/golden/make_golden.py

Inputs are re-generated from seeds by swarmmap_b200.synth, so only outputs are stored:
  extract_euroc.npz   config 1: 752x480 seed 20220404, ORBextractor(1000,1.2,8,20,7)
  extract_kitti.npz   config 2: 1241x376 seed 20220405, ORBextractor(2000,...)
  match_init.npz      config 2: SearchForInitialization(F0, Fk, prev, 100), ORBmatcher(0.9,true), k=1..3
Each also stores a SHA-256 of the input frame(s) so a drift of the generator is caught separately.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as oracle  # noqa: E402
from swarmmap_b200 import synth  # noqa: E402
from swarmmap_b200.matcher import Frame  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def extract_fixture(name, w, h, seed, nf):
    img = synth.make_frame(w, h, seed)
    ex = oracle.Extractor(nf, 1.2, 8, 20, 7)
    kps, desc = ex(img)
    lvl = {f"fast_count_{l}": len(ex.level_fast(l)) for l in range(8)}
    planes = {f"plane_sha_{l}": sha(ex.level(l, 0)) for l in range(8)}
    blur = {f"blur_sha_{l}": sha(ex.level(l, 1)) for l in range(8)}
    np.savez_compressed(os.path.join(HERE, name), frame_sha=sha(img), kps=kps, desc=desc, w=w, h=h, seed=seed,
                        nfeatures=nf, **lvl, **planes, **blur)
    print(name, len(kps), "keypoints")


def match_fixture():
    w, h, seed = 1241, 376, 20220405
    seq = synth.make_sequence(4, w, h, seed)
    ex = oracle.Extractor(4000, 1.2, 8, 20, 7)
    sf = oracle.scale_tables(1.2, 8)[0]
    fs = [Frame.from_keypoints(*ex(img), w, h, sf) for img in seq]
    prev = np.stack([fs[0].x, fs[0].y], 1).astype(np.float32)
    out = {"seq_sha": sha(seq), "n_kp": np.array([f.N for f in fs])}
    for k in (1, 2, 3):
        n, m12, prev = oracle.search_for_initialization(fs[0], fs[k], prev, 100, 0.9, True)
        out[f"n_{k}"] = n
        out[f"m12_{k}"] = m12
        out[f"prev_{k}"] = prev.copy()
        print("match_init k", k, n)
    np.savez_compressed(os.path.join(HERE, "match_init.npz"), **out)


if __name__ == "__main__":
    extract_fixture("extract_euroc.npz", 752, 480, 20220404, 1000)
    extract_fixture("extract_kitti.npz", 1241, 376, 20220405, 2000)
    match_fixture()
