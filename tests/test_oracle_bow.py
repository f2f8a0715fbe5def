"""The oracle's DBoW2 restatement (oracle/orb_oracle_match.cpp: orc_vocab_load / orc_bow_transform) against a
second, independent pure-Python reading of the reference sources (TemplatedVocabulary.h:1151-1283, :1478-1522,
BowVector.cpp, FeatureVector.cpp) on a small synthetic vocabulary.  The real ORBvoc.bin is not shipped with the
reference and DBoW2 needs OpenCV to compile, so there is no golden vector from the reference itself: parity for this
row is "unpinned by the reference's own tests" like the rest of the path."""
import numpy as np
import pytest

from swarmmap_b200 import synth

REC = np.dtype([("parent", "<i4"), ("desc", "u1", 32), ("weight", "<f4"), ("leaf", "u1")])


def py_transform(blob, desc, levelsup):
    hdr = np.frombuffer(blob[:24], np.int32)
    L, scoring, weighting = int(hdr[3]), int(hdr[4]), int(hdr[5])
    rec = np.frombuffer(blob[24:], REC)
    n_nodes = len(rec) + 1
    children = [[] for _ in range(n_nodes)]
    word = {}
    for r in range(len(rec)):
        children[rec["parent"][r]].append(r + 1)
        if rec["leaf"][r]:
            word[r + 1] = len(word)
    bits = np.unpackbits(rec["desc"], axis=1)
    bow, fv = {}, {}
    for i, d in enumerate(desc):
        db = np.unpackbits(d)
        node, level, nid = 0, 0, (0 if L - levelsup <= 0 else None)
        while children[node]:
            level += 1
            dist = [int((bits[c - 1] != db).sum()) for c in children[node]]
            node = children[node][int(np.argmin(dist))]  # first minimum = strict '<' scan
            if level == L - levelsup:
                nid = node
        if nid is None:
            nid = node
        w = float(rec["weight"][node - 1])
        if w > 0:
            if word[node] in bow:
                if weighting in (0, 1):
                    bow[word[node]] += w
            else:
                bow[word[node]] = w
            fv.setdefault(nid, []).append(i)
    ids = sorted(bow)
    vals = [bow[k] for k in ids]
    if scoring == 5:
        if weighting in (0, 1) and ids:
            vals = [v / float(len(ids)) for v in vals]
    else:
        norm = 0.0
        for v in vals:
            norm = norm + (abs(v) if scoring != 1 else v * v)
        if scoring == 1:
            norm = float(np.sqrt(norm))
        if norm > 0:
            vals = [v / norm for v in vals]
    return ids, vals, {k: fv[k] for k in sorted(fv)}


@pytest.mark.parametrize("weighting,scoring", [(0, 0), (1, 1), (2, 0), (3, 5), (0, 5)])
def test_oracle_transform_vs_python(oracle, weighting, scoring):
    blob = synth.make_vocabulary(6, 3, seed=11, early_leaf=0.1, zero_weight=0.1, weighting=weighting, scoring=scoring)
    voc = oracle.Vocabulary(blob)
    rng = np.random.default_rng(2)
    rec = np.frombuffer(blob[24:], REC)
    leaf = rec[rec["leaf"] == 1]["desc"]
    desc = leaf[rng.integers(0, len(leaf), 400)].copy()
    desc[:300, :3] ^= rng.integers(0, 256, (300, 3), dtype=np.uint8)
    for levelsup in (1, 2, 4):
        w, wv, nid, off, feats = voc.transform(desc, levelsup)
        ids, vals, fv = py_transform(blob, desc, levelsup)
        np.testing.assert_array_equal(w, ids)
        np.testing.assert_array_equal(wv, np.array(vals, np.float64))
        assert list(nid) == list(fv)
        for j, k in enumerate(fv):
            assert list(feats[off[j]:off[j + 1]]) == fv[k]
    if scoring == 0:
        assert abs(wv.sum() - 1.0) < 1e-12
