"""DBoW2 transform on the GPU against the oracle's std::map restatement: word ids,
TF-IDF values (bit-exact doubles), FeatureVector nodes / feature lists, for every weighting / scoring type, irregular
trees (early leaves), stopped words (weight 0), ragged batches and resident frames."""
import numpy as np
import pytest

from swarmmap_b200 import synth

pytestmark = pytest.mark.gpu


def _descriptors(blob, n, seed):
    """Descriptors near random vocabulary leaves + pure noise + exact leaf copies (distance ties between siblings)."""
    rng = np.random.default_rng(seed)
    rec = np.frombuffer(blob[24:], np.dtype([("parent", "<i4"), ("desc", "u1", 32), ("weight", "<f4"), ("leaf", "u1")]))
    leaf = rec[rec["leaf"] == 1]["desc"]
    d = leaf[rng.integers(0, len(leaf), n)].copy()
    d[: n // 2, :6] ^= rng.integers(0, 256, (n // 2, 6), dtype=np.uint8)
    d[n // 2: n // 2 + n // 8] = rng.integers(0, 256, (n // 8, 32), dtype=np.uint8)
    return d


def _check(res, ref):
    w, wv, nid, off, feats = ref
    np.testing.assert_array_equal(res.word_ids, w)
    np.testing.assert_array_equal(res.values, wv)  # doubles, bit for bit
    np.testing.assert_array_equal(res.node_ids, nid)
    np.testing.assert_array_equal(res.offsets, off)
    np.testing.assert_array_equal(res.feats, feats)


@pytest.mark.parametrize("k,L,levelsup", [(10, 4, 2), (10, 3, 4), (7, 5, 4), (20, 2, 1)])
def test_transform_matches_oracle(oracle, swm, k, L, levelsup):
    from swarmmap_b200.bow import ORBVocabulary
    blob = synth.make_vocabulary(k, L, seed=k * 10 + L)
    ref = oracle.Vocabulary(blob)
    voc = ORBVocabulary(blob)
    assert (voc.k, voc.L, voc.n_nodes, voc.n_words) == (ref.k, ref.L, ref.n_nodes, ref.n_words)
    for n in (1, 37, 1000, 4003):
        d = _descriptors(blob, n, n)
        _check(voc.transform(d, levelsup), ref.transform(d, levelsup))
    assert voc.transform(np.zeros((0, 32), np.uint8)).word_ids.size == 0


@pytest.mark.parametrize("weighting", [0, 1, 2, 3])
@pytest.mark.parametrize("scoring", [0, 1, 2, 5])
def test_weighting_and_scoring_types(oracle, swm, weighting, scoring):
    """TF_IDF / TF accumulate, IDF / BINARY keep the first weight (BowVector.cpp:33-58); L1 / L2 / no normalisation by
    scoring object (ScoringObject.h:76-91)."""
    from swarmmap_b200.bow import ORBVocabulary
    blob = synth.make_vocabulary(10, 3, seed=5, weighting=weighting, scoring=scoring)
    ref, voc = oracle.Vocabulary(blob), ORBVocabulary(blob)
    d = _descriptors(blob, 2500, 9)
    _check(voc.transform(d, 2), ref.transform(d, 2))


def test_batch_and_resident_frame(oracle, swm):
    from swarmmap_b200.bow import ORBVocabulary
    from swarmmap_b200.matcher import ORBmatcher, ResidentFrame
    from swarmmap_b200.orb import ORBextractor
    blob = synth.make_vocabulary(10, 4, seed=77)
    ref, voc = oracle.Vocabulary(blob), ORBVocabulary(blob)
    seq = synth.make_sequence(3, 752, 480, 5)
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=3)
    kps, desc, n = ex.extract_batch(np.stack(seq))
    out = voc.transform_batch(desc, n, 4)            # ragged batch straight from the extractor's host output
    for b in range(3):
        _check(out[b], ref.transform(desc[b, :n[b]], 4))
    frames = [ResidentFrame().from_extractor(ex, b, None, np.array([0, 752, 0, 480], np.float32)) for b in range(3)]
    res = [voc.transform_frame(f, 4) for f in frames]  # descriptors never leave the device
    for b in range(3):
        _check(res[b], ref.transform(desc[b, :n[b]], 4))
    # the FeatureVectors feed SearchByBoW on the resident frames; same matches as the oracle on host arrays
    from swarmmap_b200.matcher import Frame
    host = [Frame.from_keypoints(kps[b, :n[b]], desc[b, :n[b]], 752, 480, ex.GetScaleFactors()) for b in range(3)]
    m = ORBmatcher(0.7, True)
    valid = np.ones(host[0].N, np.uint8)
    got = m.SearchByBoW(frames[0], res[0].feature_vector(), valid, frames[1], res[1].feature_vector())
    exp = oracle.search_by_bow(host[0], res[0].feature_vector(), valid, host[1], res[1].feature_vector(), None, 0, 0.7, True)
    assert got[0] == exp[0]
    np.testing.assert_array_equal(got[1], exp[1])


def test_vocab_errors(swm):
    from swarmmap_b200._lib import SwmError
    from swarmmap_b200.bow import ORBVocabulary
    with pytest.raises(SwmError):
        ORBVocabulary(b"\x00" * 10)
    blob = bytearray(synth.make_vocabulary(5, 2, seed=1))
    blob[4:8] = np.array([40], np.uint32).tobytes()  # wrong record size
    with pytest.raises(SwmError):
        ORBVocabulary(bytes(blob))
    voc = ORBVocabulary(synth.make_vocabulary(5, 2, seed=1))
    with pytest.raises(SwmError):
        voc.transform_batch(np.zeros((1, 9000, 32), np.uint8), np.array([9000], np.int32))
