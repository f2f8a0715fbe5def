"""Closed-form tables derivable from the reference source."""
import numpy as np


def test_umax(oracle):
    # derived from ORBextractor.cc:386-401
    assert oracle.umax().tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def test_level_quotas(oracle):
    # ORBextractor.cc:367-378
    assert oracle.level_quotas(1000).tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert oracle.level_quotas(2000).tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert oracle.level_quotas(4000).tolist() == [869, 724, 603, 503, 419, 349, 291, 242]
    for n in (500, 1000, 1500, 2000):
        assert oracle.level_quotas(n).sum() == n


def test_level_sizes(oracle):
    # cvRound(cols * invScale), ORBextractor.cc:825-826
    ws, hs = oracle.level_sizes(752, 480)
    assert ws.tolist() == [752, 627, 522, 435, 363, 302, 252, 210]
    assert hs.tolist() == [480, 400, 333, 278, 231, 193, 161, 134]
    assert int((ws.astype(np.int64) * hs).sum()) == 1117367
    assert int(((ws + 38).astype(np.int64) * (hs + 38)).sum()) == 1344493
    ws, hs = oracle.level_sizes(1241, 376)
    assert ws.tolist() == [1241, 1034, 862, 718, 598, 499, 416, 346]
    assert hs.tolist() == [376, 313, 261, 218, 181, 151, 126, 105]
    assert int((ws.astype(np.int64) * hs).sum()) == 1444097


def test_scale_tables(oracle):
    sf, inv, s2, inv2 = oracle.scale_tables(1.2, 8)
    assert sf[0] == 1.0 and abs(sf[7] - 1.2 ** 7) < 1e-5
    np.testing.assert_allclose(sf * inv, 1.0, rtol=1e-6)
    np.testing.assert_array_equal(s2, sf * sf)
    # float chain with a double-held scale factor: level 1 is exactly float(1.2f as double)
    assert sf[1] == np.float32(np.float64(np.float32(1.2)))


def test_popcount_matches_numpy(oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = rng.integers(0, 256, 32, dtype=np.uint8)
        b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert oracle.hamming256(a, b) == int(np.unpackbits(a ^ b).sum())


def test_pattern_tables_identical_and_checksum():
    import hashlib
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = []
    for rel in ("oracle/orb_pattern_data.inc", "swarmmap_b200/csrc/orb_pattern.inc"):
        src = open(os.path.join(root, rel)).read()
        vals = [int(t) for t in re.findall(r"-?\d+", src[src.index("{"):src.index("};")])]
        assert len(vals) == 1024 and max(abs(v) for v in vals) == 13
        digests.append(hashlib.sha256(bytes((v + 256) % 256 for v in vals)).hexdigest())
    assert digests[0] == digests[1] == "2164181aea6ff9ac426ca512d5130d15e1f6e3cd47b1cbdd568bbe1e55d49023"
