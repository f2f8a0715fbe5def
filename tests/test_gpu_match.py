"""GPU parity of the Hamming matchers against the CPU oracle (bit-exact distances and match indices)."""
import numpy as np
import pytest

from swarmmap_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frames(oracle):
    """Keypoints/descriptors of a short KITTI-shaped and a EuRoC-shaped synthetic sequence (CPU oracle)."""
    from swarmmap_b200.matcher import Frame
    out = {}
    for name, (w, h, nf, seed) in {"kitti": (1241, 376, 4000, 20220405), "euroc": (752, 480, 1000, 20220406)}.items():
        seq = synth.make_sequence(4, w, h, seed)
        ex = oracle.Extractor(nf, 1.2, 8, 20, 7)
        sf = oracle.scale_tables(1.2, 8)[0]
        fs = []
        for img in seq:
            k, d = ex(img)
            fs.append(Frame.from_keypoints(k, d, w, h, sf))
        out[name] = fs
    return out


def test_descriptor_distance(oracle, swm):
    from swarmmap_b200.matcher import ORBmatcher
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (517, 32), dtype=np.uint8)
    b[:50] = a[:50]
    b[50:100] = a[50:100] ^ np.uint8(1)
    m = ORBmatcher()
    ref = np.unpackbits(a[:, None, :] ^ b[None, :, :], axis=2).sum(2)
    np.testing.assert_array_equal(m.distance_matrix(a, b), ref)
    pair = m.DescriptorDistance(a, b[:300])
    np.testing.assert_array_equal(pair, np.diag(ref[:, :300]))
    assert m.DescriptorDistance(a[0], a[0]) == 0
    assert m.DescriptorDistance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256
    for i in range(20):
        assert oracle.hamming256(a[i], b[i]) == ref[i, i]


def test_grid_matches_oracle(oracle, swm, frames):
    from swarmmap_b200.matcher import ORBmatcher
    m = ORBmatcher()
    for f in frames["kitti"][:2] + frames["euroc"][:1]:
        s, it = m.grid(f)
        os_, oit = oracle.grid_csr(f)
        np.testing.assert_array_equal(s, os_)
        np.testing.assert_array_equal(it, oit)


@pytest.mark.parametrize("name,window", [("kitti", 100), ("euroc", 100), ("euroc", 30)])
def test_search_for_initialization(oracle, swm, frames, name, window):
    from swarmmap_b200.matcher import ORBmatcher
    fs = frames[name]
    m = ORBmatcher(0.9, True)
    f1 = fs[0]
    prev_g = np.stack([f1.x, f1.y], 1).astype(np.float32).copy()
    prev_o = prev_g.copy()
    for k in (1, 2, 3):  # as Tracking::MonocularInitialization does: same F1, successive F2, prev carried over
        n, m12 = m.SearchForInitialization(f1, fs[k], prev_g, window)
        on, om12, prev_o = oracle.search_for_initialization(f1, fs[k], prev_o, window, 0.9, True)
        assert n == on
        np.testing.assert_array_equal(m12, om12)
        np.testing.assert_array_equal(prev_g, prev_o)
        assert n > 50
    m2 = ORBmatcher(0.9, False)
    prev = np.stack([f1.x, f1.y], 1).astype(np.float32).copy()
    n, m12 = m2.SearchForInitialization(f1, fs[1], prev, window)
    on, om12, _ = oracle.search_for_initialization(f1, fs[1], np.stack([f1.x, f1.y], 1), window, 0.9, False)
    assert n == on and (m12 == om12).all()


def _window_case(rng, src, tgt, th, band, jitter):
    sf = tgt.mvScaleFactors
    u = src.x + rng.normal(0, jitter, src.N).astype(np.float32)
    v = src.y + rng.normal(0, jitter, src.N).astype(np.float32)
    radius = (np.float32(th) * sf[src.octave]).astype(np.float32)
    lo, hi = src.octave + band[0], src.octave + band[1]
    valid = (rng.random(src.N) < 0.8).astype(np.uint8)
    blocks = (rng.random(src.N) < 0.7).astype(np.uint8)
    return u, v, radius, lo, hi, valid, blocks


@pytest.mark.parametrize("ratio_mode,ori,th_dist", [(0, True, 100), (1, False, 100), (0, False, 50), (0, True, 64)])
def test_match_window(oracle, swm, frames, ratio_mode, ori, th_dist):
    from swarmmap_b200.matcher import ORBmatcher
    rng = np.random.default_rng(ratio_mode * 7 + th_dist)
    m = ORBmatcher(0.8, ori)
    for name in ("euroc", "kitti"):
        src, tgt = frames[name][0], frames[name][1]
        band = (-1, 0) if ratio_mode else (-1, 1)
        u, v, radius, lo, hi, valid, blocks = _window_case(rng, src, tgt, 15 if ratio_mode == 0 else 4, band, 3.0)
        tgt_blocked = (rng.random(tgt.N) < 0.1).astype(np.uint8)
        asg0 = np.full(tgt.N, -1, np.int32)
        asg0[rng.random(tgt.N) < 0.05] = 7  # pre-existing assignments must survive unless overwritten
        n, asg = m.match_window(tgt, src.desc, u, v, radius, lo, hi, valid, blocks, th_dist, ratio_mode, src.angle,
                                tgt_blocked, asg0.copy(), check_ori=ori)
        on, oasg = oracle.match_window(tgt, src.desc, u, v, radius, lo, hi, valid, blocks, th_dist, ratio_mode, 0.8,
                                       ori, src.angle, tgt_blocked, asg0)
        assert n == on
        np.testing.assert_array_equal(asg, oasg)
        assert n > 20
    # level filter disabled (-1,-1) and a window that leaves the image
    src, tgt = frames["euroc"][0], frames["euroc"][2]
    u = src.x.copy(); v = src.y.copy()
    u[:50] = -500; v[50:100] = 5000; u[100:150] = 751.5
    neg = np.full(src.N, -1, np.int32)
    ones = np.ones(src.N, np.uint8)
    rad = np.full(src.N, 20, np.float32)
    n, asg = m.match_window(tgt, src.desc, u, v, rad, neg, neg, ones, ones, th_dist, ratio_mode, src.angle,
                            check_ori=ori)
    on, oasg = oracle.match_window(tgt, src.desc, u, v, rad, neg, neg, ones, ones, th_dist, ratio_mode, 0.8, ori,
                                   src.angle)
    assert n == on and (asg == oasg).all()


def test_search_by_projection_wrappers(oracle, swm, frames):
    from swarmmap_b200.matcher import ORBmatcher
    last, cur = frames["euroc"][0], frames["euroc"][1]
    rng = np.random.default_rng(5)
    u = last.x + rng.normal(0, 2, last.N).astype(np.float32)
    v = last.y + rng.normal(0, 2, last.N).astype(np.float32)
    valid = np.ones(last.N, np.uint8)
    m = ORBmatcher(0.9, True)
    n, asg = m.SearchByProjectionLastFrame(cur, last, u, v, valid, 15)
    sf = cur.mvScaleFactors
    on, oasg = oracle.match_window(cur, last.desc, u, v, (np.float32(15) * sf[last.octave]).astype(np.float32),
                                   last.octave - 1, last.octave + 1, valid, valid, 100, 0, 0.9, True, last.angle)
    assert n == on and (asg == oasg).all() and n > 200
    m3 = ORBmatcher(0.8, True)
    cosv = rng.uniform(0.99, 1.0, last.N).astype(np.float32)
    n, asg = m3.SearchByProjectionMapPoints(cur, last.desc, u, v, last.octave, cosv, valid, th=1.0)
    r = np.where(cosv.astype(np.float64) > 0.998, np.float32(2.5), np.float32(4.0)).astype(np.float32) * sf[last.octave]
    on, oasg = oracle.match_window(cur, last.desc, u, v, r.astype(np.float32), last.octave - 1, last.octave, valid,
                                   valid, 100, 1, 0.8, False)
    assert n == on and (asg == oasg).all() and n > 100


def test_search_by_projection_reloc_and_loop_wrappers(oracle, swm, frames):
    from swarmmap_b200.matcher import ORBmatcher
    kf, cur = frames["kitti"][0], frames["kitti"][2]
    rng = np.random.default_rng(9)
    u = kf.x + rng.normal(0, 3, kf.N).astype(np.float32)
    v = kf.y + rng.normal(0, 3, kf.N).astype(np.float32)
    valid = (rng.random(kf.N) < 0.9).astype(np.uint8)
    pred = np.clip(kf.octave + rng.integers(-1, 2, kf.N), 0, 7).astype(np.int32)
    sf = cur.mvScaleFactors
    has_mp = (rng.random(cur.N) < 0.2).astype(np.uint8)
    ones = np.ones(kf.N, np.uint8)
    for th, orbdist in ((10, 100), (3, 64)):  # Tracking.cc:1235 / :1248
        m = ORBmatcher(0.9, True)
        n, asg = m.SearchByProjectionKeyFrame(cur, kf.desc, kf.angle, u, v, pred, valid, th, orbdist, has_mp)
        on, oasg = oracle.match_window(cur, kf.desc, u, v, (np.float32(th) * sf[pred]).astype(np.float32), pred - 1,
                                       pred + 1, valid, ones, orbdist, 0, 0.9, True, kf.angle, has_mp)
        assert n == on and (asg == oasg).all() and n > 50
    m = ORBmatcher(0.75, True)
    found = (rng.random(cur.N) < 0.3).astype(np.uint8)
    n, asg = m.SearchByProjectionSim3(cur, kf.desc, u, v, pred, valid, 10, found)
    on, oasg = oracle.match_window(cur, kf.desc, u, v, (np.float32(10) * sf[pred]).astype(np.float32), pred - 1, pred,
                                   valid, ones, 50, 0, 0.75, False, None, found)
    assert n == on and (asg == oasg).all() and n > 50


def _buckets(desc, seed=7, n_nodes=1000):
    """Synthetic vocabulary buckets (ORBvoc.bin is a missing blob): node = hash of 10 descriptor bits."""
    rng = np.random.default_rng(seed)
    bits = np.unpackbits(desc, axis=1)
    pick = rng.choice(256, 10, replace=False)
    node = (bits[:, pick].astype(np.int64) * (1 << np.arange(10))).sum(1) % n_nodes
    return node


@pytest.mark.parametrize("mode,ratio", [(0, 0.7), (0, 0.75), (1, 0.75)])
def test_search_by_bow(oracle, swm, frames, mode, ratio):
    from swarmmap_b200.matcher import ORBmatcher, FeatureVector
    rng = np.random.default_rng(11 + mode)
    m = ORBmatcher(ratio, True)
    for name in ("euroc", "kitti"):
        kf, f = frames[name][0], frames[name][1]
        # coarse buckets (few bits) so that nodes hold several features on both sides
        fv1 = FeatureVector(_buckets(kf.desc, 7, 64))
        fv2 = FeatureVector(_buckets(f.desc, 7, 64))
        v1 = (rng.random(kf.N) < 0.85).astype(np.uint8)
        v2 = (rng.random(f.N) < 0.85).astype(np.uint8) if mode == 1 else None
        n, out = m.SearchByBoW(kf, fv1, v1, f, fv2, v2)
        on, oout = oracle.search_by_bow(kf, fv1, v1, f, fv2, v2, mode, ratio, True)
        assert n == on
        np.testing.assert_array_equal(out, oout)
        assert n > 10
    # disjoint vocabularies -> no shared node, zero matches
    kf, f = frames["euroc"][0], frames["euroc"][1]
    fa = FeatureVector(_buckets(kf.desc, 7, 64))
    fb = FeatureVector(_buckets(f.desc, 7, 64) + 1000)
    n, out = m.SearchByBoW(kf, fa, np.ones(kf.N, np.uint8), f, fb, None if mode == 0 else np.ones(f.N, np.uint8))
    assert n == 0 and (out == -1).all()


def _fundamental(dx, dy, fx=458.654, fy=457.296, cx=367.215, cy=248.375):
    """F12 for a pure translation (tx, ty, tz) between two views of the same camera: x2' F12' ... built the way
    LocalMapping::ComputeF12 does (K^-T [t]x R K^-1), float32."""
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    t = np.array([dx, dy, 0.02])
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Ki = np.linalg.inv(K)
    return (Ki.T @ tx @ Ki).astype(np.float32)


@pytest.mark.parametrize("check_ori", [True, False])
@pytest.mark.parametrize("name", ["euroc", "kitti"])
def test_search_for_triangulation(oracle, swm, frames, name, check_ori):
    """SearchForTriangulation (ORBmatcher.cc:599-749): match indices and count bit-exact against the oracle, host
    arrays and resident frames, with MapPoint masks on both sides and an epipole inside the image."""
    from swarmmap_b200.matcher import FeatureVector, ORBmatcher, ResidentFrame
    f1, f2 = frames[name][0], frames[name][1]
    rng = np.random.default_rng(7)
    fv1 = FeatureVector(_buckets(f1.desc, 5, 40))
    fv2 = FeatureVector(_buckets(f2.desc, 5, 40))
    v1 = (rng.random(f1.N) < 0.7).astype(np.uint8)
    v2 = (rng.random(f2.N) < 0.8).astype(np.uint8)
    F12 = _fundamental(0.3, 0.05)
    sf, _, s2, _ = oracle.scale_tables(1.2, 8)
    ex, ey = 400.0, 200.0
    m = ORBmatcher(0.6, check_ori)
    n, out = m.SearchForTriangulation(f1, fv1, v1, f2, fv2, v2, F12, ex, ey, sf, s2)
    rn, rout = oracle.search_for_triangulation(f1, fv1, v1, f2, fv2, v2, F12, ex, ey, sf, s2, check_ori)
    assert n == rn and n > 5, (n, rn)
    np.testing.assert_array_equal(out, rout)
    r1, r2 = ResidentFrame().upload(f1), ResidentFrame().upload(f2)
    n2, out2 = m.SearchForTriangulation(r1, fv1, v1, r2, fv2, v2, F12, ex, ey, sf, s2)
    assert n2 == rn
    np.testing.assert_array_equal(out2, rout)
    # nothing to match: all KF1 keypoints already have MapPoints
    n3, out3 = m.SearchForTriangulation(f1, fv1, np.zeros(f1.N, np.uint8), f2, fv2, v2, F12, ex, ey, sf, s2)
    assert n3 == 0 and (out3 == -1).all()


@pytest.mark.parametrize("chi2", [5.99, 0.0])
@pytest.mark.parametrize("name", ["euroc", "kitti"])
def test_window_best_fuse_loop(oracle, swm, frames, name, chi2):
    """The search loop of Fuse / SearchBySim3: best index and distance per projected MapPoint, bit-exact against the
    oracle's GetFeaturesInArea-based loop (levels [pred - 1, pred], reprojection gate, first-wins ties)."""
    from swarmmap_b200.matcher import ORBmatcher, ResidentFrame
    src, tgt = frames[name][0], frames[name][1]
    rng = np.random.default_rng(13)
    sf, _, _, inv_s2 = oracle.scale_tables(1.2, 8)
    u = src.x + rng.normal(0, 1.5, src.N).astype(np.float32)
    v = src.y + rng.normal(0, 1.5, src.N).astype(np.float32)
    pred = np.clip(src.octave + rng.integers(-1, 2, src.N), 0, 7).astype(np.int32)
    radius = (np.float32(12.0) * sf[pred]).astype(np.float32)
    valid = (rng.random(src.N) < 0.9).astype(np.uint8)
    m = ORBmatcher()
    bi, bd = m.window_best(tgt, src.desc, u, v, radius, pred, valid, inv_s2, chi2)
    ri, rd = oracle.window_best(tgt, src.desc, u, v, radius, pred, valid, inv_s2, chi2)
    np.testing.assert_array_equal(bi, ri)
    np.testing.assert_array_equal(bd, rd)
    assert (bi >= 0).sum() > src.N // 20
    if chi2 > 0:
        assert (bi >= 0).sum() < (m.window_best(tgt, src.desc, u, v, radius, pred, valid)[0] >= 0).sum() + 1
    bi2, bd2 = m.window_best(ResidentFrame().upload(tgt), src.desc, u, v, radius, pred, valid, inv_s2, chi2)
    np.testing.assert_array_equal(bi2, ri)
    np.testing.assert_array_equal(bd2, rd)


def test_compute_distinctive_descriptors(oracle, swm):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:361-391), batched: best index and median per MapPoint equal
    the oracle's sort-based version (even / odd N, single observation, empty point, duplicates -> first wins)."""
    from swarmmap_b200.matcher import ORBmatcher
    rng = np.random.default_rng(11)
    sizes = [1, 2, 3, 4, 7, 0, 16, 33, 100, 257, 5, 1024] + list(rng.integers(1, 60, 200))
    chunks = []
    for n in sizes:
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        d = np.repeat(base[None], n, 0)
        if n:
            bits = np.unpackbits(d, axis=1)
            bits ^= (rng.random(bits.shape) < 0.08).astype(np.uint8)  # observations = noisy copies of one descriptor
            d = np.packbits(bits, axis=1)
            if n > 3:
                d[n // 2] = d[1]  # exact duplicate rows: equal medians, the first must win
        chunks.append(d)
    desc = np.concatenate(chunks)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    best, med = ORBmatcher().ComputeDistinctiveDescriptors(desc, offsets)
    rbest, rmed = oracle.distinctive_descriptors(desc, offsets)
    np.testing.assert_array_equal(best, rbest)
    np.testing.assert_array_equal(med, rmed)


def _db_case(rng, nq, ndb):
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    db = rng.integers(0, 256, (ndb, 32), dtype=np.uint8)
    for i in range(0, nq, 3):  # planted noisy copies + exact duplicates (tie-break on index)
        j = int(rng.integers(0, ndb))
        bits = np.unpackbits(q[i])
        bits[rng.choice(256, 20, replace=False)] ^= 1
        db[j] = np.packbits(bits)
    if ndb > 800:
        db[777] = db[123]
        q[5] = db[123]
    # extreme popcounts on both sides: distance 0 and 256, all-zero / all-one rows
    q[1] = 0
    q[2] = 255
    db[ndb - 1] = 0
    db[ndb // 2] = 255
    return q, db


@pytest.mark.parametrize("kernel", ["umma", "imma", "popc"])
@pytest.mark.parametrize("nq,ndb", [(300, 20000), (256, 128 * 40), (17, 77), (513, 2), (40, 70001)])
def test_db_top2_shard(oracle, swm, monkeypatch, kernel, nq, ndb):
    """Shard scan == oracle brute force (distance, index, tie-break, votes) for each of the three kernels:
    tcgen05 int8 (default), legacy mma.sync, CUDA-core POPC."""
    import ctypes as C
    import torch
    from swarmmap_b200 import _lib
    lib = _lib.load()
    monkeypatch.setenv("SWM_DB_KERNEL", kernel)
    q, db = _db_case(np.random.default_rng(99 + nq), nq, ndb)
    ref = oracle.bruteforce_top2(q, db)
    h = C.c_void_p()
    first_kf, per_kf = 1000, 8
    n_kf = (ndb + per_kf - 1) // per_kf
    assert lib.swm_db_create(0, _lib.ptr(db), ndb, per_kf, first_kf, C.byref(h)) == 0
    dq = torch.from_numpy(q).cuda()
    topk = torch.zeros((nq, 2), dtype=torch.int64, device="cuda")
    votes = torch.zeros(n_kf, dtype=torch.int32, device="cuda")
    rc = lib.swm_db_query_device(h, dq.data_ptr(), nq, 2, topk.data_ptr(), votes.data_ptr(), 50, None)
    assert rc == 0
    torch.cuda.synchronize()
    t = topk.cpu().numpy().astype(np.uint64)
    dist = (t >> np.uint64(48)).astype(np.int64)
    idx = (t & np.uint64((1 << 48) - 1)).astype(np.int64) - first_kf * per_kf
    np.testing.assert_array_equal(dist[:, 0], ref[:, 0])
    np.testing.assert_array_equal(idx[:, 0], ref[:, 1])
    np.testing.assert_array_equal(dist[:, 1], ref[:, 2])
    np.testing.assert_array_equal(idx[:, 1], ref[:, 3])
    exp_votes = np.bincount(ref[ref[:, 0] <= 50, 1] // per_kf, minlength=n_kf)
    np.testing.assert_array_equal(votes.cpu().numpy(), exp_votes)
    assert lib.swm_db_size(h) == ndb
    lib.swm_db_destroy(h)


def test_db_top2_config5_scale(oracle, swm):
    """BASELINE config 5 shape at a size the oracle still finishes in seconds: 2000 queries x 4096 keyframes x 256
    descriptors (1 048 576), 1 % of the keyframes hold noisy copies of query descriptors.  The whole query batch runs on
    the tcgen05 shard scan; 32 sampled queries (every planted one among them) are checked against the oracle's brute
    force -- distance and index of best and second best -- and the vote histogram lands on the planted keyframes only."""
    import ctypes as C
    import torch
    from swarmmap_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(555)
    nq, n_kf, per_kf, first_kf = 2000, 4096, 256, 70000
    ndb = n_kf * per_kf
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    db = rng.integers(0, 256, (ndb, 32), dtype=np.uint8)
    planted_kf = rng.choice(n_kf, n_kf // 100, replace=False)
    planted_q = rng.choice(nq, len(planted_kf), replace=False)
    for kf, qi in zip(planted_kf, planted_q):
        bits = np.unpackbits(q[qi])
        bits[rng.choice(256, 12, replace=False)] ^= 1
        db[kf * per_kf + int(rng.integers(0, per_kf))] = np.packbits(bits)
    h = C.c_void_p()
    assert lib.swm_db_create(0, _lib.ptr(db), ndb, per_kf, first_kf, C.byref(h)) == 0
    dq = torch.from_numpy(q).cuda()
    topk = torch.zeros((nq, 2), dtype=torch.int64, device="cuda")
    votes = torch.zeros(n_kf, dtype=torch.int32, device="cuda")
    assert lib.swm_db_query_device(h, dq.data_ptr(), nq, 2, topk.data_ptr(), votes.data_ptr(), 50, None) == 0
    torch.cuda.synchronize()
    t = topk.cpu().numpy().astype(np.uint64)
    dist = (t >> np.uint64(48)).astype(np.int64)
    idx = (t & np.uint64((1 << 48) - 1)).astype(np.int64) - first_kf * per_kf
    sample = np.unique(np.concatenate([planted_q[:16], rng.choice(nq, 16, replace=False)]))
    ref = oracle.bruteforce_top2(q[sample], db)
    np.testing.assert_array_equal(dist[sample, 0], ref[:, 0])
    np.testing.assert_array_equal(idx[sample, 0], ref[:, 1])
    np.testing.assert_array_equal(dist[sample, 1], ref[:, 2])
    np.testing.assert_array_equal(idx[sample, 1], ref[:, 3])
    assert (dist[planted_q, 0] == 12).all() and (dist[np.setdiff1d(np.arange(nq), planted_q), 0] > 50).all()
    v = votes.cpu().numpy()
    exp = np.zeros(n_kf, np.int64)
    exp[planted_kf] = 1
    np.testing.assert_array_equal(v, exp)
    lib.swm_db_destroy(h)


@pytest.mark.parametrize("k", [1, 2])
def test_db_merge_gathered_equals_single_shard(oracle, swm, k):
    """The multi-GPU exchange step on one device: two shards scanned separately, their (nq, k) key blocks stacked the
    way all_gather_into_tensor lays them out, merged by swm_db_merge_gathered on each 'rank' -> the keys of the
    single-shard scan, and the two ranks' votes sum to the single-shard vote histogram."""
    import ctypes as C
    import torch
    from swarmmap_b200 import _lib, place
    lib = _lib.load()
    rng = np.random.default_rng(21)
    q, db = _db_case(rng, 333, 24000)
    per_kf = 8
    full = place.PlaceShard(db, per_kf, 0)
    fk, fv = full.query_local(torch.from_numpy(q), k, 50)
    cut = 11000 // per_kf * per_kf
    shards = [place.PlaceShard(db[:cut], per_kf, 0), place.PlaceShard(db[cut:], per_kf, cut // per_kf)]
    local = [s.query_local(torch.from_numpy(q), k, 50)[0] for s in shards]
    gathered = torch.stack(local, 0).contiguous()
    votes_sum = torch.zeros(len(db) // per_kf, dtype=torch.int32, device="cuda")
    for r, s in enumerate(shards):
        merged = torch.empty_like(local[0])
        votes = torch.zeros(s.n_kf, dtype=torch.int32, device="cuda")
        rc = lib.swm_db_merge_gathered(s._h, gathered.data_ptr(), 2, len(q), k, merged.data_ptr(), votes.data_ptr(), 50, None)
        assert rc == 0
        torch.cuda.synchronize()
        assert torch.equal(merged, fk)
        votes_sum[s.first_kf:s.first_kf + s.n_kf] += votes
    assert torch.equal(votes_sum, fv)


@pytest.mark.parametrize("world,k", [(1, 2), (2, 2), (3, 1), (4, 2)])
def test_db_query_peers_equals_single_shard(oracle, swm, world, k):
    """The peer-memory exchange (db_merge_peers_kernel: push into every rank's window, flags, wait, merge) with `world`
    shards driven from ONE process on one GPU, each on its own stream so that the kernels really wait for each other:
    every rank gets the keys of the single-shard scan, the votes sum to the single-shard histogram, and a second and
    third query reuse the windows (sequence parity)."""
    import os
    import torch
    from swarmmap_b200 import place
    if os.environ.get("CUDA_LAUNCH_BLOCKING") == "1":
        pytest.skip("the ranks' kernels wait for each other: they cannot run with serialised launches")
    rng = np.random.default_rng(31 + world)
    per_kf = 8
    q, db = _db_case(rng, 500, 32000)
    full = place.PlaceShard(db, per_kf, 0)
    n_kf = len(db) // per_kf
    cuts = [0] + sorted(int(c) // per_kf * per_kf for c in rng.choice(np.arange(4000, 28000), world - 1, replace=False)) + [len(db)]
    shards = [place.PlaceShard(db[cuts[r]:cuts[r + 1]], per_kf, cuts[r] // per_kf) for r in range(world)]
    for sh in shards:
        sh.enable_peers(nq_max=512, same_process=shards)
    streams = [torch.cuda.Stream() for _ in range(world)]
    for rep, nq in enumerate((500, 77, 500)):
        qq = torch.from_numpy(q[:nq])
        fk, fv = full.query_local(qq, k, 50)
        outs = []
        for r in (list(range(world)) if rep != 1 else list(reversed(range(world)))):  # any enqueue order
            with torch.cuda.stream(streams[r]):
                outs.append((r, shards[r].query_peers(qq, k, 50)))
        torch.cuda.synchronize()
        votes_sum = torch.zeros(n_kf, dtype=torch.int32, device="cuda")
        for r, (keys, votes) in outs:
            assert torch.equal(keys, fk), f"rank {r} rep {rep}"
            votes_sum[shards[r].first_kf:shards[r].first_kf + shards[r].n_kf] += votes
        assert torch.equal(votes_sum, fv)
