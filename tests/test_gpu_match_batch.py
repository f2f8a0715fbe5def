"""GPU parity of the BATCHED matchers (swm_match_window_batch / _init_batch / _bow_batch, swm_frames_from_extractor)
against the CPU oracle: every job of a batch must give exactly the indices and counts of the oracle's single call on
the same inputs -- ragged batches (different sizes, modes, ratio settings per job), host-array and resident frames,
more jobs than SMs, and a candidate-buffer regrowth in the middle of a run."""
import numpy as np
import pytest

from swarmmap_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frames(oracle):
    from swarmmap_b200.matcher import Frame
    out = {}
    for name, (w, h, nf, seed, cnt) in {"kitti": (1241, 376, 4000, 20220405, 3), "euroc": (752, 480, 1000, 20220406, 6),
                                        "small": (320, 240, 300, 5, 4)}.items():
        seq = synth.make_sequence(cnt, w, h, seed)
        ex = oracle.Extractor(nf, 1.2, 8, 20, 7)
        sf = oracle.scale_tables(1.2, 8)[0]
        out[name] = [Frame.from_keypoints(*ex(img), w, h, sf) for img in seq]
    return out


def _window_job(rng, src, tgt, th, band, jitter, th_dist, ratio_mode, ori):
    sf = tgt.mvScaleFactors
    u = src.x + rng.normal(0, jitter, src.N).astype(np.float32)
    v = src.y + rng.normal(0, jitter, src.N).astype(np.float32)
    radius = (np.float32(th) * sf[src.octave]).astype(np.float32)
    asg0 = np.full(tgt.N, -1, np.int32)
    asg0[rng.random(tgt.N) < 0.05] = 7
    return dict(tgt=tgt, desc=src.desc, u=u, v=v, radius=radius, min_level=src.octave + band[0],
                max_level=src.octave + band[1], valid=(rng.random(src.N) < 0.8).astype(np.uint8),
                blocks=(rng.random(src.N) < 0.7).astype(np.uint8), th_dist=th_dist, ratio_mode=ratio_mode,
                angle=src.angle, tgt_blocked=(rng.random(tgt.N) < 0.1).astype(np.uint8), assignment=asg0,
                check_ori=ori)


def _check_window(oracle, job, got, nnratio):
    n, asg = got
    a0 = job["assignment_in"]
    on, oasg = oracle.match_window(job["tgt_host"], job["desc"], job["u"], job["v"], job["radius"], job["min_level"],
                                   job["max_level"], job["valid"], job["blocks"], job["th_dist"], job["ratio_mode"],
                                   nnratio, job["check_ori"], job["angle"], job["tgt_blocked"], a0)
    assert n == on
    np.testing.assert_array_equal(asg, oasg)
    return n


def test_window_batch_ragged(oracle, swm, frames):
    from swarmmap_b200.matcher import ORBmatcher, ResidentFrame
    rng = np.random.default_rng(11)
    m = ORBmatcher(0.8, True)
    jobs = []
    specs = [("euroc", 0, 1, 15, (-1, 1), 3.0, 100, 0, True), ("kitti", 0, 1, 15, (-1, 1), 3.0, 100, 0, True),
             ("euroc", 1, 2, 4, (-1, 0), 2.0, 100, 1, False), ("small", 0, 1, 10, (-1, 1), 2.0, 64, 0, True),
             ("euroc", 2, 3, 30, (-1, 1), 6.0, 50, 0, False), ("small", 2, 3, 4, (-1, 0), 1.0, 100, 1, False),
             ("kitti", 1, 2, 4, (-1, 0), 2.0, 100, 1, False)]
    resident = {}
    for k, (name, a, b, th, band, jit, thd, rm, ori) in enumerate(specs):
        src, tgt = frames[name][a], frames[name][b]
        job = _window_job(rng, src, tgt, th, band, jit, thd, rm, ori)
        job["tgt_host"] = tgt
        job["assignment_in"] = job["assignment"].copy()
        if k % 2 == 1:  # every other job matches against a device-resident target
            key = (name, b)
            if key not in resident:
                resident[key] = ResidentFrame().upload(tgt)
            job["tgt"] = resident[key]
        jobs.append(job)
    call = [{k: v for k, v in j.items() if k not in ("tgt_host", "assignment_in")} for j in jobs]
    got = m.match_window_batch(call)
    total = sum(_check_window(oracle, j, g, 0.8) for j, g in zip(jobs, got))
    assert total > 500
    # the same batch again (buffers are reused) and in reverse order
    for j in jobs:
        j["assignment"] = j["assignment_in"].copy()
    call = [{k: v for k, v in j.items() if k not in ("tgt_host", "assignment_in")} for j in reversed(jobs)]
    got = m.match_window_batch(call)
    for j, g in zip(reversed(jobs), got):
        _check_window(oracle, j, g, 0.8)


def test_window_batch_more_jobs_than_sms_and_regrowth(oracle, swm, frames):
    """300 jobs (two resolve waves on 148 SMs); the second call uses far wider windows than the first, so the
    candidate buffer sized by the first call overflows and the call is repeated internally."""
    from swarmmap_b200.matcher import ORBmatcher
    rng = np.random.default_rng(12)
    m = ORBmatcher(0.9, True)
    fs = frames["euroc"]
    for th in (6, 400):
        jobs = []
        for k in range(300 if th == 6 else 40):
            src, tgt = fs[k % 5], fs[(k + 1) % 5]
            j = _window_job(rng, src, tgt, th, (-1, 1), 2.0, 100, 0, True)
            j["tgt_host"] = tgt
            j["assignment_in"] = j["assignment"].copy()
            jobs.append(j)
        got = m.match_window_batch([{k: v for k, v in j.items() if k not in ("tgt_host", "assignment_in")} for j in jobs])
        for j, g in zip(jobs[::7], got[::7]):
            _check_window(oracle, j, g, 0.9)
        # every job against the single-call entry point
        for j, g in list(zip(jobs, got))[:12]:
            n1, a1 = m.match_window(j["tgt_host"], j["desc"], j["u"], j["v"], j["radius"], j["min_level"],
                                    j["max_level"], j["valid"], j["blocks"], j["th_dist"], j["ratio_mode"], j["angle"],
                                    j["tgt_blocked"], j["assignment_in"].copy(), check_ori=j["check_ori"])
            assert n1 == g[0] and np.array_equal(a1, g[1])


def test_init_batch(oracle, swm, frames):
    from swarmmap_b200.matcher import ORBmatcher, ResidentFrame
    m = ORBmatcher(0.9, True)
    pairs, exp = [], []
    for name, window in (("kitti", 100), ("euroc", 100), ("euroc", 30), ("small", 50)):
        fs = frames[name]
        for k in (1, 2):
            prev = np.stack([fs[0].x, fs[0].y], 1).astype(np.float32).copy()
            pairs.append((fs[0], fs[k], prev, window))
            exp.append(oracle.search_for_initialization(fs[0], fs[k], prev.copy(), window, 0.9, True))
    for window in (100, 30, 50):
        sel = [i for i, p in enumerate(pairs) if p[3] == window]
        got = m.SearchForInitializationBatch([pairs[i][:3] for i in sel], window)
        for i, (n, m12) in zip(sel, got):
            on, om12, oprev = exp[i]
            assert n == on and n > 20
            np.testing.assert_array_equal(m12, om12)
            np.testing.assert_array_equal(pairs[i][2], oprev)
    # resident frames, prev carried over three calls as Tracking::MonocularInitialization does
    fs = frames["euroc"]
    res = [ResidentFrame().upload(f) for f in fs[:4]]
    prev_g = np.stack([fs[0].x, fs[0].y], 1).astype(np.float32).copy()
    prev_o = prev_g.copy()
    for k in (1, 2, 3):
        (n, m12), = m.SearchForInitializationBatch([(res[0], res[k], prev_g)], 100)
        on, om12, prev_o = oracle.search_for_initialization(fs[0], fs[k], prev_o, 100, 0.9, True)
        assert n == on
        np.testing.assert_array_equal(m12, om12)
        np.testing.assert_array_equal(prev_g, prev_o)


def _nodes(rng, f, n_nodes, drop=0.03):
    from swarmmap_b200.matcher import FeatureVector
    node = (f.desc[:, 3].astype(np.int64) >> 2) % n_nodes + 1000 * (f.desc[:, 9] & 1).astype(np.int64)
    node[rng.random(f.N) < drop] = -1
    return FeatureVector(node)


def test_bow_batch(oracle, swm, frames):
    from swarmmap_b200.matcher import ORBmatcher, ResidentFrame
    rng = np.random.default_rng(13)
    for ratio, ori in ((0.7, True), (0.75, False)):
        m = ORBmatcher(ratio, ori)
        jobs, exp = [], []
        for name, a, b, nn in (("euroc", 0, 1, 40), ("kitti", 0, 1, 64), ("small", 1, 2, 8), ("euroc", 2, 3, 3),
                               ("euroc", 3, 4, 64), ("small", 0, 3, 64)):
            kf, F = frames[name][a], frames[name][b]
            fv1, fv2 = _nodes(rng, kf, nn), _nodes(rng, F, nn)
            v1 = (rng.random(kf.N) < 0.8).astype(np.uint8)
            for mode in (0, 1):
                v2 = (rng.random(F.N) < 0.8).astype(np.uint8) if mode else None
                jobs.append((kf, fv1, v1, F, fv2, v2))
                exp.append(oracle.search_by_bow(kf, fv1, v1, F, fv2, v2, mode, ratio, ori))
        got = m.SearchByBoWBatch(jobs)
        total = 0
        for (n, out), (on, oout) in zip(got, exp):
            assert n == on
            np.testing.assert_array_equal(out, oout)
            total += n
        assert total > 300
        # resident operands + a job without any shared node
        kf, F = frames["euroc"][0], frames["euroc"][1]
        rk, rF = ResidentFrame().upload(kf), ResidentFrame().upload(F)
        fv1, fv2 = _nodes(rng, kf, 40), _nodes(rng, F, 40)
        from swarmmap_b200.matcher import FeatureVector
        lonely = FeatureVector(np.full(F.N, 99999))
        v1 = np.ones(kf.N, np.uint8)
        got = m.SearchByBoWBatch([(rk, fv1, v1, rF, fv2, None), (rk, fv1, v1, rF, lonely, None)])
        on, oout = oracle.search_by_bow(kf, fv1, v1, F, fv2, None, 0, ratio, ori)
        assert got[0][0] == on and np.array_equal(got[0][1], oout)
        assert got[1][0] == 0 and (got[1][1] == -1).all()


def test_frames_from_extractor_batch_equals_single(oracle, swm):
    from swarmmap_b200.matcher import Camera, ResidentFrame, resident_frames_from_extractor
    from swarmmap_b200.orb import ORBextractor
    imgs = synth.make_batch(5, 752, 480, 20220410)
    gpu = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=5)
    gpu.extract_batch(imgs)
    cam = Camera(458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)
    bounds = cam.bounds(752, 480)
    batch = resident_frames_from_extractor(gpu, 5, cam, bounds)
    for b in range(5):
        one = ResidentFrame().from_extractor(gpu, b, cam, bounds)
        g, s = batch[b].download(grid=True), one.download(grid=True)
        assert batch[b].N == one.N and one.N > 500
        for k in ("x", "y", "octave", "angle", "desc", "starts", "items"):
            np.testing.assert_array_equal(g[k], s[k], err_msg=k)
