"""Pins the oracle's restatement of the three OpenCV primitives the reference delegates to, plus the
FAST-9/16 score, against cv2 (4.13 in the image).  The reference ships no tests for this path
(SURVEY.md F4), so cv2 is the executable arbiter of cv::resize / copyMakeBorder / GaussianBlur."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _images(rng, n, lo=40, hi=700):
    for t in range(n):
        w, h = int(rng.integers(lo, hi)), int(rng.integers(lo, hi))
        kind = t % 3
        if kind == 0:
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        elif kind == 1:
            img = cv2.GaussianBlur(rng.integers(0, 256, (h, w), dtype=np.uint8), (0, 0), 2.5)
        else:
            img = (np.add.outer(np.arange(h) * 3, np.arange(w) * 5) % 256).astype(np.uint8)
        yield img


def test_resize_linear_bit_exact(oracle):
    rng = np.random.default_rng(1)
    for i, img in enumerate(_images(rng, 60)):
        h, w = img.shape
        if i % 2 == 0:  # the pyramid's own ratio
            dw, dh = int(np.rint(np.float32(w) * np.float32(1 / 1.2))), int(np.rint(np.float32(h) * np.float32(1 / 1.2)))
        else:
            dw, dh = int(rng.integers(20, w + 1)), int(rng.integers(20, h + 1))
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        np.testing.assert_array_equal(oracle.resize_linear(img, dw, dh), ref, err_msg=f"{w}x{h}->{dw}x{dh}")


def test_pyramid_chain_bit_exact(oracle):
    from swarmmap_b200 import synth
    for (w, h, seed) in ((752, 480, 20220404), (1241, 376, 20220405)):
        img = synth.make_frame(w, h, seed)
        ws, hs = oracle.level_sizes(w, h)
        ex = oracle.Extractor(1000)
        ex(img)
        cur = img
        for l in range(8):
            if l:
                cur = cv2.resize(cur, (int(ws[l]), int(hs[l])), interpolation=cv2.INTER_LINEAR)
            np.testing.assert_array_equal(ex.level(l, 0), cv2.copyMakeBorder(cur, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))
            np.testing.assert_array_equal(ex.level(l, 1), cv2.GaussianBlur(cur, (7, 7), 2, sigmaY=2,
                                                                           borderType=cv2.BORDER_REFLECT_101))


def test_gaussian_blur_bit_exact(oracle):
    rng = np.random.default_rng(2)
    for img in _images(rng, 40, 8, 500):
        ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        np.testing.assert_array_equal(oracle.gauss7(img), ref)


def test_border_reflect101(oracle):
    rng = np.random.default_rng(3)
    for img in _images(rng, 12, 20, 300):
        np.testing.assert_array_equal(oracle.border_reflect101(img, 19),
                                      cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))


@pytest.mark.parametrize("th", [7, 20])
def test_fast_score_matches_cv2(oracle, th):
    rng = np.random.default_rng(4 + th)
    det_all = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=False,
                                             type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det_nms = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                             type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    for img in _images(rng, 12, 40, 300):
        s = oracle.fast_score_map(img, th)
        corners = np.zeros(img.shape, bool)
        for k in det_all.detect(img):
            corners[int(k.pt[1]), int(k.pt[0])] = True
        np.testing.assert_array_equal(s > 0, corners)
        for k in det_nms.detect(img):  # with NMS cv2 reports response = the corner score
            assert s[int(k.pt[1]), int(k.pt[0])] == int(k.response)


def test_tile_select_properties(oracle):
    """Deterministic lock-step tile semantics on a score map."""
    from swarmmap_b200 import synth
    img = synth.make_frame(400, 300, 9)
    roi = img[16:-16, 16:-16]
    s = oracle.fast_score_map(np.ascontiguousarray(roi), 7)
    pts, retry = oracle.fast_tile_select(s, 20, want_retry=True)
    assert len(pts) > 100
    keys = pts["y"].astype(np.int64) * 4096 + pts["x"]
    assert (np.diff(keys) > 0).all(), "raster order"
    sc = s[pts["y"], pts["x"]]
    np.testing.assert_array_equal(sc, pts["score"])
    ty, tx = (pts["y"] - 3) // 32, (pts["x"] - 3) // 32
    lowtile = retry[ty, tx] == 1
    assert (sc[~lowtile] >= 20).all(), "non-retry tiles only emit iniThFAST corners"
    assert (sc >= 7).all()
    # strict 3x3 maxima can never be adjacent
    occ = np.zeros(s.shape, bool)
    occ[pts["y"], pts["x"]] = True
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                assert not (occ & np.roll(np.roll(occ, dy, 0), dx, 1)).any()


# EuRoC cam0, TUM1 (5 coefficients), KITTI-like (no distortion)
CAMERAS = {
    "euroc": [458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0],
    "tum1": [517.306408, 516.469215, 318.643040, 255.313989, 0.262383, -0.953104, -0.005358, 0.002628, 1.163314],
    "kitti": [718.856, 718.856, 607.1928, 185.2157, 0.0, 0.0, 0.0, 0.0, 0.0],
}


@pytest.mark.parametrize("name", sorted(CAMERAS))
def test_undistort_points_bit_exact(oracle, name):
    """Frame::UndistortKeyPoints (Frame.cc:454-484) = cv::undistortPoints(K, D, P = K): the oracle's restatement
    of the 5-iteration scheme equals cv2 bit for bit, and so do the image bounds built from it (:486-514)."""
    cam = np.array(CAMERAS[name], np.float32)
    w, h = 752, 480
    rng = np.random.default_rng(5)
    xy = np.stack([rng.uniform(0, w, 6000), rng.uniform(0, h, 6000)], 1).astype(np.float32)
    xy[:2000] = np.round(xy[:2000])  # level-0 keypoints are integers
    K = np.array([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], np.float32)
    got = oracle.undistort_points(xy, cam)
    if cam[4] == 0:
        np.testing.assert_array_equal(got, xy)  # mvKeysUn = mvKeys (:456-460)
        np.testing.assert_array_equal(oracle.image_bounds(w, h, cam), [0, w, 0, h])
        return
    ref = cv2.undistortPoints(xy.reshape(-1, 1, 2), K, cam[4:9], None, K).reshape(-1, 2)
    np.testing.assert_array_equal(got, ref)
    corners = np.array([[0, 0], [w, 0], [0, h], [w, h]], np.float32)
    m = cv2.undistortPoints(corners.reshape(-1, 1, 2), K, cam[4:9], None, K).reshape(-1, 2)
    exp = [min(m[0, 0], m[2, 0]), max(m[1, 0], m[3, 0]), min(m[0, 1], m[1, 1]), max(m[2, 1], m[3, 1])]
    np.testing.assert_array_equal(oracle.image_bounds(w, h, cam), np.array(exp, np.float32))
