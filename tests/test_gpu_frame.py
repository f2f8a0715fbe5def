"""Resident frames: extractor output -> undistort -> grid -> matchers without a host
round trip.  Parity: undistorted coordinates bit-exact against the oracle (itself bit-exact against cv2), grid CSR
identical, and every matcher returns exactly what it returns for the same frames passed as host arrays."""
import numpy as np
import pytest

from swarmmap_b200 import synth

pytestmark = pytest.mark.gpu

CAMERAS = {
    "euroc": [458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0],
    "tum1": [517.306408, 516.469215, 318.643040, 255.313989, 0.262383, -0.953104, -0.005358, 0.002628, 1.163314],
    "none": [458.654, 457.296, 367.215, 248.375, 0.0, 0.0, 0.0, 0.0, 0.0],
}


@pytest.fixture(scope="module")
def extracted(swm):
    """Four 752x480 frames extracted on the GPU; the batch stays resident in the extractor."""
    from swarmmap_b200.orb import ORBextractor
    seq = synth.make_sequence(4, 752, 480, 77)
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=4)
    kps, desc, n = ex.extract_batch(np.stack(seq))
    return ex, kps, desc, n


@pytest.mark.parametrize("name", sorted(CAMERAS))
def test_frame_from_extractor(oracle, swm, extracted, name):
    from swarmmap_b200.matcher import Camera, ResidentFrame
    ex, kps, desc, n = extracted
    cam9 = np.array(CAMERAS[name], np.float32)
    cam = Camera(*[float(c) for c in cam9])
    bounds = cam.bounds(752, 480)
    np.testing.assert_array_equal(bounds, oracle.image_bounds(752, 480, cam9))
    for b in range(4):
        f = ResidentFrame().from_extractor(ex, b, cam, bounds)
        nb = int(n[b])
        assert f.N == nb
        got = f.download(grid=True)
        k = kps[b, :nb]
        xy = np.stack([k["x"], k["y"]], 1).astype(np.float32)
        ref = oracle.undistort_points(xy, cam9)
        np.testing.assert_array_equal(got["x"], ref[:, 0])
        np.testing.assert_array_equal(got["y"], ref[:, 1])
        np.testing.assert_array_equal(got["octave"], k["octave"])
        np.testing.assert_array_equal(got["angle"], k["angle"])
        np.testing.assert_array_equal(got["desc"], desc[b, :nb])
        from swarmmap_b200.matcher import Frame
        host = Frame(ref[:, 0], ref[:, 1], k["octave"], k["angle"], desc[b, :nb],
                     (bounds[0], bounds[2], bounds[1], bounds[3]))
        st, it = oracle.grid_csr(host)
        np.testing.assert_array_equal(got["starts"], st)
        np.testing.assert_array_equal(got["items"], it)
        f.close()


def test_camera_none_is_identity(swm, extracted):
    from swarmmap_b200.matcher import ResidentFrame
    ex, kps, desc, n = extracted
    f = ResidentFrame().from_extractor(ex, 1, None, np.array([0, 752, 0, 480], np.float32))
    got = f.download()
    np.testing.assert_array_equal(got["x"], kps[1, :n[1]]["x"])
    np.testing.assert_array_equal(got["y"], kps[1, :n[1]]["y"])


def _host_frames(ex, kps, desc, n, cam9, bounds, oracle):
    from swarmmap_b200.matcher import Frame
    out = []
    for b in range(len(n)):
        k = kps[b, :n[b]]
        ref = oracle.undistort_points(np.stack([k["x"], k["y"]], 1).astype(np.float32), cam9)
        out.append(Frame(ref[:, 0], ref[:, 1], k["octave"], k["angle"], desc[b, :n[b]],
                         (bounds[0], bounds[2], bounds[1], bounds[3]), ex.GetScaleFactors()))
    return out


def test_resident_matchers_equal_host_matchers(oracle, swm, extracted):
    """extract -> ResidentFrame -> SearchForInitialization / SearchByProjection / SearchByBoW give the results of
    the host-array forms (which test_gpu_match.py pins to the oracle)."""
    from swarmmap_b200.matcher import Camera, FeatureVector, ORBmatcher, ResidentFrame
    ex, kps, desc, n = extracted
    cam9 = np.array(CAMERAS["euroc"], np.float32)
    cam = Camera(*[float(c) for c in cam9])
    bounds = cam.bounds(752, 480)
    host = _host_frames(ex, kps, desc, n, cam9, bounds, oracle)
    res = [ResidentFrame().from_extractor(ex, b, cam, bounds) for b in range(4)]
    m = ORBmatcher(0.9, True)
    # SearchForInitialization
    prev_h = np.stack([host[0].x, host[0].y], 1).astype(np.float32).copy()
    prev_r = prev_h.copy()
    nh, mh = m.SearchForInitialization(host[0], host[1], prev_h, 100)
    nr, mr = m.SearchForInitialization(res[0], res[1], prev_r, 100)
    assert nh == nr and nh > 50
    np.testing.assert_array_equal(mh, mr)
    np.testing.assert_array_equal(prev_h, prev_r)
    ref_n, ref_m, _ = oracle.search_for_initialization(host[0], host[1], np.stack([host[0].x, host[0].y], 1), 100, 0.9, True)
    assert ref_n == nr
    np.testing.assert_array_equal(ref_m, mr)
    # SearchByProjection (last frame -> current frame), projections = last positions + drift
    last, cur_h, cur_r = host[1], host[2], res[2]
    u = last.x + 3.0
    v = last.y - 2.0
    valid = np.ones(last.N, np.uint8)
    m2 = ORBmatcher(0.9, True)
    a = m2.SearchByProjectionLastFrame(cur_h, last, u, v, valid, 15)
    b = m2.SearchByProjectionLastFrame(cur_r, last, u, v, valid, 15)
    assert a[0] == b[0] and a[0] > 20
    np.testing.assert_array_equal(a[1], b[1])
    # SearchByBoW KF -> Frame with synthetic vocabulary buckets
    rng = np.random.default_rng(3)
    fa = FeatureVector(host[2].desc[:, 0].astype(np.int64) % 64)
    fb = FeatureVector(host[3].desc[:, 0].astype(np.int64) % 64)
    va = (rng.random(host[2].N) < 0.8).astype(np.uint8)
    m3 = ORBmatcher(0.7, True)
    c = m3.SearchByBoW(host[2], fa, va, host[3], fb)
    d = m3.SearchByBoW(res[2], fa, va, res[3], fb)
    assert c[0] == d[0]
    np.testing.assert_array_equal(c[1], d[1])
    for f in res:
        f.close()


def test_frame_upload_roundtrip(oracle, swm, extracted):
    from swarmmap_b200.matcher import ResidentFrame
    ex, kps, desc, n = extracted
    host = _host_frames(ex, kps, desc, n, np.array(CAMERAS["none"], np.float32), [0, 752, 0, 480], oracle)
    f = ResidentFrame().upload(host[0])
    got = f.download(grid=True)
    np.testing.assert_array_equal(got["x"], host[0].x)
    np.testing.assert_array_equal(got["desc"], host[0].desc)
    st, it = oracle.grid_csr(host[0])
    np.testing.assert_array_equal(got["starts"], st)
    np.testing.assert_array_equal(got["items"], it)
    empty = ResidentFrame()
    assert empty.N == 0


def test_frame_errors(swm, extracted):
    from swarmmap_b200.matcher import ResidentFrame
    from swarmmap_b200._lib import SwmError
    ex, kps, desc, n = extracted
    f = ResidentFrame()
    with pytest.raises(SwmError):
        f.from_extractor(ex, 99, None, np.array([0, 752, 0, 480], np.float32))
    with pytest.raises(SwmError):
        f.from_extractor(ex, 0, None, np.array([0, 0, 0, 480], np.float32))


def test_keyframe_slab_roundtrip(oracle, swm, extracted):
    """Binary keyframe-feature slab: export -> bytes -> import on another frame gives the
    same arrays and grid, 32 + 48 n bytes; corrupted headers are refused."""
    from swarmmap_b200._lib import SwmError
    from swarmmap_b200.matcher import Camera, ResidentFrame
    ex, kps, desc, n = extracted
    cam = Camera(*[float(c) for c in CAMERAS["euroc"]])
    bounds = cam.bounds(752, 480)
    a = ResidentFrame().from_extractor(ex, 2, cam, bounds)
    blob = a.export_slab()
    assert len(blob) == 32 + 48 * a.N and blob[:4] == b"SWKF"
    b = ResidentFrame().import_slab(blob)
    da, db = a.download(grid=True), b.download(grid=True)
    for k in da:
        np.testing.assert_array_equal(da[k], db[k])
    from swarmmap_b200.matcher import Frame
    z = np.zeros(0, np.float32)
    empty = ResidentFrame().upload(Frame(z, z, np.zeros(0, np.int32), z, np.zeros((0, 32), np.uint8), (0, 0, 752, 480)))
    assert ResidentFrame().import_slab(empty.export_slab()).N == 0
    with pytest.raises(SwmError):
        ResidentFrame().export_slab()  # never built
    with pytest.raises(SwmError):
        ResidentFrame().import_slab(b"XXXX" + blob[4:])
    with pytest.raises(SwmError):
        ResidentFrame().import_slab(blob[:100])
