"""ctypes binding of oracle/_ref/liborbmatcher_ref.so: the REFERENCE's own ORBmatcher.cc (compiled unmodified by
`make -C oracle ref`, see oracle/ref_orbmatcher_wrap.cpp) behind flat arrays.  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "liborbmatcher_ref.so")

EUROC_K = (458.654, 457.296, 367.215, 248.375)


class RFrame(C.Structure):
    _fields_ = [("n", C.c_int32), ("x", C.c_void_p), ("y", C.c_void_p), ("octave", C.c_void_p), ("angle", C.c_void_p),
                ("desc", C.c_void_p), ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float),
                ("max_y", C.c_float), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("nlevels", C.c_int32), ("log_scale_factor", C.c_float), ("scale_factors", C.c_void_p),
                ("level_sigma2", C.c_void_p), ("inv_level_sigma2", C.c_void_p), ("Tcw", C.c_void_p),
                ("mp_index", C.c_void_p), ("outlier", C.c_void_p), ("n_nodes", C.c_int32), ("node_ids", C.c_void_p),
                ("node_off", C.c_void_p), ("node_feats", C.c_void_p)]


class RPoints(C.Structure):
    _fields_ = [("n", C.c_int32), ("pos", C.c_void_p), ("normal", C.c_void_p), ("desc", C.c_void_p),
                ("nobs", C.c_void_p), ("bad", C.c_void_p), ("min_dist", C.c_void_p), ("max_dist", C.c_void_p),
                ("track_in_view", C.c_void_p), ("proj_x", C.c_void_p), ("proj_y", C.c_void_p),
                ("view_cos", C.c_void_p), ("track_level", C.c_void_p)]


_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p).value


def _arr(a, dt):
    return None if a is None else np.ascontiguousarray(a, dt)


class Scales:
    """The scale tables a Frame copies from its extractor (Frame.cc:111-117), float32 chain as ORBextractor.cc:346-362."""

    def __init__(self, scale_factor=1.2, nlevels=8):
        sf = np.ones(nlevels, np.float32)
        for i in range(1, nlevels):
            sf[i] = np.float32(sf[i - 1] * np.float32(scale_factor))
        self.sf = sf
        self.sigma2 = (sf * sf).astype(np.float32)
        self.inv_sigma2 = (np.float32(1.0) / self.sigma2).astype(np.float32)
        self.nlevels = nlevels
        self.log_sf = float(np.float32(np.log(np.float32(scale_factor))))  # Frame.cc:112 log(float) -> logf


def frame(f, scales, K=EUROC_K, Tcw=None, mp_index=None, outlier=None, fv=None):
    """f: swarmmap_b200.matcher.Frame-like (x, y, octave, angle, desc, bounds = (minx, miny, maxx, maxy)).
    Returns (RFrame, keepalive)."""
    keep = dict(x=_arr(f.x, np.float32), y=_arr(f.y, np.float32), octave=_arr(f.octave, np.int32),
                angle=_arr(f.angle, np.float32), desc=_arr(f.desc, np.uint8), sf=scales.sf, s2=scales.sigma2,
                is2=scales.inv_sigma2, T=_arr(Tcw, np.float32), mp=_arr(mp_index, np.int32),
                out=_arr(outlier, np.uint8))
    if fv is not None:
        keep.update(ids=_arr(fv.node_ids, np.uint32), off=_arr(fv.offsets, np.int32), feats=_arr(fv.feats, np.uint32))
    minx, miny, maxx, maxy = f.bounds
    r = RFrame(len(keep["x"]), _p(keep["x"]), _p(keep["y"]), _p(keep["octave"]), _p(keep["angle"]), _p(keep["desc"]),
               minx, maxx, miny, maxy, K[0], K[1], K[2], K[3], scales.nlevels, scales.log_sf, _p(keep["sf"]),
               _p(keep["s2"]), _p(keep["is2"]), _p(keep["T"]), _p(keep["mp"]), _p(keep["out"]),
               len(keep["ids"]) if fv is not None else 0, _p(keep.get("ids")), _p(keep.get("off")),
               _p(keep.get("feats")))
    return r, keep


def points(pos, desc, nobs=None, bad=None, normal=None, min_dist=None, max_dist=None, track=None):
    """track: dict(in_view, proj_x, proj_y, view_cos, level) or None."""
    keep = dict(pos=_arr(pos, np.float32), desc=_arr(desc, np.uint8), nobs=_arr(nobs, np.int32),
                bad=_arr(bad, np.uint8), normal=_arr(normal, np.float32), mind=_arr(min_dist, np.float32),
                maxd=_arr(max_dist, np.float32))
    t = track or {}
    keep.update(tv=_arr(t.get("in_view"), np.uint8), px=_arr(t.get("proj_x"), np.float32),
                py=_arr(t.get("proj_y"), np.float32), vc=_arr(t.get("view_cos"), np.float32),
                tl=_arr(t.get("level"), np.int32))
    r = RPoints(len(keep["pos"]), _p(keep["pos"]), _p(keep["normal"]), _p(keep["desc"]), _p(keep["nobs"]),
                _p(keep["bad"]), _p(keep["mind"]), _p(keep["maxd"]), _p(keep["tv"]), _p(keep["px"]), _p(keep["py"]),
                _p(keep["vc"]), _p(keep["tl"]))
    return r, keep


def descriptor_distance(a, b):
    a = _arr(a, np.uint8)
    b = _arr(b, np.uint8)
    return lib().refm_descriptor_distance(C.c_void_p(_p(a)), C.c_void_p(_p(b)))


def features_in_area(rf, x, y, r, min_level=-1, max_level=-1):
    out = np.zeros(max(rf.n, 1), np.int32)
    fn = lib().refm_features_in_area
    fn.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]
    n = fn(C.byref(rf), x, y, r, min_level, max_level, _p(out), len(out))
    return out[:n].copy()


def kf_features_in_area(rf, x, y, r):
    out = np.zeros(max(rf.n, 1), np.int32)
    fn = lib().refm_kf_features_in_area
    fn.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int]
    n = fn(C.byref(rf), x, y, r, _p(out), len(out))
    return out[:n].copy()


def grid_csr(rf):
    starts = np.zeros(64 * 48 + 1, np.int32)
    items = np.zeros(max(rf.n, 1), np.int32)
    fn = lib().refm_grid_csr
    fn.restype = None
    fn.argtypes = [C.c_void_p] * 3
    fn(C.byref(rf), _p(starts), _p(items))
    return starts, items[:starts[-1]]


def three_maxima(sizes):
    sizes = _arr(sizes, np.int32)
    out = np.zeros(3, np.int32)
    fn = lib().refm_three_maxima
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    fn(_p(sizes), len(sizes), _p(out))
    return out


def predict_scale(max_distance, dist, log_sf, nlevels):
    fn = lib().refm_predict_scale
    fn.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
    return np.array([fn(float(m), float(d), log_sf, nlevels) for m, d in zip(max_distance, dist)], np.int32)


def is_in_frustum(rf, rp, cos_limit):
    n = rp.n
    iv = np.zeros(n, np.uint8); px = np.zeros(n, np.float32); py = np.zeros(n, np.float32)
    vc = np.zeros(n, np.float32); lv = np.zeros(n, np.int32)
    fn = lib().refm_is_in_frustum
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_float] + [C.c_void_p] * 5
    fn(C.byref(rf), C.byref(rp), cos_limit, _p(iv), _p(px), _p(py), _p(vc), _p(lv))
    return dict(in_view=iv, proj_x=px, proj_y=py, view_cos=vc, level=lv)


def search_for_initialization(rf1, rf2, prev_xy, window, nnratio, check_ori):
    prev = np.ascontiguousarray(prev_xy, np.float32).copy()
    m12 = np.full(rf1.n, -1, np.int32)
    fn = lib().refm_search_for_initialization
    fn.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_float, C.c_int]
    n = fn(C.byref(rf1), C.byref(rf2), _p(prev), _p(m12), window, nnratio, int(check_ori))
    return n, m12, prev


def search_by_projection_last(cur, last, rp, th, nnratio, check_ori):
    out = np.full(cur.n, -1, np.int32)
    fn = lib().refm_search_by_projection_last
    fn.argtypes = [C.c_void_p] * 3 + [C.c_float, C.c_float, C.c_int, C.c_void_p]
    n = fn(C.byref(cur), C.byref(last), C.byref(rp), th, nnratio, int(check_ori), _p(out))
    return n, out


def search_by_projection_points(rf, rp, order, th, nnratio):
    order = _arr(order, np.int32)
    out = np.full(rf.n, -1, np.int32)
    fn = lib().refm_search_by_projection_points
    fn.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_float, C.c_float, C.c_void_p]
    n = fn(C.byref(rf), C.byref(rp), _p(order), len(order), th, nnratio, _p(out))
    return n, out


def search_by_projection_reloc(cur, kf, rp, found, th, orb_dist, nnratio, check_ori):
    found = _arr(found, np.int32)
    out = np.full(cur.n, -1, np.int32)
    fn = lib().refm_search_by_projection_reloc
    fn.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.c_void_p]
    n = fn(C.byref(cur), C.byref(kf), C.byref(rp), _p(found), len(found), th, orb_dist, nnratio, int(check_ori), _p(out))
    return n, out


def search_by_projection_sim3(kf, Scw, rp, order, matched, th):
    order = _arr(order, np.int32)
    Scw = _arr(Scw, np.float32)
    m = np.ascontiguousarray(matched, np.int32).copy()
    fn = lib().refm_search_by_projection_sim3
    fn.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_int]
    n = fn(C.byref(kf), _p(Scw), C.byref(rp), _p(order), len(order), _p(m), int(th))
    return n, m


def search_by_bow_kf_f(kf, f, rp, nnratio, check_ori):
    out = np.full(f.n, -1, np.int32)
    fn = lib().refm_search_by_bow_kf_f
    fn.argtypes = [C.c_void_p] * 3 + [C.c_float, C.c_int, C.c_void_p]
    n = fn(C.byref(kf), C.byref(f), C.byref(rp), nnratio, int(check_ori), _p(out))
    return n, out


def search_by_bow_kf_kf(k1, k2, rp, nnratio, check_ori):
    out = np.full(k1.n, -1, np.int32)
    fn = lib().refm_search_by_bow_kf_kf
    fn.argtypes = [C.c_void_p] * 3 + [C.c_float, C.c_int, C.c_void_p]
    n = fn(C.byref(k1), C.byref(k2), C.byref(rp), nnratio, int(check_ori), _p(out))
    return n, out


def search_for_triangulation(k1, k2, rp, F12, check_ori):
    F12 = _arr(F12, np.float32)
    pairs = np.zeros((max(k1.n, 1), 2), np.int32)
    fn = lib().refm_search_for_triangulation
    fn.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_int]
    n = fn(C.byref(k1), C.byref(k2), C.byref(rp), _p(F12), int(check_ori), _p(pairs), len(pairs))
    return n, pairs[:max(n, 0)].copy()


def fuse(kf, rp, order, th):
    order = _arr(order, np.int32)
    asg = np.full(kf.n, -1, np.int32)
    rep = np.full(rp.n, -1, np.int32)
    fn = lib().refm_fuse
    fn.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    n = fn(C.byref(kf), C.byref(rp), _p(order), len(order), th, _p(asg), _p(rep))
    return n, asg, rep


def fuse_sim3(kf, Scw, rp, order, th):
    order = _arr(order, np.int32)
    Scw = _arr(Scw, np.float32)
    asg = np.full(kf.n, -1, np.int32)
    rep = np.full(len(order), -1, np.int32)
    fn = lib().refm_fuse_sim3
    fn.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    n = fn(C.byref(kf), _p(Scw), C.byref(rp), _p(order), len(order), th, _p(asg), _p(rep))
    return n, asg, rep


def search_by_sim3(k1, k2, rp, matches12, s12, R12, t12, th):
    m = np.ascontiguousarray(matches12, np.int32).copy()
    R12 = _arr(R12, np.float32)
    t12 = _arr(t12, np.float32)
    fn = lib().refm_search_by_sim3
    fn.argtypes = [C.c_void_p] * 4 + [C.c_float, C.c_void_p, C.c_void_p, C.c_float]
    n = fn(C.byref(k1), C.byref(k2), C.byref(rp), _p(m), s12, _p(R12), _p(t12), th)
    return n, m


# ---- numpy mirror of the cv::Mat arithmetic ORBmatcher.cc applies to poses (oracle/ref_shim_matcher's numerics,
# pinned to cv2 in tests/test_ref_orbmatcher.py): used to produce the flat inputs the oracle / C ABI take.

def f32(x):
    return np.float32(x)


def gemm_small(R, x, t=None):
    """R (3x3) * x (3,) + t with cv::gemm's small-matrix path: float products and sums left to right."""
    R = np.asarray(R, np.float32); x = np.asarray(x, np.float32)
    out = np.empty(3, np.float32)
    for r in range(3):
        acc = f32(R[r, 0] * x[0])
        acc = f32(acc + f32(R[r, 1] * x[1]))
        acc = f32(acc + f32(R[r, 2] * x[2]))
        out[r] = f32(np.float64(acc) + np.float64(t[r])) if t is not None else acc
    return out


def gemm_t(R, t, alpha):
    """alpha * R.t() * t with the generic path: accumulation in double."""
    out = np.empty(3, np.float32)
    for r in range(3):
        acc = np.float64(R[0, r]) * np.float64(t[0]) + np.float64(R[1, r]) * np.float64(t[1]) \
            + np.float64(R[2, r]) * np.float64(t[2])
        out[r] = f32(alpha * acc)
    return out


def norm3(v):
    v = np.asarray(v, np.float32)
    return f32(np.sqrt(np.float64(v[0]) * np.float64(v[0]) + np.float64(v[1]) * np.float64(v[1])
                       + np.float64(v[2]) * np.float64(v[2])))


def project(K, pc, inv_double=True):
    """u = fx * xc * invz + cx with invz = (float)(1.0 / z) (ORBmatcher.cc:1257-1263, :1388-1391)."""
    fx, fy, cx, cy = [f32(k) for k in K]
    invz = f32(1.0 / np.float64(pc[2])) if inv_double else f32(f32(1.0) / pc[2])
    u = f32(f32(f32(fx * pc[0]) * invz) + cx)
    v = f32(f32(f32(fy * pc[1]) * invz) + cy)
    return u, v, invz


def stereo_matches(kl, dl, kr, dr, planes_l, planes_r, sf, inv_sf, mbf, mb):
    """Frame::ComputeStereoMatches, the reference's own body (Frame.cc:516-690).  kl / kr: keypoint record arrays
    (x, y, octave); planes_*: per level the BORDERED un-blurred plane (19 px) as a 2-D uint8 array.  -> (uRight, depth)."""
    nl, nr, nlev = len(kl), len(kr), len(planes_l)
    keep = []

    def col(a, name, dt):
        v = np.ascontiguousarray(a[name], dt)
        keep.append(v)
        return _p(v)

    def rois(planes):
        ptrs = (C.c_void_p * nlev)()
        for l, pl in enumerate(planes):
            pl = np.ascontiguousarray(pl, np.uint8)
            keep.append(pl)
            ptrs[l] = pl.ctypes.data + 19 * pl.shape[1] + 19
        return ptrs

    pl_, pr_ = rois(planes_l), rois(planes_r)
    stride = np.array([p.shape[1] for p in planes_l], np.int32)
    w = np.array([p.shape[1] - 38 for p in planes_l], np.int32)
    h = np.array([p.shape[0] - 38 for p in planes_l], np.int32)
    dl, dr = _arr(dl, np.uint8), _arr(dr, np.uint8)
    sf, inv_sf = _arr(sf, np.float32), _arr(inv_sf, np.float32)
    ur, dep = np.zeros(nl, np.float32), np.zeros(nl, np.float32)
    fn = lib().refm_stereo_matches
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 5 + [C.c_int] + \
                  [C.c_void_p] * 2 + [C.c_float] * 2 + [C.c_void_p] * 2
    n = fn(col(kl, "x", np.float32), col(kl, "y", np.float32), col(kl, "octave", np.int32), _p(dl), nl,
           col(kr, "x", np.float32), col(kr, "y", np.float32), col(kr, "octave", np.int32), _p(dr), nr,
           C.cast(pl_, C.c_void_p), C.cast(pr_, C.c_void_p), _p(stride), _p(w), _p(h), nlev, _p(sf), _p(inv_sf),
           mbf, mb, _p(ur), _p(dep))
    return ur, dep, n
