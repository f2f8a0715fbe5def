"""ctypes binding of the CPU oracle (oracle/liborb_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package (swarmmap_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")


class Keypoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32), ("class_id", C.c_int32)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
FASTPT_DTYPE = np.dtype([("x", "<i2"), ("y", "<i2"), ("score", "<i4")])


class Frame(C.Structure):
    _fields_ = [("n", C.c_int), ("x", C.c_void_p), ("y", C.c_void_p), ("octave", C.c_void_p),
                ("angle", C.c_void_p), ("desc", C.c_void_p), ("min_x", C.c_float), ("min_y", C.c_float),
                ("max_x", C.c_float), ("max_y", C.c_float)]


class WindowQuery(C.Structure):
    _fields_ = [("m", C.c_int), ("desc", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p),
                ("radius", C.c_void_p), ("min_level", C.c_void_p), ("max_level", C.c_void_p),
                ("valid", C.c_void_p), ("angle", C.c_void_p), ("blocks", C.c_void_p)]


class FeatVec(C.Structure):
    _fields_ = [("n_nodes", C.c_int), ("node_ids", C.c_void_p), ("offsets", C.c_void_p), ("feats", C.c_void_p)]


_lib = None


def build(force=False):
    """SWM_ORACLE_VARIANT=O1 selects the -O1 build (the reference's release level, CMakeLists.txt:24-25) for the CPU
    baseline runs of tools/cpu_baseline.py; the default -O2 library is the one every test uses."""
    name = "liborb_oracle_O1.so" if os.environ.get("SWM_ORACLE_VARIANT") == "O1" else "liborb_oracle.so"
    so = os.path.join(ORACLE_DIR, name)
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".h", ".inc"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", ORACLE_DIR, name], check=True, capture_output=True)
    return so


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.orc_scale_tables.argtypes = [cf, ci, vp, vp, vp, vp]
        L.orc_level_quotas.argtypes = [ci, cf, ci, vp]
        L.orc_umax.argtypes = [vp]
        L.orc_level_sizes.argtypes = [ci, ci, cf, ci, vp, vp]
        L.orc_resize_linear_u8.argtypes = [vp, ci, ci, ci, vp, ci, ci, ci]
        L.orc_border_reflect101.argtypes = [vp, ci, ci, ci, vp, ci, ci]
        L.orc_gauss7_u8.argtypes = [vp, ci, ci, ci, vp, ci]
        L.orc_fast_score_map.argtypes = [vp, ci, ci, ci, ci, vp, ci]
        L.orc_fast_tile_select.argtypes = [vp, ci, ci, ci, ci, vp, ci, vp]
        L.orc_fast_tile_select.restype = ci
        L.orc_octree.argtypes = [vp, ci, ci, ci, ci, ci, ci, vp, ci]
        L.orc_octree.restype = ci
        L.orc_ic_angle.argtypes = [vp, ci, ci, ci]
        L.orc_ic_angle.restype = cf
        L.orc_ic_moments.argtypes = [vp, ci, ci, ci, vp, vp]
        L.orc_rbrief.argtypes = [vp, ci, ci, ci, cf, vp]
        L.orc_hamming256.argtypes = [vp, vp]
        L.orc_hamming256.restype = ci
        L.orc_extractor_create.argtypes = [ci, cf, ci, ci, ci]
        L.orc_extractor_create.restype = vp
        L.orc_extractor_destroy.argtypes = [vp]
        L.orc_extract.argtypes = [vp, vp, ci, ci, ci, vp, vp, ci]
        L.orc_extract.restype = ci
        L.orc_extractor_level.argtypes = [vp, ci, ci, vp, vp, vp, vp]
        L.orc_extractor_level.restype = ci
        L.orc_extractor_level_fast.argtypes = [vp, ci, vp]
        L.orc_extractor_level_fast.restype = ci
        L.orc_extractor_level_selected.argtypes = [vp, ci, vp]
        L.orc_extractor_level_selected.restype = ci
        L.orc_time_extract.argtypes = [vp, vp, ci, ci, ci, ci]
        L.orc_time_extract.restype = C.c_double
        L.orc_grid_build.argtypes = [vp]
        L.orc_grid_build.restype = vp
        L.orc_grid_destroy.argtypes = [vp]
        L.orc_grid_csr.argtypes = [vp, vp, vp]
        L.orc_grid_query.argtypes = [vp, vp, cf, cf, cf, ci, ci, vp, ci]
        L.orc_grid_query.restype = ci
        L.orc_search_for_initialization.argtypes = [vp, vp, vp, vp, ci, cf, ci]
        L.orc_search_for_initialization.restype = ci
        L.orc_match_window.argtypes = [vp, vp, vp, ci, ci, cf, ci, vp]
        L.orc_match_window.restype = ci
        L.orc_search_by_bow.argtypes = [vp, vp, vp, vp, vp, vp, ci, cf, ci, vp]
        L.orc_search_by_bow.restype = ci
        L.orc_bruteforce_top2.argtypes = [vp, ci, vp, C.c_int64, vp]
        L.orc_undistort_points.argtypes = [vp, ci, vp, vp]
        L.orc_vocab_load.restype = vp
        L.orc_vocab_load.argtypes = [vp, C.c_size_t]
        L.orc_vocab_destroy.argtypes = [vp]
        L.orc_vocab_info.argtypes = [vp, vp, vp, vp, vp]
        L.orc_bow_transform.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, vp, vp]
        L.orc_time_bow_transform.restype = C.c_double
        L.orc_time_bow_transform.argtypes = [vp, vp, ci, ci, ci]
        L.orc_image_bounds.argtypes = [ci, ci, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- numpy-friendly wrappers ---------------------------------------------------------------------

def scale_tables(scale_factor=1.2, nlevels=8):
    out = [np.zeros(nlevels, np.float32) for _ in range(4)]
    lib().orc_scale_tables(scale_factor, nlevels, *[_p(o) for o in out])
    return out


def level_quotas(nfeatures, scale_factor=1.2, nlevels=8):
    q = np.zeros(nlevels, np.int32)
    lib().orc_level_quotas(nfeatures, scale_factor, nlevels, _p(q))
    return q


def umax():
    u = np.zeros(16, np.int32)
    lib().orc_umax(_p(u))
    return u


def level_sizes(w, h, scale_factor=1.2, nlevels=8):
    ws, hs = np.zeros(nlevels, np.int32), np.zeros(nlevels, np.int32)
    lib().orc_level_sizes(w, h, scale_factor, nlevels, _p(ws), _p(hs))
    return ws, hs


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), dw, dh, dw)
    return dst


def border_reflect101(src, border=19):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.empty((h + 2 * border, w + 2 * border), np.uint8)
    lib().orc_border_reflect101(_p(src), w, h, w, _p(dst), w + 2 * border, border)
    return dst


def gauss7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().orc_gauss7_u8(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), src.shape[1])
    return dst


def fast_score_map(img, min_th=7):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros_like(img)
    lib().orc_fast_score_map(_p(img), img.shape[1], img.shape[0], img.shape[1], min_th, _p(out), img.shape[1])
    return out


def fast_tile_select(score, ini_th=20, cap=10000, want_retry=False):
    score = np.ascontiguousarray(score, np.uint8)
    h, w = score.shape
    out = np.zeros(cap, FASTPT_DTYPE)
    retry = np.zeros(((h + 31) // 32, (w + 31) // 32), np.uint8)
    n = lib().orc_fast_tile_select(_p(score), w, h, w, ini_th, _p(out), cap, _p(retry))
    return (out[:n], retry) if want_retry else out[:n]


def octree(pts, min_x, max_x, min_y, max_y, target):
    pts = np.ascontiguousarray(pts, FASTPT_DTYPE)
    out = np.zeros(len(pts) + 16, FASTPT_DTYPE)
    n = lib().orc_octree(_p(pts), len(pts), min_x, max_x, min_y, max_y, target, _p(out), len(out))
    return out[:n]


def ic_angle(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    return float(lib().orc_ic_angle(_p(img), img.shape[1], int(x), int(y)))


def ic_moments(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    a, b = C.c_int(), C.c_int()
    lib().orc_ic_moments(_p(img), img.shape[1], int(x), int(y), C.byref(a), C.byref(b))
    return a.value, b.value


def rbrief(img, x, y, angle_deg):
    img = np.ascontiguousarray(img, np.uint8)
    d = np.zeros(32, np.uint8)
    lib().orc_rbrief(_p(img), img.shape[1], int(x), int(y), float(angle_deg), _p(d))
    return d


def hamming256(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return int(lib().orc_hamming256(_p(a), _p(b)))


class Extractor:
    """CPU oracle of ORBextractor (reference ORBextractor.cc:340-855)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.h = lib().orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        if not self.h:
            raise ValueError("bad extractor parameters")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_extractor_destroy(self.h)
            self.h = None

    def __call__(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        cap = 4 * self.nfeatures + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = lib().orc_extract(self.h, _p(img), img.shape[1], img.shape[0], img.shape[1], _p(kps), _p(desc), cap)
        if n < 0:
            raise ValueError("orc_extract failed")
        return kps[:n].copy(), desc[:n].copy()

    def level(self, level, which):
        data, w, h, s = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
        rc = lib().orc_extractor_level(self.h, level, which, C.byref(data), C.byref(w), C.byref(h), C.byref(s))
        assert rc == 0
        buf = (C.c_uint8 * (s.value * h.value)).from_address(data.value)
        return np.frombuffer(buf, np.uint8).reshape(h.value, s.value)[:, :w.value].copy()

    def _pts(self, fn, level):
        p = C.c_void_p()
        n = fn(self.h, level, C.byref(p))
        assert n >= 0
        if n == 0:
            return np.zeros(0, FASTPT_DTYPE)
        buf = (C.c_uint8 * (n * FASTPT_DTYPE.itemsize)).from_address(p.value)
        return np.frombuffer(buf, FASTPT_DTYPE).copy()

    def level_fast(self, level):
        return self._pts(lib().orc_extractor_level_fast, level)

    def level_selected(self, level):
        return self._pts(lib().orc_extractor_level_selected, level)

    def time(self, imgs, iters):
        imgs = np.ascontiguousarray(imgs, np.uint8)
        n, h, w = imgs.shape
        return lib().orc_time_extract(self.h, _p(imgs), n, w, h, iters)


def make_frame(x, y, octave, angle, desc, bounds):
    """Returns (Frame struct, keepalive tuple)."""
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    octave = np.ascontiguousarray(octave, np.int32)
    angle = np.ascontiguousarray(angle, np.float32)
    desc = np.ascontiguousarray(desc, np.uint8)
    f = Frame(len(x), _p(x).value, _p(y).value, _p(octave).value, _p(angle).value, _p(desc).value,
              *[float(b) for b in bounds])
    return f, (x, y, octave, angle, desc)


# ---- matcher wrappers (frames are swarmmap_b200.matcher.Frame-like objects: x,y,octave,angle,desc,bounds) ----

def _frame(f):
    return make_frame(f.x, f.y, f.octave, f.angle, f.desc, f.bounds)


def grid_csr(f):
    fr, keep = _frame(f)
    g = lib().orc_grid_build(C.byref(fr))
    starts = np.zeros(64 * 48 + 1, np.int32)
    items = np.zeros(max(f.N, 1), np.int32)
    lib().orc_grid_csr(g, _p(starts), _p(items))
    lib().orc_grid_destroy(g)
    return starts, items[:starts[-1]]


def search_for_initialization(f1, f2, prev_xy, window, nnratio, check_ori):
    a, k1 = _frame(f1)
    b, k2 = _frame(f2)
    prev = np.ascontiguousarray(prev_xy, np.float32).copy()
    m12 = np.full(f1.N, -1, np.int32)
    n = lib().orc_search_for_initialization(C.byref(a), C.byref(b), _p(prev), _p(m12), int(window), float(nnratio),
                                            int(check_ori))
    return n, m12, prev


def match_window(tgt, desc, u, v, radius, min_level, max_level, valid, blocks, th_dist, ratio_mode, nnratio,
                 check_ori, angle=None, tgt_blocked=None, assignment=None):
    t, keep = _frame(tgt)
    desc = np.ascontiguousarray(desc, np.uint8)
    arr = [np.ascontiguousarray(u, np.float32), np.ascontiguousarray(v, np.float32),
           np.ascontiguousarray(radius, np.float32), np.ascontiguousarray(min_level, np.int32),
           np.ascontiguousarray(max_level, np.int32), np.ascontiguousarray(valid, np.uint8)]
    ang = np.ascontiguousarray(angle, np.float32) if angle is not None else np.zeros(len(u), np.float32)
    blk = np.ascontiguousarray(blocks, np.uint8)
    q = WindowQuery(len(u), _p(desc).value, *[_p(a).value for a in arr], _p(ang).value, _p(blk).value)
    asg = np.full(tgt.N, -1, np.int32) if assignment is None else assignment.copy()
    tb = np.ascontiguousarray(tgt_blocked, np.uint8) if tgt_blocked is not None else None
    n = lib().orc_match_window(C.byref(t), C.byref(q), _p(tb) if tb is not None else None, int(th_dist),
                               int(ratio_mode), float(nnratio), int(check_ori), _p(asg))
    return n, asg


def search_by_bow(f1, fv1, valid1, f2, fv2, valid2, mode, nnratio, check_ori):
    a, k1 = _frame(f1)
    b, k2 = _frame(f2)
    v1 = np.ascontiguousarray(valid1, np.uint8)
    v2 = np.ascontiguousarray(valid2, np.uint8) if valid2 is not None else np.ones(f2.N, np.uint8)
    fa = FeatVec(len(fv1.node_ids), _p(fv1.node_ids).value, _p(fv1.offsets).value, _p(fv1.feats).value)
    fb = FeatVec(len(fv2.node_ids), _p(fv2.node_ids).value, _p(fv2.offsets).value, _p(fv2.feats).value)
    out = np.full(f2.N if mode == 0 else f1.N, -1, np.int32)
    n = lib().orc_search_by_bow(C.byref(a), C.byref(fa), _p(v1), C.byref(b), C.byref(fb), _p(v2), int(mode),
                                float(nnratio), int(check_ori), _p(out))
    return n, out


def search_for_triangulation(f1, fv1, valid1, f2, fv2, valid2, F12, ex, ey, scale_factors2, level_sigma2, check_ori):
    a, k1 = _frame(f1)
    b, k2 = _frame(f2)
    v1 = np.ascontiguousarray(valid1, np.uint8)
    v2 = np.ascontiguousarray(valid2, np.uint8)
    fa = FeatVec(len(fv1.node_ids), _p(fv1.node_ids).value, _p(fv1.offsets).value, _p(fv1.feats).value)
    fb = FeatVec(len(fv2.node_ids), _p(fv2.node_ids).value, _p(fv2.offsets).value, _p(fv2.feats).value)
    F = np.ascontiguousarray(F12, np.float32).reshape(9)
    sf = np.ascontiguousarray(scale_factors2, np.float32)
    s2 = np.ascontiguousarray(level_sigma2, np.float32)
    out = np.full(f1.N, -1, np.int32)
    fn = lib().orc_search_for_triangulation
    fn.argtypes = [C.c_void_p] * 7 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    n = fn(C.byref(a), C.byref(fa), _p(v1), C.byref(b), C.byref(fb), _p(v2), _p(F), float(ex), float(ey), _p(sf), _p(s2),
           int(check_ori), _p(out))
    return n, out


def window_best(tgt, desc, u, v, radius, pred_level, valid, inv_level_sigma2, chi2):
    t, keep = _frame(tgt)
    desc = np.ascontiguousarray(desc, np.uint8)
    arr = [np.ascontiguousarray(a, np.float32) for a in (u, v, radius)]
    pl = np.ascontiguousarray(pred_level, np.int32)
    va = np.ascontiguousarray(valid, np.uint8)
    s2 = np.ascontiguousarray(inv_level_sigma2, np.float32)
    m = len(pl)
    bi = np.zeros(m, np.int32)
    bd = np.zeros(m, np.int32)
    fn = lib().orc_window_best
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_float, C.c_void_p, C.c_void_p]
    fn(C.byref(t), m, _p(desc), _p(arr[0]), _p(arr[1]), _p(arr[2]), _p(pl), _p(va), _p(s2), float(chi2), _p(bi), _p(bd))
    return bi, bd


def distinctive_descriptors(desc, offsets):
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    offsets = np.ascontiguousarray(offsets, np.int32)
    n = len(offsets) - 1
    best = np.zeros(n, np.int32)
    med = np.zeros(n, np.int32)
    fn = lib().orc_distinctive_descriptors
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    fn.restype = None
    fn(_p(desc), _p(offsets), n, _p(best), _p(med))
    return best, med


class Pyramid(C.Structure):
    _fields_ = [("nlevels", C.c_int), ("plane", C.c_void_p), ("stride", C.c_void_p), ("w", C.c_void_p), ("h", C.c_void_p)]


def stereo_matches(kl, dl, kr, dr, planes_l, planes_r, sf, inv_sf, mbf, mb):
    """Oracle of Frame::ComputeStereoMatches (Frame.cc:516-690).  kl / kr: KP_DTYPE arrays; planes_*: per level the
    BORDERED un-blurred plane (19 px) as a 2-D uint8 array.  -> (uRight, depth, kept)."""
    nlev = len(planes_l)
    keep = []

    def pyr(planes):
        ptrs = (C.c_void_p * nlev)()
        for l, pl in enumerate(planes):
            pl = np.ascontiguousarray(pl, np.uint8)
            keep.append(pl)
            ptrs[l] = pl.ctypes.data + 19 * pl.shape[1] + 19
        stride = np.array([p.shape[1] for p in planes], np.int32)
        w = np.array([p.shape[1] - 38 for p in planes], np.int32)
        h = np.array([p.shape[0] - 38 for p in planes], np.int32)
        keep.extend([ptrs, stride, w, h])
        return Pyramid(nlev, C.cast(ptrs, C.c_void_p), _p(stride), _p(w), _p(h))

    kl = np.ascontiguousarray(kl, KP_DTYPE)
    kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl, dr = np.ascontiguousarray(dl, np.uint8), np.ascontiguousarray(dr, np.uint8)
    sf, inv_sf = np.ascontiguousarray(sf, np.float32), np.ascontiguousarray(inv_sf, np.float32)
    pl, pr = pyr(planes_l), pyr(planes_r)
    ur, dep = np.zeros(len(kl), np.float32), np.zeros(len(kl), np.float32)
    fn = lib().orc_stereo_matches
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    n = fn(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), C.byref(pl), C.byref(pr), _p(sf), _p(inv_sf), mbf, mb,
           _p(ur), _p(dep))
    return ur, dep, n


def bruteforce_top2(q, db):
    q = np.ascontiguousarray(q, np.uint8)
    db = np.ascontiguousarray(db, np.uint8)
    out = np.zeros((len(q), 4), np.int32)
    lib().orc_bruteforce_top2(_p(q), len(q), _p(db), len(db), _p(out))
    return out


def undistort_points(xy, cam9):
    xy = np.ascontiguousarray(xy, np.float32)
    cam = np.ascontiguousarray(cam9, np.float32)
    out = np.zeros_like(xy)
    lib().orc_undistort_points(_p(xy), len(xy), _p(cam), _p(out))
    return out


def image_bounds(cols, rows, cam9):
    cam = np.ascontiguousarray(cam9, np.float32)
    out = np.zeros(4, np.float32)
    lib().orc_image_bounds(cols, rows, _p(cam), _p(out))
    return out


class Vocabulary:
    """Oracle restatement of DBoW2's TemplatedVocabulary<FORB> (load from the ORBvoc.bin layout + transform)."""

    def __init__(self, blob):
        self._blob = np.frombuffer(blob, np.uint8).copy()
        self._h = lib().orc_vocab_load(_p(self._blob), len(self._blob))
        if not self._h:
            raise ValueError("bad vocabulary blob")
        a = [C.c_int32() for _ in range(4)]
        lib().orc_vocab_info(self._h, *[C.byref(x) for x in a])
        self.k, self.L, self.n_nodes, self.n_words = [x.value for x in a]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_vocab_destroy(self._h)
            self._h = None

    def transform(self, desc, levelsup=4):
        """-> (word_ids, word_values, node_ids, offsets, feats): BowVector + FeatureVector (CSR) of one frame."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        wid = np.zeros(n + 1, np.uint32); wv = np.zeros(n + 1, np.float64)
        nid = np.zeros(n + 1, np.uint32); off = np.zeros(n + 2, np.int32); feats = np.zeros(n + 1, np.uint32)
        nn = C.c_int32(0)
        nw = lib().orc_bow_transform(self._h, _p(desc), n, int(levelsup), _p(wid), _p(wv), _p(nid), _p(off), _p(feats),
                                     C.byref(nn))
        return wid[:nw], wv[:nw], nid[:nn.value], off[:nn.value + 1], feats[:off[nn.value]]

    def time_transform(self, desc, levelsup, iters):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        return lib().orc_time_bow_transform(self._h, _p(desc), len(desc), int(levelsup), int(iters))
