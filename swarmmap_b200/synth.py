"""Seeded synthetic frames for the BASELINE.json configs.

No dataset is reachable from the build or GPU boxes, so every test and bench input is generated
here: "EuRoC-shaped" 752x480 and "KITTI-shaped" 1241x376 8-bit frames with enough corner texture
that level 0 yields a few thousand FAST candidates (below the reference's 10 000 cap, Fast.hpp:30).
"""
import numpy as np

try:
    import cv2
except ImportError as e:  # pragma: no cover - cv2 ships in the image
    raise ImportError("swarmmap_b200.synth needs cv2 (present in the build and GPU images)") from e

EUROC = (752, 480)
KITTI = (1241, 376)


def _box_noise(rng, h, w, amp, box):
    n = rng.uniform(-amp, amp, (h // box + 2, w // box + 2)).astype(np.float32)
    n = cv2.resize(n, ((w // box + 2) * box, (h // box + 2) * box), interpolation=cv2.INTER_LINEAR)
    return n[:h, :w]


def make_canvas(w, h, seed, n_rect=600, n_tri=300):
    """Texture canvas: mid-gray + random rectangles/triangles + 3 octaves of noise, lightly blurred."""
    rng = np.random.default_rng(seed)
    scale = (w * h) / float(EUROC[0] * EUROC[1])
    img = np.full((h, w), 110, np.float32)
    for _ in range(int(n_rect * scale)):
        x0, y0 = rng.integers(0, w), rng.integers(0, h)
        sw, sh = rng.integers(6, 91, 2)
        img[y0:y0 + sh, x0:x0 + sw] = rng.uniform(20, 235)
    tri = img.copy()
    for _ in range(int(n_tri * scale)):
        c = np.array([rng.integers(0, w), rng.integers(0, h)])
        pts = (c + rng.integers(-45, 46, (3, 2))).astype(np.int32)
        cv2.fillConvexPoly(tri, pts, float(rng.uniform(20, 235)))
    img = tri
    for amp, box in ((24, 16), (12, 8), (6, 4)):
        img += _box_noise(rng, h, w, amp, box)
    img = cv2.GaussianBlur(img, (0, 0), 0.8)
    return img


def make_frame(w=EUROC[0], h=EUROC[1], seed=20220404):
    """One 8-bit frame (config 1: 752x480 seed 20220404)."""
    rng = np.random.default_rng(seed + 7919)
    img = make_canvas(w, h, seed)
    img = img + rng.normal(0, 2.0, img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def make_sequence(n, w=EUROC[0], h=EUROC[1], seed=20220405, return_h=False):
    """n frames: a larger canvas seen through a slowly moving homography + fresh sensor noise
    (config 2: 1241x376 seed 20220405; configs 3/4: 752x480 seeds 20220406..)."""
    rng = np.random.default_rng(seed + 104729)
    cw, ch = int(w * 1.3) + 32, int(h * 1.4) + 32
    canvas = make_canvas(cw, ch, seed)
    frames = np.empty((n, h, w), np.uint8)
    hs = []
    tx, ty, ang, sc = (cw - w) / 2.0, (ch - h) / 2.0, 0.0, 1.0
    for i in range(n):
        c, s = np.cos(ang) * sc, np.sin(ang) * sc
        cx, cy = w / 2.0, h / 2.0
        # frame pixel -> canvas pixel
        m = np.array([[c, -s, tx + cx - c * cx + s * cy], [s, c, ty + cy - s * cx - c * cy]], np.float64)
        warped = cv2.warpAffine(canvas, m, (w, h), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP,
                                borderMode=cv2.BORDER_REFLECT_101)
        warped = warped + rng.normal(0, 2.0, warped.shape).astype(np.float32)
        frames[i] = np.clip(np.rint(warped), 0, 255).astype(np.uint8)
        hs.append(m.copy())
        tx = float(np.clip(tx + rng.uniform(-6, 6), 8, cw - w - 8))
        ty = float(np.clip(ty + rng.uniform(-6, 6), 8, ch - h - 8))
        ang += float(np.deg2rad(rng.uniform(-0.3, 0.3)))
        sc *= float(1 + rng.uniform(-0.002, 0.002))
    return (frames, hs) if return_h else frames


def make_batch(n, w=EUROC[0], h=EUROC[1], seed=20220410):
    """n independent frames (cheap variant for throughput runs: one canvas, n shifted crops + noise)."""
    rng = np.random.default_rng(seed + 15485863)
    cw, ch = w + 256, h + 256
    canvas = make_canvas(cw, ch, seed)
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        ox, oy = rng.integers(0, 257, 2)
        f = canvas[oy:oy + h, ox:ox + w] + rng.normal(0, 2.0, (h, w)).astype(np.float32)
        out[i] = np.clip(np.rint(f), 0, 255).astype(np.uint8)
    return out


def make_stereo_pair(w=EUROC[0], h=EUROC[1], seed=20220420, d_near=40.0, d_far=4.0, right_shift=0):
    """A rectified stereo pair: the right view is the left view with a row-dependent horizontal disparity (far at the
    top, near at the bottom, like a ground plane; sub-pixel through linear interpolation) + fresh sensor noise.
    right_shift moves the right view the other way (negative disparities, for the matcher's rejection paths)."""
    rng = np.random.default_rng(seed + 32452843)
    canvas = make_canvas(w + 256, h, seed)
    left = canvas[:, 64:64 + w]
    right = np.empty_like(left)
    xs = np.arange(w, dtype=np.float32)
    for y in range(h):
        d = d_far + (d_near - d_far) * y / (h - 1) - right_shift
        src = xs + 64 + d  # right(x) = left(x + d): a point at left column u appears at right column u - d
        x0 = np.floor(src).astype(np.int32)
        f = src - x0
        right[y] = canvas[y, x0] * (1 - f) + canvas[y, x0 + 1] * f
    out = []
    for img in (left, right):
        out.append(np.clip(np.rint(img + rng.normal(0, 2.0, img.shape).astype(np.float32)), 0, 255).astype(np.uint8))
    return out[0], out[1]


def make_vocabulary(k=10, L=3, seed=1, early_leaf=0.03, zero_weight=0.02, weighting=0, scoring=0):
    """A synthetic DBoW2 ORB vocabulary in the binary layout TemplatedVocabulary::loadFromBinaryFile reads
    (reference code/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1478-1522; the real ORBvoc.bin is not shipped):
    header (nb_nodes, size_node = 41, k, L, scoring, weighting) + per node (int32 parent, 32 descriptor bytes,
    float weight, uint8 is_leaf), ids in creation order (a parent's k children are consecutive, then depth first,
    like HKmeansStep).  Children are the parent's descriptor with random bit flips, a few inner nodes stop early and
    a few words carry weight 0 ("stopped" words are dropped by transform)."""
    rng = np.random.default_rng(seed)
    parents, descs, weights, leaves = [], [], [], []

    def add_children(parent_id, parent_desc, depth):
        first = len(parents) + 1
        mine = []
        for _ in range(k):
            bits = np.unpackbits(parent_desc) if parent_desc is not None else rng.integers(0, 2, 256, dtype=np.uint8)
            if parent_desc is not None:
                flip = rng.choice(256, max(4, 96 >> depth), replace=False)
                bits = bits.copy()
                bits[flip] ^= 1
            d = np.packbits(bits)
            parents.append(parent_id)
            descs.append(d)
            weights.append(0.0)
            leaves.append(0)
            mine.append(d)
        for c in range(k):
            nid = first + c
            if depth == L or (depth < L and depth >= 1 and rng.random() < early_leaf):
                leaves[nid - 1] = 1
                weights[nid - 1] = 0.0 if rng.random() < zero_weight else float(np.float32(rng.uniform(0.05, 12.0)))
            else:
                add_children(nid, mine[c], depth + 1)

    add_children(0, None, 1)
    n = len(parents)
    out = bytearray()
    # nb_nodes counts the root too, like the reference's writer (m_nodes.size(), TemplatedVocabulary.h:1530); its reader
    # sizes m_nodes from it, and the parsers here take the record count from the file size
    out += np.array([n + 1, 41], np.uint32).tobytes()
    out += np.array([k, L, scoring, weighting], np.int32).tobytes()
    rec = np.zeros(n, np.dtype([("parent", "<i4"), ("desc", "u1", 32), ("weight", "<f4"), ("leaf", "u1")]))
    assert rec.dtype.itemsize == 41
    rec["parent"] = parents
    rec["desc"] = np.stack(descs)
    rec["weight"] = np.array(weights, np.float32)
    rec["leaf"] = leaves
    out += rec.tobytes()
    return bytes(out)
