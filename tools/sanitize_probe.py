"""Small run of every kernel for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from swarmmap_b200 import synth, place
from swarmmap_b200.orb import ORBextractor
from swarmmap_b200.matcher import Frame, ORBmatcher, FeatureVector
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=2)
imgs = synth.make_batch(3, 752, 480, 1)
k, d, n = ex.extract_batch(imgs)
rng = np.random.default_rng(0)
noise = rng.integers(0, 256, (333, 257), dtype=np.uint8)   # odd size, > 10000 candidates on level 0? (cap path on big noise)
ex2 = ORBextractor(500, 1.2, 8, 20, 7)
ex2(noise)
big = rng.integers(0, 256, (480, 752), dtype=np.uint8)
ex(big)
fs = [Frame.from_keypoints(k[i, :n[i]], d[i, :n[i]], 752, 480, ex.GetScaleFactors()) for i in range(2)]
m = ORBmatcher(0.9, True)
prev = np.stack([fs[0].x, fs[0].y], 1).astype(np.float32).copy()
print("init", m.SearchForInitialization(fs[0], fs[1], prev, 100)[0])
print("proj", m.SearchByProjectionLastFrame(fs[1], fs[0], fs[0].x, fs[0].y, np.ones(fs[0].N, np.uint8), 15)[0])
node = lambda f: (f.desc[:, 0] & 31).astype(np.int64)
print("bow", m.SearchByBoW(fs[0], FeatureVector(node(fs[0])), np.ones(fs[0].N, np.uint8), fs[1], FeatureVector(node(fs[1])))[0])
db = torch.randint(0, 256, (5000, 32), dtype=torch.uint8, device="cuda")
q = torch.randint(0, 256, (300, 32), dtype=torch.uint8, device="cuda")
s = place.PlaceShard(db, 8, 0)
keys, votes = s.query(q, 2, 50)
torch.cuda.synchronize()
# resident frames + vocabulary transform
from swarmmap_b200.matcher import Camera, ResidentFrame
from swarmmap_b200.bow import ORBVocabulary
cam = Camera(458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)
b4 = cam.bounds(752, 480)
ex.extract_batch(imgs[:2])
rf = [ResidentFrame().from_extractor(ex, i, cam, b4) for i in range(2)]
print("resident proj", m.SearchByProjectionLastFrame(rf[1], fs[0], fs[0].x, fs[0].y, np.ones(fs[0].N, np.uint8), 15)[0])
voc = ORBVocabulary(synth.make_vocabulary(10, 3, seed=2))
r0, r1 = voc.transform_frame(rf[0], 2), voc.transform_frame(rf[1], 2)
print("bow words", len(r0.word_ids), "nodes", len(r0.node_ids), "batch", len(voc.transform_batch(d[:2], n[:2], 2)))
print("resident bow", m.SearchByBoW(rf[0], r0.feature_vector(), np.ones(rf[0].N, np.uint8), rf[1], r1.feature_vector())[0])
# stereo front-end (two extractors)
sl, sr = synth.make_stereo_pair(752, 480, 5)
exl, exr = ORBextractor(800, 1.2, 8, 20, 7, max_batch=2), ORBextractor(800, 1.2, 8, 20, 7, max_batch=2)
exl.extract_batch(np.stack([sl, sl]))
exr.extract_batch(np.stack([sr, sr]))
u, z = exl.stereo_match(exr, 47.9, 47.9 / 458.654, 2)
print("stereo", int((u >= 0).sum()))
# (the peer-memory place exchange is NOT run here: compute-sanitizer serialises kernel launches, and the merge kernels of
# two ranks wait for each other -- tests/test_gpu_match.py::test_db_query_peers_equals_single_shard covers it; a world
# of one rank still runs the kernel's push / flag / wait / merge path)
one = [place.PlaceShard(db.cpu().numpy(), 8, 0)]
one[0].enable_peers(nq_max=512, same_process=one)
k1, v1 = one[0].query_peers(q, 2, 50)
torch.cuda.synchronize()
assert torch.equal(k1, keys)
print("peers (world 1) ok")
# batched matchers (one resolve CTA per job)
res = m.SearchForInitializationBatch([(fs[0], fs[1], prev.copy()), (fs[1], fs[0], np.ascontiguousarray(np.stack([fs[1].x, fs[1].y], 1).astype(np.float32)))], 100)
print("init batch", [int(r[0]) for r in res])
print("ok", int(n.sum()))
