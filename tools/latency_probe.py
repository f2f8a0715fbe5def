"""Single-frame operator() latency (host image in, host keypoints out), with / without SWM_NO_GRAPH."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swarmmap_b200 import synth
from swarmmap_b200.orb import ORBextractor
for (w, h, nf) in ((752, 480, 1000), (1241, 376, 2000)):
    frames = synth.make_batch(8, w, h, 3)
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_batch=1)
    ref = [ex(f) for f in frames]      # call 1 normal, call 2 captures, later calls replay
    for _ in range(20):
        ex(frames[0])
    t0 = time.perf_counter()
    for i in range(200):
        k, d = ex(frames[i % 8])
    dt = (time.perf_counter() - t0) / 200 * 1e3
    same = all(np.array_equal(ex(frames[i])[0], ref[i][0]) and np.array_equal(ex(frames[i])[1], ref[i][1]) for i in range(8))
    print(f"{w}x{h} nfeat {nf}: {dt:.4f} ms per operator() call, results identical to the first calls: {same}")
