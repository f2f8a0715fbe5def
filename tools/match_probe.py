import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swarmmap_b200 import synth
from swarmmap_b200.orb import ORBextractor
from swarmmap_b200.matcher import Frame, ORBmatcher
seq = synth.make_sequence(2, 1241, 376, 20220405)
ex = ORBextractor(4000, 1.2, 8, 20, 7)
fs = [Frame.from_keypoints(*ex(img), 1241, 376, ex.GetScaleFactors()) for img in seq]
m = ORBmatcher(0.9, True)
u, v = fs[0].x.copy(), fs[0].y.copy()
valid = np.ones(fs[0].N, np.uint8)
for _ in range(3):
    m.SearchByProjectionLastFrame(fs[1], fs[0], u, v, valid, 15)
os.environ["SWM_MATCH_PROFILE"] = "1"
t = time.perf_counter()
m.SearchByProjectionLastFrame(fs[1], fs[0], u, v, valid, 15)
print("total call ms", (time.perf_counter() - t) * 1e3)
