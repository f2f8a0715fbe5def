"""Small profiling driver: stereo front-end on 32 rectified pairs (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swarmmap_b200 import synth
from swarmmap_b200.orb import ORBextractor
B = 32
l, r = synth.make_stereo_pair(752, 480, 20220421)
exl, exr = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=B), ORBextractor(1000, 1.2, 8, 20, 7, max_batch=B)
exl.extract_batch(np.repeat(l[None], B, 0))
exr.extract_batch(np.repeat(r[None], B, 0))
for _ in range(3):
    u, z = exl.stereo_match(exr, 47.90639384423901, 47.90639384423901 / 458.654, B)
print("matches per pair", int((u[0] >= 0).sum()))
