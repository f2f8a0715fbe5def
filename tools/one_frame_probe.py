import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from swarmmap_b200 import synth
from swarmmap_b200.orb import ORBextractor
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=1)
fr = synth.make_frame()
try:
    k, d = ex(fr)
    print("ok", len(k))
except Exception as e:
    print("ERR", e)
