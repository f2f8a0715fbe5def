"""One single-frame operator() sequence for an ncu launch list (SWM_NO_GRAPH=1 so that kernels are listed one by one)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swarmmap_b200 import synth
from swarmmap_b200.orb import ORBextractor
frames = synth.make_batch(4, 752, 480, 3)
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=1)
for f in frames:
    ex(f)
