"""Batch of 64 frames through the vocabulary transform, for an ncu launch list / capture of the two bow kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swarmmap_b200 import synth
from swarmmap_b200.orb import ORBextractor
from swarmmap_b200.bow import ORBVocabulary
voc = ORBVocabulary(synth.make_vocabulary(10, 5, seed=20220407))
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=64)
k, d, n = ex.extract_batch(synth.make_batch(64, 752, 480, 1))
for _ in range(3):
    out = voc.transform_batch(d, n, 4)
print("ok", len(out[0].word_ids), len(out[0].node_ids))
