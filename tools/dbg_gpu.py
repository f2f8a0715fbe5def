import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import oracle_lib as o
from swarmmap_b200 import synth
from swarmmap_b200.orb import ORBextractor
g = ORBextractor(1000, 1.2, 8, 20, 7, debug_score=True)
c = o.Extractor(1000)
img = synth.make_frame()
k, d = g(img)
ok, od = c(img)
for l in range(8):
    for which, name in ((0, 'plain'), (2, 'score'), (1, 'blur')):
        a = g.debug_plane(0, l, which); b = c.level(l, which)
        bad = a != b
        print(l, name, a.shape, 'mismatch', int(bad.sum()))
        if bad.sum() and l == 0:
            ys, xs = np.nonzero(bad)
            print('  first', list(zip(ys[:12].tolist(), xs[:12].tolist())))
            print('  gpu', a[ys[:12], xs[:12]].tolist(), 'cpu', b[ys[:12], xs[:12]].tolist())
            print('  x%64 hist', np.bincount(xs % 64, minlength=64).tolist())
            print('  y%32 hist', np.bincount(ys % 32, minlength=32).tolist())
            print('  gpu nonzero', int((a > 0).sum()), 'cpu nonzero', int((b > 0).sum()), 'gpu>0&cpu==0', int(((a > 0) & (b == 0)).sum()), 'gpu==0&cpu>0', int(((a == 0) & (b > 0)).sum()))
    gp = g.debug_points(0, l, 0); op = c.level_fast(l)
    gs = g.debug_points(0, l, 1); os_ = c.level_selected(l)
    print(l, 'fast', len(gp), len(op), 'sel', len(gs), len(os_))
print(len(k), len(ok))
