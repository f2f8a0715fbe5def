import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import numpy as np, torch
import oracle_lib as o
from swarmmap_b200 import _lib
lib = _lib.load()
rng = np.random.default_rng(1)
q = rng.integers(0, 256, (40, 32), dtype=np.uint8)
db = rng.integers(0, 256, (300, 32), dtype=np.uint8)
db[7] = q[3]; db[200] = q[3]
ref = o.bruteforce_top2(q, db)
h = C.c_void_p()
assert lib.swm_db_create(0, _lib.ptr(db), len(db), 4, 0, C.byref(h)) == 0
dq = torch.from_numpy(q).cuda()
topk = torch.zeros((40, 2), dtype=torch.int64, device="cuda")
rc = lib.swm_db_query_device(h, dq.data_ptr(), 40, 2, topk.data_ptr(), None, 50, None)
torch.cuda.synchronize()
t = topk.cpu().numpy().astype(np.uint64)
dist = (t >> np.uint64(48)).astype(np.int64); idx = (t & np.uint64((1 << 48) - 1)).astype(np.int64)
for i in range(12):
    print(i, 'gpu', dist[i].tolist(), idx[i].tolist(), 'ref', ref[i].tolist())
true = np.unpackbits(q[:, None, :] ^ db[None, :, :], axis=2).sum(2)
print('true dist of gpu idx', [int(true[i, idx[i, 0]]) if 0 <= idx[i,0] < 300 else None for i in range(12)])
