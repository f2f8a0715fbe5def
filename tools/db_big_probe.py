"""Hamming top-2 at the bench size (2000 queries x 2.1 M descriptors) for an ncu capture of db_top2_mma_kernel."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swarmmap_b200 import _lib
lib = _lib.load()
g = torch.Generator(device="cuda"); g.manual_seed(1)
ndb, nq = 8192 * 256, 2000
db = torch.randint(0, 256, (ndb, 32), dtype=torch.uint8, device="cuda", generator=g)
q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device="cuda", generator=g)
h = C.c_void_p()
assert lib.swm_db_create_device(0, db.data_ptr(), ndb, 256, 0, C.byref(h)) == 0
topk = torch.zeros((nq, 2), dtype=torch.int64, device="cuda")
for _ in range(3):
    assert lib.swm_db_query_device(h, q.data_ptr(), nq, 2, topk.data_ptr(), None, 50, None) == 0
torch.cuda.synchronize()
print("ok", int(topk[0, 0]) >> 48)
