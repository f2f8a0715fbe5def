"""Small profiling driver: a few device-resident extract steps (for ncu)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swarmmap_b200 import synth
from swarmmap_b200.orb import ORBextractor
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
frames = synth.make_batch(B, 752, 480, 20220410)
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=B)
cap = ex.max_keypoints()
d_img = torch.from_numpy(frames).to(dev)
d_kps = torch.empty((B, cap, 7), dtype=torch.float32, device=dev)
d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
d_n = torch.zeros(B, dtype=torch.int32, device=dev)
st = torch.cuda.Stream(dev)
for _ in range(steps):
    ex.extract_batch_device(d_img.data_ptr(), B, 752, 480, 752, 752 * 480, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr(), C.c_void_p(st.cuda_stream))
torch.cuda.synchronize()
print("kp/frame", d_n.float().mean().item())
