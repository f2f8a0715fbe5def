"""Stage times (ms per 256 frames, CUDA events) of the extractor: pyramid, FAST, quadtree, describe."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swarmmap_b200 import synth, _lib
from swarmmap_b200.orb import ORBextractor
B = 256
dev = torch.device("cuda", 0)
frames = synth.make_batch(B, 752, 480, 20220410)
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=B)
cap = ex.max_keypoints()
d_img = torch.from_numpy(frames).to(dev)
d_kps = torch.empty((B, cap, 7), dtype=torch.float32, device=dev)
d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
d_n = torch.zeros(B, dtype=torch.int32, device=dev)
st = torch.cuda.Stream(dev)
torch.cuda.set_stream(st)
sp = C.c_void_p(st.cuda_stream)
ex.extract_batch_device(d_img.data_ptr(), B, 752, 480, 752, 752 * 480, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr(), sp)
torch.cuda.synchronize()
out = []
for name, mask in (("pyr", _lib.STAGE_PYRAMID), ("fast", _lib.STAGE_NMS), ("pyr+fast", _lib.STAGE_PYRAMID | _lib.STAGE_NMS),
                   ("octree", _lib.STAGE_OCTREE), ("describe", _lib.STAGE_DESCRIBE), ("all", 15)):
    for _ in range(3):
        ex.run_stage(mask, B, sp)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ex.run_stage(mask, B, sp)
    b.record()
    torch.cuda.synchronize()
    out.append(f"{name} {a.elapsed_time(b) / 20:.4f}")
print(" ".join(out), "kp", int(d_n.sum().item()))
