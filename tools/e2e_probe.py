"""e2e diagnostics: raw pinned H2D/D2H bandwidth and the e2e pipeline with different chunk sizes / handle counts."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from swarmmap_b200 import synth
from swarmmap_b200.orb import ORBextractor, KP_DTYPE
W, H = 752, 480
B = 2048
frames = np.tile(synth.make_batch(256, W, H, 20220410), (B // 256, 1, 1))
h_img = torch.from_numpy(frames).pin_memory()
d = torch.empty_like(h_img, device="cuda")
for _ in range(3): d.copy_(h_img, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): d.copy_(h_img, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("raw pinned H2D GB/s", 10 * h_img.numel() / dt / 1e9)
h_np = h_img.numpy()
for nslot, eb in ((8, 32), (8, 64), (6, 64), (12, 32), (16, 32), (4, 128), (12, 64)):
    exs = [ORBextractor(1000, 1.2, 8, 20, 7, max_batch=eb) for _ in range(nslot)]
    cap = exs[0].max_keypoints()
    outs = []
    for _ in range(nslot):
        k = torch.empty((eb, cap, 7), dtype=torch.float32).pin_memory(); dd = torch.empty((eb, cap, 32), dtype=torch.uint8).pin_memory(); n = torch.zeros(eb, dtype=torch.int32).pin_memory()
        outs.append((k.numpy().view(KP_DTYPE).reshape(eb, cap), dd.numpy(), n.numpy(), (k, dd, n)))
    chunks = [(i, min(eb, B - i)) for i in range(0, B, eb)]
    def step():
        pend = [None] * nslot
        for ci, (f0, nb) in enumerate(chunks):
            s = ci % nslot
            if pend[s] is not None: exs[s].sync()
            exs[s].extract_batch_async(h_np[f0:f0 + nb], (outs[s][0][:nb], outs[s][1][:nb], outs[s][2][:nb]))
            pend[s] = nb
        for s in range(nslot):
            if pend[s] is not None: exs[s].sync()
    for _ in range(3): step()
    t = time.perf_counter()
    for _ in range(10): step()
    dt = time.perf_counter() - t
    print(f"slots {nslot} chunk {eb}: {10 * B / dt:.0f} frames/s")
    del exs
