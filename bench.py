#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native ORB front-end (BASELINE.json metric).

A "step" is one pass of extract+describe (ORBextractor::operator(), 8 levels, scale 1.2, FAST 20/7,
1000 features) over one batch of synthetic 752x480 EuRoC-shaped frames on every rank's own GPU
(one agent stream set per GPU, no data-path collective: the path shards by agent -> "weak" scaling).

  value      frames/s, whole job, inputs and outputs resident in HBM (CUDA events on the launching stream)
  e2e        frames/s through the public host-buffer API (pinned host frames in, keypoints +
             descriptors out, H2D/D2H inside the timed region, two handles double-buffering)
  roofline   pyramid+FAST kernels (fused level kernels + NMS), algorithmic bytes / live event time
  cpu_baseline  the CPU oracle (port of the reference path) on this box's host cores, bounded sample

`--impl reference` times the reference path's CPU restatement (oracle/, the reference itself cannot
be built: SURVEY.md F3) with all host threads on the same workload and prints the same JSON line.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H, NFEAT = 752, 480, 1000
PYR_FAST_BYTES = 3912047   # pyramid R+W + FAST R per 752x480 frame
BLUR_BYTES = 2234734       # blur R+W per frame (fused into the same kernels)
METRIC = "orb_extract_frames_per_sec"  # BASELINE.json: ORB frames/sec (1000 feat, 752x480), whole job over all GPUs
UNIT = "frames/s"
WORKLOAD = "EuRoC-shaped 752x480 8-bit frames, ORBextractor(1000, 1.2, 8, 20, 7), extract+describe"


def _traffic_per_frame():
    """DRAM bytes per frame of the roofline kernels from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r1k_traffic.json")
    try:
        return float(json.load(open(p))["dram_bytes_per_frame"])
    except Exception:
        return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _reference_dbow2_ms(vblob, desc, levelsup, iters):
    """ms per Frame::ComputeBoW with the reference's own DBoW2 compiled unmodified (oracle/_ref/libdbow2_ref.so);
    None when that library is not there.  cpu_baseline leg only."""
    lib_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", "libdbow2_ref.so")
    if not os.path.exists(lib_path):
        return None
    try:
        import ctypes as C
        import tempfile
        L = C.CDLL(lib_path)
        L.ref_vocab_load.restype = C.c_void_p
        L.ref_vocab_load.argtypes = [C.c_char_p]
        L.ref_vocab_destroy.argtypes = [C.c_void_p]
        L.ref_bow_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
        with tempfile.NamedTemporaryFile(suffix=".bin") as f:
            f.write(vblob)
            f.flush()
            h = L.ref_vocab_load(f.name.encode())
        if not h:
            return None
        d = np.ascontiguousarray(desc, np.uint8)
        n = len(d)
        bufs = [np.zeros(n + 2, t) for t in (np.uint32, np.float64, np.uint32, np.int32, np.uint32)]
        nn = C.c_int32(0)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        t0 = time.perf_counter()
        for _ in range(iters):
            L.ref_bow_transform(h, ptr(d), n, levelsup, *[ptr(b) for b in bufs], C.byref(nn))
        ms = (time.perf_counter() - t0) / iters * 1e3
        L.ref_vocab_destroy(h)
        return ms
    except Exception:  # a baseline that cannot be measured is omitted, never fatal
        return None


def cpu_oracle_rate(frames, seconds_budget, threads):
    """frames/s of the CPU oracle on `threads` host threads over a bounded sample."""
    import oracle_lib
    from concurrent.futures import ThreadPoolExecutor
    ex = [oracle_lib.Extractor(NFEAT, 1.2, 8, 20, 7) for _ in range(threads)]
    t0 = time.perf_counter()
    ex[0].time(frames[:1], 1)
    per = max(time.perf_counter() - t0, 1e-3)
    iters = max(2, min(200, int(seconds_budget / per)))

    def run(i):
        return ex[i].time(frames[i % len(frames):i % len(frames) + 1], iters)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(run, range(threads)))
    dt = time.perf_counter() - t0
    return threads * iters / dt, threads * iters


def run_reference(args):
    from swarmmap_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    frames = synth.make_batch(max(8, threads), W, H, 20220410)
    for _ in range(max(args.warmup, 1) - 1):
        cpu_oracle_rate(frames, 0.5, threads)
    rates, n_total = [], 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r, n = cpu_oracle_rate(frames, 4.0, threads)
        rates.append(r)
        n_total += n
    dt = time.perf_counter() - t0
    value = statistics.median(rates)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frame": [W, H], "nfeatures": NFEAT},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n_total} frames over {args.steps} steps, oracle -O2, one extractor per thread"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="frames per step per GPU")
    ap.add_argument("--impl", default="swm", choices=["swm", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "swm" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    from swarmmap_b200 import build, synth
    build.build()
    from swarmmap_b200.orb import ORBextractor, KP_DTYPE
    from swarmmap_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ORB front-end has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    B = args.batch

    # ---- synthetic input: B distinct frames per rank (one agent stream per GPU)
    frames = synth.make_batch(B, W, H, 20220410 + rank)
    ex = ORBextractor(NFEAT, 1.2, 8, 20, 7, device=local_rank, max_batch=B)
    cap = ex.max_keypoints()

    # ---- device-resident arm
    d_img = torch.from_numpy(frames).to(dev)
    d_kps = torch.empty((B, cap, 7), dtype=torch.float32, device=dev)
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(B, dtype=torch.int32, device=dev)
    # a dedicated (non-default) torch stream: the kernels are enqueued on it through the C ABI and the
    # torch.cuda.Events below are recorded on the same stream.
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    assert sptr.value, "need a non-default stream handle"

    # The step's B frames are split over NH extractor handles (B/NH frames each), every handle on its own
    # stream, as co-located agents would run: the short quadtree/describe tails of one handle overlap the
    # pyramid/FAST kernels of another.  Timing: e0 on `stream`, every work stream waits for it, runs its K
    # steps, and `stream` waits for all of them before e1 (device time of the whole job, no host clock).
    NH = 4 if B % 4 == 0 and B >= 64 else 1
    hb = B // NH
    dex = [ORBextractor(NFEAT, 1.2, 8, 20, 7, device=local_rank, max_batch=hb) for _ in range(NH)]
    wstreams = [torch.cuda.Stream(dev) for _ in range(NH)]
    wptrs = [C.c_void_p(ws.cuda_stream) for ws in wstreams]

    def step_device():
        for i in range(NH):
            f0 = i * hb
            dex[i].extract_batch_device(d_img[f0:].data_ptr(), hb, W, H, W, W * H, d_kps[f0:].data_ptr(),
                                        d_desc[f0:].data_ptr(), cap, d_n[f0:].data_ptr(), wptrs[i])

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    launches_per_step = NH * dex[0].last_launches()  # kernels only (8 pyr + 2 fast + quadtree + describe per handle)
    # one single-handle pass so that `ex` holds the whole batch for the per-stage timings below
    ex.extract_batch_device(d_img.data_ptr(), B, W, H, W, W * H, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                            d_n.data_ptr(), sptr)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for ws in wstreams:
        ws.wait_event(e0)
    for _ in range(args.steps):
        step_device()
    for ws in wstreams:
        done = torch.cuda.Event()
        done.record(ws)
        stream.wait_event(done)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    n_kp = int(d_n.sum().item())

    # ---- per-stage device times (events around each stage on the same stream)
    stage_ms = {}
    for name, mask in (("pyramid_fast_blur", _lib.STAGE_PYRAMID), ("nms", _lib.STAGE_NMS),
                       ("octree", _lib.STAGE_OCTREE), ("describe", _lib.STAGE_DESCRIBE)):
        reps = max(3, min(args.steps, 10))
        ex.run_stage(mask, B, sptr)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            ex.run_stage(mask, B, sptr)
        b.record()
        torch.cuda.synchronize()
        stage_ms[name] = a.elapsed_time(b) / reps

    # ---- end-to-end arm: pinned host frames in, host keypoints/descriptors out, two handles
    nslot = 8
    eb = min(B, 32)
    exs = [ORBextractor(NFEAT, 1.2, 8, 20, 7, device=local_rank, max_batch=eb) for _ in range(nslot)]
    h_img = torch.from_numpy(frames).pin_memory()
    h_np = h_img.numpy()
    outs = []
    for _ in range(nslot):
        k = torch.empty((eb, cap, 7), dtype=torch.float32).pin_memory()
        d = torch.empty((eb, cap, 32), dtype=torch.uint8).pin_memory()
        n = torch.zeros(eb, dtype=torch.int32).pin_memory()
        outs.append((k.numpy().view(KP_DTYPE).reshape(eb, cap), d.numpy(), n.numpy(), (k, d, n)))
    chunks = [(i, min(eb, B - i)) for i in range(0, B, eb)]

    def run_e2e(steps):
        """`steps` passes over the B pinned host frames as one stream of chunks: each chunk is uploaded,
        extracted and its keypoints/descriptors/counts downloaded; a slot is synchronised (and its result
        consumed on the host) only when it is reused, and everything is drained at the end."""
        total = 0
        pending = [None] * nslot
        ci = 0
        for _ in range(steps):
            for f0, nb in chunks:
                s = ci % nslot
                ci += 1
                if pending[s] is not None:
                    exs[s].sync()
                    total += int(outs[s][2][:pending[s]].sum())
                exs[s].extract_batch_async(h_np[f0:f0 + nb], (outs[s][0][:nb], outs[s][1][:nb], outs[s][2][:nb]))
                pending[s] = nb
        for s in range(nslot):
            if pending[s] is not None:
                exs[s].sync()
                total += int(outs[s][2][:pending[s]].sum())
        return total

    run_e2e(args.warmup)
    barrier()
    t0 = time.perf_counter()
    kp_e2e = run_e2e(args.steps)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    assert kp_e2e == n_kp * args.steps, (kp_e2e, n_kp, args.steps)  # the streamed path produced the same keypoints
    clocks = sampler.stop()

    # ---- Hamming matches/s: server-side place recognition shard (BASELINE config 5, scaled to one step):
    # 2000 query descriptors against this rank's shard of the keyframe-descriptor database, top-2 per
    # query, then the all-gather + merge of the per-shard candidates when world > 1.
    from swarmmap_b200 import place
    NQ, DPK, KF_PER_RANK = 2000, 256, 8192          # 2.1 M descriptors (67 MB) per GPU
    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + rank)
    db = torch.randint(0, 256, (KF_PER_RANK * DPK, 32), dtype=torch.uint8, device=dev, generator=gen)
    qd = torch.randint(0, 256, (NQ, 32), dtype=torch.uint8, device=dev, generator=gen)
    shard = place.PlaceShard(db, DPK, rank * KF_PER_RANK, device=local_rank)
    for _ in range(2):
        shard.query(qd, 2, 50)
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    h0.record()
    for _ in range(reps):
        keys, votes = shard.query(qd, 2, 50)
    h1.record()
    barrier()
    ms_ham = h0.elapsed_time(h1) / reps

    # ---- single-frame latency of the drop-in operator() path and per-call matcher times (rank 0, informational)
    extras = {}
    if rank == 0:
        from swarmmap_b200.matcher import Frame, ORBmatcher
        ex1 = ORBextractor(NFEAT, 1.2, 8, 20, 7, device=local_rank, max_batch=1)
        for _ in range(5):
            ex1(frames[0])
        t0 = time.perf_counter()
        for i in range(50):
            ex1(frames[i % B])
        extras["single_frame_operator_ms"] = (time.perf_counter() - t0) / 50 * 1e3
        # KITTI-shaped extraction throughput (north_star's second frame size), device-resident, B=128
        kb = 128
        kf = synth.make_batch(kb, 1241, 376, 20220405)
        exk = ORBextractor(2000, 1.2, 8, 20, 7, device=local_rank, max_batch=kb)
        kcap = exk.max_keypoints()
        dk_img = torch.from_numpy(kf).to(dev)
        dk_kps = torch.empty((kb, kcap, 7), dtype=torch.float32, device=dev)
        dk_desc = torch.empty((kb, kcap, 32), dtype=torch.uint8, device=dev)
        dk_n = torch.zeros(kb, dtype=torch.int32, device=dev)
        def kstep():
            exk.extract_batch_device(dk_img.data_ptr(), kb, 1241, 376, 1241, 1241 * 376, dk_kps.data_ptr(),
                                     dk_desc.data_ptr(), kcap, dk_n.data_ptr(), sptr)
        for _ in range(3):
            kstep()
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(10):
            kstep()
        k1.record()
        torch.cuda.synchronize()
        extras["kitti_1241x376_2000feat_frames_per_s"] = 10 * kb / (k0.elapsed_time(k1) * 1e-3)
        extras["kitti_keypoints_per_frame"] = float(dk_n.float().mean().item())
        del exk, dk_img, dk_kps, dk_desc
        seq = synth.make_sequence(3, 1241, 376, 20220405)
        ex4k = ORBextractor(4000, 1.2, 8, 20, 7, device=local_rank, max_batch=1)
        fs = [Frame.from_keypoints(*ex4k(img), 1241, 376, ex4k.GetScaleFactors()) for img in seq]
        m = ORBmatcher(0.9, True, device=local_rank)
        prev0 = np.stack([fs[0].x, fs[0].y], 1).astype(np.float32)
        for _ in range(3):
            m.SearchForInitialization(fs[0], fs[1], prev0.copy(), 100)
        t0 = time.perf_counter()
        for i in range(20):
            nm, _ = m.SearchForInitialization(fs[0], fs[1 + i % 2], prev0.copy(), 100)
        extras["search_for_initialization_ms"] = (time.perf_counter() - t0) / 20 * 1e3
        extras["search_for_initialization_cfg"] = f"KITTI-shaped 1241x376, {fs[0].N} x {fs[1].N} keypoints, window 100, host arrays in/out, {nm} matches"
        sf = ex4k.GetScaleFactors()
        u, v = fs[0].x.copy(), fs[0].y.copy()
        valid = np.ones(fs[0].N, np.uint8)
        fs[1].mvScaleFactors = sf
        for _ in range(3):
            m.SearchByProjectionLastFrame(fs[1], fs[0], u, v, valid, 15)
        t0 = time.perf_counter()
        for i in range(20):
            m.SearchByProjectionLastFrame(fs[1], fs[0], u, v, valid, 15)
        extras["search_by_projection_ms"] = (time.perf_counter() - t0) / 20 * 1e3
        # tracking step with the frame kept on the device: operator() on a host image,
        # undistort + grid on the GPU, SearchByProjection against the resident frame -- versus the same step through
        # host arrays (keypoints back to the host, Frame built there, uploaded again by the matcher)
        from swarmmap_b200.matcher import Camera, ResidentFrame
        cam = Camera(718.856, 718.856, 607.1928, 185.2157, -0.2834, 0.0739, 0.0002, 1.76e-05, 0.0)
        bounds = cam.bounds(1241, 376, local_rank)
        rf = ResidentFrame(local_rank)
        def track_resident(img):
            ex4k(img)
            rf.from_extractor(ex4k, 0, cam, bounds)
            return m.SearchByProjectionLastFrame(rf, fs[0], u, v, valid, 15)
        def track_host(img):
            f = Frame.from_keypoints(*ex4k(img), 1241, 376, sf)
            return m.SearchByProjectionLastFrame(f, fs[0], u, v, valid, 15)
        for fn, key in ((track_resident, "track_step_resident_ms"), (track_host, "track_step_host_arrays_ms")):
            for _ in range(3):
                fn(seq[1])
            t0 = time.perf_counter()
            for i in range(20):
                fn(seq[1 + i % 2])
            extras[key] = (time.perf_counter() - t0) / 20 * 1e3
        extras["track_step_cfg"] = ("1241x376, 4000 features: operator() from a host image + frame build + "
                                    "SearchByProjection(th=15) of 4000 last-frame points; resident = undistort/grid on the "
                                    "GPU and match in place, host_arrays = no undistortion, Frame rebuilt on the host")
        # DBoW2 transform: Frame::ComputeBoW on the extractor's descriptors, synthetic
        # k = 10, L = 5 vocabulary in the ORBvoc.bin layout (the real file is not shipped), levelsup = 4
        from swarmmap_b200.bow import ORBVocabulary
        vblob = synth.make_vocabulary(10, 5, seed=20220407)
        voc = ORBVocabulary(vblob, device=local_rank)
        exb = ORBextractor(NFEAT, 1.2, 8, 20, 7, device=local_rank, max_batch=64)
        bk, bd, bn = exb.extract_batch(frames[:64])
        for _ in range(2):
            voc.transform_batch(bd, bn, 4)
        t0 = time.perf_counter()
        for _ in range(5):
            voc.transform_batch(bd, bn, 4)
        t_b = (time.perf_counter() - t0) / 5
        one = bd[0, :bn[0]]
        for _ in range(3):
            voc.transform(one, 4)
        t0 = time.perf_counter()
        for _ in range(20):
            voc.transform(one, 4)
        extras["bow_transform"] = {"frames_per_s_batch64": 64 / t_b, "ms_single_frame": (time.perf_counter() - t0) / 20 * 1e3,
                                   "features_per_frame": float(bn.mean()),
                                   "cfg": f"synthetic vocabulary k=10 L=5 ({voc.n_nodes} nodes, {voc.n_words} words), "
                                          "host descriptors in, BowVector + FeatureVector out, levelsup 4"}
        if not args.no_cpu_baseline:
            import oracle_lib
            ov = oracle_lib.Vocabulary(vblob)
            extras["bow_transform"]["cpu_oracle_ms_single_frame"] = ov.time_transform(one, 4, 20) / 20 * 1e3
            ref_ms = _reference_dbow2_ms(vblob, one, 4, 20)
            if ref_ms is not None:  # the reference's own DBoW2 (oracle/_ref, built from /root/reference where present)
                extras["bow_transform"]["cpu_reference_dbow2_ms_single_frame"] = ref_ms
        if not args.no_cpu_baseline:
            import oracle_lib
            t0 = time.perf_counter()
            for i in range(5):
                oracle_lib.search_for_initialization(fs[0], fs[1 + i % 2], prev0, 100, 0.9, True)
            extras["search_for_initialization_cpu_oracle_ms"] = (time.perf_counter() - t0) / 5 * 1e3

    # ---- reduce over ranks (max time), rank 0 prints
    t = torch.tensor([ms_total, t_e2e * 1e3, stage_ms["pyramid_fast_blur"] + stage_ms["nms"], ms_ham],
                     dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(n_kp)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total, ms_e2e, ms_pf, ms_ham = [float(x) for x in t.tolist()]
    frames_total = world * B * args.steps
    value = frames_total / (ms_total * 1e-3)
    e2e_value = frames_total / (ms_e2e * 1e-3)
    peak, peak_src = _peaks()
    achieved = PYR_FAST_BYTES * B / (ms_pf * 1e-3) / 1e9
    h2d = B * W * H
    d2h = B * (cap * (28 + 32) + 4)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        v, n = cpu_oracle_rate(frames[:8], 12.0, 1)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{n} frames of the same workload, single thread, oracle -O2 (reference extractor is single-threaded per agent)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frame": [W, H], "nfeatures": NFEAT, "batch_per_gpu": B,
                       "agents": world, "handles_per_gpu": NH, "cache": "working set per step (frames + 3 plane sets) "
                       f"{(B * (W * H + 3 * 1.45e6)) / 1e6:.0f} MB > 126 MB L2, no reuse across steps"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": f"pinned host buffers, {nslot} extractor handles x {eb}-frame chunks in flight"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (_traffic_per_frame() * B if _traffic_per_frame() else None),
                         "traffic_source": "ncu dram__bytes_read+write, same 10 launches at B=256 (profiles/r1k_traffic.json)",
                         "peak_source": peak_src,
                         "kernel": "pyr_kernel x8 (pyramid+border, blur fused) + fast_kernel x2 (FAST score+tile retry+NMS)",
                         "bytes_per_frame": PYR_FAST_BYTES, "ms_per_launch_set": ms_pf,
                         "frac_counting_fused_blur_bytes": (PYR_FAST_BYTES + BLUR_BYTES) * B / (ms_pf * 1e-3) / 1e9 / peak},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "stages_ms_per_step": stage_ms,
            "keypoints_per_frame": float(cnt.item()) / (world * B),
            "latency": extras,
            "hamming": {"metric": "hamming_matches_per_sec", "value": world * NQ * KF_PER_RANK * DPK / (ms_ham * 1e-3),
                        "unit": "256-bit pairs/s", "ms_per_query_batch": ms_ham,
                        "kernel": "db_top2_umma_kernel (tcgen05 kind::i8, A in TMEM, accumulators in TMEM) + db_merge_kernel",
                        "roofline": {"bound": "tensor", "unit": "TOP/s (int8)",
                                     "achieved": NQ * KF_PER_RANK * DPK * 512 / (ms_ham * 1e-3) / 1e12,
                                     "peak": 4500.0,
                                     "frac": NQ * KF_PER_RANK * DPK * 512 / (ms_ham * 1e-3) / 1e12 / 4500.0,
                                     "peak_source": "nominal dense int8 per GPU (MEASURED_PEAKS.json has no int8 figure); "
                                                    "algorithmic ops = 2 x 256 per pair, the kernel issues 2 x 288 "
                                                    "(constant K block); per GPU, timed with CUDA events incl. merge"},
                        "config": f"{NQ} queries x {KF_PER_RANK * DPK} descriptors per GPU shard ({KF_PER_RANK} keyframes x {DPK}), "
                                  f"top-2 + votes" + (", all_gather of 32 KB key blocks + merge" if world > 1 else "")},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
