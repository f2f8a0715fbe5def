#!/usr/bin/env python3
"""CPU baseline of the hot path per BASELINE.md section 3 / SURVEY.md 8(d): the oracle (the restated reference path;
the reference itself cannot be built as a whole) timed on this box's host cores.

  * extraction at 752x480 / 1000 features: A agents on A pinned threads (one extractor each, like one Tracking thread
    per agent), per-frame wall time -> median and p95 ms, aggregate frames/s;
  * each matcher, single thread: SearchForInitialization (1241x376, 4000 features, window 100),
    SearchByProjection(cur, last, 15), SearchByProjection(F, local map points, th 1), SearchByBoW(KF, F);
  * SWM_ORACLE_VARIANT=O1 in the environment selects the -O1 build (the reference's release level).

Prints one JSON object.  Only bench.py's cpu_baseline leg runs this (test infrastructure, never the product)."""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(round(q * (len(v) - 1))))]


def extract_agents(oracle, frames, agents, seconds, cores):
    exs = [oracle.Extractor(1000, 1.2, 8, 20, 7) for _ in range(agents)]
    times = [[] for _ in range(agents)]
    start = threading.Barrier(agents)

    def run(a):
        if cores:
            try:
                os.sched_setaffinity(0, {cores[a % len(cores)]})
            except OSError:
                pass
        exs[a](frames[a % len(frames)])  # warm
        start.wait()
        t_end = time.perf_counter() + seconds
        i = a
        while time.perf_counter() < t_end:
            t0 = time.perf_counter()
            exs[a](frames[i % len(frames)])
            times[a].append(time.perf_counter() - t0)
            i += agents
    th = [threading.Thread(target=run, args=(a,)) for a in range(agents)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    wall = time.perf_counter() - t0
    allt = [x for v in times for x in v]
    return {"agents": agents, "frames": len(allt), "frames_per_s": len(allt) / wall,
            "ms_median": 1e3 * statistics.median(allt), "ms_p95": 1e3 * pct(allt, 0.95)}


def timed(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return {"ms_median": 1e3 * statistics.median(ts), "ms_p95": 1e3 * pct(ts, 0.95), "reps": reps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", default="1,2,4,8")
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--matchers", type=int, default=1)
    args = ap.parse_args()
    import oracle_lib as oracle
    from swarmmap_b200 import synth
    from swarmmap_b200.matcher import FeatureVector, Frame
    try:
        cores = sorted(os.sched_getaffinity(0))
    except AttributeError:
        cores = []
    model = ""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    out = {"variant": os.environ.get("SWM_ORACLE_VARIANT", "O2"), "nproc": os.cpu_count(), "cores_available": len(cores),
           "cpu_model": model, "extract": [], "matchers": {}}
    frames = synth.make_batch(16, 752, 480, 20220410)
    for a in [int(x) for x in args.agents.split(",") if x]:
        if cores and a > len(cores):
            continue
        out["extract"].append(extract_agents(oracle, frames, a, args.seconds, cores))
    if args.matchers:
        sf = oracle.scale_tables(1.2, 8)[0]
        ex4 = oracle.Extractor(4000, 1.2, 8, 20, 7)
        seq = synth.make_sequence(2, 1241, 376, 20220405)
        k = [Frame.from_keypoints(*ex4(im), 1241, 376, sf) for im in seq]
        prev = np.stack([k[0].x, k[0].y], 1).astype(np.float32)
        out["matchers"]["SearchForInitialization_1241x376_4000"] = timed(
            lambda: oracle.search_for_initialization(k[0], k[1], prev, 100, 0.9, True), 8)
        ex1 = oracle.Extractor(1000, 1.2, 8, 20, 7)
        seq = synth.make_sequence(2, 752, 480, 20220406)
        e = [Frame.from_keypoints(*ex1(im), 752, 480, sf) for im in seq]
        last, cur = e
        ones = np.ones(last.N, np.uint8)
        rad = (np.float32(15) * sf[last.octave]).astype(np.float32)
        out["matchers"]["SearchByProjection_last_frame_752x480_1000"] = timed(
            lambda: oracle.match_window(cur, last.desc, last.x, last.y, rad, last.octave - 1, last.octave + 1, ones, ones,
                                        100, 0, 0.9, True, last.angle), 20)
        r2 = (np.float32(4.0) * sf[last.octave]).astype(np.float32)
        out["matchers"]["SearchByProjection_map_points_752x480_1000"] = timed(
            lambda: oracle.match_window(cur, last.desc, last.x, last.y, r2, last.octave - 1, last.octave, ones, ones,
                                        100, 1, 0.8, False), 20)
        fv = [FeatureVector((f.desc[:, 3].astype(np.int64) >> 2) % 64) for f in e]
        out["matchers"]["SearchByBoW_kf_frame_752x480_1000"] = timed(
            lambda: oracle.search_by_bow(last, fv[0], ones, cur, fv[1], None, 0, 0.7, True), 20)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
