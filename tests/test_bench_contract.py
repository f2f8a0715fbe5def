"""The bench.py JSON contract, checked on the committed lines (profiles/r2am_bench.json from `python bench.py` on one
B200, profiles/r2ae_bench_reference.json from `python bench.py --impl reference`), and the parts of bench.py that run
without a GPU (argument parsing, the reference arm on a tiny sample)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV_LINE = "r2am_bench.json"


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_device_arm_line_has_every_contract_key():
    d = _line(DEV_LINE)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    # BASELINE.json's metric is prose ("ORB frames/sec/GPU (1000 feat, 752x480) ...; Hamming matches/sec"): the line
    # carries the first as `value` (whole-job frames/s) and the second under "hamming"
    assert "ORB frames/sec" in base["metric"] and "Hamming matches/sec" in base["metric"]
    assert d["metric"] == "orb_extract_frames_per_sec" and d["unit"] == "frames/s"
    assert d["hamming"]["metric"] == "hamming_matches_per_sec" and d["hamming"]["value"] > 0
    assert d["config"]["frame"] == [752, 480] and d["config"]["nfeatures"] == 1000
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "u8" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    frames_per_step = d["config"]["batch_per_gpu"] * d["n_gpus"]
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] / frames_per_step - 1) < 0.02
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    assert r["peak"] == peaks["hbm_gbs"] and r["traffic"] is not None
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    k = d["clocks"]
    assert k["sm_mhz"] > 0 and k["sm_max_mhz"] >= k["sm_mhz"] and not set(k["reasons"]) & {
        "hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    h = d["hamming"]
    assert h["roofline"]["bound"] == "tensor" and 0 < h["roofline"]["frac"] < 1.2
    assert d["config"]["timed_region_s"] >= 1.0 and e["timed_region_s"] >= 1.0
    assert 0 < e["h2d_ceiling"]["e2e_frac_of_ceiling"] <= 1.05
    # one record per remaining BASELINE.json config, each with its own workload, device value, e2e and a parity check
    w = d["workloads"]
    k2 = w["config2_kitti"]
    assert k2["config"]["frame"] == [1241, 376] and k2["extract"]["value"] > 0 and k2["extract"]["e2e"]["value"] > 0
    assert k2["init_matching"]["value"] > 0 and k2["init_matching"]["parity"]["identical"] is True
    t = w["config34_tracking"]
    assert t["parity"]["identical"] is True
    for variant in ("projection_only", "with_bow"):
        v = t["variants"][variant]
        assert v["value"] >= 5000 and v["e2e"]["value"] >= 5000  # north_star's floor per GPU, with bit-exact matches
    p5 = w["config5_place"]
    assert "100 000 keyframes x 256" in p5["config"]["workload"] and p5["parity"]["votes_on_planted_keyframes"] is True
    assert p5["parity"]["oracle_sample"]["identical"] is True and p5["value"] > 1e12
    proto = d["cpu_baseline_protocol"]
    for variant in ("O2", "O1"):
        assert [x["agents"] for x in proto[variant]["extract"]][:2] == [1, 2] and proto[variant]["matchers"]


def test_reference_arm_line():
    d = _line("r2ae_bench_reference.json")
    dev = _line(DEV_LINE)
    assert d["impl"] == "reference" and d["metric"] == dev["metric"] and d["unit"] == dev["unit"]
    assert d["config"]["workload"] == dev["config"]["workload"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1


def test_reference_arm_runs_without_a_gpu():
    """`bench.py --impl reference` is the CPU oracle on the host cores: it must run on a box without a GPU."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0
