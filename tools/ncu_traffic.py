"""DRAM traffic and per-launch summary of the roofline launch set from an ncu --set full capture (run here, no GPU):
    python tools/ncu_traffic.py gpurun_out/X.ncu-rep BATCH profiles/rN_traffic.json profiles/rN_extract_ncu_full.csv
The capture is `ncu --set full --clock-control none -k regex:"pyr_walk|fast_tile" -s 10 -c 10 python tools/prof_bench.py 256 2`
(the 8 pyramid + 2 FAST launches of the second step)."""
import csv, io, json, subprocess, sys
rep, batch, out_json, out_csv = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
cols = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
idx = [hdr.index(c) for c in cols]
def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
rd = wr = t = 0.0
with open(out_csv, "w", newline="") as f:
    wcsv = csv.writer(f)
    wcsv.writerow(cols)
    wcsv.writerow([units[i] for i in idx])
    for r in rows[2:]:
        wcsv.writerow([r[i] for i in idx])
        rd += to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr += to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        t += float(r[hdr.index("gpu__time_duration.sum")])
json.dump({"source": f"ncu --set full, B={batch} frames 752x480, pyr_walk_kernel x8 + fast_tile_kernel x2 ({out_csv})", "batch": batch,
           "launches": len(rows) - 2, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_frame": (rd + wr) / batch,
           "sum_gpu_time_us_under_ncu": t}, open(out_json, "w"), indent=1)
print(open(out_json).read())
