"""Key per-launch metrics of an .ncu-rep (run here, no GPU): python tools/ncu_summary.py gpurun_out/X.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("smsp__inst_executed.sum", "winst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit"), ("lts__t_sector_hit_rate.pct", "l2hit"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wf%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"), ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uni%"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu_c%"), ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_c%"),
        ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "fmah_c%")]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    out = []
    for k, n in want:
        if k in hdr:
            v = r[hdr.index(k)]
            try: v = f"{float(v):.4g}"
            except ValueError: v = v[:36]
            out.append(f"{n}={v}")
    print(" ".join(out))
    st = sorted(((float(r[hdr.index(h)]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stalls if r[hdr.index(h)]), reverse=True)[:6]
    print("    stalls/issue:", " ".join(f"{n}={v:.2f}" for v, n in st))
