"""Per-source-line totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` output
(instructions executed and stall samples), to find where a kernel's issue slots go."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[2]
inst = collections.Counter(); samp = collections.Counter(); src = {}
line = None
for r in rows[3:]:
    if r and r[0].strip().isdigit():
        line = int(r[0]); src[line] = ",".join(r[1:-4])[:110] if len(r) > 6 else r[1][:110]
        continue
    if line is None or len(r) < 8:
        continue
    try:
        inst[line] += int(r[7]); samp[line] += int(r[4])
    except ValueError:
        pass
tot_i = sum(inst.values()); tot_s = sum(samp.values())
print(f"total warp instructions {tot_i}, stall samples {tot_s}")
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 10**9)
for ln in sorted(inst):
    if lo <= ln <= hi and (inst[ln] > 0.004 * tot_i or samp[ln] > 0.004 * tot_s):
        print(f"{ln:5d} {100*inst[ln]/tot_i:5.1f}%i {100*samp[ln]/max(tot_s,1):5.1f}%s  {src[ln]}")
