#!/usr/bin/env python3
"""Per-kernel SASS instruction counts of libswm_orb.so (cuobjdump -sass), for profiles/: total instructions and the
mnemonics that prove the Blackwell paths (UTCIMMA / LDTM / STTM / UTCBAR = tcgen05 + TMEM, UBLKCP / UTMALDG = TMA,
SYNCS = mbarrier, IDP = dp2a/dp4a, VIMNMX / VIADDMNMX = DPX 16x2 SIMD, POPC, REDUX)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "swarmmap_b200", "libswm_orb.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCIMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "IDP", "VIMNMX", "VIADDMNMX", "VABSDIFF4",
        "POPC", "REDUX", "IMMA", "LDG", "STG", "LDS", "STS", "BAR"]
name, counts = None, {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and name:
        op = m.group(1)
        counts[name]["total"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[name][k] += 1
print(f"# cuobjdump -sass {os.path.basename(lib)}: instruction counts per kernel (sm_100a)")
for fn in sorted(counts, key=lambda f: -counts[f]["total"]):
    dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
    c = counts[fn]
    rest = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
    print(f"{dem:60s} total={c['total']:6d} {rest}")
