"""Executed warp instructions between barrier instructions of one kernel (SASS order): python scripts/ncu_segments.py report.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[1], rows[2:]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
seg, smp, cnt, tot = 0, 0, 0, sum(int(r[iE] or 0) for r in data)
for k, r in enumerate(data):
    e = int(r[iE] or 0)
    seg += e; smp += int(r[iSm] or 0); cnt += 1
    s = r[iS]
    if "BAR" in s or "EXIT" in s or "ARRIVE" in s.upper() or "SYNCS" in s:
        print(f"{k:5d} {cnt:5d} sass  {seg:12d} ({100*seg/tot:5.1f}%) samples {smp:6d}   {s.strip()[:70]}")
        seg = smp = cnt = 0
