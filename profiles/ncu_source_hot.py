#!/usr/bin/env python
"""Hottest CUDA source lines of a captured kernel: ncu_source_hot.py file.ncu-rep [N]   (needs -lineinfo + --import-source on)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, acc = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if cur and hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":   # a source-line summary row
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        try:
            acc.append((int(d["Instructions Executed"]), int(d["# Samples"] or 0), float(d["Avg. Threads Executed"] or 0), cur, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(a[0] for a in acc) or 1
smp = sum(a[1] for a in acc) or 1
acc.sort(reverse=True)
print("total warp-instructions %d, stall samples %d" % (tot, smp))
print("%10s %6s %6s %5s  %s" % ("inst", "inst%", "smp%", "thr", "source"))
for a in acc[:top]:
    print("%10d %5.1f%% %5.1f%% %5.1f  %s:%d  %s" % (a[0], 100 * a[0] / tot, 100 * a[1] / smp, a[2], a[3], a[4], a[5]))
