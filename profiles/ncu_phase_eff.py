#!/usr/bin/env python
"""SIMT efficiency per source region of a captured kernel: ncu_phase_eff.py file.ncu-rep FILE:LO-HI=name ...
Sums warp-instructions and thread-instructions of the source lines in each region (lines outside every region -> 'other')."""
import csv
import subprocess
import sys

rep = sys.argv[1]
regions = []
for a in sys.argv[2:]:
    loc, name = a.split("=")
    f, r = loc.split(":")
    lo, hi = r.split("-")
    regions.append((f, int(lo), int(hi), name))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur, hdr = None, None
acc = {}
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if cur and hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        line = int(r[0])
        name = "other(%s)" % cur
        for f, lo, hi, nm in regions:
            if f == cur and lo <= line <= hi:
                name = nm
                break
        a = acc.setdefault(name, [0, 0, 0])
        a[0] += int(d["Instructions Executed"] or 0)
        a[1] += int(d["Thread Instructions Executed"] or 0)
        a[2] += int(d["# Samples"] or 0)
tot = sum(a[0] for a in acc.values()) or 1
smp = sum(a[2] for a in acc.values()) or 1
print("%-34s %12s %6s %6s %8s" % ("region", "warp-inst", "inst%", "smp%", "thr/inst"))
for name, a in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print("%-34s %12d %5.1f%% %5.1f%% %8.1f" % (name, a[0], 100 * a[0] / tot, 100 * a[2] / smp, a[1] / max(1, a[0])))
