#!/usr/bin/env python
"""Per-source-line summary of an `ncu --set full --import-source on` capture.
Usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > s.csv; python profiles/srclines.py s.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = ""
hdr = None
data = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 10 or r[2] != "-":   # only the per-source-line aggregate rows
        continue
    try:
        d = dict(zip(hdr[4:], r[4:]))
        data.append((cur_file, int(r[0]), r[1], int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0),
                     int(d.get("stall_long_sb", 0) or 0), int(d.get("stall_wait", 0) or 0),
                     int(d.get("stall_math", 0) or 0), int(d.get("stall_lg", 0) or 0), int(d.get("stall_short_sb", 0) or 0)))
    except Exception:
        pass
ts = sum(d[3] for d in data) or 1
ti = sum(d[4] for d in data) or 1
print("total samples %d, warp instructions %d" % (ts, ti))
print("%-14s %5s %7s %7s  %6s %6s %6s %6s %6s  %s" % ("file", "line", "samp%", "inst%", "longsb", "wait", "math", "lg", "shortsb", "source"))
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print("%-14s %5d %6.2f%% %6.2f%%  %6d %6d %6d %6d %6d  %s" % (d[0][:14], d[1], 100 * d[3] / ts, 100 * d[4] / ti, d[5], d[6], d[7], d[8], d[9], d[2].strip()[:90]))
