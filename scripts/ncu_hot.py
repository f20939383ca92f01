#!/usr/bin/env python
"""Summarise an ncu report's source page by CUDA source line: top lines by stall samples.  usage: ncu_hot.py report.ncu-rep [topN]"""
import csv, os, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname = "?"
hdr = None
items = []
tot = 0
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1])
        continue
    if r[0] == "Line No":
        hdr = r
        ci = {}
        for i, h in enumerate(hdr):
            ci.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    try:
        s = int(r[ci["# Samples"]])
    except ValueError:
        continue
    tot += s
    stalls = {h[6:]: int(r[i] or 0) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
    items.append((s, fname, r[0], r[1].strip(), stalls, r[ci["Instructions Executed"]]))
items.sort(key=lambda x: -x[0])
print("total samples", tot)
for s, f, line, src, stalls, inst in items[:top]:
    topst = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
    print("%6.2f%% %s:%-4s inst=%-9s %-34s %s" % (100.0 * s / max(tot, 1), f[:14], line, inst, " ".join("%s:%d" % kv for kv in topst if kv[1]), src[:100]))
