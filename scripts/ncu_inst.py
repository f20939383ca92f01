#!/usr/bin/env python
"""Top source lines of an ncu report by executed warp instructions.  usage: ncu_inst.py report.ncu-rep [topN]"""
import csv, os, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname = "?"; hdr = None; items = []; tot = 0
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; ci = {}
        for i, h in enumerate(hdr): ci.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue
    try: n = int(r[ci["Instructions Executed"]]); t = int(r[ci["Thread Instructions Executed"]])
    except ValueError: continue
    tot += n; items.append((n, t, fname, r[0], r[1].strip()))
items.sort(key=lambda x: -x[0])
print("total warp instr", tot)
for n, t, f, l, src in items[:top]:
    print("%5.2f%% lanes=%4.1f %s:%s %s" % (100 * n / tot, t / max(n, 1), f[:12], l, src[:100]))
