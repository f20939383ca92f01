#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_summary.py launches.csv [split_by_grid]
With a second argument, launches of one kernel are also split by launch order inside a step (k_dense_level: one line per octree level)."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
tot = collections.OrderedDict()
seq = collections.defaultdict(list)
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    us = float(r[vi].replace(",", "")) / 1000.0
    name = r[ki].split("(")[0]
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += us
    seq[name].append(us)
allus = sum(t[1] for n, t in tot.items() if "u3d" in n)
for n, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-44s launches=%4d total_us=%10.1f avg_us=%9.1f share_of_u3d=%5.1f%%" % (n[:44], t[0], t[1], t[1] / t[0], 100.0 * t[1] / allus))
if len(sys.argv) > 2:
    for n, v in seq.items():
        if "k_dense_level" in n:
            per = int(sys.argv[2])
            for l in range(per):
                xs = v[l::per]
                print("  level %d: avg_us=%9.1f over %d launches" % (l + 1, sum(xs) / len(xs), len(xs)))
