#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares from an ncu report:
    python tools/ncu_hotspots.py report.ncu-rep [kernel-substring] [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
agg = collections.OrderedDict()
cur_file, hdr, fn, use = None, None, "", True


def num(x):
    try:
        return int(x)
    except (TypeError, ValueError):
        return 0


for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        fn = r[1]
        use = filt in fn
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not use:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr, r))
    a = agg.setdefault((cur_file, line), [0, 0, r[1]])
    a[0] += num(d.get("Instructions Executed"))
    a[1] += num(d.get("# Samples"))
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[1] for v in agg.values()) or 1
print("total warp instructions", tot_i, "samples", tot_s)
for (f, l), (i, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]:
    print(f"{f}:{l:4d} inst {100 * i / tot_i:5.1f}% samples {100 * s / tot_s:5.1f}%  {src.strip()[:100]}")
