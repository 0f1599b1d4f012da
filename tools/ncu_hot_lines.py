#!/usr/bin/env python
"""Hot CUDA source lines of one kernel from an .ncu-rep (needs -lineinfo and --import-source on).

usage: ncu_hot_lines.py REPORT KERNEL_REGEX [launch_skip] [top_n]
Columns: stall samples (share), warp instructions executed (share), file:line, source text.
"""
import csv
import io
import os
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 35
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{rx}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
data, cur_file, h = [], "", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]) if len(r) > 1 else ""
        continue
    if r[0] == "Line No":
        h = r
        si, ie = h.index("# Samples"), h.index("Instructions Executed")
        continue
    if h is None or len(r) <= ie or not r[0].isdigit():
        continue
    try:
        data.append((int(r[si] or 0), int(r[ie] or 0), f"{cur_file}:{r[0]}", r[1].strip()[:110]))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1
toti = sum(d[1] for d in data) or 1
print(f"# samples={tot} warp-instructions={toti}")
key = (lambda d: d[1]) if os.environ.get("BY_INST") else (lambda d: d[0])
for s, n, loc, text in sorted(data, key=key, reverse=True)[:top]:
    print(f"{s:7d} {100 * s / tot:5.1f}%  inst={n:9d} {100 * n / toti:5.1f}%  {loc:26s} {text}")
