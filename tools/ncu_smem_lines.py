#!/usr/bin/env python
"""Shared-memory wavefronts per CUDA source line of one kernel in an .ncu-rep.  usage: ncu_smem_lines.py REPORT KERNEL_REGEX [top]"""
import csv, io, os, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{rx}",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
data, cur, h = [], "", None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1]) if len(r) > 1 else ""; continue
    if r[0] == "Line No":
        h = r; iw, ii, ie = h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Ideal"), h.index("L1 Wavefronts Shared Excessive"); continue
    if h is None or not r[0].isdigit() or len(r) <= iw:
        continue
    try:
        data.append((int(r[iw] or 0), int(r[ii] or 0), int(r[ie] or 0), f"{cur}:{r[0]}", r[1].strip()[:100]))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1
print(f"# shared wavefronts={tot} ideal={sum(d[1] for d in data)} excessive={sum(d[2] for d in data)}")
for w, i, e, loc, text in sorted(data, reverse=True)[:top]:
    print(f"{w:10d} {100 * w / tot:5.1f}%  ideal={i:10d} excess={e:9d}  {loc:30s} {text}")
