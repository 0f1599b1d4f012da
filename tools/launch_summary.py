#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the launches of the LAST
step (from the last select_expand_kernel on) in order, then totals per kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    name = re.sub(r"^void ", "", row["Kernel Name"]).replace("(anonymous namespace)::", "")
    name = re.sub(r"\(.*", "", name)[:64]
    v = float(row["Metric Value"]); u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    rows.append((name, v))
idx = [i for i, (s, _) in enumerate(rows) if "select_expand" in s]
start = idx[-1] if idx else 0
tot = sum(v for _, v in rows[start:])
print(f"# last step: {len(rows) - start} launches, {tot:.1f} us (cold-cache, serialised)")
agg = collections.OrderedDict()
for s, v in rows[start:]:
    a = agg.setdefault(s, [0, 0.0]); a[0] += 1; a[1] += v
for s, (c, v) in agg.items():
    print(f"{s:66s} n={c:3d} {v:10.1f} us {100 * v / tot:5.1f}%")
