#!/usr/bin/env python
"""Summarise `-Xptxas -v` logs: registers, spills, stack, static smem per kernel."""
import re, sys, glob, os
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?")
for f in sorted(glob.glob(os.path.join(sys.argv[1] if len(sys.argv) > 1 else ".", "*.ptxas.log"))):
    for m in pat.finditer(open(f).read()):
        name = re.sub(r"_ZN\d+_GLOBAL__N__\w+?_cu_\w{8}", "", m.group(1))
        if len(sys.argv) > 2 and not re.search(sys.argv[2], name):
            continue
        print(f"{name[:70]:72s} stack={m.group(2):>4s} spill={m.group(3):>3s}/{m.group(4):>3s} regs={m.group(5):>3s} smem={m.group(7)}")
