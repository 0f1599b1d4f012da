#!/usr/bin/env python
"""SASS evidence of the Blackwell-native instructions per kernel of libfairguide.so (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP, mbarrier -> SYNCS, warp reductions -> REDUX, 3-input integer min -> VIMNMX3).
usage: sass_evidence.py ROUND_TAG   ->  profiles/<tag>_sass_{head,sample,image_grad,solver,peer}.txt  (runs without a GPU)"""
import collections, os, re, subprocess, sys
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "finetune-fair-diffusion_b200", "csrc", "libfairguide.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        funcs[cur].append(line)
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
PAT = re.compile(r"\b(STG\.E\.STRONG\.SYS|LDG\.E\.STRONG\.SYS|MEMBAR\.ALL\.SYS|UTC\w*MMA\w*|UTCBAR\w*|UTCATOM\w*|LDTM\w*|STTM\w*|UTMALDG\w*|UTMASTG\w*|UBLKCP\w*|SYNCS\w*|REDUX\w*|VIMNMX3\w*|HMMA\w*|LDGSTS\w*|UTMAPF\w*)")
groups = {"head": ["head_gemm", "head_", "split_tf32", "face_search_tc"], "sample": ["sample_fwd"], "image_grad": ["image_grad"],
          "solver": ["ot_solve", "cost_hist", "ot_targets", "compact"], "peer": ["peer_"]}
for g, keys in groups.items():
    out = [f"# cuobjdump -sass libfairguide.so (architectures in the library: {', '.join(arch)}), kernels matching {keys}",
           "# per kernel: instruction count, then every Blackwell-specific / async / tensor mnemonic with its count and one example line"]
    for name, lines in funcs.items():
        d = demangle(name)
        if not any(k in d for k in keys):
            continue
        cnt, ex = collections.Counter(), {}
        for l in lines:
            m = PAT.search(l)
            if m:
                op = m.group(1).split(".")[0]
                cnt[op] += 1
                ex.setdefault(op, re.sub(r"\s+", " ", l.split("*/", 1)[1]).strip()[:110])
        out.append("")
        out.append(re.sub(r"\(anonymous namespace\)::", "", d)[:200])
        out.append(f"    {len(lines)} SASS instructions ({len(lines) * 16 // 1024} KB)")
        for op, c in cnt.most_common():
            out.append(f"    {op:14s} x{c:<5d} e.g. {ex[op]}")
    open(os.path.join(root, "profiles", f"{tag}_sass_{g}.txt"), "w").write("\n".join(out) + "\n")
    print(g, sum(1 for l in out if l and not l.startswith((" ", "#"))), "kernels")
