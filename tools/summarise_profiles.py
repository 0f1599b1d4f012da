#!/usr/bin/env python
"""Turns the raw captures of one round into the committed summaries under profiles/.

usage: summarise_profiles.py ROUND_TAG LAUNCH_CSV NCU_REP [WORKLOAD_NOTE]
  LAUNCH_CSV  ncu --metrics gpu__time_duration.sum --clock-control none --csv launch list of `bench.py --profile-only --steps 2 --warmup 1`
  NCU_REP     ncu --set full --clock-control none --import-source on capture of the main kernels of one step
writes profiles/<tag>_launches.csv, <tag>_launch_summary.txt, <tag>_ncu_summary.txt, <tag>_traffic.json
(tools/gpu_profile.sh produces the two inputs on the GPU box)
"""
import csv, io, json, os, re, shutil, subprocess, sys
tag, launch_csv, rep = sys.argv[1:4]
note = sys.argv[4] if len(sys.argv) > 4 else "workload C5, 1024 bf16 images"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prof = os.path.join(root, "profiles")
shutil.copy(launch_csv, os.path.join(prof, f"{tag}_launches.csv"))
out = subprocess.run([sys.executable, os.path.join(root, "tools", "launch_summary.py"), launch_csv], capture_output=True, text=True).stdout
open(os.path.join(prof, f"{tag}_launch_summary.txt"), "w").write(out)

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
def to_bytes(v, u):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
txt = [f"# ncu --set full --clock-control none, one step of bench.py ({note}); {tag}"]
traffic = {}
ni = h.index("Kernel Name")
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ni]).replace("void ", "").replace("<unnamed>::", "")
    txt.append("----")
    txt.append(f"  {'Kernel Name':66s} {r[ni][:110]}")
    for k in keys:
        if k in h:
            i = h.index(k)
            txt.append(f"  {k:66s} {r[i]} {units[i]}")
    rd, wr = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    t = h.index("gpu__time_duration.sum")
    traffic.setdefault(name, []).append({"dram_read_bytes": to_bytes(r[rd], units[rd]), "dram_write_bytes": to_bytes(r[wr], units[wr]),
                                         "duration": f"{r[t]} {units[t]}"})
seen = " ".join(r[ni] for r in rows[2:])
for rx in [x for x in ("image_grad_staged", "image_grad_quad", "sample_fwd_tiled", "ot_solve", "head_gemm_tma") if x in seen]:
    for by in ("", "1"):
        env = dict(os.environ, **({"BY_INST": "1"} if by else {}))
        o = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_hot_lines.py"), rep, rx, "0", "14"], capture_output=True, text=True, env=env).stdout
        txt.append(f"## hot lines ({'by instructions' if by else 'by stall samples'}): {rx}")
        txt += [l[:170] for l in o.splitlines()]
open(os.path.join(prof, f"{tag}_ncu_summary.txt"), "w").write("\n".join(txt) + "\n")
json.dump({"source": f"ncu --set full --clock-control none, profiles/{tag}_ncu_summary.txt; per launch, {note}, 1 GPU",
           "kernels": traffic}, open(os.path.join(prof, f"{tag}_traffic.json"), "w"), indent=1)
print("\n".join(txt[:60]))
