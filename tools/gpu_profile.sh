#!/bin/bash
# The round's committed profiles: launch list + one `ncu --set full` capture of the main kernels, for C5 (bf16) and C3 (fp32).
# usage: gpu_profile.sh TAG [WORKLOADS]   (one workload per gpurun call: a --set full report is ~35 MB and gpurun_out/ returns <= 64 MiB)
TAG=${1:-r02}; WL=${2:-"C5 C3"}
mkdir -p gpurun_out
for W in $WL; do
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_$W.csv \
      python bench.py --workload $W --profile-only --steps 2 --warmup 1 > gpurun_out/${TAG}_launch_run_$W.log 2>&1
  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'sample_fwd|image_grad|head_gemm|ot_solve|cost_hist|split' -c 9 -f -o gpurun_out/prof_${TAG}_$W \
      python bench.py --workload $W --profile-only --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_run_$W.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu_run_$W.log
done
ls -la gpurun_out/prof_${TAG}_*
