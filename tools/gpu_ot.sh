#!/bin/bash
# solver iteration: assignment parity tests, timings at the 1/2/4/8-GPU batch sizes, per-phase cycle counts of a few draws
TAG=${1:-ot}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_nextrows.py -m gpu -x -q -k "assign or ot_ or race" > gpurun_out/${TAG}_test.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_test.log
for n in 965 1930 3860 7720; do python tools/bench_ot.py $n 16 | tail -1; done | tee gpurun_out/${TAG}_ot.txt
python tools/bench_ot.py 3900 8 | tail -1 | tee -a gpurun_out/${TAG}_ot.txt
FG_LIB=finetune-fair-diffusion_b200/csrc/libfairguide_otprof.so python tools/bench_ot.py 7720 16 2>&1 | grep otprof | awk "NR%97==1" | head -12 | tee -a gpurun_out/${TAG}_ot.txt
