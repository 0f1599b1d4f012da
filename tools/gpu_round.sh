#!/bin/bash
# One gpurun call of the round: GPU tests, the default bench line, the other BASELINE configs, the next-row kernels.  usage: tools/gpu_round.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_test_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_test_gpu.log
tail -8 gpurun_out/${TAG}_test_gpu.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; head -c 1500 gpurun_out/${TAG}_bench.json; echo
for W in C1 C2 C3 C4; do
  timeout 300 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
  echo "bench $W rc=$?"; head -c 300 gpurun_out/${TAG}_bench_$W.json; echo
done
timeout 300 python tools/bench_next.py > gpurun_out/${TAG}_bench_next.json 2> gpurun_out/${TAG}_bench_next.err; tail -3 gpurun_out/${TAG}_bench_next.json
