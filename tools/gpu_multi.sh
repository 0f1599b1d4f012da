#!/bin/bash
# multi-GPU call: N = number of visible GPUs.  2 GPUs: also the NCCL pytest.  Then bench.py at N ranks (weak headline + strong_scaling block).
TAG=${1:-m}; N=${2:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/${TAG}_test_multigpu.log 2>&1
  echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_test_multigpu.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
echo "bench rc=$?"; head -c 2500 gpurun_out/${TAG}_bench_${N}gpu.json; tail -5 gpurun_out/${TAG}_bench_${N}gpu.err
