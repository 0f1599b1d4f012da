#!/bin/bash
# A/B of the tuning builds of libfairguide (csrc/libfairguide_*.so): kernel micro-benchmark, bf16, 3 repetitions each
TAG=${1:-ab}
mkdir -p gpurun_out
for rep in 1 2; do
for lib in finetune-fair-diffusion_b200/csrc/libfairguide.so $(ls finetune-fair-diffusion_b200/csrc/libfairguide_*.so 2>/dev/null); do
  FG_LIB=$lib python tools/bench_kernels.py 1024 bfloat16 2>&1 | tail -1
done; done | tee gpurun_out/${TAG}_kernels.txt
FG_BWD_QUAD=1 python tools/bench_kernels.py 1024 bfloat16 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_kernels.txt
