#!/bin/bash
# kernel iteration call: the backward / forward parity tests, then the kernel micro-benchmark for each tuning build
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_headline_shape.py tests/test_gpu_parity.py -m gpu -x -q -k "image_grad or staged or sampler or crop or tiled or whole_step or adjoint or nonfinite" > gpurun_out/${TAG}_test.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_test.log
for lib in "" $(ls finetune-fair-diffusion_b200/csrc/libfairguide_*.so 2>/dev/null); do
  for dt in bfloat16 float32; do
    FG_LIB=${lib:-finetune-fair-diffusion_b200/csrc/libfairguide.so} python tools/bench_kernels.py 1024 $dt 2>&1 | tail -1
  done
done | tee gpurun_out/${TAG}_kernels.txt
FG_BWD_QUAD=1 python tools/bench_kernels.py 1024 bfloat16 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_kernels.txt
