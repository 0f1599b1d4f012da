#!/bin/bash
# backward-kernel iteration: parity tests of the image gradient + head, micro-benchmark per tuning build, one ncu capture
TAG=${1:-b}; NCULIB=${2:-finetune-fair-diffusion_b200/csrc/libfairguide.so}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_headline_shape.py tests/test_gpu_parity.py -m gpu -q -k "image_grad or staged or nonfinite or adjoint or head or whole_step or tiled" > gpurun_out/${TAG}_test.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_test.log
for lib in finetune-fair-diffusion_b200/csrc/libfairguide.so $(ls finetune-fair-diffusion_b200/csrc/libfairguide_*.so 2>/dev/null); do
  for dt in bfloat16 float32; do FG_LIB=$lib python tools/bench_kernels.py 1024 $dt 2>&1 | tail -1; done
done | tee gpurun_out/${TAG}_kernels.txt
for dt in bfloat16 float32; do python tools/bench_head.py 965 $dt 2>&1 | tail -1; done | tee -a gpurun_out/${TAG}_kernels.txt
FG_LIB=$NCULIB ncu --set full --clock-control none --import-source on -k regex:image_grad_quad -s 3 -c 1 -f -o gpurun_out/${TAG}_ncu python tools/bench_kernels.py 1024 bfloat16 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
