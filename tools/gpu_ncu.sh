#!/bin/bash
# usage: tools/gpu_ncu.sh TAG KERNEL_REGEX [LIB] [extra env as VAR=VAL ...] -- one `ncu --set full` capture of a kernel in tools/bench_kernels.py
TAG=$1; RX=$2; LIB=${3:-finetune-fair-diffusion_b200/csrc/libfairguide.so}; shift 3
mkdir -p gpurun_out
env FG_LIB=$LIB "$@" ncu --set full --clock-control none --import-source on -k regex:$RX -s 3 -c 1 -f -o gpurun_out/${TAG} python tools/bench_kernels.py 1024 bfloat16 > gpurun_out/${TAG}.log 2>&1
tail -3 gpurun_out/${TAG}.log
