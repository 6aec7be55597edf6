#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/micro_gemm3.log
for d in 0 1 2 3 4; do DG_GEMM_DEBUG=$d timeout -k 5 120 python tools/gemm_micro3.py >> gpurun_out/micro_gemm3.log 2>&1; echo "debug $d exit $?"; done
timeout -k 5 200 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv_down or conv_up or pairs or splitk or fused" 2>&1 | tail -3
timeout -k 5 300 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline > gpurun_out/b512_bn.log 2>&1; tail -n 1 gpurun_out/b512_bn.log | cut -c1-200
