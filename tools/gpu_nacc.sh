#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/micro_nacc.log
for n in 1 2 4; do DG_GEMM_NACC=$n timeout -k 5 120 python tools/gemm_micro3.py >> gpurun_out/micro_nacc.log 2>&1; echo "nacc $n exit $?"; done
DG_GEMM_NACC=4 DG_GEMM_DEBUG=4 timeout -k 5 120 python tools/gemm_micro3.py >> gpurun_out/micro_nacc.log 2>&1
timeout -k 5 200 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv_down or conv_up or pairs or splitk or fused or masked" 2>&1 | tail -3
