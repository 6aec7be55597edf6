#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_kernels_gpu.py -q -x -k "role_swapped" 2>&1 | tail -15
DG_GEMM_SWAP=0 timeout -k 5 100 python tools/gemm_micro4.py 2>&1 | grep -v "^swap" | tee gpurun_out/micro_swap.log
DG_GEMM_SWAP=1 timeout -k 5 100 python tools/gemm_micro4.py 2>&1 | grep "^swap" | tee -a gpurun_out/micro_swap.log
