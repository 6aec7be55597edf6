#!/bin/bash
mkdir -p gpurun_out
DG_GEMM_DEBUG=6 timeout -k 5 100 python tools/gemm_micro4.py 2>&1 | grep "^swap" | sed 's/^/noprefetch /' | tee gpurun_out/micro_mask.log
timeout -k 5 100 python tools/gemm_micro4.py 2>&1 | tee -a gpurun_out/micro_mask.log
timeout -k 5 200 python -m pytest tests/test_kernels_gpu.py -q -x -k "role_swapped or masked" 2>&1 | tail -2
