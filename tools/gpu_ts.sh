#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/micro_pn.log
for pair in 0 1; do for n in 1 4; do PAIR=$pair DG_GEMM_NACC=$n timeout -k 5 120 python tools/gemm_micro3.py 2>&1 | grep -v "bn=256" >> gpurun_out/micro_pn.log; done; done
timeout -k 5 200 python -m pytest tests/test_kernels_gpu.py -q -x -k "pairs" 2>&1 | tail -2
