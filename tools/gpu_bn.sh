#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/bn_micro.log
for m in 0 1; do for shape in "2097152 64" "524288 128" "131072 256"; do echo "stream=$m $shape" >> gpurun_out/bn_micro.log; DG_BN_STREAM=$m timeout -k 5 100 python tools/bn_micro.py $shape >> gpurun_out/bn_micro.log 2>&1; done; done
timeout -k 5 200 python -m pytest tests/test_kernels_gpu.py -q -x -k "bn" 2>&1 | tail -3
