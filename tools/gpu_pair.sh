#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout -k 5 300 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 12 gpurun_out/$name.log | cut -c1-300; }
run t_pairs python -m pytest tests/test_kernels_gpu.py -q -x -k "cta_pairs or splitk"
run micro_gemm python tools/gemm_micro.py
