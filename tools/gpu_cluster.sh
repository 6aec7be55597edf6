#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout -k 5 240 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log | cut -c1-300; }
run t_wgrad python -m pytest tests/test_kernels_gpu.py -q -x -k "wgrad"
DG_WGRAD_CLUSTER=0 run micro_c0 python tools/wgrad_micro.py
DG_WGRAD_CLUSTER=1 run micro_c1 python tools/wgrad_micro.py
