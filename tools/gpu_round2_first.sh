#!/bin/bash
# First GPU call of round 2: validate the experimental cta_group::2 wgrad kernel (DG_WGRAD_PAIR=1, written at the end of
# round 1 without hardware) against the existing tests, then A/B it.  One box, no profiler.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout -k 5 400 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log | cut -c1-260; }
DG_WGRAD_PAIR=1 run t_wgrad_pair python -m pytest tests/test_kernels_gpu.py -q -x -k "wgrad"
DG_WGRAD_PAIR=0 run micro_wgrad_pair0 python tools/wgrad_micro.py
DG_WGRAD_PAIR=1 run micro_wgrad_pair1 python tools/wgrad_micro.py
DG_WGRAD_PAIR=1 run t_model_pair python -m pytest tests/test_model_gpu.py -q -x
DG_WGRAD_PAIR=0 run b512_pair0 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
DG_WGRAD_PAIR=1 run b512_pair1 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
