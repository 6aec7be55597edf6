#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout -k 5 400 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 4 gpurun_out/$name.log | cut -c1-250; }
run micro_gemm2 python tools/gemm_micro2.py
run t_all python -m pytest tests/ -x -q -m gpu
DG_GEMM_PAIR=0 run b512_pair0 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
DG_GEMM_PAIR=1 run b512_pair1 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
DG_GEMM_PAIR=0 run b64_pair0 python bench.py --steps 60 --warmup 9 --also-512 0 --no-roofline --no-cpu-baseline
DG_GEMM_PAIR=1 run b64_pair1 python bench.py --steps 60 --warmup 9 --also-512 0 --no-roofline --no-cpu-baseline
