#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout -k 5 400 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; grep '^{' gpurun_out/$name.log | cut -c1-190; }
timeout -k 5 300 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
DG_GEMM_SWAP=1 timeout -k 5 100 python tools/gemm_micro4.py 2>&1 | grep "^swap"
DG_GEMM_SWAP=0 run b512_swap0 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
DG_GEMM_SWAP=1 run b512_swap1 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
DG_GEMM_SWAP=0 run b64_swap0 python bench.py --steps 60 --warmup 9 --also-512 0 --no-roofline --no-cpu-baseline
DG_GEMM_SWAP=1 run b64_swap1 python bench.py --steps 60 --warmup 9 --also-512 0 --no-roofline --no-cpu-baseline
