#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; t0=$SECONDS; timeout -k 5 400 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$name.log | head -1)"; }
for mk in 12 16 20 24 32; do DG_SPLITK_GFLOP=1 DG_SPLITK_MINK=$mk run b64_sk_1_$mk python bench.py --steps 60 --warmup 9 --also-512 0 --no-roofline --no-cpu-baseline; done
DG_SPLITK_GFLOP=0.2 DG_SPLITK_MINK=16 run b64_sk_02_16 python bench.py --steps 60 --warmup 9 --also-512 0 --no-roofline --no-cpu-baseline
run b64_base python bench.py --steps 60 --warmup 9 --also-512 0 --no-roofline --no-cpu-baseline
DG_SPLITK_GFLOP=1 DG_SPLITK_MINK=16 run b512_sk_1_16 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
run b512_base python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
DG_SPLITK_GFLOP=1 DG_SPLITK_MINK=16 timeout -k 5 300 python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -2
