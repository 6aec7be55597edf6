#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout -k 5 400 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; grep '^{' gpurun_out/$name.log | cut -c1-190; }
for rep in 1 2; do
DG_GEMM_BNPOLICY=0 run b512_old_$rep python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
DG_GEMM_BNPOLICY=1 run b512_new_$rep python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
DG_GEMM_BNPOLICY=0 run b64_old_$rep python bench.py --steps 60 --warmup 9 --also-512 0 --no-roofline --no-cpu-baseline
DG_GEMM_BNPOLICY=1 run b64_new_$rep python bench.py --steps 60 --warmup 9 --also-512 0 --no-roofline --no-cpu-baseline
done
DG_GEMM_BNPOLICY=1 DG_GEMM_PAIR=0 run b512_new_nopair python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
