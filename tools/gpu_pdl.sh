#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout -k 5 400 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 4 gpurun_out/$name.log | cut -c1-400; }
run t_all python -m pytest tests/ -x -q -m gpu
DG_PDL=0 run b64_pdl0 python bench.py --steps 60 --warmup 9 --also-512 0 --no-roofline --no-cpu-baseline
DG_PDL=1 run b64_pdl1 python bench.py --steps 60 --warmup 9 --also-512 0 --no-roofline --no-cpu-baseline
DG_PDL=0 run b512_pdl0 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
DG_PDL=1 run b512_pdl1 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
DG_PDL=1 DG_WGRAD_CLUSTER=0 run b512_pdl1_c0 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
