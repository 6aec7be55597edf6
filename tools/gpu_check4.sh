#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 4 gpurun_out/$name.log | cut -c1-300; }
run t_kernels python -m pytest tests/test_kernels_gpu.py -q -k "fused_bn or pack"
run t_model python -m pytest tests/test_model_gpu.py -q
run bench64 python bench.py --steps 30 --warmup 6 --also-512 0
run bench512 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline
grep '^{' gpurun_out/bench64.log > gpurun_out/bench64.json
grep '^{' gpurun_out/bench512.log > gpurun_out/bench512.json
