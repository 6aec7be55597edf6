#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout -k 5 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; grep '^{' gpurun_out/$name.log | cut -c1-200; }
run bench64_r python bench.py --steps 30 --warmup 6 --also-512 0 --no-cpu-baseline
run bench512_r python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline
for f in bench64_r bench512_r; do grep '^{' gpurun_out/$f.log > gpurun_out/$f.json; done
