#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 4 gpurun_out/$name.log | cut -c1-300; }
run t_kernels python -m pytest tests/test_kernels_gpu.py -q -k "not simt"
DG_WGRAD_SHARE=1 run t_share1 python -m pytest tests/test_kernels_gpu.py -q -k "conv_wgrad and tc"
DG_WGRAD_SHARE=2 run t_share2 python -m pytest tests/test_kernels_gpu.py -q -k "conv_wgrad and tc"
run bench64 python bench.py --steps 30 --warmup 6 --also-512 0 --no-cpu-baseline
run bench512 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline
DG_WGRAD_SHARE=1 run bench512_share1 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline
DG_WGRAD_SHARE=2 run bench512_share2 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline
for f in bench64 bench512 bench512_share1 bench512_share2; do grep '^{' gpurun_out/$f.log > gpurun_out/$f.json; done
