#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run t_model python -m pytest tests/test_model_gpu.py -q
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench64 python bench.py --steps 30 --warmup 6 --also-512 0
run bench512 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline
DISCOGAN_B200_WGRAD_LANES=1 run bench512_wl python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline --no-roofline
for f in bench64 bench512 bench512_wl; do grep '^{' gpurun_out/$f.log > gpurun_out/$f.json; done
