#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout -k 5 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run t_all python -m pytest tests/ -x -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench_default python bench.py
grep '^{' gpurun_out/bench_default.log > gpurun_out/bench_default.json
run bench_ref python bench.py --impl reference --steps 6 --warmup 2
run bench512 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline
grep '^{' gpurun_out/bench512.log > gpurun_out/bench512.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --also-512 0 --no-roofline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench64_final.csv $CMD > gpurun_out/ncu_final.log 2>&1; echo "ncu exit $?"
