#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout -k 5 400 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 2 gpurun_out/$name.log | cut -c1-260; }
run t_all python -m pytest tests/ -x -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench_default python bench.py
grep '^{' gpurun_out/bench_default.log > gpurun_out/bench_default.json
